# round 2, call AB: tensor-core train step on the library's streams, optionally as a CUDA graph
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_entrypoints.py tests/test_gpu_misc.py -q -m gpu 2>&1 | tail -3
SSE_TRAIN_GRAPH=1 timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu 2>&1 | tail -3
for gr in 0 1; do
  export SSE_TRAIN_GRAPH=$gr
  timeout 900 python bench.py --steps 10 --warmup 3 --train-steps 30 --no-real-regime --no-cpu-baseline > gpurun_out/bench_ab_train_$gr.json 2> gpurun_out/bench_ab_train_$gr.err; tail -2 gpurun_out/bench_ab_train_$gr.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ab_train_$gr.json')); print('graph=$gr train 1024 rows %.1f step/s %.3f ms' % (d['train']['value'], d['train']['ms_per_step']))"
  timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --train-steps 30 --no-real-regime --no-cpu-baseline > gpurun_out/bench_ab_c4_$gr.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_ab_c4_$gr.json')); print('graph=$gr c4 1536 rows %.1f step/s %.3f ms' % (d['value'], d['ms_per_step']))"
done
