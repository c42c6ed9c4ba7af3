# round 2, call AE: two row-half chains per tower in the train step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_entrypoints.py -q -m gpu 2>&1 | tail -3
SSE_TRAIN_GRAPH=0 timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu 2>&1 | tail -2
for ch in 0 1; do
  export SSE_TRAIN_CHAINS=$ch; [ $ch = 0 ] && unset SSE_TRAIN_CHAINS
  timeout 900 python bench.py --steps 10 --warmup 3 --train-steps 30 --no-real-regime --no-cpu-baseline > gpurun_out/bench_ae_train_$ch.json 2> gpurun_out/bench_ae_train_$ch.err; tail -2 gpurun_out/bench_ae_train_$ch.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ae_train_$ch.json')); print('single_chain=$ch train 1024 rows %.1f step/s %.3f ms' % (d['train']['value'], d['train']['ms_per_step']))"
  timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --train-steps 30 --no-real-regime --no-cpu-baseline > gpurun_out/bench_ae_c4_$ch.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_ae_c4_$ch.json')); print('single_chain=$ch c4 1536 rows %.1f step/s %.3f ms' % (d['value'], d['ms_per_step']))"
done
