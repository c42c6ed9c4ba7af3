# round 2, call AA: the two towers of the tensor-core train step on two streams
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_gemm_tc.py tests/test_gpu_entrypoints.py tests/test_gpu_misc.py -q -m gpu 2>&1 | tail -3
for sm in 0 1; do
  export SSE_TRAIN_STREAMS=$sm; [ $sm = 0 ] && unset SSE_TRAIN_STREAMS
  timeout 900 python bench.py --steps 10 --warmup 3 --train-steps 20 --no-real-regime --no-cpu-baseline > gpurun_out/bench_aa_train_$sm.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_aa_train_$sm.json')); print('one_stream=$sm train 1024 rows %.1f step/s %.3f ms' % (d['train']['value'], d['train']['ms_per_step']))"
  timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --train-steps 20 --no-real-regime --no-cpu-baseline > gpurun_out/bench_aa_c4_$sm.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_aa_c4_$sm.json')); print('one_stream=$sm c4 1536 rows %.1f step/s %.3f ms' % (d['value'], d['ms_per_step']))"
done
