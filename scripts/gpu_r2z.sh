# round 2, call Z: 64-column GEMM tiles for sub-wave grids: train step, H=512 tower
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_train.py tests/test_gpu_encoder.py -q -m gpu -x -k "gemm or train or cnn" 2>&1 | tail -3
for bn in 0 128; do
  export SSE_GEMM_BN=$bn; [ $bn = 0 ] && unset SSE_GEMM_BN
  timeout 900 python bench.py --steps 10 --warmup 3 --train-steps 10 --no-real-regime --no-cpu-baseline > gpurun_out/bench_z_train_$bn.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_z_train_$bn.json')); print('BN=$bn train 1024 rows %.1f step/s %.3f ms' % (d['train']['value'], d['train']['ms_per_step']))"
  timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --train-steps 10 --no-real-regime --no-cpu-baseline > gpurun_out/bench_z_c4_$bn.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_z_c4_$bn.json')); print('BN=$bn c4 1536 rows %.1f step/s %.3f ms' % (d['value'], d['ms_per_step']))"
  timeout 900 python bench.py --config c5 --targets 400000 --steps 10 --warmup 3 --train-steps 0 --no-real-regime --no-cpu-baseline > gpurun_out/bench_z_c5_$bn.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_z_c5_$bn.json')); print('BN=$bn c5 enc %.4f ms, step %.4f' % (d['roofline']['encoder']['ms'], d['ms_per_step']))"
done
