# round 2, call X: GEMM ring depth for multi-wave grids (two CTAs per SM): CNN tower, train step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_train.py tests/test_gpu_encoder.py -q -m gpu -x -k "gemm or train or cnn" 2>&1 | tail -3
show() { python - $1 <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
r=d['roofline']
if 'encoder' in r: print(sys.argv[1], 'q/s %.0f ms/step %.4f | enc %.4f ms %.1f TF/s | index_build_s %.3f' % (d['value'], d['ms_per_step'], r['encoder']['ms'], r['encoder']['achieved_tflops'], d['config'].get('index_build_s', 0)))
else: print(sys.argv[1], d['metric'], '%.1f' % d['value'], 'ms %.3f' % d['ms_per_step'])
PY
}
for st in 0 6; do
  export SSE_GEMM_STAGES=$st; [ $st = 0 ] && unset SSE_GEMM_STAGES
  timeout 900 python bench.py --config c3 --steps 20 --warmup 3 --no-cpu-baseline --train-steps 0 --no-real-regime > gpurun_out/bench_x_c3_$st.json 2> gpurun_out/bench_x_c3_$st.err; show gpurun_out/bench_x_c3_$st.json
  timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --train-steps 10 --no-real-regime --no-cpu-baseline > gpurun_out/bench_x_c4_$st.json 2> gpurun_out/bench_x_c4_$st.err; show gpurun_out/bench_x_c4_$st.json
done
