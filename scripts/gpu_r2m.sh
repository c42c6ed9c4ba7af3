# round 2, call M: 64-row clusters for query batches; pipelining choices
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_misc.py tests/test_gpu_headline.py -q -m gpu -x 2>&1 | tail -5
LSTM_KERNELS=3 LSTM_DBG=1 timeout 300 python scripts/lstm_debug.py 600 1200 2>&1 | grep -v "^\[lstm ptable dbg\] -" | grep -v "step 2[23]"
B="timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --train-steps 0"
run() { name=$1; shift; $B "$@" > gpurun_out/bench_m_$name.json 2> gpurun_out/bench_m_$name.err; python - $name <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_m_%s.json'%sys.argv[1]))
e=d['roofline']['encoder']
print(sys.argv[1], 'q/s %.0f ms/step %.4f e2e %.0f | search %.4f frac %.3f | enc alone %.4f (128-row %.4f) | real %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], e['ms'], e['ms_128_row_clusters'], d['regimes']['real']['value'] if d.get('regimes') else 0))
PY
}
run default
run c68r64 --search-ctas 68 --cluster-rows 64
run c0r64 --search-ctas 0 --cluster-rows 64
run c0r128 --search-ctas 0 --cluster-rows 128
run c116r128 --search-ctas 116 --cluster-rows 128
run nopipe --no-pipeline
