# round 2, call O: finalize buffers sized by slots; sample fraction at the 8-GPU per-rank shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_headline.py -q -m gpu -x 2>&1 | tail -3
SCAN_CONFIGS=auto,div8,div12,div24 timeout 900 python scripts/scan_configs.py 4800x125000 2400x250000 600x1000000 2>&1 | tee gpurun_out/scan_div_sweep.log
B="timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --train-steps 0 --no-real-regime"
run() { name=$1; shift; $B "$@" > gpurun_out/bench_o_$name.json 2> gpurun_out/bench_o_$name.err; python - $name <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_o_%s.json'%sys.argv[1]))
e=d['roofline']['encoder']
print(sys.argv[1], 'q/s %.0f ms/step %.4f e2e %.0f | search %.4f frac %.3f | enc alone %.4f (128-row %.4f)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], e['ms'], e['ms_128_row_clusters']))
PY
}
run e8_c0_r64 --emulate-world 8 --search-ctas 0 --cluster-rows 64
SSE_SCAN_SAMPLE_DIV=8 run e8_c0_r64_div8 --emulate-world 8 --search-ctas 0 --cluster-rows 64
run e4_c0_r64 --emulate-world 4 --search-ctas 0 --cluster-rows 64
run e4_c108 --emulate-world 4 --search-ctas 108 --cluster-rows 128
run n1_late24 --search-late 24
run n1_late32 --search-late 32
run n1_late16 --search-late 16
