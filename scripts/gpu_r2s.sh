# round 2, call S: remainder group follows the heavy groups' tile order (one HBM sweep of the index)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_headline.py tests/test_gpu_entrypoints.py -q -m gpu -x 2>&1 | tail -3
SCAN_CONFIGS=auto,nofollow,follow_cost650,follow_cost500 timeout 900 python scripts/scan_configs.py 600x1000000 300x1000000 2>&1 | tee gpurun_out/scan_follow.log
SCAN_Q=600 timeout 300 python scripts/scan_debug.py 1000000 2>&1 | grep -E "group|search call|mma_total|mma_wait|==" | head -9
timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:scan_kernel -s 7 -c 2 python scripts/search_probe.py 600x1000000 2>&1 | grep -E "dram__bytes_read|gpu__time|hit_rate|scan_kernel" | head -12
B="timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --train-steps 0"
run() { name=$1; shift; $B "$@" > gpurun_out/bench_s_$name.json 2> gpurun_out/bench_s_$name.err; python - $name <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_s_%s.json'%sys.argv[1]))
e=d['roofline']['encoder']
print(sys.argv[1], 'q/s %.0f ms/step %.4f e2e %.0f | search %.4f frac %.3f | enc alone %.4f (128-row %.4f) | real %.0f | verify %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], e['ms'], e['ms_128_row_clusters'], d['regimes']['real']['value'] if d.get('regimes') else 0, d['verify']['index_agreement']))
PY
}
run follow
SSE_SCAN_FOLLOW=0 run nofollow
