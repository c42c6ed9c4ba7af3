set -x
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], 'search ms',d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'])" $1; }
timeout 600 python -m pytest tests/test_gpu_search.py -x -q 2>&1 | tail -3
for pk in 0 1; do
echo "== pack $pk"
SSE_SCAN_PACK=$pk timeout 300 python scripts/search_probe.py 600x1000000 257x300000 1200x500000 2400x250000 4800x125000 2>&1 | grep "Q="
SSE_SCAN_PACK=$pk timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/bench_pack$pk.json 2> gpurun_out/bench_pack$pk.err; show gpurun_out/bench_pack$pk.json
done
