set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 300 python scripts/scan_debug.py 2>&1 | grep -v "prod_total\|mma_wait_a\|epi_total" | head -60
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2300 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
