set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tcmb scripts/tc_microbench.cu 2>/dev/null; timeout 100 /tmp/tcmb > gpurun_out/tc_microbench.log 2>&1; tail -4 gpurun_out/tc_microbench.log
timeout 600 python -m pytest tests/test_gpu_search.py -x -q 2>&1 | tail -8
timeout 300 python scripts/scan_debug.py 2>&1 | grep -v "prod_total\|mma_wait_a\|epi_total" | head -40
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2700 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
