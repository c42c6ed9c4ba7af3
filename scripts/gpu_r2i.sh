# round 2, call I: scan cost-model sweep for the remainder group
mkdir -p gpurun_out
SCAN_CONFIGS=cost100,cost500,cost780,cost1000,cost1300,cost780pdl timeout 900 python scripts/scan_configs.py 600x1000000 300x1000000 1200x500000 2>&1 | tee gpurun_out/scan_cost_sweep.log
SCAN_Q=600 timeout 300 python scripts/scan_debug.py 1000000 2>&1 | grep -E "group|search call|mma_total|==" | head -8
