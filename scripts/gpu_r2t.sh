# round 2, call T: following order with 32-bit tile arithmetic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_headline.py -q -m gpu -x 2>&1 | tail -3
SCAN_CONFIGS=auto,nofollow,follow_cost650,follow_cost500 timeout 900 python scripts/scan_configs.py 600x1000000 2>&1 | tee gpurun_out/scan_follow.log
SCAN_Q=600 timeout 300 python scripts/scan_debug.py 1000000 2>&1 | grep -E "group|search call|==" | head -5
timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:scan_kernel -s 7 -c 1 python scripts/search_probe.py 600x1000000 2>&1 | grep -E "dram__bytes_read|gpu__time|hit_rate" | head -4
