# round 2, call A: new cluster-multicast scan -- parity tests, then the configuration sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -x -q 2>&1 | tail -5
timeout 1200 python scripts/scan_configs.py 600x1000000 4800x125000 2400x250000 300x1000000 2>&1 | tee gpurun_out/scan_configs.log
SCAN_Q=600 SSE_SCAN_MTG=1 timeout 300 python scripts/scan_debug.py 1000000 2> gpurun_out/scan_debug_mtg1.log; tail -30 gpurun_out/scan_debug_mtg1.log
SCAN_Q=600 SSE_SCAN_MTG=2 timeout 300 python scripts/scan_debug.py 1000000 2> gpurun_out/scan_debug_mtg2.log; tail -30 gpurun_out/scan_debug_mtg2.log
