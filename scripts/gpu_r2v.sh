# round 2, call V: smoke(); ncu --set full at the 8-GPU per-rank shape (4800 x 125k): filter scan, sample scan, finalize
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 6 -c 2 -o gpurun_out/prof_scan_4800 -f python scripts/search_probe.py 4800x125000 > gpurun_out/ncu_scan_4800.log 2>&1; tail -1 gpurun_out/ncu_scan_4800.log | cut -c1-150
timeout 600 ncu --set full --clock-control none -k regex:"finalize_kernel|select_tau" -s 6 -c 2 -o gpurun_out/prof_fin_4800 -f python scripts/search_probe.py 4800x125000 > gpurun_out/ncu_fin_4800.log 2>&1; tail -1 gpurun_out/ncu_fin_4800.log | cut -c1-150
SCAN_Q=4800 timeout 300 python scripts/scan_debug.py 125000 2>&1 | grep -E "scan dbg|search call|==" | head -14
