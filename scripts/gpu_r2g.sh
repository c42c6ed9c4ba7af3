# round 2, call G: in-kernel cycle counters (scan, LSTM) + ncu --set full of the dominant kernels of the r02 build
mkdir -p gpurun_out
SCAN_Q=600 timeout 300 python scripts/scan_debug.py 1000000 2> gpurun_out/scan_dbg_600.log; tail -60 gpurun_out/scan_dbg_600.log
SCAN_Q=4800 timeout 300 python scripts/scan_debug.py 125000 2> gpurun_out/scan_dbg_4800.log; tail -40 gpurun_out/scan_dbg_4800.log
LSTM_DBG=1 timeout 300 python scripts/lstm_debug.py 600 2> gpurun_out/lstm_dbg_600.log; tail -40 gpurun_out/lstm_dbg_600.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 7 -c 1 -o gpurun_out/prof_scan_600 python scripts/search_probe.py 600x1000000 > gpurun_out/ncu_scan_600.log 2>&1; tail -1 gpurun_out/ncu_scan_600.log | cut -c1-150
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ptable_kernel -s 2 -c 1 -o gpurun_out/prof_lstm python bench.py --steps 2 --warmup 3 --repeats 1 --no-pipeline --no-cpu-baseline --train-steps 0 --no-real-regime --index gaussian > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log | cut -c1-200
timeout 600 ncu --set full --clock-control none -k regex:"finalize_kernel|select_tau|tok_prep|prep_queries" -s 8 -c 4 -o gpurun_out/prof_small python bench.py --steps 2 --warmup 3 --repeats 1 --no-pipeline --no-cpu-baseline --train-steps 0 --no-real-regime --index gaussian > gpurun_out/ncu_full3.log 2>&1; tail -1 gpurun_out/ncu_full3.log | cut -c1-200
ls -la gpurun_out | head -50
