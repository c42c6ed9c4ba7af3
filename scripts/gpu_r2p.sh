# round 2, call P: evidence run on one GPU -- all GPU tests, default bench, reference arm, launch list, ncu --set full of the dominant kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; head -c 400 gpurun_out/bench_n1.json; echo; tail -3 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 900 gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 3 --no-pipeline --no-cpu-baseline --train-steps 0 > gpurun_out/bench_n1_nopipe.json 2> gpurun_out/bench_n1_nopipe.err; head -c 300 gpurun_out/bench_n1_nopipe.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_n1.csv python bench.py --steps 2 --warmup 1 --repeats 1 --no-pipeline --no-cpu-baseline --train-steps 0 --no-real-regime > gpurun_out/ncu_n1.log 2>&1; python scripts/launch_table.py gpurun_out/launches_n1.csv 2>/dev/null | head -16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 7 -c 1 -o gpurun_out/prof_scan_600 -f python scripts/search_probe.py 600x1000000 > gpurun_out/ncu_scan_600.log 2>&1; tail -1 gpurun_out/ncu_scan_600.log | cut -c1-150
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ptable_kernel -s 2 -c 1 -o gpurun_out/prof_lstm -f python bench.py --steps 2 --warmup 3 --repeats 1 --no-pipeline --no-cpu-baseline --train-steps 0 --no-real-regime --index gaussian > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log | cut -c1-200
ls -la gpurun_out | head -60
