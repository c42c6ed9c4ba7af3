# one GPU round: tests, bench, launch list, full ncu capture of the dominant kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 2 -o gpurun_out/prof_scan python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_kernel -s 2 -c 1 -o gpurun_out/prof_lstm_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log | cut -c1-300
ls -la gpurun_out
