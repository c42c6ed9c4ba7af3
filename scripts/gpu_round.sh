# one GPU round: tests, bench, launch list, full ncu capture of the dominant kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --steps 10 --warmup 3 --targets 100000 --no-cpu-baseline > gpurun_out/bench_n1_100k.json 2>&1; tail -c 1500 gpurun_out/bench_n1_100k.json
timeout 600 python bench.py --steps 5 --warmup 3 --search 1 --no-cpu-baseline > gpurun_out/bench_n1_simt.json 2>&1; tail -c 1500 gpurun_out/bench_n1_simt.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 2 -o gpurun_out/prof_scan python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_step_kernel -s 60 -c 2 -o gpurun_out/prof_lstm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; tail -3 gpurun_out/ncu_full2.log
ls -la gpurun_out
