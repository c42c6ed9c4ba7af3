# round 2, call C: full GPU tests; scan sweep after the ACC1 / finalize / packing fixes; kernel launch tables; bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -25
timeout 900 python scripts/scan_configs.py 600x1000000 4800x125000 2400x250000 300x1000000 2>&1 | tee gpurun_out/scan_configs_c.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c.csv python bench.py --steps 2 --warmup 1 --repeats 1 --no-pipeline --no-cpu-baseline --train-steps 0 --no-real-regime > gpurun_out/ncu_c.log 2>&1; python scripts/launch_table.py gpurun_out/launches_c.csv | head -30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1400 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 1 --warmup 1 --repeats 1 --no-pipeline --no-cpu-baseline --train-steps 1 --no-real-regime > gpurun_out/ncu_train.log 2>&1; python scripts/launch_table.py gpurun_out/launches_train.csv | head -30
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --steps 20 --warmup 3 --search-late 40 --no-cpu-baseline --train-steps 0 --no-real-regime > gpurun_out/bench_n1_late40.json 2> gpurun_out/bench_n1_late40.err; tail -c 800 gpurun_out/bench_n1_late40.json | head -c 400
timeout 600 python bench.py --steps 20 --warmup 3 --no-pipeline --no-cpu-baseline --train-steps 0 --no-real-regime > gpurun_out/bench_n1_nopipe.json 2> gpurun_out/bench_n1_nopipe.err; head -c 400 gpurun_out/bench_n1_nopipe.json
timeout 600 python bench.py --steps 20 --warmup 3 --emulate-world 8 --no-cpu-baseline --train-steps 0 > gpurun_out/bench_emu8.json 2> gpurun_out/bench_emu8.err; head -c 500 gpurun_out/bench_emu8.json; tail -5 gpurun_out/bench_emu8.err
