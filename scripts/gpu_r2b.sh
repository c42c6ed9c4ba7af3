# round 2, call B: full GPU test suite, the scan sweep, first bench lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 900 python scripts/scan_configs.py 600x1000000 4800x125000 2400x250000 300x1000000 2>&1 | tee gpurun_out/scan_configs_b.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --steps 20 --warmup 3 --search-late 40 --no-cpu-baseline --train-steps 0 --no-real-regime > gpurun_out/bench_n1_late40.json 2> gpurun_out/bench_n1_late40.err; tail -c 1200 gpurun_out/bench_n1_late40.json; tail -3 gpurun_out/bench_n1_late40.err
timeout 600 python bench.py --steps 20 --warmup 3 --search-ctas 124 --search-late 24 --no-cpu-baseline --train-steps 0 --no-real-regime > gpurun_out/bench_n1_c124.json 2> gpurun_out/bench_n1_c124.err; tail -c 600 gpurun_out/bench_n1_c124.json | head -c 400
timeout 600 python bench.py --steps 20 --warmup 3 --emulate-world 8 --no-cpu-baseline --train-steps 0 > gpurun_out/bench_emu8.json 2> gpurun_out/bench_emu8.err; tail -c 1500 gpurun_out/bench_emu8.json; tail -5 gpurun_out/bench_emu8.err
timeout 600 python bench.py --steps 20 --warmup 3 --emulate-world 8 --search-ctas 108 --search-late 40 --no-cpu-baseline --train-steps 0 --no-real-regime > gpurun_out/bench_emu8_late.json 2> gpurun_out/bench_emu8_late.err; tail -c 600 gpurun_out/bench_emu8_late.json | head -c 400
