# round 2, call B: full GPU test suite on the new token pre-pass / scan defaults, the scan sweep, first bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 900 python scripts/scan_configs.py 600x1000000 4800x125000 2400x250000 300x1000000 200x1000000 2>&1 | tee gpurun_out/scan_configs_b.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --steps 20 --warmup 3 --emulate-world 8 --no-cpu-baseline --train-steps 0 > gpurun_out/bench_emu8.json 2> gpurun_out/bench_emu8.err; tail -c 1500 gpurun_out/bench_emu8.json; tail -5 gpurun_out/bench_emu8.err
