set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], 'search ms',d['roofline']['ms_per_launch'], 'frac',d['roofline']['frac'], 'enc ms',d['roofline']['encoder']['ms'], d['train']['value'])"; tail -5 gpurun_out/bench_n1.err
timeout 900 ncu --set full --clock-control none -k regex:scan_kernel -s 4 -c 2 -o gpurun_out/prof_scan python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-100
