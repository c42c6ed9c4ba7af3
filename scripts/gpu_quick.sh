set -x
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], 'search ms',d['roofline']['ms_per_launch'], 'clk', d['clocks'])" $1; }
for g in 2 4; do for c in 108 0; do
timeout 900 python bench.py --steps 20 --warmup 3 --emulate-world $g --search-ctas $c --no-cpu-baseline --train-steps 0 > gpurun_out/bench_emu${g}_c$c.json 2> gpurun_out/bench_emu${g}_c$c.err; show gpurun_out/bench_emu${g}_c$c.json; tail -2 gpurun_out/bench_emu${g}_c$c.err
done; done
