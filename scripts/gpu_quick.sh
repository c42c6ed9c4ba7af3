set -x
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], 'search ms',d['roofline']['ms_per_launch'], 'frac',d['roofline']['frac'], 'enc ms',d['roofline']['encoder']['ms'], d['train'] and d['train']['value'])" $1; }
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 --no-pipeline --no-cpu-baseline --train-steps 0 > gpurun_out/bench_n1_nopipe.json 2> gpurun_out/bench_n1_nopipe.err; show gpurun_out/bench_n1_nopipe.json
for c in 148 126 108; do
timeout 900 python bench.py --steps 20 --warmup 3 --search-ctas $c --no-cpu-baseline --train-steps 0 > gpurun_out/bench_n1_c$c.json 2> gpurun_out/bench_n1_c$c.err; show gpurun_out/bench_n1_c$c.json
done
timeout 900 python bench.py --steps 20 --warmup 3 --queries 4800 --targets 125000 --no-pipeline --no-cpu-baseline --train-steps 0 > gpurun_out/bench_q4800_np.json 2> gpurun_out/bench_q4800_np.err; show gpurun_out/bench_q4800_np.json
timeout 900 python bench.py --steps 20 --warmup 3 --queries 4800 --targets 125000 --no-cpu-baseline --train-steps 0 > gpurun_out/bench_q4800.json 2> gpurun_out/bench_q4800.err; show gpurun_out/bench_q4800.json
