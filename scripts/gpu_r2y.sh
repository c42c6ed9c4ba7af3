# round 2, call Y: full GPU suite + default bench + configs 3 / 4 after the GEMM / gather changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json')); e=d['roofline']['encoder']
print('headline q/s %.0f ms/step %.4f e2e %.0f | search %.4f frac %.3f | enc %.4f (128-row %.4f) | real %.0f | train %.1f | cpu %.1f | verify %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], e['ms'], e['ms_128_row_clusters'], d['regimes']['real']['value'], d['train']['value'], d['cpu_baseline']['value'], d['verify']))
PY
show() { python - $1 <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
r=d['roofline']
if 'encoder' in r: print(sys.argv[1], 'q/s %.0f ms/step %.4f | search frac %.3f | enc %.4f ms %.1f TF/s | index_build_s %.3f' % (d['value'], d['ms_per_step'], r['frac'], r['encoder']['ms'], r['encoder']['achieved_tflops'], d['config'].get('index_build_s', 0)))
else: print(sys.argv[1], d['metric'], '%.1f' % d['value'], 'ms %.3f' % d['ms_per_step'], r)
PY
}
timeout 900 python bench.py --config c3 --steps 20 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; show gpurun_out/bench_c3.json
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --train-steps 10 --no-real-regime > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; show gpurun_out/bench_c4.json
timeout 900 python bench.py --steps 10 --warmup 3 --train-steps 10 --no-real-regime --no-cpu-baseline > gpurun_out/bench_n1_train.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_n1_train.json')); print('train 1024 rows', d['train'])"
