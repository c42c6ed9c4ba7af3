# round 2, call AC: full GPU suite + default bench + config 4 with the two-stream / graph train step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json')); e=d['roofline']['encoder']
print('headline q/s %.0f ms/step %.4f e2e %.0f | search %.4f frac %.3f | enc %.4f (128-row %.4f) | real %.0f | train %.1f | cpu %.1f | launches %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], e['ms'], e['ms_128_row_clusters'], d['regimes']['real']['value'], d['train']['value'], d['cpu_baseline']['value'], d['gpu_launches']))
PY
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --train-steps 30 --no-real-regime > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4 %.1f step/s %.3f ms' % (d['value'], d['ms_per_step']), d['roofline'])"
