# round 2, call AI: why the search-alone timing differs between the full default run and the quick runs
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py --steps 20 --warmup 3 "$@" > gpurun_out/bench_ai_$name.json 2> gpurun_out/bench_ai_$name.err; python - $name <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_ai_%s.json'%sys.argv[1]))
print(sys.argv[1], 'q/s %.0f ms/step %.4f | search %.4f frac %.3f | enc %.4f (%.4f)' % (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['roofline']['encoder']['ms'], d['roofline']['encoder']['ms_128_row_clusters']))
PY
}
run quick --no-cpu-baseline --train-steps 0 --no-real-regime
run real --no-cpu-baseline --train-steps 0
run real_train --no-cpu-baseline
run full
run quick2 --no-cpu-baseline --train-steps 0 --no-real-regime
