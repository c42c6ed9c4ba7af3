# round 2, call AF: late scan CTAs -- count x share sweep at the headline shape
mkdir -p gpurun_out
B="timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --train-steps 0 --no-real-regime"
run() { name=$1; shift; $B "$@" > gpurun_out/bench_af_$name.json 2> gpurun_out/bench_af_$name.err; python - $name <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_af_%s.json'%sys.argv[1]))
print(sys.argv[1], 'q/s %.0f ms/step %.4f e2e %.0f | search %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch']))
PY
}
run l32_s40
run l32_s55 --search-late-share 55
run l32_s70 --search-late-share 70
run l40_s40 --search-late 40
run l40_s55 --search-late 40 --search-late-share 55
run l40_s70 --search-late 40 --search-late-share 70
run c104_l36_s55 --search-ctas 104 --search-late 36 --search-late-share 55
