# round 2, call N: one rank's share of an 8-GPU step (emulated) under different pipelining choices; N=1 cap sweep
mkdir -p gpurun_out
B="timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --train-steps 0 --no-real-regime"
run() { name=$1; shift; $B "$@" > gpurun_out/bench_n_$name.json 2> gpurun_out/bench_n_$name.err; python - $name <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_n_%s.json'%sys.argv[1]))
e=d['roofline']['encoder']
print(sys.argv[1], 'q/s %.0f ms/step %.4f e2e %.0f | search %.4f frac %.3f | enc alone %.4f (128-row %.4f)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], e['ms'], e['ms_128_row_clusters']))
PY
}
run n1_c100 --search-ctas 100
run n1_c108
run n1_c116 --search-ctas 116
run n1_c108_late --search-late 24
run e8_c108 --emulate-world 8 --search-ctas 108 --cluster-rows 128
run e8_c0_r64 --emulate-world 8 --search-ctas 0 --cluster-rows 64
run e8_c0_r128 --emulate-world 8 --search-ctas 0 --cluster-rows 128
run e8_c124 --emulate-world 8 --search-ctas 124 --cluster-rows 128
run e8_c68_r64 --emulate-world 8 --search-ctas 68 --cluster-rows 64
run e4_c108 --emulate-world 4 --search-ctas 108 --cluster-rows 128
run e2_c108 --emulate-world 2 --search-ctas 108 --cluster-rows 128
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_emu8.csv python bench.py --steps 2 --warmup 1 --repeats 1 --emulate-world 8 --no-pipeline --no-cpu-baseline --train-steps 0 --no-real-regime > gpurun_out/ncu_emu8.log 2>&1; python scripts/launch_table.py gpurun_out/launches_emu8.csv 2>/dev/null | head -16
