# round 2, call W: pipelined step with the 20-SM encoder (variant 3, 64 units per CTA) and a larger scan grid
mkdir -p gpurun_out
B="timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --train-steps 0 --no-real-regime"
run() { name=$1; shift; $B "$@" > gpurun_out/bench_w_$name.json 2> gpurun_out/bench_w_$name.err; python - $name <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_w_%s.json'%sys.argv[1]))
    e=d['roofline']['encoder']
    print(sys.argv[1], 'q/s %.0f ms/step %.4f e2e %.0f | search %.4f frac %.3f | enc alone %.4f (128-row %.4f)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], e['ms'], e['ms_128_row_clusters']))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
}
run base
export SSE_LSTM_VARIANT=3 SSE_LSTM_UNITS=64
run w_c108_l32
run w_c128_l0 --search-ctas 128 --search-late 0
run w_c128_l16 --search-ctas 128 --search-late 16
run w_c124_l24 --search-ctas 124 --search-late 24
run w_c120_l24 --search-ctas 120 --search-late 24
run w_c132_l0 --search-ctas 132 --search-late 0
