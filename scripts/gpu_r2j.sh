# round 2, call J: wide LSTM kernel (64 units per CTA, clusters of H/64): parity, error, timeline, sizes; bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_misc.py tests/test_gpu_headline.py -q -m gpu -x 2>&1 | tail -8
for v in "64 1" "64 0" "64 2" "32 1"; do set -- $v
  echo "#### UNITS=$1 GATE=$2"
  SSE_LSTM_UNITS=$1 SSE_LSTM_GATE_MATH=$2 timeout 300 python tests/probe_lstm_error.py 2>&1 | grep "kernel 3"
  SSE_LSTM_UNITS=$1 SSE_LSTM_GATE_MATH=$2 LSTM_KERNELS=3 LSTM_DBG=1 timeout 300 python scripts/lstm_debug.py 600 2>&1 | grep -v "^\[lstm ptable dbg\] -"
done
echo "#### sizes"
LSTM_KERNELS=1,3 timeout 300 python scripts/lstm_debug.py 128 1200 4800 18944 2>&1 | grep encode
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/bench_n1_j.json 2> gpurun_out/bench_n1_j.err; tail -3 gpurun_out/bench_n1_j.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1_j.json'))
print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['roofline']['encoder'], d['regimes'], d['config']['pipeline'])
PY
