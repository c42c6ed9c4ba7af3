# round 2, call H: LSTM table-kernel variants (epilogue warps x gate math x poll back-off): parity, error, timeline; scan per-group view
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_misc.py tests/test_gpu_headline.py -q -m gpu -x 2>&1 | tail -5
for v in "8 0" "16 0" "16 1" "16 2" "8 1" "8 2"; do set -- $v
  echo "#### EW=$1 GATE=$2"
  SSE_LSTM_EW=$1 SSE_LSTM_GATE_MATH=$2 timeout 300 python tests/probe_lstm_error.py 2>&1 | grep "kernel 3"
  SSE_LSTM_EW=$1 SSE_LSTM_GATE_MATH=$2 LSTM_KERNELS=3 LSTM_DBG=1 timeout 300 python scripts/lstm_debug.py 600 2>&1 | grep -v "^\[lstm ptable dbg\] -"
done
echo "#### EW=16 GATE=1 POLL=0"
SSE_LSTM_POLL=0 LSTM_KERNELS=3 LSTM_DBG=1 timeout 300 python scripts/lstm_debug.py 600 4800 2>&1 | grep -v "^\[lstm ptable dbg\] -"
echo "#### gate 2 tests"
SSE_LSTM_GATE_MATH=2 timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_misc.py -q -m gpu 2>&1 | tail -5
SCAN_Q=600 timeout 300 python scripts/scan_debug.py 1000000 2>&1 | grep -E "group|search call|mma_total|==" 
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/bench_n1_h.json 2> gpurun_out/bench_n1_h.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1_h.json'))
print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['roofline']['encoder'], d['regimes'])
PY
