set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python scripts/lstm_debug.py 2>&1 | grep "encode"
for r in 64 128; do SSE_LSTM_ROWS=$r timeout 300 python scripts/lstm_debug.py 2>&1 | grep "encode" | head -1; done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], 'search ms',d['roofline']['ms_per_launch'], 'frac',d['roofline']['frac'], 'enc ms',d['roofline']['encoder']['ms'], d['train']['value'])"; tail -5 gpurun_out/bench_n1.err
