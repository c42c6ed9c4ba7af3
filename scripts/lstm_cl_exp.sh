timeout 300 python -m pytest tests/test_gpu_encoder.py -x -q 2>&1 | tail -3
LSTM_DBG=1 timeout 120 python scripts/lstm_debug.py 600 2>&1 | grep "kernel 3\|ptable dbg"
timeout 120 python scripts/lstm_debug.py 64 600 2048 4800 18944 2>&1 | grep "kernel 3\|B=18944 kernel 1"
