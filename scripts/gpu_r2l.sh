# round 2, call L: what slows the h exchange when the table rows are staged (producer experiments)
for f in 0 1 2 4 6; do
  echo "#### FLAGS=$f"
  SSE_LSTM_FLAGS=$f LSTM_KERNELS=3 LSTM_DBG=1 timeout 300 python scripts/lstm_debug.py 600 2>&1 | grep -v "^\[lstm ptable dbg\] -" | grep -v "step 2[23]"
done
