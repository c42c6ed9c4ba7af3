# LSTM kernel variant sweep (experiments): prints TF/s for B=600 and a full wave
for cfg in "1 1 1 4" "1 1 0 4" "2 0 1 2" "2 0 0 2" "1 1 1 2" "1 1 0 2" "2 0 1 1"; do
  set -- $cfg
  echo "== xbufs=$1 bias_smem=$2 gate_math=$3 slot_kb=$4"
  SSE_LSTM_XBUFS=$1 SSE_LSTM_BIAS_SMEM=$2 SSE_LSTM_GATE_MATH=$3 SSE_LSTM_SLOT_KB=$4 timeout 120 python scripts/lstm_debug.py 2>&1 | grep "encode\|wait_acce\|wait_wfull\|wait_xfull\|wait_hfull" | awk '{printf "%s | ", $0} END{print ""}' | cut -c1-900
done
