"""Probe (not collected by pytest -- no test_ prefix; lives under tests/ because it uses the oracle): error of the
tcgen05 LSTM kernels against the numpy oracle at the headline shape.  python tests/probe_lstm_error.py"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sequence-semantic-embedding_b200")); sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle"))
import sse_ffi, sse_oracle as O
mode, V, We, H, E, T, B = "dual-encoder", 5000, 256, 256, 256, 50, 300
p = O.init_params(mode, V, We, E, H, H, seed=1)
rng = np.random.default_rng(3)
tok = np.concatenate([O.synth_tokens(rng, B // 2, T, V, "full"), O.synth_tokens(rng, B // 2, T, V, "real", 4.0)])
want = O.encode(p, mode, "src", tok, True)
for kern in (1, 2, 3):
    h = sse_ffi.Handle(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_TC)
    h.set_params(p); h.set_option("encoder", 2); h.set_option("lstm_kernel", kern)
    got = h.encode_host(0, tok, True)
    err = np.abs(got - want)
    print("kernel %d gate_math=%s: max err %.2e  mean err %.2e" % (kern, os.environ.get("SSE_LSTM_GATE_MATH", "0"), err.max(), err.mean()), flush=True)
    h.close()
