"""Timing experiment on the tcgen05 LSTM tower (not a test, not a benchmark)."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sequence-semantic-embedding_b200")); sys.path.insert(0, REPO)
import sse_ffi, bench
h = sse_ffi.Handle("dual-encoder", bench.V, 256, 256, 256, 256, 50, precision=sse_ffi.PRECISION_TC)
h.set_params(bench.init_weights())
for B in (600, 148 * 128):
    tok = torch.from_numpy(bench.synth_tokens(np.random.default_rng(1), B)).cuda()
    out = torch.empty(B, 256, device="cuda")
    for _ in range(2): h.encode(0, tok, B, out, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.encode(0, tok, B, out, True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("B=%d encode %.3f ms -> %.1f TFLOP/s" % (B, ms, B * bench.F_LSTM / ms / 1e9), file=sys.stderr, flush=True)
    os.environ["SSE_LSTM_DEBUG"] = "1"
    h.encode(0, tok, B, out, True); torch.cuda.synchronize()
    os.environ.pop("SSE_LSTM_DEBUG")
