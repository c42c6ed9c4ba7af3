"""Timing experiment on the tcgen05 LSTM towers (not a test, not a benchmark).
usage: lstm_debug.py [B ...]   -- times both kernels (1 = weight streaming, 2 = cluster) per batch size."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sequence-semantic-embedding_b200")); sys.path.insert(0, REPO)
import sse_ffi, bench
CFG = bench.CONFIGS["headline"]
h = bench.make_handle(CFG, 0, sse_ffi)
h.set_params(bench.init_weights(CFG))
sizes = [int(a) for a in sys.argv[1:]] or [600, 148 * 128]
ref = {}
for B in sizes:
    tok = torch.from_numpy(bench.synth_tokens(np.random.default_rng(1), B, CFG)).cuda()
    for kern in [int(x) for x in os.environ.get("LSTM_KERNELS", "1,2,3").split(",")]:
        h.set_option("lstm_kernel", kern)
        out = torch.empty(B, 256, device="cuda")
        for _ in range(2): h.encode(0, tok, B, out, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): h.encode(0, tok, B, out, True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        diff = (out - ref[B]).abs().max().item() if B in ref else 0.0
        ref.setdefault(B, out.clone())
        print("B=%d kernel %d encode %.3f ms -> %.1f TFLOP/s   max|diff vs kernel 1| %.2e" % (B, kern, ms, B * bench.encoder_flops(CFG) / ms / 1e9, diff), file=sys.stderr, flush=True)
        if os.environ.get("LSTM_DBG"):
            os.environ["SSE_LSTM_DEBUG"] = "1"
            h.encode(0, tok, B, out, True); torch.cuda.synchronize()
            os.environ.pop("SSE_LSTM_DEBUG")
