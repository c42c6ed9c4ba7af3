"""Timing experiments on the tcgen05 scan (not a test, not a benchmark): per-role wait cycles."""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sequence-semantic-embedding_b200"))
import sse_ffi
E, N, k = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 10
h = sse_ffi.Handle("dual-encoder", 50, 8, E, 8, 8, 8, precision=sse_ffi.PRECISION_TC)
g = torch.Generator(device="cuda").manual_seed(7)
idx = torch.randn(N, E, device="cuda", generator=g); idx /= idx.norm(dim=1, keepdim=True)
h.index_set(idx, N, 0)
for Q in [int(x) for x in os.environ.get("SCAN_Q", "600,128,256").split(",")]:
    q = torch.randn(Q, E, device="cuda", generator=g); q /= q.norm(dim=1, keepdim=True)
    s = torch.empty(Q, k, device="cuda"); i = torch.empty(Q, k, device="cuda", dtype=torch.int32)
    for flags in ("0", "2"):
        os.environ.pop("SSE_SCAN_DEBUG", None)
        for _ in range(2): h.search(q, Q, k, s, i)
        torch.cuda.synchronize()
        os.environ["SSE_SCAN_DEBUG"] = "1"; os.environ["SSE_SCAN_FLAGS"] = flags
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        print("==== Q=%d N=%d flags=%s" % (Q, N, flags), file=sys.stderr, flush=True)
        t0.record(); h.search(q, Q, k, s, i); t1.record(); torch.cuda.synchronize()
        print("   search call %.3f ms (includes the debug sync)" % t0.elapsed_time(t1), file=sys.stderr, flush=True)
