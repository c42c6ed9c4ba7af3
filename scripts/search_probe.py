"""Probe (not a test, not a benchmark): search time and parity vs torch over (Q, N) shapes."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sequence-semantic-embedding_b200"))
import sse_ffi
E, k = 256, 10
shapes = [(600, 1000000), (1200, 500000), (2400, 250000), (4800, 125000), (4800, 1000000)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for Q, N in shapes:
    h = sse_ffi.Handle("dual-encoder", 50, 8, E, 8, 8, 8, precision=sse_ffi.PRECISION_TC)
    g = torch.Generator(device="cuda").manual_seed(7)
    idx = torch.randn(N, E, device="cuda", generator=g); idx /= idx.norm(dim=1, keepdim=True)
    h.index_set(idx, N, 0)
    q = torch.randn(Q, E, device="cuda", generator=g); q /= q.norm(dim=1, keepdim=True)
    s = torch.empty(Q, k, device="cuda"); i = torch.empty(Q, k, device="cuda", dtype=torch.int32)
    for _ in range(3): h.search(q, Q, k, s, i)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10): h.search(q, Q, k, s, i)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    # parity on a slice of the queries
    qs = slice(0, min(Q, 256))
    ref = (q[qs].double() @ idx.double().T).topk(k, dim=1)
    same = (ref.indices.int() == i[qs]).float().mean().item()
    print("Q=%5d N=%8d  search %.3f ms  (%.0f TF/s)  idx-parity %.4f" % (Q, N, ms, 2.0 * Q * N * E / ms / 1e9, same), flush=True)
    del h, idx
