"""2-GPU tests (NCCL): row-sharded retrieval == single-GPU retrieval; data-parallel train step ==
single-GPU train step.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import sse_dist
    import sse_ffi
    import sse_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        E, N, Q, k = 128, 40000, 300, 10
        rng = np.random.default_rng(3)
        tgt = rng.standard_normal((N, E)).astype(np.float32); tgt /= np.linalg.norm(tgt, axis=1, keepdims=True)
        q = rng.standard_normal((Q, E)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
        h = sse_ffi.Handle("dual-encoder", 50, 8, E, 8, 8, 8, device=rank)
        sh = sse_dist.ShardedIndex(h, N)
        sh.set_local(tgt[sh.lo:sh.hi])
        s, i = sh.search(torch.from_numpy(q).cuda(), k)
        torch.cuda.synchronize()
        res = {"idx": i.cpu().numpy(), "sc": s.cpu().numpy()}
        # throughput flow: every rank brings its own 150 queries; packed scan + all-to-all of the per-owner blocks + merge
        Qr = Q // world
        s2, i2 = sh.search_distributed_queries(torch.from_numpy(q[rank * Qr:(rank + 1) * Qr]).cuda(), k)
        torch.cuda.synchronize()
        own_ok = bool(np.array_equal(i2.cpu().numpy(), res["idx"][rank * Qr:(rank + 1) * Qr]) and
                      np.abs(s2.cpu().numpy() - res["sc"][rank * Qr:(rank + 1) * Qr]).max() < 1e-6)
        flags = [torch.zeros(1, device="cuda") for _ in range(world)]
        dist.all_gather(flags, torch.tensor([1.0 if own_ok else 0.0], device="cuda"))
        res["ok_own_rows"] = all(f.item() == 1.0 for f in flags)
        if rank == 0:
            h.index_set(tgt, global_offset=0)
            s1 = torch.empty(Q, k, device="cuda"); i1 = torch.empty(Q, k, device="cuda", dtype=torch.int32)
            h.search(torch.from_numpy(q).cuda(), Q, k, s1, i1)
            torch.cuda.synchronize()
            res["ok_search"] = bool(np.array_equal(res["idx"], i1.cpu().numpy()) and np.abs(res["sc"] - s1.cpu().numpy()).max() < 1e-6)
        h.close()
        # ---- data-parallel train step
        mode, V, We, Ee, H, T, B = "dual-encoder", 500, 32, 32, 64, 10, 64
        p = O.init_params(mode, V, We, Ee, H, H, seed=3)
        r2 = np.random.default_rng(8)
        src = O.synth_tokens(r2, B, T, V, "real", 4.0); tg = O.synth_tokens(r2, B, T, V, "real", 6.0)
        lab = np.tile(np.array([1.0, 0.0], np.float32), B // 2)
        hd = sse_ffi.Handle(mode, V, We, Ee, H, H, T, device=rank, precision=sse_ffi.PRECISION_FP32); hd.set_params(p)
        lo, hi = rank * B // world, (rank + 1) * B // world
        l, a, g = sse_dist.allreduce_train_step(hd, torch.from_numpy(src[lo:hi]).cuda(), torch.from_numpy(tg[lo:hi]).cuda(),
                                                torch.from_numpy(lab[lo:hi]).cuda(), B)
        if rank == 0:
            hs = sse_ffi.Handle(mode, V, We, Ee, H, H, T, device=rank, precision=sse_ffi.PRECISION_FP32); hs.set_params(p)
            l1, a1, g1 = hs.train_step(src, tg, lab)
            ok = abs(l - l1) < 1e-5 and abs(a - a1) < 1e-6 and abs(g - g1) < 1e-3 * g1
            for name in p:
                ok = ok and np.abs(hd.get_param(name) - hs.get_param(name)).max() < 1e-5
            res["ok_train"] = bool(ok)
            out.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_search_and_dp_train():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert res["ok_search"], "sharded search differs from single-GPU search"
    assert res["ok_own_rows"], "own-rows exchange (all-to-all + merge_packed) differs from the all-gather path"
    assert res["ok_train"], "data-parallel train step differs from single-GPU step"
