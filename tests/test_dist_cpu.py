"""world_size-2 gloo test (CPU) of the multi-GPU retrieval plumbing: shard ranges, global ids,
packing of (score, id) into one all-gather message, rank-major unpacking and the merge order.
The merge itself is the CUDA kernel on GPUs; here a numpy stand-in with the same (score desc,
id asc) rule checks the host logic against the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sse_dist
import sse_oracle as O


def _merge_numpy(cs, ci, k):
    cs, ci = cs.numpy(), ci.numpy()
    order = np.lexsort((ci, -cs), axis=1)[:, :k]
    return torch.from_numpy(np.take_along_axis(cs, order, 1)), torch.from_numpy(np.take_along_axis(ci, order, 1))


def _worker(rank, world, port, N, Q, E, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _body(rank, world, N, Q, E, k, out)
    except Exception as e:          # surface worker failures instead of timing out
        if rank == 0:
            out.put((False, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def _body(rank, world, N, Q, E, k, out):
    rng = np.random.default_rng(5)
    tgt = rng.standard_normal((N, E)).astype(np.float32)
    tgt[7] = tgt[N - 3]                           # an exact tie across the two shards
    q = rng.standard_normal((Q, E)).astype(np.float32)
    lo, hi = sse_dist.shard_range(N, world, rank)
    d = q @ tgt[lo:hi].T
    s, i = O.top_k_tf(d, k, normalize_scores=False)          # stands in for the local fused top-k
    s_t, i_t = torch.from_numpy(s.astype(np.float32)), torch.from_numpy((i + lo).astype(np.int32))
    ms, mi = sse_dist.gather_and_merge(s_t, i_t, k, _merge_numpy)
    if rank == 0:
        ws, wi = O.top_k_tf(q @ tgt.T, k, normalize_scores=False)
        out.put((np.array_equal(mi.numpy(), wi), float(np.abs(ms.numpy() - ws).max())))


def test_sharded_topk_allgather_merge_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, 9, 16, 5, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok, err = out.get(timeout=90)
    for p in procs:
        p.join(timeout=60)
    assert ok, err
    assert err < 1e-6


def test_shard_range_partitions_exactly():
    for n, g in [(1000000, 8), (7, 3), (5, 8), (0, 2)]:
        spans = [sse_dist.shard_range(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_pack_unpack_round_trip():
    s = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    i = torch.tensor([[1, 2, 3, -1]] * 3, dtype=torch.int32)
    packed = sse_dist.pack_topk(s, i)
    g = torch.stack([packed, packed + 0])
    cs, ci = sse_dist.unpack_gathered(g, 4)
    assert cs.shape == (3, 8) and torch.equal(cs[:, :4], s) and torch.equal(ci[:, 4:], i)


def _body_distributed_queries(rank, world, N, Q, E, k, out):
    """The throughput-serving flow of bench.py --gpus N: every rank brings its own queries."""
    rng = np.random.default_rng(11)
    tgt = rng.standard_normal((N, E)).astype(np.float32)
    q_all = rng.standard_normal((world * Q, E)).astype(np.float32)
    q_mine = torch.from_numpy(q_all[rank * Q:(rank + 1) * Q].copy())
    gathered = sse_dist.allgather_rows(q_mine)
    assert np.array_equal(gathered.numpy(), q_all)                       # rank-major, bit-exact
    pre = torch.empty(world * Q, E)
    assert sse_dist.allgather_rows(q_mine, pre) is pre and np.array_equal(pre.numpy(), q_all)
    lo, hi = sse_dist.shard_range(N, world, rank)
    s, i = O.top_k_tf(gathered.numpy() @ tgt[lo:hi].T, k, normalize_scores=False)
    ms, mi = sse_dist.gather_and_merge(torch.from_numpy(s.astype(np.float32)), torch.from_numpy((i + lo).astype(np.int32)), k, _merge_numpy)
    my_s, my_i = sse_dist.rows_of_rank(ms, rank, world), sse_dist.rows_of_rank(mi, rank, world)
    ws, wi = O.top_k_tf(q_mine.numpy() @ tgt.T, k, normalize_scores=False)
    ok = np.array_equal(my_i.numpy(), wi) and float(np.abs(my_s.numpy() - ws).max()) < 1e-6
    flags = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(flags, torch.tensor([1.0 if ok else 0.0]))
    if rank == 0:
        out.put((all(f.item() == 1.0 for f in flags), 0.0))


def _worker_dq(rank, world, port, N, Q, E, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _body_distributed_queries(rank, world, N, Q, E, k, out)
    except Exception as e:
        if rank == 0:
            out.put((False, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_distributed_queries_allgather_scan_merge_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker_dq, args=(r, 2, port, 777, 6, 8, 4, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok, err = out.get(timeout=90)
    for p in procs:
        p.join(timeout=60)
    assert ok, err


def test_allgather_rows_single_process_is_identity():
    x = torch.arange(6.0).reshape(3, 2)
    assert sse_dist.allgather_rows(x) is x
    y = torch.empty(3, 2)
    assert sse_dist.allgather_rows(x, y) is y and torch.equal(x, y)
    assert torch.equal(sse_dist.rows_of_rank(torch.arange(8).reshape(4, 2), 1, 2), torch.tensor([[4, 5], [6, 7]]))


def _body_own_rows(rank, world, N, Q, E, k, out):
    """bench.py --gpus N flow: packed per-shard top-k of all G*Q rows -> all-to-all of the per-owner row blocks -> merge for the own rows."""
    rng = np.random.default_rng(21)
    tgt = rng.standard_normal((N, E)).astype(np.float32)
    tgt[3] = tgt[N - 2]                                                   # a tie across shards: the lower global id must win
    q_all = rng.standard_normal((world * Q, E)).astype(np.float32)
    lo, hi = sse_dist.shard_range(N, world, rank)
    s, i = O.top_k_tf(q_all @ tgt[lo:hi].T, k, normalize_scores=False)   # this shard's top-k for ALL rows (stands in for sse_search_packed)
    packed = sse_dist.pack_topk(torch.from_numpy(s.astype(np.float32)), torch.from_numpy((i + lo).astype(np.int32)))
    recv = sse_dist.exchange_own_rows(packed)
    ms, mi = sse_dist.merge_packed_numpy(recv, world, k)
    ws, wi = O.top_k_tf(q_all[rank * Q:(rank + 1) * Q] @ tgt.T, k, normalize_scores=False)
    ok = np.array_equal(mi.numpy(), wi) and float(np.abs(ms.numpy() - ws).max()) < 1e-6
    flags = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(flags, torch.tensor([1.0 if ok else 0.0]))
    if rank == 0:
        out.put((all(f.item() == 1.0 for f in flags), 0.0))


def _worker_own(rank, world, port, N, Q, E, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _body_own_rows(rank, world, N, Q, E, k, out)
    except Exception as e:
        if rank == 0:
            out.put((False, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_own_rows_all_to_all_then_merge_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker_own, args=(r, 2, port, 901, 7, 8, 4, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok, err = out.get(timeout=90)
    for p in procs:
        p.join(timeout=60)
    assert ok, err


def test_exchange_own_rows_single_process_is_identity():
    x = torch.arange(12.0).reshape(3, 4)
    assert sse_dist.exchange_own_rows(x) is x
    s, i = sse_dist.merge_packed_numpy(sse_dist.pack_topk(torch.tensor([[0.5, 0.1]]), torch.tensor([[4, 9]], dtype=torch.int32)), 1, 2)
    assert s.tolist() == [[0.5, 0.10000000149011612]] and i.tolist() == [[4, 9]]
