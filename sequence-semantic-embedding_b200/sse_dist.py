# coding=utf-8
"""Multi-GPU retrieval plumbing: row-sharded target index + ONE all-gather of per-shard top-k.

New design (the reference has no distributed code): rank r owns the contiguous
global rows [r*N/G, (r+1)*N/G) in its own HBM; every rank holds the same query batch; each rank
runs the fused local top-k (global ids), the packed [Q, 2k] (fp32 score | int32 id bits) tensors
are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) and merged per row with the
(score desc, id asc) order.  The merge on GPUs is the CUDA merge kernel (sse_merge_topk); the
`merge_fn` hook exists so the packing / id arithmetic can be tested with world_size-2 gloo on a
CPU-only box (tests/test_dist_cpu.py)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of `rank`; the first n_total % world ranks get one extra row."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_topk(scores: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """[Q,k] fp32 + [Q,k] int32 -> [Q,2k] fp32 (ids bit-cast), one message per rank."""
    return torch.cat([scores.contiguous(), ids.contiguous().view(torch.float32)], dim=1)


def unpack_gathered(gathered: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[G,Q,2k] -> candidate scores [Q,G*k], candidate ids [Q,G*k] (rank-major per row)."""
    G, Q, _ = gathered.shape
    cs = gathered[:, :, :k].permute(1, 0, 2).reshape(Q, G * k).contiguous()
    ci = gathered[:, :, k:].permute(1, 0, 2).reshape(Q, G * k).contiguous().view(torch.int32)
    return cs, ci


def gather_and_merge(scores: torch.Tensor, ids: torch.Tensor, k: int, merge_fn: Callable, group=None):
    """all-gather the packed local top-k and merge; returns (scores [Q,k], ids [Q,k]) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return scores, ids
    packed = pack_topk(scores, ids)
    flat = torch.empty((world * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(flat, packed, group=group)            # rank-major concatenation along dim 0
    cs, ci = unpack_gathered(flat.view(world, packed.shape[0], packed.shape[1]), k)
    return merge_fn(cs, ci, k)


def exchange_own_rows(packed: torch.Tensor, out: Optional[torch.Tensor] = None, group=None) -> torch.Tensor:
    """The exchange step when every rank only needs the results of ITS OWN queries.  `packed` [G*R, 2k] holds this rank's
    per-shard top-k of ALL G*R query rows, rank-major (rows [r*R, (r+1)*R) are rank r's queries).  One all-to-all sends
    block r to rank r; the result [G*R, 2k] holds, for this rank's R queries, the packed top-k found in shard 0..G-1
    (block g = from rank g) -- the layout sse_merge_packed / merge_packed_numpy consume.  Per rank G*R*2k*4 bytes in and
    out and a merge over R rows, instead of an all-gather of G such blocks and a merge over G*R rows."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        if out is not None:
            out.copy_(packed)
            return out
        return packed
    if out is None:
        out = torch.empty_like(packed)
    dist.all_to_all_single(out, packed.contiguous(), group=group)
    return out


def merge_packed_numpy(recv: torch.Tensor, G: int, k: int):
    """CPU stand-in of sse_merge_packed for the gloo tests: recv [G, R, 2k] -> (scores [R,k], ids [R,k]), (score desc, id asc)."""
    r3 = recv.view(G, -1, 2 * k)
    cs, ci = unpack_gathered(r3, k)
    cs, ci = cs.numpy(), ci.numpy()
    import numpy as _np
    key_s = _np.where(ci < 0, -_np.inf, cs)
    order = _np.lexsort((ci, -key_s), axis=1)[:, :k]
    return torch.from_numpy(_np.take_along_axis(cs, order, 1)), torch.from_numpy(_np.take_along_axis(ci, order, 1))


def allgather_rows(x: torch.Tensor, out: Optional[torch.Tensor] = None, group=None) -> torch.Tensor:
    """Rank-major concatenation of equally shaped per-rank row blocks: [R, C] on every rank -> [G*R, C] on every rank
    (rank r's rows land at [r*R, (r+1)*R)).  The throughput-serving flow uses it to hand every rank the query encodings
    of all ranks before the shard scan (each rank encodes only its own batch); single process: returns x."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        if out is not None:
            out.copy_(x)
            return out
        return x
    if out is None:
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def rows_of_rank(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """The block of `rank` inside a rank-major concatenation produced by allgather_rows."""
    r = x.shape[0] // world
    return x[rank * r:(rank + 1) * r]


class ShardedIndex(object):
    """Index rows partitioned over the ranks of the default process group."""

    def __init__(self, handle, n_total: int):
        self.h = handle
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.n_total = n_total
        self.lo, self.hi = shard_range(n_total, self.world, self.rank)

    def set_local(self, enc_local):
        """enc_local: fp32 [hi-lo, E] rows of this rank (numpy or device tensor)."""
        assert enc_local.shape[0] == self.hi - self.lo
        self.h.index_set(enc_local, self.hi - self.lo, global_offset=self.lo)

    def build_local(self, tgt_tokens_local, batch: int = 10000):
        assert tgt_tokens_local.shape[0] == self.hi - self.lo
        self.h.index_build(tgt_tokens_local, global_offset=self.lo, batch=batch)

    def _merge_cuda(self, cs, ci, k):
        Q = cs.shape[0]
        out_s = torch.empty(Q, k, device=cs.device)
        out_i = torch.empty(Q, k, device=cs.device, dtype=torch.int32)
        self.h.merge_topk(cs, ci, Q, cs.shape[1], k, out_s, out_i, torch.cuda.current_stream())
        return out_s, out_i

    def search(self, q_dev: torch.Tensor, k: int):
        Q = q_dev.shape[0]
        s = torch.empty(Q, k, device=q_dev.device)
        i = torch.empty(Q, k, device=q_dev.device, dtype=torch.int32)
        self.h.search(q_dev, Q, k, s, i, torch.cuda.current_stream())
        return gather_and_merge(s, i, k, self._merge_cuda)

    def search_distributed_queries(self, q_local_dev: torch.Tensor, k: int):
        """Every rank brings its OWN [Q_r, E] query vectors (same Q_r on all ranks): all-gather them, scan the local
        shard for all G*Q_r straight into one packed [G*Q_r, 2k] block (sse_search_packed), all-to-all so that every
        rank receives the G per-shard blocks of its own rows, merge G*k -> k for those rows (sse_merge_packed).
        Returns (scores [Q_r,k], ids [Q_r,k]) of this rank's queries."""
        Qr = q_local_dev.shape[0]
        q_all = allgather_rows(q_local_dev)
        Q = q_all.shape[0]
        packed = torch.empty(Q, 2 * k, device=q_all.device)
        self.h.search_packed(q_all, Q, k, packed, torch.cuda.current_stream())
        recv = exchange_own_rows(packed)
        s = torch.empty(Qr, k, device=q_all.device)
        i = torch.empty(Qr, k, device=q_all.device, dtype=torch.int32)
        self.h.merge_packed(recv, self.world, Qr, k, s, i, torch.cuda.current_stream())
        return s, i


def allreduce_train_step(handle, src, tgt, labels, b_global: int):
    """Data-parallel train step: local grads scaled by 1/B_global, all-reduce(sum) of the arena, apply."""
    handle.train_grads(src, tgt, labels, b_global, stream=torch.cuda.current_stream())
    ptr, n = handle.grad_arena()

    class _Raw:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    # wrap the arena on the HANDLE's device: with device="cuda" (torch's current device) a handle living on another GPU
    # would be silently copied, the all-reduce would run on the copy and the apply would use un-reduced gradients
    arena = torch.as_tensor(_Raw(), device=torch.device("cuda", int(handle.cfg.device)))
    if arena.data_ptr() != ptr:
        raise RuntimeError("gradient arena was copied instead of wrapped (device mismatch?)")
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(arena)
    return handle.train_apply(stream=torch.cuda.current_stream())
