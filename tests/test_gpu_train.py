"""GPU parity of the training step (C ABI sse_train_step / sse_train_grads + sse_train_apply)
against the oracle's restatement of sess.run([train, loss, train_acc]) (sse_model.py:279-302,
355-364).  fp32 both sides; tolerances are for summation-order noise only."""
import numpy as np
import pytest

import sse_ffi
import sse_oracle as O

pytestmark = pytest.mark.gpu


def batch(rng, B, T, V):
    """data.Data.get_train_batch layout (data.py:95-115): rows alternate (pos, label 1), (neg, label 0),
    the source row repeated for the pair."""
    half = B // 2
    src_half = O.synth_tokens(rng, half, T, V, "real", 4.0)
    src = np.repeat(src_half, 2, axis=0)
    tgt = O.synth_tokens(rng, B, T, V, "real", 8.0)
    labels = np.tile(np.array([1.0, 0.0], np.float32), half)
    return src, tgt, labels


@pytest.mark.parametrize("mode,V,We,E,H,T,B", [
    ("dual-encoder", 400, 50, 64, 96, 12, 10),      # classification recipe dims (makefile:5), short T
    ("shared-encoder", 300, 40, 50, 96, 9, 8),      # crosslingual recipe dims (makefile:42)
    ("dual-encoder", 900, 64, 64, 64, 20, 128),
])
def test_train_steps_match_oracle(mode, V, We, E, H, T, B):
    p = O.init_params(mode, V, We, E, H, H, seed=11)
    h = sse_ffi.Handle(mode, V, We, E, H, H, T, learning_rate=0.9, precision=sse_ffi.PRECISION_FP32)
    h.set_params(p)
    st = O.TrainState({k: v.copy() for k, v in p.items()}, learning_rate=0.9)
    rng = np.random.default_rng(5)
    for step in range(3):
        src, tgt, labels = batch(rng, B, T, V)
        loss, acc, gn = h.train_step(src, tgt, labels)
        wl, wa, wg = O.train_step(st, mode, src, tgt, labels, dtype=np.float32)
        assert abs(loss - wl) < 2e-4 * max(1.0, abs(wl)), (step, loss, wl)
        assert abs(acc - wa) < 1e-6, (step, acc, wa)
        assert abs(gn - wg) < 2e-3 * wg, (step, gn, wg)
        for name, want in st.params.items():
            got = h.get_param(name)
            assert np.abs(got - want).max() < 2e-4, (step, name, np.abs(got - want).max())
            ga = h.get_param(name + "/Adagrad")
            assert np.abs(ga - st.accum[name]).max() < 2e-3 * max(1.0, np.abs(st.accum[name]).max()), (step, name)
    lr, gs = h.scalars()
    assert gs == 3 and abs(lr - 0.9) < 1e-7
    h.lr_decay()
    assert abs(h.scalars()[0] - np.float32(0.9) * np.float32(0.99)) < 1e-7
    h.close()


def test_split_grads_plus_apply_equals_train_step_and_dp_sum():
    """Two half-batches through sse_train_grads (scaled by 1/B_global) summed == one full batch:
    the data-parallel contract behind the NCCL all-reduce of the gradient arena."""
    import torch
    mode, V, We, E, H, T, B = "dual-encoder", 500, 32, 32, 48, 10, 32
    p = O.init_params(mode, V, We, E, H, H, seed=3)
    rng = np.random.default_rng(8)
    src, tgt, labels = batch(rng, B, T, V)
    full = sse_ffi.Handle(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_FP32); full.set_params(p)
    a = sse_ffi.Handle(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_FP32); a.set_params(p)
    b = sse_ffi.Handle(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_FP32); b.set_params(p)
    lf, af, gf = full.train_step(src, tgt, labels)
    a.train_grads(src[: B // 2], tgt[: B // 2], labels[: B // 2], B)
    b.train_grads(src[B // 2:], tgt[B // 2:], labels[B // 2:], B)
    pa, na = a.grad_arena(); pb, nb = b.grad_arena()
    assert na == nb

    class _Raw:                      # view raw device memory as torch tensors
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    ta, tb = torch.as_tensor(_Raw(pa, na), device="cuda"), torch.as_tensor(_Raw(pb, nb), device="cuda")
    ta += tb                                              # what the all-reduce(sum) produces on every rank
    torch.cuda.synchronize()
    l2, a2, g2 = a.train_apply()
    assert abs(l2 - lf) < 1e-5 and abs(a2 - af) < 1e-6 and abs(g2 - gf) < 1e-3 * gf
    for name in p:
        assert np.abs(a.get_param(name) - full.get_param(name)).max() < 1e-5, name
    for hh in (full, a, b):
        hh.close()


def test_pair_score_matches_oracle():
    import torch
    mode, V, We, E, H, T, B = "shared-encoder", 300, 24, 20, 32, 8, 19
    p = O.init_params(mode, V, We, E, H, H, seed=2)
    h = sse_ffi.Handle(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_FP32); h.set_params(p)
    rng = np.random.default_rng(1)
    src, tgt = O.synth_tokens(rng, B, T, V, "real", 3.0), O.synth_tokens(rng, B, T, V, "real", 5.0)
    cos = torch.empty(B, device="cuda")
    h.pair_score(torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda(), B, cos)
    want = O.binarylogit(O.encode(p, mode, "src", src), O.encode(p, mode, "tgt", tgt))
    assert np.abs(cos.cpu().numpy() - want).max() < 1e-5
    h.close()
