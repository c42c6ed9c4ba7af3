"""GPU tests of the tcgen05 GEMM (csrc/gemm_tc.cu) and its two users: the fused CNN tower (conv + bias + ReLU + max-pool in
the GEMM epilogue; reference sse_model.py:179-214) and the tensor-core train step (bf16 operands; reference
sse_model.py:279-302,355-364).  GEMM: against float64 products of the SAME 16-bit-rounded operands (so only the fp32
accumulation differs); CNN: north_star encoder tolerance 1e-3 against the fp32 oracle; train step: loss within 1e-2
relative and gradient cosine >= 0.999 against the fp32 parity path ("Tolerance budget")."""
import numpy as np
import pytest

import sse_ffi
import sse_oracle as O

pytestmark = pytest.mark.gpu


def _round16(x, fmt):
    import torch
    t = torch.from_numpy(x)
    return (t.to(torch.bfloat16) if fmt == 1 else t.to(torch.float16)).to(torch.float64).numpy()


@pytest.mark.parametrize("M,N,K,fmt,split", [(1536, 1024, 256, 1, 1), (300, 70, 100, 0, 1), (129, 257, 64, 1, 1), (256, 1024, 9600, 1, 8),
                                             (64, 64, 1000, 0, 3), (1, 8, 8, 1, 1)])
def test_gemm_tc_matches_float64_product_of_rounded_operands(M, N, K, fmt, split):
    import torch
    h = sse_ffi.Handle("dual-encoder", 50, 8, 8, 8, 8, 8)
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((N, K)).astype(np.float32)
    d0 = rng.standard_normal((M, N)).astype(np.float32)
    da, db, dd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(d0).cuda()
    alpha, beta = (1.0, 1.0) if split > 1 else (0.5, 2.0)
    h.debug_gemm_tc(da, db, M, N, K, dd, fmt=fmt, split_k=split, alpha=alpha, beta=beta)
    torch.cuda.synchronize()
    want = alpha * (_round16(a, fmt) @ _round16(b, fmt).T) + beta * d0.astype(np.float64)
    err = np.abs(dd.cpu().numpy() - want).max()
    assert err < 2e-5 * np.sqrt(K) * 4 + 1e-5, err           # fp32 accumulation of K products of O(1) values
    h.close()


@pytest.mark.parametrize("V,We,E,T,B,ks,fs", [(2000, 64, 48, 30, 41, (3, 4, 5), (64, 32, 48)), (4000, 256, 256, 50, 300, (3, 4, 5), (256, 256, 256)),
                                              (1500, 32, 16, 100, 9, (2, 5), (8, 72))])
def test_cnn_tower_on_tensor_cores_within_tolerance(V, We, E, T, B, ks, fs):
    mode = "dual-cnn"
    p = O.init_params(mode, V, We, E, 0, 0, seed=21, cnn_filter_sizes=ks, cnn_num_filters=fs)
    h = sse_ffi.Handle(mode, V, We, E, 0, 0, T, precision=sse_ffi.PRECISION_TC, cnn_filter_sizes=ks, cnn_num_filters=fs)
    h.set_params(p)
    h.set_option("encoder", 2)                               # the tcgen05 tower or an error, never a silent SIMT run
    rng = np.random.default_rng(8)
    tok = np.concatenate([O.synth_tokens(rng, B - B // 2, T, V, "real", 9.0), O.synth_tokens(rng, B // 2, T, V, "full")])
    for side, name in ((0, "src"), (1, "tgt")):
        got = h.encode_host(side, tok, True)
        want = O.encode(p, mode, name, tok, True)
        err = np.abs(got - want).max()
        assert 0 < err < 1e-3, (name, err)
        raw = h.encode_host(side, tok, False)
        wraw = O.encode(p, mode, name, tok, False)
        assert np.abs(raw - wraw).max() < 5e-3 * np.abs(wraw).max()
    h.set_option("encoder", 1)                               # exact fp32 tower on the same handle
    assert np.abs(h.encode_host(0, tok, True) - O.encode(p, mode, "src", tok, True)).max() < 2e-5
    h.close()


def _arena(h):
    import torch
    ptr, n = h.grad_arena()

    class _Raw:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Raw(), device="cuda").clone().cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("mode,V,We,E,H,T,B", [("dual-encoder", 2000, 64, 64, 128, 20, 96), ("shared-encoder", 1500, 128, 96, 64, 12, 64),
                                               ("dual-encoder", 8000, 256, 256, 256, 50, 384)])
def test_tensor_core_train_step_against_the_fp32_parity_path(mode, V, We, E, H, T, B):
    p = O.init_params(mode, V, We, E, H, H, seed=13)
    rng = np.random.default_rng(2)
    src = np.repeat(O.synth_tokens(rng, B // 2, T, V, "real", 4.0), 2, axis=0)
    tgt = O.synth_tokens(rng, B, T, V, "real", 8.0)
    lab = np.tile(np.array([1.0, 0.0], np.float32), B // 2)
    res = {}
    for name, opt in (("fp32", 1), ("tc", 2)):
        h = sse_ffi.Handle(mode, V, We, E, H, H, T, learning_rate=0.5, precision=sse_ffi.PRECISION_TC)
        h.set_params(p)
        h.set_option("train", opt)
        loss, acc = h.train_grads(src, tgt, lab, B)
        res[name] = (loss, acc, _arena(h))
        if name == "tc":
            l2, a2, gn = h.train_apply()                     # the optimizer half is shared code: it must run on these gradients
            assert np.isfinite(gn) and gn > 0
            assert np.abs(h.get_param("word_embedding") - p["word_embedding"]).max() > 0
        h.close()
    (l0, a0, g0), (l1, a1, g1) = res["fp32"], res["tc"]
    assert abs(l1 - l0) < 1e-2 * max(abs(l0), 1e-3), (l0, l1)
    assert abs(a1 - a0) < 0.02
    n_dense = len(g0) - V * We - V - 8
    for lo, hi, what in ((0, n_dense, "dense variables"), (n_dense, n_dense + V * We, "word_embedding")):
        a, b = g0[lo:hi], g1[lo:hi]
        cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
        assert cos > 0.999, (what, cos)
        assert abs(np.linalg.norm(b) / np.linalg.norm(a) - 1) < 2e-2, what
    assert np.array_equal(g0[n_dense + V * We:n_dense + V * We + V], g1[n_dense + V * We:n_dense + V * We + V])      # touched-row counts
