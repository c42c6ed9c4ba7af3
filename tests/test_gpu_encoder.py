"""GPU parity: the CUDA encoders (called through the C ABI) against the CPU oracle on the
same seeded inputs.  Tolerances: encoder outputs within 1e-3 of the fp32 reference
(north_star); the fp32 SIMT path is expected at ~1e-6 and is held to 2e-5."""
import numpy as np
import pytest

import sse_ffi
import sse_oracle as O

pytestmark = pytest.mark.gpu

TOL_FP32 = 2e-5


def make(mode, V, We, E, Hs, Ht, T, seed=1234, precision=sse_ffi.PRECISION_FP32, **kw):
    p = O.init_params(mode, V, We, E, Hs, Ht, seed=seed, **{k: v for k, v in kw.items() if k.startswith("cnn") or k == "target_space_size"})
    rng = np.random.default_rng(seed + 1)
    for k in p:     # non-zero biases so the bias path is exercised
        if k.endswith("/bias"):
            p[k] = (rng.normal(size=p[k].shape) * 0.1).astype(np.float32)
    h = sse_ffi.Handle(mode, V, We, E, Hs, Ht, T, precision=precision,
                       target_space_size=kw.get("target_space_size", 0),
                       cnn_filter_sizes=kw.get("cnn_filter_sizes", ()), cnn_num_filters=kw.get("cnn_num_filters", ()))
    h.set_params(p)
    return h, p


# the real reference recipes (makefile:5,17,30,42) and the BASELINE shape
SHAPES = [
    ("dual-encoder", 3000, 50, 64, 96, 96, 80, 37),      # classification
    ("shared-encoder", 2500, 40, 50, 96, 96, 50, 70),    # crosslingual
    ("dual-encoder", 3000, 30, 64, 96, 96, 60, 130),     # ranking
    ("dual-encoder", 4000, 256, 256, 256, 256, 50, 150), # BASE
    ("dual-encoder", 500, 7, 5, 3, 9, 6, 1),             # odd everything, B = 1, Hs != Ht
]


@pytest.mark.parametrize("mode,V,We,E,Hs,Ht,T,B", SHAPES)
def test_lstm_encode_matches_oracle(mode, V, We, E, Hs, Ht, T, B):
    h, p = make(mode, V, We, E, Hs, Ht, T)
    rng = np.random.default_rng(42)
    for side, name in ((sse_ffi.SIDE_SRC, "src"), (sse_ffi.SIDE_TGT, "tgt")):
        tok = np.concatenate([O.synth_tokens(rng, B - B // 2, T, V, "full"), O.synth_tokens(rng, B // 2, T, V, "real", 4.0)])
        for normalize in (True, False):
            got = h.encode_host(side, tok, normalize)
            want = O.encode(p, mode, name, tok, normalize)
            scale = 1.0 if normalize else max(1.0, np.abs(want).max())
            assert np.abs(got - want).max() / scale < TOL_FP32, (side, normalize)
    h.close()


def test_edge_rows_all_pad_and_truncated():
    mode, V, We, E, H, T = "dual-encoder", 300, 16, 16, 32, 10
    h, p = make(mode, V, We, E, H, H, T)
    rows = [O.pad_tokens([], T), O.pad_tokens(list(range(2, 40)), T), O.pad_tokens([7], T)]
    tok = np.array(rows, np.int32)
    got = h.encode_host(sse_ffi.SIDE_SRC, tok, True)
    assert np.abs(got - O.encode(p, mode, "src", tok, True)).max() < TOL_FP32
    h.close()


def test_pad_prefix_skip_is_bit_identical():
    mode, V, We, E, H, T = "dual-encoder", 2000, 64, 64, 128, 50
    h, p = make(mode, V, We, E, H, H, T)
    rng = np.random.default_rng(5)
    tok = O.synth_tokens(rng, 90, T, V, "real", 3.0)          # queries: ~3 real tokens of 50
    h.set_option("pad_skip", 0)
    full = h.encode_host(sse_ffi.SIDE_SRC, tok, True)
    h.set_option("pad_skip", 1)
    skipped = h.encode_host(sse_ffi.SIDE_SRC, tok, True)
    assert np.array_equal(full, skipped)
    # weights change -> table must be invalidated
    p2 = {k: (v * 1.01).astype(np.float32) for k, v in p.items()}
    h.set_params(p2)
    s2 = h.encode_host(sse_ffi.SIDE_SRC, tok, True)
    assert np.abs(s2 - O.encode(p2, mode, "src", tok, True)).max() < TOL_FP32
    h.close()


def test_device_pointer_entry_point_matches_host_entry_point():
    import torch
    mode, V, We, E, H, T, B = "shared-encoder", 1000, 32, 48, 64, 20, 33
    h, p = make(mode, V, We, E, H, H, T)
    tok = O.synth_tokens(np.random.default_rng(2), B, T, V, "real", 6.0)
    dtok = torch.from_numpy(tok).cuda()
    out = torch.empty(B, E, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        h.encode(sse_ffi.SIDE_TGT, dtok, B, out, True, stream=st)
    st.synchronize()
    assert np.array_equal(out.cpu().numpy(), h.encode_host(sse_ffi.SIDE_TGT, tok, True))
    h.close()


@pytest.mark.parametrize("mode", ["dual-cnn", "source_only_cnn"])
def test_cnn_encode_matches_oracle(mode):
    V, We, E, T, B = 2000, 64, 48, 30, 41
    kw = dict(cnn_filter_sizes=(3, 4, 5), cnn_num_filters=(64, 32, 48)) if mode == "dual-cnn" else dict(target_space_size=11)
    h, p = make(mode, V, We, E, 0, 0, T, **kw)
    rng = np.random.default_rng(8)
    tok = O.synth_tokens(rng, B, T, V, "real", 9.0)
    sides = [("src", 0)] + ([("tgt", 1)] if mode == "dual-cnn" else [])
    for name, side in sides:
        got = h.encode_host(side, tok, True)
        assert np.abs(got - O.encode(p, mode, name, tok, True)).max() < TOL_FP32
    if mode == "source_only_cnn":
        with pytest.raises(sse_ffi.SseError):
            h.encode_host(1, tok, True)
    h.close()


def test_param_round_trip_and_errors():
    h, p = make("dual-encoder", 100, 8, 8, 16, 16, 12)
    names = dict(h.param_names())
    for k, v in p.items():
        assert names[k] == v.shape
        assert np.array_equal(h.get_param(k), v)
        assert np.allclose(h.get_param(k + "/Adagrad"), 0.1)
    with pytest.raises(sse_ffi.SseError):
        h.set_param("nope", np.zeros(3, np.float32))
    with pytest.raises(sse_ffi.SseError):
        h.set_param("word_embedding", np.zeros((3, 3), np.float32))
    with pytest.raises(sse_ffi.SseError):
        h.encode_host(0, np.zeros((2, 5), np.int32))
    h.close()


# ---------------------------------------------------------------------------------------------
# tensor-core (tcgen05, fp16 operands / fp32 accumulate + state) LSTM tower: north_star tolerance
TOL_TC = 1e-3


@pytest.mark.parametrize("We,H,E,T,B", [(256, 256, 256, 50, 300), (64, 64, 32, 12, 128), (128, 192, 64, 20, 77),
                                        (256, 128, 256, 50, 1), (192, 256, 128, 30, 513),
                                        # the real reference recipes (makefile:5,42,30): run on zero-padded tiles (We -> 64, H -> 128)
                                        (50, 96, 64, 80, 130), (40, 96, 50, 50, 70), (30, 96, 64, 60, 37)])
@pytest.mark.parametrize("kern", [1, 2, 3])   # 1 = weight-streaming kernel (lstm_tc.cu), 2 / 3 = cluster kernels (lstm_cluster.cu; 3 = tabulated input projection)
def test_tc_lstm_encode_within_tolerance(We, H, E, T, B, kern):
    if kern >= 2 and (H + 63) // 64 * 64 not in (64, 128, 256):
        pytest.skip("cluster kernels: H (padded to a multiple of 64) in {64,128,256}")
    mode, V = "dual-encoder", 5000
    h, p = make(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_TC)
    h.set_option("encoder", 2)          # force the tcgen05 path (raise if unsupported)
    h.set_option("lstm_kernel", kern)
    rng = np.random.default_rng(We + H + B)
    for side, name in ((sse_ffi.SIDE_SRC, "src"), (sse_ffi.SIDE_TGT, "tgt")):
        tok = np.concatenate([O.synth_tokens(rng, B - B // 2, T, V, "full"), O.synth_tokens(rng, B // 2, T, V, "real", 4.0)]) \
            if B > 1 else O.synth_tokens(rng, 1, T, V, "full")
        got = h.encode_host(side, tok, True)
        want = O.encode(p, mode, name, tok, True)
        err = np.abs(got - want).max()
        assert err < TOL_TC, (name, err)
        assert err > 0          # it really is the fp16-operand tensor-core path, not the fp32 one
        raw = h.encode_host(side, tok, False)
        wraw = O.encode(p, mode, name, tok, False)
        assert np.abs(raw - wraw).max() < 2e-2 * np.abs(wraw).max()
    h.close()


@pytest.mark.parametrize("kern", [1, 2, 3])
def test_tc_lstm_pad_skip_and_exact_mode_switch(kern):
    mode, V, We, H, E, T, B = "shared-encoder", 3000, 128, 128, 64, 40, 200
    h, p = make(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_TC)
    rng = np.random.default_rng(12)
    tok = O.synth_tokens(rng, B, T, V, "real", 3.0)
    want = O.encode(p, mode, "src", tok, True)
    h.set_option("encoder", 2)
    h.set_option("lstm_kernel", kern)
    h.set_option("pad_skip", 0)
    a = h.encode_host(0, tok, True)
    h.set_option("pad_skip", 1)
    b = h.encode_host(0, tok, True)           # starts every row from the pad-prefix state table
    assert np.abs(a - want).max() < TOL_TC and np.abs(b - want).max() < TOL_TC
    h.set_option("encoder", 1)                # exact fp32 SIMT mode on the same handle
    c = h.encode_host(0, tok, True)
    assert np.abs(c - want).max() < TOL_FP32
    # every LSTM shape has a tensor-core tower now (zero-padded tiles up to 256, GEMM-per-step above): forcing it never raises
    h2, p2 = make(mode, 100, 320, 64, 96, 96, 20, precision=sse_ffi.PRECISION_TC)
    h2.set_option("encoder", 2)
    t2 = O.synth_tokens(rng, 9, 20, 100, "real", 3.0)
    assert np.abs(h2.encode_host(0, t2, True) - O.encode(p2, mode, "src", t2, True)).max() < TOL_TC
    h2.close()
    h.close()


def test_tc_lstm_table_follows_parameter_updates():
    """The tabulated input projection (lstm_kernel 3) is derived state: it must be rebuilt after set_param."""
    mode, V, We, H, E, T, B = "dual-encoder", 900, 64, 64, 32, 10, 70
    h, p = make(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_TC)
    h.set_option("encoder", 2)
    h.set_option("lstm_kernel", 3)
    rng = np.random.default_rng(5)
    tok = O.synth_tokens(rng, B, T, V, "real", 2.0)
    a = h.encode_host(0, tok, True)
    assert np.abs(a - O.encode(p, mode, "src", tok, True)).max() < TOL_TC
    p2 = dict(p)
    p2["word_embedding"] = (p["word_embedding"] * 0.5 + 0.01).astype(np.float32)
    h.set_params({"word_embedding": p2["word_embedding"]})
    b = h.encode_host(0, tok, True)
    assert np.abs(b - O.encode(p2, mode, "src", tok, True)).max() < TOL_TC
    assert np.abs(a - b).max() > 1e-2
    h.close()


@pytest.mark.parametrize("mode,We,H,E,T,B,force", [("shared-encoder", 512, 512, 512, 20, 70, 0),      # BASELINE config 5 cell (makefile:17 QnA recipe, T shortened)
                                                  ("dual-encoder", 320, 384, 128, 12, 33, 0),         # wider than the resident-weight kernels hold
                                                  ("dual-encoder", 64, 64, 32, 12, 200, 4)])          # forced on a shape the other kernels also run
def test_tc_lstm_gemm_per_step_tower_within_tolerance(mode, We, H, E, T, B, force):
    V = 3000
    h, p = make(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_TC)
    h.set_option("encoder", 2)
    if force:
        h.set_option("lstm_kernel", force)
    rng = np.random.default_rng(We + H)
    tok = np.concatenate([O.synth_tokens(rng, B - B // 2, T, V, "full"), O.synth_tokens(rng, B // 2, T, V, "real", 4.0)])
    for side, name in ((sse_ffi.SIDE_SRC, "src"), (sse_ffi.SIDE_TGT, "tgt")):
        got = h.encode_host(side, tok, True)
        want = O.encode(p, mode, name, tok, True)
        err = np.abs(got - want).max()
        assert 0 < err < TOL_TC, (name, err)
    h.close()


# ---------------------------------------------------------------------------------------------
# round-2 variants of the table kernel: the cluster size must not change a single bit, the optional variants must stay in tolerance
@pytest.mark.parametrize("B", [1, 77, 600, 1300])
def test_tc_lstm_cluster_rows_64_and_128_are_bit_identical(B):
    mode, V, We, H, E, T = "dual-encoder", 5000, 256, 256, 256, 50
    h, p = make(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_TC)
    h.set_option("encoder", 2)
    h.set_option("lstm_kernel", 3)
    rng = np.random.default_rng(B)
    tok = np.concatenate([O.synth_tokens(rng, B - B // 2, T, V, "full"), O.synth_tokens(rng, B // 2, T, V, "real", 4.0)]) \
        if B > 1 else O.synth_tokens(rng, 1, T, V, "full")
    outs = {}
    for rows in (64, 128, 0):
        h.set_option("cluster_rows", rows)
        for skip in (0, 1):
            h.set_option("pad_skip", skip)
            outs[(rows, skip)] = h.encode_host(sse_ffi.SIDE_SRC, tok, True)
    ref = outs[(128, 0)]
    assert np.abs(ref - O.encode(p, mode, "src", tok, True)).max() < TOL_TC
    for key, got in outs.items():
        assert np.array_equal(got, ref), key
    with pytest.raises(Exception):
        h.set_option("cluster_rows", 96)
    h.close()


def test_query_batch_projection_matches_the_index_build_projection():
    """Query batches run projection + l2-norm in one launch (project_rows_kernel), larger batches the SIMT GEMM + l2norm_rows:
    the same rows must come out equal to fp32 rounding through both."""
    mode, V, We, H, E, T = "dual-encoder", 3000, 64, 128, 96, 20
    h, p = make(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_TC)
    rng = np.random.default_rng(5)
    tok = O.synth_tokens(rng, 4500, T, V, "real", 5.0)            # > 4096 rows: GEMM path
    big = h.encode_host(sse_ffi.SIDE_TGT, tok, True)
    small = h.encode_host(sse_ffi.SIDE_TGT, tok[:700], True)     # fused path
    raw_big = h.encode_host(sse_ffi.SIDE_TGT, tok, False)
    raw_small = h.encode_host(sse_ffi.SIDE_TGT, tok[:700], False)
    assert np.abs(big[:700] - small).max() < 2e-6
    assert np.abs(raw_big[:700] - raw_small).max() < 2e-5 * max(1.0, np.abs(raw_big).max())
    assert np.abs(small - O.encode(p, mode, "tgt", tok[:700], True)).max() < TOL_TC
    h.close()


_VARIANT_PROBE = r"""
import os, sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import sse_ffi, sse_oracle as O
mode, V, We, H, E, T, B = "dual-encoder", 4000, 256, 256, 256, 30, 333
p = O.init_params(mode, V, We, E, H, H, seed=3)
h = sse_ffi.Handle(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_TC)
h.set_params(p); h.set_option("encoder", 2); h.set_option("lstm_kernel", 3)
rng = np.random.default_rng(9)
tok = np.concatenate([O.synth_tokens(rng, B - B // 2, T, V, "full"), O.synth_tokens(rng, B // 2, T, V, "real", 4.0)])
want = O.encode(p, mode, "src", tok, True)
errs = []
for skip in (0, 1):
    h.set_option("pad_skip", skip)
    errs.append(float(np.abs(h.encode_host(0, tok, True) - want).max()))
print("ERR", max(errs))
"""


@pytest.mark.parametrize("env", [{"SSE_LSTM_VARIANT": "3", "SSE_LSTM_UNITS": "64"},
                                 {"SSE_LSTM_VARIANT": "3", "SSE_LSTM_UNITS": "32", "SSE_LSTM_STAGE": "1"},
                                 {"SSE_LSTM_VARIANT": "3", "SSE_LSTM_UNITS": "32", "SSE_LSTM_STAGE": "0"},
                                 {"SSE_LSTM_VARIANT": "2", "SSE_LSTM_EW": "8", "SSE_LSTM_GATE_MATH": "0"},
                                 {"SSE_LSTM_VARIANT": "2", "SSE_LSTM_GATE_MATH": "2"}])
def test_optional_table_kernel_variants_stay_within_tolerance(env):
    """The variants selected through the environment (read once per process, hence the subprocess): single-h-tile kernel with 64
    units per CTA / with staged table rows, 8 epilogue warps with the ex2 gate form, the f16x2 gate form."""
    import os, subprocess, sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", _VARIANT_PROBE, os.path.join(repo, "sequence-semantic-embedding_b200"), os.path.join(repo, "oracle")],
                       env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    err = float([l for l in r.stdout.splitlines() if l.startswith("ERR")][0].split()[1])
    assert 0 < err < TOL_TC, (env, err)
