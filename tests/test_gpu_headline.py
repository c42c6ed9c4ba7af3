"""GPU parity at the HEADLINE configuration (BASELINE.json metric: 256-d dual LSTM, T=50, top-10 over a 1M-target index,
600-query batches; reference sse_evaluator.py:95-114, sse_index.py:55-97, sse_model.py:217-233,344-350): the tcgen05
scan against the reference's float64 np.dot + descending sort on sampled query rows, once over a gaussian index and
once over an index BUILT BY THE TARGET ENCODER from REAL-regime target tokens (clustered encodings: the hard case for
the sampled-threshold filter)."""
import numpy as np
import pytest

import sse_ffi
import sse_oracle as O

pytestmark = pytest.mark.gpu

V, WE, H, E, T, K = 32000, 256, 256, 256, 50, 10
N, Q, SAMPLE = 1_000_000, 600, 64


def _check_rows(q, tgt, got_s, got_i, rows):
    """float64 scores of the sampled rows against ALL targets (sse_evaluator.py:110), descending order
    (data_utils.py:263-267); indices must be exact wherever the neighbouring scores are further apart than fp32 noise."""
    d = q[rows].astype(np.float64) @ tgt.astype(np.float64).T            # [SAMPLE, N]
    part = np.argpartition(-d, K + 1, axis=1)[:, :K + 2]
    ps = np.take_along_axis(d, part, 1)
    order = np.argsort(-ps, axis=1, kind="stable")
    top_i = np.take_along_axis(part, order, 1)
    top_s = np.take_along_axis(ps, order, 1)
    assert np.abs(got_s[rows] - top_s[:, :K]).max() < 1e-5 * max(1.0, np.abs(top_s).max())
    gaps = -np.diff(top_s, axis=1)                                        # [SAMPLE, K+1] gaps between consecutive ranks
    wide = np.minimum(gaps[:, :K], np.concatenate([np.full((len(rows), 1), 1.0), gaps[:, :K - 1]], 1)) > 2e-6
    assert wide.mean() > 0.9
    assert np.array_equal(got_i[rows][wide], top_i[:, :K][wide])


@pytest.fixture(scope="module")
def model():
    p = O.init_params("dual-encoder", V, WE, E, H, H, seed=1234)
    h = sse_ffi.Handle("dual-encoder", V, WE, E, H, H, T, precision=sse_ffi.PRECISION_TC)
    h.set_params(p)
    yield h, p
    h.close()


def _search(h, q):
    import torch
    dq = torch.from_numpy(q).cuda()
    s = torch.empty(q.shape[0], K, device="cuda")
    i = torch.empty(q.shape[0], K, device="cuda", dtype=torch.int32)
    h.set_option("search", 2)                                            # the tcgen05 path or an error, never a silent SIMT run
    h.search(dq, q.shape[0], K, s, i)
    torch.cuda.synchronize()
    return s.cpu().numpy(), i.cpu().numpy()


def test_headline_shape_gaussian_index(model):
    h, _p = model
    rng = np.random.default_rng(7)
    tgt = rng.standard_normal((N, E), dtype=np.float32)
    tgt /= np.linalg.norm(tgt, axis=1, keepdims=True)
    q = rng.standard_normal((Q, E), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    planted = rng.integers(0, N, size=Q // 2)
    q[:Q // 2] = tgt[planted] + 0.05 * rng.standard_normal((Q // 2, E), dtype=np.float32)
    h.index_set(tgt)
    s, i = _search(h, q)
    assert np.array_equal(i[:Q // 2, 0], planted)
    rows = rng.choice(Q, size=SAMPLE, replace=False)
    _check_rows(q, tgt, s, i, rows)


def test_headline_shape_encoder_built_index_real_regime(model):
    h, p = model
    rng = np.random.default_rng(11)
    ttok = O.synth_tokens(rng, N, T, V, "real", 8.0)                     # titles: ~8 subtokens of 50
    qtok = O.synth_tokens(rng, Q, T, V, "real", 3.0)                     # queries: ~3
    h.index_build(ttok, batch=16384)
    tgt = h.index_get(0, N)
    # the index rows are the target encoder's output (spot-check against the oracle) and unit length
    some = rng.choice(N, size=256, replace=False)
    assert np.abs(tgt[some] - O.encode(p, "dual-encoder", "tgt", ttok[some], True)).max() < 1e-3
    assert np.abs(np.linalg.norm(tgt, axis=1) - 1).max() < 1e-4
    s, i = h.query_host(qtok, K, True)                                    # tokens -> source encoder -> scan -> [Q,k]
    q = h.encode_host(sse_ffi.SIDE_SRC, qtok, True)
    rows = rng.choice(Q, size=SAMPLE, replace=False)
    _check_rows(q, tgt, s, i, rows)
    # and the encoder half of the query path against the oracle on the same rows
    assert np.abs(q[rows] - O.encode(p, "dual-encoder", "src", qtok[rows], True)).max() < 1e-3
