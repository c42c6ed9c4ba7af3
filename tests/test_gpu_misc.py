"""GPU tests of the path's smaller entry points: source-encoder-only mode (reference sse_model.py:217-233), the in-graph
top-k tensors predicted_tgts_score / predicted_labels (:344-350), sse_l2_normalize_rows (:282-283,350), the device-side
token pre-pass (range check = TF's InvalidArgument; per-row pad-prefix start, data_utils.py:149-155) and the serving
micro-batcher against concurrent single-query callers (webserver.py:124-286)."""
import threading

import numpy as np
import pytest

import sse_ffi
import sse_model
import sse_oracle as O
import sse_serve

pytestmark = pytest.mark.gpu


def _model(mode, V, We, E, Hs, Ht, T, N=0, precision=0, nbest=5):
    p = O.init_params(mode, V, We, E, Hs, Ht, seed=5, target_space_size=N)
    m = sse_model.SSEModel(dict(forward_only=True, network_mode=mode, predict_nbest=nbest, max_seq_length=T, vocab_size=V,
                                embedding_size=We, encoding_size=E, src_cell_size=Hs, tgt_cell_size=Ht, learning_rate=0.5,
                                learning_rate_decay_factor=0.99, targetSpaceSize=N), precision=precision)
    m.handle.set_params(p)
    m._initialized = True
    return m, p


def test_source_encoder_only_mode_encodes_and_serves_its_target_table():
    mode, V, We, E, H, T, N = "source-encoder-only", 900, 24, 20, 40, 12, 57
    m, p = _model(mode, V, We, E, H, H, T, N=N)
    sess = sse_model.Session()
    rng = np.random.default_rng(1)
    tok = O.synth_tokens(rng, 33, T, V, "real", 4.0)
    got = sess.run([m.norm_src_seq_embedding], m.get_source_encoding_feed_dict(tok.tolist()))[0]
    assert np.abs(got - O.encode(p, mode, "src", tok, True)).max() < 2e-5
    raw = sess.run(m.src_seq_embedding, m.get_source_encoding_feed_dict(tok.tolist()))
    assert np.abs(raw - O.encode(p, mode, "src", tok, False)).max() < 2e-5 * max(1.0, np.abs(raw).max())
    tgt = sess.run(m.norm_tgt_seq_embedding, {})                          # the whole [targetSpaceSize, E] table (:233)
    want = O.l2_normalize(p["target_embedding/tgt_seq_embedding"])
    assert tgt.shape == (N, E) and np.abs(tgt - want).max() < 1e-6
    with pytest.raises(sse_ffi.SseError):                                 # there is no target ENCODER in this mode
        m.handle.encode_host(sse_ffi.SIDE_TGT, tok, True)
    m.handle.close()


def test_predicted_scores_and_labels_follow_tf_top_k_and_leave_the_index_alone():
    mode, V, We, E, H, T = "dual-encoder", 700, 16, 24, 32, 10
    m, p = _model(mode, V, We, E, H, H, T, nbest=5)
    sess = sse_model.Session()
    rng = np.random.default_rng(2)
    src = O.synth_tokens(rng, 21, T, V, "real", 3.0)
    tgt = O.synth_tokens(rng, 40, T, V, "real", 5.0)
    tgt[7] = tgt[3]                                                       # an exact tie: the lower index wins (TF rule)
    resident = rng.standard_normal((300, E)).astype(np.float32)
    m.handle.index_set(resident)
    scores, labels = sess.run([m.predicted_tgts_score, m.predicted_labels], m.get_predict_feed_dict(src.tolist(), tgt.tolist()))
    sim = O.similarity(O.encode(p, mode, "src", src, True), O.encode(p, mode, "tgt", tgt, True))
    ws, wi = O.top_k_tf(sim, 5, normalize_scores=True)
    srt = -np.sort(-sim, axis=1)[:, :6]
    wide = np.min(np.abs(np.diff(srt, axis=1)), axis=1) > 1e-5            # rows without fp32-level near-ties besides the planted one
    assert wide.sum() >= 10
    assert np.array_equal(labels[wide], wi[wide])
    assert np.abs(scores - ws).max() < 1e-4
    assert np.allclose(np.linalg.norm(scores, axis=1), 1.0, atol=1e-5)    # l2_normalize(scores, 1), sse_model.py:350
    for r in range(len(src)):                                             # wherever rows 3 and 7 both appear, 3 comes first
        l = labels[r].tolist()
        if 3 in l and 7 in l:
            assert l.index(3) < l.index(7)
    # the index registered on the handle is still the one we put there
    assert np.array_equal(m.handle.index_get(0, 300), resident)
    m.set_top_n(41)
    with pytest.raises((ValueError, sse_ffi.SseError)):                   # TF: "input must have at least k columns"
        sess.run(m.predicted_labels, m.get_predict_feed_dict(src.tolist(), tgt.tolist()))
    m.handle.close()


def test_l2_normalize_rows_matches_tf_definition():
    import torch
    h = sse_ffi.Handle("dual-encoder", 50, 8, 8, 8, 8, 8)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((37, 50)).astype(np.float32) * 3
    x[5] = 0.0                                                            # zero row: x * rsqrt(max(0, 1e-12)) = 0
    x[6] *= 1e-8                                                          # below the epsilon: scaled by 1e6, not to unit length
    d = torch.from_numpy(x).cuda()
    h.l2_normalize_rows(d, 37, 50)
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    want = O.l2_normalize(x)
    assert np.abs(got - want).max() < 1e-6
    assert np.all(got[5] == 0)
    h.close()


@pytest.mark.parametrize("precision,kern", [(sse_ffi.PRECISION_FP32, 0), (sse_ffi.PRECISION_TC, 3)])
def test_pad_prefix_start_is_bit_identical_on_mixed_length_batches_through_sse_encode(precision, kern):
    """Device-pointer entry point, rows of every length in one batch (some tiles mix 1 and 40 leading PADs): bucketing by
    prefix length + starting each tile from the tabulated state must give the bits of the full T-step run."""
    import torch
    mode, V, We, E, H, T, B = "dual-encoder", 3000, 64, 64, 128, 50, 700
    p = O.init_params(mode, V, We, E, H, H, seed=9)
    h = sse_ffi.Handle(mode, V, We, E, H, H, T, precision=precision)
    h.set_params(p)
    if kern:
        h.set_option("encoder", 2)
        h.set_option("lstm_kernel", kern)
    rng = np.random.default_rng(6)
    tok = np.concatenate([O.synth_tokens(rng, 300, T, V, "real", 3.0), O.synth_tokens(rng, 250, T, V, "real", 9.0),
                          O.synth_tokens(rng, 149, T, V, "full"), np.array([O.pad_tokens([], T)], np.int32)])
    tok = tok[rng.permutation(len(tok))]
    d = torch.from_numpy(tok).cuda()
    outs = []
    for skip in (0, 1):
        h.set_option("pad_skip", skip)
        o = torch.empty(B, E, device="cuda")
        h.encode(sse_ffi.SIDE_SRC, d, B, o, True)
        torch.cuda.synchronize()
        outs.append(o.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    tol = 2e-5 if precision == sse_ffi.PRECISION_FP32 else 1e-3
    assert np.abs(outs[1] - O.encode(p, mode, "src", tok, True)).max() < tol
    assert h.token_errors() == 0
    h.close()


def test_out_of_range_token_ids_are_reported_not_dereferenced():
    import torch
    mode, V, We, E, H, T = "dual-encoder", 100, 16, 16, 32, 8
    p = O.init_params(mode, V, We, E, H, H, seed=1)
    h = sse_ffi.Handle(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_FP32)
    h.set_params(p)
    tok = O.synth_tokens(np.random.default_rng(0), 9, T, V, "real", 3.0)
    good = h.encode_host(0, tok, True)
    bad = tok.copy()
    bad[2, T - 2] = V                                                     # first id past the table
    bad[4, T - 3] = -7
    with pytest.raises(sse_ffi.SseError, match="token id"):
        h.encode_host(0, bad, True)
    assert np.array_equal(h.encode_host(0, tok, True), good)              # the error is not sticky across calls
    d = torch.from_numpy(bad).cuda()
    o = torch.empty(9, E, device="cuda")
    h.encode(0, d, 9, o, True)                                            # asynchronous entry point: counted, read as PAD
    assert h.token_errors() == 2 and h.token_errors() == 0
    as_pad = bad.copy()
    as_pad[2, T - 2] = 0
    as_pad[4, T - 3] = 0
    assert np.abs(o.cpu().numpy() - O.encode(p, mode, "src", as_pad, True)).max() < 2e-5
    lab = np.array([1.0, 0.0] * 4 + [1.0], np.float32)
    with pytest.raises(sse_ffi.SseError, match="token id"):
        h.train_step(bad, tok, lab)
    h.close()


def test_micro_batcher_coalesces_concurrent_single_query_callers_on_the_gpu():
    mode, V, We, E, H, T, N = "dual-encoder", 2000, 64, 64, 64, 20, 20000
    p = O.init_params(mode, V, We, E, H, H, seed=3)
    h = sse_ffi.Handle(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_TC)
    h.set_params(p)
    rng = np.random.default_rng(4)
    h.index_build(O.synth_tokens(rng, N, T, V, "real", 8.0), batch=8192)
    qtok = O.synth_tokens(rng, 64, T, V, "real", 3.0)
    want_s, want_i = h.query_host(qtok, 10, False)                       # one batched call = the reference answer per row
    lock = threading.Lock()

    def query_fn(tokens, k, normalize):                                  # a handle is single-threaded: the batcher's worker is its only user
        with lock:
            return h.query_host(tokens, k, normalize)

    results = [None] * 64
    with sse_serve.MicroBatcher(query_fn, max_batch=64, max_wait_ms=50.0) as mb:
        def client(j):
            results[j] = mb.query(qtok[j], 10, False, timeout=60)
        threads = [threading.Thread(target=client, args=(j,)) for j in range(64)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        calls = mb.calls
    assert calls <= 8                                                    # 64 concurrent Q=1 requests shared a handful of scans
    for j in range(64):
        s, i = results[j]
        assert np.array_equal(i, want_i[j])
        assert np.abs(s - want_s[j]).max() < 1e-5 * max(1.0, np.abs(want_s).max())
    h.close()


def test_device_side_train_batch_sampler_follows_the_reference_rule():
    """data.py:95-115 on the device: window of consecutive positives, rows alternate (source, a verified target, 1.0),
    (same source, a NON-verified target, 0.0); reproducible per (seed, step); feeds sse_train_step directly."""
    import torch
    mode, V, We, E, H, T = "dual-encoder", 300, 16, 16, 32, 8
    P, N, B = 500, 60, 64
    rng = np.random.default_rng(12)
    src_rows = O.synth_tokens(rng, P, T, V, "real", 3.0)
    tgt_rows = O.synth_tokens(rng, N, T, V, "full")                    # full-length rows: distinct, so a row identifies its target
    tgt_rows[:, 1] = 2 + np.arange(N)
    counts = rng.integers(1, 4, size=P)
    ver_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ver_rows = np.concatenate([rng.choice(N, size=c, replace=False) for c in counts]).astype(np.int32)
    p = O.init_params(mode, V, We, E, H, H, seed=2)
    h = sse_ffi.Handle(mode, V, We, E, H, H, T, precision=sse_ffi.PRECISION_FP32)
    h.set_params(p)
    h.sampler_set(src_rows, ver_off, ver_rows, tgt_rows)
    src = torch.empty(2 * B, T, dtype=torch.int32, device="cuda")
    tgt = torch.empty_like(src)
    lab = torch.empty(2 * B, device="cuda")
    start = 200
    n = h.sampler_batch(start, B, 7, 3, src, tgt, lab)
    torch.cuda.synchronize()
    assert n == 2 * B
    s_, t_, l_ = src.cpu().numpy(), tgt.cpu().numpy(), lab.cpu().numpy()
    assert np.array_equal(l_, np.tile(np.array([1.0, 0.0], np.float32), B))
    tgt_key = {tuple(r): j for j, r in enumerate(map(tuple, tgt_rows))}
    assert len(tgt_key) == N                                             # distinct target rows, so rows identify targets
    negs = []
    for w in range(B):
        i = start + w
        ver = set(ver_rows[ver_off[i]:ver_off[i + 1]].tolist())
        assert np.array_equal(s_[2 * w], src_rows[i]) and np.array_equal(s_[2 * w + 1], src_rows[i])
        assert tgt_key[tuple(t_[2 * w])] in ver                          # the positive pair's target is verified for this source
        assert tgt_key[tuple(t_[2 * w + 1])] not in ver                  # the negative is not
        negs.append(tgt_key[tuple(t_[2 * w + 1])])
    assert len(set(negs)) > B // 4                                       # negatives spread over the target space
    src2, tgt2, lab2 = torch.empty_like(src), torch.empty_like(tgt), torch.empty_like(lab)
    h.sampler_batch(start, B, 7, 3, src2, tgt2, lab2)
    assert torch.equal(tgt, tgt2)                                        # same (seed, step) -> same draw
    h.sampler_batch(start, B, 7, 4, src2, tgt2, lab2)
    assert not torch.equal(tgt, tgt2)
    assert h.sampler_batch(P - 10, B, 7, 5, src2, tgt2, lab2) == 20      # short window at the end of the corpus (data.py:97-98)
    loss, acc, gn = h.train_step(src, tgt, lab)                          # device buffers feed the train step as they are
    assert np.isfinite(loss) and gn > 0
    with pytest.raises(sse_ffi.SseError):
        h.sampler_batch(P, B, 7, 5, src2, tgt2, lab2)
    h.close()
