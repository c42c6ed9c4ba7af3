"""CPU tests of the oracle: (1) retrieval half against fixtures produced by the REAL
reference code (tests/golden/make_golden.py); (2) encoder / loss / optimizer half --
which lives in TensorFlow 1.x and cannot run here ("parity unpinned") -- against
hand-computable known answers and independent torch implementations."""
import os

import numpy as np
import pytest
import torch

import sse_oracle as O


# ---------------------------------------------------------------- pinned by the reference
def test_ranking_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "ranking.npz"))
    k = g["top_idx"].shape[1]
    s, idx = O.retrieve(g["src"], g["tgt"].astype(np.float64), k)
    assert np.array_equal(idx, g["top_idx"])
    assert np.allclose(s, g["top_scores"], rtol=0, atol=1e-12)
    labels = [[int(v) for v in row if v >= 0] for row in g["labels"]]
    _, ranked = O.get_sorted_results(np.dot(g["src"], g["tgt"].astype(np.float64).T))
    acc = [O.topk_tight_accuracy(n, labels, ranked) for n in (1, 3, 10)] + \
          [O.topk_accuracy(n, labels, ranked) for n in (1, 3, 10)]
    assert np.allclose(acc, g["acc"], atol=1e-12)


def test_index_row_format_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "index_rows.npz"))
    lines = open(os.path.join(golden_dir, "index_rows.tsv"), encoding="utf-8").readlines()
    for i, line in enumerate(lines):
        assert O.format_index_row("id%d" % i, "some target text %d" % i, g["enc"][i]) == line
    ids, texts, enc = O.parse_index_lines(lines)
    assert ids == ["id%d" % i for i in range(5)]
    assert enc.dtype == np.float64 and np.array_equal(enc, g["parsed"])
    assert np.array_equal(enc.astype(np.float32), g["enc"])      # float32 round trip is exact


def test_padding_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "padding.npz"))
    T = int(g["T"])
    for raw, row in zip(g["raw"], g["padded"]):
        ids = [int(v) for v in raw if v >= 0]
        assert O.pad_tokens(ids, T) == row.tolist()
        assert row[0] == O.PAD_ID and row[-1] == O.EOS_ID and len(row) == T


# ---------------------------------------------------------------- known answers (TF semantics)
def test_lstm_zero_weights_forget_bias():
    V, We, H, T = 5, 3, 2, 4
    emb = np.ones((V, We), np.float32)
    K = np.zeros((We + H, 4 * H), np.float32)
    b = np.zeros(4 * H, np.float32)
    tok = np.zeros((1, T), np.int32)
    # K = 0, b = 0: i = j = f = o = 0 -> c_t = c_{t-1} * sigmoid(1) + 0.5 * tanh(0) = 0 ; h = 0
    assert np.all(O.lstm_last_state(tok, emb, K, b) == 0)
    # candidate bias only: j = 1 -> c_1 = 0.5 * tanh(1); c_2 = c_1 * sigmoid(1) + 0.5 tanh(1)  (forget_bias = 1.0)
    b2 = b.copy(); b2[H:2 * H] = 1.0
    h = O.lstm_last_state(tok[:, :2], emb, K, b2, dtype=np.float64)
    c1 = 0.5 * np.tanh(1.0)
    c2 = c1 / (1 + np.exp(-1.0)) + 0.5 * np.tanh(1.0)
    assert np.allclose(h, np.tanh(c2) * 0.5, atol=1e-12)


def test_lstm_hand_computed_step():
    # H = We = 1, one step, scalar check of gate order (i, j, f, o) and the [x; h] row order
    emb = np.array([[2.0]], np.float64)
    K = np.array([[0.1, 0.2, 0.3, 0.4], [9, 9, 9, 9]], np.float64)   # h row irrelevant at t = 0
    b = np.array([0.01, 0.02, 0.03, 0.04], np.float64)
    h = O.lstm_last_state(np.zeros((1, 1), np.int32), emb, K, b, dtype=np.float64)
    sg = lambda v: 1 / (1 + np.exp(-v))
    c = sg(0.21) * np.tanh(0.42)
    assert np.allclose(h, np.tanh(c) * sg(0.84), atol=1e-14)


def test_l2_normalize_epsilon():
    x = np.array([[3.0, 4.0], [0.0, 0.0], [1e-8, 0.0]], np.float32)
    n = O.l2_normalize(x)
    assert np.allclose(n[0], [0.6, 0.8], atol=1e-7)
    assert np.all(n[1] == 0)
    assert np.allclose(n[2, 0], 1e-8 / 1e-6, rtol=1e-5)      # sum x^2 = 1e-16 < eps -> divide by sqrt(1e-12)


def test_top_k_tf_tie_rule():
    sim = np.array([[0.5, 0.9, 0.9, 0.1, 0.5]], np.float32)
    s, i = O.top_k_tf(sim, 4, normalize_scores=False)
    assert i.tolist() == [[1, 2, 0, 4]]
    s2, _ = O.top_k_tf(sim, 4)
    assert np.allclose(np.linalg.norm(s2, axis=1), 1.0, atol=1e-6)


def test_pair_loss_known_values():
    cos = np.array([1.0, -1.0, 0.0], np.float64)
    lab = np.array([1.0, 0.0, 1.0], np.float64)
    loss, acc = O.pair_loss_acc(cos, lab)
    # x = 64, -64, 0: wce = log1p(exp(-64)), log1p(exp(-64)), log 2
    assert np.isclose(loss, (2 * np.log1p(np.exp(-64.0)) + np.log(2.0)) / 3)
    # acc = mean(l*floor(sig+0.1)) + mean((1-l)*floor(1.1-sig)) = (1 + 0)/3 + 1/3
    assert np.isclose(acc, 2.0 / 3.0)


# ---------------------------------------------------------------- cross-checks with torch
def _rand_tokens(rng, B, T, V, mean=3.0):
    return O.synth_tokens(rng, B, T, V, "real", mean)


@pytest.mark.parametrize("We,H,E,T", [(12, 16, 10, 9), (50, 96, 64, 20)])
def test_lstm_matches_torch_lstm(We, H, E, T):
    V, B = 200, 7
    rng = np.random.default_rng(0)
    p = O.init_params("dual-encoder", V, We, E, H, H, seed=1)
    p["source_encoder/rnn/basic_lstm_cell/bias"] = (rng.normal(size=4 * H) * 0.1).astype(np.float32)
    tok = _rand_tokens(rng, B, T, V)
    out = O.encode(p, "dual-encoder", "src", tok)
    K, b = p["source_encoder/rnn/basic_lstm_cell/kernel"], p["source_encoder/rnn/basic_lstm_cell/bias"]
    perm = np.concatenate([np.arange(0, H), np.arange(2 * H, 3 * H), np.arange(H, 2 * H), np.arange(3 * H, 4 * H)])
    Kp, bp = K[:, perm], b[perm].copy()
    bp[H:2 * H] += 1.0                                      # forget_bias folded into torch's f bias
    lstm = torch.nn.LSTM(We, H, batch_first=True)
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(torch.tensor(Kp[:We].T)); lstm.weight_hh_l0.copy_(torch.tensor(Kp[We:].T))
        lstm.bias_ih_l0.copy_(torch.tensor(bp)); lstm.bias_hh_l0.zero_()
        x = torch.tensor(p["word_embedding"])[torch.tensor(tok, dtype=torch.long)]
        _, (h, _) = lstm(x)
        n = torch.nn.functional.normalize(h[0] @ torch.tensor(p["source_encoder/src_M"]), dim=-1).numpy()
    assert np.abs(n - out).max() < 2e-6


def test_cnn_matches_torch_conv():
    V, We, E, T, B = 100, 10, 8, 12, 5
    rng = np.random.default_rng(3)
    p = O.init_params("dual-cnn", V, We, E, 0, 0, seed=5, cnn_filter_sizes=(2, 3, 5), cnn_num_filters=(6, 4, 3))
    tok = _rand_tokens(rng, B, T, V, 5.0)
    out = O.encode(p, "dual-cnn", "tgt", tok, normalize=False)
    x = torch.tensor(p["word_embedding"])[torch.tensor(tok, dtype=torch.long)]          # [B,T,We]
    feats = []
    for k in (2, 3, 5):
        W = torch.tensor(p["target_cnn/conv-maxpool-%d/W" % k])                            # [k,We,1,F]
        bb = torch.tensor(p["target_cnn/conv-maxpool-%d/b" % k])
        # conv2d NHWC VALID == conv2d NCHW with in_channels=1 over the [T,We] image
        y = torch.nn.functional.conv2d(x[:, None], W.permute(3, 2, 0, 1), bb)              # [B,F,T-k+1,1]
        feats.append(torch.relu(y).amax(dim=2)[:, :, 0])
    u = torch.cat(feats, 1) @ torch.tensor(p["target_cnn/tgt_M"])
    assert np.abs(u.numpy() - out).max() < 1e-5


def _torch_train_step(p, mode, src, tgt, labels):
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}

    def enc(side, tok):
        scope, mn = O.tower_names(mode, side)
        K, b = tp[scope + "/rnn/basic_lstm_cell/kernel"], tp[scope + "/rnn/basic_lstm_cell/bias"]
        H = K.shape[1] // 4
        x_all = tp["word_embedding"][torch.tensor(tok, dtype=torch.long)]
        h = torch.zeros(tok.shape[0], H, dtype=torch.float64); c = torch.zeros_like(h)
        for t in range(tok.shape[1]):
            z = torch.cat([x_all[:, t], h], 1) @ K + b
            i, j, f, o = z.split(H, 1)
            c = c * torch.sigmoid(f + 1) + torch.sigmoid(i) * torch.tanh(j)
            h = torch.tanh(c) * torch.sigmoid(o)
        u = h @ tp[mn]
        return u * torch.rsqrt(torch.clamp((u * u).sum(-1, keepdim=True), min=1e-12))

    cos = (enc("src", src) * enc("tgt", tgt)).sum(-1)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(64 * cos, torch.tensor(labels, dtype=torch.float64))
    loss.backward()
    return loss.item(), {k: v.grad.numpy() for k, v in tp.items() if v.grad is not None}


@pytest.mark.parametrize("mode", ["dual-encoder", "shared-encoder"])
def test_train_step_gradients_match_autograd(mode):
    V, We, E, H, T, B = 60, 6, 5, 8, 7, 6
    rng = np.random.default_rng(4)
    p = O.init_params(mode, V, We, E, H, H, seed=2)
    src, tgt = _rand_tokens(rng, B, T, V), _rand_tokens(rng, B, T, V, 4.0)
    labels = (np.arange(B) % 2 == 0).astype(np.float32)
    tl, tg = _torch_train_step(p, mode, src, tgt, labels)
    st = O.TrainState({k: v.copy() for k, v in p.items()}, learning_rate=0.9)
    loss, acc, gn, g = O.train_step(st, mode, src, tgt, labels, dtype=np.float64, return_grads=True)
    assert abs(loss - tl) < 1e-10
    for k in g:
        assert np.abs(g[k] - tg[k]).max() < 1e-9 * max(1.0, np.abs(tg[k]).max()), k


def test_adagrad_and_clip_semantics():
    """dense var: acc += g^2; w -= lr*g/sqrt(acc) with acc0 = 0.1; clip scale = 5/max(gnorm,5);
    embedding: norm over un-merged slices, update over merged rows."""
    V, We, E, H, T, B = 30, 4, 3, 4, 5, 4
    rng = np.random.default_rng(9)
    p = O.init_params("dual-encoder", V, We, E, H, H, seed=3)
    src = np.tile(_rand_tokens(rng, 1, T, V), (B, 1))        # identical rows -> heavy duplication of embedding rows
    tgt = _rand_tokens(rng, B, T, V)
    labels = np.array([1, 0, 1, 0], np.float32)
    st = O.TrainState({k: v.copy() for k, v in p.items()}, learning_rate=0.5)
    loss, acc, gn, g = O.train_step(st, "dual-encoder", src, tgt, labels, dtype=np.float64, return_grads=True)
    scale = 5.0 / max(gn, 5.0)
    name = "source_encoder/src_M"
    gg = g[name] * scale
    assert np.allclose(st.params[name], p[name] - 0.5 * gg / np.sqrt(0.1 + gg * gg), atol=1e-6)
    # un-merged norm >= merged norm when rows repeat with same-sign slices; check the definition directly
    _, tg = _torch_train_step(p, "dual-encoder", src, tgt, labels)
    merged_sq = sum(float((v ** 2).sum()) for v in tg.values())
    assert gn ** 2 >= 0 and abs(gn ** 2 - merged_sq) > 0    # duplication makes the two norms differ
    touched = np.unique(np.concatenate([src.ravel(), tgt.ravel()]))
    untouched = np.setdiff1d(np.arange(V), touched)
    assert np.array_equal(st.params["word_embedding"][untouched], p["word_embedding"][untouched])
    ge = g["word_embedding"][touched] * scale
    assert np.allclose(st.params["word_embedding"][touched],
                       p["word_embedding"][touched] - 0.5 * ge / np.sqrt(0.1 + ge * ge), atol=1e-6)
    assert st.global_step == 1


def test_lr_decay_floor():
    st = O.TrainState({"w": np.zeros(1, np.float32)}, learning_rate=0.0011)
    assert O.lr_decay(st, 0.5) == pytest.approx(1e-3)
    st.learning_rate = 0.9
    assert O.lr_decay(st, 0.99) == pytest.approx(np.float32(0.9) * np.float32(0.99))


def test_train_loop_policy_matches_reference_rules():
    """Host logic of the training entry point (sse_train.py:196-215 of the reference): lr decay / BestEver / the dead
    early-stop branch, and the windowed means."""
    import sse_train
    pol = sse_train.CheckpointPolicy()
    accs = [0.5, 0.6, 0.7, 0.65, 0.66, 0.67, 0.68, 0.60, 0.69, 0.9]
    out = [pol.report(a, epoch=12) for a in accs]
    assert [o["save_best"] for o in out] == [True, True, True, False, False, False, False, False, False, True]
    # decay needs more than six earlier reports and a value below the minimum of the last five of them
    assert [o["decay_lr"] for o in out] == [False] * 7 + [True, False, False]
    assert not any(o["finished"] for o in out)          # the reference's own early stop can never fire (checked after append)
    assert out[-1]["best"] == 0.9
    w = sse_train.WindowStats(4)
    for s in range(4):
        w.add(2.0, 1.0 + s, 0.5)
    assert abs(w.step_time - 2.0) < 1e-12 and abs(w.loss - 2.5) < 1e-12 and abs(w.train_acc - 0.5) < 1e-12
    w.reset()
    assert w.loss == 0.0


def test_train_entry_point_flow_with_stubbed_compute(tmp_path, monkeypatch):
    """Drives sse_train.main() end to end on the CPU with the compute (model / session / index / evaluator / data)
    replaced by recorders: checks the order of operations, checkpoint names and the log lines the reference emits."""
    import logging
    import sse_train

    calls = []

    class T(object):
        def __init__(self, v): self.v = v
        def eval(self): return self.v

    class FakeModel(object):
        train, loss, train_acc, learning_rate_decay_op = "train", "loss", "acc", "decay"
        def __init__(self, params, device=0):
            self.global_step, self.learning_rate, self.saver = T(0), T(float(params["learning_rate"])), self
        def add_summaries(self): return "summary"
        def set_forward_only(self, f): calls.append(("forward_only", f))
        def get_train_feed_dict(self, s, t, l): return {"s": s}
        def save(self, sess, path): calls.append(("save", os.path.basename(path))); return path
        def restore(self, sess, path): calls.append(("restore", path))

    class FakeSession(object):
        def __init__(self, seed=None): self.n = 0
        def __enter__(self): return self
        def __exit__(self, *a): return False
        def run(self, fetches, feed_dict=None):
            if fetches == "decay": calls.append(("decay",)); return None
            if not isinstance(fetches, list): calls.append(("init",)); return None
            self.n += 1
            model.global_step.v = self.n
            return [None, None, 1.0 / self.n, min(0.99, 0.5 + 0.01 * self.n)]

    class FakeData(object):
        def __init__(self, *a, **k):
            self.rawTrainPosCorpus, self.rawnegSetLen, self.vocab_size, self.encoder, self.rawEvalCorpus = list(range(40)), 7, 100, None, []
        def get_train_batch(self, bs): return [[0]], [[0]], [1.0]

    model = None

    def make_model(params, device=0):
        nonlocal model
        model = FakeModel(params, device)
        return model

    class FakeEval(object):
        def __init__(self, *a): calls.append(("evaluator",))
        def eval(self): return 0.1, 0.3, 0.9

    monkeypatch.setattr(sse_train, "Data", FakeData)
    monkeypatch.setattr(sse_train.sse_model, "SSEModel", make_model)
    monkeypatch.setattr(sse_train.sse_model, "Session", FakeSession)
    monkeypatch.setattr(sse_train.sse_model, "get_checkpoint_state", lambda d: None)
    monkeypatch.setattr(sse_train.sse_index, "createIndexFile", lambda *a, **k: calls.append(("index", k.get("batchsize"))))
    monkeypatch.setattr(sse_train.sse_evaluator, "Evaluator", FakeEval)
    md = str(tmp_path / "m")
    root = logging.getLogger("")
    before = list(root.handlers)
    try:
        sse_train.main(["--data_dir", str(tmp_path / "d"), "--model_dir", md, "--batch_size", "10", "--steps_per_checkpoint", "2",
                        "--max_epoc", "2", "--task_type", "classification"])
    finally:
        for h in list(root.handlers):
            if h not in before:
                root.removeHandler(h); h.close()
    # 40 positives / batch 10 = 4 steps per epoch, report every 2 steps: 2 reports per epoch, accuracies rise -> BestEver each time
    saves = [c[1] for c in calls if c[0] == "save"]
    assert saves == ["SSE-LSTM.ckpt-BestEver", "SSE-LSTM.ckpt-BestEver", "SSE-LSTM.ckpt-epoch-0",
                     "SSE-LSTM.ckpt-BestEver", "SSE-LSTM.ckpt-BestEver", "SSE-LSTM.ckpt-epoch-1"]
    assert [c for c in calls if c[0] in ("index", "evaluator")] == [("index", 1000), ("evaluator",)] * 2
    assert ("init",) in calls and ("decay",) not in calls
    log = open(os.path.join(md, "TrainingLog.txt")).read()
    assert "Created model with fresh parameters." in log and "Better Accuracy" in log
    assert "epoc#1, task specific evaluation: top 1/3/10 accuracies: 0.100000 / 0.300000 / 0.900000" in log
    assert "global epoc: 0.500, global step 2, learning rate 0.9000" in log
    assert os.path.exists(os.path.join(md, "modelConfig.param"))
