"""CPU oracle for the SSE dual-encoder hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, the algorithm of the reference's hot path
(eBay/Sequence-Semantic-Embedding, TensorFlow-1 graph + numpy ranking).  It is
the checker the CUDA path is compared against; it is never the thing shipped
or measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.

Pinning status
--------------
* Retrieval half (``get_sorted_results``, ``topk_tight_accuracy``,
  ``topk_accuracy``, the index TSV reader/writer) is PINNED: the real reference
  code (``data_utils.getSortedResults`` / ``computeTopK_*`` /
  ``sse_evaluator.Evaluator.__init__`` parsing) was imported in the build
  container through ``oracle/tf_stub`` and its outputs are committed under
  ``tests/golden/`` (generator: ``tests/golden/make_golden.py``).
* Encoder / loss / optimizer half is **parity unpinned**: it lives inside
  TensorFlow 1.x (un-pinned in the reference's requirements.txt:1), which is
  not installed and cannot be installed offline.  Those functions follow the
  published TF-1.x op semantics and are cross-validated against independent
  implementations (``torch.nn.LSTM``, ``torch.nn.functional.conv1d``, torch
  autograd + ``torch.optim.Adagrad``) in ``tests/test_oracle.py``.

Every function cites the reference file:line it follows (paths are relative to
the reference repo root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

PAD_ID = 0  # text_encoder.py:44
EOS_ID = 1  # text_encoder.py:45

MODES = ("dual-encoder", "shared-encoder", "source-encoder-only", "source_only_cnn", "dual-cnn")


# --------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------
def pad_tokens(ids: Sequence[int], T: int) -> List[int]:
    """Left-pad / truncate rule.  data_utils.py:149-155, sse_index.py:79-85,
    sse_demo.py:115-119."""
    ids = list(ids)
    if len(ids) > T - 2:
        return [PAD_ID] + ids[: T - 2] + [EOS_ID]
    return [PAD_ID] * (T - len(ids) - 1) + ids + [EOS_ID]


def synth_tokens(rng: np.random.Generator, B: int, T: int, V: int, regime: str = "full",
                 mean_len: float = 8.0) -> np.ndarray:
    """Synthetic token rows in the reference layout.  ids follow a
    Zipf-like law over [2, V); 'full' = L = T-2 real tokens (one leading PAD),
    'real' = L ~ clip(Poisson(mean_len), 1, T-2)."""
    out = np.zeros((B, T), dtype=np.int32)
    for r in range(B):
        L = T - 2 if regime == "full" else int(np.clip(rng.poisson(mean_len), 1, T - 2))
        u = rng.random(L)
        ids = 2 + np.floor((V - 2) * u ** 3).astype(np.int64)
        ids = np.minimum(ids, V - 1)
        out[r] = pad_tokens(ids.tolist(), T)
    return out


# --------------------------------------------------------------------------
# parameters (reference initialisers)
# --------------------------------------------------------------------------
def _trunc_normal(rng, shape, std=1.0):
    """tf.truncated_normal_initializer: resample where |x| > 2*std."""
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * std).astype(np.float32)


def _glorot_uniform(rng, shape):
    lim = math.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


CNN_REF_FILTER_SIZES = (2, 3, 4, 5)       # sse_model.py:183
CNN_REF_NUM_FILTERS = (256, 128, 128, 64)  # sse_model.py:184


def init_params(mode: str, V: int, We: int, E: int, Hs: int, Ht: int, seed: int = 1234,
                target_space_size: int = 0,
                cnn_filter_sizes: Sequence[int] = CNN_REF_FILTER_SIZES,
                cnn_num_filters: Sequence[int] = CNN_REF_NUM_FILTERS) -> Dict[str, np.ndarray]:
    """Variables with the reference's names, shapes and initialisers
    (sse_model.py:159-160, 186-189, 210, 214, 222-227, 240-253, 262-270;
    BasicLSTMCell kernel = Glorot uniform (tf.get_variable default), bias = 0).
    Names follow SURVEY appendix A.6 (TF >= 1.2 checkpoint names)."""
    rng = np.random.default_rng(seed)
    p: Dict[str, np.ndarray] = {}
    p["word_embedding"] = rng.uniform(-0.25, 0.25, size=(V, We)).astype(np.float32)

    def lstm(scope, H):
        p[f"{scope}/rnn/basic_lstm_cell/kernel"] = _glorot_uniform(rng, (We + H, 4 * H))
        p[f"{scope}/rnn/basic_lstm_cell/bias"] = np.zeros(4 * H, np.float32)

    def cnn(scope):
        for k, F in zip(cnn_filter_sizes, cnn_num_filters):
            p[f"{scope}/conv-maxpool-{k}/W"] = _trunc_normal(rng, (k, We, 1, F), 0.1)
            p[f"{scope}/conv-maxpool-{k}/b"] = np.full(F, 0.1, np.float32)

    if mode == "dual-encoder":
        lstm("source_encoder", Hs)
        p["source_encoder/src_M"] = _trunc_normal(rng, (Hs, E))
        lstm("target_encoder", Ht)
        p["target_encoder/tgt_M"] = _trunc_normal(rng, (Ht, E))
    elif mode == "shared-encoder":
        lstm("shared_encoder", Hs)
        p["shared_encoder/src_M"] = _trunc_normal(rng, (Hs, E))
        p["shared_encoder/tgt_M"] = _trunc_normal(rng, (Hs, E))
    elif mode == "source-encoder-only":
        lstm("source_only_encoder", Hs)
        p["source_only_encoder/src_M"] = _trunc_normal(rng, (Hs, E))
        p["target_embedding/tgt_seq_embedding"] = rng.uniform(
            -0.25, 0.25, size=(target_space_size, E)).astype(np.float32)
    elif mode == "source_only_cnn":
        cnn("source_only_cnn")
        p["source_only_cnn/src_M"] = _trunc_normal(rng, (sum(cnn_num_filters), E))
        p["target_embedding/tgt_seq_embedding"] = rng.uniform(
            -0.25, 0.25, size=(target_space_size, E)).astype(np.float32)
    elif mode == "dual-cnn":  # BASELINE.json config 3 extension (no reference behaviour)
        cnn("source_cnn")
        p["source_cnn/src_M"] = _trunc_normal(rng, (sum(cnn_num_filters), E))
        cnn("target_cnn")
        p["target_cnn/tgt_M"] = _trunc_normal(rng, (sum(cnn_num_filters), E))
    else:
        raise ValueError("Unsupported network mode: %s" % mode)  # sse_model.py:175-177
    return p


def tower_names(mode: str, side: str) -> Tuple[str, str]:
    """(scope of the encoder variables, name of the projection matrix)."""
    s = side == "src"
    if mode == "dual-encoder":
        return ("source_encoder", "source_encoder/src_M") if s else ("target_encoder", "target_encoder/tgt_M")
    if mode == "shared-encoder":
        return ("shared_encoder", "shared_encoder/src_M" if s else "shared_encoder/tgt_M")
    if mode == "source-encoder-only":
        assert s
        return ("source_only_encoder", "source_only_encoder/src_M")
    if mode == "source_only_cnn":
        assert s
        return ("source_only_cnn", "source_only_cnn/src_M")
    if mode == "dual-cnn":
        return ("source_cnn", "source_cnn/src_M") if s else ("target_cnn", "target_cnn/tgt_M")
    raise ValueError(mode)


# --------------------------------------------------------------------------
# elementwise pieces
# --------------------------------------------------------------------------
def sigmoid(x):
    x = np.asarray(x)
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    ex = np.exp(x[~pos])
    out[~pos] = ex / (1.0 + ex)
    return out


def l2_normalize(x: np.ndarray, axis: int = -1, eps: float = 1e-12) -> np.ndarray:
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), eps)).  sse_model.py:282-283,350."""
    ss = np.sum(x * x, axis=axis, keepdims=True)
    return (x * (1.0 / np.sqrt(np.maximum(ss, x.dtype.type(eps))))).astype(x.dtype)


# --------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------
def lstm_last_state(tokens: np.ndarray, emb: np.ndarray, K: np.ndarray, b: np.ndarray,
                    dtype=np.float32, return_all: bool = False):
    """BasicLSTMCell(forget_bias=1.0) statically unrolled over all T positions
    (PADs included), zero initial state.  sse_model.py:222-224, 240-242,
    248-250, 262-264, 273-274.  Gate column blocks are i, j, f, o; the +1.0
    forget bias is added at run time (not stored in b)."""
    tokens = np.asarray(tokens)
    B, T = tokens.shape
    H = K.shape[1] // 4
    emb = emb.astype(dtype, copy=False); K = K.astype(dtype, copy=False); b = b.astype(dtype, copy=False)
    c = np.zeros((B, H), dtype); h = np.zeros((B, H), dtype)
    one = dtype(1.0)
    hs, cs, gates = [], [], []
    for t in range(T):
        x = emb[tokens[:, t]]                      # tf.nn.embedding_lookup, sse_model.py:163-164
        z = np.concatenate([x, h], axis=1) @ K + b
        i, j, f, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        si, sf, so, tj = sigmoid(i), sigmoid(f + one), sigmoid(o), np.tanh(j)
        c = c * sf + si * tj
        tc = np.tanh(c)
        h = tc * so
        if return_all:
            hs.append(h); cs.append(c); gates.append((si, tj, sf, so, tc))
    if return_all:
        return h, hs, cs, gates
    return h


def lstm_encode(tokens, emb, K, b, M, normalize: bool = True, dtype=np.float32) -> np.ndarray:
    """LSTM tower + projection (+ l2 norm).  sse_model.py:224-228 / 242-245 /
    250-254 / 264-275, 282-283."""
    h = lstm_last_state(tokens, emb, K, b, dtype)
    u = h @ M.astype(dtype)
    return l2_normalize(u) if normalize else u


def cnn_features(tokens, emb, Ws: Sequence[np.ndarray], bs: Sequence[np.ndarray], dtype=np.float32,
                 return_argmax: bool = False):
    """conv2d VALID over [T, We] with filter [k, We, 1, F] -> +b -> ReLU ->
    max over the T-k+1 positions -> concat over k.  sse_model.py:185-207."""
    tokens = np.asarray(tokens)
    B, T = tokens.shape
    X = emb.astype(dtype)[tokens]                  # [B, T, We]
    feats, argm = [], []
    for W, bb in zip(Ws, bs):
        k, We, _, F = W.shape
        Wm = W.astype(dtype).reshape(k * We, F)
        P = T - k + 1
        win = np.stack([X[:, i:i + P, :] for i in range(k)], axis=2).reshape(B, P, k * We)
        conv = win @ Wm + bb.astype(dtype)         # [B, P, F]
        act = np.maximum(conv, dtype(0))
        feats.append(act.max(axis=1))
        argm.append(act.argmax(axis=1))
    pool = np.concatenate(feats, axis=1)
    if return_argmax:
        return pool, argm
    return pool


def cnn_encode(tokens, emb, Ws, bs, M, normalize=True, dtype=np.float32):
    """CNN tower + projection.  sse_model.py:205-211."""
    u = cnn_features(tokens, emb, Ws, bs, dtype) @ M.astype(dtype)
    return l2_normalize(u) if normalize else u


def _cnn_vars(params, scope):
    ks = sorted(int(n.split("conv-maxpool-")[1].split("/")[0]) for n in params
                if n.startswith(scope + "/conv-maxpool-") and n.endswith("/W"))
    Ws = [params[f"{scope}/conv-maxpool-{k}/W"] for k in ks]
    bs = [params[f"{scope}/conv-maxpool-{k}/b"] for k in ks]
    return Ws, bs


def encode(params: Dict[str, np.ndarray], mode: str, side: str, tokens, normalize=True, dtype=np.float32):
    """model.{src,tgt}_seq_embedding / norm_*  for any mode."""
    scope, mname = tower_names(mode, side)
    emb = params["word_embedding"]
    if "cnn" in mode:
        Ws, bs = _cnn_vars(params, scope)
        return cnn_encode(tokens, emb, Ws, bs, params[mname], normalize, dtype)
    return lstm_encode(tokens, emb, params[f"{scope}/rnn/basic_lstm_cell/kernel"],
                       params[f"{scope}/rnn/basic_lstm_cell/bias"], params[mname], normalize, dtype)


# --------------------------------------------------------------------------
# similarity, loss, prediction
# --------------------------------------------------------------------------
def similarity(n_src: np.ndarray, n_tgt: np.ndarray) -> np.ndarray:
    """sse_model.py:286."""
    return n_src @ n_tgt.T


def binarylogit(n_src: np.ndarray, n_tgt: np.ndarray) -> np.ndarray:
    """sse_model.py:290."""
    return np.sum(n_src * n_tgt, axis=-1)


def pair_loss_acc(cos: np.ndarray, labels: np.ndarray) -> Tuple[float, float]:
    """weighted_cross_entropy_with_logits(64*cos, labels, pos_weight=1) mean;
    train_acc.  sse_model.py:298, 302."""
    dt = cos.dtype.type
    x = dt(64.0) * cos
    l = labels.astype(cos.dtype)
    wce = (1 - l) * x + (np.log1p(np.exp(-np.abs(x))) + np.maximum(-x, dt(0)))
    loss = wce.mean(dtype=cos.dtype)
    s = sigmoid(x)
    acc = (l * np.floor(s + dt(0.1))).mean(dtype=cos.dtype) + ((1 - l) * np.floor(dt(1.1) - s)).mean(dtype=cos.dtype)
    return float(loss), float(acc)


def top_k_tf(sim: np.ndarray, k: int, normalize_scores: bool = True):
    """tf.nn.top_k(sorted=True) (ties -> lower index first) then
    l2_normalize(scores, 1).  sse_model.py:348-350."""
    order = np.lexsort((np.broadcast_to(np.arange(sim.shape[1]), sim.shape), -sim), axis=1)[:, :k]
    scores = np.take_along_axis(sim, order, axis=1)
    if normalize_scores:
        scores = l2_normalize(scores, axis=1)
    return scores, order.astype(np.int32)


# --------------------------------------------------------------------------
# retrieval (numpy half of the reference)
# --------------------------------------------------------------------------
def get_sorted_results(scores: np.ndarray):
    """data_utils.py:263-267 (full descending sort of every row)."""
    ranked_idx = np.argsort(-scores)
    sorted_score = -np.sort(-scores, axis=1)
    return sorted_score, ranked_idx


def topk_tight_accuracy(topk: int, labels, results) -> float:
    """data_utils.py:270-286."""
    assert len(labels) == len(results)
    k = min(topk, results.shape[1])
    total = 0.0
    for i in range(results.shape[0]):
        cur = 0.0
        top = set(int(v) for v in results[i][:k])
        for lab in labels[i]:
            if lab in top:
                cur += 1.0
        total += cur / len(labels[i])
    return total / float(results.shape[0])


def topk_accuracy(topk: int, labels, results) -> float:
    """data_utils.py:289-304."""
    assert len(labels) == len(results)
    k = min(topk, results.shape[1])
    total = 0.0
    for i in range(results.shape[0]):
        top = set(int(v) for v in results[i][:k])
        for lab in labels[i]:
            if lab in top:
                total += 1.0
                break
    return total / float(results.shape[0])


def retrieve(src_enc: np.ndarray, target_encodings_f64: np.ndarray, k: int):
    """sse_evaluator.py:110-111 / sse_demo.py:126-127: float32 [Q,E] x float64
    [E,N] -> float64 distances, full argsort, first k."""
    d = np.dot(src_enc, target_encodings_f64.T)
    s, idx = get_sorted_results(d)
    return s[:, :k], idx[:, :k]


def evaluator_eval(src_enc_batches: Sequence[np.ndarray], target_encodings_f64: np.ndarray,
                   eval_labels: Sequence[Sequence[int]], top_n=(1, 3, 10), batch_size: int = 600):
    """sse_evaluator.py:95-114 with the source encodings given (the encoder is
    a separate oracle function)."""
    acc = []
    for n in top_n:
        batchacc = []
        for bi, enc in enumerate(src_enc_batches):
            d = np.dot(enc, target_encodings_f64.T)
            _, ranked = get_sorted_results(d)
            batchacc.append(topk_tight_accuracy(n, eval_labels[bi * batch_size:(bi + 1) * batch_size], ranked))
        acc.append(float(np.mean(batchacc)))
    return acc


# --------------------------------------------------------------------------
# index file format  (sse_index.py:93-95 writer, sse_evaluator.py:80-92 reader)
# --------------------------------------------------------------------------
def format_index_row(tgt_id: str, text: str, enc_row: np.ndarray) -> str:
    return tgt_id + "\t" + text + "\t" + ",".join([str(n) for n in enc_row]) + "\n"


def parse_index_lines(lines: Sequence[str]):
    ids, texts, encs = [], [], []
    for line in lines:
        info = line.strip().split("\t")
        if len(info) != 3:
            continue
        ids.append(info[0]); texts.append(info[1])
        encs.append([float(f) for f in info[2].strip().split(",")])
    return ids, texts, np.array(encs)


# --------------------------------------------------------------------------
# training step (forward + BPTT + clip + Adagrad), numpy, manual backward
# --------------------------------------------------------------------------
def _lstm_backward(tokens, emb, K, dh_last, hs, cs, gates, dtype):
    """BPTT through lstm_last_state.  Returns dK, db, and the per-(row,t)
    gradient w.r.t. the gathered embedding rows [B,T,We]."""
    B, T = tokens.shape
    H = K.shape[1] // 4
    We = emb.shape[1]
    dK = np.zeros_like(K, dtype=dtype); db = np.zeros(4 * H, dtype)
    dX = np.zeros((B, T, We), dtype)
    dh = dh_last.astype(dtype); dc = np.zeros((B, H), dtype)
    Kd = K.astype(dtype)
    for t in range(T - 1, -1, -1):
        si, tj, sf, so, tc = gates[t]
        c_prev = cs[t - 1] if t > 0 else np.zeros((B, H), dtype)
        h_prev = hs[t - 1] if t > 0 else np.zeros((B, H), dtype)
        do = dh * tc
        dc = dc + dh * so * (1 - tc * tc)
        di = dc * tj; dj = dc * si; df = dc * c_prev
        dz = np.concatenate([di * si * (1 - si), dj * (1 - tj * tj), df * sf * (1 - sf), do * so * (1 - so)], axis=1)
        x = emb.astype(dtype)[tokens[:, t]]
        xh = np.concatenate([x, h_prev], axis=1)
        dK += xh.T @ dz
        db += dz.sum(axis=0)
        dxh = dz @ Kd.T
        dX[:, t, :] = dxh[:, :We]
        dh = dxh[:, We:]
        dc = dc * sf
    return dK, db, dX


def _l2norm_backward(u, g):
    """d/du of n = u * rsqrt(max(sum u^2, eps)) given dL/dn = g (eps branch inactive)."""
    ss = np.sum(u * u, axis=-1, keepdims=True)
    inv = 1.0 / np.sqrt(np.maximum(ss, 1e-12))
    n = u * inv
    return (g - n * np.sum(g * n, axis=-1, keepdims=True)) * inv


@dataclass
class TrainState:
    params: Dict[str, np.ndarray]
    accum: Dict[str, np.ndarray] = field(default_factory=dict)
    learning_rate: float = 0.9
    global_step: int = 0

    def __post_init__(self):
        if not self.accum:  # AdagradOptimizer initial_accumulator_value = 0.1
            self.accum = {k: np.full_like(v, 0.1) for k, v in self.params.items()}


def train_step(state: TrainState, mode: str, src, tgt, labels, dtype=np.float32, max_grad_norm: float = 5.0,
               return_grads: bool = False):
    """One sess.run([train, loss, train_acc]).  sse_model.py:279-302 (loss),
    355-364 (tf.gradients -> clip_by_global_norm(5.0) -> Adagrad).  LSTM modes
    only (dual-encoder / shared-encoder), which are the modes the reference can
    actually train at HEAD.

    Embedding gradient semantics: the gradient of the shared
    word_embedding is an IndexedSlices = concat(source-lookup slices,
    target-lookup slices); the global norm is taken over the UN-merged slice
    values; Adagrad's sparse apply then sums duplicate rows before updating
    only the touched rows."""
    p = state.params
    src = np.asarray(src, np.int32); tgt = np.asarray(tgt, np.int32)
    labels = np.asarray(labels, dtype)
    B = src.shape[0]
    emb = p["word_embedding"]
    s_scope, s_M = tower_names(mode, "src")
    t_scope, t_M = tower_names(mode, "tgt")
    Ks, bs_ = p[f"{s_scope}/rnn/basic_lstm_cell/kernel"], p[f"{s_scope}/rnn/basic_lstm_cell/bias"]
    Kt, bt_ = p[f"{t_scope}/rnn/basic_lstm_cell/kernel"], p[f"{t_scope}/rnn/basic_lstm_cell/bias"]
    hS, hsS, csS, gS = lstm_last_state(src, emb, Ks, bs_, dtype, True)
    hT, hsT, csT, gT = lstm_last_state(tgt, emb, Kt, bt_, dtype, True)
    uS = hS @ p[s_M].astype(dtype); uT = hT @ p[t_M].astype(dtype)
    nS = l2_normalize(uS); nT = l2_normalize(uT)
    cos = binarylogit(nS, nT)
    loss, acc = pair_loss_acc(cos, labels)
    # d loss / d cos : x = 64 cos ; d wce/dx = sigmoid(x) - l  (pos_weight = 1)
    x = dtype(64.0) * cos
    dcos = (sigmoid(x) - labels) * dtype(64.0) / dtype(B)
    gnS = dcos[:, None] * nT; gnT = dcos[:, None] * nS
    duS = _l2norm_backward(uS, gnS).astype(dtype); duT = _l2norm_backward(uT, gnT).astype(dtype)
    grads: Dict[str, np.ndarray] = {}
    grads[s_M] = hS.T @ duS
    grads[t_M] = hT.T @ duT
    dhS = duS @ p[s_M].astype(dtype).T; dhT = duT @ p[t_M].astype(dtype).T
    dKs, dbs, dXs = _lstm_backward(src, emb, Ks, dhS, hsS, csS, gS, dtype)
    dKt, dbt, dXt = _lstm_backward(tgt, emb, Kt, dhT, hsT, csT, gT, dtype)
    kn_s, bn_s = f"{s_scope}/rnn/basic_lstm_cell/kernel", f"{s_scope}/rnn/basic_lstm_cell/bias"
    kn_t, bn_t = f"{t_scope}/rnn/basic_lstm_cell/kernel", f"{t_scope}/rnn/basic_lstm_cell/bias"
    if s_scope == t_scope:  # shared encoder: both towers' grads add into one kernel
        grads[kn_s] = dKs + dKt; grads[bn_s] = dbs + dbt
    else:
        grads[kn_s] = dKs; grads[bn_s] = dbs; grads[kn_t] = dKt; grads[bn_t] = dbt
    # IndexedSlices for the embedding: un-merged values
    slice_idx = np.concatenate([src.reshape(-1), tgt.reshape(-1)])
    slice_val = np.concatenate([dXs.reshape(-1, emb.shape[1]), dXt.reshape(-1, emb.shape[1])])
    sq = sum(float(np.sum(g.astype(np.float64) ** 2)) for g in grads.values())
    sq += float(np.sum(slice_val.astype(np.float64) ** 2))
    gnorm = math.sqrt(sq)
    scale = dtype(max_grad_norm / max(gnorm, max_grad_norm))
    lr = dtype(state.learning_rate)
    for name, g in grads.items():
        g = (g * scale).astype(dtype)
        state.accum[name] = state.accum[name] + g * g
        p[name] = (p[name] - lr * g / np.sqrt(state.accum[name])).astype(np.float32)
    # sparse Adagrad: sum duplicates, update touched rows only
    uniq, inv = np.unique(slice_idx, return_inverse=True)
    summed = np.zeros((len(uniq), emb.shape[1]), dtype)
    np.add.at(summed, inv, (slice_val * scale).astype(dtype))
    acc_rows = state.accum["word_embedding"][uniq] + summed * summed
    state.accum["word_embedding"][uniq] = acc_rows
    emb_new = p["word_embedding"].copy()
    emb_new[uniq] = emb[uniq] - lr * summed / np.sqrt(acc_rows)
    p["word_embedding"] = emb_new.astype(np.float32)
    state.global_step += 1
    if return_grads:
        dense_emb = np.zeros_like(emb, dtype=dtype)
        np.add.at(dense_emb, slice_idx, slice_val)
        grads = dict(grads); grads["word_embedding"] = dense_emb
        return loss, acc, gnorm, grads
    return loss, acc, gnorm


def lr_decay(state: TrainState, factor: float):
    """learning_rate_decay_op.  sse_model.py:123-124."""
    state.learning_rate = float(max(np.float32(state.learning_rate) * np.float32(factor), np.float32(1e-3)))
    return state.learning_rate
