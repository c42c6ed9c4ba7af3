# coding=utf-8
"""Drop-in `sse_model` for the B200-native SSE hot path.

Keeps the object protocol the reference's callers use (reference sse_model.py:90-440 and
the call sites listed in SURVEY 8b): ``SSEModel(modelParams)``, the fetchable attributes
(``src_seq_embedding``, ``norm_src_seq_embedding``, ``tgt_seq_embedding``,
``norm_tgt_seq_embedding``, ``similarity``, ``binarylogit``, ``loss``, ``train_acc``,
``train``, ``predicted_tgts_score``, ``predicted_labels``, ``learning_rate``,
``learning_rate_decay_op``, ``global_step``, ``saver``), the four ``get_*_feed_dict``
builders, ``save`` / ``load`` / ``set_top_n`` / ``set_forward_only`` / ``add_summaries``.

There is no TensorFlow graph behind it: every fetch is executed by hand-written CUDA
kernels in libsse_b200.so through the C ABI (sse_ffi.py).  ``Session.run(fetches,
feed_dict)`` is a thin dispatcher so that code written against ``tf.Session`` (index
builder, evaluator, demo, web routes, training loop) runs unchanged in shape.
"""
from __future__ import annotations

import os
import threading
from typing import Dict, List, Optional

import numpy as np

import sse_ffi

PAD_ID, EOS_ID = 0, 1

_MODEL_PARAM_KEYS = ("forward_only", "network_mode", "predict_nbest", "max_seq_length", "vocab_size", "embedding_size",
                     "encoding_size", "src_cell_size", "tgt_cell_size", "learning_rate", "learning_rate_decay_factor",
                     "targetSpaceSize")


class _Tensor(object):
    """A fetchable name (stands in for a tf.Tensor / tf.Operation / placeholder)."""

    def __init__(self, model, name):
        self.model = model
        self.name = name

    def eval(self, session=None, feed_dict=None):
        return self.model._run_one(self, feed_dict or {})

    def __repr__(self):
        return "<sse_b200 tensor %s>" % self.name

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


class Saver(object):
    """Replacement for tf.train.Saver(tf.global_variables(), max_to_keep=20) (sse_model.py:138).
    Checkpoints are ``<path>[-<step>].npz`` holding every variable, the Adagrad slots,
    ``learning_rate`` and ``global_step`` under their TF names, plus the TF-style
    ``checkpoint`` state file in the same directory (read by get_checkpoint_state)."""

    def __init__(self, model, max_to_keep=20):
        self.model = model
        self.max_to_keep = max_to_keep
        self._kept: List[str] = []

    def save(self, session, path, global_step=None):
        prefix = path if global_step is None else "%s-%d" % (path, int(global_step))
        arrays = {}
        for name, _shape in self.model.handle.param_names():
            arrays[name] = self.model.handle.get_param(name)
        lr, gs = self.model.handle.scalars()
        arrays["learning_rate"] = np.float32(lr)
        arrays["global_step"] = np.int64(gs)
        # SSE_CHECKPOINT_FORMAT: "npz" (default), "tf" = TensorFlow V2 tensor bundle (<prefix>.index + .data-00000-of-00001,
        # the files the reference's tf.train.Saver writes, sse_train.py:205,212,232), "both"
        fmt = os.environ.get("SSE_CHECKPOINT_FORMAT", "npz")
        if fmt in ("tf", "both"):
            import tf_bundle
            tf_bundle.write_bundle(prefix, arrays)
        if fmt != "tf":
            np.savez(prefix + ".npz", **arrays)
        d = os.path.dirname(prefix) or "."
        # tf.train.Saver keeps ONE entry per checkpoint name: re-saving `...-BestEver` moves it to the newest
        # position instead of filling the max_to_keep window with duplicates (which would later delete a file
        # that newer entries still name)
        if prefix in self._kept:
            self._kept.remove(prefix)
        self._kept.append(prefix)
        while len(self._kept) > self.max_to_keep:
            old = self._kept.pop(0)
            if old not in self._kept and old != prefix:
                for ext in (".npz", ".index", ".data-00000-of-00001"):
                    if os.path.exists(old + ext):
                        os.remove(old + ext)
        with open(os.path.join(d, "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s"\n' % os.path.basename(prefix))
            for k in self._kept:
                f.write('all_model_checkpoint_paths: "%s"\n' % os.path.basename(k))
        return prefix

    def restore(self, session, path):
        if not path.endswith(".npz") and not os.path.exists(path + ".npz") and os.path.exists(path + ".index"):
            return self._restore_tf_bundle(path)
        fn = path if path.endswith(".npz") else path + ".npz"
        with np.load(fn) as z:
            names = dict(self.model.handle.param_names())
            for name in z.files:
                if name in ("learning_rate", "global_step"):
                    continue
                if name not in names:
                    raise sse_ffi.SseError("checkpoint variable %s not in model" % name)
                self.model.handle.set_param(name, z[name])
            self.model.handle.set_scalars(float(z["learning_rate"]), int(z["global_step"]))
        self.model._initialized = True


    def _restore_tf_bundle(self, prefix):
        """A checkpoint written by the reference's tf.train.Saver (V2 tensor bundle: <prefix>.index + .data-*), read by
        tf_bundle.py (format-level reader, unpinned against the real library).  Variable names are TensorFlow's
        (SURVEY A.6); TF <= 1.1 LSTM names (`.../basic_lstm_cell/weights|biases`) map to `kernel|bias`.  Variables the
        model does not have (summaries, beta powers, ...) are ignored; every model variable must be present."""
        import tf_bundle
        tensors = tf_bundle.read_bundle(prefix)
        names = dict(self.model.handle.param_names())
        found = set()
        for name, arr in tensors.items():
            tgt = name.replace("basic_lstm_cell/weights", "basic_lstm_cell/kernel").replace("basic_lstm_cell/biases", "basic_lstm_cell/bias")
            if tgt in names:
                self.model.handle.set_param(tgt, np.asarray(arr, dtype=np.float32))
                found.add(tgt)
        missing = [n for n in names if n not in found and not n.endswith("/Adagrad")]
        if missing:
            raise sse_ffi.SseError("TF checkpoint %s lacks model variables: %s" % (prefix, ", ".join(sorted(missing)[:6])))
        lr, gs = self.model.handle.scalars()
        if "learning_rate" in tensors:
            lr = float(np.asarray(tensors["learning_rate"]).reshape(-1)[0])
        if "global_step" in tensors:
            gs = int(np.asarray(tensors["global_step"]).reshape(-1)[0])
        self.model.handle.set_scalars(lr, gs)
        self.model._initialized = True


class CheckpointState(object):
    def __init__(self, model_checkpoint_path):
        self.model_checkpoint_path = model_checkpoint_path


def get_checkpoint_state(model_dir):
    """tf.train.get_checkpoint_state (used at sse_train.py:110, sse_index.py:117, sse_demo.py:98)."""
    fn = os.path.join(model_dir, "checkpoint")
    if not os.path.exists(fn):
        return None
    for line in open(fn):
        if line.startswith("model_checkpoint_path:"):
            name = line.split(":", 1)[1].strip().strip('"')
            p = name if os.path.isabs(name) else os.path.join(model_dir, name)
            if os.path.exists(p + ".npz") or os.path.exists(p + ".index"):
                return CheckpointState(p)
    return None


class _InitOp(object):
    pass


def global_variables_initializer():
    """tf.global_variables_initializer() (sse_train.py:120): run it with Session.run."""
    return _InitOp()


class Session(object):
    """Minimal stand-in for tf.Session: ``run(fetches, feed_dict)`` over SSEModel tensors.
    Thread-safe like tf.Session.run (one lock per model handle)."""

    def __init__(self, config=None, seed=None):
        self.seed = seed
        self.graph = None
        self._models = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        feed_dict = feed_dict or {}
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        model = None
        for f in list(flist) + list(feed_dict.keys()):
            if isinstance(f, _Tensor):
                model = f.model
                break
        if model is None:
            if any(isinstance(f, _InitOp) for f in flist):
                for m in SSEModel._live:
                    m.initialize(self.seed)
                return None if single else [None] * len(flist)
            raise ValueError("Session.run: nothing to run")
        out = model._run(flist, feed_dict)
        return out[0] if single else out


class SSEModel(object):
    _live: List["SSEModel"] = []

    def __init__(self, modelParams, device=None, precision=None):
        """Create the Sequence Semantic Embedding model (reference sse_model.py:94-138).
        modelParams values may be strings (as read back from modelConfig.param)."""
        self._post_train_ops = []
        self.name = "SSEmodel"
        self.forward_only = bool(modelParams["forward_only"])        # same quirk as the reference (:113)
        self.network_mode = modelParams["network_mode"]
        self.TOP_N = int(modelParams["predict_nbest"])
        self.MAX_SEQ_LENGTH = int(modelParams["max_seq_length"])
        self.max_gradient_norm = 5.0
        self.vocab_size = int(modelParams["vocab_size"])
        self.word_embed_size = int(modelParams["embedding_size"])
        self.seq_embed_size = int(modelParams["encoding_size"])
        self.src_cell_size = int(modelParams["src_cell_size"])
        self.tgt_cell_size = int(modelParams["tgt_cell_size"])
        self.targetSpaceSize = int(modelParams["targetSpaceSize"])
        if self.network_mode not in sse_ffi.MODE_IDS:
            # the reference prints and exit(-1)s (sse_model.py:175-177); raise instead of killing the host
            raise ValueError("Error!! Unsupported network mode: %s. Please specify on: source-encoder-only, "
                             "dual-encoder or shared-encoder." % self.network_mode)
        if device is None:
            device = int(os.environ.get("SSE_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        if precision is None:
            precision = int(os.environ.get("SSE_PRECISION", str(sse_ffi.PRECISION_TC)))
        cnn_k = modelParams.get("cnn_filter_sizes", ())
        cnn_f = modelParams.get("cnn_num_filters", ())
        if isinstance(cnn_k, str):
            cnn_k = [int(x) for x in cnn_k.split(",") if x]
            cnn_f = [int(x) for x in str(cnn_f).split(",") if x]
        self.handle = sse_ffi.Handle(
            self.network_mode, self.vocab_size, self.word_embed_size, self.seq_embed_size, self.src_cell_size,
            self.tgt_cell_size, self.MAX_SEQ_LENGTH, predict_nbest=self.TOP_N, target_space_size=self.targetSpaceSize,
            learning_rate=float(modelParams["learning_rate"]),
            learning_rate_decay_factor=float(modelParams["learning_rate_decay_factor"]), device=device,
            precision=precision, cnn_filter_sizes=cnn_k, cnn_num_filters=cnn_f)
        self._lock = threading.Lock()
        self._initialized = False
        # placeholders (sse_model.py:153-155)
        self._src_input_data = _Tensor(self, "source_sequence")
        self._tgt_input_data = _Tensor(self, "target_sequence")
        self._labels = _Tensor(self, "targetSpace_labels")
        # fetchable tensors / ops
        for n in ("src_seq_embedding", "tgt_seq_embedding", "norm_src_seq_embedding", "norm_tgt_seq_embedding",
                  "similarity", "binarylogit", "loss", "train_acc", "train", "predicted_tgts_score",
                  "predicted_labels", "learning_rate", "learning_rate_decay_op", "global_step"):
            setattr(self, n, _Tensor(self, n))
        self._summary = _Tensor(self, "summary")
        self.saver = Saver(self, max_to_keep=20)
        SSEModel._live.append(self)

    # ------------------------------------------------------------------ variables
    def initialize(self, seed=None):
        """The reference initialisers (sse_model.py:160,188-189,210,214,227; BasicLSTMCell: Glorot
        uniform kernel, zero bias) -- what sess.run(tf.global_variables_initializer()) does."""
        rng = np.random.default_rng(seed)
        for name, shape in self.handle.param_names():
            if name.endswith("/Adagrad"):
                v = np.full(shape, 0.1, np.float32)
            elif name == "word_embedding" or name.endswith("tgt_seq_embedding"):
                v = rng.uniform(-0.25, 0.25, size=shape).astype(np.float32)
            elif name.endswith("/kernel"):
                lim = np.sqrt(6.0 / (shape[0] + shape[1]))
                v = rng.uniform(-lim, lim, size=shape).astype(np.float32)
            elif name.endswith("/bias"):
                v = np.zeros(shape, np.float32)
            elif name.endswith("_M") or name.endswith("/W"):
                std = 0.1 if name.endswith("/W") else 1.0
                x = rng.standard_normal(shape)
                bad = np.abs(x) > 2
                while bad.any():
                    x[bad] = rng.standard_normal(int(bad.sum()))
                    bad = np.abs(x) > 2
                v = (x * std).astype(np.float32)
            elif name.endswith("/b"):
                v = np.full(shape, 0.1, np.float32)
            else:
                raise sse_ffi.SseError("no initialiser for %s" % name)
            self.handle.set_param(name, v)
        self._initialized = True

    def set_top_n(self, top_n):
        self.TOP_N = top_n

    def set_forward_only(self, forward_only=True):
        self.forward_only = forward_only

    def save(self, session, path, global_step=None):
        """ Saves variables to given path """
        return self.saver.save(session, path, global_step)

    def load(self, session, path):
        """ Restores variables from given path """
        self.saver.restore(session, path)

    def add_summaries(self):
        return self._summary

    # ------------------------------------------------------------------ feed dicts (sse_model.py:401-440)
    def get_predict_feed_dict(self, srcSeqs, tgtSeqs):
        d = {}
        d[self._src_input_data] = np.array(srcSeqs, dtype=np.int32)
        d[self._tgt_input_data] = np.array(tgtSeqs, dtype=np.int32)
        return d

    def get_train_feed_dict(self, srcSeqs, tgtSeqs, labels):
        d = {}
        d[self._src_input_data] = np.array(srcSeqs, dtype=np.int32)
        d[self._labels] = np.array(labels, dtype=np.float32)
        d[self._tgt_input_data] = np.array(tgtSeqs, dtype=np.int32)
        return d

    def get_source_encoding_feed_dict(self, srcSeqs):
        d = {}
        d[self._src_input_data] = np.array(srcSeqs, dtype=np.int32)
        return d

    def get_target_encoding_feed_dict(self, tgtSeqs):
        d = {}
        d[self._tgt_input_data] = np.array(tgtSeqs, dtype=np.int32)
        return d

    # ------------------------------------------------------------------ execution
    def _feed(self, feed_dict, key, what):
        if key not in feed_dict:
            raise ValueError("You must feed a value for placeholder tensor '%s'" % what)
        a = np.ascontiguousarray(feed_dict[key])
        return a

    def _tokens(self, feed_dict, key, what):
        a = self._feed(feed_dict, key, what).astype(np.int32, copy=False)
        if a.ndim != 2 or a.shape[1] != self.MAX_SEQ_LENGTH:
            raise ValueError("%s must have shape [None, %d], got %s" % (what, self.MAX_SEQ_LENGTH, a.shape))
        return a

    def _tgt_is_table(self):
        return self.network_mode in ("source-encoder-only", "source_only_cnn")

    def _encode_tgt(self, feed_dict, normalize):
        if self._tgt_is_table():
            t = self.handle.get_param("target_embedding/tgt_seq_embedding")
            if normalize:
                t = t / np.sqrt(np.maximum((t * t).sum(-1, keepdims=True), 1e-12))
            return t.astype(np.float32)
        return self.handle.encode_host(sse_ffi.SIDE_TGT, self._tokens(feed_dict, self._tgt_input_data, "target_sequence"),
                                       normalize)

    def _encode_src(self, feed_dict, normalize):
        return self.handle.encode_host(sse_ffi.SIDE_SRC, self._tokens(feed_dict, self._src_input_data, "source_sequence"),
                                       normalize)

    def _pair_cos(self, feed_dict):
        import torch
        src = torch.from_numpy(self._tokens(feed_dict, self._src_input_data, "source_sequence")).cuda(self.handle.cfg.device)
        tgt = torch.from_numpy(self._tokens(feed_dict, self._tgt_input_data, "target_sequence")).cuda(self.handle.cfg.device)
        cos = torch.empty(src.shape[0], device=src.device)
        self.handle.pair_score(src, tgt, src.shape[0], cos)
        return cos.cpu().numpy()

    def _run_one(self, t, feed_dict):
        return self._run([t], feed_dict)[0]

    def _run(self, fetches, feed_dict):
        if not self._initialized:
            raise sse_ffi.SseError("Attempting to use uninitialized variables: run global_variables_initializer() "
                                   "or saver.restore() first")
        with self._lock:
            names = [f.name if isinstance(f, _Tensor) else None for f in fetches]
            cache: Dict[str, object] = {}

            def get(name):
                if name in cache:
                    return cache[name]
                v = self._compute(name, feed_dict, get, cache)
                cache[name] = v
                return v

            if "train" in names:                   # one fused step also yields loss / train_acc
                get("train")
            return [get(n) if n is not None else None for n in names]

    def _compute(self, name, feed_dict, get, cache):
        h = self.handle
        if name == "src_seq_embedding":
            return self._encode_src(feed_dict, False)
        if name == "norm_src_seq_embedding":
            return self._encode_src(feed_dict, True)
        if name == "tgt_seq_embedding":
            return self._encode_tgt(feed_dict, False)
        if name == "norm_tgt_seq_embedding":
            return self._encode_tgt(feed_dict, True)
        if name == "similarity":
            # [B_src, B_tgt] matrix of the batch (sse_model.py:286); no caller fetches it -- host matmul of
            # the two CUDA-encoded batches is enough for this debugging tensor.
            return get("norm_src_seq_embedding") @ get("norm_tgt_seq_embedding").T
        if name == "binarylogit":
            return self._pair_cos(feed_dict)
        if name in ("loss", "train_acc"):
            cos = get("binarylogit").astype(np.float32)
            lab = self._feed(feed_dict, self._labels, "targetSpace_labels").astype(np.float32)
            x = np.float32(64.0) * cos
            sig = 1.0 / (1.0 + np.exp(-x))
            loss = np.mean((1 - lab) * x + np.log1p(np.exp(-np.abs(x))) + np.maximum(-x, 0))
            acc = np.mean(lab * np.floor(sig + 0.1)) + np.mean((1 - lab) * np.floor(1.1 - sig))
            cache["loss"], cache["train_acc"] = np.float32(loss), np.float32(acc)
            return cache[name]
        if name == "train":
            loss, acc, _gn = h.train_step(self._tokens(feed_dict, self._src_input_data, "source_sequence"),
                                          self._tokens(feed_dict, self._tgt_input_data, "target_sequence"),
                                          self._feed(feed_dict, self._labels, "targetSpace_labels").astype(np.float32))
            cache["loss"], cache["train_acc"] = np.float32(loss), np.float32(acc)
            return None
        if name in ("predicted_tgts_score", "predicted_labels"):
            # tf.nn.top_k(similarity, TOP_N, sorted=True) then l2_normalize(scores, 1) (sse_model.py:348-350)
            import torch
            dev = torch.device("cuda", h.cfg.device)
            tgt = torch.from_numpy(get("norm_tgt_seq_embedding")).to(dev)
            q = torch.from_numpy(get("norm_src_seq_embedding")).to(dev)
            k = self.TOP_N
            if tgt.shape[0] < k:
                raise ValueError("input must have at least k columns")          # TF's error for top_k
            s = torch.empty(q.shape[0], k, device=dev)
            i = torch.empty(q.shape[0], k, device=dev, dtype=torch.int32)
            # scored against the batch's own targets without registering them: the index an Evaluator / DemoSession
            # keeps resident on this handle is left alone
            h.topk_batch(q, q.shape[0], tgt.contiguous(), tgt.shape[0], k, s, i, normalize_scores=True)
            torch.cuda.synchronize(dev)
            cache["predicted_tgts_score"], cache["predicted_labels"] = s.cpu().numpy(), i.cpu().numpy()
            return cache[name]
        if name == "learning_rate":
            return np.float32(h.scalars()[0])
        if name == "global_step":
            return h.scalars()[1]
        if name == "learning_rate_decay_op":
            h.lr_decay()
            return np.float32(h.scalars()[0])
        if name == "summary":
            return None
        raise ValueError("cannot fetch %s" % name)
