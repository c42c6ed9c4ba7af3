"""ctypes binding of libsse_b200.so (the C ABI declared in include/sse_b200.h).

This is the only place Python touches the native library.  There is no CPU
fallback: if the shared library is missing or no GPU is visible, the calls
raise.  Device memory is passed as raw pointers (``tensor.data_ptr()``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsse_b200.so")

SSE_OK = 0
MODE_IDS = {
    "dual-encoder": 0,
    "shared-encoder": 1,
    "source-encoder-only": 2,
    "source_only_cnn": 3,
    "dual-cnn": 4,
}
PRECISION_FP32, PRECISION_TC = 0, 1
SIDE_SRC, SIDE_TGT = 0, 1
MAX_CNN = 8


class SseConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("network_mode", C.c_int32),
        ("vocab_size", C.c_int32),
        ("embedding_size", C.c_int32),
        ("encoding_size", C.c_int32),
        ("src_cell_size", C.c_int32),
        ("tgt_cell_size", C.c_int32),
        ("max_seq_length", C.c_int32),
        ("predict_nbest", C.c_int32),
        ("forward_only", C.c_int32),
        ("target_space_size", C.c_int64),
        ("learning_rate", C.c_float),
        ("learning_rate_decay_factor", C.c_float),
        ("device", C.c_int32),
        ("precision", C.c_int32),
        ("n_cnn_filters", C.c_int32),
        ("cnn_filter_sizes", C.c_int32 * MAX_CNN),
        ("cnn_num_filters", C.c_int32 * MAX_CNN),
        ("reserved", C.c_int32 * 8),
    ]


class SseError(RuntimeError):
    pass


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)   -- must list every symbol include/sse_b200.h declares
    "sse_create": (C.c_int, [C.POINTER(SseConfig), C.POINTER(_P)]),
    "sse_destroy": (C.c_int, [_P]),
    "sse_last_error": (C.c_char_p, []),
    "sse_version": (C.c_char_p, []),
    "sse_set_param": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "sse_get_param": (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    "sse_param_count": (C.c_int, [_P]),
    "sse_param_info": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "sse_encode": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int, _P]),
    "sse_encode_host": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int]),
    "sse_index_set": (C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    "sse_index_build": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int]),
    "sse_index_get": (C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    "sse_search": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "sse_search_packed": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "sse_merge_packed": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "sse_search_stats": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sse_merge_topk": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "sse_query_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "sse_l2_normalize_rows": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "sse_topk_batch": (C.c_int, [_P, _P, C.c_int, _P, C.c_int64, C.c_int, C.c_int, _P, _P, _P]),
    "sse_token_errors": (C.c_int, [_P, C.POINTER(C.c_int64), _P]),
    "sse_pair_score": (C.c_int, [_P, _P, _P, C.c_int, _P, _P]),
    "sse_train_step": (C.c_int, [_P, _P, _P, _P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), _P]),
    "sse_train_grads": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), _P]),
    "sse_grad_arena": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "sse_train_apply": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), _P]),
    "sse_sampler_set": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, C.c_int64]),
    "sse_sampler_batch": (C.c_int, [_P, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, _P, _P, _P, C.POINTER(C.c_int), _P]),
    "sse_lr_decay": (C.c_int, [_P]),
    "sse_get_scalars": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "sse_set_scalars": (C.c_int, [_P, C.c_float, C.c_int64]),
    "sse_debug_gemm_tc": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _P, _P]),
    "sse_launch_count": (C.c_int64, [_P]),
    "sse_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "sse_tsv_last_error": (C.c_char_p, []),
    "sse_crc32c": (C.c_uint32, [_P, C.c_size_t, C.c_uint32]),
    "sse_tsv_format_f32": (C.c_int, [_P, C.c_int64, _P, C.c_size_t, _P]),
    "sse_tsv_write_index": (C.c_int, [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _P, C.c_int64, C.c_int, C.c_int, C.c_int]),
    "sse_tsv_parse_index": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.c_int64, _P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]),
    "sse_tok_last_error": (C.c_char_p, []),
    "sse_tok_create": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.POINTER(_P)]),
    "sse_tok_destroy": (C.c_int, [_P]),
    "sse_tok_vocab_size": (C.c_int, [_P]),
    "sse_tok_encode_batch": (C.c_int, [_P, C.POINTER(C.c_char_p), C.c_int64, C.c_int, _P, _P, C.c_int]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load_library(path: Optional[str] = None):
    """Load libsse_b200.so and attach signatures.  Raises if it is missing:
    the product path has no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise SseError(
            "libsse_b200.so not found at %s -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "There is no CPU fallback for the SSE hot path." % p)
    lib = C.CDLL(p)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _ptr(x) -> int:
    """Raw address of a numpy array / torch tensor / int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


def _stream_ptr(stream) -> Optional[int]:
    if stream is None:
        return None
    if isinstance(stream, int):
        return stream or None
    return stream.cuda_stream or None


class Handle:
    """One sse_handle (weights + optimizer state + resident index on one GPU)."""

    def __init__(self, mode: str, vocab_size: int, embedding_size: int, encoding_size: int, src_cell_size: int,
                 tgt_cell_size: int, max_seq_length: int, predict_nbest: int = 10, target_space_size: int = 0,
                 learning_rate: float = 0.9, learning_rate_decay_factor: float = 0.99, device: int = 0,
                 precision: int = PRECISION_TC, cnn_filter_sizes: Sequence[int] = (), cnn_num_filters: Sequence[int] = ()):
        self.lib = load_library()
        if mode not in MODE_IDS:
            raise SseError("Unsupported network mode: %s" % mode)
        cfg = SseConfig()
        cfg.struct_size = C.sizeof(SseConfig)
        cfg.network_mode = MODE_IDS[mode]
        cfg.vocab_size = vocab_size
        cfg.embedding_size = embedding_size
        cfg.encoding_size = encoding_size
        cfg.src_cell_size = src_cell_size
        cfg.tgt_cell_size = tgt_cell_size
        cfg.max_seq_length = max_seq_length
        cfg.predict_nbest = predict_nbest
        cfg.target_space_size = target_space_size
        cfg.learning_rate = learning_rate
        cfg.learning_rate_decay_factor = learning_rate_decay_factor
        cfg.device = device
        cfg.precision = precision
        cfg.n_cnn_filters = len(cnn_filter_sizes)
        for i, (k, f) in enumerate(zip(cnn_filter_sizes, cnn_num_filters)):
            cfg.cnn_filter_sizes[i] = k
            cfg.cnn_num_filters[i] = f
        self.cfg = cfg
        self.mode = mode
        self.T = max_seq_length
        self.E = encoding_size
        self._h = _P()
        self._check(self.lib.sse_create(C.byref(cfg), C.byref(self._h)))

    # -- plumbing
    def _check(self, rc: int):
        if rc != SSE_OK:
            msg = self.lib.sse_last_error()
            raise SseError("libsse_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))

    def close(self):
        if self._h:
            self.lib.sse_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- variables
    def param_names(self) -> List[Tuple[str, Tuple[int, ...]]]:
        out = []
        n = self.lib.sse_param_count(self._h)
        buf = C.create_string_buffer(128)
        shape = (C.c_int64 * 4)()
        nd = C.c_int()
        for i in range(n):
            self._check(self.lib.sse_param_info(self._h, i, buf, 128, shape, C.byref(nd)))
            out.append((buf.value.decode(), tuple(int(shape[d]) for d in range(nd.value))))
        return out

    def set_param(self, name: str, value):
        """value: numpy fp32 array or a torch tensor (host or device)."""
        if isinstance(value, np.ndarray):
            value = np.ascontiguousarray(value, dtype=np.float32)
            shape = value.shape
        else:
            value = value.contiguous().float()
            shape = tuple(value.shape)
        sh = (C.c_int64 * len(shape))(*shape)
        self._check(self.lib.sse_set_param(self._h, name.encode(), _ptr(value), sh, len(shape)))

    def get_param(self, name: str) -> np.ndarray:
        shapes = dict(self.param_names())
        if name not in shapes:
            raise SseError("unknown variable %s" % name)
        out = np.empty(shapes[name], dtype=np.float32)
        self._check(self.lib.sse_get_param(self._h, name.encode(), out.ctypes.data, out.nbytes))
        return out

    def set_params(self, params: Dict[str, np.ndarray]):
        for k, v in params.items():
            self.set_param(k, v)

    # -- encoders
    def encode(self, side: int, tokens_dev, B: int, out_dev, normalize: bool = True, stream=None):
        self._check(self.lib.sse_encode(self._h, side, _ptr(tokens_dev), B, _ptr(out_dev), int(normalize),
                                        _stream_ptr(stream)))

    def encode_host(self, side: int, tokens: np.ndarray, normalize: bool = True) -> np.ndarray:
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        B = tokens.shape[0]
        if tokens.ndim != 2 or tokens.shape[1] != self.T:
            raise SseError("tokens must be [B, %d]" % self.T)
        out = np.empty((B, self.E), dtype=np.float32)
        self._check(self.lib.sse_encode_host(self._h, side, tokens.ctypes.data, B, out.ctypes.data, int(normalize)))
        return out

    # -- index / retrieval
    def index_set(self, tgt, n_local: Optional[int] = None, global_offset: int = 0):
        if isinstance(tgt, np.ndarray):
            tgt = np.ascontiguousarray(tgt, dtype=np.float32)
        if len(tgt.shape) != 2 or int(tgt.shape[1]) != self.E:
            # a stale targetEncodingIndex.tsv written with another encoding_size would be read as n * E_model floats
            raise SseError("index rows must be [N, %d] (encoding_size of this model), got %s" % (self.E, tuple(tgt.shape)))
        n = int(tgt.shape[0]) if n_local is None else n_local
        self._check(self.lib.sse_index_set(self._h, _ptr(tgt), n, global_offset))

    def index_build(self, tgt_tokens, global_offset: int = 0, batch: int = 10000):
        if isinstance(tgt_tokens, np.ndarray):
            tgt_tokens = np.ascontiguousarray(tgt_tokens, dtype=np.int32)
        n = int(tgt_tokens.shape[0])
        self._check(self.lib.sse_index_build(self._h, _ptr(tgt_tokens), n, global_offset, batch))

    def index_get(self, row0: int, n: int) -> np.ndarray:
        out = np.empty((n, self.E), dtype=np.float32)
        self._check(self.lib.sse_index_get(self._h, row0, n, out.ctypes.data))
        return out

    def search(self, q_dev, Q: int, k: int, scores_dev, idx_dev, stream=None):
        self._check(self.lib.sse_search(self._h, _ptr(q_dev), Q, k, _ptr(scores_dev), _ptr(idx_dev), _stream_ptr(stream)))

    def search_packed(self, q_dev, Q: int, k: int, packed_dev, stream=None):
        """local top-k as one packed [Q,2k] fp32 block (scores | int32 global ids bit-cast): the NCCL message"""
        self._check(self.lib.sse_search_packed(self._h, _ptr(q_dev), Q, k, _ptr(packed_dev), _stream_ptr(stream)))

    def merge_packed(self, gathered_dev, G: int, Q: int, k: int, scores_dev, idx_dev, stream=None):
        """gathered [G,Q,2k] packed blocks of G shards -> k best per row"""
        self._check(self.lib.sse_merge_packed(self._h, _ptr(gathered_dev), G, Q, k, _ptr(scores_dev), _ptr(idx_dev), _stream_ptr(stream)))

    def search_stats(self) -> Dict[str, int]:
        c, r, f, it = C.c_int64(), C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.sse_search_stats(self._h, C.byref(c), C.byref(r), C.byref(f), C.byref(it)))
        return {"candidates": int(c.value), "rows": int(r.value), "fallback_rows": int(f.value), "scan_items": int(it.value)}

    def merge_topk(self, cand_s_dev, cand_i_dev, Q: int, Cn: int, k: int, scores_dev, idx_dev, stream=None):
        self._check(self.lib.sse_merge_topk(self._h, _ptr(cand_s_dev), _ptr(cand_i_dev), Q, Cn, k, _ptr(scores_dev),
                                            _ptr(idx_dev), _stream_ptr(stream)))

    def query_host(self, tokens, k: int, normalize: bool = True, scores_out=None, idx_out=None):
        """tokens: int32 [Q,T] numpy array or pinned torch tensor; returns (scores [Q,k], idx [Q,k]) numpy."""
        if isinstance(tokens, np.ndarray):
            tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        Q = int(tokens.shape[0])
        if scores_out is None:
            scores_out = np.empty((Q, k), dtype=np.float32)
        if idx_out is None:
            idx_out = np.empty((Q, k), dtype=np.int32)
        self._check(self.lib.sse_query_host(self._h, _ptr(tokens), Q, k, int(normalize), _ptr(scores_out), _ptr(idx_out)))
        return scores_out, idx_out

    def topk_batch(self, q_dev, Q: int, tgt_dev, n_tgt: int, k: int, scores_dev, idx_dev, normalize_scores: bool = True, stream=None):
        """tf.nn.top_k over the batch's own target encodings (sse_model.py:344-350); the resident index is untouched."""
        self._check(self.lib.sse_topk_batch(self._h, _ptr(q_dev), Q, _ptr(tgt_dev), n_tgt, k, int(normalize_scores),
                                            _ptr(scores_dev), _ptr(idx_dev), _stream_ptr(stream)))

    def token_errors(self, stream=None) -> int:
        """out-of-range token ids seen since the last check (synchronises the stream, resets the counter)"""
        n = C.c_int64()
        self._check(self.lib.sse_token_errors(self._h, C.byref(n), _stream_ptr(stream)))
        return int(n.value)

    def l2_normalize_rows(self, x_dev, rows: int, cols: int, stream=None):
        self._check(self.lib.sse_l2_normalize_rows(self._h, _ptr(x_dev), rows, cols, _stream_ptr(stream)))

    # -- training
    def pair_score(self, src_dev, tgt_dev, B: int, cos_dev, stream=None):
        self._check(self.lib.sse_pair_score(self._h, _ptr(src_dev), _ptr(tgt_dev), B, _ptr(cos_dev), _stream_ptr(stream)))

    def train_step(self, src, tgt, labels, stream=None, want_scalars: bool = True):
        if isinstance(src, np.ndarray):
            src = np.ascontiguousarray(src, dtype=np.int32)
            tgt = np.ascontiguousarray(tgt, dtype=np.int32)
            labels = np.ascontiguousarray(labels, dtype=np.float32)
        B = int(src.shape[0])
        loss, acc, gn = C.c_float(), C.c_float(), C.c_float()
        if want_scalars:
            self._check(self.lib.sse_train_step(self._h, _ptr(src), _ptr(tgt), _ptr(labels), B, C.byref(loss),
                                                C.byref(acc), C.byref(gn), _stream_ptr(stream)))
            return loss.value, acc.value, gn.value
        self._check(self.lib.sse_train_step(self._h, _ptr(src), _ptr(tgt), _ptr(labels), B, None, None, None,
                                            _stream_ptr(stream)))
        return None

    def train_grads(self, src, tgt, labels, B_global: int, stream=None):
        if isinstance(src, np.ndarray):
            src = np.ascontiguousarray(src, dtype=np.int32)
            tgt = np.ascontiguousarray(tgt, dtype=np.int32)
            labels = np.ascontiguousarray(labels, dtype=np.float32)
        B = int(src.shape[0])
        loss, acc = C.c_float(), C.c_float()
        self._check(self.lib.sse_train_grads(self._h, _ptr(src), _ptr(tgt), _ptr(labels), B, B_global, C.byref(loss),
                                             C.byref(acc), _stream_ptr(stream)))
        return loss.value, acc.value

    def grad_arena(self) -> Tuple[int, int]:
        p = _P()
        n = C.c_int64()
        self._check(self.lib.sse_grad_arena(self._h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def train_apply(self, stream=None, want_scalars: bool = True):
        loss, acc, gn = C.c_float(), C.c_float(), C.c_float()
        if not want_scalars:
            self._check(self.lib.sse_train_apply(self._h, None, None, None, _stream_ptr(stream)))
            return None
        self._check(self.lib.sse_train_apply(self._h, C.byref(loss), C.byref(acc), C.byref(gn), _stream_ptr(stream)))
        return loss.value, acc.value, gn.value

    def sampler_set(self, src_rows: np.ndarray, ver_off: np.ndarray, ver_rows: np.ndarray, tgt_rows: np.ndarray):
        """register the training corpus with the device-side batch sampler (arrays as built by data.Data._build_arrays)"""
        src_rows = np.ascontiguousarray(src_rows, dtype=np.int32)
        tgt_rows = np.ascontiguousarray(tgt_rows, dtype=np.int32)
        ver_off = np.ascontiguousarray(ver_off, dtype=np.int64)
        ver_rows = np.ascontiguousarray(ver_rows, dtype=np.int32)
        if src_rows.ndim != 2 or src_rows.shape[1] != self.T or tgt_rows.ndim != 2 or tgt_rows.shape[1] != self.T:
            raise SseError("sampler rows must be [*, %d]" % self.T)
        if ver_off.shape[0] != src_rows.shape[0] + 1 or int(ver_off[-1]) != ver_rows.shape[0]:
            raise SseError("ver_off / ver_rows do not describe %d positives" % src_rows.shape[0])
        self._check(self.lib.sse_sampler_set(self._h, src_rows.ctypes.data, src_rows.shape[0], ver_off.ctypes.data, ver_rows.ctypes.data,
                                             tgt_rows.ctypes.data, tgt_rows.shape[0]))

    def sampler_batch(self, start: int, batch_size: int, seed: int, step: int, src_dev, tgt_dev, labels_dev, stream=None) -> int:
        """device tensors [2*batch_size,T] int32 x2 and [2*batch_size] float32 are filled; returns the number of pair rows written"""
        n = C.c_int()
        self._check(self.lib.sse_sampler_batch(self._h, start, batch_size, seed, step, _ptr(src_dev), _ptr(tgt_dev), _ptr(labels_dev), C.byref(n),
                                               _stream_ptr(stream)))
        return int(n.value)

    def lr_decay(self):
        self._check(self.lib.sse_lr_decay(self._h))

    def scalars(self) -> Tuple[float, int]:
        lr = C.c_float()
        gs = C.c_int64()
        self._check(self.lib.sse_get_scalars(self._h, C.byref(lr), C.byref(gs)))
        return lr.value, int(gs.value)

    def set_scalars(self, lr: float, global_step: int):
        self._check(self.lib.sse_set_scalars(self._h, lr, global_step))

    def debug_gemm_tc(self, a_dev, b_dev, M: int, N: int, K: int, d_dev, fmt: int = 1, split_k: int = 1, alpha: float = 1.0, beta: float = 0.0, stream=None):
        """test hook: D[M,N] = alpha A[M,K] B[N,K]^T (+ beta D) on the internal tcgen05 GEMM (fp32 in / out, 16-bit operands)"""
        self._check(self.lib.sse_debug_gemm_tc(self._h, _ptr(a_dev), _ptr(b_dev), M, N, K, fmt, split_k, alpha, beta, _ptr(d_dev), _stream_ptr(stream)))

    def launch_count(self) -> int:
        return int(self.lib.sse_launch_count(self._h))

    def set_option(self, key: str, value: int):
        self._check(self.lib.sse_set_option(self._h, key.encode(), int(value)))


# ---------------------------------------------------------------------------------------------------------
# targetEncodingIndex.tsv fast writer / reader (host only: works without a GPU)
def _tsv_check(rc):
    if rc != SSE_OK:
        raise SseError(load_library().sse_tsv_last_error().decode("utf-8", "replace"))


def tsv_format_f32(values) -> List[str]:
    """str(np.float32(v)) for every value, formatted by the native writer (test / small-scale helper)."""
    lib = load_library()
    v = np.ascontiguousarray(values, dtype=np.float32).ravel()
    out = C.create_string_buffer(int(v.size) * 16 + 16)
    ends = np.zeros(v.size, np.int64)
    _tsv_check(lib.sse_tsv_format_f32(v.ctypes.data, v.size, C.addressof(out), len(out), ends.ctypes.data))
    raw = out.raw
    res, b = [], 0
    for e in ends.tolist():
        res.append(raw[b:e].decode("ascii")); b = e
    return res


def tsv_write_index(path: str, ids: Sequence[str], texts: Sequence[str], rows, append: bool = False, threads: int = 0) -> None:
    """Write rows `id \\t text \\t comma-joined str(np.float32)` exactly as reference sse_index.py:93-97 does."""
    lib = load_library()
    r = np.ascontiguousarray(rows, dtype=np.float32)
    n = len(ids)
    if r.ndim != 2 or r.shape[0] != n or len(texts) != n:
        raise ValueError("ids / texts / rows disagree: %d %d %s" % (n, len(texts), r.shape))
    a_ids = (C.c_char_p * n)(*[s.encode("utf-8") for s in ids])
    a_txt = (C.c_char_p * n)(*[s.encode("utf-8") for s in texts])
    _tsv_check(lib.sse_tsv_write_index(path.encode("utf-8"), a_ids, a_txt, r.ctypes.data, n, r.shape[1], 1 if append else 0, threads))


def tsv_read_index(path: str, threads: int = 0) -> Tuple[List[str], List[str], np.ndarray, int]:
    """Parse an index file as reference sse_evaluator.py:79-88 does (strip, split on tabs, rows without three
    fields are skipped).  Returns (ids, texts, float32 [N,E], n_skipped)."""
    lib = load_library()
    with open(path, "rb") as f:
        buf = f.read()
    E = 0
    for line in buf.split(b"\n", 64)[:64]:                 # E from the first well-formed row
        info = line.strip().split(b"\t")
        if len(info) == 3:
            E = info[2].count(b",") + 1
            break
    if E == 0:
        return [], [], np.zeros((0, 0), np.float32), buf.count(b"\n")
    max_rows = buf.count(b"\n") + 1
    out = np.empty((max_rows, E), np.float32)
    spans = np.empty((max_rows, 4), np.int64)
    n, skipped = C.c_int64(0), C.c_int64(0)
    _tsv_check(lib.sse_tsv_parse_index(buf, len(buf), E, max_rows, out.ctypes.data, spans.ctypes.data, C.byref(n), C.byref(skipped), threads))
    n = n.value
    sp = spans[:n].tolist()
    ids = [buf[a:b].decode("utf-8") for a, b, _c, _d in sp]
    texts = [buf[c:d].decode("utf-8") for _a, _b, c, d in sp]
    return ids, texts, out[:n], skipped.value


# ---------------------------------------------------------------------------------------------------------
# batched subword tokenizer + padder (host only)
class NativeTokenizer(object):
    """Bulk `encode + pad` of lower-cased sentences with a reference vocabulary (csrc/subword_tok.cpp)."""

    def __init__(self, subtoken_strings: Sequence[str]):
        lib = load_library()
        arr = (C.c_char_p * len(subtoken_strings))(*[s.encode("utf-8") for s in subtoken_strings])
        self._h = _P()
        if lib.sse_tok_create(arr, len(subtoken_strings), C.byref(self._h)) != SSE_OK:
            raise SseError(lib.sse_tok_last_error().decode("utf-8", "replace"))

    def encode_batch(self, texts: Sequence[str], T: int, threads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """-> (int32 [n,T] padded rows, int32 [n] subtoken counts before padding / truncation)."""
        lib = load_library()
        n = len(texts)
        rows = np.zeros((n, T), np.int32)
        lengths = np.zeros(n, np.int32)
        if n == 0:
            return rows, lengths
        arr = (C.c_char_p * n)(*[t.encode("utf-8") for t in texts])
        if lib.sse_tok_encode_batch(self._h, arr, n, T, rows.ctypes.data, lengths.ctypes.data, threads) != SSE_OK:
            raise SseError(lib.sse_tok_last_error().decode("utf-8", "replace"))
        return rows, lengths

    def close(self):
        if self._h:
            load_library().sse_tok_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def crc32c(data, seed: int = 0) -> int:
    """CRC-32C of a bytes-like object / numpy array (native, slicing-by-8)."""
    lib = load_library()
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data)
        return int(lib.sse_crc32c(a.ctypes.data, a.nbytes, seed))
    b = bytes(data)
    return int(lib.sse_crc32c(b, len(b), seed))
