# coding=utf-8
"""Subword text -> token ids for the SSE entry points (host side of the hot path's inputs).

Own implementation of the ENCODE half of the reference's Tensor2Tensor-style
``text_encoder.SubwordTextEncoder`` + ``tokenizer`` (reference text_encoder.py:427-436,
491-532, 334-356, 731-746; tokenizer.py:68-90): load a ``vocabulary.txt`` written by the
reference, split text at alphanumeric / non-alphanumeric boundaries, escape each token,
and greedily match the longest vocabulary subtoken.  Building a NEW vocabulary from a
corpus (``build_to_target_size``) is out of the hot-path scope: use the
reference's ``text_encoder_build_subword.py`` once and keep the ``vocabulary.txt``.

Checked against the real reference encoder on real data via tests/golden/subword.npz.
"""
from __future__ import annotations

import unicodedata
from functools import lru_cache
from typing import Dict, Iterable, List

PAD = "<pad>"
EOS = "<EOS>"
RESERVED_TOKENS = [PAD, EOS]
PAD_ID = 0   # reference text_encoder.py:44
EOS_ID = 1   # reference text_encoder.py:45

_ESCAPE_ALPHABET = set("\\_u;0123456789")


@lru_cache(maxsize=65536)
def _is_alnum(ch: str) -> bool:
    c = unicodedata.category(ch)
    return c[0] == "L" or c[0] == "N"


def split_tokens(text: str) -> List[str]:
    """Runs of alphanumeric / non-alphanumeric characters; a lone space between two
    alphanumeric runs is dropped (tokenizer.py:68-90)."""
    if not text:
        return []
    out: List[str] = []
    start = 0
    prev = _is_alnum(text[0])
    for pos in range(1, len(text)):
        cur = _is_alnum(text[pos])
        if cur != prev:
            piece = text[start:pos]
            if piece != " " or start == 0:
                out.append(piece)
            start = pos
            prev = cur
    out.append(text[start:])
    return out


class SubwordTextEncoder(object):
    """encode()/vocab_size of the reference class, for vocab files it wrote."""

    def __init__(self, filename: str = None):
        self._strings: List[str] = []
        self._ids: Dict[str, int] = {}
        self._alphabet = set()
        self._maxlen = 0
        self._cache: Dict[str, List[int]] = {}
        if filename is not None:
            with open(filename, "r", encoding="utf-8") as f:
                self._load(f)

    def _load(self, lines: Iterable[str]):
        strings = []
        for line in lines:
            s = line.strip()
            if len(s) >= 2 and ((s[0] == "'" and s[-1] == "'") or (s[0] == '"' and s[-1] == '"')):
                s = s[1:-1]
            strings.append(s)
        self._strings = strings
        self._maxlen = max(len(s) for s in strings) if strings else 0
        self._ids = {s: i for i, s in enumerate(strings) if s}
        self._alphabet = {c for s in strings for c in s} | _ESCAPE_ALPHABET

    @property
    def vocab_size(self) -> int:
        return len(self._strings)

    def _escape(self, token: str) -> str:
        token = token.replace("\\", "\\\\").replace("_", "\\u")
        alpha = self._alphabet
        return "".join(c if (c in alpha and c != "\n") else "\\%d;" % ord(c) for c in token) + "_"

    def _match(self, esc: str) -> List[int]:
        ids, start, n = [], 0, len(esc)
        table, maxlen = self._ids, self._maxlen
        while start < n:
            end = min(n, start + maxlen)
            while end > start:
                i = table.get(esc[start:end])
                if i is not None:
                    ids.append(i)
                    start = end
                    break
                end -= 1
            else:
                raise AssertionError("Token substring not found in subtoken vocabulary.")
        return ids

    def encode(self, raw_text: str) -> List[int]:
        out: List[int] = []
        cache = self._cache
        for tok in split_tokens(raw_text):
            ids = cache.get(tok)
            if ids is None:
                ids = self._match(self._escape(tok))
                if len(cache) < 1_000_000:
                    cache[tok] = ids
            out.extend(ids)
        return out

    def encode_batch(self, raw_texts, max_seq_length: int, threads: int = 0):
        """Bulk path: encode + pad every sentence (already lower-cased by the caller, as the entry points do) into
        int32 [n, max_seq_length] rows with the native multi-threaded tokenizer (csrc/subword_tok.cpp); also returns
        the un-padded subtoken counts.  Same ids as [pad_tokens(self.encode(t), T) for t in raw_texts]."""
        if getattr(self, "_native", None) is None:
            import sse_ffi
            self._native = sse_ffi.NativeTokenizer(self._strings)
        texts = list(raw_texts)
        rows, lengths = self._native.encode_batch([t if "\x00" not in t else "" for t in texts], max_seq_length, threads)
        for i, t in enumerate(texts):             # C strings end at NUL: such (pathological) rows take the python path
            if "\x00" in t:
                ids = self.encode(t)
                rows[i] = pad_tokens(ids, max_seq_length)
                lengths[i] = len(ids)
        return rows, lengths

    def decode_list(self, ids: Iterable[int]) -> List[str]:
        return [self._strings[i] if 0 <= i < len(self._strings) else "" for i in ids]


def pad_tokens(ids: List[int], max_seq_length: int) -> List[int]:
    """Left-pad / truncate rule of the reference (data_utils.py:149-155, sse_index.py:79-85)."""
    if len(ids) > max_seq_length - 2:
        return [PAD_ID] + list(ids[: max_seq_length - 2]) + [EOS_ID]
    return [PAD_ID] * (max_seq_length - len(ids) - 1) + list(ids) + [EOS_ID]
