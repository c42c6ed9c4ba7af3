# coding=utf-8
"""TensorFlow "V2" checkpoint (tensor bundle) reader / writer -- pure python + numpy.

The reference saves with `tf.train.Saver(tf.global_variables())` (sse_model.py:138; `SSE-LSTM.ckpt-*.index` +
`.data-00000-of-00001`, sse_train.py:205,212,232).  This module lets such a checkpoint be loaded into the B200 path
(`Saver.restore` accepts a bundle prefix) and writes bundles with the same layout.

PARITY UNPINNED: TensorFlow cannot be installed here, so no bundle written by the real library was available to pin
this against; it follows the published on-disk format:
  * `<prefix>.index` is an immutable sorted string table (the LevelDB table format TensorFlow vendors in
    `tensorflow/core/lib/io/`): data blocks of prefix-compressed entries (varint shared / non_shared / value_len, key
    suffix, value; uint32 restart offsets + count at the end), each followed by a 1-byte compression type and a masked
    crc32c; an index block mapping separator keys to block handles; a 48-byte footer (metaindex handle, index handle,
    padding, magic 0xdb4775248b80fb57).
  * key "" holds `BundleHeaderProto` (num_shards, endianness, version); every other key is a variable name whose
    value is a `BundleEntryProto` (dtype, shape, shard_id, offset, size, masked crc32c of the bytes).
  * `<prefix>.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes.
What this module writes it reads back bit-exactly (tests/test_tf_bundle.py); compatibility with the real library is
by construction only and has to be checked against a real checkpoint when one is at hand.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Tuple

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"), 6: np.dtype("i1"),
           9: np.dtype("<i8"), 10: np.dtype("bool"), 17: np.dtype("<u2"), 19: np.dtype("<f2"), 22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DTYPE_IDS = {v: k for k, v in _DTYPES.items()}


# ----------------------------------------------------------------------------------------------- crc32c (Castagnoli)
def _make_table():
    tbl = np.zeros(256, np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl[i] = c
    return tbl


_CRC_TABLE = _make_table()


def _crc32c_py(data: bytes) -> int:
    crc = 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in data:
        crc = int(tbl[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def crc32c(data: bytes) -> int:
    """Native (csrc/tsv_io.cpp: sse_crc32c) when the library is built, else the byte-wise python loop."""
    try:
        import sse_ffi
        return sse_ffi.crc32c(data)
    except Exception:
        return _crc32c_py(data)


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------- varints / protobuf wire
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_message(buf: bytes) -> Dict[int, list]:
    """field number -> list of raw values (int for varint / fixed, bytes for length-delimited)."""
    out: Dict[int, list] = {}
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _field(tag: int, wt: int) -> bytes:
    return _put_varint((tag << 3) | wt)


def _encode_shape(shape) -> bytes:
    body = b""
    for d in shape:
        dim = _field(1, 0) + _put_varint(int(d))
        body += _field(2, 2) + _put_varint(len(dim)) + dim
    return body


def _decode_shape(buf: bytes) -> Tuple[int, ...]:
    msg = _parse_message(buf)
    dims = []
    for d in msg.get(2, []):
        dm = _parse_message(d)
        dims.append(_signed64(dm.get(1, [0])[0]))
    return tuple(dims)


# ----------------------------------------------------------------------------------------------- snappy (decode only)
def _snappy_uncompress(buf: bytes) -> bytes:
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]; pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little"); pos += nb
            ln += 1
            out += buf[pos:pos + ln]; pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]; pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8); pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little"); pos += 4
        for _ in range(ln):
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy: length mismatch")
    return bytes(out)


# ----------------------------------------------------------------------------------------------- table reader
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        want = struct.unpack_from("<I", data, offset + size + 1)[0]
        if masked_crc32c(data[offset:offset + size + 1]) != want:
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    if ctype == 0:
        return raw
    if ctype == 1:
        return _snappy_uncompress(raw)
    raise ValueError("unknown block compression type %d" % ctype)


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    out, pos, key = [], 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]; pos += non_shared
        out.append((key, block[pos:pos + vlen])); pos += vlen
    return out


def read_table(path: str, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow/LevelDB table (bad magic)" % path)
    footer = data[len(data) - 48:]
    _mo, pos = _get_varint(footer, 0)
    _ms, pos = _get_varint(footer, pos)
    io, pos = _get_varint(footer, pos)
    isz, pos = _get_varint(footer, pos)
    entries: List[Tuple[bytes, bytes]] = []
    for _sep, handle in _block_entries(_read_block(data, io, isz, verify)):
        bo, p = _get_varint(handle, 0)
        bs, p = _get_varint(handle, p)
        entries.extend(_block_entries(_read_block(data, bo, bs, verify)))
    return entries


# ----------------------------------------------------------------------------------------------- bundle reader
def list_variables(prefix: str) -> Dict[str, Tuple[np.dtype, Tuple[int, ...]]]:
    out = {}
    for key, val in read_table(prefix + ".index"):
        if key == b"":
            continue
        e = _parse_message(val)
        dt = _DTYPES.get(e.get(1, [0])[0])
        out[key.decode("utf-8")] = (dt, _decode_shape(e[2][0]) if 2 in e else ())
    return out


def read_bundle(prefix: str, verify_crc: bool = False) -> Dict[str, np.ndarray]:
    """name -> array for every dense variable of the checkpoint `prefix` (without the .index / .data suffix)."""
    entries = read_table(prefix + ".index")
    num_shards = 1
    for key, val in entries:
        if key == b"":
            h = _parse_message(val)
            num_shards = h.get(1, [1])[0]
            if h.get(2, [0])[0] != 0:
                raise ValueError("big-endian bundles are not supported")
    shards: Dict[int, np.memmap] = {}
    out: Dict[str, np.ndarray] = {}
    for key, val in entries:
        if key == b"":
            continue
        e = _parse_message(val)
        name = key.decode("utf-8")
        if 7 in e:
            raise ValueError("variable %s is stored as slices (partitioned variable): not supported" % name)
        dt_id = e.get(1, [0])[0]
        if dt_id not in _DTYPES:
            continue                                            # strings / resources: not part of the model
        shape = _decode_shape(e[2][0]) if 2 in e else ()
        shard, off, size = e.get(3, [0])[0], e.get(4, [0])[0], e.get(5, [0])[0]
        if shard not in shards:
            shards[shard] = np.memmap("%s.data-%05d-of-%05d" % (prefix, shard, num_shards), dtype=np.uint8, mode="r")
        raw = bytes(shards[shard][off:off + size])
        if verify_crc and 6 in e and masked_crc32c(raw) != e[6][0]:
            raise ValueError("checksum mismatch for variable %s" % name)
        arr = np.frombuffer(raw, dtype=_DTYPES[dt_id])
        if int(np.prod(shape, dtype=np.int64)) != arr.size:
            raise ValueError("variable %s: %d bytes do not match shape %s" % (name, size, shape))
        out[name] = arr.reshape(shape).copy()
    return out


# ----------------------------------------------------------------------------------------------- writer
class _BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""
        self.interval = restart_interval

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count < self.interval:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))

    def size(self) -> int:
        return len(self.buf) + 4 * len(self.restarts) + 4


def write_table(path: str, entries: List[Tuple[bytes, bytes]], block_size: int = 4096):
    """entries must be sorted by key (bytewise)."""
    out = bytearray()

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block)
        out.append(0)                                           # kNoCompression
        out.extend(struct.pack("<I", masked_crc32c(block + b"\x00")))
        return _put_varint(off) + _put_varint(len(block))

    index = _BlockBuilder(restart_interval=1)
    cur = _BlockBuilder()
    for key, value in entries:
        cur.add(key, value)
        if cur.size() >= block_size:
            index.add(cur.last, emit(cur.finish()))             # separator = last key of the block
            cur = _BlockBuilder()
    if cur.count or not entries:
        index.add(cur.last, emit(cur.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(out))


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]):
    """Single-shard bundle: <prefix>.index + <prefix>.data-00000-of-00001."""
    names = sorted(tensors, key=lambda s: s.encode("utf-8"))
    header = _field(1, 0) + _put_varint(1) + _field(2, 0) + _put_varint(0)
    version = _field(1, 0) + _put_varint(1)                     # VersionDef.producer = 1
    header += _field(3, 2) + _put_varint(len(version)) + version
    entries = [(b"", header)]
    off = 0
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in names:
            a = np.asarray(tensors[name])
            if a.ndim and not a.flags.c_contiguous:
                a = np.ascontiguousarray(a)               # (ascontiguousarray would turn a 0-d scalar into shape (1,))
            if a.dtype.byteorder == ">":
                a = a.astype(a.dtype.newbyteorder("<"))
            dt_id = _DTYPE_IDS.get(a.dtype)
            if dt_id is None:
                raise ValueError("unsupported dtype %s for %s" % (a.dtype, name))
            raw = a.tobytes()
            f.write(raw)
            shape = _encode_shape(a.shape)
            e = _field(1, 0) + _put_varint(dt_id) + _field(2, 2) + _put_varint(len(shape)) + shape
            if off:
                e += _field(4, 0) + _put_varint(off)
            e += _field(5, 0) + _put_varint(len(raw)) + _field(6, 5) + struct.pack("<I", masked_crc32c(raw))
            entries.append((name.encode("utf-8"), e))
            off += len(raw)
    write_table(prefix + ".index", entries)
