"""TF V2 checkpoint (tensor bundle) reader / writer: format-level checks (unpinned against the real library -- see
tf_bundle.py): CRC-32C known answers, varint / protobuf wire round trips, table blocks with prefix compression and
restarts over many keys, bit-exact tensor round trip, corruption detection, snappy block decoding."""
import os
import struct

import numpy as np
import pytest

import sse_ffi
import tf_bundle as TB


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors
    for fn in (TB._crc32c_py, sse_ffi.crc32c):
        assert fn(b"123456789") == 0xE3069283
        assert fn(bytes(32)) == 0x8A9136AA
        assert fn(bytes([0xFF] * 32)) == 0x62A8AB43
        assert fn(bytes(range(32))) == 0x46DD794E
    blob = np.random.default_rng(0).integers(0, 256, 100003, dtype=np.uint8).tobytes()
    assert sse_ffi.crc32c(blob) == TB._crc32c_py(blob)
    assert sse_ffi.crc32c(blob[50000:], sse_ffi.crc32c(blob[:50000])) == sse_ffi.crc32c(blob)      # incremental
    # the masking LevelDB / TensorFlow apply to stored checksums
    assert TB.masked_crc32c(b"foo") == ((((TB.crc32c(b"foo") >> 15) | (TB.crc32c(b"foo") << 17)) + 0xa282ead8) & 0xFFFFFFFF)


def test_varint_and_wire_round_trip():
    for v in (0, 1, 127, 128, 300, 2 ** 31, 2 ** 63 - 1):
        enc = TB._put_varint(v)
        assert TB._get_varint(enc + b"\x00", 0) == (v, len(enc))
    shape = (3, 0, 1 << 40)
    assert TB._decode_shape(TB._encode_shape(shape)) == shape
    msg = TB._field(1, 0) + TB._put_varint(9) + TB._field(6, 5) + struct.pack("<I", 0xDEADBEEF) + TB._field(2, 2) + TB._put_varint(3) + b"abc"
    assert TB._parse_message(msg) == {1: [9], 6: [0xDEADBEEF], 2: [b"abc"]}


def test_table_with_many_keys_blocks_and_restarts(tmp_path):
    rng = np.random.default_rng(1)
    keys = sorted({("scope_%03d/rnn/basic_lstm_cell/kernel_%d" % (i % 37, i)).encode() for i in range(900)})
    entries = [(k, rng.integers(0, 256, int(rng.integers(0, 90)), dtype=np.uint8).tobytes()) for k in keys]
    p = str(tmp_path / "t.index")
    TB.write_table(p, entries, block_size=512)
    assert TB.read_table(p) == entries
    raw = bytearray(open(p, "rb").read())
    assert struct.unpack_from("<Q", raw, len(raw) - 8)[0] == 0xdb4775248b80fb57
    raw[100] ^= 0x40                                              # flip a bit inside the first data block
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        TB.read_table(p)


def test_bundle_round_trip_is_bit_exact(tmp_path):
    rng = np.random.default_rng(2)
    t = {"word_embedding": rng.standard_normal((501, 24)).astype(np.float32),
         "source_encoder/rnn/basic_lstm_cell/kernel": rng.standard_normal((56, 128)).astype(np.float32),
         "source_encoder/rnn/basic_lstm_cell/bias": np.zeros(128, np.float32),
         "global_step": np.array(1234, np.int64), "learning_rate": np.array(0.875, np.float32),
         "flags": np.array([True, False]), "half": rng.standard_normal(7).astype(np.float16), "empty": np.zeros((0, 3), np.float64)}
    prefix = str(tmp_path / "SSE-LSTM.ckpt-7")
    TB.write_bundle(prefix, t)
    assert sorted(os.listdir(tmp_path)) == ["SSE-LSTM.ckpt-7.data-00000-of-00001", "SSE-LSTM.ckpt-7.index"]
    got = TB.read_bundle(prefix, verify_crc=True)
    assert sorted(got) == sorted(t)
    for k in t:
        assert got[k].dtype == t[k].dtype and got[k].shape == t[k].shape and got[k].tobytes() == t[k].tobytes()
    assert TB.list_variables(prefix)["word_embedding"] == (np.dtype("<f4"), (501, 24))
    # a damaged data file is detected by the per-tensor checksum
    f = prefix + ".data-00000-of-00001"
    raw = bytearray(open(f, "rb").read()); raw[40] ^= 1; open(f, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        TB.read_bundle(prefix, verify_crc=True)


def test_snappy_block_decoder():
    # literal "abcd" then a copy (offset 4, length 8) -> "abcdabcdabcd": tag-1 copy = len-4 in bits 2..4, offset high bits 5..7
    comp = bytes([12, (4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
    assert TB._snappy_uncompress(comp) == b"abcdabcdabcd"
    # 2-byte-offset copy form and a long literal (length 61 -> one extra length byte)
    lit = bytes(range(61))
    comp = TB._put_varint(61 + 5) + bytes([60 << 2, 60]) + lit + bytes([((5 - 1) << 2) | 2, 61, 0])
    assert TB._snappy_uncompress(comp) == lit + lit[:5]
