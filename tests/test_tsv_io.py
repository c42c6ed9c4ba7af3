"""targetEncodingIndex.tsv fast writer / reader (native, host-only) against numpy's str(np.float32), the oracle's
restatement of the reference writer/reader (sse_index.py:93-97, sse_evaluator.py:79-88) and the fixture written by the
real reference code path (tests/golden/index_rows.tsv)."""
import os

import numpy as np
import pytest

import sse_ffi
import sse_oracle as O


def special_values():
    return np.array([0, -0.0, 1, -1, 1e-4, 9.999e-5, 1e-5, 1e16, 9.9999e15, 1e15, 123456789, 0.1, 0.5, 1e-45, 3.4028235e38,
                     np.inf, -np.inf, np.nan, 16777216, 1e7, 1.5e-5, 2 ** -24, 1e6, 999999.94, 999999.9, 1e5, 1.17549435e-38], np.float32)


def test_format_matches_numpy_str_float32_over_all_magnitudes():
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.standard_normal(100000).astype(np.float32) * 0.06,                                  # what an index holds
        rng.standard_normal(20000).astype(np.float32) * 10.0 ** rng.integers(-45, 39, 20000).astype(np.float32),
        special_values(),
        rng.integers(0, 2 ** 32, 200000, dtype=np.uint64).astype(np.uint32).view(np.float32)])  # random bit patterns
    got = sse_ffi.tsv_format_f32(vals)
    want = [str(v) for v in vals]
    assert got == want


def test_writer_bytes_equal_reference_writer_and_golden_fixture(tmp_path, golden_dir):
    g = np.load(os.path.join(golden_dir, "index_rows.npz"))
    enc = g["enc"].astype(np.float32)
    ids = ["id%d" % i for i in range(len(enc))]
    texts = ["some target text %d é中" % i for i in range(len(enc))]
    p = str(tmp_path / "idx.tsv")
    sse_ffi.tsv_write_index(p, ids, texts, enc, threads=3)
    want = "".join(O.format_index_row(i, t, r) for i, t, r in zip(ids, texts, enc))
    assert open(p, "rb").read() == want.encode("utf-8")
    # the float fields of the fixture produced by the real reference writer loop
    lines = open(os.path.join(golden_dir, "index_rows.tsv"), encoding="utf-8").readlines()
    gi, gt, _ = O.parse_index_lines(lines)
    sse_ffi.tsv_write_index(p, gi, gt, enc[:len(gi)], threads=2)
    assert open(p, encoding="utf-8").readlines() == lines
    # append mode continues the same file
    sse_ffi.tsv_write_index(p, gi[:2], gt[:2], enc[:2], append=True)
    assert open(p, encoding="utf-8").readlines() == lines + lines[:2]


def test_reader_round_trip_is_bit_exact_and_follows_the_three_field_rule(tmp_path):
    rng = np.random.default_rng(1)
    N, E = 5000, 64
    enc = rng.standard_normal((N, E)).astype(np.float32)
    enc /= np.linalg.norm(enc, axis=1, keepdims=True)
    enc[:30].reshape(-1)[:len(special_values())] = np.nan_to_num(special_values(), nan=0.25)
    ids = ["t%d" % i for i in range(N)]
    texts = ["text %d" % i for i in range(N)]
    p = str(tmp_path / "idx.tsv")
    sse_ffi.tsv_write_index(p, ids, texts, enc)
    raw = open(p, "rb").read().split(b"\n")
    # malformed rows as the reference tolerates them: blank line, 2 fields, 4 fields, padded line, CRLF, no final newline
    raw.insert(10, b"")
    raw.insert(20, b"only\ttwo")
    raw.insert(30, b"a\tb\t0.5\textra")
    raw[40] = b"  " + raw[40] + b" \r"
    blob = b"\n".join(raw).rstrip(b"\n")
    open(p, "wb").write(blob)
    gi, gt, ge, skipped = sse_ffi.tsv_read_index(p, threads=4)
    oi, ot, oe = O.parse_index_lines(blob.decode("utf-8").split("\n"))
    assert gi == oi and gt == ot and skipped == 3
    assert ge.dtype == np.float32 and ge.shape == (N, E)
    assert np.array_equal(ge.view(np.uint32), enc.view(np.uint32))            # bit-exact round trip
    # the reference parses the decimals with float() into float64; narrowed to the index dtype they are the same numbers
    assert np.array_equal(oe.astype(np.float32).view(np.uint32), ge.view(np.uint32))


def test_reader_rejects_a_row_with_a_broken_vector(tmp_path):
    p = str(tmp_path / "bad.tsv")
    open(p, "w").write("a\tb\t0.1,0.2\nc\td\t0.1,zzz\n")
    with pytest.raises(sse_ffi.SseError):
        sse_ffi.tsv_read_index(p)
    open(p, "w").write("a\tb\t0.1,0.2\nc\td\t0.1\n")
    with pytest.raises(sse_ffi.SseError):
        sse_ffi.tsv_read_index(p)


def test_property_format_and_round_trip_for_arbitrary_float32_bits():
    """hypothesis: every float32 bit pattern prints like numpy prints it and parses back to the same bits."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None, derandomize=True)
    @given(st.lists(st.integers(min_value=0, max_value=2 ** 32 - 1), min_size=1, max_size=64))
    def check(bits):
        v = np.array(bits, np.uint32).view(np.float32)
        got = sse_ffi.tsv_format_f32(v)
        assert got == [str(x) for x in v]
        back = np.array([np.float32(s) for s in got], np.float32)
        same = (back.view(np.uint32) == v.view(np.uint32)) | (np.isnan(back) & np.isnan(v))
        assert same.all()

    check()
