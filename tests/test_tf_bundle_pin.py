"""Pins of the TF V2 checkpoint reader / writer (tf_bundle.py; reference sse_model.py:138,379-386) against INDEPENDENT
implementations that ship in this image -- TensorFlow itself cannot be installed:
  * the masked CRC-32C (rotation + 0xa282ead8) against tensorboard's own `masked_crc32c` (the one its TFRecord event
    writer uses), and the DataType ids against tensorboard's compiled `types.proto`;
  * the two protobuf messages of a bundle (`BundleHeaderProto`, `BundleEntryProto`, tensor_bundle.proto) against the
    real protobuf runtime: message classes are built from the published schema on top of tensorboard's compiled
    `TensorShapeProto`, `DataType` and `VersionDef`; what tf_bundle writes must parse there field by field, and what the
    runtime serialises must read back through tf_bundle.
Not pinned by any third party: the LevelDB-style table framing of the .index file (only self round trips + format
constants, tests/test_tf_bundle.py) -- stated in DESIGN.md."""
import os

import numpy as np
import pytest

import tf_bundle

tb_record = pytest.importorskip("tensorboard.summary.writer.record_writer")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402
from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2  # noqa: E402


def _bundle_messages():
    """BundleHeaderProto / BundleEntryProto as published in tensorflow/core/protobuf/tensor_bundle.proto."""
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "sse_test/tensor_bundle.proto"
    fd.package = "sse_test"
    fd.syntax = "proto3"
    fd.dependency.extend([tensor_shape_pb2.DESCRIPTOR.name, types_pb2.DESCRIPTOR.name, versions_pb2.DESCRIPTOR.name])
    T = descriptor_pb2.FieldDescriptorProto
    hdr = fd.message_type.add()
    hdr.name = "BundleHeaderProto"
    en = hdr.enum_type.add()
    en.name = "Endianness"
    for n, v in (("LITTLE", 0), ("BIG", 1)):
        ev = en.value.add(); ev.name = n; ev.number = v
    for name, num, typ, tn in (("num_shards", 1, T.TYPE_INT32, None), ("endianness", 2, T.TYPE_ENUM, ".sse_test.BundleHeaderProto.Endianness"),
                               ("version", 3, T.TYPE_MESSAGE, ".tensorboard.VersionDef")):
        f = hdr.field.add(); f.name = name; f.number = num; f.type = typ; f.label = T.LABEL_OPTIONAL
        if tn:
            f.type_name = tn
    ent = fd.message_type.add()
    ent.name = "BundleEntryProto"
    for name, num, typ, tn in (("dtype", 1, T.TYPE_ENUM, ".tensorboard.DataType"), ("shape", 2, T.TYPE_MESSAGE, ".tensorboard.TensorShapeProto"),
                               ("shard_id", 3, T.TYPE_INT32, None), ("offset", 4, T.TYPE_INT64, None), ("size", 5, T.TYPE_INT64, None),
                               ("crc32c", 6, T.TYPE_FIXED32, None)):
        f = ent.field.add(); f.name = name; f.number = num; f.type = typ; f.label = T.LABEL_OPTIONAL
        if tn:
            f.type_name = tn
    pool = descriptor_pool.Default()
    try:
        fdesc = pool.Add(fd)
    except Exception:
        fdesc = pool.FindFileByName(fd.name)
    get = getattr(message_factory, "GetMessageClass", None)
    mk = get if get else message_factory.MessageFactory(pool).GetPrototype
    return mk(fdesc.message_types_by_name["BundleHeaderProto"]), mk(fdesc.message_types_by_name["BundleEntryProto"])


def test_masked_crc_and_dtype_ids_match_tensorboard():
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 64, 4097):
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert tf_bundle.masked_crc32c(data) == int(tb_record.masked_crc32c(data))
    for name, np_dt in (("DT_FLOAT", "<f4"), ("DT_DOUBLE", "<f8"), ("DT_INT32", "<i4"), ("DT_INT64", "<i8"), ("DT_UINT8", "u1"), ("DT_BOOL", "bool"),
                        ("DT_HALF", "<f2"), ("DT_INT16", "<i2"), ("DT_INT8", "i1")):
        assert tf_bundle._DTYPES[getattr(types_pb2, name)] == np.dtype(np_dt), name


def test_written_entries_parse_with_the_real_protobuf_runtime(tmp_path):
    Header, Entry = _bundle_messages()
    rng = np.random.default_rng(1)
    tensors = {"word_embedding": rng.standard_normal((7, 5)).astype(np.float32),
               "source_encoder/rnn/basic_lstm_cell/kernel": rng.standard_normal((9, 16)).astype(np.float32),
               "source_encoder/rnn/basic_lstm_cell/bias": np.zeros(16, np.float32),
               "global_step": np.array(123456789012, np.int64), "learning_rate": np.array(0.25, np.float32)}
    prefix = str(tmp_path / "SSE-LSTM.ckpt-7")
    tf_bundle.write_bundle(prefix, tensors)
    raw = open(prefix + ".data-00000-of-00001", "rb").read()
    seen = {}
    for key, val in tf_bundle.read_table(prefix + ".index"):
        if key == b"":
            h = Header.FromString(val)
            assert h.num_shards == 1 and h.endianness == 0 and h.version.producer == 1
            continue
        e = Entry.FromString(val)
        name = key.decode()
        want = tensors[name]
        assert e.dtype == {np.dtype("<f4"): types_pb2.DT_FLOAT, np.dtype("<i8"): types_pb2.DT_INT64}[want.dtype]
        assert tuple(d.size for d in e.shape.dim) == want.shape and not e.shape.unknown_rank
        assert e.shard_id == 0 and e.size == want.nbytes
        assert raw[e.offset:e.offset + e.size] == want.tobytes()
        assert e.crc32c == int(tb_record.masked_crc32c(want.tobytes()))
        assert e.SerializeToString() == val or Entry.FromString(e.SerializeToString()) == e      # canonical re-encode keeps the fields
        seen[name] = True
    assert set(seen) == set(tensors)


def test_entries_serialised_by_the_runtime_read_back_through_tf_bundle(tmp_path):
    """A bundle whose header / entry VALUES come from the real protobuf library (as TensorFlow's writer produces them),
    framed by the table writer, must load through read_bundle with every checksum verified."""
    Header, Entry = _bundle_messages()
    rng = np.random.default_rng(2)
    tensors = {"shared_encoder/src_M": rng.standard_normal((6, 4)).astype(np.float32), "global_step": np.array(42, np.int64),
               "a/Adagrad": np.full((3, 2, 2), 0.1, np.float32)}
    prefix = str(tmp_path / "ckpt")
    h = Header(); h.num_shards = 1; h.endianness = 0; h.version.producer = 1
    entries, off = [(b"", h.SerializeToString())], 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = tensors[name]
            e = Entry()
            e.dtype = types_pb2.DT_FLOAT if a.dtype == np.float32 else types_pb2.DT_INT64
            for d in a.shape:
                e.shape.dim.add().size = d
            e.shard_id = 0; e.offset = off; e.size = a.nbytes
            e.crc32c = int(tb_record.masked_crc32c(a.tobytes()))
            entries.append((name.encode(), e.SerializeToString()))
            f.write(a.tobytes()); off += a.nbytes
    tf_bundle.write_table(prefix + ".index", entries)
    got = tf_bundle.read_bundle(prefix, verify_crc=True)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)
    assert tf_bundle.list_variables(prefix)["a/Adagrad"] == (np.dtype("<f4"), (3, 2, 2))
