"""Host logic of the Saver stand-in (reference sse_model.py:138 tf.train.Saver(max_to_keep=20); call sites
sse_train.py:205,212,232): the max_to_keep window holds ONE entry per checkpoint name, as tf.train.Saver's does."""
import os

import numpy as np

import sse_model


class _StubHandle(object):
    def param_names(self):
        return [("word_embedding", (3, 2))]

    def get_param(self, name):
        return np.zeros((3, 2), np.float32)

    def scalars(self):
        return 0.5, 7


class _StubModel(object):
    handle = _StubHandle()


def test_repeated_best_ever_saves_do_not_evict_themselves(tmp_path):
    saver = sse_model.Saver(_StubModel(), max_to_keep=20)
    base = str(tmp_path / "SSE-LSTM.ckpt")
    for _ in range(15):
        saver.save(None, base + "-BestEver")
    for e in range(6):
        saver.save(None, base + "-epoch-%d" % e)
    assert os.path.exists(base + "-BestEver.npz")                    # the best model is still on disk ...
    listed = [l.split('"')[1] for l in open(tmp_path / "checkpoint") if l.startswith("all_model_checkpoint_paths")]
    assert listed.count("SSE-LSTM.ckpt-BestEver") == 1              # ... listed once ...
    for name in listed:                                              # ... and everything listed exists
        assert os.path.exists(str(tmp_path / name) + ".npz")
    assert len(listed) == 7


def test_window_rotation_deletes_only_unreferenced_files(tmp_path):
    saver = sse_model.Saver(_StubModel(), max_to_keep=3)
    base = str(tmp_path / "m.ckpt")
    for e in range(5):
        saver.save(None, base, global_step=e)
        saver.save(None, base + "-BestEver")
    kept = sorted(f for f in os.listdir(tmp_path) if f.endswith(".npz"))
    assert kept == ["m.ckpt-3.npz", "m.ckpt-4.npz", "m.ckpt-BestEver.npz"]
    assert sse_model.get_checkpoint_state(str(tmp_path)).model_checkpoint_path.endswith("m.ckpt-BestEver")


def test_tf_bundle_checkpoint_format(tmp_path, monkeypatch):
    import tf_bundle
    monkeypatch.setenv("SSE_CHECKPOINT_FORMAT", "tf")
    saver = sse_model.Saver(_StubModel(), max_to_keep=2)
    base = str(tmp_path / "SSE-LSTM.ckpt")
    for e in range(3):
        saver.save(None, base, global_step=e)
    files = sorted(os.listdir(tmp_path))
    assert files == ["SSE-LSTM.ckpt-1.data-00000-of-00001", "SSE-LSTM.ckpt-1.index", "SSE-LSTM.ckpt-2.data-00000-of-00001", "SSE-LSTM.ckpt-2.index", "checkpoint"]
    got = tf_bundle.read_bundle(base + "-2", verify_crc=True)
    assert got["word_embedding"].shape == (3, 2) and int(got["global_step"]) == 7 and abs(float(got["learning_rate"]) - 0.5) < 1e-7
    assert sse_model.get_checkpoint_state(str(tmp_path)).model_checkpoint_path.endswith("SSE-LSTM.ckpt-2")
