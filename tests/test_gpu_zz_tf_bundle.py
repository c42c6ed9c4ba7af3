"""Loading a TensorFlow V2 checkpoint bundle into the model (Saver.restore on a `<prefix>.index` + `.data-*` pair):
variables by their TF names, the TF<=1.1 LSTM names, scalars, and the error for an incomplete checkpoint.  The bundle
is produced by tf_bundle.write_bundle (the format itself is unpinned against the real library, see tf_bundle.py)."""
import os

import numpy as np
import pytest

import sse_ffi
import sse_model
import sse_oracle as O
import tf_bundle

pytestmark = pytest.mark.gpu

CFG = dict(network_mode="dual-encoder", vocab_size=300, embedding_size=16, encoding_size=12, src_cell_size=24, tgt_cell_size=24,
           max_seq_length=10, predict_nbest=3, forward_only=True, targetSpaceSize=10, learning_rate=0.5, learning_rate_decay_factor=0.9)


def test_restore_from_tf_bundle(tmp_path):
    p = O.init_params("dual-encoder", 300, 16, 12, 24, 24, seed=9)
    tensors = {k: np.asarray(v, np.float32) for k, v in p.items()}
    # the reference's Saver also stores these; old TF spelled the LSTM variables weights / biases
    old_style = {k.replace("basic_lstm_cell/kernel", "basic_lstm_cell/weights").replace("basic_lstm_cell/bias", "basic_lstm_cell/biases"): v
                 for k, v in tensors.items()}
    old_style["global_step"] = np.array(4321, np.int64)
    old_style["learning_rate"] = np.array(0.125, np.float32)
    old_style["beta1_power"] = np.array(0.9, np.float32)                 # a variable the model does not have: ignored
    prefix = str(tmp_path / "SSE-LSTM.ckpt-4321")
    tf_bundle.write_bundle(prefix, old_style)
    with open(str(tmp_path / "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "SSE-LSTM.ckpt-4321"\n')
    ck = sse_model.get_checkpoint_state(str(tmp_path))
    assert ck is not None and ck.model_checkpoint_path == prefix
    with sse_model.Session() as sess:
        m = sse_model.SSEModel(dict(CFG))
        m.saver.restore(sess, ck.model_checkpoint_path)
        for name, want in tensors.items():
            assert np.array_equal(m.handle.get_param(name), want), name
        lr, gs = m.handle.scalars()
        assert abs(lr - 0.125) < 1e-7 and gs == 4321
        # and the restored model computes what the oracle computes with these weights
        rng = np.random.default_rng(0)
        tok = O.synth_tokens(rng, 9, 10, 300, "real", 3.0)
        got = m.handle.encode_host(sse_ffi.SIDE_SRC, tok, True)
        assert np.abs(got - O.encode(p, "dual-encoder", "src", tok, True)).max() < 1e-3
        m.handle.close()
    # an incomplete checkpoint is refused
    del old_style["word_embedding"]
    tf_bundle.write_bundle(prefix, old_style)
    with sse_model.Session() as sess:
        m = sse_model.SSEModel(dict(CFG))
        with pytest.raises(sse_ffi.SseError):
            m.saver.restore(sess, prefix)
        m.handle.close()
