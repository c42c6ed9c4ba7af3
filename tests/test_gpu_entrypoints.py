"""End to end through the re-hosted reference entry points (sse_index.index -> targetEncodingIndex.tsv ->
sse_evaluator.Evaluator.eval -> sse_demo.DemoSession.query) on a small model directory laid out like the reference's
(vocabulary.txt, modelConfig.param, targetIDs, checkpoint), checked against the oracle (reference sse_index.py:55-97,
sse_evaluator.py:61-114, sse_demo.py:79-127)."""
import json
import os
import shutil

import numpy as np
import pytest

import data_utils
import sse_demo
import sse_evaluator
import sse_index
import sse_model
import sse_oracle as O
import text_encoder

pytestmark = pytest.mark.gpu


@pytest.fixture()
def model_dir(tmp_path, golden_dir, monkeypatch):
    monkeypatch.setenv("SSE_PRECISION", "0")                      # exact fp32 kernels: the file must match the oracle to 2e-5
    d = str(tmp_path / "model")
    os.makedirs(d)
    shutil.copy(os.path.join(golden_dir, "subword_vocab.txt"), os.path.join(d, "vocabulary.txt"))
    enc = data_utils.load_vocabulary(d)
    samples = json.load(open(os.path.join(golden_dir, "subword_samples.json"), encoding="utf-8"))
    texts = [s["text"] if isinstance(s, dict) else s for s in (samples["samples"] if isinstance(samples, dict) else samples)]
    texts = [t.replace("\t", " ").strip() for t in texts if t.strip()]
    rng = np.random.default_rng(0)
    words = sorted({w for t in texts for w in t.split() if w.isalpha()})
    targets = []
    for i in range(120):
        targets.append((" ".join(rng.choice(words, size=int(rng.integers(2, 7)))), "T%03d" % i))
    with open(os.path.join(d, "targetIDs"), "w", encoding="utf-8") as f:
        for t, i in targets:
            f.write("%s\t%s\n" % (t, i))
        f.write("a malformed line without an id\n")                # reference: reported and skipped (sse_index.py:72-74)
    cfg = dict(network_mode="dual-encoder", vocab_size=enc.vocab_size, embedding_size=32, encoding_size=24, src_cell_size=40,
               tgt_cell_size=40, max_seq_length=16, predict_nbest=5, forward_only=True, targetSpaceSize=len(targets),
               learning_rate=0.5, learning_rate_decay_factor=0.99)
    data_utils.save_model_configs(d, cfg)
    p = O.init_params("dual-encoder", enc.vocab_size, 32, 24, 40, 40, seed=4)
    with sse_model.Session() as sess:
        m = sse_model.SSEModel(data_utils.load_model_configs(d))
        sess.run(sse_model.global_variables_initializer())
        m.handle.set_params(p)
        m.saver.save(sess, os.path.join(d, "SSE-LSTM.ckpt"), global_step=7)
        m.handle.close()
    return d, enc, targets, p


def test_index_then_evaluate_then_demo(model_dir):
    d, enc, targets, p = model_dir
    idx_file = os.path.join(d, "targetEncodingIndex.tsv")
    sse_index.index(d, os.path.join(d, "targetIDs"), idx_file, batchsize=50)
    # the file: one row per well-formed target, in targetIDs order, vectors = oracle target encoder (normalised)
    lines = open(idx_file, encoding="utf-8").readlines()
    ids, texts, vecs = O.parse_index_lines(lines)
    assert ids == [i for _t, i in targets] and texts == [t for t, _i in targets]
    T = 16
    ttok = np.array([text_encoder.pad_tokens(enc.encode(t.lower()), T) for t, _ in targets], np.int32)
    want = O.encode(p, "dual-encoder", "tgt", ttok, True)
    assert np.abs(vecs - want).max() < 2e-5
    # every float is numpy's shortest round-trip decimal of a float32
    row0 = lines[0].rstrip("\n").split("\t")[2].split(",")
    assert row0 == [str(np.float32(float(x))) for x in row0]

    # evaluator: queries = noisy copies of some targets, labels = their ids
    rng = np.random.default_rng(1)
    corpus = []
    for j in rng.choice(len(targets), size=40, replace=False):
        words = targets[j][0].split()
        q = " ".join(words[: max(1, len(words) - 1)])
        corpus.append((text_encoder.pad_tokens(enc.encode(q.lower()), T), [targets[j][1]]))
    with sse_model.Session() as sess:
        m = sse_model.SSEModel(data_utils.load_model_configs(d))
        m.saver.restore(sess, sse_model.get_checkpoint_state(d).model_checkpoint_path)
        ev = sse_evaluator.Evaluator(m, corpus, idx_file, sess)
        got = ev.eval(top_n=(1, 3, 10))
        qtok = np.array([c[0] for c in corpus], np.int32)
        qenc = O.encode(p, "dual-encoder", "src", qtok, True)
        labels = [[ids.index(l) for l in c[1]] for c in corpus]
        want_acc = O.evaluator_eval([qenc], vecs, labels, top_n=(1, 3, 10))
        assert np.allclose(got, want_acc, atol=1e-6)
        m.handle.close()

    # demo: un-normalised source against the normalised index (sse_demo.py:121-127)
    ds = sse_demo.DemoSession(d, "targetEncodingIndex.tsv")
    sent = targets[5][0]
    res = ds.query(sent, 5)
    u = O.encode(p, "dual-encoder", "src", np.array([text_encoder.pad_tokens(enc.encode(sent.lower()), T)], np.int32), False)
    sc, ix = O.retrieve(u, vecs, 5)
    assert [r[0] for r in res] == [ids[j] for j in ix[0]]
    assert np.allclose([r[1] for r in res], sc[0], atol=2e-4)
    ds.model.handle.close()
