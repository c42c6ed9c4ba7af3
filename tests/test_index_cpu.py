"""sse_index.createIndexFile on the CPU with the model replaced by a recorder: batching, the skipped malformed lines,
the token rows handed to the encoder (native batched tokenizer == python encoder + reference pad rule) and the bytes of
the index file (native writer == reference row format)."""
import os

import numpy as np

import sse_index
import sse_oracle as O
import text_encoder


class FakeModel(object):
    norm_tgt_seq_embedding = "norm_tgt"

    def __init__(self):
        self.fed = []

    def get_target_encoding_feed_dict(self, rows):
        rows = np.array(rows, dtype=np.int32)
        self.fed.append(rows)
        return {"tgt": rows}


class FakeSession(object):
    def run(self, fetches, feed_dict=None):
        rows = feed_dict["tgt"]
        rng = np.random.default_rng(int(rows.sum()) % 1000)
        e = rng.standard_normal((rows.shape[0], 8)).astype(np.float32)
        self.last = e
        return [e]


def test_create_index_file_flow(tmp_path, golden_dir, capsys):
    enc = text_encoder.SubwordTextEncoder(os.path.join(golden_dir, "subword_vocab.txt"))
    raw = tmp_path / "targetIDs"
    lines = ["Vacation Settings\tid1\n", "how do I FILTER listings?\tid2\n", "broken line without id\n",
             "a very long target " + "word " * 40 + "\tid3\n", "naïve café – 日本語\tid4\n", "last one\tid5\n"]
    raw.write_text("".join(lines), encoding="utf-8")
    out = str(tmp_path / "idx.tsv")
    model, sess = FakeModel(), FakeSession()
    T = 12
    sse_index.createIndexFile(model, enc, str(raw), T, out, sess, batchsize=4)
    printed = capsys.readouterr().out
    assert "Missing field with error line" in printed and "Error Detected!!!" in printed and "total count:6" in printed
    # two batches: the first holds lines 0-3 minus the malformed one, the second the rest
    assert [f.shape for f in model.fed] == [(3, T), (2, T)]
    good = [l.rstrip("\n").split("\t") for l in lines if len(l.strip().split("\t")) == 2]
    want_rows = np.array([text_encoder.pad_tokens(enc.encode(t.lower()), T) for t, _ in good], np.int32)
    assert np.array_equal(np.vstack(model.fed), want_rows)
    ids, texts, vecs = O.parse_index_lines(open(out, encoding="utf-8").readlines())
    assert ids == [i for _t, i in good] and texts == [t for t, _i in good]
    assert vecs.shape == (5, 8)
    # the last batch's rows are the last encodings the "model" produced, formatted like the reference formats them
    tail = open(out, encoding="utf-8").readlines()[-2:]
    assert tail == [O.format_index_row(i, t, r) for (t, i), r in zip(good[-2:], sess.last)]


def test_evaluator_host_logic_with_a_numpy_handle(tmp_path):
    """sse_evaluator.Evaluator (reference sse_evaluator.py:61-114) with the device calls replaced by numpy: index file
    parsing incl. a malformed row and a duplicated id, label mapping, 600-row batching, top-n accuracies."""
    import sse_evaluator
    import sse_ffi

    rng = np.random.default_rng(0)
    N, E, T = 700, 16, 5
    tgt = rng.standard_normal((N, E)).astype(np.float32)
    tgt /= np.linalg.norm(tgt, axis=1, keepdims=True)
    ids = ["t%d" % i for i in range(N)]
    ids[11] = ids[3]                                               # duplicated id: the later row wins the label map (reference dict semantics)
    p = str(tmp_path / "idx.tsv")
    sse_ffi.tsv_write_index(p, ids, ["text %d" % i for i in range(N)], tgt)
    with open(p, "a", encoding="utf-8") as f:
        f.write("broken row\n")

    class Handle(object):
        def index_set(self, enc, global_offset=0):
            self.enc = np.asarray(enc, np.float32)

        def query_host(self, toks, k, normalize=True):
            q = tgt[toks[:, 0]] + 0.05 * np.random.default_rng(1).standard_normal((len(toks), E)).astype(np.float32)
            self.calls = getattr(self, "calls", 0) + 1
            s, i = O.top_k_tf(q.astype(np.float64) @ self.enc.astype(np.float64).T, k, normalize_scores=False)
            return s.astype(np.float32), i.astype(np.int32)

    class Model(object):
        handle = Handle()
        def set_forward_only(self, f): pass

    corpus = [([int(j)] * T, [ids[j]]) for j in rng.integers(0, N, size=1300)]
    ev = sse_evaluator.Evaluator(Model(), corpus, p, session=None)
    assert len(ev.targetIDs) == N and ev.idLabelMap[ids[3]] == 11 and ev.targetEncodings.dtype == np.float64
    acc = ev.eval(top_n=(1, 3, 10))
    assert Model.handle.calls == 3                                 # 1300 queries in batches of 600
    assert 0.9 < acc[0] <= acc[1] <= acc[2] <= 1.0
