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
