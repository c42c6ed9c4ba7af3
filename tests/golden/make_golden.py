"""Generate golden fixtures by running the REAL reference code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference's TF-free modules are imported through oracle/tf_stub; their
outputs on seeded inputs are frozen into tests/golden/*.npz / *.tsv.  The GPU
box has no /root/reference: tests only read the committed fixtures.

Covers the retrieval half of the hot path (the only half runnable without
TensorFlow): data_utils.getSortedResults, computeTopK_TightVersion_accuracy,
computeTopK_accuracy (data_utils.py:263-304), the targetEncodingIndex.tsv row
format written by sse_index.createIndexFile (sse_index.py:93-95) and parsed by
sse_evaluator.Evaluator.__init__ (sse_evaluator.py:80-92), and the pad rule of
data_utils.gen_postive_corpus (data_utils.py:147-155) via text_encoder ids.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle", "tf_stub"))
sys.path.insert(0, "/root/reference")

import data_utils  # noqa: E402  (the real reference module)
import text_encoder  # noqa: E402


def ranking_fixture():
    rng = np.random.default_rng(7)
    Q, N, E, k = 24, 3000, 64, 20
    tgt = rng.standard_normal((N, E)).astype(np.float32)
    tgt /= np.linalg.norm(tgt, axis=1, keepdims=True)
    src = rng.standard_normal((Q, E)).astype(np.float32)
    src /= np.linalg.norm(src, axis=1, keepdims=True)
    # planted near-duplicates so top-1 has a wide margin
    planted = rng.integers(0, N, size=Q)
    src[: Q // 2] = tgt[planted[: Q // 2]] + 0.05 * rng.standard_normal((Q // 2, E)).astype(np.float32)
    src /= np.linalg.norm(src, axis=1, keepdims=True)
    tgt64 = tgt.astype(np.float64)          # Evaluator parses the TSV into float64 (sse_evaluator.py:87,92)
    distances = np.dot(src, tgt64.T)        # sse_evaluator.py:110
    sortedScore, rankedIdx = data_utils.getSortedResults(distances)   # data_utils.py:263-267
    labels = [[int(planted[i])] + [int(x) for x in rng.integers(0, N, size=int(rng.integers(0, 3)))] for i in range(Q)]
    acc = {}
    for n in (1, 3, 10):
        acc["tight%d" % n] = data_utils.computeTopK_TightVersion_accuracy(n, labels, rankedIdx)
        acc["any%d" % n] = data_utils.computeTopK_accuracy(n, labels, rankedIdx)
    lab_flat = np.full((Q, 3), -1, np.int64)
    for i, l in enumerate(labels):
        lab_flat[i, : len(l)] = l
    np.savez_compressed(os.path.join(HERE, "ranking.npz"), src=src, tgt=tgt, top_scores=sortedScore[:, :k],
                        top_idx=rankedIdx[:, :k], labels=lab_flat,
                        acc=np.array([acc["tight1"], acc["tight3"], acc["tight10"], acc["any1"], acc["any3"], acc["any10"]]))
    print("ranking.npz:", acc)


def index_format_fixture():
    """Rows exactly as sse_index.createIndexFile writes them (sse_index.py:93-95) and the
    float64 matrix Evaluator.__init__ parses back (sse_evaluator.py:80-92)."""
    rng = np.random.default_rng(11)
    E = 8
    enc = rng.standard_normal((5, E)).astype(np.float32)
    enc /= np.linalg.norm(enc, axis=1, keepdims=True)
    enc[0, 0] = np.float32(1e-8); enc[1, 1] = np.float32(-0.5); enc[2, 2] = np.float32(1.0)
    ids = ["id%d" % i for i in range(5)]
    texts = ["some target text %d" % i for i in range(5)]
    path = os.path.join(HERE, "index_rows.tsv")
    with open(path, "w", encoding="utf-8") as f:
        for i in range(5):
            f.write(ids[i] + "\t" + texts[i] + "\t" + ",".join([str(n) for n in enc[i]]) + "\n")   # sse_index.py:93-95
    parsed = []
    for line in open(path, "r", encoding="utf-8").readlines():   # sse_evaluator.py:80-87
        info = line.strip().split("\t")
        parsed.append([float(f) for f in info[2].strip().split(",")])
    parsed = np.array(parsed)
    np.savez_compressed(os.path.join(HERE, "index_rows.npz"), enc=enc, parsed=parsed)
    print("index_rows: float32 round trip exact:", bool(np.all(parsed.astype(np.float32) == enc)))


def padding_fixture():
    """pad / truncate rule applied by the reference to token id lists (data_utils.py:147-155)."""
    T = 12
    cases = [[], [5], [5, 6, 7], list(range(2, 12)), list(range(2, 13)), list(range(2, 30))]
    out = []
    for ids in cases:
        seqlen = len(ids)
        if seqlen > T - 2:
            toks = [text_encoder.PAD_ID] + ids[: T - 2] + [text_encoder.EOS_ID]
        else:
            toks = [text_encoder.PAD_ID] * (T - seqlen - 1) + ids + [text_encoder.EOS_ID]
        out.append(toks)
    raw = np.full((len(cases), 32), -1, np.int32)
    for i, c in enumerate(cases):
        raw[i, : len(c)] = c
    np.savez_compressed(os.path.join(HERE, "padding.npz"), T=T, raw=raw, padded=np.array(out, dtype=np.int32))
    print("padding.npz written; PAD/EOS =", text_encoder.PAD_ID, text_encoder.EOS_ID)




def subword_fixture():
    """Build a small vocabulary with the REAL reference builder on the qna data set and freeze
    (vocabulary file, sample lines, ids) so the repo's own encoder can be checked against it."""
    import tarfile
    import tempfile
    import tokenizer
    tmp = tempfile.mkdtemp()
    with tarfile.open("/root/reference/rawdata-qna/DataSet.tar.gz") as t:
        t.extractall(tmp)
    root = tmp
    for dp, dn, fn in os.walk(tmp):
        if "targetIDs" in fn:
            root = dp
    lines = []
    for name in ("TrainPairs", "targetIDs"):
        for line in open(os.path.join(root, name), encoding="utf-8"):
            lines.append(line.strip().split("\t")[0].lower())
    corpus = os.path.join(tmp, "c.Corpus")
    with open(corpus, "w", encoding="utf-8") as f:
        f.write("\n".join(lines))
    counts = tokenizer.corpus_token_counts(corpus, 2000000, split_on_newlines=True)
    enc = text_encoder.SubwordTextEncoder.build_to_target_size(600, counts, 2, 1000)
    vocab_path = os.path.join(HERE, "subword_vocab.txt")
    enc.store_to_file(vocab_path)
    sample = lines[:40] + ["Dude - that's so cool.", "snow_man \\ x_y 3.14 é中 zzzqqq", "", " ", "a  b"]
    ids = [enc.encode(s) for s in sample]
    width = max(len(i) for i in ids)
    arr = np.full((len(ids), width), -1, np.int32)
    for r, i in enumerate(ids):
        arr[r, : len(i)] = i
    import json
    with open(os.path.join(HERE, "subword_samples.json"), "w", encoding="utf-8") as f:
        json.dump(sample, f, ensure_ascii=False)
    np.savez_compressed(os.path.join(HERE, "subword.npz"), ids=arr, vocab_size=enc.vocab_size)
    print("subword fixture: vocab", enc.vocab_size, "samples", len(sample))


if __name__ == "__main__":
    ranking_fixture()
    index_format_fixture()
    padding_fixture()
    subword_fixture()
