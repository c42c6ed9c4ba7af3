"""Index-file I/O measurement (host only): native writer / reader (csrc/tsv_io.cpp) next to the reference's python loops
(restated inline).  usage: tsv_bench.py [rows] [E]"""
import json, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "sequence-semantic-embedding_b200"))
import sse_ffi


def py_format_row(tgt_id, text, enc_row):            # the reference's writer loop body (sse_index.py:93-97)
    return tgt_id + "\t" + text + "\t" + ",".join([str(n) for n in enc_row]) + "\n"


def py_parse_lines(lines):                            # the reference's reader loop (sse_evaluator.py:79-88)
    ids, encs = [], []
    for line in lines:
        info = line.strip().split("\t")
        if len(info) != 3:
            continue
        ids.append(info[0]); encs.append([float(f) for f in info[2].strip().split(",")])
    return ids, np.array(encs)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
E = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rng = np.random.default_rng(0)
enc = rng.standard_normal((N, E)).astype(np.float32); enc /= np.linalg.norm(enc, axis=1, keepdims=True)
ids = ["id%d" % i for i in range(N)]; texts = ["target text number %d" % i for i in range(N)]
p = ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp") + "/sse_idx_bench.tsv"
t = time.perf_counter(); sse_ffi.tsv_write_index(p, ids, texts, enc); tw = time.perf_counter() - t
t = time.perf_counter(); gi, gt, ge, _ = sse_ffi.tsv_read_index(p); tr = time.perf_counter() - t
assert np.array_equal(ge.view(np.uint32), enc.view(np.uint32)) and gi == ids
n2 = min(N, 5000)
t = time.perf_counter(); s = "".join(py_format_row(i, tx, r) for i, tx, r in zip(ids[:n2], texts[:n2], enc[:n2])); tpw = time.perf_counter() - t
lines = s.splitlines(True)
t = time.perf_counter(); py_parse_lines(lines); tpr = time.perf_counter() - t
# batched tokenizer + padder (csrc/subword_tok.cpp) vs the python encoder, product-title-like sentences
import text_encoder
enc_tok = text_encoder.SubwordTextEncoder(os.path.join(REPO, "tests", "golden", "subword_vocab.txt"))
words = ["alpha", "beta", "gamma", "delta", "shoes", "women", "iphone", "case", "black", "2019", "new", "size", "xl", "usb-c", "naïve", "日本"]
sent = [" ".join(rng.choice(words, size=int(rng.integers(3, 12)))) for _ in range(200000)]
t = time.perf_counter(); rows_tok, _ = enc_tok.encode_batch(sent, 50); tt = time.perf_counter() - t
t = time.perf_counter(); want = [text_encoder.pad_tokens(enc_tok.encode(x), 50) for x in sent[:5000]]; tp = time.perf_counter() - t
assert rows_tok[:5000].tolist() == want
tok = {"sentences": len(sent), "native_sentences_per_s": len(sent) / tt, "python_sentences_per_s": 5000 / tp}
print(json.dumps({"tokenizer": tok, "rows": N, "E": E, "file_MB": os.path.getsize(p) / 1e6, "host_threads": os.cpu_count(),
                  "native_write_rows_per_s": N / tw, "native_read_rows_per_s": N / tr,
                  "python_write_rows_per_s": n2 / tpw, "python_read_rows_per_s": n2 / tpr, "round_trip": "bit-exact"}))
os.remove(p)
