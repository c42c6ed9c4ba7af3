# coding=utf-8
"""Interactive n-best demo (re-host of reference sse_demo.py:59-138).  The query is encoded with
the UN-normalised source embedding (sse_demo.py:123) and scored against the normalised index;
scoring + ranking is the fused GPU top-k instead of np.dot + full sort."""
from __future__ import print_function

import argparse
import codecs
import os
import sys

import numpy as np

import data_utils
import sse_ffi
import sse_model
import text_encoder


def load_index(path):
    """id, text, encoding rows of targetEncodingIndex.tsv (sse_demo.py:79-90)."""
    ids, texts, encs, skipped = sse_ffi.tsv_read_index(path)      # native reader (csrc/tsv_io.cpp), same row rules
    if skipped:
        print("Error in targetIndexFile! %d malformed line(s) skipped" % skipped)
    return ids, dict(zip(ids, texts)), encs


class DemoSession(object):
    def __init__(self, model_dir, indexFile):
        if not os.path.exists(os.path.join(model_dir, indexFile)):
            raise IOError("Index file does not exist!!!")
        self.encoder = data_utils.load_vocabulary(model_dir)
        self.targetIDs, self.names, encs = load_index(os.path.join(model_dir, indexFile))
        self.cfg = data_utils.load_model_configs(model_dir)
        self.sess = sse_model.Session()
        self.model = sse_model.SSEModel(self.cfg)
        ckpt = sse_model.get_checkpoint_state(model_dir)
        if not ckpt:
            raise IOError("Error!!!Could not load any model from specified folder: %s" % model_dir)
        print("Reading model parameters from %s" % ckpt.model_checkpoint_path)
        self.model.saver.restore(self.sess, ckpt.model_checkpoint_path)
        self.model.handle.index_set(encs, global_offset=0)
        self.T = int(self.cfg["max_seq_length"])

    def batcher(self, max_batch=64, max_wait_ms=2.0):
        """A micro-batcher over this session's device index for concurrent callers (web-server threads): requests
        arriving within max_wait_ms share one encode + scan (sse_serve.MicroBatcher)."""
        import sse_serve
        return sse_serve.MicroBatcher(lambda toks, k, normalize: self.model.handle.query_host(toks, k, normalize=normalize),
                                      max_batch=max_batch, max_wait_ms=max_wait_ms)

    def tokens(self, sentence):
        ids = self.encoder.encode(sentence.lower())
        return np.array(text_encoder.pad_tokens(ids, self.T), dtype=np.int32)

    def query(self, sentence, nbest, normalize=False):
        ids = self.encoder.encode(sentence.lower())
        if len(ids) > self.T - 2:
            print("Input sentence too long, max allowed is %d. Try to increase limit!!!!" % self.T)
        toks = np.array([text_encoder.pad_tokens(ids, self.T)], dtype=np.int32)
        k = min(nbest, len(self.targetIDs), 128)
        scores, idx = self.model.handle.query_host(toks, k, normalize=normalize)
        return [(self.targetIDs[j], float(s), self.names[self.targetIDs[j]]) for s, j in zip(scores[0], idx[0])]


def demo(nbest, model_dir, indexFile):
    d = DemoSession(model_dir, indexFile)
    sys.stdout.write("\n\nPlease type some keywords to get related task results.\nType 'exit' to quit demo.\n > ")
    sys.stdout.flush()
    sentence = sys.stdin.readline()
    while sentence and sentence.strip().lower() != "exit":
        res = d.query(sentence.strip(), nbest)
        print("Top %s Prediction results are:\n" % nbest)
        for i, (tid, conf, name) in enumerate(res):
            print("top%d:  %s , %f ,  %s " % (i + 1, tid, conf, name))
        print("> ", end="")
        sys.stdout.flush()
        sentence = sys.stdin.readline()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("nbest", type=int, nargs="?", default=10)
    ap.add_argument("--model_dir", default="models-classification")
    ap.add_argument("--indexFile", default="targetEncodingIndex.tsv")
    a = ap.parse_args()
    demo(a.nbest, a.model_dir, a.indexFile)
