# coding=utf-8
"""Top-n accuracy evaluator (re-host of reference sse_evaluator.py:52-114).

Same constructor and ``eval(top_n)`` contract.  Differences, all result-preserving:
  * the parsed index is registered once with the GPU (fp32; the values in the TSV are exact
    float32 decimals, so float64 parsing adds nothing) instead of a float64 numpy matrix;
  * per 600-row batch the sources are encoded and searched with the fused top-k
    (k = max(top_n)) instead of np.dot + a full argsort of [600, N];
  * the three redundant passes (one per n in top_n, reference :103-113) collapse into one."""
from __future__ import print_function

import codecs
import math

import numpy as np

import data_utils
import sse_ffi


class Evaluator(object):
    def __init__(self, model, eval_corpus, tgtIndexFile, session):
        self.model = model
        self.srcSeq_batch = [entry[0] for entry in eval_corpus]
        self.session = session
        # reference sse_evaluator.py:79-88: strip, split on tabs, rows without three fields are reported and skipped.
        # Parsed by the native reader (csrc/tsv_io.cpp): float32, bit-exact inverse of the writer.
        self.targetIDs, _texts, enc32, skipped = sse_ffi.tsv_read_index(tgtIndexFile)
        if skipped:
            print("Error in targetIndexFile! %d malformed line(s) skipped" % skipped)
        self.idLabelMap = {tgtid: idx for idx, tgtid in enumerate(self.targetIDs)}
        self.eval_Labels = [[self.idLabelMap[tgtid] for tgtid in entry[1]] for entry in eval_corpus]
        self.targetEncodings = enc32.astype(np.float64)             # float64 [N,E], the dtype the reference keeps
        self.model.handle.index_set(enc32, global_offset=0)

    def eval(self, top_n=(1, 3, 10)):
        acc = [[] for _ in top_n]
        self.model.set_forward_only(True)
        batchSize = 600
        kmax = min(max(top_n), len(self.targetIDs), 128)
        for batchId in range(math.ceil(len(self.srcSeq_batch) / batchSize)):
            rows = self.srcSeq_batch[batchId * batchSize:(batchId + 1) * batchSize]
            toks = np.array(rows, dtype=np.int32)
            _scores, rankedIdx = self.model.handle.query_host(toks, kmax, normalize=True)
            labels = self.eval_Labels[batchId * batchSize:(batchId + 1) * batchSize]
            for j, n in enumerate(top_n):
                acc[j].append(data_utils.computeTopK_TightVersion_accuracy(n, labels, rankedIdx))
        return [float(np.mean(a)) for a in acc]
