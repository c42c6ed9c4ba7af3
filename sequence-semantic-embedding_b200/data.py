# coding=utf-8
"""Training-batch sampler (re-host of the reference's data.py).  Output layout of
get_train_batch is the input contract of the train step: rows alternate
(positive pair, 1.0), (random non-positive pair, 0.0) for a contiguous window of positives."""
from __future__ import print_function

import json
import os

import numpy

import data_utils


class Data(object):
    def __init__(self, work_dir, rawdata_dir, rawvocabsize, max_seq_length, seed=None):
        json_path = work_dir + "/compressed"
        self.rng = numpy.random.RandomState(seed) if seed is not None else numpy.random
        if os.path.exists(json_path):
            print("loading saved json data from %s" % json_path)
            with open(json_path, "r") as fin:
                for name, val in json.load(fin).items():
                    setattr(self, name, val)
            self.encoder = data_utils.load_vocabulary(work_dir)
            self.max_seq_length = int(self.max_seq_length)
            self.vocab_size = self.encoder.vocab_size
        else:
            print("generating data from data path: %s" % rawdata_dir)
            encoder, trainCorpus, evalCorpus, encodedFullTargetSpace, tgtIdNameMap = data_utils.prepare_raw_data(
                rawdata_dir, work_dir, rawvocabsize, max_seq_length)
            self.encoder = encoder
            self.rawTrainPosCorpus = trainCorpus
            self.rawEvalCorpus = evalCorpus
            self.max_seq_length = max_seq_length
            self.encodedFullTargetSpace = encodedFullTargetSpace
            self.tgtIdNameMap = tgtIdNameMap
            self.vocab_size = encoder.vocab_size
            self.fullSetTargetIds = list(encodedFullTargetSpace.keys())
            self.rawnegSetLen = len(self.fullSetTargetIds)
            gdict = {k: v for k, v in self.__dict__.items() if k not in ("encoder", "rng") and not callable(v)}
            with open(json_path, "w") as fout:
                json.dump(gdict, fout)
            print("Processed data dumped")
        print("-\nVocab size:", self.vocab_size, "unique words\n-\nMax allowed sequence length:", self.max_seq_length, "\n-")

    def get_train_batch(self, batch_size):
        """Reference data.py:95-115 (including its window quirk: never starts in the first
        batch_size samples, short windows near the end)."""
        num_samples = len(self.rawTrainPosCorpus)
        idx = self.rng.randint(0, num_samples - batch_size) + batch_size
        source_inputs, tgt_inputs, labels = [], [], []
        for source_tokens, verifiedTgtIds in self.rawTrainPosCorpus[idx:idx + batch_size]:
            curPos = verifiedTgtIds[self.rng.randint(0, len(verifiedTgtIds))]
            posSet = set(verifiedTgtIds)
            source_inputs.append(source_tokens)
            tgt_inputs.append(self.encodedFullTargetSpace[curPos])
            labels.append(1.0)
            neg = self.fullSetTargetIds[self.rng.randint(0, self.rawnegSetLen)]
            while neg in posSet:
                neg = self.fullSetTargetIds[self.rng.randint(0, self.rawnegSetLen)]
            source_inputs.append(source_tokens)
            tgt_inputs.append(self.encodedFullTargetSpace[neg])
            labels.append(0.0)
        return source_inputs, tgt_inputs, labels

    # ------------------------------------------------------------------ vectorised sampler (host side)
    def _build_arrays(self):
        """Corpus as arrays: source rows [P,T], verified-target lists in CSR form over target ROW numbers, target rows
        [N,T].  Built once; the python objects stay the source of truth for the reference-shaped API above."""
        tid_row = {tid: j for j, tid in enumerate(self.fullSetTargetIds)}
        self._tgt_rows = numpy.array([self.encodedFullTargetSpace[t] for t in self.fullSetTargetIds], dtype=numpy.int32)
        self._src_rows = numpy.array([src for src, _ in self.rawTrainPosCorpus], dtype=numpy.int32)
        counts = numpy.array([len(v) for _, v in self.rawTrainPosCorpus], dtype=numpy.int64)
        self._ver_off = numpy.concatenate([[0], numpy.cumsum(counts)])
        self._ver_rows = numpy.array([tid_row[t] for _, v in self.rawTrainPosCorpus for t in v], dtype=numpy.int64)
        # membership test "target row j is verified for positive i" as a sorted key array (i * N + j)
        owner = numpy.repeat(numpy.arange(len(counts), dtype=numpy.int64), counts)
        self._ver_keys = numpy.sort(owner * len(self.fullSetTargetIds) + self._ver_rows)

    def get_train_batch_arrays(self, batch_size):
        """Same sampling rule and row layout as get_train_batch (reference data.py:95-115) -- window of consecutive
        positives that never starts in the first batch_size samples, one uniformly chosen verified target per positive,
        one uniformly chosen NON-verified target as the negative (rejection), rows alternating (pos, 1.0), (neg, 0.0)
        -- drawn with array operations: returns int32 [2B,T], int32 [2B,T], float32 [2B].  The random stream differs
        from the python loop's (the reference's is unseeded); the distribution is the same."""
        if not hasattr(self, "_src_rows"):
            self._build_arrays()
        rng = self.rng
        P, N = self._src_rows.shape[0], self._tgt_rows.shape[0]
        start = int(rng.randint(0, P - batch_size)) + batch_size
        rows = numpy.arange(start, min(start + batch_size, P), dtype=numpy.int64)
        B = rows.shape[0]
        cnt = self._ver_off[rows + 1] - self._ver_off[rows]
        pos = self._ver_rows[self._ver_off[rows] + (rng.random_sample(B) * cnt).astype(numpy.int64)]
        neg = rng.randint(0, N, size=B).astype(numpy.int64)
        keys = self._ver_keys
        while True:
            k = rows * N + neg
            j = numpy.searchsorted(keys, k)
            clash = (j < keys.shape[0]) & (keys[numpy.minimum(j, keys.shape[0] - 1)] == k)
            if not clash.any():
                break
            neg[clash] = rng.randint(0, N, size=int(clash.sum()))
        src = numpy.repeat(self._src_rows[rows], 2, axis=0)
        tgt = numpy.empty((2 * B, self._tgt_rows.shape[1]), numpy.int32)
        tgt[0::2] = self._tgt_rows[pos]
        tgt[1::2] = self._tgt_rows[neg]
        labels = numpy.tile(numpy.array([1.0, 0.0], numpy.float32), B)
        return src, tgt, labels

    def device_sampler(self, handle, seed=0):
        """Device-side sampler: the corpus arrays go to the GPU once; every call of the returned object's
        next_batch() is ONE kernel writing the step's pair rows into device buffers that sse_train_step consumes."""
        if not hasattr(self, "_src_rows"):
            self._build_arrays()
        return DeviceSampler(handle, self._src_rows, self._ver_off, self._ver_rows, self._tgt_rows, self.rng, seed)

    def get_test_batch(self, batch_size):
        num_samples = len(self.rawEvalCorpus)
        idx = self.rng.randint(0, num_samples - batch_size) + batch_size
        return self.rawEvalCorpus[idx:idx + batch_size]


class DeviceSampler(object):
    """Train batches drawn on the GPU (csrc/tok_prep.cu: sample_train_batch_kernel).  Same rule and row layout as
    Data.get_train_batch (reference data.py:95-115); the window start is drawn on the host (one integer), positives /
    negatives by a counter-based hash of (seed, step, positive) on the device."""

    def __init__(self, handle, src_rows, ver_off, ver_rows, tgt_rows, rng, seed=0):
        import torch
        self.h, self.rng, self.seed, self.step = handle, rng, int(seed), 0
        self.P, self.T = int(src_rows.shape[0]), int(src_rows.shape[1])
        handle.sampler_set(src_rows, ver_off, ver_rows, tgt_rows)
        self._torch = torch
        self._buf = {}

    def next_batch(self, batch_size, start=None, stream=None):
        """-> (src int32 [2B,T], tgt int32 [2B,T], labels float32 [2B]) device tensors (views of reused buffers)"""
        torch = self._torch
        if batch_size not in self._buf:
            dev = torch.device("cuda", int(self.h.cfg.device))
            self._buf[batch_size] = (torch.empty(2 * batch_size, self.T, dtype=torch.int32, device=dev),
                                     torch.empty(2 * batch_size, self.T, dtype=torch.int32, device=dev),
                                     torch.empty(2 * batch_size, dtype=torch.float32, device=dev))
        src, tgt, lab = self._buf[batch_size]
        if start is None:        # reference window rule: never starts in the first batch_size samples, short windows near the end
            start = int(self.rng.randint(0, self.P - batch_size)) + batch_size
        n = self.h.sampler_batch(start, batch_size, self.seed, self.step, src, tgt, lab, stream)
        self.step += 1
        return src[:n], tgt[:n], lab[:n]
