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

    def get_test_batch(self, batch_size):
        num_samples = len(self.rawEvalCorpus)
        idx = self.rng.randint(0, num_samples - batch_size) + batch_size
        return self.rawEvalCorpus[idx:idx + batch_size]
