# coding=utf-8
"""Host-side helpers of the SSE entry points (B200 re-host).

Same function names / file formats as the reference's data_utils.py so
model directories are interchangeable: ``modelConfig.param`` (key=value), ``targetIDs``
(text \\t id), ``TrainPairs`` / ``EvalPairs`` (text \\t id1|id2...), ``encoded.FullTargetSpace``
(id \\t text \\t comma-separated ids), ``vocabulary.txt``.  Ranking metrics restate
data_utils.py:263-304 but take the top-k ids straight from the fused GPU top-k instead of a
full argsort of a [Q,N] matrix.
"""
from __future__ import annotations

import codecs
import os
import sys
import tarfile

import numpy as np

import text_encoder


def get_data_set(rawDir, processedDir):
    """Unpack <rawDir>/DataSet.tar.gz into processedDir once (reference data_utils.py:82-112)."""
    need = [os.path.join(processedDir, n) for n in ("TrainPairs", "EvalPairs", "targetIDs")]
    if all(os.path.exists(p) for p in need):
        return
    tar = os.path.join(rawDir, "DataSet.tar.gz")
    if not os.path.exists(tar):
        raise ValueError("Could not find %s" % tar)
    os.makedirs(processedDir, exist_ok=True)
    with tarfile.open(tar) as t:
        for m in t.getmembers():
            base = os.path.basename(m.name)
            if m.isfile() and base in ("TrainPairs", "EvalPairs", "targetIDs", "vocabulary.txt"):
                m.name = base
                t.extract(m, processedDir)


def _pad(ids, max_seq_length, what, text):
    if len(ids) > max_seq_length - 2:
        print('Warning: %s:\n %s \n Its seq length is:%d,  which is longer than MAX_SEQ_LENTH of %d. Try to increase limit!!!!'
              % (what, text, len(ids), max_seq_length))
    return text_encoder.pad_tokens(ids, max_seq_length)


def gen_postive_corpus(pairfilename, encodedTargetSpace, encoder, max_seq_length):
    """[(source_tokens, verifiedTgtIds)]  (reference data_utils.py:115-157)."""
    corpus = []
    known = set(encodedTargetSpace.keys())
    for line in codecs.open(pairfilename, "r", "utf-8"):
        info = line.strip().split("\t")
        if len(info) != 2:
            print("File %s has Bad line of training data:\n %s" % (pairfilename, line))
            continue
        srcSeq, tgtIds = info
        verified = [t for t in tgtIds.split("|") if t in known]
        if len(verified) != len(tgtIds.split("|")):
            print("Warning! trouble in finding targetID in target Space file!! %s" % line)
        if not verified:
            print("Not found any verified tgtIDs in line:%s" % line)
            continue
        corpus.append((_pad(encoder.encode(srcSeq.lower()), max_seq_length, "Source Seq", srcSeq), verified))
    return corpus


def load_vocabulary(processed_data_dir):
    vocabFile = os.path.join(processed_data_dir, "vocabulary.txt")
    if not os.path.exists(vocabFile):
        raise ValueError(
            "Error!! Could not find vocabulary.txt in %s. Building a subword vocabulary from a corpus is outside "
            "the B200 hot path: build it once with the reference's text_encoder_build_subword.py (or copy the one "
            "from a reference model directory); the format is unchanged." % processed_data_dir)
    return text_encoder.SubwordTextEncoder(filename=vocabFile)


def prepare_raw_data(raw_data_dir, processed_data_dir, vocabulary_size, max_seq_length):
    """Reference data_utils.py:160-213 (vocabulary must already exist, see load_vocabulary)."""
    get_data_set(raw_data_dir, processed_data_dir)
    encoder = load_vocabulary(processed_data_dir)
    encodedFullTargetSpace, tgtIdNameMap = {}, {}
    with codecs.open(os.path.join(processed_data_dir, "encoded.FullTargetSpace"), "w", "utf-8") as out:
        for line in codecs.open(os.path.join(processed_data_dir, "targetIDs"), "r", "utf-8"):
            tgtSeq, tid = line.strip().split("\t")
            ids = _pad(encoder.encode(tgtSeq.lower()), max_seq_length, "Target", tgtSeq)
            encodedFullTargetSpace[tid] = ids
            tgtIdNameMap[tid] = tgtSeq
            out.write(tid + "\t" + tgtSeq.strip() + "\t" + ",".join([str(i) for i in ids]) + "\n")
    evalCorpus = gen_postive_corpus(os.path.join(processed_data_dir, "EvalPairs"), encodedFullTargetSpace, encoder, max_seq_length)
    trainCorpus = gen_postive_corpus(os.path.join(processed_data_dir, "TrainPairs"), encodedFullTargetSpace, encoder, max_seq_length)
    return encoder, trainCorpus, evalCorpus, encodedFullTargetSpace, tgtIdNameMap


def load_encodedTargetSpace(processed_data_dir):
    """Reference data_utils.py:217-239."""
    encoder = load_vocabulary(processed_data_dir)
    fn = os.path.join(processed_data_dir, "encoded.FullTargetSpace")
    if not os.path.exists(fn):
        raise ValueError("Error! could not found encoded.FullTargetSpace in model folder.")
    encodedTgtSpace, names = {}, {}
    for line in codecs.open(fn, "r", "utf-8"):
        tgtId, tgtName, tgtEncoding = line.strip().split("\t")
        names[tgtId] = tgtName
        encodedTgtSpace[tgtId] = [int(i) for i in tgtEncoding.split(",")]
    return encoder, encodedTgtSpace, names


def save_model_configs(processed_data_dir, configs):
    """key=value lines (reference data_utils.py:244-249)."""
    with codecs.open(os.path.join(processed_data_dir, "modelConfig.param"), "w", "utf-8") as f:
        for key in configs.keys():
            f.write(str(key) + "=" + str(configs[key]) + "\n")


def load_model_configs(processed_data_dir):
    """All values come back as strings (reference data_utils.py:252-259)."""
    cfg = {}
    for line in codecs.open(os.path.join(processed_data_dir, "modelConfig.param"), "r", "utf-8").readlines():
        if "=" not in line.strip():
            continue
        key, value = line.strip().split("=")
        cfg[key] = value
    return cfg


def getSortedResults(scores):
    """Reference data_utils.py:263-267 (kept for callers that still hold a dense score matrix)."""
    rankedIdx = np.argsort(-scores)
    sortedScore = -np.sort(-scores, axis=1)
    return sortedScore, rankedIdx


def computeTopK_TightVersion_accuracy(topk, labels, results):
    """Reference data_utils.py:270-286.  `results` = ranked ids per row; only the first k columns matter,
    so a [Q,k'] (k' >= topk) matrix from the fused GPU top-k gives the same value as the full argsort."""
    assert len(labels) == len(results)
    k = min(topk, results.shape[1])
    total = 0.0
    for i in range(results.shape[0]):
        top = set(int(v) for v in results[i][:k])
        cur = sum(1.0 for lab in labels[i] if lab in top)
        total += cur / len(labels[i])
    return total / float(results.shape[0])


def computeTopK_accuracy(topk, labels, results):
    """Reference data_utils.py:289-304."""
    assert len(labels) == len(results)
    k = min(topk, results.shape[1])
    total = 0.0
    for i in range(results.shape[0]):
        top = set(int(v) for v in results[i][:k])
        if any(lab in top for lab in labels[i]):
            total += 1.0
    return total / float(results.shape[0])
