# coding=utf-8
"""Build targetEncodingIndex.tsv with the B200 target encoder (re-host of reference sse_index.py).

File format is byte-compatible with the reference writer (sse_index.py:93-95):
``targetId \\t raw target text \\t E comma-separated str(np.float32) values``, one row per
well-formed ``targetIDs`` line, in file order (malformed lines are counted and skipped,
sse_index.py:69-74).  The encodings come from sess.run([model.norm_tgt_seq_embedding])."""
from __future__ import print_function

import argparse
import codecs
import math
import os
import sys

import numpy as np

import data_utils
import sse_model
import sse_ffi
import text_encoder


def createIndexFile(model, encoder, rawfile, max_seq_len, encodeIndexFile, session, batchsize=10000):
    if not os.path.exists(rawfile):
        raise IOError("Error!! Could not find raw target file to be indexed!! :%s" % rawfile)
    codecs.open(encodeIndexFile, "w", "utf-8").close()          # truncate; batches are appended by the native writer
    rawdata = codecs.open(rawfile, "r", "utf-8").readlines()
    cnt = 0
    print("Start indexing whole target space entries with current model ...")
    for batchId in range(math.ceil(len(rawdata) / batchsize)):
        tgtInputs, tgtIds, tgtSentences = [], [], []
        for line in rawdata[batchId * batchsize:(batchId + 1) * batchsize]:
            cnt += 1
            info = line.strip().split("\t")
            if len(info) != 2:
                print("Missing field with error line in raw target file: %s " % line)
                continue
            tgtSentence, tgtId = info[0], info[1]
            tgtIds.append(tgtId)
            tgtSentences.append(tgtSentence)
        if tgtSentences:
            # tokenise + pad the whole batch natively (csrc/subword_tok.cpp): same ids / row rule as the reference loop
            # encoder.encode(tgtSentence.lower()) + [PAD]*(T-len-1) + ids + [EOS] (sse_index.py:77-85)
            tgtInputs, lens = encoder.encode_batch([t.lower() for t in tgtSentences], max_seq_len)
            for tgtSentence, n in zip(tgtSentences, lens.tolist()):
                if n > max_seq_len - 2:
                    print("Error Detected!!! \n Target:\n %s \n Its seq length is:%d,  which is longer than MAX_SEQ_LENTH of %d. "
                          "Try to increase limit!!!!" % (tgtSentence, n, max_seq_len))
        if not tgtSentences:
            continue
        feed = model.get_target_encoding_feed_dict(tgtInputs)
        targetsEncodings = np.vstack(session.run([model.norm_tgt_seq_embedding], feed_dict=feed))
        # same bytes as the reference loop (tgtId \t sentence \t ','.join(str(np.float32))), written by the native
        # multi-threaded formatter (csrc/tsv_io.cpp; python: 5.6 k rows/s at E=256)
        sse_ffi.tsv_write_index(encodeIndexFile, tgtIds, tgtSentences, targetsEncodings, append=True)
    print("Done of all indexing total count:%d" % cnt)


def index(model_dir, rawfile, encodeIndexFile, batchsize=10000):
    if not os.path.exists(model_dir):
        raise IOError("Error! Model folder does not exist!! : %s" % model_dir)
    encoder = data_utils.load_vocabulary(model_dir)
    print("Loaded  vocab size is: %d" % encoder.vocab_size)
    with sse_model.Session() as sess:
        modelConfigs = data_utils.load_model_configs(model_dir)
        model = sse_model.SSEModel(modelConfigs)
        ckpt = sse_model.get_checkpoint_state(model_dir)
        if not ckpt:
            raise IOError("Error!!!Could not load any model from specified folder: %s" % model_dir)
        print("Reading model parameters from %s" % ckpt.model_checkpoint_path)
        model.saver.restore(sess, ckpt.model_checkpoint_path)
        createIndexFile(model, encoder, rawfile, int(modelConfigs["max_seq_length"]), encodeIndexFile, sess, batchsize)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--idx_model_dir", default="models-classification", help="Trained model directory.")
    ap.add_argument("--idx_rawfilename", default="targetIDs", help="raw target sequence file to be indexed")
    ap.add_argument("--idx_encodedIndexFile", default="targetEncodingIndex.tsv", help="target sequece encoding index file.")
    a = ap.parse_args(argv)
    index(a.idx_model_dir, os.path.join(a.idx_model_dir, a.idx_rawfilename), os.path.join(a.idx_model_dir, a.idx_encodedIndexFile))


if __name__ == "__main__":
    main()
