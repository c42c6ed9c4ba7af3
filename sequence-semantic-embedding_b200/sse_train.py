# coding=utf-8
"""Training entry point (re-host of reference sse_train.py: same flags, defaults, log lines,
checkpoint names and per-epoch index + evaluation; the compute is the CUDA train step).

    python sse_train.py --task_type=classification --data_dir=rawdata-classification \
        --model_dir=models-classification --learning_rate=0.9 --max_epoc=50 --steps_per_checkpoint=200
"""
from __future__ import print_function

import argparse
import logging
import os
import sys
import time
from logging import handlers

import data_utils
import sse_evaluator
import sse_index
import sse_model
from data import Data

FLAGS = None


def parse_flags(argv=None):
    """Flag names and defaults of reference sse_train.py:60-88."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--learning_rate", type=float, default=0.9, help="Learning rate.")
    ap.add_argument("--learning_rate_decay_factor", type=float, default=0.99, help="Learning rate decays by this much.")
    ap.add_argument("--batch_size", type=int, default=64, help="Batch size to use during training(positive pair count based).")
    ap.add_argument("--embedding_size", type=int, default=50, help="Size of word embedding vector.")
    ap.add_argument("--encoding_size", type=int, default=64, help="Size of sequence encoding vector.")
    ap.add_argument("--src_cell_size", type=int, default=96, help="LSTM cell size in source RNN model.")
    ap.add_argument("--tgt_cell_size", type=int, default=96, help="LSTM cell size in target RNN model.")
    ap.add_argument("--num_layers", type=int, default=1, help="Number of layers in the model (unused, as in the reference).")
    ap.add_argument("--vocab_size", type=int, default=32000)
    ap.add_argument("--max_seq_length", type=int, default=80)
    ap.add_argument("--max_epoc", type=int, default=30)
    ap.add_argument("--predict_nbest", type=int, default=10)
    ap.add_argument("--task_type", default="classification")
    ap.add_argument("--data_dir", default="rawdata-classification")
    ap.add_argument("--model_dir", default="models-classification")
    ap.add_argument("--rawfilename", default="targetIDs")
    ap.add_argument("--encodedIndexFile", default="targetEncodingIndex.tsv")
    ap.add_argument("--device", default="0")
    ap.add_argument("--network_mode", default="dual-encoder")
    ap.add_argument("--steps_per_checkpoint", type=int, default=200)
    ap.add_argument("--seed", type=int, default=None, help="(new) seed for the initialisers and the batch sampler")
    ap.add_argument("--max_steps", type=int, default=0, help="(new) stop after this many steps (0 = run max_epoc epochs)")
    return ap.parse_args(argv)


def create_model(session, targetSpaceSize, vocabsize, forward_only):
    """Create SSE model and initialize or load parameters (reference sse_train.py:96-121)."""
    modelParams = {"max_seq_length": FLAGS.max_seq_length, "vocab_size": vocabsize,
                   "embedding_size": FLAGS.embedding_size, "encoding_size": FLAGS.encoding_size,
                   "learning_rate": FLAGS.learning_rate, "learning_rate_decay_factor": FLAGS.learning_rate_decay_factor,
                   "src_cell_size": FLAGS.src_cell_size, "tgt_cell_size": FLAGS.tgt_cell_size,
                   "network_mode": FLAGS.network_mode, "predict_nbest": FLAGS.predict_nbest,
                   "targetSpaceSize": targetSpaceSize, "forward_only": forward_only}
    data_utils.save_model_configs(FLAGS.model_dir, modelParams)
    model = sse_model.SSEModel(modelParams, device=int(str(FLAGS.device).split(",")[0]))
    ckpt = sse_model.get_checkpoint_state(FLAGS.model_dir)
    if ckpt:
        logging.info("Reading model parameters from %s" % ckpt.model_checkpoint_path)
        model.saver.restore(session, ckpt.model_checkpoint_path)
    else:
        if forward_only:
            logging.error("Error!!!Could not load any model from specified folder: %s" % FLAGS.model_dir)
            sys.exit(-1)
        logging.info("Created model with fresh parameters.")
        session.run(sse_model.global_variables_initializer())
    return model


def set_up_logging():
    os.makedirs(FLAGS.model_dir, exist_ok=True)
    log = logging.getLogger("")
    log.setLevel(logging.DEBUG)
    fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s", datefmt="%m/%d/%Y %I:%M:%S %p")
    ch = logging.StreamHandler(sys.stdout)
    ch.setFormatter(fmt)
    log.addHandler(ch)
    fh = handlers.RotatingFileHandler(FLAGS.model_dir + "/TrainingLog.txt", maxBytes=(1048576 * 20), backupCount=7)
    fh.setFormatter(fmt)
    log.addHandler(fh)


class WindowStats(object):
    """Means over the current reporting window of `steps_per_checkpoint` steps (reference sse_train.py:170-174, 215:
    the reference accumulates x / steps_per_checkpoint per step and zeroes the sums after each report)."""

    def __init__(self, window):
        self.window = float(window)
        self.reset()

    def reset(self):
        self.step_time = self.loss = self.train_acc = 0.0

    def add(self, seconds, loss, acc):
        self.step_time += seconds / self.window
        self.loss += loss / self.window
        self.train_acc += acc / self.window


class CheckpointPolicy(object):
    """What to do at a reporting point, given the window's training accuracy (reference sse_train.py:196-212):
      * decay the learning rate when more than six reports exist and this one is below the minimum of the last five;
      * the accuracy joins the history; equal to the best so far -> save `-BestEver`, else report and skip;
      * `finished` (save `-final`, leave the batch loop) when epoch > 10 and the accuracy is below the minimum of the
        last five reports -- evaluated AFTER the append, as the reference does, so the window's own value is part of
        that minimum and the branch cannot fire; kept for parity of behaviour."""

    def __init__(self):
        self.history = []

    def report(self, acc, epoch):
        decay = len(self.history) > 6 and acc < min(self.history[-5:])
        self.history.append(acc)
        return {"decay_lr": decay, "save_best": acc == max(self.history), "best": max(self.history),
                "finished": epoch > 10 and acc < min(self.history[-5:])}


def _evaluate_epoch(model, data, sess, epoch):
    """Index the whole target space with the current weights and report top-1/3/10 (reference sse_train.py:221-227)."""
    model.set_forward_only(True)
    idx_file = os.path.join(FLAGS.model_dir, FLAGS.encodedIndexFile)
    sse_index.createIndexFile(model, data.encoder, os.path.join(FLAGS.model_dir, FLAGS.rawfilename), FLAGS.max_seq_length, idx_file,
                              sess, batchsize=1000)
    acc1, acc3, acc10 = sse_evaluator.Evaluator(model, data.rawEvalCorpus, idx_file, sess).eval()
    logging.info("epoc#%d, task specific evaluation: top 1/3/10 accuracies: %f / %f / %f \n\n\n" % (epoch, acc1, acc3, acc10))


def train():
    logging.info("Preparing Train & Eval data in %s" % FLAGS.data_dir)
    for d in (FLAGS.data_dir, FLAGS.model_dir):
        os.makedirs(d, exist_ok=True)
    data = Data(FLAGS.model_dir, FLAGS.data_dir, FLAGS.vocab_size, FLAGS.max_seq_length, seed=FLAGS.seed)
    epoc_steps = len(data.rawTrainPosCorpus) / FLAGS.batch_size
    logging.info("Training Data: %d total positive samples, each epoch need %d steps" % (len(data.rawTrainPosCorpus), epoc_steps))
    ckpt_prefix = os.path.join(FLAGS.model_dir, "SSE-LSTM.ckpt")
    with sse_model.Session(seed=FLAGS.seed) as sess:
        model = create_model(sess, data.rawnegSetLen, data.vocab_size, False)
        summary_op = model.add_summaries()
        fetches = [model.train, summary_op, model.loss, model.train_acc]
        stats, policy = WindowStats(FLAGS.steps_per_checkpoint), CheckpointPolicy()
        # same sampling rule and row layout as data.get_train_batch, drawn with array operations (8 ms -> 1.3 ms per 512-pair batch)
        next_batch = getattr(data, "get_train_batch_arrays", data.get_train_batch)
        steps_done, out_of_budget = 0, False
        for epoch in range(FLAGS.max_epoc):
            t_epoch = time.time()
            for _ in range(int(epoc_steps)):
                t0 = time.time()
                src, tgt, labels = next_batch(FLAGS.batch_size)
                model.set_forward_only(False)
                _, _summary, step_loss, step_acc = sess.run(fetches, feed_dict=model.get_train_feed_dict(src, tgt, labels))
                stats.add(time.time() - t0, step_loss, step_acc)
                steps_done += 1
                out_of_budget = bool(FLAGS.max_steps) and steps_done >= FLAGS.max_steps
                if steps_done % FLAGS.steps_per_checkpoint == 0:
                    gs = model.global_step.eval()
                    logging.info("global epoc: %.3f, global step %d, learning rate %.4f step-time:%.2f loss:%.4f train_binary_acc:%.4f " %
                                 (float(gs) / float(epoc_steps), gs, model.learning_rate.eval(), stats.step_time, step_loss, stats.train_acc))
                    verdict = policy.report(stats.train_acc, epoch)
                    if verdict["decay_lr"]:
                        sess.run(model.learning_rate_decay_op)
                    if verdict["save_best"]:
                        logging.info("Better Accuracy %.4f found. Saving current best model ..." % stats.train_acc)
                        model.save(sess, ckpt_prefix + "-BestEver")
                    else:
                        logging.info("Best Accuracy is: %.4f, while current round is: %.4f" % (verdict["best"], stats.train_acc))
                        logging.info("skip saving model ...")
                    if verdict["finished"]:
                        where = model.save(sess, ckpt_prefix + "-final")
                        logging.info("After around %d Epocs no further improvement, Training finished, wrote checkpoint to %s." % (epoch, where))
                        break
                    stats.reset()
                if out_of_budget:
                    break
            logging.info("\n\n\nepoch# %d  took %f hours" % (epoch, (time.time() - t_epoch) / 3600.0))
            if FLAGS.task_type not in ("ranking", "crosslingual") or (epoch + 1) % 20 == 0 or out_of_budget:
                _evaluate_epoch(model, data, sess, epoch)
            model.save(sess, ckpt_prefix + "-epoch-%d" % epoch)
            if policy.history:
                logging.info("So far best ever model training binary accuracy is: %.4f " % max(policy.history))
            if out_of_budget:
                break


def main(argv=None):
    global FLAGS
    FLAGS = parse_flags(argv)
    set_up_logging()
    if not FLAGS.data_dir or not FLAGS.model_dir:
        logging.error("--data_dir and --model_dir must be specified.")
        sys.exit(1)
    train()


if __name__ == "__main__":
    main()
