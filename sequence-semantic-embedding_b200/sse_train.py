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


def train():
    logging.info("Preparing Train & Eval data in %s" % FLAGS.data_dir)
    for d in (FLAGS.data_dir, FLAGS.model_dir):
        os.makedirs(d, exist_ok=True)
    data = Data(FLAGS.model_dir, FLAGS.data_dir, FLAGS.vocab_size, FLAGS.max_seq_length, seed=FLAGS.seed)
    epoc_steps = len(data.rawTrainPosCorpus) / FLAGS.batch_size
    logging.info("Training Data: %d total positive samples, each epoch need %d steps" % (len(data.rawTrainPosCorpus), epoc_steps))
    with sse_model.Session(seed=FLAGS.seed) as sess:
        model = create_model(sess, data.rawnegSetLen, data.vocab_size, False)
        summary_op = model.add_summaries()
        step_time, loss, train_acc = 0.0, 0.0, 0.0
        current_step = 0
        previous_accuracies = []
        stop = False
        for epoch in range(FLAGS.max_epoc):
            epoc_start_Time = time.time()
            for _batchId in range(int(epoc_steps)):
                start_time = time.time()
                source_inputs, tgt_inputs, labels = data.get_train_batch(FLAGS.batch_size)
                model.set_forward_only(False)
                d = model.get_train_feed_dict(source_inputs, tgt_inputs, labels)
                _, _summary, step_loss, step_train_acc = sess.run([model.train, summary_op, model.loss, model.train_acc], feed_dict=d)
                step_time += (time.time() - start_time) / FLAGS.steps_per_checkpoint
                loss += step_loss / FLAGS.steps_per_checkpoint
                train_acc += step_train_acc / FLAGS.steps_per_checkpoint
                current_step += 1
                if FLAGS.max_steps and current_step >= FLAGS.max_steps:
                    stop = True
                if current_step % FLAGS.steps_per_checkpoint == 0:
                    logging.info("global epoc: %.3f, global step %d, learning rate %.4f step-time:%.2f loss:%.4f train_binary_acc:%.4f " %
                                 (float(model.global_step.eval()) / float(epoc_steps), model.global_step.eval(),
                                  model.learning_rate.eval(), step_time, step_loss, train_acc))
                    checkpoint_path = os.path.join(FLAGS.model_dir, "SSE-LSTM.ckpt")
                    # Decrease learning rate if no improvement was seen over last 5 times.
                    if len(previous_accuracies) > 6 and train_acc < min(previous_accuracies[-5:]):
                        sess.run(model.learning_rate_decay_op)
                    previous_accuracies.append(train_acc)
                    if train_acc == max(previous_accuracies):
                        logging.info("Better Accuracy %.4f found. Saving current best model ..." % train_acc)
                        model.save(sess, checkpoint_path + "-BestEver")
                    else:
                        logging.info("Best Accuracy is: %.4f, while current round is: %.4f" % (max(previous_accuracies), train_acc))
                        logging.info("skip saving model ...")
                    if epoch > 10 and train_acc < min(previous_accuracies[-5:]):
                        p = model.save(sess, checkpoint_path + "-final")
                        logging.info("After around %d Epocs no further improvement, Training finished, wrote checkpoint to %s." % (epoch, p))
                        stop = True
                    step_time, loss, train_acc = 0.0, 0.0, 0.0
                if stop:
                    break
            logging.info("\n\n\nepoch# %d  took %f hours" % (epoch, (time.time() - epoc_start_Time) / 3600.0))
            if (FLAGS.task_type not in ["ranking", "crosslingual"]) or ((epoch + 1) % 20 == 0) or stop:
                model.set_forward_only(True)
                idx_file = os.path.join(FLAGS.model_dir, FLAGS.encodedIndexFile)
                sse_index.createIndexFile(model, data.encoder, os.path.join(FLAGS.model_dir, FLAGS.rawfilename),
                                          FLAGS.max_seq_length, idx_file, sess, batchsize=1000)
                evaluator = sse_evaluator.Evaluator(model, data.rawEvalCorpus, idx_file, sess)
                acc1, acc3, acc10 = evaluator.eval()
                logging.info("epoc#%d, task specific evaluation: top 1/3/10 accuracies: %f / %f / %f \n\n\n" % (epoch, acc1, acc3, acc10))
            model.save(sess, os.path.join(FLAGS.model_dir, "SSE-LSTM.ckpt") + "-epoch-%d" % epoch)
            if previous_accuracies:
                logging.info("So far best ever model training binary accuracy is: %.4f " % max(previous_accuracies))
            if stop:
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
