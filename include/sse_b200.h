/*
 * sse_b200.h -- C ABI of libsse_b200.so, the B200-native replacement for the
 * hot path of eBay/Sequence-Semantic-Embedding (reference = TensorFlow-1 graph
 * in sse_model.py + numpy ranking in sse_evaluator.py / data_utils.py).
 *
 * The reference has no FFI of its own: its de-facto boundary is
 * `tf.Session.run(model.<tensor>, feed_dict=model.get_*_feed_dict(...))`.  Each
 * entry point below names the reference tensor / call site it replaces
 * (file:line in the reference repo).  The Python facade
 * (sequence-semantic-embedding_b200/sse_model.py) binds these with ctypes and
 * re-exposes the reference's SSEModel attribute names.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SSE_E* code on failure;
 *     nothing throws or exits across the ABI; sse_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - plain pointers and sizes only.  "dev" pointers are CUDA device pointers
 *     on the handle's device (e.g. torch.Tensor.data_ptr()); "host" pointers are
 *     ordinary (ideally pinned) host memory.  The caller owns every buffer it
 *     passes; the library owns weights, optimizer slots, workspaces, the fp16
 *     copy of the index.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls
 *     taking a stream are asynchronous on it; *_host calls synchronise before
 *     returning (they include H2D and D2H copies).
 *   - a handle is single-stream / single-threaded; concurrent callers need one
 *     handle each or an external lock (the facade holds a mutex, mirroring the
 *     thread-safety of tf.Session.run used by webserver.py:107-121).
 */
#ifndef SSE_B200_H_
#define SSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSE_OK          0
#define SSE_EINVAL     -1   /* bad argument / unknown name / shape mismatch   */
#define SSE_ECUDA      -2   /* CUDA runtime or driver error                   */
#define SSE_ENOMEM     -3
#define SSE_ESTATE     -4   /* call not valid in this mode / missing index    */
#define SSE_ENODEVICE  -5   /* no CUDA device: the product path has NO CPU fallback */

/* network_mode, sse_model.py:166-177 (+ "dual-cnn", the BASELINE.json config-3 extension) */
#define SSE_MODE_DUAL_ENCODER        0
#define SSE_MODE_SHARED_ENCODER      1
#define SSE_MODE_SOURCE_ENCODER_ONLY 2
#define SSE_MODE_SOURCE_ONLY_CNN     3
#define SSE_MODE_DUAL_CNN            4

/* arithmetic of the tensor-core paths (the SIMT fp32 path is always available) */
#define SSE_PRECISION_FP32  0   /* every kernel in fp32 SIMT (exact mode)              */
#define SSE_PRECISION_TC    1   /* fp16-operand tcgen05 scan + exact fp32 re-rank; fp16-operand tcgen05 encoders */

#define SSE_SIDE_SRC 0
#define SSE_SIDE_TGT 1

#define SSE_MAX_CNN_FILTERS 8
#define SSE_MAX_TOPK 128

typedef struct sse_handle sse_handle;

/* The 12 modelParams keys of SSEModel.__init__ (sse_model.py:113-126) plus
 * placement / precision.  Zero-initialise, set struct_size = sizeof(sse_config). */
typedef struct sse_config {
  int32_t struct_size;
  int32_t network_mode;             /* SSE_MODE_*                               */
  int32_t vocab_size;               /* modelParams['vocab_size']                */
  int32_t embedding_size;           /* word_embed_size  (We)                    */
  int32_t encoding_size;            /* seq_embed_size   (E)                     */
  int32_t src_cell_size;            /* Hs                                       */
  int32_t tgt_cell_size;            /* Ht                                       */
  int32_t max_seq_length;           /* T                                        */
  int32_t predict_nbest;            /* TOP_N                                    */
  int32_t forward_only;
  int64_t target_space_size;        /* targetSpaceSize (source-only modes)      */
  float   learning_rate;
  float   learning_rate_decay_factor;
  int32_t device;                   /* CUDA ordinal                             */
  int32_t precision;                /* SSE_PRECISION_*                          */
  int32_t n_cnn_filters;            /* CNN modes: number of n-gram widths (0 -> reference 2,3,4,5 x 256,128,128,64) */
  int32_t cnn_filter_sizes[SSE_MAX_CNN_FILTERS];
  int32_t cnn_num_filters[SSE_MAX_CNN_FILTERS];
  int32_t reserved[8];
} sse_config;

/* ---- lifetime ----------------------------------------------------------- */
/* replaces SSEModel(modelParams) graph construction, sse_model.py:94-138.
 * Allocates every variable of the mode (uninitialised weights, Adagrad slots
 * at 0.1, global_step 0).  Fails with SSE_ENODEVICE when no GPU is present. */
int sse_create(const sse_config* cfg, sse_handle** out);
int sse_destroy(sse_handle* h);
const char* sse_last_error(void);
/* library / build identification: "sse_b200 <version> sm_100a" */
const char* sse_version(void);

/* ---- variables (tf.train.Saver / tf.global_variables, sse_model.py:138) --- */
/* name = TF variable name, e.g. "word_embedding",
 * "source_encoder/rnn/basic_lstm_cell/kernel", "target_encoder/tgt_M",
 * "<var>/Adagrad".  Data is fp32 row-major in the TF layout; ptr may be a host
 * or a device pointer.  shape/ndim must match the variable. */
int sse_set_param(sse_handle* h, const char* name, const void* ptr, const int64_t* shape, int ndim);
int sse_get_param(sse_handle* h, const char* name, void* host_dst, int64_t nbytes);
/* enumerate variables: returns count; fills name (<=127 chars) and shape of variable i */
int sse_param_count(sse_handle* h);
int sse_param_info(sse_handle* h, int i, char* name_out, int name_cap, int64_t* shape_out /*[4]*/, int* ndim_out);

/* ---- encoders ------------------------------------------------------------ */
/* replaces sess.run(model.{src,tgt}_seq_embedding | model.norm_{src,tgt}_seq_embedding)
 * (sse_index.py:90-91, sse_evaluator.py:107-108, sse_demo.py:121-123,
 * webserver.py:144-146,183-184): embedding gather -> LSTM over all T positions
 * (or n-gram CNN + max-pool) -> x M -> optional l2-normalise.
 * tokens: int32 [B, T] row-major, left-padded with PAD_ID=0, ending in EOS_ID=1.
 * out: fp32 [B, E]. */
int sse_encode(sse_handle* h, int side, const int32_t* tokens_dev, int B, float* out_dev,
               int normalize, void* stream);
/* same, host buffers (H2D of tokens and D2H of encodings inside the call). */
int sse_encode_host(sse_handle* h, int side, const int32_t* tokens_host, int B, float* out_host,
                    int normalize);

/* ---- index + retrieval --------------------------------------------------- */
/* replaces Evaluator.__init__'s targetEncodings matrix (sse_evaluator.py:80-92):
 * registers this rank's shard of the target index.  tgt: fp32 [N_local, E],
 * host or device; the library keeps its own fp32 copy (and an fp16 copy for the
 * tensor-core scan).  global_offset = global id of local row 0. */
int sse_index_set(sse_handle* h, const float* tgt, int64_t n_local, int64_t global_offset);
/* encode this rank's target rows straight into the resident index
 * (createIndexFile without the TSV, sse_index.py:76-92).  tokens: int32 [N_local,T] (host or device). */
int sse_index_build(sse_handle* h, const int32_t* tgt_tokens, int64_t n_local, int64_t global_offset,
                    int batch);
/* copy the resident fp32 index rows [row0,row0+n) to host (for the TSV writer) */
int sse_index_get(sse_handle* h, int64_t row0, int64_t n, float* host_dst);

/* replaces np.dot(sourceEncodings, targetEncodings.T) + data_utils.getSortedResults
 * (sse_evaluator.py:110-111, data_utils.py:263-267, sse_demo.py:126-127) and
 * tf.nn.top_k(similarity, TOP_N) (sse_model.py:348): scores of q [Q,E] against
 * the local shard and the k best per row, descending, ties -> lower index.
 * scores: fp32 [Q,k]; idx: int32 [Q,k] GLOBAL ids.  Rows of shards with fewer
 * than k targets are padded with (-inf, -1). */
int sse_search(sse_handle* h, const float* q_dev, int Q, int k, float* scores_dev, int32_t* idx_dev,
               void* stream);
/* Multi-GPU form of sse_search: the local shard's top-k of every row as ONE packed block
 * packed [Q, 2k] fp32 -- columns [0,k) scores, columns [k,2k) the int32 GLOBAL ids bit-cast -- i.e. exactly the
 * message a rank contributes to the NCCL exchange (all-gather, or all-to-all of the row blocks each peer owns). */
int sse_search_packed(sse_handle* h, const float* q_dev, int Q, int k, float* packed_dev, void* stream);
/* ... and the step after the exchange: gathered [G, Q, 2k] = the packed blocks of G shards for the same Q rows
 * (as NCCL lays them out, rank-major) -> the k best of the G*k candidates per row, (score desc, id asc). */
int sse_merge_packed(sse_handle* h, const float* gathered_dev, int G, int Q, int k, float* scores_dev,
                     int32_t* idx_dev, void* stream);
/* merge C candidates per row (e.g. the all-gathered per-shard top-k) into the
 * k best: the cross-shard step after the NCCL all-gather. */
int sse_merge_topk(sse_handle* h, const float* cand_scores_dev, const int32_t* cand_idx_dev, int Q, int C,
                   int k, float* scores_dev, int32_t* idx_dev, void* stream);
/* end-to-end query: host tokens [Q,T] -> H2D -> source encoder -> search of the
 * local shard -> D2H of [Q,k] scores and ids.  normalize=1 for the evaluator /
 * classify route, 0 for demo/search/qna/crosslingual (sse_demo.py:123). */
int sse_query_host(sse_handle* h, const int32_t* tokens_host, int Q, int k, int normalize,
                   float* scores_host, int32_t* idx_host);
/* replaces model.predicted_tgts_score / model.predicted_labels (sse_model.py:286,344-350):
 * tf.nn.top_k(similarity, TOP_N, sorted=True) of q [Q,E] against the BATCH's own target encodings tgt [n_tgt,E]
 * (device pointers; nothing is registered, the resident index is untouched), ids = row numbers of tgt, ties -> lower
 * index; normalize_scores = 1 applies l2_normalize(scores, 1) as :350 does.  n_tgt < k is TF's "input must have at
 * least k columns" error. */
int sse_topk_batch(sse_handle* h, const float* q_dev, int Q, const float* tgt_dev, int64_t n_tgt, int k,
                   int normalize_scores, float* scores_dev, int32_t* idx_dev, void* stream);
/* x * rsqrt(max(sum x^2, 1e-12)) over each row, in place (tf.nn.l2_normalize,
 * sse_model.py:282-283,350). */
int sse_l2_normalize_rows(sse_handle* h, float* x_dev, int rows, int cols, void* stream);

/* ---- training ------------------------------------------------------------ */
/* replaces model.binarylogit (sse_model.py:290): per-pair cosine [B]. */
int sse_pair_score(sse_handle* h, const int32_t* src_dev, const int32_t* tgt_dev, int B, float* cos_dev,
                   void* stream);
/* replaces sess.run([model.train, model.loss, model.train_acc], get_train_feed_dict(...))
 * (sse_train.py:170-172; sse_model.py:279-302,355-364): forward both towers,
 * pair loss on 64*cos, BPTT, clip_by_global_norm(5.0), Adagrad, global_step++.
 * src,tgt: int32 [B,T]; labels: fp32 [B]; host or device pointers.
 * loss_host / acc_host / gnorm_host (optional, may be NULL) receive the scalars
 * (this synchronises the stream). */
int sse_train_step(sse_handle* h, const int32_t* src, const int32_t* tgt, const float* labels, int B,
                   float* loss_host, float* acc_host, float* gnorm_host, void* stream);
/* data-parallel variant, phase 1: forward + backward into the gradient arena
 * without applying; the host all-reduces the arena over NCCL, then phase 2. */
int sse_train_grads(sse_handle* h, const int32_t* src, const int32_t* tgt, const float* labels, int B,
                    int B_global, float* loss_host, float* acc_host, void* stream);
int sse_grad_arena(sse_handle* h, float** dev_ptr_out, int64_t* n_floats_out);
int sse_train_apply(sse_handle* h, float* loss_host, float* acc_host, float* gnorm_host, void* stream);
/* Device-side train-batch sampler (replaces the python loop of data.py:95-115 -> Data.get_train_batch).
 * sse_sampler_set copies the training corpus to the device once: src_rows int32 [n_pos, T] (padded source sequences),
 * the verified targets of positive i as ver_rows[ver_off[i] .. ver_off[i+1]) (ROW numbers into tgt_rows, CSR; every
 * positive needs at least one), tgt_rows int32 [n_tgt, T] (the encoded full target space).  All host pointers.
 * sse_sampler_batch writes the pair rows of the window of positives [start, start + batch_size) -- rows alternate
 * (source, one verified target, 1.0), (same source, one uniformly drawn NON-verified target, 0.0), exactly the layout
 * sse_train_step consumes -- into device buffers sized [2*batch_size, T] / [2*batch_size]; *rows_out = rows written
 * (shorter near the end of the corpus, as data.py:97-98).  Draws are a hash of (seed, step, positive): reproducible. */
int sse_sampler_set(sse_handle* h, const int32_t* src_rows, int64_t n_pos, const int64_t* ver_off, const int32_t* ver_rows,
                    const int32_t* tgt_rows, int64_t n_tgt);
int sse_sampler_batch(sse_handle* h, int64_t start, int batch_size, uint64_t seed, uint64_t step, int32_t* src_dev,
                      int32_t* tgt_dev, float* labels_dev, int* rows_out, void* stream);
/* learning_rate_decay_op, sse_model.py:123-124 */
int sse_lr_decay(sse_handle* h);
int sse_get_scalars(sse_handle* h, float* learning_rate, int64_t* global_step);
int sse_set_scalars(sse_handle* h, float learning_rate, int64_t global_step);

/* Token ids outside [0, vocab_size) (TensorFlow's embedding_lookup raises InvalidArgument, sse_model.py:163-164):
 * every encoder / train entry point checks the ids on the device; offenders are read as PAD_ID and counted.  Entry
 * points that synchronise anyway (*_host, sse_index_build, train calls that return scalars) fail with SSE_EINVAL;
 * after asynchronous calls this returns (and resets) the count seen so far -- it synchronises `stream`. */
int sse_token_errors(sse_handle* h, int64_t* count_out, void* stream);

/* ---- introspection for benchmarks / tests -------------------------------- */
/* bookkeeping of the LAST tcgen05 search on this handle (synchronises the device): candidates that passed the sampled
 * threshold summed over all rows, rows searched, rows that overflowed into the brute-force fallback, scan work items */
int sse_search_stats(sse_handle* h, int64_t* candidates, int* rows, int* fallback_rows, int* items);
/* test hook of the internal tcgen05 GEMM (csrc/gemm_tc.cu) behind the tensor-core train step and the CNN tower:
 * D[M,N] = alpha * A[M,K] B[N,K]^T (+ beta D) with the fp32 inputs rounded to fp16 (fmt 0) or bf16 (fmt 1), fp32
 * accumulation; split_k > 1 adds the partial sums atomically into D (beta must be 1 or D pre-zeroed). */
int sse_debug_gemm_tc(sse_handle* h, const float* a_dev, const float* b_dev, int M, int N, int K, int fmt, int split_k,
                      float alpha, float beta, float* d_dev, void* stream);
/* number of kernels this library launched on the handle since creation */
int64_t sse_launch_count(sse_handle* h);
/* select kernel variants at run time: key in {"search", "encoder", "lstm_kernel", "cluster_rows", "pad_skip", "search_ctas", "search_late_ctas", "search_late_share", "train"};
 * search: 0 auto, 1 simt-fp32, 2 tcgen05-fp16;  encoder: 0 auto, 1 simt-fp32, 2 tcgen05;
 * lstm_kernel (tcgen05 encoder only): 0 auto, 1 weight-streaming kernel, 2 cluster kernel (weights resident in
 * the shared memory of a thread-block cluster), 3 cluster kernel with the input projection tabulated per
 * vocabulary entry (V x 4H fp32 table, rebuilt when parameters change; the default when it fits in 2 GiB);
 * cluster_rows: batch rows per thread-block cluster of the table LSTM kernel: 0 (default) = 64 when every cluster still gets its own
 * SMs (query batches: half the per-step work per SM, lowest latency), else 128; 64 / 128 force (128 keeps a 600-row batch on 40 SMs
 * for a step that overlaps it with a capped scan);
 * pad_skip: 1 (default) rows are bucketed by their number of leading PADs on the device and every kernel tile starts
 * from the tabulated pad-prefix state instead of running the PAD steps (all entry points), 0 off;  search_ctas: cap on the scan grid (0 = all SMs),
 * so that an encoder launched on another stream can run concurrently on the remaining SMs;  search_late_ctas /
 * search_late_share (with a cap): that many EXTRA scan CTAs, numbered last so that they start when the concurrent
 * kernel frees its SMs, each taking `share` percent of a regular CTA's tile range (default 40);
 * train: 0 auto (tensor cores when the handle was created with SSE_PRECISION_TC and the shapes allow), 1 fp32 SIMT
 * (parity mode), 2 tensor cores (bf16 operands, fp32 accumulation / state / optimizer; error if unsupported). */
int sse_set_option(sse_handle* h, const char* key, int value);
/* names + durations of the last timed kernels are not kept here: time with CUDA events on `stream`. */

/* ---- targetEncodingIndex.tsv fast writer / reader (host only) -------------------------------
 * Replaces the python loops of reference sse_index.py:93-97 (writer: id \t text \t ','.join(str(np.float32)))
 * and sse_evaluator.py:79-88 (reader: line.strip().split('\t'), rows without three fields are skipped).
 * Formatting is numpy's str(np.float32) byte for byte; parsing is correctly rounded, so a written index reads
 * back bit-exactly.  Rows are split over `threads` host threads (<= 0: all hardware threads). */
const char* sse_tsv_last_error(void);
/* CRC-32C (Castagnoli) of n bytes, continuing from `seed` (0 to start): checksum of TF table blocks / bundle entries. */
uint32_t sse_crc32c(const void* data, size_t n, uint32_t seed);
/* n floats -> their decimal strings packed back to back (no separators); ends[i] = end offset of value i. */
int sse_tsv_format_f32(const float* values, int64_t n, char* out, size_t cap, int64_t* ends);
int sse_tsv_write_index(const char* path, const char* const* ids, const char* const* texts, const float* rows,
                        int64_t n_rows, int E, int append, int threads);
/* buf = whole file contents.  out: [max_rows, E]; spans: [max_rows, 4] byte offsets (id_begin, id_end, text_begin,
 * text_end) of every accepted row, in file order; n_skipped counts rows rejected by the three-field rule.  A row
 * with three fields whose vector does not hold exactly E floats is an error (the reference raises there). */
int sse_tsv_parse_index(const char* buf, size_t len, int E, int64_t max_rows, float* out, int64_t* spans,
                        int64_t* n_rows, int64_t* n_skipped, int threads);

/* ---- batched subword tokenizer + padder (host only) ---------------------------------------------
 * The ENCODE half of the reference's SubwordTextEncoder for a vocabulary.txt it wrote (text_encoder.py:334-356,
 * 427-436, 491-532; tokenizer.py:68-90) plus the row rule of data_utils.py:149-155: n lower-cased utf-8 sentences
 * -> int32 [n,T] rows ([PAD]*(T-len-1) + ids + [EOS], or [PAD] + ids[:T-2] + [EOS]); lengths[i] = subtokens before
 * padding / truncation.  subtokens_utf8 = the vocabulary lines with their quotes removed, in file order. */
typedef struct sse_tokenizer sse_tokenizer;
const char* sse_tok_last_error(void);
int sse_tok_create(const char* const* subtokens_utf8, int n_subtokens, sse_tokenizer** out);
int sse_tok_destroy(sse_tokenizer* t);
int sse_tok_vocab_size(const sse_tokenizer* t);
int sse_tok_encode_batch(const sse_tokenizer* t, const char* const* texts_utf8, int64_t n, int T, int32_t* rows,
                         int32_t* lengths, int threads);

#ifdef __cplusplus
}
#endif
#endif /* SSE_B200_H_ */
