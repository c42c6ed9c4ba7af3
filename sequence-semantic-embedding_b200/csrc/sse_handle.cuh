// The opaque handle behind the C ABI (internal).
#pragma once
#include "sse_common.cuh"

namespace sse {
struct PadTable {
  Scratch buf;
  const float* h = nullptr;   // [T][H]: state after t+1 leading PADs
  const float* c = nullptr;
  bool valid = false;
};
}  // namespace sse

namespace sse {
struct PaddedWeights {            // zero-padded fp32 copies of an LSTM tower's kernel / bias for shapes the tensor-core tiles do not divide
  float* K = nullptr;             // [We' + H', 4H']
  float* b = nullptr;             // [4H']
  bool valid = false;
};
}  // namespace sse

struct sse_handle {
  sse_config cfg;
  int num_sms = 148;
  int cc_major = 0;
  std::vector<sse::Param> params;   // variables first (n_vars), then their "/Adagrad" slots in the same order
  int n_vars = 0;
  int emb_param = -1;
  int tgt_table_param = -1;
  sse::LstmTower lstm[2];           // [SSE_SIDE_SRC], [SSE_SIDE_TGT]
  sse::CnnTower cnn[2];
  sse::CnnTc cnn_tc[2];             // fp16 K-major copies of the CNN filters / projection (lazy, invalidated with the weights)
  sse::PadTable pad[2];             // pad-prefix state tables, fp32 SIMT arithmetic
  sse::PadTable pad_tc[2];          // the same in the arithmetic of lstm_ptable_kernel
  sse::GemmTower gemm_tw[2];        // K^T fp16 of the GEMM-per-step LSTM tower (lstm_gemm.cu)
  sse::TcTower tct[2];              // tensor-core copies of the LSTM weights (lazy, invalidated with the weights)
  __half* emb_f16 = nullptr;
  bool emb_f16_valid = false;
  float* emb_pad = nullptr;         // [V, We'] zero-padded fp32 embedding (padded shapes only)
  bool emb_pad_valid = false;
  sse::PaddedWeights padw[2];
  float learning_rate = 0.f;
  int64_t global_step = 0;

  // resident target index (this rank's shard)
  float* index_f32 = nullptr;
  bool index_owned = false;
  int64_t index_n = 0, index_off = 0;
  sse::TcIndex tc;

  // workspaces
  sse::Scratch enc_ws, search_ws, io_ws, train_ws, train_tc_ws, tok_ws, proj_ws;
  cudaStream_t train_side = nullptr;        // the target tower of the tensor-core train step runs here, next to the source tower
  cudaStream_t train_main = nullptr;        // ... and the source tower + loss here (forked from / joined into the caller's stream, so that the
                                            // step can be captured into a graph even when the caller uses the legacy default stream)
  cudaEvent_t train_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t train_chain[2] = {nullptr, nullptr};   // second row-half chain of each tower's recurrence
  cudaEvent_t train_ev2[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
  // the tensor-core train step as a CUDA graph (SSE_TRAIN_GRAPH=1): replayed while batch size and every buffer address stay the same
  cudaGraphExec_t train_graph = nullptr;
  unsigned long long train_graph_key[6] = {0, 0, 0, 0, 0, 0};
  unsigned long long train_seen_key[6] = {0, 0, 0, 0, 0, 0};
  long long train_graph_launches = 0;
  bool train_graph_failed = false;
  int* tok_bad = nullptr;           // device counter: out-of-range token ids seen by the pre-pass (sticky until read)
  float* grad_arena = nullptr;      // dense gradients, laid out by Param::grad_off, then the dense embedding gradient
  int64_t grad_floats = 0;          // dense (non-embedding) part
  int64_t arena_floats = 0;

  // device-resident training corpus of the batch sampler (sse_sampler_set)
  int32_t *smp_src = nullptr, *smp_ver = nullptr, *smp_tgt = nullptr;
  int64_t* smp_off = nullptr;
  int64_t smp_P = 0, smp_N = 0;

  int opt_search = 0, opt_encoder = 0;
  int opt_train = 0;                // 0 = auto (tensor cores with SSE_PRECISION_TC), 1 = fp32 SIMT (parity mode), 2 = tensor cores (bf16 operands)
  int opt_lstm_kernel = 0;          // 0 = auto; 1 = weight-streaming kernel (lstm_tc.cu); 2 / 3 = cluster kernels (lstm_cluster.cu); 4 = GEMM per step (lstm_gemm.cu)
  int opt_search_ctas = 0;          // 0 = all SMs; else cap on the scan grid (leaves SMs to a concurrent encoder)
  int opt_search_late = 0;          // with a cap: extra LATE scan items for the SMs the concurrent kernel frees mid-scan ...
  int opt_search_late_share = 40;   // ... each taking this percentage of a regular item's tile range
  int opt_cluster_rows = 0;         // batch rows per cluster of the table LSTM kernel: 0 = auto (64 when every cluster still gets its own SMs), 64, 128
  bool opt_pad_skip = true;         // per-tile pad-prefix start of the LSTM towers (tok_prep.cu)
  int64_t launches = 0;
};

namespace sse {
int find_param(sse_handle* h, const char* name);
void refresh_pointers(sse_handle* h);
bool side_is_cnn(const sse_handle* h, int side);
void invalidate_derived(sse_handle* h);
int encode_device(sse_handle* h, int side, const int32_t* tokens, int B, float* out, int normalize, cudaStream_t st);
int take_token_errors(sse_handle* h, cudaStream_t st, int* count);
}  // namespace sse
