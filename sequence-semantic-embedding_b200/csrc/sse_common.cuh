// Shared declarations of libsse_b200.so (internal; the public ABI is include/sse_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include <algorithm>

#include "../../include/sse_b200.h"

namespace sse {

void set_error(const char* fmt, ...);
const char* get_error();

#define SSE_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::sse::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SSE_ECUDA;                                                                \
    }                                                                                  \
  } while (0)

#define SSE_TRY(expr)              \
  do {                             \
    int _r = (expr);               \
    if (_r != SSE_OK) return _r;   \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct Param {
  std::string name;
  std::vector<int64_t> shape;
  float* dev = nullptr;
  int64_t numel = 0;
  bool trainable = true;
  int64_t grad_off = -1;   // offset (floats) of the dense gradient in the arena, -1 = none (sparse / non-trainable)
};

// One LSTM tower (TF layouts, device pointers into Params).
struct LstmTower {
  const float* K = nullptr;   // [We+H, 4H], column blocks i,j,f,o
  const float* b = nullptr;   // [4H]
  const float* M = nullptr;   // [H, E]
  int H = 0;
  int kparam = -1, bparam = -1, mparam = -1;   // indices into params
};

struct CnnTower {
  int nf = 0;
  int ksize[SSE_MAX_CNN_FILTERS];
  int nfilt[SSE_MAX_CNN_FILTERS];
  const float* W[SSE_MAX_CNN_FILTERS];   // [k, We, 1, F]
  const float* b[SSE_MAX_CNN_FILTERS];   // [F]
  int wparam[SSE_MAX_CNN_FILTERS], bparam[SSE_MAX_CNN_FILTERS];
  const float* M = nullptr;              // [sumF, E]
  int mparam = -1;
  int sumF = 0;
};

// growable device scratch buffer
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes);
  void release();
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// ---- kernels (implemented in the .cu files) --------------------------------
// lstm_simt.cu
int lstm_forward_simt(const int32_t* tokens, int B, int T, int t_start, const float* emb, int We,
                      const LstmTower& tw, float* h0, float* h1, float* c,      // [B,H] each
                      const float* init_h, const float* init_c,                // optional [H] broadcast initial state (pad-prefix table row)
                      float* save_h, float* save_c, float* save_g,             // optional training stash: [T,B,H],[T,B,H],[T,B,5H]
                      float** h_final, cudaStream_t st, int64_t* launches,
                      const int32_t* lead = nullptr, const float* pad_h = nullptr, const float* pad_c = nullptr);   // per-row pad-prefix start
int sgemm(bool ta, bool tb, int M, int N, int Kd, float alpha, const float* A, int lda, const float* Bm, int ldb,
          float beta, float* C, int ldc, cudaStream_t st, int64_t* launches, bool allow_split = false);
int l2norm_rows(float* x, int rows, int cols, cudaStream_t st, int64_t* launches);
int l2norm_rows_out(const float* x, float* y, int rows, int cols, cudaStream_t st, int64_t* launches);

// cnn.cu
int cnn_forward_ws(const int32_t* tokens, int B, int T, const float* emb, int We, const CnnTower& tw, float* xg,
                   float* conv, float* pool /*[B,sumF]*/, int32_t* argmax /*optional [B,sumF]*/, cudaStream_t st,
                   int64_t* launches);

// gemm_tc.cu: tcgen05 GEMM  D[M,N] = alpha A[M,K] B[N,K]^T (+ beta D), 16-bit K-major operands (fmt 0 fp16, 1 bf16)
int gemm_tc(const uint16_t* A, int64_t lda, const uint16_t* B, int64_t ldb, int M, int N, int K, float alpha, float beta, float* D, int64_t ldd,
            int fmt, int split_k, uint16_t* D16, int64_t ldd16, cudaStream_t st, int64_t* launches);
int convert_to_16(const float* src, int64_t rows, int cols, int64_t lds, uint16_t* dst, int64_t ldd, int fmt, cudaStream_t st, int64_t* launches);
int transpose_to_16(const float* src, int rows, int cols, int64_t lds, uint16_t* dst, int64_t ldd, int fmt, cudaStream_t st, int64_t* launches);
int gather_rows_16(const int32_t* tokens, int64_t n_tok, const float* emb, int We, int ld, uint16_t* out, int fmt, cudaStream_t st, int64_t* launches);
int cnn_conv_pool_tc(const uint16_t* X, int n_seq, int T, int ldx, int kf, const uint16_t* Wt, int F, const float* bias, float* pool, int pool_ld,
                     int pool_off, int fmt, cudaStream_t st, int64_t* launches);
// lstm_gemm.cu: LSTM tower as one gemm_tc call per time step (any We / H multiple of 8: cells wider than 256)
struct GemmTower {
  __half* kT16 = nullptr;   // [4H, We+H] fp16 = K^T
  bool valid = false;
};
bool lstm_gemm_supported(int We, int H);
int lstm_gemm_prepare(GemmTower& gt, const float* K, int We, int H, cudaStream_t st, int64_t* launches);
void lstm_gemm_release(GemmTower& gt);
size_t lstm_gemm_ws_bytes(int B, int T, int We, int H);
int lstm_forward_gemm(const int32_t* tokens, int B, int T, const float* emb, int We, int H, const GemmTower& gt, const float* bias, void* ws,
                      float* h_out, int ldh, cudaStream_t st, int64_t* launches);
// cnn.cu, tensor-core path
struct CnnTc {
  uint16_t* wt[SSE_MAX_CNN_FILTERS] = {};   // [F, k*We] fp16
  uint16_t* mt = nullptr;                   // [E, sumF] fp16
  bool valid = false;
};
bool cnn_tc_supported(int We, int T, const CnnTower& tw);
int cnn_tc_prepare(CnnTc& ct, const CnnTower& tw, int We, int E, cudaStream_t st, int64_t* launches);
void cnn_tc_release(CnnTc& ct);
size_t cnn_tc_ws_bytes(int nb, int T, int We, const CnnTower& tw);
int cnn_forward_tc(const int32_t* tokens, int nb, int T, const float* emb, int We, int E, const CnnTower& tw, const CnnTc& ct, void* ws,
                   float* proj, cudaStream_t st, int64_t* launches);

// search_simt.cu
int search_simt(const float* q, int Q, int E, const float* index, int64_t N, int64_t global_offset, int k,
                float* out_scores, int32_t* out_idx, Scratch& ws, int num_sms, cudaStream_t st, int64_t* launches,
                int out_stride = 0);   // row stride of out_scores / out_idx in elements (0 = k)
int merge_topk_strided(const float* cand_s, const int32_t* cand_i, int Q, int n_groups, int64_t group_stride, int row_stride,
                       int k_in, int k, float* out_s, int32_t* out_i, int out_stride, cudaStream_t st, int64_t* launches);
int merge_topk(const float* cand_s, const int32_t* cand_i, int Q, int C, int k, float* out_s, int32_t* out_i,
               cudaStream_t st, int64_t* launches);

// search_tc.cu (tcgen05 fp16 scan + exact fp32 re-rank)
struct TcIndex {
  __half* h16 = nullptr;    // [N, E] row-major fp16
  int64_t N = 0;
  int E = 0;
  alignas(64) unsigned char tmap[128];     // CUtensorMap of the fp16 index, box 64 x 128 rows
  alignas(64) unsigned char tmap64[128];   // same, box 64 x 64 rows
  alignas(64) unsigned char tmap3d[128];   // 3-D (k-in-block, row, k-block), box = whole [128 x E] tile
  alignas(64) unsigned char tmap3d64[128]; // 3-D, box = whole [64 x E] tile
  bool use3d = false;
  bool tmap_ok = false;
  unsigned* group_ctr = nullptr;    // fused scan: per-m-group barrier counters (device, zero between searches)
  // layout of the last search_tc call's candidate bookkeeping (search_tc_stats)
  const int32_t* last_cnt = nullptr; int64_t last_cnt_n = 0; const int32_t* last_overflow = nullptr; int last_Q = 0, last_items = 0;
};
bool search_tc_supported(int E, int64_t N, int k);
int search_tc_max_rows(int E);   // query rows one search_tc call accepts
int search_tc_prepare(TcIndex& ti, const float* index_f32, int64_t N, int E, cudaStream_t st, int64_t* launches);
void search_tc_release(TcIndex& ti);
int search_tc(const float* q, int Q, int E, const float* index_f32, TcIndex& ti, int64_t global_offset, int k,
              float* out_scores, int32_t* out_idx, Scratch& ws, int num_sms, cudaStream_t st, int64_t* launches,
              int out_stride = 0, int late_ctas = 0, int late_share = 0);
// candidates / fallback rows of the LAST search_tc call on this index (synchronises; for benchmarks and tests)
struct SearchStats { long long candidates = 0; int rows = 0, fallback_rows = 0, items = 0; };
int search_tc_stats(const TcIndex& ti, SearchStats* out);

// tok_prep.cu: token range check + pad-prefix bucketing (device pre-pass of the LSTM encoders)
struct TokPrep {
  int32_t* stok = nullptr;          // [B,T] sanitised tokens, rows in SORTED order
  int32_t* perm = nullptr;          // [B] original row of sorted position
  int32_t* lead_sorted = nullptr;   // [B] leading PADs (<= T-1) of each sorted row, descending when `sorted`
  bool sorted = false;
};
size_t tok_prep_ws_bytes(int B, int T);
int tok_prep(const int32_t* tokens, int B, int T, int V, bool sort, void* ws, TokPrep* out, int* bad_total, cudaStream_t st,
             int64_t* launches);
int sanitize_tokens_inplace(int32_t* tokens, int64_t n, int V, int* bad_total, cudaStream_t st, int64_t* launches);
int sample_train_batch(const int32_t* src_rows, const int64_t* ver_off, const int32_t* ver_rows, const int32_t* tgt_rows, int64_t N, int T, int64_t start,
                       int B, uint64_t seed, uint64_t step, int32_t* src_out, int32_t* tgt_out, float* lab_out, cudaStream_t st, int64_t* launches);
int unpermute_rows(const float* x, float* y, const int32_t* perm, int rows, int cols, int normalize, cudaStream_t st, int64_t* launches);
// out[perm ? perm[r] : r] = (l2-normalised) hmat[r] M  -- projection + normalisation + un-permutation of a query batch in one launch
bool project_rows_supported(int rows, int H, int E);
int project_rows(const float* hmat, int ldh, const float* M, int H, int E, int rows, const int32_t* perm, int normalize, float* out,
                 cudaStream_t st, int64_t* launches);
// per-tile pad-prefix start of the tensor-core LSTM kernels (all null = off)
struct PadSkip {
  const int32_t* lead_sorted = nullptr;
  const float* pad_h = nullptr;     // [T][H], entry t = state after t+1 leading PADs
  const float* pad_c = nullptr;
  float* dump_h = nullptr;          // table generation (lstm_ptable_kernel only)
  float* dump_c = nullptr;
};

// lstm_tc.cu (tcgen05 LSTM tower, fp16 operands / fp32 accumulate and state)
struct TcTower {
  __half* wt = nullptr;     // [4H, We+H] chunk-major transposed weights
  float* bias_r = nullptr;         // [4H] chunk-major bias (+1 folded into the forget gate)
  alignas(64) unsigned char tmap[128];     // 3-D map (k-in-block, gate row, k-block), box = [128 rows x slot_kb k-blocks]
  int slot_kb = 2;
  alignas(64) unsigned char tmap2d[128];   // 2-D map, box = one [128 x 64] sub-tile (fallback)
  bool use3d = true;
  bool valid = false;
  float* ptable = nullptr;         // [V, 4H] fp32 input-projection table (lstm_cluster.cu variant 2); rebuilt when the weights change
  int64_t ptable_rows = 0;
  float* wxp = nullptr;            // [We, 4H] permuted + scaled W_x (GEMM operand of the table build)
  bool ptable_valid = false;
  int ptable_mode = 0;             // gate-math variant the table was scaled for
  int ptable_wide = 32;            // 0: table laid out for lstm_ptable_kernel; 32 / 64: for lstm_wide_kernel with that many units per CTA
  int ptable_ew = 16;              // epilogue warps of the kernel variant the table's column order was built for
};
bool lstm_tc_supported(int We, int H);
int lstm_tc_prepare(TcTower& tt, const float* K, const float* b, int We, int H, cudaStream_t st, int64_t* launches);
void lstm_tc_release(TcTower& tt);
int lstm_forward_tc(const int32_t* tokens, int B, int T, int t_start, const __half* emb_f16, int We, int H,
                    const TcTower& tt, const float* init_h, const float* init_c, const PadSkip& ps, float* c_scratch, float* h_out,
                    cudaStream_t st, int64_t* launches);

// cluster variant for small / medium batches (lstm_cluster.cu): weights resident in the shared memory of a cluster
bool lstm_cluster_supported(int We, int H);
int lstm_forward_cluster(const int32_t* tokens, int B, int T, int t_start, const __half* emb_f16, int We, int H,
                         const TcTower& tt, const float* init_h, const float* init_c, const PadSkip& ps, float* h_out, cudaStream_t st,
                         int64_t* launches);

// variant 2 (lstm_cluster.cu): input projection tabulated per vocabulary entry, 128 rows per cluster
bool lstm_ptable_supported(int64_t V, int We, int H);
int lstm_ptable_prepare(TcTower& tt, const float* emb, int64_t V, const float* K, int We, int H, cudaStream_t st, int64_t* launches);
void lstm_ptable_release(TcTower& tt);
int lstm_forward_ptable(const int32_t* tokens, int B, int T, int t_start, int We, int H, const TcTower& tt, const float* init_h,
                        const float* init_c, const PadSkip& ps, float* h_out, cudaStream_t st, int64_t* launches, int cluster_rows = 0,
                        int num_sms = 148);

// small utilities (util.cu)
int fill_f32(float* p, int64_t n, float v, cudaStream_t st, int64_t* launches);
int f32_to_f16(const float* src, __half* dst, int64_t n, cudaStream_t st, int64_t* launches);
int pad_cols(const float* src, int64_t R, int C, float* dst, int Cp, cudaStream_t st, int64_t* launches);
int pad_lstm_weights(const float* K, const float* b, int We, int H, int Wp, int Hp, float* Kp, float* bp, cudaStream_t st, int64_t* launches);

}  // namespace sse
