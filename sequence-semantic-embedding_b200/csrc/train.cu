// Training step of the dual / shared LSTM encoder (fp32).
//
// Restates sess.run([model.train, model.loss, model.train_acc]) of the reference
// (sse_train.py:170-172):  forward both towers -> l2-normalise -> per-pair cosine
// (sse_model.py:282-290) -> mean weighted sigmoid CE on 64*cos, train_acc (:298,302) ->
// tf.gradients (BPTT through static_rnn) -> clip_by_global_norm(5.0) -> Adagrad
// (initial accumulator 0.1, no epsilon) -> global_step += 1 (:355-364).
//
// Gradient arena (one flat fp32 buffer, the unit the data-parallel all-reduce moves):
//   [ dense grads of every trainable variable except word_embedding | dense word_embedding
//     grad [V,We] | touched-row counts [V] | scalars[8] ]
//   scalars: 0 sum of squares (dense grads + UN-merged embedding slices)
//            1 loss (already / B_global)  2 acc-positive term  3 acc-negative term
#include "sse_common.cuh"
#include "sse_handle.cuh"
#include <math_constants.h>
#include <cuda_bf16.h>

using namespace sse;

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// warp per pair row: normalise both encodings, cosine, loss terms and d loss / d u.
__global__ void pair_loss_kernel(const float* __restrict__ uS, const float* __restrict__ uT,
                                 const float* __restrict__ labels, int B, int E, float inv_bglobal,
                                 float* __restrict__ duS, float* __restrict__ duT, float* __restrict__ row_loss,
                                 float* __restrict__ row_accp, float* __restrict__ row_accn, float* __restrict__ cos_out) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* a = uS + (size_t)row * E;
  const float* b = uT + (size_t)row * E;
  float ssa = 0.f, ssb = 0.f, dab = 0.f;
  for (int j = lane; j < E; j += 32) {
    float x = a[j], y = b[j];
    ssa = fmaf(x, x, ssa); ssb = fmaf(y, y, ssb); dab = fmaf(x, y, dab);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    ssa += __shfl_xor_sync(0xffffffffu, ssa, o);
    ssb += __shfl_xor_sync(0xffffffffu, ssb, o);
    dab += __shfl_xor_sync(0xffffffffu, dab, o);
  }
  float ia = 1.0f / sqrtf(fmaxf(ssa, 1e-12f)), ib = 1.0f / sqrtf(fmaxf(ssb, 1e-12f));
  float cosv = dab * ia * ib;
  if (cos_out && lane == 0) cos_out[row] = cosv;
  if (!duS) return;
  float l = labels[row];
  float x = 64.0f * cosv;
  float s = sigmoidf_(x);
  if (lane == 0) {
    // weighted_cross_entropy_with_logits, pos_weight = 1: (1-l)*x + log1p(exp(-|x|)) + max(-x, 0)
    row_loss[row] = ((1.0f - l) * x + log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f)) * inv_bglobal;
    row_accp[row] = l * floorf(s + 0.1f) * inv_bglobal;
    row_accn[row] = (1.0f - l) * floorf(1.1f - s) * inv_bglobal;
  }
  float dcos = (s - l) * 64.0f * inv_bglobal;
  // n = u * inv ; dn_a = dcos * n_b ; du_a = (dn_a - n_a * <dn_a, n_a>) * inv_a = dcos*(n_b - n_a*cos)*inv_a
  for (int j = lane; j < E; j += 32) {
    float na = a[j] * ia, nb = b[j] * ib;
    duS[(size_t)row * E + j] = dcos * (nb - na * cosv) * ia;
    duT[(size_t)row * E + j] = dcos * (na - nb * cosv) * ib;
  }
}

// deterministic sums of three [B] arrays into scalars[1..3]
__global__ void reduce_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                   int n, float* __restrict__ scalars) {
  __shared__ float sh[3][256];
  float sa = 0.f, sb = 0.f, sc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) { sa += a[i]; sb += b[i]; sc += c[i]; }
  sh[0][threadIdx.x] = sa; sh[1][threadIdx.x] = sb; sh[2][threadIdx.x] = sc;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; sh[2][threadIdx.x] += sh[2][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { scalars[1] += sh[0][0]; scalars[2] += sh[1][0]; scalars[3] += sh[2][0]; }
}

// one step of the backward recurrence: dz[t] from (dh, dc) and the stashed activations.
__global__ void lstm_bwd_gates_kernel(const float* __restrict__ dh, int ldh, float* __restrict__ dc,
                                      const float* __restrict__ g /*[B,5H] at t: si,tj,sf,so,tc*/,
                                      const float* __restrict__ c_prev /*[B,H] or null*/, int B, int H,
                                      float* __restrict__ dz /*[B,4H]*/) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  int b = i / H, u = i - b * H;
  const float* gr = g + (size_t)b * 5 * H + u;
  float si = gr[0], tj = gr[H], sf = gr[2 * H], so = gr[3 * H], tc = gr[4 * H];
  float dhv = dh[(size_t)b * ldh + u];
  float dcv = dc[i] + dhv * so * (1.f - tc * tc);
  float cp = c_prev ? c_prev[i] : 0.f;
  float* z = dz + (size_t)b * 4 * H + u;
  z[0] = dcv * tj * si * (1.f - si);
  z[H] = dcv * si * (1.f - tj * tj);
  z[2 * H] = dcv * cp * sf * (1.f - sf);
  z[3 * H] = dhv * tc * so * (1.f - so);
  dc[i] = dcv * sf;
}

// out[j] += sum_r x[r][j]   (block = 32 columns x 8 row lanes; deterministic)
__global__ void colsum_kernel(const float* __restrict__ x, int64_t rows, int cols, float* __restrict__ out) {
  __shared__ float sh[8][33];
  int col = blockIdx.x * 32 + (threadIdx.x & 31);
  int rl = threadIdx.x >> 5;
  float s = 0.f;
  if (col < cols)
    for (int64_t r = rl; r < rows; r += 8) s += x[(size_t)r * cols + col];
  sh[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x & 31];
    out[col] += t;
  }
}

// time-major gather: xg[(t*B + b)*We + e] = emb[tok[b][t]][e]
__global__ void gather_time_major_kernel(const int32_t* __restrict__ tok, int B, int T, const float* __restrict__ emb,
                                         int We, float* __restrict__ xg) {
  int64_t total = (int64_t)B * T * We;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / We;
    int e = (int)(i - r * We);
    int t = (int)(r / B), b = (int)(r - (int64_t)t * B);
    xg[i] = __ldg(emb + (size_t)tok[(size_t)b * T + t] * We + e);
  }
}

// embedding IndexedSlices: G[tok] += dX (merged, for Adagrad), touched[tok] += 1, sumsq += |dX|^2 (UN-merged)
__global__ void embed_scatter_kernel(const int32_t* __restrict__ tok, int B, int T, const float* __restrict__ dxh, int ld,
                                     int We, float* __restrict__ G, float* __restrict__ touched, float* __restrict__ sumsq) {
  __shared__ float sh[256];
  int64_t total = (int64_t)B * T * We;
  float ss = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / We;
    int e = (int)(i - r * We);
    int t = (int)(r / B), b = (int)(r - (int64_t)t * B);
    int id = tok[(size_t)b * T + t];
    float v = dxh[(size_t)r * ld + e];
    ss = fmaf(v, v, ss);
    atomicAdd(G + (size_t)id * We + e, v);
    if (e == 0) atomicAdd(touched + id, 1.0f);
  }
  sh[threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(sumsq, sh[0]);
}

__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ sumsq) {
  __shared__ float sh[256];
  float ss = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { float v = x[i]; ss = fmaf(v, v, ss); }
  sh[threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(sumsq, sh[0]);
}

// g <- g * 5/max(|g|,5) ; acc += g^2 ; w -= lr * g / sqrt(acc)
__global__ void adagrad_dense_kernel(float* __restrict__ w, float* __restrict__ acc, const float* __restrict__ g, int64_t n,
                                     const float* __restrict__ scalars, float lr, float max_norm) {
  float gn = sqrtf(scalars[0]);
  float scale = max_norm / fmaxf(gn, max_norm);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gv = g[i] * scale;
    float a = acc[i] + gv * gv;
    acc[i] = a;
    w[i] -= lr * gv / sqrtf(a);
  }
}

// sparse apply: only rows that appeared in a lookup (duplicates already summed in G)
__global__ void adagrad_rows_kernel(float* __restrict__ w, float* __restrict__ acc, const float* __restrict__ G,
                                    const float* __restrict__ touched, int V, int We, const float* __restrict__ scalars,
                                    float lr, float max_norm) {
  float gn = sqrtf(scalars[0]);
  float scale = max_norm / fmaxf(gn, max_norm);
  int64_t total = (int64_t)V * We;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int row = (int)(i / We);
    if (touched[row] == 0.f) continue;
    float gv = G[i] * scale;
    float a = acc[i] + gv * gv;
    acc[i] = a;
    w[i] -= lr * gv / sqrtf(a);
  }
}

struct TowerStash { float *sh, *sc, *sg, *u, *h0, *h1, *c, *hlast; };

int ensure_arena(sse_handle* h) {
  if (h->grad_arena) return SSE_OK;
  const int64_t V = h->cfg.vocab_size, We = h->cfg.embedding_size;
  h->arena_floats = h->grad_floats + V * We + V + 8;
  cudaError_t e = cudaMalloc(&h->grad_arena, (size_t)h->arena_floats * 4);
  if (e != cudaSuccess) { set_error("cudaMalloc(grad arena) failed: %s", cudaGetErrorString(e)); return SSE_ENOMEM; }
  return SSE_OK;
}

bool trainable_mode(const sse_handle* h) {
  return h->cfg.network_mode == SSE_MODE_DUAL_ENCODER || h->cfg.network_mode == SSE_MODE_SHARED_ENCODER;
}

// ======================================================================================================================
// Tensor-core training path (bf16 operands, fp32 accumulate / state / stash; BASELINE config 4).  Same graph as the fp32
// path; every contraction is a gemm_tc.cu GEMM over TIME-MAJOR 16-bit operand copies:
//   forward   ZX = X Wx^T (one GEMM over all T*B rows), then per step z_t = ZX[t] += h_{t-1} Wh^T and a gate kernel;
//   backward  per step dz_t (gate kernel, written as bf16) and dh_{t-1} = dz_t K[We:]^T; after the loop
//             dX = dZ K[:We]^T (one GEMM), dK[:We] += X^T dZ and dK[We:] += Hprev^T dZ (split-K GEMMs over T*B rows with
//             atomic fp32 accumulation into the gradient arena), db = column sums of dZ.
struct TcTrainBufs {
  uint16_t *x16, *h16, *dz16, *xT, *hT, *dzT, *kT16, *k16;
  float *zx, *dx, *dh, *dc, *c_zero;
  int32_t* tok_tm;
};

__global__ void tokens_time_major_kernel(const int32_t* __restrict__ tok, int B, int T, int32_t* __restrict__ out) {
  const int64_t total = (int64_t)B * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / B), b = (int)(i - (int64_t)t * B);
    out[i] = tok[(size_t)b * T + t];
  }
}

__device__ __forceinline__ uint16_t bf16_bits(float v) {
  __nv_bfloat16 b = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&b);
}

// z [B,4H] (pre-activations without bias, TF gate order i,j,f,o) -> gates stash, c, h (bf16 operand copy + optional fp32)
__global__ void lstm_fwd_gates_kernel(const float* __restrict__ z, const float* __restrict__ bias, const float* __restrict__ c_prev, int B, int H,
                                      float* __restrict__ sg /*[B,5H]*/, float* __restrict__ c_out /*[B,H]*/, uint16_t* __restrict__ h16 /*[B,H]*/,
                                      float* __restrict__ h32 /*optional [B,H]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, u = i - b * H;
  const float* zr = z + (size_t)b * 4 * H + u;
  const float si = sigmoidf_(zr[0] + __ldg(bias + u));
  const float tj = tanhf(zr[H] + __ldg(bias + H + u));
  const float sf = sigmoidf_(zr[2 * H] + __ldg(bias + 2 * H + u) + 1.0f);      // forget_bias = 1.0 added at run time
  const float so = sigmoidf_(zr[3 * H] + __ldg(bias + 3 * H + u));
  const float cn = (c_prev ? c_prev[i] : 0.f) * sf + si * tj;
  const float tc = tanhf(cn);
  const float hn = tc * so;
  float* g = sg + (size_t)b * 5 * H + u;
  g[0] = si; g[H] = tj; g[2 * H] = sf; g[3 * H] = so; g[4 * H] = tc;
  c_out[i] = cn;
  h16[i] = bf16_bits(hn);
  if (h32) h32[i] = hn;
}

// backward gate step writing dz as the bf16 operand of the three gradient GEMMs
__global__ void lstm_bwd_gates16_kernel(const float* __restrict__ dh, int ldh, float* __restrict__ dc, const float* __restrict__ g,
                                        const float* __restrict__ c_prev, int B, int H, uint16_t* __restrict__ dz16 /*[B,4H]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, u = i - b * H;
  const float* gr = g + (size_t)b * 5 * H + u;
  const float si = gr[0], tj = gr[H], sf = gr[2 * H], so = gr[3 * H], tc = gr[4 * H];
  const float dhv = dh[(size_t)b * ldh + u];
  const float dcv = dc[i] + dhv * so * (1.f - tc * tc);
  const float cp = c_prev ? c_prev[i] : 0.f;
  uint16_t* zr = dz16 + (size_t)b * 4 * H + u;
  zr[0] = bf16_bits(dcv * tj * si * (1.f - si));
  zr[H] = bf16_bits(dcv * si * (1.f - tj * tj));
  zr[2 * H] = bf16_bits(dcv * cp * sf * (1.f - sf));
  zr[3 * H] = bf16_bits(dhv * tc * so * (1.f - so));
  dc[i] = dcv * sf;
}

// 16-bit tiled transpose: dst[c][r] = src[r][c], dst leading dimension ldd >= rows (pad zeroed)
__global__ void transpose_16_kernel(const uint16_t* __restrict__ s, int64_t rows, int cols, int64_t lds, uint16_t* __restrict__ d, int64_t ldd) {
  __shared__ uint16_t tile[32][34];
  const int64_t r0 = (int64_t)blockIdx.y * 32;
  const int c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? s[(size_t)r * lds + c] : (uint16_t)0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t r = r0 + threadIdx.x;
    if (c < cols && r < ldd) d[(size_t)c * ldd + r] = tile[threadIdx.x][i];
  }
}

// out[j] += sum_r x[r][j] over bf16 rows (db = column sums of dZ); rows are split over blockIdx.y (atomic accumulation:
// 32 column blocks alone left the other 100+ SMs idle for 0.9 ms)
__global__ void colsum16_kernel(const uint16_t* __restrict__ x, int64_t rows, int cols, float* __restrict__ out) {
  __shared__ float sh[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  float s = 0.f;
  if (col < cols)
    for (int64_t r = (int64_t)blockIdx.y * 8 + rl; r < rows; r += (int64_t)gridDim.y * 8) {
      const uint32_t bits = (uint32_t)x[(size_t)r * cols + col] << 16;
      s += __uint_as_float(bits);
    }
  sh[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x & 31];
    atomicAdd(out + col, t);
  }
}

bool train_tc_supported(const sse_handle* h, int B) {
  const int We = h->cfg.embedding_size;
  return We % 8 == 0 && h->lstm[0].H % 8 == 0 && h->lstm[1].H % 8 == 0 && B % 8 == 0;
}

static int transpose16(const uint16_t* s, int64_t rows, int cols, int64_t lds, uint16_t* d, int64_t ldd, cudaStream_t st, int64_t* launches) {
  dim3 grid(cdiv(cols, 32), (unsigned)cdiv64(ldd, 32)), block(32, 8);
  transpose_16_kernel<<<grid, block, 0, st>>>(s, rows, cols, lds, d, ldd);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

// forward + backward of ONE tower on the tensor cores.  u [B,E] out (forward), then du -> grads into the arena.
// Split in two calls because the pair loss needs both towers' forward results first.
struct TcTowerState {
  float *sg, *sc, *u, *hlast;      // stash [T,B,5H], [T,B,H]; projection; fp32 h_T
  uint16_t* h16;                   // [T,B,H] bf16
  uint16_t* x16;                   // [T,B,We] bf16
};

// `st_b` (optional): the recurrence of the second half of the batch rows runs there -- rows are independent, so a tower's T-step
// chain of small dependent launches becomes two chains that fill each other's bubbles (ev: fork / join events)
int train_tc_forward(sse_handle* h, int s, const int32_t* tok, int B, uint8_t* ws, size_t* ws_off, TcTowerState* ts, uint16_t* kT16,
                     float* zx, cudaStream_t st, cudaStream_t st_b = nullptr, cudaEvent_t* ev = nullptr) {
  const sse_config& c = h->cfg;
  const int T = c.max_seq_length, We = c.embedding_size, E = c.encoding_size;
  const LstmTower& tw = h->lstm[s];
  const int H = tw.H, ld = We + H;
  const int64_t TB = (int64_t)T * B;
  auto carve = [&](size_t bytes) { size_t o = *ws_off; *ws_off = (o + bytes + 255) / 256 * 256; return ws + o; };
  ts->sg = reinterpret_cast<float*>(carve((size_t)TB * 5 * H * 4));
  ts->sc = reinterpret_cast<float*>(carve((size_t)TB * H * 4));
  ts->u = reinterpret_cast<float*>(carve((size_t)B * E * 4));
  ts->hlast = reinterpret_cast<float*>(carve((size_t)B * H * 4));
  ts->h16 = reinterpret_cast<uint16_t*>(carve((size_t)TB * H * 2));
  ts->x16 = reinterpret_cast<uint16_t*>(carve((size_t)TB * We * 2));
  int32_t* tok_tm = reinterpret_cast<int32_t*>(carve((size_t)TB * 4));
  tokens_time_major_kernel<<<(int)std::min<int64_t>(cdiv64(TB, 256), 148 * 4), 256, 0, st>>>(tok, B, T, tok_tm);
  ++h->launches;
  SSE_TRY(gather_rows_16(tok_tm, TB, h->params[h->emb_param].dev, We, We, ts->x16, 1, st, &h->launches));
  // K^T [4H, We+H] bf16: B operand of both forward GEMMs
  SSE_TRY(transpose_to_16(tw.K, ld, 4 * H, 4 * H, kT16, ld, 1, st, &h->launches));
  // ZX = X Wx^T over all T*B rows
  SSE_TRY(gemm_tc(ts->x16, We, kT16, ld, (int)TB, 4 * H, We, 1.f, 0.f, zx, 4 * H, 1, 1, nullptr, 0, st, &h->launches));
  const bool chains = st_b != nullptr && ev != nullptr && B >= 32;
  const int nrow[2] = {chains ? (B / 2 + 7) / 8 * 8 : B, chains ? B - (B / 2 + 7) / 8 * 8 : 0};
  cudaStream_t cst[2] = {st, chains ? st_b : st};
  if (chains) { SSE_CUDA_OK(cudaEventRecord(ev[0], st)); SSE_CUDA_OK(cudaStreamWaitEvent(st_b, ev[0], 0)); }
  for (int t = 0; t < T; ++t) {
    for (int cix = 0; cix < (chains ? 2 : 1); ++cix) {
      const size_t b0 = cix ? (size_t)nrow[0] : 0;
      const int nb = nrow[cix];
      float* z_t = zx + ((size_t)t * B + b0) * 4 * H;
      if (t > 0)   // z_t += h_{t-1} Wh^T
        SSE_TRY(gemm_tc(ts->h16 + ((size_t)(t - 1) * B + b0) * H, H, kT16 + We, ld, nb, 4 * H, H, 1.f, 1.f, z_t, 4 * H, 1, 1, nullptr, 0, cst[cix], &h->launches));
      lstm_fwd_gates_kernel<<<cdiv(nb * H, 256), 256, 0, cst[cix]>>>(z_t, tw.b, t > 0 ? ts->sc + ((size_t)(t - 1) * B + b0) * H : nullptr, nb, H,
                                                                  ts->sg + ((size_t)t * B + b0) * 5 * H, ts->sc + ((size_t)t * B + b0) * H,
                                                                  ts->h16 + ((size_t)t * B + b0) * H, t == T - 1 ? ts->hlast + b0 * H : nullptr);
      ++h->launches;
    }
  }
  if (chains) { SSE_CUDA_OK(cudaEventRecord(ev[1], st_b)); SSE_CUDA_OK(cudaStreamWaitEvent(st, ev[1], 0)); }
  if (project_rows_supported(B, H, E))       // u = h_T M in one launch (the SIMT GEMM takes 68 us for 1536 x 256 x 256)
    SSE_TRY(project_rows(ts->hlast, H, tw.M, H, E, B, nullptr, 0, ts->u, st, &h->launches));
  else
    SSE_TRY(sgemm(false, false, B, E, H, 1.f, ts->hlast, H, tw.M, E, 0.f, ts->u, E, st, &h->launches, true));
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

int train_tc_backward(sse_handle* h, int s, const int32_t* tok, int B, const TcTowerState& ts, const float* du, uint8_t* ws, size_t ws_off,
                      uint16_t* k16, float* G, float* touched, float* scalars, cudaStream_t st, cudaStream_t st_b = nullptr, cudaEvent_t* ev = nullptr) {
  const sse_config& c = h->cfg;
  const int T = c.max_seq_length, We = c.embedding_size, E = c.encoding_size;
  const LstmTower& tw = h->lstm[s];
  const int H = tw.H, ld = We + H;
  const int64_t TB = (int64_t)T * B;
  const int64_t ldT = (TB + 7) / 8 * 8;
  float* arena = h->grad_arena;
  float* gM = arena + h->params[tw.mparam].grad_off;
  float* gK = arena + h->params[tw.kparam].grad_off;
  float* gb = arena + h->params[tw.bparam].grad_off;
  auto carve = [&](size_t bytes) { size_t o = ws_off; ws_off = (o + bytes + 255) / 256 * 256; return ws + o; };
  uint16_t* dz16 = reinterpret_cast<uint16_t*>(carve((size_t)TB * 4 * H * 2));
  uint16_t* dzT = reinterpret_cast<uint16_t*>(carve((size_t)4 * H * ldT * 2));
  uint16_t* xT = reinterpret_cast<uint16_t*>(carve((size_t)We * ldT * 2));
  uint16_t* hT = reinterpret_cast<uint16_t*>(carve((size_t)H * ldT * 2));
  float* dx = reinterpret_cast<float*>(carve((size_t)TB * We * 4));
  float* dh = reinterpret_cast<float*>(carve((size_t)2 * B * H * 4));
  float* dc = reinterpret_cast<float*>(carve((size_t)B * H * 4));
  // dM += h_last^T du ; dh_T = du M^T   (small: fp32 SIMT)
  SSE_TRY(sgemm(true, false, H, E, B, 1.f, ts.hlast, H, du, E, 1.f, gM, E, st, &h->launches, true));
  SSE_TRY(sgemm(false, true, B, H, E, 1.f, du, E, tw.M, E, 0.f, dh, H, st, &h->launches, true));
  SSE_CUDA_OK(cudaMemsetAsync(dc, 0, (size_t)B * H * 4, st));
  // K [We+H, 4H] bf16: B operand (N = rows of K, K-dim = 4H) of dz K^T
  SSE_TRY(convert_to_16(tw.K, ld, 4 * H, 4 * H, k16, 4 * H, 1, st, &h->launches));
  float* dh_cur = dh;
  float* dh_next = dh + (size_t)B * H;
  const bool chains = st_b != nullptr && ev != nullptr && B >= 32;
  const int nrow[2] = {chains ? (B / 2 + 7) / 8 * 8 : B, chains ? B - (B / 2 + 7) / 8 * 8 : 0};
  cudaStream_t cst[2] = {st, chains ? st_b : st};
  if (chains) { SSE_CUDA_OK(cudaEventRecord(ev[2], st)); SSE_CUDA_OK(cudaStreamWaitEvent(st_b, ev[2], 0)); }
  for (int t = T - 1; t >= 0; --t) {
    for (int cix = 0; cix < (chains ? 2 : 1); ++cix) {
      const size_t b0 = cix ? (size_t)nrow[0] : 0;
      const int nb = nrow[cix];
      uint16_t* dz_t = dz16 + ((size_t)t * B + b0) * 4 * H;
      lstm_bwd_gates16_kernel<<<cdiv(nb * H, 256), 256, 0, cst[cix]>>>(dh_cur + b0 * H, H, dc + b0 * H, ts.sg + ((size_t)t * B + b0) * 5 * H,
                                                                    t > 0 ? ts.sc + ((size_t)(t - 1) * B + b0) * H : nullptr, nb, H, dz_t);
      ++h->launches;
      if (t > 0)    // dh_{t-1} = dz_t K[We:]^T
        SSE_TRY(gemm_tc(dz_t, 4 * H, k16 + (size_t)We * 4 * H, 4 * H, nb, H, 4 * H, 1.f, 0.f, dh_next + b0 * H, H, 1, 1, nullptr, 0, cst[cix], &h->launches));
    }
    if (t > 0) std::swap(dh_cur, dh_next);
  }
  if (chains) { SSE_CUDA_OK(cudaEventRecord(ev[3], st_b)); SSE_CUDA_OK(cudaStreamWaitEvent(st, ev[3], 0)); }
  // dX = dZ K[:We]^T over all rows; embedding IndexedSlices from it
  SSE_TRY(gemm_tc(dz16, 4 * H, k16, 4 * H, (int)TB, We, 4 * H, 1.f, 0.f, dx, We, 1, 1, nullptr, 0, st, &h->launches));
  embed_scatter_kernel<<<148 * 4, 256, 0, st>>>(tok, B, T, dx, We, We, G, touched, scalars);
  ++h->launches;
  // dK[:We] += X^T dZ ;  dK[We:] += Hprev^T dZ (t >= 1) ; db += colsum dZ : K-major copies = transposes over the T*B rows
  SSE_TRY(transpose16(dz16, TB, 4 * H, 4 * H, dzT, ldT, st, &h->launches));
  SSE_TRY(transpose16(ts.x16, TB, We, We, xT, ldT, st, &h->launches));
  const int splits = 8;
  SSE_TRY(gemm_tc(xT, ldT, dzT, ldT, We, 4 * H, (int)TB, 1.f, 1.f, gK, 4 * H, 1, splits, nullptr, 0, st, &h->launches));
  if (T > 1) {
    SSE_TRY(transpose16(ts.h16, TB - B, H, H, hT, ldT, st, &h->launches));
    SSE_TRY(gemm_tc(hT, ldT, dzT + B, ldT, H, 4 * H, (int)(TB - B), 1.f, 1.f, gK + (size_t)We * 4 * H, 4 * H, 1, splits, nullptr, 0, st, &h->launches));
  }
  colsum16_kernel<<<dim3(cdiv(4 * H, 32), 32), 256, 0, st>>>(dz16, TB, 4 * H, gb);
  ++h->launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

}  // namespace

extern "C" {

int sse_pair_score(sse_handle* h, const int32_t* src_dev, const int32_t* tgt_dev, int B, float* cos_dev, void* stream) {
  if (!h || !src_dev || !tgt_dev || !cos_dev || B < 0) { set_error("sse_pair_score: bad argument"); return SSE_EINVAL; }
  if (B == 0) return SSE_OK;
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int E = h->cfg.encoding_size;
  SSE_TRY(h->train_ws.ensure((size_t)2 * B * E * 4));
  float* uS = h->train_ws.as<float>();
  float* uT = uS + (size_t)B * E;
  SSE_TRY(encode_device(h, SSE_SIDE_SRC, src_dev, B, uS, 0, st));
  if (h->tgt_table_param >= 0) { set_error("sse_pair_score: target side is a table in this mode"); return SSE_ESTATE; }
  SSE_TRY(encode_device(h, SSE_SIDE_TGT, tgt_dev, B, uT, 0, st));
  pair_loss_kernel<<<cdiv(B, 8), 256, 0, st>>>(uS, uT, nullptr, B, E, 0.f, nullptr, nullptr, nullptr, nullptr, nullptr, cos_dev);
  ++h->launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

int sse_train_grads(sse_handle* h, const int32_t* src, const int32_t* tgt, const float* labels, int B, int B_global,
                    float* loss_host, float* acc_host, void* stream) {
  if (!h || !src || !tgt || !labels || B < 1 || B_global < B) { set_error("sse_train_grads: bad argument"); return SSE_EINVAL; }
  if (!trainable_mode(h)) { set_error("training is implemented for dual-encoder / shared-encoder (the modes the reference can train at HEAD)"); return SSE_ESTATE; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const sse_config& c = h->cfg;
  const int T = c.max_seq_length, We = c.embedding_size, E = c.encoding_size, V = c.vocab_size;
  SSE_TRY(ensure_arena(h));
  float* arena = h->grad_arena;
  float* G = arena + h->grad_floats;
  float* touched = G + (size_t)V * We;
  float* scalars = touched + V;
  SSE_CUDA_OK(cudaMemsetAsync(arena, 0, (size_t)h->arena_floats * 4, st));

  const int Hmax = std::max(h->lstm[0].H, h->lstm[1].H);
  // ---- workspace carve
  size_t off = 0;
  auto carve = [&](size_t floats) { size_t o = off; off += (floats + 63) / 64 * 64; return o; };
  size_t o_tok = carve((size_t)2 * B * T), o_lab = carve(B);
  size_t o_st[2][7];
  for (int s = 0; s < 2; ++s) {
    int H = h->lstm[s].H;
    o_st[s][0] = carve((size_t)T * B * H);       // save_h
    o_st[s][1] = carve((size_t)T * B * H);       // save_c
    o_st[s][2] = carve((size_t)T * B * 5 * H);   // save_g
    o_st[s][3] = carve((size_t)B * E);           // u
    o_st[s][4] = carve((size_t)B * H);           // h0
    o_st[s][5] = carve((size_t)B * H);           // h1
    o_st[s][6] = carve((size_t)B * H);           // c
  }
  size_t o_du0 = carve((size_t)B * E), o_du1 = carve((size_t)B * E);
  size_t o_rl = carve(B), o_rp = carve(B), o_rn = carve(B);
  size_t o_dz = carve((size_t)T * B * 4 * Hmax);
  size_t o_dxh = carve((size_t)T * B * (We + Hmax));
  size_t o_xg = carve((size_t)T * B * We);
  size_t o_dh = carve((size_t)B * Hmax), o_dc = carve((size_t)B * Hmax);
  SSE_TRY(h->train_ws.ensure(off * 4));
  float* w = h->train_ws.as<float>();
  int32_t* d_src = reinterpret_cast<int32_t*>(w + o_tok);
  int32_t* d_tgt = d_src + (size_t)B * T;
  float* d_lab = w + o_lab;
  SSE_CUDA_OK(cudaMemcpyAsync(d_src, src, (size_t)B * T * 4, cudaMemcpyDefault, st));
  SSE_CUDA_OK(cudaMemcpyAsync(d_tgt, tgt, (size_t)B * T * 4, cudaMemcpyDefault, st));
  SSE_CUDA_OK(cudaMemcpyAsync(d_lab, labels, (size_t)B * 4, cudaMemcpyDefault, st));
  // ids outside [0, V) would gather past the embedding table and scatter gradients out of bounds (TF raises
  // InvalidArgument): replaced by PAD and counted; reported when the step hands scalars back / by sse_token_errors
  SSE_TRY(sanitize_tokens_inplace(d_src, (int64_t)2 * B * T, V, h->tok_bad, st, &h->launches));
  const float* emb = h->params[h->emb_param].dev;

  // ---- tensor-core path (bf16 operands): option train = 2, or auto with the tensor-core precision and supported shapes
  const bool want_tc = h->opt_train == 2 || (h->opt_train == 0 && c.precision == SSE_PRECISION_TC);
  if (want_tc && !train_tc_supported(h, B)) {
    if (h->opt_train == 2) { set_error("tensor-core train step needs We%%8==0, H%%8==0, B%%8==0 (We=%d B=%d)", We, B); return SSE_EINVAL; }
  } else if (want_tc) {
    const int64_t TB = (int64_t)T * B;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t need = 2 * (al((size_t)TB * 4 * Hmax * 4) + al((size_t)4 * Hmax * (We + Hmax) * 2) * 2) + 2 * al((size_t)B * E * 4) + 3 * al((size_t)B * 4);
    for (int s2 = 0; s2 < 2; ++s2) {
      const int H = h->lstm[s2].H;
      need += al((size_t)TB * 5 * H * 4) + al((size_t)TB * H * 4) + al((size_t)B * E * 4) + al((size_t)B * H * 4) + al((size_t)TB * H * 2) + al((size_t)TB * We * 2) + al((size_t)TB * 4);
    }
    const int64_t ldT = (TB + 7) / 8 * 8;
    const size_t bwd_bytes = al((size_t)TB * 4 * Hmax * 2) + al((size_t)4 * Hmax * ldT * 2) + al((size_t)We * ldT * 2) + al((size_t)Hmax * ldT * 2) + al((size_t)TB * We * 4) +
                             al((size_t)2 * B * Hmax * 4) + al((size_t)B * Hmax * 4) + 4096;
    need += 2 * bwd_bytes;
    SSE_TRY(h->train_tc_ws.ensure(need));
    uint8_t* tw8 = h->train_tc_ws.as<uint8_t>();
    size_t o2 = 0;
    auto carve8 = [&](size_t bytes) { size_t o = o2; o2 = al(o2 + bytes); return tw8 + o; };
    // The two towers are independent chains of 2T small dependent launches each (GEMM + gate kernel per step): they run side by
    // side on two streams (fork after the token check, join before the pair loss, fork / join again around the backward), each
    // with its own scratch.  SSE_TRAIN_STREAMS=1 puts both back on the caller's stream.
    static const bool two_streams = !(getenv("SSE_TRAIN_STREAMS") && atoi(getenv("SSE_TRAIN_STREAMS")) == 1);
    if (two_streams && !h->train_side) {
      SSE_CUDA_OK(cudaStreamCreateWithFlags(&h->train_side, cudaStreamNonBlocking));
      SSE_CUDA_OK(cudaStreamCreateWithFlags(&h->train_main, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        SSE_CUDA_OK(cudaStreamCreateWithFlags(&h->train_chain[i], cudaStreamNonBlocking));
        for (int j = 0; j < 4; ++j) SSE_CUDA_OK(cudaEventCreateWithFlags(&h->train_ev2[i][j], cudaEventDisableTiming));
      }
      for (int i = 0; i < 6; ++i) SSE_CUDA_OK(cudaEventCreateWithFlags(&h->train_ev[i], cudaEventDisableTiming));
    }
    cudaStream_t const st_user = st;
    if (two_streams) {                 // the step runs on the library's own pair of streams, forked from the caller's stream ...
      SSE_CUDA_OK(cudaEventRecord(h->train_ev[4], st_user));
      SSE_CUDA_OK(cudaStreamWaitEvent(h->train_main, h->train_ev[4], 0));
      st = h->train_main;
    }
    cudaStream_t st2[2] = {st, two_streams ? h->train_side : st};
    // SSE_TRAIN_CHAINS=2: additionally split each tower's recurrence into two row-half chains (measured 374 -> 377 steps/s at 1024
    // rows, 289 -> 297 at 1536: the two towers already saturate the GPU; off by default)
    static const bool env_chains = getenv("SSE_TRAIN_CHAINS") && atoi(getenv("SSE_TRAIN_CHAINS")) == 2;
    const bool row_chains = two_streams && env_chains;
    float* zx_s[2];
    uint16_t *kT16_s[2], *k16_s[2];
    for (int s2 = 0; s2 < 2; ++s2) {
      zx_s[s2] = reinterpret_cast<float*>(carve8((size_t)TB * 4 * Hmax * 4));
      kT16_s[s2] = reinterpret_cast<uint16_t*>(carve8((size_t)4 * Hmax * (We + Hmax) * 2));
      k16_s[s2] = reinterpret_cast<uint16_t*>(carve8((size_t)4 * Hmax * (We + Hmax) * 2));
    }
    float* du_tc[2] = {reinterpret_cast<float*>(carve8((size_t)B * E * 4)), reinterpret_cast<float*>(carve8((size_t)B * E * 4))};
    float* rl = reinterpret_cast<float*>(carve8((size_t)B * 4));
    float* rp = reinterpret_cast<float*>(carve8((size_t)B * 4));
    float* rn = reinterpret_cast<float*>(carve8((size_t)B * 4));
    // everything below addresses fixed buffers (workspace, arena, weights): with SSE_TRAIN_GRAPH=1 the ~530 launches of a step are
    // captured once per (batch size, buffer addresses) and replayed -- the step is bound by the host's launch rate otherwise
    auto issue = [&]() -> int {
    TcTowerState tst[2];
      if (two_streams) {
        SSE_CUDA_OK(cudaEventRecord(h->train_ev[0], st));                 // tokens, labels and the zeroed arena are in place
        SSE_CUDA_OK(cudaStreamWaitEvent(h->train_side, h->train_ev[0], 0));
      }
      for (int s2 = 0; s2 < 2; ++s2)
        SSE_TRY(train_tc_forward(h, s2, s2 == 0 ? d_src : d_tgt, B, tw8, &o2, &tst[s2], kT16_s[s2], zx_s[s2], st2[s2], row_chains ? h->train_chain[s2] : nullptr, h->train_ev2[s2]));
      if (two_streams) {
        SSE_CUDA_OK(cudaEventRecord(h->train_ev[1], h->train_side));
        SSE_CUDA_OK(cudaStreamWaitEvent(st, h->train_ev[1], 0));
      }
      pair_loss_kernel<<<cdiv(B, 8), 256, 0, st>>>(tst[0].u, tst[1].u, d_lab, B, E, 1.0f / (float)B_global, du_tc[0], du_tc[1], rl, rp, rn, nullptr);
      ++h->launches;
      reduce_rows_kernel<<<1, 256, 0, st>>>(rl, rp, rn, B, scalars);
      ++h->launches;
      if (two_streams) {
        SSE_CUDA_OK(cudaEventRecord(h->train_ev[2], st));
        SSE_CUDA_OK(cudaStreamWaitEvent(h->train_side, h->train_ev[2], 0));
      }
      const size_t o_bwd = o2;
      for (int s2 = 0; s2 < 2; ++s2)      // each tower's backward scratch: its own region behind the forward state
        SSE_TRY(train_tc_backward(h, s2, s2 == 0 ? d_src : d_tgt, B, tst[s2], du_tc[s2], tw8, o_bwd + (size_t)s2 * bwd_bytes, k16_s[s2], G, touched, scalars, st2[s2], row_chains ? h->train_chain[s2] : nullptr, h->train_ev2[s2]));
      if (two_streams) {
        SSE_CUDA_OK(cudaEventRecord(h->train_ev[3], h->train_side));
        SSE_CUDA_OK(cudaStreamWaitEvent(st, h->train_ev[3], 0));
      }
      return SSE_OK;
    };
    static const bool env_graph = !(getenv("SSE_TRAIN_GRAPH") && atoi(getenv("SSE_TRAIN_GRAPH")) == 0);      // default on; =0 disables
    const bool use_graph = env_graph && !h->train_graph_failed;
    const unsigned long long key[6] = {(unsigned long long)B, (unsigned long long)(uintptr_t)tw8, (unsigned long long)(uintptr_t)arena,
                                       (unsigned long long)(uintptr_t)d_src, (unsigned long long)(uintptr_t)d_lab, (unsigned long long)B_global};
    bool same_graph = h->train_graph != nullptr, seen = true;
    for (int i = 0; i < 6; ++i) { same_graph = same_graph && h->train_graph_key[i] == key[i]; seen = seen && h->train_seen_key[i] == key[i]; }
    if (use_graph && two_streams && same_graph) {
      SSE_CUDA_OK(cudaGraphLaunch(h->train_graph, st));
      h->launches += h->train_graph_launches;
    } else if (use_graph && two_streams && seen) {
      // second call with this key (the first ran eagerly: every lazy one-time set-up has happened): capture
      if (h->train_graph) { cudaGraphExecDestroy(h->train_graph); h->train_graph = nullptr; }
      const long long l0 = h->launches;
      SSE_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      const size_t o2_before = o2;
      const int rc = issue();
      cudaGraph_t g = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(st, &g);
      cudaError_t ie = cudaErrorUnknown;
      if (rc == SSE_OK && ce == cudaSuccess && g) ie = cudaGraphInstantiate(&h->train_graph, g, 0);
      if (g) cudaGraphDestroy(g);
      if (ie == cudaSuccess) {
        h->train_graph_launches = h->launches - l0;
        for (int i = 0; i < 6; ++i) h->train_graph_key[i] = key[i];
        SSE_CUDA_OK(cudaGraphLaunch(h->train_graph, st));
      } else {
        // capture is an optimisation only: whatever went wrong (a context that does not allow it, an invalidated capture),
        // this handle runs its steps eagerly from now on
        h->train_graph = nullptr;
        h->train_graph_failed = true;
        cudaGetLastError();
        h->launches = l0;
        o2 = o2_before;
        SSE_TRY(issue());
      }
    } else {
      for (int i = 0; i < 6; ++i) h->train_seen_key[i] = key[i];
      SSE_TRY(issue());
    }
    if (two_streams) {                 // ... and joined back into it
      SSE_CUDA_OK(cudaEventRecord(h->train_ev[5], st));
      SSE_CUDA_OK(cudaStreamWaitEvent(st_user, h->train_ev[5], 0));
      st = st_user;
    }
    SSE_CUDA_OK(cudaGetLastError());
    if (loss_host || acc_host) {
      float sc[4];
      SSE_CUDA_OK(cudaMemcpyAsync(sc, scalars, 16, cudaMemcpyDeviceToHost, st));
      SSE_CUDA_OK(cudaStreamSynchronize(st));
      if (loss_host) *loss_host = sc[1];
      if (acc_host) *acc_host = sc[2] + sc[3];
      int bad = 0;
      SSE_TRY(take_token_errors(h, st, &bad));
      if (bad) { set_error("train step: %d token id(s) outside [0, vocab_size=%d)", bad, V); return SSE_EINVAL; }
    }
    return SSE_OK;
  }

  // ---- forward with stash
  TowerStash ts[2];
  for (int s = 0; s < 2; ++s) {
    const LstmTower& tw = h->lstm[s];
    ts[s] = {w + o_st[s][0], w + o_st[s][1], w + o_st[s][2], w + o_st[s][3], w + o_st[s][4], w + o_st[s][5], w + o_st[s][6], nullptr};
    SSE_TRY(lstm_forward_simt(s == 0 ? d_src : d_tgt, B, T, 0, emb, We, tw, ts[s].h0, ts[s].h1, ts[s].c, nullptr, nullptr,
                              ts[s].sh, ts[s].sc, ts[s].sg, &ts[s].hlast, st, &h->launches));
    SSE_TRY(sgemm(false, false, B, E, tw.H, 1.f, ts[s].hlast, tw.H, tw.M, E, 0.f, ts[s].u, E, st, &h->launches));
  }
  float* du[2] = {w + o_du0, w + o_du1};
  pair_loss_kernel<<<cdiv(B, 8), 256, 0, st>>>(ts[0].u, ts[1].u, d_lab, B, E, 1.0f / (float)B_global, du[0], du[1], w + o_rl,
                                               w + o_rp, w + o_rn, nullptr);
  ++h->launches;
  reduce_rows_kernel<<<1, 256, 0, st>>>(w + o_rl, w + o_rp, w + o_rn, B, scalars);
  ++h->launches;

  // ---- backward
  float* dZ = w + o_dz;
  float* dXH = w + o_dxh;
  float* xg = w + o_xg;
  float* dh0 = w + o_dh;
  float* dc = w + o_dc;
  for (int s = 0; s < 2; ++s) {
    const LstmTower& tw = h->lstm[s];
    const int H = tw.H, ld = We + H;
    const int32_t* tok = s == 0 ? d_src : d_tgt;
    float* gM = arena + h->params[tw.mparam].grad_off;
    float* gK = arena + h->params[tw.kparam].grad_off;
    float* gb = arena + h->params[tw.bparam].grad_off;
    // dM += h_last^T du ; dh = du M^T
    SSE_TRY(sgemm(true, false, H, E, B, 1.f, ts[s].hlast, H, du[s], E, 1.f, gM, E, st, &h->launches));
    SSE_TRY(sgemm(false, true, B, H, E, 1.f, du[s], E, tw.M, E, 0.f, dh0, H, st, &h->launches));
    SSE_CUDA_OK(cudaMemsetAsync(dc, 0, (size_t)B * H * 4, st));
    const float* dh = dh0;
    int ldh = H;
    for (int t = T - 1; t >= 0; --t) {
      float* dz_t = dZ + (size_t)t * B * 4 * H;
      float* dxh_t = dXH + (size_t)t * B * ld;
      lstm_bwd_gates_kernel<<<cdiv(B * H, 256), 256, 0, st>>>(dh, ldh, dc, ts[s].sg + (size_t)t * B * 5 * H,
                                                              t > 0 ? ts[s].sc + (size_t)(t - 1) * B * H : nullptr, B, H, dz_t);
      ++h->launches;
      // d[x;h] = dz K^T
      SSE_TRY(sgemm(false, true, B, ld, 4 * H, 1.f, dz_t, 4 * H, tw.K, 4 * H, 0.f, dxh_t, ld, st, &h->launches));
      dh = dxh_t + We;
      ldh = ld;
    }
    // dK[:We] += X^T dZ over all (t,b);  dK[We:] += H_prev^T dZ (t >= 1);  db += colsum dZ
    gather_time_major_kernel<<<148 * 8, 256, 0, st>>>(tok, B, T, emb, We, xg);
    ++h->launches;
    SSE_TRY(sgemm(true, false, We, 4 * H, T * B, 1.f, xg, We, dZ, 4 * H, 1.f, gK, 4 * H, st, &h->launches));
    if (T > 1)
      SSE_TRY(sgemm(true, false, H, 4 * H, (T - 1) * B, 1.f, ts[s].sh, H, dZ + (size_t)B * 4 * H, 4 * H, 1.f,
                    gK + (size_t)We * 4 * H, 4 * H, st, &h->launches));
    colsum_kernel<<<cdiv(4 * H, 32), 256, 0, st>>>(dZ, (int64_t)T * B, 4 * H, gb);
    ++h->launches;
    embed_scatter_kernel<<<148 * 4, 256, 0, st>>>(tok, B, T, dXH, ld, We, G, touched, scalars);
    ++h->launches;
  }
  // NOTE: the dense part of the global norm is taken in sse_train_apply, AFTER any cross-rank
  // all-reduce of the arena (|sum g|^2 != sum |g|^2); the un-merged embedding slices are additive.
  SSE_CUDA_OK(cudaGetLastError());
  if (loss_host || acc_host) {
    float sc[4];
    SSE_CUDA_OK(cudaMemcpyAsync(sc, scalars, 16, cudaMemcpyDeviceToHost, st));
    SSE_CUDA_OK(cudaStreamSynchronize(st));
    if (loss_host) *loss_host = sc[1];
    if (acc_host) *acc_host = sc[2] + sc[3];
    int bad = 0;
    SSE_TRY(take_token_errors(h, st, &bad));
    if (bad) { set_error("train step: %d token id(s) outside [0, vocab_size=%d)", bad, V); return SSE_EINVAL; }
  }
  return SSE_OK;
}

int sse_grad_arena(sse_handle* h, float** dev_ptr_out, int64_t* n_floats_out) {
  if (!h) return SSE_EINVAL;
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  SSE_TRY(ensure_arena(h));
  if (dev_ptr_out) *dev_ptr_out = h->grad_arena;
  if (n_floats_out) *n_floats_out = h->arena_floats;
  return SSE_OK;
}

int sse_train_apply(sse_handle* h, float* loss_host, float* acc_host, float* gnorm_host, void* stream) {
  if (!h || !h->grad_arena) { set_error("sse_train_apply: no gradients computed"); return SSE_ESTATE; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int V = h->cfg.vocab_size, We = h->cfg.embedding_size;
  float* arena = h->grad_arena;
  float* G = arena + h->grad_floats;
  float* touched = G + (size_t)V * We;
  float* scalars = touched + V;
  const float lr = h->learning_rate, max_norm = 5.0f;   // self.max_gradient_norm, sse_model.py:117
  sumsq_kernel<<<148 * 2, 256, 0, st>>>(arena, h->grad_floats, scalars);
  ++h->launches;
  for (int i = 0; i < h->n_vars; ++i) {
    Param& p = h->params[i];
    if (p.grad_off < 0) continue;
    Param& a = h->params[h->n_vars + i];
    int blocks = (int)std::min<int64_t>(cdiv64(p.numel, 256), 148 * 8);
    adagrad_dense_kernel<<<blocks, 256, 0, st>>>(p.dev, a.dev, arena + p.grad_off, p.numel, scalars, lr, max_norm);
    ++h->launches;
  }
  {
    Param& p = h->params[h->emb_param];
    Param& a = h->params[h->n_vars + h->emb_param];
    adagrad_rows_kernel<<<148 * 8, 256, 0, st>>>(p.dev, a.dev, G, touched, V, We, scalars, lr, max_norm);
    ++h->launches;
  }
  SSE_CUDA_OK(cudaGetLastError());
  h->global_step += 1;
  invalidate_derived(h);
  if (loss_host || acc_host || gnorm_host) {
    float sc[4];
    SSE_CUDA_OK(cudaMemcpyAsync(sc, scalars, 16, cudaMemcpyDeviceToHost, st));
    SSE_CUDA_OK(cudaStreamSynchronize(st));
    if (loss_host) *loss_host = sc[1];
    if (acc_host) *acc_host = sc[2] + sc[3];
    if (gnorm_host) *gnorm_host = sqrtf(sc[0]);
    int bad = 0;                               // the batch of this step was checked on the device (sse_train_grads)
    SSE_TRY(take_token_errors(h, st, &bad));
    if (bad) { set_error("train step: %d token id(s) outside [0, vocab_size=%d)", bad, V); return SSE_EINVAL; }
  }
  return SSE_OK;
}

int sse_train_step(sse_handle* h, const int32_t* src, const int32_t* tgt, const float* labels, int B, float* loss_host,
                   float* acc_host, float* gnorm_host, void* stream) {
  SSE_TRY(sse_train_grads(h, src, tgt, labels, B, B, nullptr, nullptr, stream));
  return sse_train_apply(h, loss_host, acc_host, gnorm_host, stream);
}

}  // extern "C"
