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
  }
  return SSE_OK;
}

int sse_train_step(sse_handle* h, const int32_t* src, const int32_t* tgt, const float* labels, int B, float* loss_host,
                   float* acc_host, float* gnorm_host, void* stream) {
  SSE_TRY(sse_train_grads(h, src, tgt, labels, B, B, nullptr, nullptr, stream));
  return sse_train_apply(h, loss_host, acc_host, gnorm_host, stream);
}

}  // extern "C"
