// n-gram CNN tower (fp32 SIMT): conv2d VALID over [T,We] with filter [k,We,1,F] -> +b ->
// ReLU -> max over the T-k+1 positions -> concat over k  (reference sse_model.py:185-207).
//
// A window (b,p) of width k is the contiguous slice X[b].flat[p*We : (p+k)*We] of the
// gathered embeddings, so the convolution is ONE GEMM per filter width over the
// flattened [B*T, We] activations with leading dimension We (overlapping rows); rows
// whose window crosses a sequence boundary are computed and ignored by the pooling.
#include "sse_common.cuh"
#include <math_constants.h>

namespace sse {
namespace {

__global__ void gather_rows_kernel(const int32_t* __restrict__ tokens, int64_t n_tok, const float* __restrict__ emb,
                                   int We, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n_tok * We;
  for (; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / We;
    int e = (int)(i - r * We);
    out[i] = __ldg(emb + (size_t)tokens[r] * We + e);
  }
}

// conv: [B*T (rows b*T+p), F] raw GEMM output; pool[b, off+f] = relu(max_p conv + bias[f])
__global__ void bias_relu_maxpool_kernel(const float* __restrict__ conv, int B, int T, int P, int F,
                                         const float* __restrict__ bias, float* __restrict__ pool, int sumF, int off,
                                         int32_t* __restrict__ argmax) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  int b = blockIdx.y;
  if (f >= F || b >= B) return;
  const float* c = conv + ((size_t)b * T) * F + f;
  float m = -CUDART_INF_F;
  int am = 0;
  for (int p = 0; p < P; ++p) {
    float v = c[(size_t)p * F];
    if (v > m) { m = v; am = p; }
  }
  float r = m + __ldg(bias + f);
  // relu(max) == max(relu); np.argmax of the ReLU output picks the first max, which for an
  // all-non-positive column is position 0 (all zeros after ReLU).
  if (!(r > 0.f)) { r = 0.f; am = 0; }
  pool[(size_t)b * sumF + off + f] = r;
  if (argmax) argmax[(size_t)b * sumF + off + f] = am;
}

}  // namespace

// scratch layout is owned by the caller (sse_api.cu): xg [B*T*We] then conv [B*T*maxF]
int cnn_forward_ws(const int32_t* tokens, int B, int T, const float* emb, int We, const CnnTower& tw, float* xg,
                   float* conv, float* pool, int32_t* argmax, cudaStream_t st, int64_t* launches) {
  int64_t n_tok = (int64_t)B * T;
  int blocks = (int)std::min<int64_t>(cdiv64(n_tok * We, 256), 148 * 16);
  gather_rows_kernel<<<blocks, 256, 0, st>>>(tokens, n_tok, emb, We, xg);
  if (launches) ++*launches;
  int off = 0;
  for (int i = 0; i < tw.nf; ++i) {
    int k = tw.ksize[i], F = tw.nfilt[i];
    int P = T - k + 1;
    if (P <= 0) { set_error("cnn: filter width %d > max_seq_length %d", k, T); return SSE_EINVAL; }
    int64_t rows = n_tok - (k - 1);
    SSE_TRY(sgemm(false, false, (int)rows, F, k * We, 1.f, xg, We, tw.W[i], F, 0.f, conv, F, st, launches));
    dim3 grid(cdiv(F, 128), B);
    bias_relu_maxpool_kernel<<<grid, 128, 0, st>>>(conv, B, T, P, F, tw.b[i], pool, tw.sumF, off, argmax);
    if (launches) ++*launches;
    off += F;
  }
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

// ---- tensor-core path (gemm_tc.cu): fp16 operands / fp32 accumulate; the [B*T, F] convolution output never exists
void cnn_tc_release(CnnTc& ct) {
  for (int i = 0; i < SSE_MAX_CNN_FILTERS; ++i) { if (ct.wt[i]) cudaFree(ct.wt[i]); ct.wt[i] = nullptr; }
  if (ct.mt) cudaFree(ct.mt);
  ct.mt = nullptr; ct.valid = false;
}

bool cnn_tc_supported(int We, int T, const CnnTower& tw) {
  if (We % 8 != 0 || tw.sumF % 8 != 0) return false;
  for (int i = 0; i < tw.nf; ++i)
    if (T - tw.ksize[i] + 1 > 128 || T - tw.ksize[i] + 1 <= 0) return false;
  return true;
}

// 16-bit K-major copies of the filters ([F, k*We]) and of the projection ([E, sumF]); rebuilt when the weights change
int cnn_tc_prepare(CnnTc& ct, const CnnTower& tw, int We, int E, cudaStream_t st, int64_t* launches) {
  for (int i = 0; i < tw.nf; ++i) {
    const int K = tw.ksize[i] * We, F = tw.nfilt[i];
    if (!ct.wt[i]) SSE_CUDA_OK(cudaMalloc(&ct.wt[i], (size_t)F * K * 2));
    SSE_TRY(transpose_to_16(tw.W[i], K, F, F, ct.wt[i], K, 0, st, launches));       // W [k*We, F] -> [F, k*We]
  }
  if (!ct.mt) SSE_CUDA_OK(cudaMalloc(&ct.mt, (size_t)E * tw.sumF * 2));
  SSE_TRY(transpose_to_16(tw.M, tw.sumF, E, E, ct.mt, tw.sumF, 0, st, launches));   // M [sumF, E] -> [E, sumF]
  ct.valid = true;
  return SSE_OK;
}

size_t cnn_tc_ws_bytes(int nb, int T, int We, const CnnTower& tw) {
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  return al((size_t)nb * T * We * 2) + al((size_t)nb * tw.sumF * 4) + al((size_t)nb * tw.sumF * 2);
}

// tokens [nb,T] -> proj [nb,E] (un-normalised encodings): gather (fp16) -> per filter width fused conv+bias+ReLU+max-pool
// GEMM -> fp16 copy of the pooled features -> projection GEMM
int cnn_forward_tc(const int32_t* tokens, int nb, int T, const float* emb, int We, int E, const CnnTower& tw, const CnnTc& ct, void* ws,
                   float* proj, cudaStream_t st, int64_t* launches) {
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  uint8_t* w = static_cast<uint8_t*>(ws);
  uint16_t* x16 = reinterpret_cast<uint16_t*>(w);
  float* pool = reinterpret_cast<float*>(w + al((size_t)nb * T * We * 2));
  uint16_t* pool16 = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(pool) + al((size_t)nb * tw.sumF * 4));
  SSE_TRY(gather_rows_16(tokens, (int64_t)nb * T, emb, We, We, x16, 0, st, launches));
  SSE_CUDA_OK(cudaMemsetAsync(pool, 0, (size_t)nb * tw.sumF * 4, st));
  int off = 0;
  for (int i = 0; i < tw.nf; ++i) {
    SSE_TRY(cnn_conv_pool_tc(x16, nb, T, We, tw.ksize[i], ct.wt[i], tw.nfilt[i], tw.b[i], pool, tw.sumF, off, 0, st, launches));
    off += tw.nfilt[i];
  }
  SSE_TRY(convert_to_16(pool, nb, tw.sumF, tw.sumF, pool16, tw.sumF, 0, st, launches));
  return gemm_tc(pool16, tw.sumF, ct.mt, tw.sumF, nb, E, tw.sumF, 1.f, 0.f, proj, E, 0, 1, nullptr, 0, st, launches);
}

}  // namespace sse
