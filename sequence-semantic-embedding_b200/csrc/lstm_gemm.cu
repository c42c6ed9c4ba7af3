// LSTM tower as one tensor-core GEMM per time step (any We / H that are multiples of 8; the path for cells wider than the
// resident-weight kernels hold: H = 512 of BASELINE config 5, reference recipe makefile:17).
// Restates BasicLSTMCell(forget_bias=1) + static_rnn (reference sse_model.py:222-224, 240-250, 262-275):
//     z_t = [x_t, h_{t-1}] K + b ;  c = c sigmoid(f + 1) + sigmoid(i) tanh(j) ;  h = tanh(c) sigmoid(o)
//   * operand buffer A [T, B, We+H] fp16, time-major: the gather writes x_t into columns [0, We) of block t, the gate
//     kernel of step t writes h_t into columns [We, We+H) of block t+1 -- so block t IS the [B, We+H] A operand of step t
//     and the whole step is ONE gemm_tc call (M = B, N = 4H, K = We+H) against K^T [4H, We+H] fp16;
//   * c stays fp32 in HBM ([B,H], L2-resident at query batch sizes), gate math in fp32 (expf / tanhf);
//   * numerics: fp16 operands, fp32 accumulate and state -- the same contract as the other tensor-core towers (<= 1e-3
//     on the normalised encodings against the fp32 oracle).
// Per step the GEMM streams the weights once per 128-row m-tile from L2; for the large batches of an index build it runs
// at GEMM speed, for a 600-row query batch the 2 T launches (~3 us each) bound it -- still ~50x the fp32 SIMT path at H=512.
#include "sse_common.cuh"

namespace sse {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void tokens_tm_kernel(const int32_t* __restrict__ tok, int B, int T, int32_t* __restrict__ out) {
  const int64_t total = (int64_t)B * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / B), b = (int)(i - (int64_t)t * B);
    out[i] = tok[(size_t)b * T + t];
  }
}

// z [B,4H] (TF gate order i,j,f,o, no bias) -> c (in place), h as fp16 into the next step's operand block, fp32 h at the end
__global__ void lstm_gemm_gates_kernel(const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ c, int first, int B, int H,
                                       __half* __restrict__ h16_next /*row stride ld, may be null*/, int ld, float* __restrict__ h32 /*[B,ldh] or null*/, int ldh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, u = i - b * H;
  const float* zr = z + (size_t)b * 4 * H + u;
  const float si = sigmoidf_(zr[0] + __ldg(bias + u));
  const float tj = tanhf(zr[H] + __ldg(bias + H + u));
  const float sf = sigmoidf_(zr[2 * H] + __ldg(bias + 2 * H + u) + 1.0f);      // forget_bias = 1.0 added at run time
  const float so = sigmoidf_(zr[3 * H] + __ldg(bias + 3 * H + u));
  const float cn = (first ? 0.f : c[i]) * sf + si * tj;
  c[i] = cn;
  const float hn = tanhf(cn) * so;
  if (h16_next) h16_next[(size_t)b * ld + u] = __float2half_rn(hn);
  if (h32) h32[(size_t)b * ldh + u] = hn;
}

}  // namespace

bool lstm_gemm_supported(int We, int H) { return We % 8 == 0 && H % 8 == 0 && We >= 8 && H >= 8; }

void lstm_gemm_release(GemmTower& gt) {
  if (gt.kT16) cudaFree(gt.kT16);
  gt.kT16 = nullptr; gt.valid = false;
}

// K [We+H, 4H] fp32 -> K^T [4H, We+H] fp16 (B operand of every step), once per weight version
int lstm_gemm_prepare(GemmTower& gt, const float* K, int We, int H, cudaStream_t st, int64_t* launches) {
  if (!gt.kT16) SSE_CUDA_OK(cudaMalloc(&gt.kT16, (size_t)4 * H * (We + H) * 2));
  SSE_TRY(transpose_to_16(K, We + H, 4 * H, 4 * H, reinterpret_cast<uint16_t*>(gt.kT16), We + H, 0, st, launches));
  gt.valid = true;
  return SSE_OK;
}

size_t lstm_gemm_ws_bytes(int B, int T, int We, int H) {
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  return al((size_t)T * B * (We + H) * 2) + al((size_t)B * 4 * H * 4) + al((size_t)B * H * 4) + al((size_t)T * B * 4) + al((size_t)T * B * We * 2);
}

// tokens [B,T] (sanitised) -> h_out [B, ldh] fp32 (last step); emb fp32 [V, We], bias fp32 [4H] in TF order
int lstm_forward_gemm(const int32_t* tokens, int B, int T, const float* emb, int We, int H, const GemmTower& gt, const float* bias, void* ws,
                      float* h_out, int ldh, cudaStream_t st, int64_t* launches) {
  if (B <= 0) return SSE_OK;
  if (!gt.valid) { set_error("lstm_forward_gemm: weights not prepared"); return SSE_ESTATE; }
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const int ld = We + H;
  uint8_t* w = static_cast<uint8_t*>(ws);
  __half* A = reinterpret_cast<__half*>(w); w += al((size_t)T * B * ld * 2);
  float* z = reinterpret_cast<float*>(w); w += al((size_t)B * 4 * H * 4);
  float* c = reinterpret_cast<float*>(w); w += al((size_t)B * H * 4);
  int32_t* tok_tm = reinterpret_cast<int32_t*>(w); w += al((size_t)T * B * 4);
  uint16_t* x16 = reinterpret_cast<uint16_t*>(w);
  const int64_t TB = (int64_t)T * B;
  tokens_tm_kernel<<<(int)std::min<int64_t>(cdiv64(TB, 256), 148 * 4), 256, 0, st>>>(tokens, B, T, tok_tm);
  if (launches) ++*launches;
  // x_t into columns [0, We) of every block: gather to a dense [T*B, We] buffer, then a strided 2-D copy into A
  SSE_TRY(gather_rows_16(tok_tm, TB, emb, We, We, x16, 0, st, launches));
  SSE_CUDA_OK(cudaMemcpy2DAsync(A, (size_t)ld * 2, x16, (size_t)We * 2, (size_t)We * 2, (size_t)TB, cudaMemcpyDeviceToDevice, st));
  for (int t = 0; t < T; ++t) {
    const uint16_t* At = reinterpret_cast<const uint16_t*>(A + (size_t)t * B * ld);
    // step 0: h_{-1} = 0, only the x part contributes (K = We); later steps read the full [x_t, h_{t-1}] block
    SSE_TRY(gemm_tc(At, ld, reinterpret_cast<const uint16_t*>(gt.kT16), ld, B, 4 * H, t == 0 ? We : ld, 1.f, 0.f, z, 4 * H, 0, 1, nullptr, 0, st, launches));
    const bool last = t == T - 1;
    lstm_gemm_gates_kernel<<<cdiv(B * H, 256), 256, 0, st>>>(z, bias, c, t == 0 ? 1 : 0, B, H, last ? nullptr : A + (size_t)(t + 1) * B * ld + We, ld,
                                                             last ? h_out : nullptr, ldh);
    if (launches) ++*launches;
  }
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

}  // namespace sse
