#include "sse_common.cuh"
#include <stdarg.h>

namespace sse {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int Scratch::ensure(size_t bytes) {
  if (bytes <= cap) return SSE_OK;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); return SSE_ENOMEM; }
  cap = want;
  return SSE_OK;
}
void Scratch::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

namespace {
__global__ void fill_kernel(float* p, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void f32_to_f16_kernel(const float* __restrict__ s, __half* __restrict__ d, int64_t n4) {
  // n4 = number of float4 groups
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(s)[i];
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(d)[i] = o;
  }
}
__global__ void f32_to_f16_tail(const float* __restrict__ s, __half* __restrict__ d, int64_t from, int64_t n) {
  int64_t i = from + threadIdx.x;
  if (i < n) d[i] = __float2half_rn(s[i]);
}
// dst [R, Cp] = src [R, C] with zero columns appended
__global__ void pad_cols_kernel(const float* __restrict__ s, int64_t R, int C, float* __restrict__ d, int Cp) {
  const int64_t total = R * Cp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Cp;
    const int c = (int)(i - r * Cp);
    d[i] = c < C ? s[r * C + c] : 0.f;
  }
}
// LSTM kernel [We+H, 4H] (rows x then h; column blocks i,j,f,o) -> [Wp+Hp, 4Hp] with zero rows / columns for the padding
__global__ void pad_lstm_kernel(const float* __restrict__ K, const float* __restrict__ b, int We, int H, int Wp, int Hp,
                                float* __restrict__ Kp, float* __restrict__ bp) {
  const int64_t total = (int64_t)(Wp + Hp) * 4 * Hp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (4 * Hp)), cc = (int)(i - (int64_t)r * 4 * Hp);
    const int g = cc / Hp, u = cc - g * Hp;
    int rs = -1;
    if (r < We) rs = r;
    else if (r >= Wp && r - Wp < H) rs = We + (r - Wp);
    Kp[i] = (rs >= 0 && u < H) ? K[(size_t)rs * 4 * H + g * H + u] : 0.f;
    if (r == 0) bp[cc] = u < H ? b[g * H + u] : 0.f;
  }
}
}  // namespace

int pad_cols(const float* src, int64_t R, int C, float* dst, int Cp, cudaStream_t st, int64_t* launches) {
  pad_cols_kernel<<<(int)std::min<int64_t>(cdiv64(R * Cp, 256), 148 * 8), 256, 0, st>>>(src, R, C, dst, Cp);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}
int pad_lstm_weights(const float* K, const float* b, int We, int H, int Wp, int Hp, float* Kp, float* bp, cudaStream_t st, int64_t* launches) {
  pad_lstm_kernel<<<148 * 2, 256, 0, st>>>(K, b, We, H, Wp, Hp, Kp, bp);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

int fill_f32(float* p, int64_t n, float v, cudaStream_t st, int64_t* launches) {
  if (n <= 0) return SSE_OK;
  int blocks = (int)std::min<int64_t>(cdiv64(n, 256), 148 * 8);
  fill_kernel<<<blocks, 256, 0, st>>>(p, n, v);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

int f32_to_f16(const float* src, __half* dst, int64_t n, cudaStream_t st, int64_t* launches) {
  if (n <= 0) return SSE_OK;
  int64_t n4 = n / 4;
  if (n4 > 0) {
    int blocks = (int)std::min<int64_t>(cdiv64(n4, 256), 148 * 16);
    f32_to_f16_kernel<<<blocks, 256, 0, st>>>(src, dst, n4);
    if (launches) ++*launches;
  }
  if (n4 * 4 < n) {
    f32_to_f16_tail<<<1, 4, 0, st>>>(src, dst, n4 * 4, n);
    if (launches) ++*launches;
  }
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

}  // namespace sse
