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
}  // namespace

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
