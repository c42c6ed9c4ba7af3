// Microbenchmark (not product code): MUFU throughput per SM sub-partition for ex2 / rcp / tanh on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mufu_mb scripts/mufu_microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int OP>
__global__ void k(int iters, float seed, long long* out, float* sink) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = seed + 0.01f * i + 0.001f * threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = op<OP>(v[i]);
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
template <int OP>
void run(const char* name, long long* d, float* sink) {
  for (int warps : {4, 8, 16}) {
    k<OP><<<148, warps * 32>>>(1000, 0.3f, d, sink);
    k<OP><<<148, warps * 32>>>(1000, 0.3f, d, sink);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    // per SMSP: warps/4 warps x 8000 MUFU warp-instructions
    printf("%s warps/SM %2d: %.2f cycles per MUFU warp-instruction per SMSP\n", name, warps, (double)h / (8000.0 * warps / 4));
  }
}
int main() {
  long long* d; float* sink; cudaMalloc(&d, 64); cudaMalloc(&sink, 148 * 512 * 4);
  run<0>("ex2 ", d, sink); run<1>("rcp ", d, sink); run<2>("tanh", d, sink);
  return 0;
}
