// Microbenchmark (not product code): issue cost of tcgen05.mma / tcgen05.commit from one thread.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tc_microbench scripts/tc_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = clock64();
  while (!done) {
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\nselp.u32 %0, 1, 0, P1;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    if (!done && clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %1;\n@%%px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// MODE 0 = SS, 1 = TS.  reps x (8 MMAs unrolled with constant operands) then ONE commit + wait.
// All operands are warp-uniform / loop-invariant so the issue loop is nothing but UTCHMMA.
template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128, 1) bench(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x >= 64 && threadIdx.x < 96) {
    const uint64_t a = make_desc(smem_u32(smem)), b = make_desc(smem_u32(smem + 32768));
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t d0 = tmem + 256, d1 = tmem + 256 + (NACC > 1 ? 128 : 0);
    long long t0 = clock64();
    if (elect_one_sync()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t d = (i & 1) ? d1 : d0;
          if (MODE == 0) mma_ss(d, a + 2 * (i & 3), b + 2 * (i & 3), idesc, 1);
          else mma_ts(d, tmem + 8 * (i & 3), b + 2 * (i & 3), idesc, 1);
        }
      }
      tc_commit(smem_u32(&bar));
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 64) { out[0] = t2 - t0; out[1] = t1 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)); }
}

template <int MODE, int N, int NACC>
void run(long long* d) {
  cudaFuncSetAttribute(bench<MODE, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 128;
  for (int w = 0; w < 2; ++w) bench<MODE, N, NACC><<<148, 128, 100 * 1024>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%s N=%3d acc=%d : %7.1f cyc/MMA to completion, %7.1f cyc/MMA issue (floor %d) %s\n", MODE ? "TS" : "SS", N, NACC,
         (double)h[0] / (reps * 8), (double)h[1] / (reps * 8), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  run<0, 64, 1>(d); run<0, 128, 1>(d); run<0, 256, 1>(d); run<0, 64, 2>(d); run<0, 128, 2>(d);
  run<1, 64, 1>(d); run<1, 128, 1>(d); run<1, 256, 1>(d); run<1, 64, 2>(d); run<1, 128, 2>(d);
  return 0;
}
