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

// mode: 0 = SS, 1 = TS.  n_mma MMAs per commit, reps commits.  D alternates over n_acc accumulators.
__global__ void __launch_bounds__(128, 1) bench(int mode, int N, int n_mma, int reps, int n_acc, int commit_every_rep, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem = tslot;
  if (threadIdx.x >= 64 && threadIdx.x < 96) {   // warp 2, converged; one elected lane issues (warp-uniform operands)
    uint64_t a = make_desc(smem_u32(smem)), b = make_desc(smem_u32(smem + 32768));
    uint32_t idesc = make_idesc(128, N);
    uint32_t ncommit = 0;
    long long t_issue = 0, t_commit = 0, t_wait = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      long long c0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        uint32_t d = tmem + 256 + (uint32_t)((i % n_acc) * 64 % 256);
        if (elect_one_sync()) {
          if (mode == 0) mma_ss(d, a + 2 * (i & 3), b + 2 * (i & 3), idesc, 1);
          else mma_ts(d, tmem + 8 * (i & 3), b + 2 * (i & 3), idesc, 1);
        }
        __syncwarp();
      }
      long long c1 = clock64();
      t_issue += c1 - c0;
      if (commit_every_rep || r == reps - 1) {
        if (elect_one_sync()) tc_commit(smem_u32(&bar));
        __syncwarp(); ++ncommit;
        long long c2 = clock64();
        t_commit += c2 - c1;
        if (commit_every_rep == 2 || r == reps - 1) {   // 2 = also wait for completion every rep
          mbar_wait(smem_u32(&bar), (ncommit - 1) & 1);   // phase k completes at the k-th commit (count 1)
          t_wait += clock64() - c2;
        }
      }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 64) { out[0] = t1 - t0; out[1] = t_issue; out[2] = t_commit; out[3] = t_wait; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int Ns[3] = {64, 128, 256};
  for (int grid : {148})
    for (int mode = 0; mode < 2; ++mode)
      for (int ni = 0; ni < 3; ++ni)
        for (int n_acc : {1})
          for (int cfg = 0; cfg < 4; ++cfg) {
            int N = Ns[ni];
            if (N == 256 && n_acc == 2) continue;
            int n_mma = cfg == 0 ? 256 : (cfg == 1 ? 32 : (cfg == 2 ? 8 : 8));
            int reps = cfg == 0 ? 4 : (cfg == 1 ? 32 : 128);
            int ce = cfg == 0 ? 0 : (cfg == 3 ? 2 : 1);
            for (int w = 0; w < 2; ++w) bench<<<grid, 128, 100 * 1024>>>(mode, N, n_mma, reps, n_acc, ce, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
            long long total_mma = (long long)n_mma * reps;
            printf("grid %3d %s N=%3d n_acc=%d mma/commit=%3d commits=%3d wait_each=%d : total %8lld cyc  per-mma %6.1f (floor %d)  issue/mma %6.1f  commit %6.1f  wait %7.1f  %s\n",
                   grid, mode ? "TS" : "SS", N, n_acc, n_mma, ce ? reps : 1, ce == 2, h[0], (double)h[0] / total_mma, N / 2,
                   (double)h[1] / total_mma, (double)h[2] / (ce ? reps : 1), (double)h[3] / (ce == 2 ? reps : 1), e == cudaSuccess ? "" : cudaGetErrorString(e));
            fflush(stdout);
            if (e != cudaSuccess) return 1;
          }
  return 0;
}
