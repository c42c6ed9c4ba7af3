// Microbenchmark (not product code): TMA streaming throughput per SM into a shared-memory ring.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmamb scripts/tma_microbench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = clock64();
  while (!done) {
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %1;\n@%%px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}

// mode 0: 2D tensor-map loads, box = 64 cols x box_rows of a [total_rows, row_elems] fp16 matrix (k-block kb = column offset)
// mode 1: 1D bulk copies of stage_bytes contiguous bytes
__global__ void __launch_bounds__(64, 1) stream(const __grid_constant__ CUtensorMap tm, const uint8_t* base, int mode, int ns, int box_rows,
                                                int total_rows, int kblocks, int n_loads, size_t per_cta_offset_rows, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[32];
  const uint32_t stage_bytes = (uint32_t)box_rows * 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ns; ++s) mbar_init(smem_u32(&full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    long long t0 = clock64();
    const int tiles = total_rows / box_rows;
    const int tile0 = (int)((per_cta_offset_rows * blockIdx.x / box_rows) % tiles);
    // prologue: fill the ring
    for (int i = 0; i < ns && i < n_loads; ++i) {
      if (elect_one_sync()) {
        mbar_expect_tx(smem_u32(&full[i]), stage_bytes);
        int tile = (tile0 + i / kblocks) % tiles, kb = i % kblocks;
        if (mode == 0) tma_load_2d(smem_u32(smem + (size_t)i * stage_bytes), &tm, smem_u32(&full[i]), kb * 64, tile * box_rows);
        else bulk_load_1d(smem_u32(smem + (size_t)i * stage_bytes), base + ((size_t)tile * kblocks + kb) * stage_bytes, stage_bytes, smem_u32(&full[i]));
      }
      __syncwarp();
    }
    for (int i = 0; i < n_loads; ++i) {
      const int s = i % ns;
      mbar_wait(smem_u32(&full[s]), (i / ns) & 1);
      const int nx = i + ns;
      if (nx < n_loads && elect_one_sync()) {
        mbar_expect_tx(smem_u32(&full[s]), stage_bytes);
        int tile = (tile0 + nx / kblocks) % tiles, kb = nx % kblocks;
        if (mode == 0) tma_load_2d(smem_u32(smem + (size_t)s * stage_bytes), &tm, smem_u32(&full[s]), kb * 64, tile * box_rows);
        else bulk_load_1d(smem_u32(smem + (size_t)s * stage_bytes), base + ((size_t)tile * kblocks + kb) * stage_bytes, stage_bytes, smem_u32(&full[s]));
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fnp;
  long long* d_out; cudaMalloc(&d_out, 148 * 8);
  cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  struct Case { const char* name; size_t rows; int row_elems; };
  Case cases[2] = {{"L2 1MB", 2048, 256}, {"HBM 512MB", 1000000, 256}};
  for (auto& cs : cases) {
    uint8_t* buf; size_t bytes = cs.rows * cs.row_elems * 2; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    for (int box_rows : {64, 128}) {
      CUtensorMap tm;
      cuuint64_t gdim[2] = {(cuuint64_t)cs.row_elems, (cuuint64_t)cs.rows}; cuuint64_t gstr[1] = {(cuuint64_t)cs.row_elems * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
      enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      for (int mode = 0; mode < 2; ++mode)
        for (int grid : {1, 148})
          for (int ns : {2, 6, 13}) {
            size_t stage = (size_t)box_rows * 128;
            if (ns * stage > 220 * 1024) continue;
            int kblocks = cs.row_elems / 64;
            int n_loads = 2048;
            size_t per_cta_rows = cs.rows / 148;
            for (int w = 0; w < 2; ++w)
              stream<<<grid, 64, ns * stage + 1024>>>(tm, buf, mode, ns, box_rows, (int)cs.rows / box_rows * box_rows, kblocks, n_loads, per_cta_rows, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<long long> h(grid); cudaMemcpy(h.data(), d_out, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
            double bpc = (double)n_loads * stage / mx;
            printf("%-9s %s box=%3d grid=%3d ring=%2dx%5zu: %6.1f B/cyc/SM %5.2f TB/s %s\n", cs.name, mode ? "bulk1D" : "tma2D ",
                   box_rows, grid, ns, stage, bpc, bpc * grid * 1.9e9 / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
            fflush(stdout);
            if (e != cudaSuccess) return 1;
          }
    }
    cudaFree(buf);
  }
  return 0;
}
