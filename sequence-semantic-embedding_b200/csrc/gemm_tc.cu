// General tensor-core GEMM for sm_100a (tcgen05 + TMEM + TMA), used by the training step and the CNN tower:
//     D[M,N] (fp32)  =  alpha * A[M,K] * B[N,K]^T  (+ beta * D)          A, B: 16-bit (fp16 or bf16), K contiguous
// i.e. both operands K-major, exactly what the tensor core reads from SWIZZLE_128B shared-memory tiles.  Callers keep
// 16-bit K-major copies of their operands (conversion / transposition kernels below); outputs are fp32.
//   * one CTA per 128 x BN output tile (and per K split): warp 0 = TMA producer (6-stage ring of [128x64] + [BNx64]
//     tiles), warp 1 = MMA issuer (tcgen05.mma kind::f16, M=128, N=BN, accumulator in TMEM; each stage is handed back
//     to the producer by a tcgen05.commit on its `empty` barrier), warps 2-5 = epilogue (thread == output row:
//     tcgen05.ld 32 columns at a time).
//   * epilogues: STORE (alpha/beta, optional 16-bit copy for the next GEMM), ATOMIC (split-K partial sums into a
//     pre-initialised D: the gradient reductions over T*B rows), POOL (CNN: + bias, ReLU, max over the positions of a
//     sequence -- rows of a tile are whole sequences, warp REDUX max on the non-negative float bits -- so the
//     [B*T, F] convolution output never exists; reference sse_model.py:190-202).
// Reference ops replaced: tf.matmul / conv2d inside the TF-1 graph of sse_model.py:185-211 (CNN) and the GEMMs of
// tf.gradients over static_rnn (sse_model.py:355-364).
#include "sse_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <math_constants.h>
#include <stdlib.h>
#include <algorithm>

namespace sse {

namespace {

constexpr int KBLK = 64;               // 16-bit elements per 128-byte swizzle row
constexpr int G_THREADS = 192;
constexpr int G_STAGES_MAX = 6;

struct GemmParams {
  int M, N, K;
  int kb_per_split;                    // k-blocks per blockIdx.z
  float alpha, beta;
  float* D; int64_t ldd;
  uint16_t* D16; int64_t ldd16; int d16_fmt;      // optional 16-bit copy of the result (STORE)
  int fmt;                             // operands: 0 fp16, 1 bf16
  int n_stages;
  // POOL epilogue (CNN): a tile holds 128 / rows_per_seq sequences of rows_per_seq rows each
  int rows_per_seq, valid_pos, n_seq;  // valid_pos = T - k + 1 positions per sequence
  const float* bias; float* pool; int pool_ld, pool_off;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(32);
    if ((spins & 0xfff) == 0xfff) {          // watchdog: a protocol bug becomes a trap, not a hung GPU
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %1;\n@%%px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row x 128-byte atoms, SBO = 1024 B, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = f32, A/B = fp16 (fmt 0) or bf16 (fmt 1), both K-major, M x N
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_ss(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 ad, bd;\nsetp.ne.b32 p, %6, 0;\nmov.b64 ad, {%1, %2};\nmov.b64 bd, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n}\n" ::"r"(d),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
      : "memory");
}
#define G_TMEM_LD_32(taddr, v)                                                                                     \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
      "%29,%30,%31}, [%32];"                                                                                       \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),     \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),    \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                  \
      : "r"(taddr))

__device__ __forceinline__ uint16_t to16(float v, int fmt) {
  if (fmt == 1) { __nv_bfloat16 b = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&b); }
  __half hh = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&hh);
}

enum { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_POOL = 2 };

template <int BN, int EPI>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ GemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t A_BYTES = 128 * KBLK * 2, B_BYTES = BN * KBLK * 2, STAGE = A_BYTES + B_BYTES;
  const int NS = P.n_stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NS * STAGE);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NS), bar_acc = smem_u32(bars + 2 * NS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 1);
  const int m0 = blockIdx.y * 128, n0 = blockIdx.x * BN;
  const int nkb_total = (P.K + KBLK - 1) / KBLK;
  const int kb0 = blockIdx.z * P.kb_per_split;
  const int kb1 = min(nkb_total, kb0 + P.kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN < 32 ? 32 : BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    for (int kb = kb0; kb < kb1; ++kb) {
      const int i = kb - kb0;
      const uint32_t s = (uint32_t)i % NS, ph = ((uint32_t)i / NS) & 1;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      if (elect_one_sync()) {
        const uint32_t a_dst = smem_u32(smem + (size_t)s * STAGE), b_dst = a_dst + A_BYTES;
        mbar_expect_tx(bar_full + 8 * s, STAGE);
        if (EPI == EPI_POOL) tma_load_3d(a_dst, &tmapA, bar_full + 8 * s, kb * KBLK, 0, blockIdx.y * (128 / P.rows_per_seq));
        else tma_load_2d(a_dst, &tmapA, bar_full + 8 * s, kb * KBLK, m0);
        tma_load_2d(b_dst, &tmapB, bar_full + 8 * s, kb * KBLK, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc(128, BN, P.fmt);
    for (int kb = kb0; kb < kb1; ++kb) {
      const int i = kb - kb0;
      const uint32_t s = (uint32_t)i % NS, ph = ((uint32_t)i / NS) & 1;
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t ad = make_sw128_desc(smem_u32(smem + (size_t)s * STAGE));
        const uint64_t bd = make_sw128_desc(smem_u32(smem + (size_t)s * STAGE + A_BYTES));
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)       // 4 x (K = 16): both descriptors advance 32 bytes
          tc_mma_ss(tmem_base, (uint32_t)ad + 2 * k4, (uint32_t)(ad >> 32), (uint32_t)bd + 2 * k4, (uint32_t)(bd >> 32), idesc, (i | k4) ? 1u : 0u);
        tc_commit(bar_empty + 8 * s);        // stage free once these MMAs have read it
        if (kb == kb1 - 1) tc_commit(bar_acc);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: thread == output row =====
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int row = m0 + r;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    if (kb1 > kb0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      if (kb1 > kb0) {
        G_TMEM_LD_32(lane_base + (uint32_t)c, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (EPI == EPI_POOL) {
        // row r of the tile = position p of sequence (tile * seqs_per_tile + r / rows_per_seq); positions >= valid_pos and
        // sequences >= n_seq were zero-filled by TMA and are masked to 0 here (ReLU output is >= 0, so 0 never wins wrongly)
        const int rps = P.rows_per_seq;
        const int seq = blockIdx.y * (128 / rps) + r / rps, pos = r % rps;
        const bool live = seq < P.n_seq && pos < P.valid_pos;
        uint32_t keep = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c + j;
          float x = 0.f;
          if (live && col < P.N) x = fmaxf(__uint_as_float(v[j]) + __ldg(P.bias + col), 0.f);
          // non-negative floats order like their bit patterns: one REDUX per column over the 32 positions of this warp
          const uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint(x));
          if (lane == j) keep = mx;
        }
        // lane j holds the warp's max of column c + j; a sequence spans rows_per_seq / 32 warps -> atomic max on the bits
        const int col = n0 + c + lane;
        const int wseq = blockIdx.y * (128 / rps) + (quarter * 32) / rps;
        if (col < P.N && wseq < P.n_seq)
          atomicMax(reinterpret_cast<unsigned int*>(P.pool + (size_t)wseq * P.pool_ld + P.pool_off + col), keep);
      } else if (row < P.M) {
        float* drow = P.D + (size_t)row * P.ldd + n0 + c;
        if (EPI == EPI_ATOMIC) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c + j < P.N) atomicAdd(drow + j, P.alpha * __uint_as_float(v[j]));
        } else {
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = P.alpha * __uint_as_float(v[j]);
          const bool full = n0 + c + 32 <= P.N && ((reinterpret_cast<uintptr_t>(drow) & 15) == 0);
          if (full) {
            if (P.beta != 0.f) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 old = *reinterpret_cast<const float4*>(drow + j);
                o[j] = fmaf(P.beta, old.x, o[j]); o[j + 1] = fmaf(P.beta, old.y, o[j + 1]); o[j + 2] = fmaf(P.beta, old.z, o[j + 2]); o[j + 3] = fmaf(P.beta, old.w, o[j + 3]);
              }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(drow + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c + j < P.N) drow[j] = P.beta != 0.f ? fmaf(P.beta, drow[j], o[j]) : o[j];
          }
          if (P.D16) {
            uint16_t* d16 = P.D16 + (size_t)row * P.ldd16 + n0 + c;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c + j < P.N) d16[j] = to16(o[j], P.d16_fmt);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN < 32 ? 32 : BN));
  }
}

// ---------------------------------------------------------------- conversion kernels
__global__ void f32_to_16_kernel(const float* __restrict__ s, int64_t rows, int cols, int64_t lds, uint16_t* __restrict__ d, int64_t ldd, int fmt) {
  const int64_t total = rows * ldd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ldd;
    const int c = (int)(i - r * ldd);
    d[i] = c < cols ? to16(s[r * lds + c], fmt) : (uint16_t)0;
  }
}
// dst[c][r] (16-bit, leading dimension ldd >= rows, zero padded) = src[r][c] (fp32): 32 x 32 tiles through shared memory
__global__ void transpose_to_16_kernel(const float* __restrict__ s, int rows, int cols, int64_t lds, uint16_t* __restrict__ d, int64_t ldd, int fmt) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? s[(size_t)r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < ldd) d[(size_t)c * ldd + r] = to16(tile[threadIdx.x][i], fmt);
  }
}
// rows of the embedding table -> 16-bit [n_tok, ld] (zero padded to ld)
__global__ void gather_rows_16_kernel(const int32_t* __restrict__ tokens, int64_t n_tok, const float* __restrict__ emb, int We, int ld,
                                      uint16_t* __restrict__ out, int fmt) {
  const int64_t total = n_tok * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld;
    const int e = (int)(i - r * ld);
    out[i] = e < We ? to16(__ldg(emb + (size_t)tokens[r] * We + e), fmt) : (uint16_t)0;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}
// 16-bit [rows, K] with row stride ld (elements, multiple of 8), box = 64 k x box_rows rows, 128-byte swizzle, OOB -> 0
int make_map_2d(CUtensorMap* tm, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows, int fmt) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return SSE_ECUDA; }
  if (ld % 8 != 0 || (reinterpret_cast<uintptr_t>(base) & 15)) { set_error("gemm_tc: operand leading dimension %lld / base not 16-byte aligned", (long long)ld); return SSE_EINVAL; }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {KBLK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld ld=%lld", (int)r, (long long)rows, (long long)K, (long long)ld); return SSE_ECUDA; }
  return SSE_OK;
}

}  // namespace

int convert_to_16(const float* src, int64_t rows, int cols, int64_t lds, uint16_t* dst, int64_t ldd, int fmt, cudaStream_t st, int64_t* launches) {
  if (rows <= 0) return SSE_OK;
  f32_to_16_kernel<<<(int)std::min<int64_t>(cdiv64(rows * ldd, 256), 148 * 16), 256, 0, st>>>(src, rows, cols, lds, dst, ldd, fmt);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}
int transpose_to_16(const float* src, int rows, int cols, int64_t lds, uint16_t* dst, int64_t ldd, int fmt, cudaStream_t st, int64_t* launches) {
  if (rows <= 0 || cols <= 0) return SSE_OK;
  dim3 grid(cdiv(cols, 32), (unsigned)cdiv64(ldd, 32)), block(32, 8);
  transpose_to_16_kernel<<<grid, block, 0, st>>>(src, rows, cols, lds, dst, ldd, fmt);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}
// the common case (We = ld, multiple of 8, 16-byte aligned rows): 8 elements per thread -- two float4 loads, one 16-byte store
__global__ void gather_rows_16_vec_kernel(const int32_t* __restrict__ tokens, int64_t n_tok, const float* __restrict__ emb, int We,
                                          uint16_t* __restrict__ out, int fmt) {
  const int per_row = We >> 3;
  const int64_t total = n_tok * per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / per_row;
    const int e = (int)(i - r * per_row) << 3;
    const float* src = emb + (size_t)__ldg(tokens + r) * We + e;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    uint4 o;
    if (fmt == 1) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w), p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
      o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1); o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
    } else {
      __half2 p0 = __floats2half2_rn(a.x, a.y), p1 = __floats2half2_rn(a.z, a.w), p2 = __floats2half2_rn(b.x, b.y), p3 = __floats2half2_rn(b.z, b.w);
      o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1); o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
    }
    *reinterpret_cast<uint4*>(out + (size_t)r * We + e) = o;
  }
}

int gather_rows_16(const int32_t* tokens, int64_t n_tok, const float* emb, int We, int ld, uint16_t* out, int fmt, cudaStream_t st, int64_t* launches) {
  if (n_tok <= 0) return SSE_OK;
  if (ld == We && (We & 7) == 0 && ((reinterpret_cast<uintptr_t>(emb) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    gather_rows_16_vec_kernel<<<(int)std::min<int64_t>(cdiv64(n_tok * (We >> 3), 256), 148 * 16), 256, 0, st>>>(tokens, n_tok, emb, We, out, fmt);
    if (launches) ++*launches;
    SSE_CUDA_OK(cudaGetLastError());
    return SSE_OK;
  }
  gather_rows_16_kernel<<<(int)std::min<int64_t>(cdiv64(n_tok * ld, 256), 148 * 16), 256, 0, st>>>(tokens, n_tok, emb, We, ld, out, fmt);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

// D[M,N] = alpha * A[M,K] B[N,K]^T (+ beta D); split_k > 1: partial sums are atomically ADDED to D (beta is taken as 1)
// One CTA computes one output tile and nothing overlaps its prologue (barriers, TMEM allocation, first TMA round trip) and
// its epilogue with a mainloop -- unless a second CTA lives on the same SM.  A grid of several waves therefore runs with a
// 3-stage ring (<= 113 KB of shared memory per CTA: two CTAs per SM, 128 TMEM columns each) instead of the deepest ring that
// fits alone.  SSE_GEMM_STAGES=n forces the ring depth (experiments).
static int multiwave_stages(int n_stages, int64_t n_ctas, size_t stage_bytes) {
  static const int env = getenv("SSE_GEMM_STAGES") ? atoi(getenv("SSE_GEMM_STAGES")) : 0;
  if (env > 0) return std::max(2, std::min(G_STAGES_MAX, std::min(env, n_stages)));
  if (n_ctas > 148) {
    const int fit2 = (int)((113 * 1024 - 1024 - 256) / stage_bytes);
    if (fit2 >= 2) return std::min(n_stages, fit2);
  }
  return n_stages;
}

int gemm_tc(const uint16_t* A, int64_t lda, const uint16_t* B, int64_t ldb, int M, int N, int K, float alpha, float beta, float* D, int64_t ldd,
            int fmt, int split_k, uint16_t* D16, int64_t ldd16, cudaStream_t st, int64_t* launches) {
  if (M <= 0 || N <= 0) return SSE_OK;
  if (K <= 0) { set_error("gemm_tc: K must be positive"); return SSE_EINVAL; }
  CUtensorMap ta, tb;
  // 64-column tiles when 128-column tiles would leave SMs without a CTA (the per-step GEMMs of the train step and of the wide
  // LSTM tower: 96 / 80 tiles): twice the CTAs, half the epilogue each, and two of them fit an SM.  SSE_GEMM_BN=128 restores.
  static const int env_bn = getenv("SSE_GEMM_BN") ? atoi(getenv("SSE_GEMM_BN")) : 0;
  const int64_t tiles128 = (int64_t)cdiv(N, 128) * cdiv(M, 128) * std::max(1, split_k);
  const int BN = N <= 64 ? 64 : (env_bn == 128 ? 128 : (env_bn == 64 || tiles128 < 148 ? 64 : 128));
  SSE_TRY(make_map_2d(&ta, A, M, K, lda, 128, fmt));
  SSE_TRY(make_map_2d(&tb, B, N, K, ldb, BN, fmt));
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.D = D; p.ldd = ldd; p.D16 = D16; p.ldd16 = ldd16; p.d16_fmt = fmt; p.fmt = fmt;
  const int nkb = cdiv(K, KBLK);
  if (split_k < 1) split_k = 1;
  if (split_k > nkb) split_k = nkb;
  p.kb_per_split = cdiv(nkb, split_k);
  split_k = cdiv(nkb, p.kb_per_split);
  const size_t stage = (size_t)128 * KBLK * 2 + (size_t)BN * KBLK * 2;
  p.n_stages = std::min(G_STAGES_MAX, std::max(2, p.kb_per_split));
  dim3 grid(cdiv(N, BN), cdiv(M, 128), split_k);
  p.n_stages = multiwave_stages(p.n_stages, (int64_t)grid.x * grid.y * grid.z, stage);
  const size_t smem = 1024 + (size_t)p.n_stages * stage + 256;
  const bool atomic = split_k > 1;
#define LAUNCH_GEMM(BN_, EPI_)                                                                                               \
  do {                                                                                                                        \
    static bool attr_set = false;          /* once per instantiation: the call costs about as much as a small GEMM */        \
    if (!attr_set) {                                                                                                          \
      SSE_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN_, EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(1024 + G_STAGES_MAX * ((size_t)128 * KBLK * 2 + (size_t)BN_ * KBLK * 2) + 256))); \
      attr_set = true;                                                                                                        \
    }                                                                                                                         \
    gemm_tc_kernel<BN_, EPI_><<<grid, G_THREADS, smem, st>>>(ta, tb, p);                                                      \
  } while (0)
  if (BN == 64) { if (atomic) LAUNCH_GEMM(64, EPI_ATOMIC); else LAUNCH_GEMM(64, EPI_STORE); }
  else { if (atomic) LAUNCH_GEMM(128, EPI_ATOMIC); else LAUNCH_GEMM(128, EPI_STORE); }
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

// CNN n-gram convolution + bias + ReLU + max over positions, fused (reference sse_model.py:190-202):
//   pool[b, off + f] = max_p relu( sum_{j<kf, e<We} X[b, p + j, e] * Wt[f, j * ldx + e] + bias[f] ),  p in [0, T - kf + 1)
// X: 16-bit [n_seq, T, ldx] gathered embeddings (ldx = We padded to a multiple of 8, pad = 0); a window of kf positions is
// the CONTIGUOUS slice of kf * ldx elements starting at (b, p): the A operand is a 3-D tensor map (k, position, sequence)
// with overlapping rows, no im2col.  Wt: 16-bit [F, kf * ldx] (filter transposed, same padding).  pool must be zeroed.
int cnn_conv_pool_tc(const uint16_t* X, int n_seq, int T, int ldx, int kf, const uint16_t* Wt, int F, const float* bias, float* pool, int pool_ld,
                     int pool_off, int fmt, cudaStream_t st, int64_t* launches) {
  const int P = T - kf + 1;
  if (P <= 0) { set_error("cnn: filter width %d > max_seq_length %d", kf, T); return SSE_EINVAL; }
  if (P > 128) { set_error("cnn_conv_pool_tc: %d positions per sequence exceed one 128-row tile", P); return SSE_EINVAL; }
  const int rps = P <= 32 ? 32 : (P <= 64 ? 64 : 128);
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return SSE_ECUDA; }
  const int K = kf * ldx;
  CUtensorMap ta, tb;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)P, (cuuint64_t)n_seq};
    cuuint64_t gstr[2] = {(cuuint64_t)ldx * 2, (cuuint64_t)T * ldx * 2};
    cuuint32_t box[3] = {KBLK, (cuuint32_t)rps, (cuuint32_t)(128 / rps)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&ta, fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(X), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cnn_conv_pool_tc: tensor map failed (%d)", (int)r); return SSE_ECUDA; }
  }
  const int BN = F <= 64 ? 64 : 128;
  SSE_TRY(make_map_2d(&tb, Wt, F, K, K, BN, fmt));
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = n_seq * rps; p.N = F; p.K = K; p.alpha = 1.f; p.fmt = fmt;
  p.rows_per_seq = rps; p.valid_pos = P; p.n_seq = n_seq; p.bias = bias; p.pool = pool; p.pool_ld = pool_ld; p.pool_off = pool_off;
  const int nkb = cdiv(K, KBLK);
  p.kb_per_split = nkb;
  const size_t stage = (size_t)128 * KBLK * 2 + (size_t)BN * KBLK * 2;
  p.n_stages = std::min(G_STAGES_MAX, std::max(2, nkb));
  dim3 grid(cdiv(F, BN), cdiv(n_seq, 128 / rps), 1);
  p.n_stages = multiwave_stages(p.n_stages, (int64_t)grid.x * grid.y, stage);
  const size_t smem = 1024 + (size_t)p.n_stages * stage + 256;
  if (BN == 64) LAUNCH_GEMM(64, EPI_POOL); else LAUNCH_GEMM(128, EPI_POOL);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

}  // namespace sse
