// LSTM tower for small / medium batches (query encoding): a thread-block CLUSTER of H/32 CTAs owns 64 batch
// rows for all T steps, and the fp16 weights never move after the first microsecond.
// Restates BasicLSTMCell(forget_bias=1)+static_rnn (reference sse_model.py:222-224,240-242,248-250,262-264,
// 273-274), same numerics as lstm_tc.cu (fp16 operands, fp32 accumulate, fp32 cell state and gate math).
//
// Why a second kernel: lstm_tc.cu streams the whole [4H, We+H] weight matrix (1 MB at We=H=256) through every
// CTA on every step.  That is the right trade when all 148 SMs hold full 128-row tiles (index build), but a
// 600-row query batch then runs at the per-SM L2->smem rate: ~16 us per step, 0.8 ms per batch.  Here
//   * CTA `rank` of the cluster keeps the weight slice of hidden units [32 rank, 32 rank + 32) -- 128 gate
//     columns (i|j|f|o x 32), [128 x (We+H)] fp16 = 128 KB -- RESIDENT in shared memory (one TMA load),
//   * per step it issues 4 (We/64 + H/64) tcgen05.mma (M=128 with 64 live rows, N=128, A and B from smem),
//   * its four epilogue warps (thread == batch row x 16 units, c in registers) apply the gate math and store the new
//     h slice (fp16) straight into the h_t operand tile of EVERY CTA of the cluster through distributed shared
//     memory, already in the 128B-swizzled K-major layout the next step's MMA reads,
//   * one remote mbarrier arrival per (warp, destination) publishes it; h tiles ping-pong so there is exactly
//     one cluster-wide dependency per step and no cluster barrier.
// The x_{t+1} part of the next step is issued before h_t is complete (it does not depend on the recurrence),
// so the tensor pipe works through the exchange latency.
//
// shared memory (We=H=256): x tile 32 KB | h ping 32 KB | h pong 32 KB | Wx slice 64 KB | Wh slice 64 KB | bias
// TMEM: two [128 lanes x 128 col] fp32 accumulators.
// warps: 0,1,4,5 epilogue (TMEM lanes 0-63, 16 units per thread) | 2 MMA issuer + TMEM alloc + weight load |
//        3,7 embedding gather | 6 idle (so that gather / MMA / epilogue warps sit on different SM sub-partitions).
#include "sse_common.cuh"
#include <algorithm>
#include <cuda.h>
#include <math_constants.h>
#include <stdlib.h>
#include <vector>

namespace sse {

namespace {

constexpr int KBLK = 64;
constexpr int W_TILE_BYTES = 128 * KBLK * 2;   // [128 gate cols x 64 k] fp16, SW128
constexpr int CL_ROWS = 64;                    // batch rows per cluster
constexpr int A_TILE_BYTES = CL_ROWS * 128;    // [64 rows x 64 k] fp16, SW128
constexpr int CL_THREADS = 256;
constexpr int SLICE_BYTES = CL_ROWS * 64;       // one CTA's h slice: [64 rows x 32 units] fp16, 64B-swizzled K-major atom

struct LstmClParams {
  const int32_t* tokens;      // [B, T]
  const __half* emb;          // [V, We] fp16
  const float* bias_r;        // [4H] chunk-major, pre-scaled for the ex2 gate math (lstm_tc.cu prep_weights_kernel)
  const float* init_h;        // optional [H] broadcast initial state (pad-prefix table row)
  const float* init_c;
  const int32_t* lead_sorted; // optional [B]: leading PADs of each (sorted) row -> per-tile start step (tok_prep.cu)
  const float* pad_h;         // with lead_sorted: pad-prefix state table [T][H] (entry t = state after t+1 PADs)
  const float* pad_c;
  float* h_out;               // [B, H] fp32 (last step)
  int B, T, t_start, We, H;
  long long* dbg;             // optional [grid][8] cycle counters
  int dbg_flags;              // timing experiments only: 1 = store h to the own CTA only, 2 = skip the MUFU gate math
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
// arrival on a barrier that lives in ANOTHER CTA of the cluster; releases this thread's (and, after a
// __syncwarp, its warp's) distributed-shared-memory stores at cluster scope
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// CLUSTER = true: acquire at cluster scope (the barrier is completed by remote arrivals that publish DSMEM stores)
template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t backoff_ns = 32) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    if (CLUSTER)
      asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    else
      asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (backoff_ns) __nanosleep(backoff_ns);
    if ((spins & 0xfff) == 0xfff) {          // watchdog: a protocol bug becomes a trap, not a hung GPU
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}
template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait_timed(uint32_t bar, uint32_t parity, long long& acc, uint32_t backoff_ns = 32) {
  long long t = clock64();
  mbar_wait<CLUSTER>(bar, parity, backoff_ns);
  acc += clock64() - t;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %1;\n@%%px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// K-major tile whose rows are 64 bytes (32 fp16): SWIZZLE_64B atoms, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// K-major tile whose rows are 32 bytes (16 fp16 = one k-step): SWIZZLE_32B atoms, 8-row groups 256 B apart
__device__ __forceinline__ uint64_t make_sw32_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
// shared memory of this CTA -> shared memory of a peer CTA through the bulk-copy engine; completes (bytes) on the
// PEER's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t peer_dst, uint32_t local_src, uint32_t bytes, uint32_t peer_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(peer_dst), "r"(local_src),
               "r"(bytes), "r"(peer_bar)
               : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // D=f32, A=B=f16
}
// descriptor given as (lo, hi) words: only the low word (address field) varies per MMA
// A and B descriptors with different high words (different swizzle modes)
__device__ __forceinline__ void tc_mma_ss3(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 ad, bd;\nsetp.ne.b32 p, %6, 0;\nmov.b64 ad, {%1, %2};\nmov.b64 bd, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n}\n" ::"r"(d),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_mma_ss2(uint32_t d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 ad, bd;\nsetp.ne.b32 p, %5, 0;\nmov.b64 ad, {%1, %3};\nmov.b64 bd, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, p;\n}\n" ::"r"(d),
      "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void lds_v4(uint32_t addr, float* v) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}

// 256-bit read-only global load (sm_100: LDG.E.256): one 32-byte sector per lane per instruction
__device__ __forceinline__ void ldg_v8(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}

#define TMEM_LD_8(taddr, v)                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                             \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])      \
               : "r"(taddr))

// ------------------------------------------------------------------ the kernel
// grid = n_clusters * CL CTAs, cluster = CL = H / 32 CTAs.  Cluster q owns batch rows [64 q, 64 q + 64).
__global__ void __launch_bounds__(CL_THREADS, 1)
lstm_cluster_kernel(const __grid_constant__ CUtensorMap tmap_w2d, const __grid_constant__ LstmClParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KBx = P.We / KBLK, KBh = P.H / KBLK;
  const int CL = P.H / 32;
  const uint32_t rank = cluster_ctarank();
  const int row0 = (blockIdx.x / CL) * CL_ROWS;
  // per-tile pad-prefix start (tok_prep.cu): rows are sorted by descending number of leading PADs, so the tile's LAST row
  // has the shortest prefix; the tile starts at t0 from the tabulated state after t0 PADs
  int t0 = P.t_start;
  const float* init_h = P.init_h;
  const float* init_c = P.init_c;
  if (P.lead_sorted) {
    t0 = min(__ldg(P.lead_sorted + min(row0 + CL_ROWS - 1, P.B - 1)), P.T - 1);
    init_h = t0 > 0 ? P.pad_h + (size_t)(t0 - 1) * P.H : nullptr;
    init_c = t0 > 0 ? P.pad_c + (size_t)(t0 - 1) * P.H : nullptr;
  }
  const bool has_init = init_h != nullptr;
  const int nsteps = P.T - t0;

  // The A tiles hold 64 rows; every MMA still reads 128 rows per k-block, i.e. 8 KB past the tile: those bytes are
  // the next tile / the weight slices (finite fp16 bit patterns) and only feed accumulator lanes 64..127 nobody reads.
  uint8_t* x_smem = smem;                                          // [KBx] tiles [64 x 64] fp16 SW128
  uint8_t* h_smem = x_smem + (size_t)KBx * A_TILE_BYTES;           // [2 tiles][CL slices] of [64 rows x 32 k] fp16 SW64 (slice q = CTA q's units)
  uint8_t* wx_smem = h_smem + (size_t)2 * CL * SLICE_BYTES;        // [KBx] tiles [128 x 64]
  uint8_t* wh_smem = wx_smem + (size_t)KBx * W_TILE_BYTES;         // [KBh] tiles
  float* bias_s = reinterpret_cast<float*>(wh_smem + (size_t)KBh * W_TILE_BYTES);   // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 128);
  const uint32_t bar_wf = smem_u32(bars + 0);
  const uint32_t bar_xf = smem_u32(bars + 1);
  const uint32_t bar_xe = smem_u32(bars + 2);
  const uint32_t bar_accf = smem_u32(bars + 3);    // [2]
  const uint32_t bar_acce = smem_u32(bars + 5);    // [2]
  const uint32_t bar_hr = smem_u32(bars + 7);      // [2] h tile complete: own slice in place + expect_tx of the CL-1 peer slices
  const uint32_t bar_sl = smem_u32(bars + 9);      // [2] own slice written by the 4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  if (threadIdx.x < 128) bias_s[threadIdx.x] = P.bias_r[rank * 128 + threadIdx.x];
  if (threadIdx.x == 0) {
    mbar_init(bar_wf, 1);
    mbar_init(bar_xf, 64);
    mbar_init(bar_xe, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accf + 8 * b, 1);
      mbar_init(bar_acce + 8 * b, 4);
      mbar_init(bar_hr + 8 * b, 2);
      mbar_init(bar_sl + 8 * b, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();                    // every CTA's barriers exist before anyone stores / arrives remotely
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 2) {
    // ===== weight slice load (once) + MMA issuer: warp converged, one elected lane issues =====
    if (elect_one_sync()) {
      mbar_expect_tx(bar_wf, (uint32_t)(KBx + KBh) * W_TILE_BYTES);
      for (int kb = 0; kb < KBx; ++kb) tma_load_2d(smem_u32(wx_smem + (size_t)kb * W_TILE_BYTES), &tmap_w2d, bar_wf, kb * KBLK, (int)rank * 128);
      for (int kb = 0; kb < KBh; ++kb) tma_load_2d(smem_u32(wh_smem + (size_t)kb * W_TILE_BYTES), &tmap_w2d, bar_wf, (KBx + kb) * KBLK, (int)rank * 128);
    }
    __syncwarp();
    const uint32_t idesc = make_idesc_f16(128, 128);
    const uint64_t xd = make_sw128_desc(smem_u32(x_smem)), wxd = make_sw128_desc(smem_u32(wx_smem)), whd = make_sw128_desc(smem_u32(wh_smem));
    const uint32_t hi = (uint32_t)(xd >> 32);
    const uint32_t x_lo = (uint32_t)xd, wx_lo = (uint32_t)wxd, wh_lo = (uint32_t)whd;
    const uint64_t hd = make_sw64_desc(smem_u32(h_smem));
    const uint32_t h_lo0 = (uint32_t)hd, h_hi = (uint32_t)(hd >> 32);
    long long w_hr = 0, w_xf = 0, w_acce = 0, w_ih = 0, w_ix = 0, t_begin = clock64();
    mbar_wait<false>(bar_wf, 0);

    // x part of step s (local index) into accumulator s & 1
    auto issue_x = [&](int s) {
      const int buf = s & 1;
      if (s >= 2) { mbar_wait_timed<false>(bar_acce + 8 * buf, (uint32_t)(((s >> 1) - 1) & 1), w_acce); }
      mbar_wait_timed<false>(bar_xf, (uint32_t)(s & 1), w_xf);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t d = tmem_base + (uint32_t)(buf * 128);
        for (int kb = 0; kb < KBx; ++kb)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            tc_mma_ss2(d, x_lo + (uint32_t)(kb * (A_TILE_BYTES >> 4) + 2 * k4), wx_lo + (uint32_t)(kb * (W_TILE_BYTES >> 4) + 2 * k4), hi, idesc,
                       (kb | k4) ? 1u : 0u);
        tc_commit(bar_xe);                                         // x tile consumed once these retire
        if (s == 0 && !has_init) tc_commit(bar_accf + 8 * buf);    // no recurrent part at the very first step
      }
      __syncwarp();
    };

    issue_x(0);
    for (int s = 0; s < nsteps; ++s) {
      const int buf = s & 1;
      if (s > 0 || has_init) {
        // h_{s-1} was written into tile (s-1)&1 (the initial state counts as step -1 -> tile 1)
        const int hb = (s - 1) & 1;
        const uint32_t n = s == 0 ? 0u : (uint32_t)(((s - 1) >> 1) + ((hb == 1 && has_init) ? 1 : 0));
        // this phase of h_ready needs: the own slice (copy-issuer warp's arrival) + the CL-1 peer slices (bytes)
        if (elect_one_sync()) {
          if (s == 0) mbar_arrive(bar_hr + 8 * hb);                       // initial state: filled locally, no copies
          else mbar_expect_tx(bar_hr + 8 * hb, (uint32_t)(CL - 1) * SLICE_BYTES);
        }
        __syncwarp();
        mbar_wait_timed<false>(bar_hr + 8 * hb, n & 1, w_hr);
        long long ti0 = clock64();
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t d = tmem_base + (uint32_t)(buf * 128);
          const uint32_t h_lo = h_lo0 + (uint32_t)(hb * CL * (SLICE_BYTES >> 4));
          // k-step m covers hidden units [16 m, 16 m + 16): A = slice m/2 (SW64 rows, half m%2), B = k-block m/4 (SW128 rows, quarter m%4)
          for (int m = 0; m < 2 * CL; ++m)
            tc_mma_ss3(d, h_lo + (uint32_t)((m >> 1) * (SLICE_BYTES >> 4) + 2 * (m & 1)), h_hi,
                       wh_lo + (uint32_t)((m >> 2) * (W_TILE_BYTES >> 4) + 2 * (m & 3)), hi, idesc, 1u);
          tc_commit(bar_accf + 8 * buf);
        }
        __syncwarp();
        w_ih += clock64() - ti0;
      }
      if (s + 1 < nsteps) { long long tx0 = clock64(); issue_x(s + 1); w_ix += clock64() - tx0; }
    }
    if (P.dbg && lane == 0) {
      long long* o = P.dbg + blockIdx.x * 16;
      o[0] = w_hr; o[1] = w_xf; o[2] = w_acce; o[3] = clock64() - t_begin; o[10] = w_ih; o[11] = w_ix;
    }
  } else if (warp == 3 || warp == 7) {
    // ===== embedding gather (1 row per lane): cp.async 16 B chunks into the 128B-swizzled x tile =====
    const int r = (warp == 7 ? 32 : 0) + lane;
    const int grow = min(row0 + r, P.B - 1);
    const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const uint32_t sw = (uint32_t)(r & 7);
    const uint32_t xbase = smem_u32(x_smem) + row_off;
    int tok = __ldg(P.tokens + (size_t)grow * P.T + t0);
    for (int s = 0; s < nsteps; ++s) {
      if (s > 0) mbar_wait<false>(bar_xe, (uint32_t)((s - 1) & 1), 64);
      const __half* src = P.emb + (size_t)tok * P.We;
      for (int kb = 0; kb < KBx; ++kb) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t dst = xbase + (uint32_t)kb * A_TILE_BYTES + (((uint32_t)j ^ sw) * 16);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + kb * KBLK + j * 8) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (s + 1 < nsteps) tok = __ldg(P.tokens + (size_t)grow * P.T + t0 + s + 1);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_xf);
    }
  } else if (warp == 6) {
    // ===== slice exchange: once the four epilogue warps have written this CTA's h_s slice into the local tile,
    //       push it to the same place in every peer with the bulk-copy engine (completes on the peer's h_ready)
    uint32_t peer_h = map_to_cta(smem_u32(h_smem), (uint32_t)(lane < CL ? lane : 0));
    uint32_t peer_bar = map_to_cta(bar_hr, (uint32_t)(lane < CL ? lane : 0));
    if (has_init) {
      mbar_wait<false>(bar_sl + 8, 0);
      if (lane == 0) mbar_arrive(bar_hr + 8);
    }
    for (int s = 0; s + 1 < nsteps; ++s) {
      const int tb = s & 1;
      const uint32_t n = (uint32_t)((s >> 1) + ((tb == 1 && has_init) ? 1 : 0));
      mbar_wait<false>(bar_sl + 8 * tb, n & 1);
      const uint32_t off = (uint32_t)(tb * CL * SLICE_BYTES) + rank * SLICE_BYTES;
      if (lane < CL && lane != (int)rank) bulk_copy_to_peer(peer_h + off, smem_u32(h_smem) + off, SLICE_BYTES, peer_bar + 8 * tb);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hr + 8 * tb);
    }
  } else if ((warp & 3) < 2) {
    // ===== epilogue: warps 0,1,4,5; thread == (batch row, 16 of this CTA's 32 hidden units), c in registers.
    //       Two warps per SM sub-partition hide each other's MUFU / TMEM latencies.
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const int grow = row0 + r;
    const bool valid = grow < P.B;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const uint32_t sw = (uint32_t)(r & 7);
    const int ub = half * 16;                      // first of this thread's units within the CTA slice
    float c[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) c[u] = has_init ? __ldg(init_c + rank * 32 + ub + u) : 0.f;
    // row r of a slice: 64 bytes at r*64, 16-byte chunk c stored at c ^ ((r >> 1) & 3)  (SWIZZLE_64B)
    const uint32_t row_off64 = (uint32_t)(r * 64);
    const uint32_t sw64 = (uint32_t)((r >> 1) & 3);
    if (has_init) {
      // the broadcast initial state is the same for every row: each CTA fills its own tile 1 (= "step -1"), all slices
      const uint32_t t1 = smem_u32(h_smem) + (uint32_t)(CL * SLICE_BYTES) + row_off64;
      for (int q = half; q < CL; q += 2)
        for (int j = 0; j < 4; ++j) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) pk[e] = pack_f16x2(__ldg(init_h + q * 32 + j * 8 + 2 * e), __ldg(init_h + q * 32 + j * 8 + 2 * e + 1));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t1 + (uint32_t)q * SLICE_BYTES + (((uint32_t)j ^ sw64) * 16)),
                       "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sl + 8);       // completion 0 of slice_ready[1] == "initial state staged"
    }
    // this thread's 16 units inside the CTA's own slice: 16-byte chunks half*2 + p
    const uint32_t own_slice = smem_u32(h_smem) + rank * SLICE_BYTES + row_off64;
    const uint32_t bias_a = smem_u32(bias_s) + (uint32_t)(ub * 4);
    constexpr float NL2E = -1.4426950408889634f;
    long long w_accf = 0, w_ld = 0, w_math = 0, w_st = 0, w_pub = 0, tq = 0, t_begin = clock64();
    const bool dbgt = P.dbg != nullptr;
    for (int s = 0; s < nsteps; ++s) {
      const int buf = s & 1;
      const bool last = s == nsteps - 1;
      mbar_wait_timed<false>(bar_accf + 8 * buf, (uint32_t)((s >> 1) & 1), w_accf);
      tc_fence_after();
      const uint32_t acc = lane_base + (uint32_t)(buf * 128 + ub);
      const uint32_t tile_own = own_slice + (uint32_t)(buf * CL * SLICE_BYTES);        // h_s goes to tile s & 1
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t vi[8], vj[8], vf[8], vo[8];
        if (dbgt) tq = clock64();
        TMEM_LD_8(acc + p * 8, vi);
        TMEM_LD_8(acc + 32 + p * 8, vj);
        TMEM_LD_8(acc + 64 + p * 8, vf);
        TMEM_LD_8(acc + 96 + p * 8, vo);
        // this pass's 32 bias values (warp-uniform addresses: broadcast reads)
        float bi[8], bj[8], bf[8], bo[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          lds_v4(bias_a + (uint32_t)((p * 8 + q * 4) * 4), bi + q * 4);
          lds_v4(bias_a + (uint32_t)((32 + p * 8 + q * 4) * 4), bj + q * 4);
          lds_v4(bias_a + (uint32_t)((64 + p * 8 + q * 4) * 4), bf + q * 4);
          lds_v4(bias_a + (uint32_t)((96 + p * 8 + q * 4) * 4), bo + q * 4);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (p == 1) {                       // the accumulator sits in registers: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acce + 8 * buf);
        }
        if (dbgt) { long long t = clock64(); w_ld += t - tq; tq = t; }
        // gate math, 8 MUFU per (row, unit) -- see lstm_tc.cu: sig(i) tanh(j) = (1-Ej)/((1+Ei)(1+Ej)) etc.,
        // the bias table is pre-scaled by -log2 e (-2 log2 e for j, forget bias folded in)
        float hv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int u = p * 8 + i;
          if (P.dbg_flags & 2) { c[u] += __uint_as_float(vi[i]) + __uint_as_float(vj[i]); hv[i] = c[u] * __uint_as_float(vf[i]) + __uint_as_float(vo[i]); continue; }
          const float ei = ex2_approx(fminf(fmaf(__uint_as_float(vi[i]), NL2E, bi[i]), 57.f));
          const float ej = ex2_approx(fminf(fmaf(__uint_as_float(vj[i]), 2.f * NL2E, bj[i]), 57.f));
          const float ef = ex2_approx(fminf(fmaf(__uint_as_float(vf[i]), NL2E, bf[i]), 57.f));
          const float eo = ex2_approx(fminf(fmaf(__uint_as_float(vo[i]), NL2E, bo[i]), 57.f));
          const float pj = (1.f - ej) * rcp_approx((1.f + ei) * (1.f + ej));
          c[u] = fmaf(c[u], rcp_approx(1.f + ef), pj);
          const float ec = ex2_approx(fminf(c[u] * (2.f * NL2E), 57.f));
          hv[i] = (1.f - ec) * rcp_approx((1.f + ec) * (1.f + eo));
        }
        if (dbgt) { long long t = clock64(); w_math += t - tq; tq = t; }
        if (!last) {
          const uint32_t p0 = pack_f16x2(hv[0], hv[1]), p1 = pack_f16x2(hv[2], hv[3]), p2 = pack_f16x2(hv[4], hv[5]), p3 = pack_f16x2(hv[6], hv[7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_own + ((((uint32_t)(half * 2 + p)) ^ sw64) * 16)),
                       "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
        } else if (valid) {
          float* ho = P.h_out + (size_t)grow * P.H + rank * 32 + ub + p * 8;
          *reinterpret_cast<float4*>(ho) = make_float4(hv[0], hv[1], hv[2], hv[3]);
          *reinterpret_cast<float4*>(ho + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
        }
        if (dbgt) { long long t = clock64(); w_st += t - tq; tq = t; }
      }
      if (!last) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the bulk-copy engine / tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sl + 8 * buf);
        if (dbgt) { long long t = clock64(); w_pub += t - tq; tq = t; }
      }
    }
    if (P.dbg && lane == 0 && warp == 0) {
      long long* o = P.dbg + blockIdx.x * 16;
      o[4] = w_accf; o[5] = clock64() - t_begin; o[6] = w_ld; o[7] = w_math; o[8] = w_st; o[9] = w_pub;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // nobody leaves while a peer may still address its shared memory
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}


// =====================================================================================================
// Variant 2: input projection from a table.  x_t W_x + b depends only on the token id, so it is tabulated once
// per weight set: P[v, :] = scale_g * (Emb[v] W_x + b) for every vocabulary entry v (fp32, chunk-major gate
// columns, already in the ex2 argument scale).  The step then needs NO x tile, NO W_x slice and no gather warps:
//   z = scale_g * (h_{t-1} W_h)  +  P[token]          (one FMA per gate value in the epilogue)
// which frees enough shared memory for 128 batch rows per cluster (all four TMEM lane quarters / SM sub-partitions
// carry live rows: 8 epilogue warps, 16 units per thread), halves the tensor work per step and makes the x part
// exact fp32.  Everything else is the scheme above: W_h slice resident, h slices exchanged with the bulk-copy
// engine into SWIZZLE_64B K-major tiles, one dependency per step.
// shared memory (H=256): h ping 64 KB | h pong 64 KB | Wh slice 64 KB.  TMEM: one [128 x 128] fp32 accumulator.
// warps: 0-7 epilogue (warp w: lane quarter w%4, unit half w/4) | 8 MMA issuer + TMEM alloc + weight load | 9 slice exchange
constexpr int P2_ROWS = 128;
constexpr int P2_SUB_BYTES = P2_ROWS * 32;      // one sub-slice: [128 rows x 16 units] fp16, SWIZZLE_32B K-major
constexpr int P2_TRACE_STEPS = 48;              // debug: per-step timestamps of CTA 0
constexpr int P2_DEFAULT_EW = 16, P2_DEFAULT_GATE = 1;

struct LstmP2Params {
  const int32_t* tokens;      // [B, T]
  const float* ptable;        // [V, 4H] fp32, scaled, bias folded in; column order: see ptab_col()
  const float* init_h;
  const float* init_c;
  const int32_t* lead_sorted; // optional [B]: leading PADs of each (sorted) row -> per-tile start step (tok_prep.cu)
  const float* pad_h;         // with lead_sorted: pad-prefix state table [T][H] (entry t = state after t+1 PADs)
  const float* pad_c;
  float* dump_h;              // optional [T][H]: row 0's state after every step (h as the fp16 value the recurrence carries, c fp32):
  float* dump_c;              // run on one all-PAD row this IS the pad-prefix table of this kernel's own arithmetic
  float* h_out;               // [B, H]
  int B, T, t_start, H;
  int poll_ns;                // back-off of the MMA / exchange warps' barrier polls (0 = none)
  int flags;                  // timing experiments (SSE_LSTM_FLAGS): 1 = producer skips its loads, 2 = producer throttled, 4 = producer fetches after the exchange
  long long* dbg;             // optional [grid][16] cycle counters, then [P2_TRACE_STEPS][8] timestamps of CTA 0
};

// Column order of the table (and of W_x / the bias it is built from).  The epilogue thread (row r, unit group `sub`) of
// pass p needs, for its UPP = 16 / NSUB units, all four gate values: they sit in ONE contiguous run of 4 * UPP floats
//   col(c, p, sub, g, i) = c * 128 + p * 64 + sub * 4 * UPP + g * UPP + i        (unit = 32 c + 16 p + UPP sub + i)
// so a pass is one 128-byte (8 epilogue warps) or 64-byte (16 warps) gather per thread instead of four 32-byte pieces of
// four different lines (the L1 data pipe was ~40 % busy with the chunk-major order).
__host__ __device__ __forceinline__ int ptab_col(int c, int g, int j, int upp) {
  const int p = j >> 4, w = j & 15;
  return c * 128 + p * 64 + (w / upp) * 4 * upp + g * upp + (w % upp);
}

#define TMEM_LD_4(taddr, v)                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"                       \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])                                    \
               : "r"(taddr))

__device__ __forceinline__ uint32_t tanh_f16x2(uint32_t x) {
  uint32_t y;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

// EW: epilogue warps (8: thread == batch row x 16 units, 16: x 8 units -- four warps per SM sub-partition keep the MUFU
// pipe fed).  GATE: 0 = ex2 / rcp form (8 MUFU per unit), 1 = tanh.approx form (sigmoid(z) = 0.5 tanh(z/2) + 0.5: 5 MUFU),
// 2 = form 1 with tanh(o/2) and tanh(c) of a unit computed by ONE tanh.approx.f16x2 (4 MUFU; their product h is rounded
// to fp16 for the recurrence anyway).
// warps: 0..EW-1 epilogue (warp w: TMEM lane quarter w % 4, unit group w / 4) | EW: MMA issuer + TMEM alloc + weight
// load | EW+1: slice exchange
// ROWS: batch rows per cluster.  128 fills every TMEM lane quarter; 64 (query batches small enough to give every cluster its own
// SMs) halves a CTA's epilogue, gather and exchange work per step -- the M = 128 MMAs then read 64 rows of whatever follows
// the sub-slice (finite fp16 bits feeding accumulator lanes nobody reads) and the warps of lane quarters 2 / 3 have nothing to do.
template <int EW, int GATE, int ROWS>
__global__ void __launch_bounds__((EW + 2) * 32, 1)
lstm_ptable_kernel(const __grid_constant__ CUtensorMap tmap_w2d, const __grid_constant__ LstmP2Params P, int kb_first) {
  constexpr int NSUB = EW / 4;          // unit groups per pass
  constexpr int UPP = 16 / NSUB;        // units per thread and pass
  constexpr int UPT = 2 * UPP;          // units per thread
  constexpr int W_MMA = EW, W_XCH = EW + 1;
  constexpr int SUB_BYTES = ROWS * 32;  // one sub-slice: [ROWS rows x 16 units] fp16, SWIZZLE_32B K-major
  constexpr int LIVE_EW = EW * ROWS / 128;   // epilogue warps whose TMEM lane quarter holds batch rows
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KBh = P.H / KBLK;
  const int CL = P.H / 32;
  const uint32_t rank = cluster_ctarank();
  const int row0 = (blockIdx.x / CL) * ROWS;
  // per-tile pad-prefix start (tok_prep.cu): rows are sorted by descending number of leading PADs, so the tile's LAST row
  // has the shortest prefix; the tile starts at t0 from the tabulated state after t0 PADs
  int t0 = P.t_start;
  const float* init_h = P.init_h;
  const float* init_c = P.init_c;
  if (P.lead_sorted) {
    t0 = min(__ldg(P.lead_sorted + min(row0 + ROWS - 1, P.B - 1)), P.T - 1);
    init_h = t0 > 0 ? P.pad_h + (size_t)(t0 - 1) * P.H : nullptr;
    init_c = t0 > 0 ? P.pad_c + (size_t)(t0 - 1) * P.H : nullptr;
  }
  const bool has_init = init_h != nullptr;
  const int nsteps = P.T - t0;
  long long* trace = (P.dbg && blockIdx.x == 0) ? P.dbg + (size_t)gridDim.x * 16 : nullptr;

  // h tile = 2*CL sub-slices of 16 hidden units: sub-slice m = units [16 m, 16 m + 16) = CTA m/2's pass m%2,
  // [128 rows x 32 B] fp16, SWIZZLE_32B K-major -- exactly the A operand of k-step m.
  uint8_t* h_smem = smem;                                            // [2 tiles][2 CL sub-slices]
  uint8_t* wh_smem = h_smem + (size_t)2 * 2 * CL * SUB_BYTES;     // [KBh] tiles [128 x 64] SW128
  uint64_t* bars = reinterpret_cast<uint64_t*>(wh_smem + (size_t)KBh * W_TILE_BYTES);
  const uint32_t bar_wf = smem_u32(bars + 0);
  const uint32_t bar_accf = smem_u32(bars + 1);    // [2 accumulators]
  const uint32_t bar_hr = smem_u32(bars + 3);      // [2 tiles][2 passes]: all CL sub-slices of that pass have landed
  const uint32_t bar_sl = smem_u32(bars + 7);      // [2 tiles][2 passes]: own sub-slice written by the EW epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  if (threadIdx.x == 0) {
    mbar_init(bar_wf, 1);
    for (int b = 0; b < 2; ++b) mbar_init(bar_accf + 8 * b, 1);
    for (int b = 0; b < 4; ++b) { mbar_init(bar_hr + 8 * b, 2); mbar_init(bar_sl + 8 * b, LIVE_EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_MMA) {
    // ===== W_h slice load (once) + MMA issuer =====
    if (elect_one_sync()) {
      mbar_expect_tx(bar_wf, (uint32_t)KBh * W_TILE_BYTES);
      for (int kb = 0; kb < KBh; ++kb) tma_load_2d(smem_u32(wh_smem + (size_t)kb * W_TILE_BYTES), &tmap_w2d, bar_wf, (kb_first + kb) * KBLK, (int)rank * 128);
    }
    __syncwarp();
    const uint32_t idesc = make_idesc_f16(128, 128);
    const uint64_t whd = make_sw128_desc(smem_u32(wh_smem)), hd = make_sw32_desc(smem_u32(h_smem));
    const uint32_t wh_lo = (uint32_t)whd, wh_hi = (uint32_t)(whd >> 32), h_lo0 = (uint32_t)hd, h_hi = (uint32_t)(hd >> 32);
    long long w_hr = 0, t_begin = clock64();
    mbar_wait<false>(bar_wf, 0);
    for (int s = has_init ? 0 : 1; s < nsteps; ++s) {
      const int hb = (s - 1) & 1;
      const uint32_t n = s == 0 ? 0u : (uint32_t)(((s - 1) >> 1) + ((hb == 1 && has_init) ? 1 : 0));
      const uint32_t d = tmem_base + (uint32_t)((s & 1) * 128);
      const uint32_t h_lo = h_lo0 + (uint32_t)(hb * 2 * CL * (SUB_BYTES >> 4));
      for (int p = 0; p < 2; ++p) {
        // pass-p sub-slices of h_{s-1}: the k-steps m = 2 q + p can start while the peers still compute / send pass 1
        const uint32_t bar = bar_hr + 8 * (hb * 2 + p);
        if (elect_one_sync()) {
          if (s == 0) mbar_arrive(bar);
          else mbar_expect_tx(bar, (uint32_t)(CL - 1) * SUB_BYTES);
        }
        __syncwarp();
        mbar_wait_timed<false>(bar, n & 1, w_hr, (uint32_t)P.poll_ns);
        tc_fence_after();
        if (trace && lane == 0 && s < P2_TRACE_STEPS) trace[s * 8 + 4 + p] = clock64();
        if (elect_one_sync()) {
          for (int q = 0; q < CL; ++q) {
            const int m = 2 * q + p;
            tc_mma_ss3(d, h_lo + (uint32_t)(m * (SUB_BYTES >> 4)), h_hi, wh_lo + (uint32_t)((m >> 2) * (W_TILE_BYTES >> 4) + 2 * (m & 3)), wh_hi, idesc,
                       (p | q) ? 1u : 0u);
          }
          if (p == 1) tc_commit(bar_accf + 8 * (s & 1));
        }
        __syncwarp();
      }
    }
    if (P.dbg && lane == 0) { P.dbg[blockIdx.x * 16 + 0] = w_hr; P.dbg[blockIdx.x * 16 + 3] = clock64() - t_begin; }
  } else if (warp == W_XCH) {
    // ===== sub-slice exchange =====
    const uint32_t peer_h = map_to_cta(smem_u32(h_smem), (uint32_t)(lane < CL ? lane : 0));
    const uint32_t peer_bar = map_to_cta(bar_hr, (uint32_t)(lane < CL ? lane : 0));
    if (has_init) {
      for (int p = 0; p < 2; ++p) {
        mbar_wait<false>(bar_sl + 8 * (2 + p), 0);
        if (lane == 0) mbar_arrive(bar_hr + 8 * (2 + p));
      }
    }
    for (int s = 0; s + 1 < nsteps; ++s) {
      const int tb = s & 1;
      const uint32_t n = (uint32_t)((s >> 1) + ((tb == 1 && has_init) ? 1 : 0));
      for (int p = 0; p < 2; ++p) {
        mbar_wait<false>(bar_sl + 8 * (tb * 2 + p), n & 1, (uint32_t)P.poll_ns);
        const uint32_t off = (uint32_t)((tb * 2 * CL + 2 * (int)rank + p) * SUB_BYTES);
        if (lane < CL && lane != (int)rank) bulk_copy_to_peer(peer_h + off, smem_u32(h_smem) + off, SUB_BYTES, peer_bar + 8 * (tb * 2 + p));
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_hr + 8 * (tb * 2 + p));
        if (trace && lane == 0 && s < P2_TRACE_STEPS) trace[s * 8 + 6 + p] = clock64();
      }
    }
  } else if ((warp & 3) < ROWS / 32) {
    // ===== epilogue: thread == (batch row, UPT of this CTA's 32 hidden units), c in registers.
    //       pass p of unit group `sub` covers units 16 p + UPP sub + [0, UPP) of the slice, so that the NSUB groups of
    //       pass p together complete sub-slice 2 rank + p, which leaves for the peers while pass 1 is still computing
    const int quarter = warp & 3, sub = warp >> 2;
    const int r = quarter * 32 + lane;
    const int grow = row0 + r;
    const bool valid = grow < P.B;
    const int32_t* trow = P.tokens + (size_t)min(grow, P.B - 1) * P.T + t0;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t row_off32 = (uint32_t)(r * 32);
    const uint32_t sw32 = (uint32_t)((r >> 2) & 1);            // SWIZZLE_32B: 16-byte chunk c of row r sits at c ^ ((r >> 2) & 1)
    float c[UPT];
#pragma unroll
    for (int u = 0; u < UPT; ++u) c[u] = has_init ? __ldg(init_c + rank * 32 + (u / UPP) * 16 + sub * UPP + (u % UPP)) : 0.f;
    if (has_init) {
      // the broadcast initial state is the same for every row: each CTA fills its own tile 1 (= "step -1"), all sub-slices
      const uint32_t t1 = smem_u32(h_smem) + (uint32_t)(2 * CL * SUB_BYTES) + row_off32;
      for (int m = sub; m < 2 * CL; m += NSUB)
        for (int j = 0; j < 2; ++j) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) pk[e] = pack_f16x2(__ldg(init_h + m * 16 + j * 8 + 2 * e), __ldg(init_h + m * 16 + j * 8 + 2 * e + 1));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t1 + (uint32_t)m * SUB_BYTES + (((uint32_t)j ^ sw32) * 16)),
                       "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_sl + 8 * 2); mbar_arrive(bar_sl + 8 * 3); }
    }
    // byte offset of this thread's UPP units inside the 32-byte row of a sub-slice (16-byte chunks are XOR-swizzled)
    const uint32_t ubyte = (uint32_t)(sub * UPP * 2);
    const uint32_t own_sub = smem_u32(h_smem) + (2 * rank) * SUB_BYTES + row_off32 + (((ubyte >> 4) ^ sw32) * 16) + (ubyte & 15);
    const size_t pcol = (size_t)rank * 128 + sub * 4 * UPP;
    constexpr float NL2E = -1.4426950408889634f;
    long long w_accf = 0, w_math = 0, tq = 0, t_begin = clock64();
    const bool dbgt = P.dbg != nullptr;
    // table values of step s: pg[p][g * UPP + i] = P[token][gate g, unit 16 p + UPP sub + i].  Each pass's run of the NEXT
    // step's row is fetched right after the corresponding pass of this step has been published: it lands in registers
    // that just died, and no global load is in flight when the next publish executes its memory barrier
    // (fence.proxy.async lowers to MEMBAR.ALL.CTA, which would otherwise wait out the L2 latency on the critical path).
    float pg[2][4 * UPP];
    int tok_next = nsteps > 1 ? __ldg(trow + 1) : 0;
    {
      const float* prow = P.ptable + (size_t)__ldg(trow) * 4 * P.H + pcol;
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int x = 0; x < 4 * UPP; x += 8) ldg_v8(prow + p * 64 + x, pg[p] + x);
    }
    for (int s = 0; s < nsteps; ++s) {
      const bool last = s == nsteps - 1;
      const bool has_state = s > 0 || has_init;
      if (has_state) {
        mbar_wait_timed<false>(bar_accf + 8 * (s & 1), (uint32_t)(((s - (has_init ? 0 : 1)) >> 1) & 1), w_accf);
        tc_fence_after();
      }
      if (dbgt) tq = clock64();
      if (trace && warp == 0 && lane == 0 && s < P2_TRACE_STEPS) trace[s * 8 + 0] = tq;
      const uint32_t tile_sub = own_sub + (uint32_t)((s & 1) * 2 * CL * SUB_BYTES);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t v[4][UPP];
        if (has_state) {
          const uint32_t acc = lane_base + (uint32_t)((s & 1) * 128 + p * 16 + sub * UPP);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (UPP == 8) { TMEM_LD_8(acc + 32 * g, v[g]); }
            else { TMEM_LD_4(acc + 32 * g, v[g]); }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < UPP; ++i) v[g][i] = 0u;
        }
        float hv[UPP];
#pragma unroll
        for (int i = 0; i < UPP; ++i) {
          const int u = p * UPP + i;
          const float bi = pg[p][i], bj = pg[p][UPP + i], bf = pg[p][2 * UPP + i], bo = pg[p][3 * UPP + i];
          const float ai = __uint_as_float(v[0][i]), aj = __uint_as_float(v[1][i]), af = __uint_as_float(v[2][i]), ao = __uint_as_float(v[3][i]);
          if (GATE == 0) {
            const float ei = ex2_approx(fminf(fmaf(ai, NL2E, bi), 57.f));
            const float ej = ex2_approx(fminf(fmaf(aj, 2.f * NL2E, bj), 57.f));
            const float ef = ex2_approx(fminf(fmaf(af, NL2E, bf), 57.f));
            const float eo = ex2_approx(fminf(fmaf(ao, NL2E, bo), 57.f));
            const float pj = (1.f - ej) * rcp_approx((1.f + ei) * (1.f + ej));
            c[u] = fmaf(c[u], rcp_approx(1.f + ef), pj);
            const float ec = ex2_approx(fminf(c[u] * (2.f * NL2E), 57.f));
            hv[i] = (1.f - ec) * rcp_approx((1.f + ec) * (1.f + eo));
          } else {
            // table scaled by 1/2 (1 for the candidate gate): sigmoid(z) = 0.5 tanh(z / 2) + 0.5
            const float ti = tanh_approx(fmaf(ai, 0.5f, bi));
            const float tj = tanh_approx(aj + bj);
            const float tf = tanh_approx(fmaf(af, 0.5f, bf));
            c[u] = fmaf(c[u], fmaf(0.5f, tf, 0.5f), fmaf(0.5f, ti, 0.5f) * tj);
            if (GATE == 1) {
              const float to = tanh_approx(fmaf(ao, 0.5f, bo));
              hv[i] = fmaf(0.5f, to, 0.5f) * tanh_approx(c[u]);
            } else {
              const uint32_t t2 = tanh_f16x2(pack_f16x2(fmaf(ao, 0.5f, bo), c[u]));       // low half: o gate, high half: cell
              const float2 tt = __half22float2(*reinterpret_cast<const __half2*>(&t2));
              hv[i] = fmaf(0.5f, tt.x, 0.5f) * tt.y;
            }
          }
        }
        if (P.dump_h && grow == 0) {        // pad-prefix table generation: the state this kernel itself carries past step t0 + s
          float* dh = P.dump_h + (size_t)(t0 + s) * P.H + rank * 32 + p * 16 + sub * UPP;
          float* dc = P.dump_c + (size_t)(t0 + s) * P.H + rank * 32 + p * 16 + sub * UPP;
#pragma unroll
          for (int i = 0; i < UPP; ++i) { dh[i] = __half2float(__float2half_rn(hv[i])); dc[i] = c[p * UPP + i]; }
        }
        if (!last) {
          if (UPP == 8) {
            const uint32_t p0 = pack_f16x2(hv[0], hv[1]), p1 = pack_f16x2(hv[2], hv[3]), p2 = pack_f16x2(hv[4 % UPP], hv[5 % UPP]), p3 = pack_f16x2(hv[6 % UPP], hv[7 % UPP]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_sub + (uint32_t)(p * SUB_BYTES)), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
          } else {
            const uint32_t p0 = pack_f16x2(hv[0], hv[1]), p1 = pack_f16x2(hv[2], hv[3]);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(tile_sub + (uint32_t)(p * SUB_BYTES)), "r"(p0), "r"(p1) : "memory");
          }
          // this pass's part of the sub-slice is in place (and, for p == 1, the accumulator fully read): publish it
          if (p == 1) tc_fence_before();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_sl + 8 * ((s & 1) * 2 + p));
          if (trace && warp == 0 && lane == 0 && s < P2_TRACE_STEPS) trace[s * 8 + 1 + p] = clock64();
          {
            const float* prow = P.ptable + (size_t)tok_next * 4 * P.H + pcol + p * 64;
#pragma unroll
            for (int x = 0; x < 4 * UPP; x += 8) ldg_v8(prow + x, pg[p] + x);
          }
        } else if (valid) {
          float* ho = P.h_out + (size_t)grow * P.H + rank * 32 + p * 16 + sub * UPP;
#pragma unroll
          for (int i = 0; i < UPP; i += 4) *reinterpret_cast<float4*>(ho + i) = make_float4(hv[i], hv[i + 1], hv[i + 2], hv[i + 3]);
        }
      }
      if (s + 2 < nsteps) tok_next = __ldg(trow + s + 2);
      if (dbgt) w_math += clock64() - tq;
    }
    if (P.dbg && lane == 0 && warp == 0) {
      long long* o = P.dbg + blockIdx.x * 16;
      o[4] = w_accf; o[5] = clock64() - t_begin; o[7] = w_math;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

// =====================================================================================================
// Variant 3 ("wide"): 64 hidden units per CTA, clusters of H/64 CTAs.
// Variant 2's step is bound by the h exchange, not by the tensor or MUFU pipes: with 8 CTAs x 32 units every CTA pushes
// 7 x 8 KB = 56 KB of slices per step through distributed shared memory, which sustains ~17 B/cycle per SM (measured with
// the timeline trace: 28 KB per pass land 1.65-2.0 k cycles after the copies were issued).  The bytes a CTA sends per
// (row, unit) it computes are 2 (CL - 1), so halving the cluster halves them: here a CTA owns 64 units (256 gate columns,
// N = 256 MMAs, W_h slice 128 KB resident), computes twice the work per step and sends 3 x 16 KB = 48 KB.
// What makes it fit: ONE h operand tile (64 KB) instead of a ping-pong pair.  A peer may overwrite this CTA's tile with
// h_s slices only after this CTA's MMAs of step s (which read h_{s-1}) have completed; that is published explicitly: when
// the epilogue sees the accumulator of step s it arrives on the `free` barrier of every CTA of the cluster, and an exchange
// warp waits for all CL arrivals of step s before it sends the first slice of h_s.  The signal is sent ~a pass before the
// first slice is ready, so it costs nothing in steady state.
// Four passes of 16 units per step (thread == batch row x 4 units per pass, 16 epilogue warps): pass p's sub-slice leaves
// while passes p+1.. compute, and the k-steps of pass p of the next step start when all CL pass-p sub-slices have landed.
// Table values are fetched two passes ahead into a two-slot register ring.
// shared memory (H=256): h tile 64 KB | W_h slice 128 KB.  TMEM: two [128 x 256] fp32 accumulators (512 columns).
// warps: 0-15 epilogue | 16 MMA issuer + TMEM alloc + weight load | 17 slice exchange
// The same kernel also runs with 32 units per CTA (clusters of H/32, N = 128): the single h tile then leaves room to
// STAGE the table rows in shared memory.  Why: the per-thread gathers (thread == batch row, every lane another table row)
// cost ~47 L1 data-pipe wavefronts per LDG.256 -- 64 (128 with 64 units) such loads per step kept the LSU pipe busy for
// ~3 k (~6 k) cycles per step, more than the MUFU work.  With STAGE a producer warp copies each row's run with coalesced
// 16-byte cp.async (two whole rows per warp instruction) into a swizzled [128 rows x 256 B] region per pass, one step ahead;
// the epilogue threads read their 64 bytes with four conflict-free LDS.128.
constexpr int P3_EW = 16, P3_UPP = 4;
constexpr int P3_DEFAULT_UNITS = 32;
constexpr int P3_TAB_REGION = P2_ROWS * 256;   // one pass of staged table rows: [128 rows][16 chunks of 16 B], chunk ^ (row & 15)

// column of (gate g, hidden unit `unit`) in the table of this kernel: the 16 floats a thread needs in a pass are contiguous,
// the 64 a batch row needs in a pass (all four unit groups) too
__host__ __device__ __forceinline__ int ptab_col_wide(int unit, int g, int units_per_cta) {
  const int c = unit / units_per_cta, j = unit % units_per_cta;
  return c * 4 * units_per_cta + (j >> 4) * 64 + ((j & 15) >> 2) * 16 + g * 4 + (j & 3);
}

template <int GATE, int UNITS, bool STAGE>
__global__ void __launch_bounds__((P3_EW + 2 + (STAGE ? 1 : 0)) * 32, 1)
lstm_wide_kernel(const __grid_constant__ CUtensorMap tmap_w2d, const __grid_constant__ LstmP2Params P, int kb_first) {
  constexpr int EW = P3_EW, NPASS = UNITS / 16, UPP = P3_UPP, UPT = NPASS * UPP;
  constexpr int NCOL = 4 * UNITS;                   // gate columns of the slice = MMA N = accumulator columns
  constexpr int WK_BYTES = NCOL * KBLK * 2;         // one k-block of the slice: [NCOL gate rows x 64 k] fp16, SW128
  constexpr int W_MMA = EW, W_XCH = EW + 1, W_TAB = EW + 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KBh = P.H / KBLK;
  const int CL = P.H / UNITS;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const int row0 = (blockIdx.x / CL) * P2_ROWS;
  int t0 = P.t_start;
  const float* init_h = P.init_h;
  const float* init_c = P.init_c;
  if (P.lead_sorted) {
    t0 = min(__ldg(P.lead_sorted + min(row0 + P2_ROWS - 1, P.B - 1)), P.T - 1);
    init_h = t0 > 0 ? P.pad_h + (size_t)(t0 - 1) * P.H : nullptr;
    init_c = t0 > 0 ? P.pad_c + (size_t)(t0 - 1) * P.H : nullptr;
  }
  const bool has_init = init_h != nullptr;
  const int hi = has_init ? 1 : 0;
  const int nsteps = P.T - t0;
  long long* trace = (P.dbg && blockIdx.x == 0) ? P.dbg + (size_t)gridDim.x * 16 : nullptr;

  // h tile = H/16 sub-slices of 16 hidden units: sub-slice m = units [16 m, 16 m + 16) = CTA m/NPASS's pass m%NPASS,
  // [128 rows x 32 B] fp16, SWIZZLE_32B K-major -- exactly the A operand of k-step m.
  uint8_t* h_smem = smem;
  uint8_t* wh_smem = h_smem + (size_t)(P.H / 16) * P2_SUB_BYTES;     // [KBh] tiles [NCOL x 64] SW128
  uint8_t* tab_smem = wh_smem + (size_t)KBh * WK_BYTES;              // STAGE: [NPASS] regions
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab_smem + (STAGE ? (size_t)NPASS * P3_TAB_REGION : 0));
  const uint32_t bar_wf = smem_u32(bars + 0);
  const uint32_t bar_accf = smem_u32(bars + 1);    // [2 accumulators]
  const uint32_t bar_hr = smem_u32(bars + 3);      // [NPASS]: all CL sub-slices of that pass have landed
  const uint32_t bar_sl = smem_u32(bars + 7);      // [NPASS]: own sub-slice written by the EW epilogue warps
  const uint32_t bar_free = smem_u32(bars + 11);   // every CTA of the cluster has finished reading the previous h from its tile
  const uint32_t bar_tfull = smem_u32(bars + 12);  // [NPASS] STAGE: the region holds the table rows of the coming step
  const uint32_t bar_tempty = smem_u32(bars + 16); // [NPASS] STAGE: all epilogue warps have taken their values out of the region
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  if (threadIdx.x == 0) {
    mbar_init(bar_wf, 1);
    for (int b = 0; b < 2; ++b) mbar_init(bar_accf + 8 * b, 1);
    for (int b = 0; b < NPASS; ++b) {
      mbar_init(bar_hr + 8 * b, 2); mbar_init(bar_sl + 8 * b, EW);
      mbar_init(bar_tfull + 8 * b, 32); mbar_init(bar_tempty + 8 * b, EW);
    }
    mbar_init(bar_free, (uint32_t)CL);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * NCOL));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CL > 1) cluster_sync_all();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_MMA) {
    // ===== W_h slice load (once) + MMA issuer =====
    if (elect_one_sync()) {
      mbar_expect_tx(bar_wf, (uint32_t)KBh * WK_BYTES);
      for (int kb = 0; kb < KBh; ++kb)
        for (int hf = 0; hf < NCOL / 128; ++hf)
          tma_load_2d(smem_u32(wh_smem + (size_t)kb * WK_BYTES + (size_t)hf * W_TILE_BYTES), &tmap_w2d, bar_wf, (kb_first + kb) * KBLK, (int)rank * NCOL + hf * 128);
    }
    __syncwarp();
    const uint32_t idesc = make_idesc_f16(128, NCOL);
    const uint64_t whd = make_sw128_desc(smem_u32(wh_smem)), hd = make_sw32_desc(smem_u32(h_smem));
    const uint32_t wh_lo = (uint32_t)whd, wh_hi = (uint32_t)(whd >> 32), h_lo = (uint32_t)hd, h_hi = (uint32_t)(hd >> 32);
    long long w_hr = 0, t_begin = clock64();
    mbar_wait<false>(bar_wf, 0);
    for (int s = has_init ? 0 : 1; s < nsteps; ++s) {
      const uint32_t n = (uint32_t)(s - 1 + hi);            // completion index of h_{s-1} (the initial state is index 0)
      const uint32_t d = tmem_base + (uint32_t)((s & 1) * NCOL);
      for (int p = 0; p < NPASS; ++p) {
        const uint32_t bar = bar_hr + 8 * p;
        if (elect_one_sync()) {
          if (s == 0 || CL == 1) mbar_arrive(bar);
          else mbar_expect_tx(bar, (uint32_t)(CL - 1) * P2_SUB_BYTES);
        }
        __syncwarp();
        mbar_wait_timed<false>(bar, n & 1, w_hr, (uint32_t)P.poll_ns);
        tc_fence_after();
        if (trace && lane == 0 && s < P2_TRACE_STEPS && p == 0) trace[s * 8 + 4] = clock64();
        if (trace && lane == 0 && s < P2_TRACE_STEPS && p == NPASS - 1) trace[s * 8 + 5] = clock64();
        if (elect_one_sync()) {
          for (int q = 0; q < CL; ++q) {
            const int m = NPASS * q + p;
            tc_mma_ss3(d, h_lo + (uint32_t)(m * (P2_SUB_BYTES >> 4)), h_hi, wh_lo + (uint32_t)((m >> 2) * (WK_BYTES >> 4) + 2 * (m & 3)), wh_hi, idesc,
                       (p | q) ? 1u : 0u);
          }
          if (p == NPASS - 1) tc_commit(bar_accf + 8 * (s & 1));
        }
        __syncwarp();
      }
    }
    if (P.dbg && lane == 0) { P.dbg[blockIdx.x * 16 + 0] = w_hr; P.dbg[blockIdx.x * 16 + 3] = clock64() - t_begin; }
  } else if (warp == W_XCH) {
    // ===== sub-slice exchange =====
    const uint32_t peer_h = map_to_cta(smem_u32(h_smem), (uint32_t)(lane < CL ? lane : 0));
    const uint32_t peer_bar = map_to_cta(bar_hr, (uint32_t)(lane < CL ? lane : 0));
    if (has_init) {
      for (int p = 0; p < NPASS; ++p) {
        mbar_wait<false>(bar_sl + 8 * p, 0);
        if (lane == 0) mbar_arrive(bar_hr + 8 * p);
      }
    }
    for (int s = 0; s + 1 < nsteps; ++s) {
      const uint32_t n = (uint32_t)(s + hi);
      for (int p = 0; p < NPASS; ++p) {
        mbar_wait<false>(bar_sl + 8 * p, n & 1, (uint32_t)P.poll_ns);
        if (p == 0 && CL > 1) mbar_wait<true>(bar_free, (uint32_t)s & 1, (uint32_t)P.poll_ns);     // every peer's tile may be overwritten now
        const uint32_t off = (uint32_t)((NPASS * (int)rank + p) * P2_SUB_BYTES);
        if (lane < CL && lane != (int)rank) bulk_copy_to_peer(peer_h + off, smem_u32(h_smem) + off, P2_SUB_BYTES, peer_bar + 8 * p);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_hr + 8 * p);
        if (trace && lane == 0 && s < P2_TRACE_STEPS && p == 0) trace[s * 8 + 6] = clock64();
        if (trace && lane == 0 && s < P2_TRACE_STEPS && p == NPASS - 1) trace[s * 8 + 7] = clock64();
      }
    }
  } else if (STAGE && warp == W_TAB) {
    // ===== table-row producer: region p of step s is refilled as soon as every epilogue warp has taken step s-1's values out
    const int chunk = lane & 15, rsel = lane >> 4;
    for (int s = 0; s < nsteps; ++s) {
      for (int p = 0; p < NPASS; ++p) {
        if (s > 0) mbar_wait<false>(bar_tempty + 8 * p, (uint32_t)(s - 1) & 1, 64);
        if ((P.flags & 4) && s > 0 && p == 0)      // experiment: fetch only once h_{s-1} has completely landed here (the previous step's exchange is over)
          mbar_wait<false>(bar_hr + 8 * (NPASS - 1), (uint32_t)(s - 1 + hi) & 1, 64);
        const uint32_t reg_base = smem_u32(tab_smem) + (uint32_t)(p * P3_TAB_REGION);
        const float* src0 = P.ptable + (size_t)rank * NCOL + p * 64 + chunk * 4;
        if (!(P.flags & 1)) {
#pragma unroll 8
          for (int i = 0; i < P2_ROWS / 2; ++i) {
            const int row = 2 * i + rsel;
            const int tok = __ldg(P.tokens + (size_t)min(row0 + row, P.B - 1) * P.T + t0 + s);
            const uint32_t dst = reg_base + (uint32_t)(row * 256 + ((chunk ^ (row & 15)) << 4));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src0 + (size_t)tok * 4 * P.H) : "memory");
            if ((P.flags & 2) && (i & 7) == 7) __nanosleep(100);
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_tfull + 8 * p) : "memory");
      }
    }
  } else {
    // ===== epilogue: thread == (batch row, UPT of this CTA's UNITS hidden units: 4 per pass), c in registers
    const int quarter = warp & 3, sub = warp >> 2;
    const int r = quarter * 32 + lane;
    const int grow = row0 + r;
    const bool valid = grow < P.B;
    const int32_t* trow = P.tokens + (size_t)min(grow, P.B - 1) * P.T + t0;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t row_off32 = (uint32_t)(r * 32);
    const uint32_t sw32 = (uint32_t)((r >> 2) & 1);            // SWIZZLE_32B: 16-byte chunk c of row r sits at c ^ ((r >> 2) & 1)
    const int ubase = (int)rank * UNITS + sub * UPP;           // + 16 p + i
    float c[UPT];
#pragma unroll
    for (int u = 0; u < UPT; ++u) c[u] = has_init ? __ldg(init_c + ubase + (u / UPP) * 16 + (u % UPP)) : 0.f;
    if (has_init) {
      // the broadcast initial state is the same for every row: each CTA fills its own tile, all sub-slices
      const uint32_t t1 = smem_u32(h_smem) + row_off32;
      for (int m = sub; m < P.H / 16; m += EW / 4)
        for (int j = 0; j < 2; ++j) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) pk[e] = pack_f16x2(__ldg(init_h + m * 16 + j * 8 + 2 * e), __ldg(init_h + m * 16 + j * 8 + 2 * e + 1));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t1 + (uint32_t)m * P2_SUB_BYTES + (((uint32_t)j ^ sw32) * 16)),
                       "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0)
        for (int p = 0; p < NPASS; ++p) mbar_arrive(bar_sl + 8 * p);
    }
    const uint32_t ubyte = (uint32_t)(sub * UPP * 2);
    const uint32_t own_sub = smem_u32(h_smem) + (uint32_t)(NPASS * (int)rank) * P2_SUB_BYTES + row_off32 + (((ubyte >> 4) ^ sw32) * 16) + (ubyte & 15);
    const size_t pcol = (size_t)rank * NCOL + sub * 16;        // + 64 p: the pass's run of 16 floats (4 gates x 4 units)
    const uint32_t tab_row = smem_u32(tab_smem) + (uint32_t)(r * 256);
    // remote `free` barriers (one elected thread of the CTA signals all CL CTAs, itself included)
    const bool signaller = warp == 0 && lane < CL && CL > 1;
    const uint32_t free_remote = signaller ? map_to_cta(bar_free, (uint32_t)lane) : 0u;
    long long w_accf = 0, w_math = 0, tq = 0, t_begin = clock64();
    const bool dbgt = P.dbg != nullptr;
    // !STAGE: table values come straight from global memory into a two-slot register ring, slot p & 1 holds pass p's run;
    // it is refilled for the pass two positions later right after pass p has been published
    float pg[2][4 * UPP];
    int tok_cur = __ldg(trow), tok_next = nsteps > 1 ? __ldg(trow + 1) : 0;
    if (!STAGE) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float* prow = P.ptable + (size_t)(q < NPASS ? tok_cur : tok_next) * 4 * P.H + pcol + (q % NPASS) * 64;
#pragma unroll
        for (int x = 0; x < 4 * UPP; x += 8) ldg_v8(prow + x, pg[q] + x);
      }
    }
    for (int s = 0; s < nsteps; ++s) {
      const bool last = s == nsteps - 1;
      const bool has_state = s > 0 || has_init;
      if (has_state) {
        mbar_wait_timed<false>(bar_accf + 8 * (s & 1), (uint32_t)(((s - (has_init ? 0 : 1)) >> 1) & 1), w_accf);
        tc_fence_after();
      }
      // this CTA's MMAs of step s are complete: its h tile may now be overwritten with h_s
      if (signaller && !last) mbar_arrive_remote(free_remote);
      if (dbgt) tq = clock64();
      if (trace && warp == 0 && lane == 0 && s < P2_TRACE_STEPS) trace[s * 8 + 0] = tq;
#pragma unroll
      for (int p = 0; p < NPASS; ++p) {
        float* pgp = pg[p & 1];
        if (STAGE) {
          mbar_wait<false>(bar_tfull + 8 * p, (uint32_t)s & 1);
#pragma unroll
          for (int g = 0; g < 4; ++g) lds_v4(tab_row + (uint32_t)(p * P3_TAB_REGION + ((((sub << 2) + g) ^ (r & 15)) << 4)), pgp + g * UPP);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * p);
        }
        uint32_t v[4][UPP];
        if (has_state) {
          const uint32_t acc = lane_base + (uint32_t)((s & 1) * NCOL + (p >> 1) * 128 + (p & 1) * 16 + sub * UPP);
#pragma unroll
          for (int g = 0; g < 4; ++g) TMEM_LD_4(acc + 32 * g, v[g]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < UPP; ++i) v[g][i] = 0u;
        }
        float hv[UPP];
#pragma unroll
        for (int i = 0; i < UPP; ++i) {
          const int u = p * UPP + i;
          const float bi = pgp[i], bj = pgp[UPP + i], bf = pgp[2 * UPP + i], bo = pgp[3 * UPP + i];
          const float ai = __uint_as_float(v[0][i]), aj = __uint_as_float(v[1][i]), af = __uint_as_float(v[2][i]), ao = __uint_as_float(v[3][i]);
          if (GATE == 0) {
            constexpr float NL2E = -1.4426950408889634f;
            const float ei = ex2_approx(fminf(fmaf(ai, NL2E, bi), 57.f));
            const float ej = ex2_approx(fminf(fmaf(aj, 2.f * NL2E, bj), 57.f));
            const float ef = ex2_approx(fminf(fmaf(af, NL2E, bf), 57.f));
            const float eo = ex2_approx(fminf(fmaf(ao, NL2E, bo), 57.f));
            const float pj = (1.f - ej) * rcp_approx((1.f + ei) * (1.f + ej));
            c[u] = fmaf(c[u], rcp_approx(1.f + ef), pj);
            const float ec = ex2_approx(fminf(c[u] * (2.f * NL2E), 57.f));
            hv[i] = (1.f - ec) * rcp_approx((1.f + ec) * (1.f + eo));
          } else {
            const float ti = tanh_approx(fmaf(ai, 0.5f, bi));
            const float tj = tanh_approx(aj + bj);
            const float tf = tanh_approx(fmaf(af, 0.5f, bf));
            c[u] = fmaf(c[u], fmaf(0.5f, tf, 0.5f), fmaf(0.5f, ti, 0.5f) * tj);
            if (GATE == 1) {
              const float to = tanh_approx(fmaf(ao, 0.5f, bo));
              hv[i] = fmaf(0.5f, to, 0.5f) * tanh_approx(c[u]);
            } else {
              const uint32_t t2 = tanh_f16x2(pack_f16x2(fmaf(ao, 0.5f, bo), c[u]));
              const float2 tt = __half22float2(*reinterpret_cast<const __half2*>(&t2));
              hv[i] = fmaf(0.5f, tt.x, 0.5f) * tt.y;
            }
          }
        }
        if (P.dump_h && grow == 0) {
          float* dh = P.dump_h + (size_t)(t0 + s) * P.H + ubase + p * 16;
          float* dc = P.dump_c + (size_t)(t0 + s) * P.H + ubase + p * 16;
#pragma unroll
          for (int i = 0; i < UPP; ++i) { dh[i] = __half2float(__float2half_rn(hv[i])); dc[i] = c[p * UPP + i]; }
        }
        if (!last) {
          const uint32_t p0 = pack_f16x2(hv[0], hv[1]), p1 = pack_f16x2(hv[2], hv[3]);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(own_sub + (uint32_t)(p * P2_SUB_BYTES)), "r"(p0), "r"(p1) : "memory");
          if (p == NPASS - 1) tc_fence_before();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_sl + 8 * p);
          if (trace && warp == 0 && lane == 0 && s < P2_TRACE_STEPS && p == 0) trace[s * 8 + 1] = clock64();
          if (trace && warp == 0 && lane == 0 && s < P2_TRACE_STEPS && p == NPASS - 1) trace[s * 8 + 2] = clock64();
        } else if (valid) {
          float* ho = P.h_out + (size_t)grow * P.H + ubase + p * 16;
          *reinterpret_cast<float4*>(ho) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        }
        if (!STAGE && (p + 2 < NPASS || !last)) {
          // refill this slot for the pass two positions later: of this step, or of the next one
          const int q = p + 2;
          const float* prow = P.ptable + (size_t)(q < NPASS ? tok_cur : tok_next) * 4 * P.H + pcol + (q % NPASS) * 64;
#pragma unroll
          for (int x = 0; x < 4 * UPP; x += 8) ldg_v8(prow + x, pgp + x);
        }
      }
      tok_cur = tok_next;
      if (s + 2 < nsteps) tok_next = __ldg(trow + s + 2);
      if (dbgt) w_math += clock64() - tq;
    }
    if (P.dbg && lane == 0 && warp == 0) {
      long long* o = P.dbg + blockIdx.x * 16;
      o[4] = w_accf; o[5] = clock64() - t_begin; o[7] = w_math;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * NCOL));
  }
}

// Wxp[k][col] = scale_g * K[k][g*H + 32c + j], col = ptab_col(c, g, j, upp)  (the table's column order), k < We
__global__ void ptable_wx_kernel(const float* __restrict__ K, int We, int H, int tanh_form, int upp, int wide, float* __restrict__ Wxp) {
  const int64_t total = (int64_t)We * 4 * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / (4 * H)), np = (int)(i - (int64_t)k * 4 * H);
    const int c = np >> 7, g = (np >> 5) & 3, j = np & 31;
    const float sc = tanh_form ? (g == 1 ? 1.f : 0.5f) : (g == 1 ? -2.885390081777927f : -1.4426950408889634f);
    Wxp[(size_t)k * 4 * H + (wide ? ptab_col_wide(c * 32 + j, g, wide) : ptab_col(c, g, j, upp))] = sc * K[(size_t)k * 4 * H + g * H + c * 32 + j];
  }
}
// P[v][col] = bias_r[c*128 + g*32 + j]  (the GEMM then accumulates the projection on top)
__global__ void ptable_bias_kernel(const float* __restrict__ bias_r, int64_t V, int H4, int tanh_form, int upp, int wide, float* __restrict__ Pt) {
  const int64_t total = V * H4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int np = (int)(i % H4);
    const int c = np >> 7, g = (np >> 5) & 3, j = np & 31;
    float b = bias_r[np];                       // -log2e (-2 log2e for g == 1) x (b [+1 forget])
    if (tanh_form) b *= -0.34657359027997264f;  // (-log2e)(b) -> b/2 for i,f,o and (-2 log2e)(b) -> b for j: the same factor -ln2/2
    Pt[i - np + (wide ? ptab_col_wide(c * 32 + j, g, wide) : ptab_col(c, g, j, upp))] = b;
  }
}

}  // namespace

// H = 64 / 128 / 256 (cluster of 2 / 4 / 8 CTAs), We a multiple of 64 up to 256
bool lstm_cluster_supported(int We, int H) { return We % 64 == 0 && We >= 64 && We <= 256 && (H == 64 || H == 128 || H == 256); }

int lstm_forward_cluster(const int32_t* tokens, int B, int T, int t_start, const __half* emb_f16, int We, int H,
                         const TcTower& tt, const float* init_h, const float* init_c, const PadSkip& ps, float* h_out, cudaStream_t st,
                         int64_t* launches) {
  if (B <= 0 || T - t_start <= 0) return SSE_OK;
  LstmClParams p;
  p.tokens = tokens; p.emb = emb_f16; p.bias_r = tt.bias_r; p.init_h = init_h; p.init_c = init_c; p.h_out = h_out;
  p.lead_sorted = ps.lead_sorted; p.pad_h = ps.pad_h; p.pad_c = ps.pad_c;
  p.B = B; p.T = T; p.t_start = t_start; p.We = We; p.H = H; p.dbg = nullptr;
  p.dbg_flags = getenv("SSE_LSTM_CL_FLAGS") ? atoi(getenv("SSE_LSTM_CL_FLAGS")) : 0;
  const int CL = H / 32, KBx = We / KBLK, KBh = H / KBLK;
  const int n_clusters = cdiv(B, CL_ROWS);
  const int grid = n_clusters * CL;
  const size_t smem = 1024 + (size_t)KBx * A_TILE_BYTES + (size_t)2 * KBh * A_TILE_BYTES + (size_t)(KBx + KBh) * W_TILE_BYTES + 512 + 256;
  const bool want_dbg = getenv("SSE_LSTM_DEBUG") != nullptr;
  long long* d_dbg = nullptr;
  if (want_dbg) { cudaMalloc(&d_dbg, (size_t)grid * 128); cudaMemset(d_dbg, 0, (size_t)grid * 128); p.dbg = d_dbg; }
  SSE_CUDA_OK(cudaFuncSetAttribute(lstm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(CL_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SSE_CUDA_OK(cudaLaunchKernelEx(&cfg, lstm_cluster_kernel, *reinterpret_cast<const CUtensorMap*>(tt.tmap2d), p));
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  if (want_dbg) {
    std::vector<long long> hd((size_t)grid * 16);
    cudaStreamSynchronize(st);
    cudaMemcpy(hd.data(), d_dbg, hd.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_dbg);
    const char* nm[12] = {"mma_wait_hready", "mma_wait_xfull", "mma_wait_acce", "mma_total", "epi_wait_accf", "epi_total",
                          "epi_tmem_ld", "epi_math", "epi_store", "epi_publish", "mma_issue_h", "mma_issue_x(+waits)"};
    for (int c = 0; c < 12; ++c) {
      long long sm = 0;
      for (int i = 0; i < grid; ++i) sm += hd[(size_t)i * 16 + c];
      fprintf(stderr, "[lstm cluster dbg] %-16s avg %10lld cycles  (grid %d = %d clusters x %d, steps %d)\n", nm[c], sm / grid, grid, n_clusters, CL, T - t_start);
    }
  }
  return SSE_OK;
}

// ---- variant 2 host side ------------------------------------------------------------------------------
bool lstm_ptable_supported(int64_t V, int We, int H) {
  (void)We;
  return (H == 64 || H == 128 || H == 256) && V * 4 * H * 4 <= ((int64_t)2 << 30);      // table <= 2 GiB
}

void lstm_ptable_release(TcTower& tt) {
  if (tt.ptable) cudaFree(tt.ptable);
  if (tt.wxp) cudaFree(tt.wxp);
  tt.wxp = nullptr;
  tt.ptable = nullptr; tt.ptable_valid = false; tt.ptable_rows = 0;
}

// P = scale * (Emb W_x + b), fp32 (one SIMT GEMM per weight set; the recurrent part keeps using tt.wt / tt.bias_r)
int lstm_ptable_prepare(TcTower& tt, const float* emb, int64_t V, const float* K, int We, int H, cudaStream_t st, int64_t* launches) {
  if (!tt.valid) { set_error("lstm_ptable_prepare: tower weights not prepared"); return SSE_ESTATE; }
  if (tt.ptable && tt.ptable_rows != V) { cudaFree(tt.ptable); tt.ptable = nullptr; }
  if (!tt.ptable) { SSE_CUDA_OK(cudaMalloc(&tt.ptable, (size_t)V * 4 * H * 4)); tt.ptable_rows = V; }
  if (!tt.wxp) SSE_CUDA_OK(cudaMalloc(&tt.wxp, (size_t)We * 4 * H * 4));
  float* wxp = tt.wxp;
  // kernel variant (fixed per table: it sets the table's scaling and column order): SSE_LSTM_GATE_MATH 0 / 1 / 2,
  // SSE_LSTM_EW 8 / 16 epilogue warps
  const int gate_math = getenv("SSE_LSTM_GATE_MATH") ? std::max(0, std::min(2, atoi(getenv("SSE_LSTM_GATE_MATH")))) : P2_DEFAULT_GATE;
  const int ew = getenv("SSE_LSTM_EW") ? (atoi(getenv("SSE_LSTM_EW")) == 8 ? 8 : 16) : P2_DEFAULT_EW;
  // SSE_LSTM_UNITS = 32: variant 2 (clusters of H/32 CTAs), 64: variant 3 (clusters of H/64 CTAs, half the exchange per unit)
  // tt.ptable_wide: 0 = variant 2 (lstm_ptable_kernel), 32 / 64 = variant 3 (lstm_wide_kernel) with that many units per CTA
  const int variant = getenv("SSE_LSTM_VARIANT") ? atoi(getenv("SSE_LSTM_VARIANT")) : 2;     // measured: variant 3 is slower in every configuration tried (see DESIGN)
  const int units = getenv("SSE_LSTM_UNITS") ? atoi(getenv("SSE_LSTM_UNITS")) : P3_DEFAULT_UNITS;
  const int wide = variant == 2 ? 0 : (units == 64 ? 64 : 32);
  tt.ptable_mode = gate_math;
  tt.ptable_ew = ew;
  tt.ptable_wide = wide;
  const int upp = 64 / ew;
  ptable_wx_kernel<<<148, 256, 0, st>>>(K, We, H, gate_math != 0, upp, wide, wxp);
  if (launches) ++*launches;
  ptable_bias_kernel<<<148 * 8, 256, 0, st>>>(tt.bias_r, V, 4 * H, gate_math != 0, upp, wide, tt.ptable);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  // rows in slabs that keep the GEMM's M within int range and friendly to the SIMT kernel
  for (int64_t v0 = 0; v0 < V; v0 += 65536) {
    const int nv = (int)std::min<int64_t>(65536, V - v0);
    SSE_TRY(sgemm(false, false, nv, 4 * H, We, 1.f, emb + (size_t)v0 * We, We, wxp, 4 * H, 1.f, tt.ptable + (size_t)v0 * 4 * H, 4 * H, st, launches));
  }
  tt.ptable_valid = true;
  return SSE_OK;
}

int lstm_forward_ptable(const int32_t* tokens, int B, int T, int t_start, int We, int H, const TcTower& tt, const float* init_h,
                        const float* init_c, const PadSkip& ps, float* h_out, cudaStream_t st, int64_t* launches, int cluster_rows, int num_sms) {
  if (B <= 0 || T - t_start <= 0) return SSE_OK;
  if (!tt.ptable_valid) { set_error("lstm_forward_ptable: table not prepared"); return SSE_ESTATE; }
  LstmP2Params p;
  p.tokens = tokens; p.ptable = tt.ptable; p.init_h = init_h; p.init_c = init_c; p.h_out = h_out;
  p.lead_sorted = ps.lead_sorted; p.pad_h = ps.pad_h; p.pad_c = ps.pad_c; p.dump_h = ps.dump_h; p.dump_c = ps.dump_c;
  p.B = B; p.T = T; p.t_start = t_start; p.H = H; p.dbg = nullptr;
  static const int poll_ns = getenv("SSE_LSTM_POLL") ? atoi(getenv("SSE_LSTM_POLL")) : 32;
  p.poll_ns = poll_ns;
  p.flags = getenv("SSE_LSTM_FLAGS") ? atoi(getenv("SSE_LSTM_FLAGS")) : 0;
  typedef void (*p2_fn)(const CUtensorMap, const LstmP2Params, int);
  const int wide = tt.ptable_wide;
  // 64-row clusters (variant 2, 16 epilogue warps): when every cluster then still gets its own SMs (B <= 64 * floor(SMs / CL)),
  // unless the caller asks for 128 (cluster_rows; a concurrent scan wants the encoder on as few SMs as possible)
  static const int env_rows = getenv("SSE_LSTM_CL_ROWS") ? atoi(getenv("SSE_LSTM_CL_ROWS")) : 0;
  const int want_rows = env_rows ? env_rows : cluster_rows;
  const bool rows64 = wide == 0 && tt.ptable_ew == 16 && want_rows != 128 && (want_rows == 64 || cdiv(B, 64) * (H / 32) <= num_sms);
  static const bool stage = getenv("SSE_LSTM_STAGE") ? atoi(getenv("SSE_LSTM_STAGE")) != 0 : true;
  const bool staged = wide == 32 && stage;
  const int ew = wide ? P3_EW : tt.ptable_ew;
  const int gm = tt.ptable_mode;
  p2_fn fn = wide == 64 ? (gm == 0 ? lstm_wide_kernel<0, 64, false> : gm == 1 ? lstm_wide_kernel<1, 64, false> : lstm_wide_kernel<2, 64, false>)
           : wide == 32 ? (staged ? (gm == 0 ? lstm_wide_kernel<0, 32, true> : gm == 1 ? lstm_wide_kernel<1, 32, true> : lstm_wide_kernel<2, 32, true>)
                                  : (gm == 0 ? lstm_wide_kernel<0, 32, false> : gm == 1 ? lstm_wide_kernel<1, 32, false> : lstm_wide_kernel<2, 32, false>))
           : ew == 8 ? (gm == 0 ? lstm_ptable_kernel<8, 0, 128> : gm == 1 ? lstm_ptable_kernel<8, 1, 128> : lstm_ptable_kernel<8, 2, 128>)
                     : rows64 ? (gm == 0 ? lstm_ptable_kernel<16, 0, 64> : gm == 1 ? lstm_ptable_kernel<16, 1, 64> : lstm_ptable_kernel<16, 2, 64>)
                              : (gm == 0 ? lstm_ptable_kernel<16, 0, 128> : gm == 1 ? lstm_ptable_kernel<16, 1, 128> : lstm_ptable_kernel<16, 2, 128>);
  const int CL = wide ? H / wide : H / 32, KBh = H / KBLK;
  const int rows_cl = rows64 ? 64 : P2_ROWS;
  const int n_clusters = cdiv(B, rows_cl);
  const int grid = n_clusters * CL;
  const size_t smem = wide ? 1024 + (size_t)(H / 16) * P2_SUB_BYTES + (size_t)KBh * (4 * wide * KBLK * 2) + (staged ? (size_t)(wide / 16) * P3_TAB_REGION : 0) + 256
                           : 1024 + (size_t)2 * 2 * CL * (rows_cl * 32) + (size_t)KBh * W_TILE_BYTES + 256;
  const bool want_dbg = getenv("SSE_LSTM_DEBUG") != nullptr;
  long long* d_dbg = nullptr;
  const size_t dbg_bytes = (size_t)grid * 128 + (size_t)P2_TRACE_STEPS * 64;
  if (want_dbg) { cudaMalloc(&d_dbg, dbg_bytes); cudaMemset(d_dbg, 0, dbg_bytes); p.dbg = d_dbg; }
  SSE_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)(ew + 2 + (staged ? 1 : 0)) * 32, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SSE_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, *reinterpret_cast<const CUtensorMap*>(tt.tmap2d), p, We / KBLK));
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  if (want_dbg) {
    std::vector<long long> hd((size_t)grid * 16 + (size_t)P2_TRACE_STEPS * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(hd.data(), d_dbg, hd.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_dbg);
    fprintf(stderr, "[lstm ptable dbg] variant: %s, %d rows per cluster, %d units per CTA%s, %d epilogue warps, gate math %d, poll back-off %d ns\n", wide ? "3 (single h tile)" : "2", rows_cl, wide ? wide : 32, staged ? ", staged table rows" : "", ew, tt.ptable_mode, poll_ns);
    {   // CTA 0's timeline of a few mid-sequence steps, relative to the step's accumulator-ready time
      const long long* tr = hd.data() + (size_t)grid * 16;
      for (int s = 20; s < std::min(24, T - t_start - 1); ++s) {
        const long long a = tr[s * 8 + 0];
        fprintf(stderr, "[lstm ptable trace] step %2d: acc ready +0 | pass0 published %+lld | pass1 published %+lld | own copies issued p0 %+lld p1 %+lld | "
                        "NEXT step: h pass0 landed %+lld, pass1 landed %+lld, acc ready %+lld\n",
                s, tr[s * 8 + 1] - a, tr[s * 8 + 2] - a, tr[s * 8 + 6] - a, tr[s * 8 + 7] - a, tr[(s + 1) * 8 + 4] - a, tr[(s + 1) * 8 + 5] - a, tr[(s + 1) * 8 + 0] - a);
      }
    }
    const char* nm[8] = {"mma_wait_hready", "-", "-", "mma_total", "epi_wait_accf", "epi_total", "-", "epi_ld+math+store"};
    for (int c = 0; c < 8; ++c) {
      if (nm[c][0] == '-') continue;
      long long sm = 0;
      for (int i = 0; i < grid; ++i) sm += hd[(size_t)i * 16 + c];
      fprintf(stderr, "[lstm ptable dbg] %-18s avg %10lld cycles  (grid %d = %d clusters x %d, steps %d)\n", nm[c], sm / grid, grid, n_clusters, CL, T - t_start);
    }
  }
  return SSE_OK;
}

}  // namespace sse
