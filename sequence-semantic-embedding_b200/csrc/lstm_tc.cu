// LSTM tower on the 5th-gen tensor cores (sm_100a): one persistent CTA per 128 batch rows runs
// all T steps on chip.  Restates BasicLSTMCell(forget_bias=1)+static_rnn (reference
// sse_model.py:222-224,240-242,248-250,262-264,273-274) with fp16 operands / fp32 accumulate;
// cell state c and all gate math stay fp32.  Measured deviation from the fp32 oracle on the
// normalised encodings: <= ~7e-4 abs (north_star tolerance 1e-3); the fp32 SIMT path
// (lstm_simt.cu) remains the exact mode.
//
// Per step t and per chunk of 32 hidden units (128 gate columns = i|j|f|o x 32):
//     z[128 rows, 128] = x_t[128, We] * Wx_chunk  (A from smem, gathered fp16 embedding rows)
//                      + h_{t-1}[128, H] * Wh_chunk (A from TENSOR MEMORY, written by the epilogue)
//   the fp16 weights stream through a TMA ring (B operand, [128 gate cols x 64 k] SW128 tiles of
//   the pre-transposed, chunk-major matrix Wt[4H, We+H]); accumulators are double-buffered in
//   TMEM; the epilogue (thread == batch row) applies bias, sigmoid/tanh (tanh.approx), updates
//   c (fp32, L2-resident scratch) and writes h_t back to TMEM as packed fp16 for the next step.
// TMEM columns: h ping [0,128) | h pong [128,256) | acc0 [256,384) | acc1 [384,512).
// warps 0-7 epilogue (warp w: TMEM lane quarter w%4, unit half w/4 of each 32-unit chunk) | warp 8 MMA issuer |
// warp 9 TMA producer | warp 10 TMEM alloc | warp 11 embedding gather (cp.async, 4 rows per lane, manual 128B
// swizzle).  Single-thread roles run warp-converged and issue through elect_one_sync().
#include "sse_common.cuh"
#include <cuda.h>
#include <math_constants.h>
#include <stdlib.h>

namespace sse {

namespace {

constexpr int KBLK = 64;
constexpr int TILE_BYTES = 128 * KBLK * 2;   // 16 KB
constexpr int LSTM_THREADS = 384;
constexpr int CHUNK_UNITS = 32;

struct LstmTcParams {
  const int32_t* tokens;      // [B, T]
  const __half* emb;   // [V, We] fp16
  const float* bias_r;        // [4H] chunk-major, +1 folded into the forget gate
  const float* init_h;        // optional [H] broadcast initial state (pad-prefix table row)
  const float* init_c;
  const int32_t* lead_sorted; // optional [B]: leading PADs of each (sorted) row -> per-tile start step (tok_prep.cu)
  const float* pad_h;         // with lead_sorted: pad-prefix state table [T][H] (entry t = state after t+1 PADs)
  const float* pad_c;
  float* c_scratch;           // [Bpad, H] fp32
  float* h_out;               // [B, H] fp32 (last step)
  int B, T, t_start, We, H, n_stages;
  long long* dbg;             // optional [grid][8] cycle counters
  int slot_kb;                // k-blocks per weight-ring slot
  int rows_per_cta;           // 128, or 64 / 32 for small batches (more CTAs, idle TMEM lane quarters skip the epilogue)
  int xbufs;                  // 1 or 2 x_t buffers in shared memory
  int bias_smem;              // 1: bias table staged in shared memory, 0: read through L1
  int gate_math;              // 0: ex2/rcp merged form (8 MUFU per unit), 1: tanh.approx (5 MUFU.TANH per unit)
  int use3d;                  // 1: one 3-D TMA per slot; 0: KB 2-D loads per slot (fallback if the 3-D map is rejected)
};

// ---------------------------------------------------------------- PTX helpers (see search_tc.cu)
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t backoff_ns = 32) {
  // Poll with try_wait and SLEEP between polls (nanosleep): a busy-polling waiter steals issue slots from
  // the warps that are doing the work on its SM sub-partition (the waiters include the high-priority
  // single-thread roles, so a spinning producer starved the epilogue warps -- measured 2x slowdown), while
  // the try_wait suspend-time hint showed a ~700-cycle wake-up quantum.  A watchdog turns a protocol bug
  // (a wait that can never complete) into a trap instead of a hung GPU.
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (backoff_ns) __nanosleep(backoff_ns);
    if ((spins & 0xfff) == 0xfff) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();   // ~2 s at 2 GHz
    }
  }
}
__device__ __forceinline__ void mbar_wait_timed(uint32_t bar, uint32_t parity, long long& acc, uint32_t backoff_ns = 32) {
  long long t = clock64();
  mbar_wait(bar, parity, backoff_ns);
  acc += clock64() - t;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// One elected lane of a CONVERGED warp (cutlass elect_one_sync).  The single-thread roles must run their
// loops warp-converged with warp-uniform operands and wrap only the issue in this predicate: under a plain
// `if (lane == 0)` branch nvcc cannot keep descriptors in uniform registers and emits an ELECT/R2UR retry loop
// around every UTCHMMA / UTMALDG -- measured ~208 cycles per tcgen05.mma issue instead of the tensor-pipe floor.
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %1;\n@%%px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
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
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // D=f32, A=B=f16
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh on the FMA pipe: odd rational minimax p(x)/q(x) on the clamped argument (the classic float
// approximation used by Eigen/XLA, |rel err| ~ 1e-7) -- 13 FMA-pipe ops + one MUFU.RCP.  MUFU.TANH has a
// much lower issue rate; mixing the two keeps the MUFU and FMA pipes both under the MMA time per chunk.
__device__ __forceinline__ float tanh_rational(float x) {
  x = fminf(fmaxf(x, -7.90531110763549805f), 7.90531110763549805f);
  const float x2 = x * x;
  float p = fmaf(x2, -2.76076847742355e-16f, 2.00018790482477e-13f);
  p = fmaf(x2, p, -8.60467152213735e-11f);
  p = fmaf(x2, p, 5.12229709037114e-08f);
  p = fmaf(x2, p, 1.48572235717979e-05f);
  p = fmaf(x2, p, 6.37261928875436e-04f);
  p = fmaf(x2, p, 4.89352455891786e-03f);
  p = x * p;
  float q = fmaf(x2, 1.19825839466702e-06f, 1.18534705686654e-04f);
  q = fmaf(x2, q, 2.26843463243900e-03f);
  q = fmaf(x2, q, 4.89352518554385e-03f);
  return __fdividef(p, q);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_approx(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
  return r;
}

#define TMEM_LD_32(taddr, v)                                                                                       \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
      "%29,%30,%31}, [%32];"                                                                                       \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),     \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),    \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                  \
      : "r"(taddr))

#define TMEM_LD_16(taddr, v)                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),      \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])  \
               : "r"(taddr))

#define TMEM_ST_8(taddr, v)                                                                                  \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), \
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])                     \
               : "memory")

#define TMEM_ST_16(taddr, v)                                                                                   \
  asm volatile(                                                                                                \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),    \
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])             \
      : "memory")

// descriptor given as (lo, hi) words: only the low word (address field) varies per MMA
__device__ __forceinline__ void tc_mma_ss2(uint32_t d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 ad, bd;\nsetp.ne.b32 p, %5, 0;\nmov.b64 ad, {%1, %3};\nmov.b64 bd, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, p;\n}\n" ::"r"(d),
      "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_mma_ts2(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 bd;\nsetp.ne.b32 p, %5, 0;\nmov.b64 bd, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n}\n" ::"r"(d),
      "r"(a_tmem), "r"(blo), "r"(hi), "r"(idesc), "r"(acc)
      : "memory");
}

// KXT / KHT: compile-time We/64 and H/64 (0 = run-time)
template <int KXT, int KHT>
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap tmap_w3, const __grid_constant__ CUtensorMap tmap_w2d,
               const __grid_constant__ LstmTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KBx = KXT ? KXT : P.We / KBLK, KBh = KHT ? KHT : P.H / KBLK, NC = P.H / CHUNK_UNITS, NS = P.n_stages;
  const int SKB = P.slot_kb;                                  // k-blocks per weight-ring slot (2, or 1 for odd KB)
  const int nsx = KBx / SKB, nsh = KBh / SKB;                 // slots per chunk: x part, h part
  const int RPC = P.rows_per_cta;
  const int n_act = RPC / 32;                                 // active TMEM lane quarters
  const int row0 = blockIdx.x * RPC;
  // per-tile pad-prefix start (tok_prep.cu): rows are sorted by descending number of leading PADs, so the tile's LAST row
  // has the shortest prefix; the tile starts at t0 from the tabulated state after t0 PADs
  int t0 = P.t_start;
  const float* init_h = P.init_h;
  const float* init_c = P.init_c;
  if (P.lead_sorted) {
    t0 = min(__ldg(P.lead_sorted + min(row0 + RPC - 1, P.B - 1)), P.T - 1);
    init_h = t0 > 0 ? P.pad_h + (size_t)(t0 - 1) * P.H : nullptr;
    init_c = t0 > 0 ? P.pad_c + (size_t)(t0 - 1) * P.H : nullptr;
  }
  const bool has_init = init_h != nullptr;
  const uint32_t slot_bytes = (uint32_t)SKB * TILE_BYTES;

  const int XB = P.xbufs;
  // x tiles hold only this CTA's RPC rows: the MMA still reads 128 rows per tile, the rows past RPC are whatever
  // follows in shared memory (finite fp16 weight bits) and only feed accumulator lanes nobody reads.
  const uint32_t x_tile_bytes = (uint32_t)RPC * 128;
  uint8_t* x_smem = smem;                                     // [XB buffers][KBx] tiles [RPC x 64] fp16 SW128
  uint8_t* w_smem = x_smem + (size_t)XB * KBx * x_tile_bytes; // [NS] slots of SKB tiles
  float* bias_s = reinterpret_cast<float*>(w_smem + (size_t)NS * slot_bytes);   // [4H] (only if P.bias_smem)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + (P.bias_smem ? (size_t)4 * P.H * 4 : 0));
  // bars: full[NS], empty[NS], x_full[2], x_empty[2], h_full, acc_full[2], acc_empty[2]
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + NS);
  const uint32_t bar_xf = smem_u32(bars + 2 * NS);
  const uint32_t bar_xe = smem_u32(bars + 2 * NS + 2);
  const uint32_t bar_hf = smem_u32(bars + 2 * NS + 4);
  const uint32_t bar_accf = smem_u32(bars + 2 * NS + 5);
  const uint32_t bar_acce = smem_u32(bars + 2 * NS + 7);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 9);

  if (P.bias_smem)
    for (int i = threadIdx.x; i < 4 * P.H; i += LSTM_THREADS) bias_s[i] = P.bias_r[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_xf + 8 * b, 64);
      mbar_init(bar_xe + 8 * b, 1);
      mbar_init(bar_accf + 8 * b, 1);
      mbar_init(bar_acce + 8 * b, 2 * n_act);
    }
    mbar_init(bar_hf, 2 * n_act);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 10) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // ===== TMA producer: weights, chunk-major; whole warp converged, one elected lane issues =====
    // slot order per step (must mirror the MMA warp): for each PAIR of chunks (c, c+1):
    //   x(c), x(c+1), h(c), h(c+1)  -- the x parts of two chunks do not need h_{t-1}, so they fill the
    //   tensor pipe while the epilogue of the previous step's last chunks is still producing h.
    uint32_t it = 0;
    for (int t = t0; t < P.T; ++t) {
      const bool has_state = has_init || t > t0;
      for (int cp = 0; cp < NC; cp += 2) {
        const int nphase = has_state ? 4 : 2;
        for (int phs = 0; phs < nphase; ++phs) {
          const int c = cp + (phs & 1);
          const int nsl = phs < 2 ? nsx : nsh;
          const int kbase = phs < 2 ? 0 : KBx;
          for (int sl = 0; sl < nsl; ++sl, ++it) {
            const uint32_t s = it % NS, ph = (it / NS) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1, 64);
            if (elect_one_sync()) {
              const uint32_t dst = smem_u32(w_smem + (size_t)s * slot_bytes);
              const int kb0 = kbase + sl * SKB;
              mbar_expect_tx(bar_full + 8 * s, slot_bytes);
              if (P.use3d) {
                tma_load_3d(dst, &tmap_w3, bar_full + 8 * s, 0, c * 128, kb0);
              } else {
                for (int kb = 0; kb < SKB; ++kb)
                  tma_load_2d(dst + (uint32_t)kb * TILE_BYTES, &tmap_w2d, bar_full + 8 * s, (kb0 + kb) * KBLK, c * 128);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 8) {
    // ===== MMA issuer: whole warp converged, one elected lane issues (warp-uniform operands) =====
    const uint32_t idesc = make_idesc_f16(128, 128);
    uint32_t it = 0, gchunk = 0;
    long long w_acce = 0, w_xf = 0, w_full = 0, w_hf = 0, t_begin = clock64();
    for (int t = t0; t < P.T; ++t) {
      const int step = t - t0;
      const int xb = XB == 2 ? (step & 1) : 0;
      const uint32_t xuse = XB == 2 ? (uint32_t)(step >> 1) : (uint32_t)step;
      const bool has_state = has_init || t > t0;
      // h_{t-1} lives in h buffer (t+1)&1 ; epilogue of step t writes buffer t&1
      const uint32_t h_src = tmem_base + (uint32_t)(((t + 1) & 1) * 128);
      const uint32_t x_lo = (uint32_t)make_sw128_desc(smem_u32(x_smem + (size_t)xb * KBx * x_tile_bytes));
      for (int cp = 0; cp < NC; cp += 2, gchunk += 2) {
        const uint32_t use = gchunk >> 1;                 // NC is even: chunk c always uses accumulator c & 1
        // ---- x parts of chunks cp (acc 0) and cp+1 (acc 1)
        for (int b2 = 0; b2 < 2; ++b2) {
          const int c = cp + b2;
          const uint32_t d = tmem_base + 256u + (uint32_t)(b2 * 128);
          mbar_wait_timed(bar_acce + 8 * b2, (use & 1) ^ 1, w_acce);
          tc_fence_after();
          if (c == 0) { mbar_wait_timed(bar_xf + 8 * xb, xuse & 1, w_xf); tc_fence_after(); }
          for (int sl = 0; sl < nsx; ++sl, ++it) {
            const uint32_t s = it % NS, ph = (it / NS) & 1;
            mbar_wait_timed(bar_full + 8 * s, ph, w_full);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint64_t b0 = make_sw128_desc(smem_u32(w_smem + (size_t)s * slot_bytes));
              const uint32_t blo = (uint32_t)b0, hi = (uint32_t)(b0 >> 32);
              const uint32_t alo = x_lo + (uint32_t)(sl * SKB) * (x_tile_bytes >> 4);
#pragma unroll 2
              for (int kb = 0; kb < SKB; ++kb)
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                  tc_mma_ss2(d, alo + (uint32_t)(kb * (x_tile_bytes >> 4) + 2 * k4), blo + (uint32_t)(kb * (TILE_BYTES >> 4) + 2 * k4), hi, idesc,
                             (sl | kb | k4) ? 1u : 0u);
              tc_commit(bar_empty + 8 * s);
              if (c == NC - 1 && sl == nsx - 1) tc_commit(bar_xe + 8 * xb);   // x_t fully consumed once these MMAs retire
              if (!has_state && sl == nsx - 1) tc_commit(bar_accf + 8 * b2);  // no recurrent part at the very first step
            }
            __syncwarp();
          }
        }
        // ---- h parts (A operand = h_{t-1} in tensor memory)
        if (has_state) {
          if (cp == 0) {
            // completions of h_full: [initial state staged (only with init)], end of step t0, t0+1, ...
            const uint32_t idx = has_init ? (uint32_t)step : (uint32_t)(step - 1);
            mbar_wait_timed(bar_hf, idx & 1, w_hf);
            tc_fence_after();
          }
          for (int b2 = 0; b2 < 2; ++b2) {
            const uint32_t d = tmem_base + 256u + (uint32_t)(b2 * 128);
            for (int sl = 0; sl < nsh; ++sl, ++it) {
              const uint32_t s = it % NS, ph = (it / NS) & 1;
              mbar_wait_timed(bar_full + 8 * s, ph, w_full);
              tc_fence_after();
              if (elect_one_sync()) {
                const uint64_t b0 = make_sw128_desc(smem_u32(w_smem + (size_t)s * slot_bytes));
                const uint32_t blo = (uint32_t)b0, hi = (uint32_t)(b0 >> 32);
                const uint32_t a0 = h_src + (uint32_t)(sl * SKB * 32);
#pragma unroll 2
                for (int kb = 0; kb < SKB; ++kb)
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4)
                    tc_mma_ts2(d, a0 + (uint32_t)(kb * 32 + k4 * 8), blo + (uint32_t)(kb * (TILE_BYTES >> 4) + 2 * k4), hi, idesc, 1u);
                tc_commit(bar_empty + 8 * s);
                if (sl == nsh - 1) tc_commit(bar_accf + 8 * b2);
              }
              __syncwarp();
            }
          }
        }
      }
    }
    if (P.dbg && lane == 0) {
      long long* o = P.dbg + blockIdx.x * 8;
      o[0] = w_acce; o[1] = w_xf; o[2] = w_full; o[3] = w_hf; o[4] = clock64() - t_begin;
    }
  } else if (warp >= 10) {
    // ===== embedding gather (warps 10-11, 2 rows per lane): cp.async 16 B chunks into the 128B-swizzled x tiles,
    //       double-buffered: x_{t+1} is fetched while step t computes =====
    const int gl = (warp - 10) * 32 + lane;        // 0..63
    int grow[2];
    uint32_t row_off[2], sw[2];
    int tok[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = q * 64 + gl;
      grow[q] = min(row0 + min(r, RPC - 1), P.B - 1);
      row_off[q] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
      sw[q] = (uint32_t)(r & 7);
      tok[q] = __ldg(P.tokens + (size_t)grow[q] * P.T + t0);
    }
    for (int t = t0; t < P.T; ++t) {
      const int step = t - t0;
      const int xb = XB == 2 ? (step & 1) : 0;
      const uint32_t xuse = XB == 2 ? (uint32_t)(step >> 1) : (uint32_t)step;
      mbar_wait(bar_xe + 8 * xb, (xuse & 1) ^ 1, 128);
      const uint32_t xbase = smem_u32(x_smem + (size_t)xb * KBx * x_tile_bytes);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q * 64 + gl >= RPC) continue;                 // rows beyond this CTA's tile: MMA rows are independent, leave them
        const __half* src = P.emb + (size_t)tok[q] * P.We;
        for (int kb = 0; kb < KBx; ++kb) {
          const uint32_t tile = xbase + (uint32_t)kb * x_tile_bytes + row_off[q];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t dst = tile + (uint32_t)(((j ^ sw[q]) * 16));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + kb * KBLK + j * 8) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (t + 1 < P.T) {          // next step's token ids while the rows are in flight
#pragma unroll
        for (int q = 0; q < 2; ++q) tok[q] = __ldg(P.tokens + (size_t)grow[q] * P.T + t + 1);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      mbar_arrive(bar_xf + 8 * xb);
    }
  } else if (warp < 8 && (warp & 3) < n_act) {
    // ===== epilogue: up to 8 warps; thread == (batch row, half of the chunk's 32 hidden units) =====
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const int grow = row0 + r;
    const bool valid = grow < P.B;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float* crow = P.c_scratch + (size_t)(row0 + r) * P.H;
    if (has_init) {
      // stage the broadcast initial state: h -> TMEM buffer (t0+1)&1 (read by step t0), c -> scratch
      const uint32_t hdst = lane_base + (uint32_t)(((t0 + 1) & 1) * 128);
      for (int u0 = half * 16; u0 < P.H; u0 += 32) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_f16x2(__ldg(init_h + u0 + 2 * i), __ldg(init_h + u0 + 2 * i + 1));
        TMEM_ST_8(hdst + u0 / 2, pk);
#pragma unroll
        for (int i = 0; i < 16; ++i) crow[u0 + i] = __ldg(init_c + u0 + i);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hf);      // phase 0 of h_full == "initial state staged"
    }
    uint32_t gchunk = 0;
    long long w_accf = 0, t_begin_e = clock64();
    for (int t = t0; t < P.T; ++t) {
      const bool has_c = has_init || t > t0;
      const bool last = t == P.T - 1;
      const uint32_t hdst = lane_base + (uint32_t)((t & 1) * 128);
      for (int c = 0; c < NC; ++c, ++gchunk) {
        const int buf = gchunk & 1;
        const uint32_t use = gchunk >> 1;
        const int u0 = c * CHUNK_UNITS + half * 16;       // first hidden unit of this thread's 16
        // previous cell state and the biases of this half chunk (issued before the accumulator wait)
        float cold[16];
        if (has_c) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 v = *reinterpret_cast<const float4*>(crow + u0 + q * 4);
            cold[q * 4] = v.x; cold[q * 4 + 1] = v.y; cold[q * 4 + 2] = v.z; cold[q * 4 + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) cold[i] = 0.f;
        }
        const float* bs = (P.bias_smem ? bias_s : P.bias_r) + c * 128 + half * 16;   // warp-uniform addresses (broadcast)
        mbar_wait_timed(bar_accf + 8 * buf, use & 1, w_accf);
        tc_fence_after();
        const uint32_t acc = lane_base + 256u + (uint32_t)(buf * 128 + half * 16);
        uint32_t vi[16], vj[16], vf[16], vo[16];
        TMEM_LD_16(acc, vi);
        TMEM_LD_16(acc + 32, vj);
        TMEM_LD_16(acc + 64, vf);
        TMEM_LD_16(acc + 96, vo);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // accumulator buffer is free as soon as it sits in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acce + 8 * buf);
        // Gate math with 8 MUFU ops per (row, unit): exponentials via ex2, and ONE reciprocal per product of
        // denominators:  sig(i)*tanh(j) = (1-Ej) / ((1+Ei)(1+Ej)),  tanh(c)*sig(o) = (1-Ec) / ((1+Ec)(1+Eo))
        // with Ei = e^-zi, Ej = e^-2zj, ...  bias_r is pre-scaled by -log2(e) (-2 log2(e) for the candidate
        // gate) so every ex2 argument is one FMA; arguments are capped at 57 (e^40) so products stay finite.
        constexpr float NL2E = -1.4426950408889634f;
        float hv[16];
        if (P.gate_math == 0) {
          float ei[16], ej[16], ef[16], eo[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bi = *reinterpret_cast<const float4*>(bs + q * 4);
            const float4 bj = *reinterpret_cast<const float4*>(bs + 32 + q * 4);
            const float4 bf = *reinterpret_cast<const float4*>(bs + 64 + q * 4);
            const float4 bo = *reinterpret_cast<const float4*>(bs + 96 + q * 4);
            const float bia[4] = {bi.x, bi.y, bi.z, bi.w}, bja[4] = {bj.x, bj.y, bj.z, bj.w};
            const float bfa[4] = {bf.x, bf.y, bf.z, bf.w}, boa[4] = {bo.x, bo.y, bo.z, bo.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = q * 4 + e;
              ei[i] = ex2_approx(fminf(fmaf(__uint_as_float(vi[i]), NL2E, bia[e]), 57.f));
              ej[i] = ex2_approx(fminf(fmaf(__uint_as_float(vj[i]), 2.f * NL2E, bja[e]), 57.f));
              ef[i] = ex2_approx(fminf(fmaf(__uint_as_float(vf[i]), NL2E, bfa[e]), 57.f));     // forget bias +1 folded in
              eo[i] = ex2_approx(fminf(fmaf(__uint_as_float(vo[i]), NL2E, boa[e]), 57.f));
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p = (1.f - ej[i]) * rcp_approx((1.f + ei[i]) * (1.f + ej[i]));
            cold[i] = fmaf(cold[i], rcp_approx(1.f + ef[i]), p);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float ec = ex2_approx(fminf(cold[i] * (2.f * NL2E), 57.f));
            hv[i] = (1.f - ec) * rcp_approx((1.f + ec) * (1.f + eo[i]));
          }
        } else {
          // tanh.approx variant: bias table is scaled by -log2e (-2 log2e for j): undo the scale in the argument
          constexpr float INV = -0.6931471805599453f;          // 1 / (-log2 e)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float zi = fmaf(bs[i], INV, __uint_as_float(vi[i]));
            const float zj = fmaf(bs[32 + i], 0.5f * INV, __uint_as_float(vj[i]));
            const float zf = fmaf(bs[64 + i], INV, __uint_as_float(vf[i]));
            const float zo = fmaf(bs[96 + i], INV, __uint_as_float(vo[i]));
            const float p = sigmoid_approx(zi) * tanh_approx(zj);
            cold[i] = fmaf(cold[i], sigmoid_approx(zf), p);
            hv[i] = tanh_approx(cold[i]) * sigmoid_approx(zo);
          }
        }
        if (!last) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = pack_f16x2(hv[2 * i], hv[2 * i + 1]);
          TMEM_ST_8(hdst + (uint32_t)(u0 / 2), pk);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(crow + u0 + q * 4) = make_float4(cold[q * 4], cold[q * 4 + 1], cold[q * 4 + 2], cold[q * 4 + 3]);
        } else if (valid) {
          float* ho = P.h_out + (size_t)grow * P.H + u0;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(ho + q * 4) = make_float4(hv[q * 4], hv[q * 4 + 1], hv[q * 4 + 2], hv[q * 4 + 3]);
        }
      }
      if (!last) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_hf);    // h_t complete in TMEM (this warp's rows / unit halves)
      }
    }
    if (P.dbg && warp == 0 && lane == 0) { P.dbg[blockIdx.x * 8 + 5] = w_accf; P.dbg[blockIdx.x * 8 + 6] = clock64() - t_begin_e; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// Wt[n'][k] = K[k][g*H + 32c + j],  n' = c*128 + g*32 + j ; bias_r[n'] = -log2e-scaled b[g*H + 32c + j] (+1 for g == 2)
__global__ void prep_weights_kernel(const float* __restrict__ K, const float* __restrict__ b, int We, int H,
                                    __half* __restrict__ Wt, float* __restrict__ bias_r) {
  const int Kd = We + H;
  int64_t total = (int64_t)4 * H * Kd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int np = (int)(i / Kd), k = (int)(i - (int64_t)np * Kd);
    int c = np >> 7, g = (np >> 5) & 3, j = np & 31;
    int col = g * H + c * CHUNK_UNITS + j;
    Wt[i] = __float2half_rn(K[(size_t)k * 4 * H + col]);
    // pre-scaled for the ex2-based gate math: -log2(e) * (b [+1 forget bias]); -2 log2(e) * b for the candidate gate
    if (k == 0) bias_r[np] = (g == 1 ? -2.885390081777927f : -1.4426950408889634f) * (b[col] + (g == 2 ? 1.0f : 0.0f));
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

bool lstm_tc_supported(int We, int H) { return We % 64 == 0 && We >= 64 && We <= 256 && H % 64 == 0 && H >= 64 && H <= 256; }

void lstm_tc_release(TcTower& tt) {
  if (tt.wt) cudaFree(tt.wt);
  if (tt.bias_r) cudaFree(tt.bias_r);
  tt.wt = nullptr; tt.bias_r = nullptr; tt.valid = false;
}

int lstm_tc_prepare(TcTower& tt, const float* K, const float* b, int We, int H, cudaStream_t st, int64_t* launches) {
  if (!tt.wt) {
    SSE_CUDA_OK(cudaMalloc(&tt.wt, (size_t)4 * H * (We + H) * 2));
    SSE_CUDA_OK(cudaMalloc(&tt.bias_r, (size_t)4 * H * 4));
  }
  prep_weights_kernel<<<148 * 4, 256, 0, st>>>(K, b, We, H, tt.wt, tt.bias_r);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return SSE_ECUDA;
  }
  // 3-D view of Wt[4H, We+H] fp16: (k within a 64-wide block, gate row, k-block); one box = a half chunk
  // [128 gate rows x KBx (or KBh) k-blocks], landing in smem as consecutive [128 x 64] SWIZZLE_128B tiles.
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(p);
  tt.use3d = getenv("SSE_LSTM_NO3D") == nullptr;
  const int KBx = We / KBLK, KBh = H / KBLK;
  cuuint64_t gdim[3] = {(cuuint64_t)KBLK, (cuuint64_t)(4 * H), (cuuint64_t)(KBx + KBh)};
  cuuint64_t gstr[2] = {(cuuint64_t)(We + H) * 2, (cuuint64_t)KBLK * 2};
  cuuint32_t estr[3] = {1, 1, 1};
  {
    tt.slot_kb = getenv("SSE_LSTM_SLOT_KB") ? atoi(getenv("SSE_LSTM_SLOT_KB")) : 4;
    if (tt.slot_kb < 1) tt.slot_kb = 1;
    while (tt.slot_kb > 1 && (KBx % tt.slot_kb || KBh % tt.slot_kb)) tt.slot_kb >>= 1;
    cuuint32_t box[3] = {KBLK, 128, (cuuint32_t)tt.slot_kb};
    CUresult r = enc(reinterpret_cast<CUtensorMap*>(tt.tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, tt.wt, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) tt.use3d = false;
  }
  {
    cuuint64_t gdim2[2] = {(cuuint64_t)(We + H), (cuuint64_t)(4 * H)};
    cuuint64_t gstr2[1] = {(cuuint64_t)(We + H) * 2};
    cuuint32_t box2[2] = {KBLK, 128};
    cuuint32_t estr2[2] = {1, 1};
    CUresult r = enc(reinterpret_cast<CUtensorMap*>(tt.tmap2d), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, tt.wt, gdim2, gstr2, box2, estr2,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(Wt) failed (%d)", (int)r); return SSE_ECUDA; }
  }
  tt.valid = true;
  return SSE_OK;
}

int lstm_forward_tc(const int32_t* tokens, int B, int T, int t_start, const __half* emb_f16, int We, int H,
                    const TcTower& tt, const float* init_h, const float* init_c, const PadSkip& ps, float* c_scratch, float* h_out,
                    cudaStream_t st, int64_t* launches) {
  if (B <= 0) return SSE_OK;
  LstmTcParams p;
  p.tokens = tokens; p.emb = emb_f16; p.bias_r = tt.bias_r; p.init_h = init_h; p.init_c = init_c;
  p.lead_sorted = ps.lead_sorted; p.pad_h = ps.pad_h; p.pad_c = ps.pad_c;
  p.c_scratch = c_scratch; p.h_out = h_out; p.B = B; p.T = T; p.t_start = t_start; p.We = We; p.H = H;
  // small batches: 64- / 32-row tiles spread the rows over more SMs (the step latency, not the MMA rate, bounds them)
  const int rpc_auto = (cdiv(B, 128) * 2 >= 148) ? 128 : ((cdiv(B, 64) * 2 >= 148) ? 64 : 32);
  p.rows_per_cta = getenv("SSE_LSTM_ROWS") ? atoi(getenv("SSE_LSTM_ROWS")) : rpc_auto;
  // shared-memory plan (tunable for experiments): x buffers, weight-ring slot size, bias table placement
  auto envi = [](const char* k, int dflt) { const char* v = getenv(k); return v ? atoi(v) : dflt; };
  p.xbufs = envi("SSE_LSTM_XBUFS", 1);
  p.bias_smem = envi("SSE_LSTM_BIAS_SMEM", 1);
  p.gate_math = envi("SSE_LSTM_GATE_MATH", 0);
  const int slot_kb = tt.slot_kb;
  const size_t x_bytes = (size_t)p.xbufs * (We / KBLK) * p.rows_per_cta * 128;
  // the last x tile is read 128 rows deep: keep that window inside the allocation (the ring follows it)
  const size_t fixed = 1024 + x_bytes + (p.bias_smem ? (size_t)4 * H * 4 : 0) + 512;
  const size_t slot = (size_t)slot_kb * TILE_BYTES;
  int NS = (int)((232448 - fixed) / slot);
  if (NS > 8) NS = 8;
  if (NS < 2) { set_error("lstm_tc: shared memory budget"); return SSE_EINVAL; }
  p.n_stages = NS;
  p.use3d = tt.use3d ? 1 : 0;
  p.slot_kb = slot_kb;
  p.dbg = nullptr;
  const bool want_dbg = getenv("SSE_LSTM_DEBUG") != nullptr;
  long long* d_dbg = nullptr;
  const int grid = cdiv(B, p.rows_per_cta);
  if (want_dbg) { cudaMalloc(&d_dbg, (size_t)grid * 64); cudaMemset(d_dbg, 0, (size_t)grid * 64); p.dbg = d_dbg; }
  const size_t smem = fixed + (size_t)NS * slot;
  typedef void (*lstm_fn)(const CUtensorMap, const CUtensorMap, const LstmTcParams);
  lstm_fn fn = (We == 256 && H == 256) ? lstm_tc_kernel<4, 4> : lstm_tc_kernel<0, 0>;
  SSE_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  fn<<<grid, LSTM_THREADS, smem, st>>>(*reinterpret_cast<const CUtensorMap*>(tt.tmap),
                                               *reinterpret_cast<const CUtensorMap*>(tt.tmap2d), p);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  if (want_dbg) {
    std::vector<long long> hd((size_t)grid * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(hd.data(), d_dbg, hd.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_dbg);
    const char* nm[7] = {"mma_wait_acce", "mma_wait_xfull", "mma_wait_wfull", "mma_wait_hfull", "mma_total", "epi_wait_accf", "epi_total"};
    for (int c = 0; c < 7; ++c) {
      long long sm = 0;
      for (int i = 0; i < grid; ++i) sm += hd[(size_t)i * 8 + c];
      fprintf(stderr, "[lstm dbg] %-15s avg %10lld cycles  (grid %d, steps %d, chunks/step %d, NS %d, use3d %d)\n", nm[c], sm / grid, grid,
              T - t_start, H / CHUNK_UNITS, NS, p.use3d);
    }
  }
  return SSE_OK;
}

}  // namespace sse
