// Cosine scoring + row top-k on the 5th-gen tensor cores (sm_100a): fp16 tcgen05
// scan of the whole index with a threshold-filter epilogue, then an EXACT fp32
// re-score of the few survivors.  Replaces np.dot + full argsort
// (reference sse_evaluator.py:110-111, data_utils.py:263-267) for E % 64 == 0.
//
// Pipeline per query batch (all on one stream, no host sync):
//   1. prep_queries      q fp32 [Q,E] -> fp16 [Qp,E] (zero padded) + row norms
//   2. scan<TILEMAX>     tcgen05 GEMM over a strided SAMPLE of 128-row index tiles;
//                        epilogue = per (row, tile) max            (no divergence)
//   3. select_tau        tau[r] = k-th largest sampled tile max - margin[r]
//                        (a valid lower bound of the k-th best approximate score)
//   4. scan<FILTER>      tcgen05 GEMM over ALL tiles; epilogue compares every score
//                        with tau[r] in registers and appends the rare survivors
//                        to a private per-(CTA,row) candidate list
//   5. finalize          per row: sort candidates, keep those within the fp16 error
//                        margin of the k-th best, re-score them in fp32 against the
//                        fp32 index, sort by (score desc, idx asc), emit k
//   6. fallback          rows whose candidate lists overflowed (pathological ties)
//                        are recomputed by brute force in fp32
// Exactness: |approx - exact| <= eps_r = EPS_REL*|q_r|*max|t|, EPS_REL = 0.0011 (fp16 operand rounding 2*2^-11
// plus fp32 accumulation slack; derivation next to EPS_REL).  Every exact top-k element has approx >= A_k - 2 eps >= tau, so it
// survives 4 and 5; the final order/scores come from fp32 arithmetic only.
//
// GEMM mapping: queries are the A operand (M = 128 rows per m-tile, 1 or 2 m-tiles
// resident in smem per CTA), the index streams through a TMA ring as the B operand
// (N = 128 index rows per tile, K = E), fp32 accumulators are double-buffered in
// TMEM; TMEM lane == query row, so each epilogue thread owns one row.
#include "sse_common.cuh"
#include <cuda.h>
#include <math_constants.h>
#include <math.h>
#include <stdlib.h>

namespace sse {

namespace {

constexpr int TILE_M = 128;          // query rows per m-tile
constexpr int KBLK = 64;             // fp16 elements per 128-byte swizzle row
constexpr int CAND_CAP = 128;        // candidates per (CTA item, row); compacted in-kernel when fewer than 32 slots remain
constexpr int MAX_GROUPS = 64;
constexpr int DBG_N = 12;            // debug cycle counters per work item
constexpr int SCAN_THREADS = 384;
constexpr int FIN_MAXC_SMALL = 2048, FIN_MAXC_LARGE = 4096;   // candidates per row the finalize kernel can sort (k <= 32 / larger k)
constexpr int FIN_MAXR = 256;        // candidates per row re-scored exactly
// Bound of |fp16-operand dot - exact dot| / (|q| |t|), valid for every supported E (<= 512):
//   operand rounding: q_i(1+a_i) t_i(1+b_i), |a_i|,|b_i| <= u = 2^-11  ->  sum |q_i t_i| (2u + u^2) <= (2u + u^2) |q||t|
//                     = 0.000977 |q||t|                                   (Cauchy-Schwarz);
//   fp32 accumulation in the tensor core: E/16 k-steps, each adding a 16-term partial sum: <= (E/16 + 16) 2^-23 |q||t|
//                     = 5.8e-6 at E = 512 (2.0e-6 at 256);
//   fp16 subnormal index elements (|t_i| < 6.1e-5, absolute rounding error <= 2^-25): <= sqrt(E) 2^-25 |q| = 6.8e-7 |q|
//                     at E = 512 -- negligible for the normalised index rows the path produces (|t| = 1).  Queries
//                     are rescaled to max |q_i| in [0.5, 1) (prep_queries_kernel), so they have no subnormal issue.
// Sum < 0.000985; EPS_REL = 0.0011 leaves > 10 % slack at E = 512.
constexpr float EPS_REL = 0.0011f;

enum { MODE_TILEMAX = 0, MODE_FILTER = 1, MODE_FUSED = 2 };
constexpr int FUSED_MAX_K = 16;     // the fused scan keeps a row's k best sampled tile maxima in registers

struct ScanParams {
  int n_groups;
  int mtg;                         // m-tiles per group (smem / TMEM are sized for this)
  int kb;                          // E / 64
  int n_stages;
  int tn;                          // index rows per MMA tile (64 or 128)
  int a_cols;                      // TMEM columns holding the fp16 queries: mtg * E / 2
  const __half* qb;         // [Qp, E] fp16 queries (TILEMAX / FILTER; the fused scan converts q32 itself)
  // fused scan (MODE_FUSED): sample pass + threshold selection + filter pass in ONE launch
  const float* q32;                // [Q, E_true] fp32 queries
  int Q, E_true, gstride, rpg;     // padded row p = g * gstride + l  <->  query row g * rpg + l
  const float* tnorm_max;          // max |t| of the index
  float* margin_out;               // [Qp] 2*eps per row, written for finalize
  float* tau_out;                  // [Qp] (debug / tests)
  unsigned* group_ctr;             // [n_groups] arrival counters of the per-group barrier (zero between searches)
  int n_s, s_step;                 // sample tiles: tile = js * s_step, js in [0, n_s)
  long long* dbg;                  // optional [items][DBG_N] cycle counters (nullptr = off)
  int dbg_flags;                   // timing experiments only: 1 = no TMA loads, 2 = no MMA issue
  int use3d;                       // one 3-D TMA instruction per index tile (else KB 2-D loads)
  int cs;                          // CTAs per thread-block cluster (1 = no cluster): the CTAs of a cluster hold
                                   // different m-groups and sweep the SAME tile range; every index tile is fetched
                                   // once by the cluster leader and TMA-multicast into all cs shared memories
  int R;                           // clusters per super-group (a super-group = cs consecutive m-groups)
  int group_first_item[MAX_GROUPS];  // cs == 1: items of group g are [first, first + items)
  int group_items[MAX_GROUPS];
  // LATE items (cs == 1, optional): extra CTAs numbered after all regular items, meant for SMs that a concurrent kernel
  // (the encoder of the next batch, on another stream) still holds when the scan starts.  A late item takes late_share/256
  // of a regular item's tile range, skips the sample pass and never arrives at the group barrier -- it only waits for it.
  // SPANNING items (cs == 1, equal packing, not fused): the (group, tile) pairs of ALL m-groups form one list of
  // n_groups * n_j units that is cut into gridDim.x equal pieces, so every CTA carries the same work whatever the ratio of
  // CTAs to groups (19 groups on 148 SMs: 7 items per group used 133 SMs).  A piece that crosses a group boundary has two
  // SEGMENTS: the CTA restages its queries (TMEM A operand) between them and keeps a separate candidate slot per segment.
  int span;
  // FOLLOWING last group (filter pass, full packing, cs == 1): the remainder m-group (one live m-tile) does not sweep its own
  // contiguous tile ranges -- that reads the whole index from HBM a second time -- but walks the tiles of the FIRST group's
  // regular items in the order those items reach them (unit q = (round, lead item): tile = lead item's first tile + round;
  // the group's regular items take the units round-robin), so that its reads find the tiles in L2.  Its late items split the
  // tail [n_reg_lead, n_j) that the leader's late items cover.
  int follow;
  int n_early;                       // items [0, n_early) are regular, [n_early, grid) late
  int group_first_late[MAX_GROUPS];
  int group_late[MAX_GROUPS];
  int late_share;                    // 0..256
  int group_mt[MAX_GROUPS];        // valid m-tiles in the group
  int n_j;                         // tiles visited: tile = tile_base + j * tile_step, j in [0, n_j)
  int tile_step;
  int64_t N;                       // index rows
  int Qp;                          // padded query rows
  int64_t global_offset;
  const float* tau;                // [Qp]        (FILTER)
  const float* margin;             // [Qp]        (FILTER) 2*eps per row, for the in-kernel threshold tightening
  int k;
  float* tilemax;                  // [n_j][Qp]   (TILEMAX)
  float* cand_s;                   // [items][mtg*128][CAND_CAP]
  int32_t* cand_i;
  int32_t* cand_cnt;               // [items][mtg*128]
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
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B smem matrix descriptor (cute::UMMA::SmemDescriptor): 8-row x 128-byte
// swizzle atoms, SBO = 1024 B between 8-row groups, LBO unused (1), version 1 (sm_100).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=fp16, both K-major, M x N
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // D=f32, A=B=f16
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

// ------------------------------------------------------------------ the scan
// grid = number of work items; item -> (group g, range r of the j loop).
// warps 0..7: epilogue (warp e: m-tile e/4, TMEM lane quarter e%4), warp 8: MMA issuer,
// warp 9: TMA producer (index tiles), warp 10: TMEM allocator.  The two single-thread roles sit on
// the HIGHEST warp ids: the SM sub-partition arbiter favours high warp ids, and they are the critical path.
// The fp16 queries (A operand) live in TMEM ([lane = row][col = 2 consecutive k]); every byte of
// shared memory is the TMA ring of index tiles, so ~200 KB of loads are in flight per SM.
#define TMEM_ST_32(taddr, v)                                                                                        \
  asm volatile(                                                                                                     \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                               \
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29," \
      "%30,%31,%32};" ::"r"(taddr),                                                                                 \
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),   \
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),    \
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),    \
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])                                                                \
      : "memory")

// D[tmem] (+)= A[tmem] * B[smem desc]   (A from tensor memory, "TS" form)
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc given as (lo, hi) words]: only the low word (address field) varies per MMA
__device__ __forceinline__ void tc_mma_f16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 bd;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 bd, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Padded row p = g * gstride + l  <->  query row g * rpg + l  (l < rpg): -1 for the padding rows of a group
__device__ __forceinline__ int padded_row_to_query(int p, int gstride, int rpg, int Q) {
  const int g = p / gstride, l = p - g * gstride;
  const int r = g * rpg + l;
  return (l < rpg && r < Q) ? r : -1;
}

// KBT / TNT: compile-time E/64 and tile width (0 = take them from the params at run time)
// A row's candidate list is about to run out of slots (rare: the sampled threshold was loose for this row).
// Move its k best to the front, tighten the row's threshold to (k-th best approximate score - 2 eps) -- still a
// lower bound of what any true top-k member scores in fp16 -- and drop everything below.  cnt = CAND_CAP + 1 is
// the overflow sentinel (-> exact fallback) when that frees no room (mass ties).
struct Compacted { int cnt; float thr; };
__device__ __noinline__ Compacted compact_candidates(float* s, int32_t* id, int cnt, int k, float mg, float thr) {
  Compacted out;
  out.thr = thr;
  out.cnt = CAND_CAP + 1;
  if (k > CAND_CAP / 4) return out;
  for (int r = 0; r < k; ++r) {
    int best = r;
    float bs = s[r];
    for (int x = r + 1; x < cnt; ++x) { float v = s[x]; if (v > bs) { bs = v; best = x; } }
    if (best != r) { float t = s[r]; s[r] = bs; s[best] = t; int32_t ti = id[r]; id[r] = id[best]; id[best] = ti; }
  }
  const float nt = s[k - 1] - mg;
  if (nt > thr) thr = nt;
  int w = k;
  for (int x = k; x < cnt; ++x) {
    float v = s[x];
    if (v > thr) { s[w] = v; id[w] = id[x]; ++w; }
  }
  out.thr = thr;
  if (w <= CAND_CAP - 64) out.cnt = w;
  return out;
}

// max of 32 fp32 accumulator values held as raw bits
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32]) {
  float m[8];
#pragma unroll
  for (int o = 0; o < 8; ++o)
    m[o] = fmaxf(fmaxf(__uint_as_float(v[o * 4 + 0]), __uint_as_float(v[o * 4 + 1])),
                 fmaxf(__uint_as_float(v[o * 4 + 2]), __uint_as_float(v[o * 4 + 3])));
  return fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
}

// predicated (branch-free) append of one value: no divergence stack traffic inside the candidate path
__device__ __forceinline__ int append_if(float s, float thr, float* cs, int32_t* ci, int cnt, int32_t id) {
  uint32_t took;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.gt.f32 p, %1, %2;\n"
      "@p st.global.f32 [%3], %1;\n"
      "@p st.global.s32 [%4], %5;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(took)
      : "f"(s), "f"(thr), "l"(cs + cnt), "l"(ci + cnt), "r"(id)
      : "memory");
  return cnt + (int)took;
}

// append every value of the 32-wide chunk that beats thr to the row's candidate list; one 8-wide octet at a time,
// only the octets whose maximum beats thr (thread == query row, so lanes diverge only at octet granularity)
__device__ __forceinline__ int chunk_filter(const uint32_t (&v)[32], float thr, float* cs, int32_t* ci, int cnt, int32_t id0) {
  float mq[4];
#pragma unroll
  for (int qd = 0; qd < 4; ++qd) {
    float a = fmaxf(__uint_as_float(v[qd * 8 + 0]), __uint_as_float(v[qd * 8 + 1]));
    float b = fmaxf(__uint_as_float(v[qd * 8 + 2]), __uint_as_float(v[qd * 8 + 3]));
    float c = fmaxf(__uint_as_float(v[qd * 8 + 4]), __uint_as_float(v[qd * 8 + 5]));
    float d = fmaxf(__uint_as_float(v[qd * 8 + 6]), __uint_as_float(v[qd * 8 + 7]));
    mq[qd] = fmaxf(fmaxf(a, b), fmaxf(c, d));
  }
  const float m = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
  if (m > thr) {
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
      if (mq[qd] > thr) {
#pragma unroll
        for (int i = qd * 8; i < qd * 8 + 8; ++i) cnt = append_if(__uint_as_float(v[i]), thr, cs, ci, cnt, id0 + i);
      }
    }
  }
  return cnt;
}

// ---- programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start once every block of its predecessor has executed launch_dependents (or exited); it must execute wait before it
// touches anything the predecessor writes.  Both are no-ops in a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- cluster helpers (multicast variant)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar), "r"(rank)
      : "memory");
}
// wait on a local mbarrier whose arrivals come from other CTAs of the cluster
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, uint32_t backoff_ns = 32) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
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
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// MODE: tile maxima (sample pass) or threshold filter.  KBT / TNT: compile-time E/64 and tile width (0 = run-time).
// ACC1: accumulator scheme.  false: one [mt_count x TN] accumulator set per tile, double-buffered over tiles (one commit
// per tile).  true (two m-tiles, TN = 128): ONE accumulator per m-tile; the MMAs of m-tile 1 run while the epilogue
// drains m-tile 0 and vice versa (one commit per m-tile), which buys 128-column tiles -- half the MMA instructions and
// commits per flop -- inside the same 512 TMEM columns.
template <int MODE, int KBT, int TNT, bool ACC1>
__global__ void __launch_bounds__(SCAN_THREADS, 1)
scan_kernel(const __grid_constant__ CUtensorMap tmap_idx, const __grid_constant__ ScanParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- which item: cluster `cid` = (super-group sg, tile range r_in_g); CTA `rank` of the cluster owns m-group sg*cs+rank
  const int item = blockIdx.x;
  const int cs = P.cs;
  const uint32_t rank = cs > 1 ? cluster_ctarank() : 0u;
  int g, r_in_g, R;
  bool late = false;
  int R_late = 0;
  if (cs > 1) {
    const int cid = item / cs;
    R = P.R;
    const int sg = cid / R;
    r_in_g = cid - sg * R;
    g = sg * cs + (int)rank;
  } else if (item < P.n_early) {
    g = 0;
    while (g + 1 < P.n_groups && item >= P.group_first_item[g + 1]) ++g;
    r_in_g = item - P.group_first_item[g];
    R = P.group_items[g];
    R_late = P.group_late[g];
  } else {
    late = true;
    g = 0;
    while (g + 1 < P.n_groups && item >= P.group_first_late[g + 1]) ++g;
    r_in_g = item - P.group_first_late[g];
    R = P.group_items[g];
    R_late = P.group_late[g];
  }
  const int mt_count = P.group_mt[g];
  // tile range of this item: the group's tiles are split in proportion to the item weights (regular 256, late late_share)
  int j0, j1;
  int g2 = 0, k0 = 0, k1 = 0;         // second segment (spanning items only): group g2, tiles [k0, k1)
  int slot = item, slot2 = 0;         // candidate slots of the two segments
  if (P.span) {
    const int64_t U = (int64_t)P.n_groups * P.n_j;
    const int64_t u0 = U * item / gridDim.x, u1 = U * (item + 1) / gridDim.x;
    g = (int)(u0 / P.n_j);
    j0 = (int)(u0 - (int64_t)g * P.n_j);
    const int64_t gend = (int64_t)(g + 1) * P.n_j;
    j1 = (int)(min(u1, gend) - (int64_t)g * P.n_j);
    if (u1 > gend) { g2 = g + 1; k0 = 0; k1 = (int)(u1 - gend); }
    // slot numbering: one slot per (item, segment), in item order; every group boundary that falls strictly inside an
    // earlier item (or this one) adds a slot
    int extra = 0;
    for (int b = 1; b < P.n_groups; ++b) {
      const int64_t ub = (int64_t)b * P.n_j;
      const int ii = (int)((ub * gridDim.x + gridDim.x - 1) / U) - 1;        // candidate item containing unit ub
      for (int c = max(0, ii - 1); c <= min((int)gridDim.x - 1, ii + 1); ++c) {
        const int64_t c0 = U * c / gridDim.x, c1 = U * (c + 1) / gridDim.x;
        if (c0 < ub && ub < c1 && c < item) ++extra;
      }
    }
    slot = item + extra;
    slot2 = slot + 1;
    r_in_g = 0; R = 1;
  } else {
    const int64_t wsum = (int64_t)R * 256 + (int64_t)R_late * P.late_share;
    const int64_t w0 = late ? (int64_t)R * 256 + (int64_t)r_in_g * P.late_share : (int64_t)r_in_g * 256;
    const int64_t w1 = w0 + (late ? P.late_share : 256);
    j0 = (int)(((int64_t)P.n_j * w0) / wsum);
    j1 = (int)(((int64_t)P.n_j * w1) / wsum);
  }
  // following group (see ScanParams::follow): leader = group 0; lead item i owns tiles [lead_j0(i), lead_j0(i + 1))
  const bool follower = MODE == MODE_FILTER && P.follow && g == P.n_groups - 1 && !P.span && cs == 1;
  const int lead_R = P.group_items[0];
  const int64_t lead_wsum = (int64_t)lead_R * 256 + (int64_t)P.group_late[0] * P.late_share;
  auto lead_j0 = [&](int i) { return (int)(((int64_t)P.n_j * ((int64_t)i * 256)) / lead_wsum); };
  const int n_reg_lead = lead_j0(lead_R);
  int f_rounds = 0;
  if (follower) {
    if (!late) {
      // rounds = longest lead range (ranges differ by at most one tile); units of the last round that do not exist are
      // scanned as an out-of-range tile (TMA zero fill, every column masked)
      f_rounds = n_reg_lead / lead_R + 1;
      const int64_t U = (int64_t)f_rounds * lead_R;
      j0 = 0;
      j1 = (int)((U - r_in_g + R - 1) / R);          // units r_in_g, r_in_g + R, ... < U
    } else {
      const int tail = P.n_j - n_reg_lead;
      j0 = n_reg_lead + (int)(((int64_t)tail * r_in_g) / max(R_late, 1));
      j1 = n_reg_lead + (int)(((int64_t)tail * (r_in_g + 1)) / max(R_late, 1));
    }
  }
  const int na = j1 - j0;             // tiles of the first segment
  // fused scan: this item first visits its share of the SAMPLE tiles (tile maxima), then -- after the per-group barrier
  // and the threshold selection in the epilogue -- its share of all tiles (filter)
  const int js0 = (MODE == MODE_FUSED && !late) ? (int)(((int64_t)P.n_s * r_in_g) / R) : 0;
  const int js1 = (MODE == MODE_FUSED && !late) ? (int)(((int64_t)P.n_s * (r_in_g + 1)) / R) : 0;
  const int ns_loc = js1 - js0;
  const int n_loc = ns_loc + na + (k1 - k0);
  // (32-bit arithmetic: every role evaluates this once per tile, and 64-bit divisions cost the epilogue ~600 cycles per tile;
  // the host enables `follow` only when n_j * items * 256 fits)
  const uint32_t lead_wsum32 = (uint32_t)lead_wsum, lead_nj256 = (uint32_t)P.n_j * 256u;
  auto tile_of = [&](int jj) {
    if (follower && !late) {
      const uint32_t q = (uint32_t)jj * (uint32_t)R + (uint32_t)r_in_g;
      const uint32_t rd = q / (uint32_t)lead_R, i = q - rd * (uint32_t)lead_R;
      const uint32_t t0 = (lead_nj256 * i) / lead_wsum32, t1 = (lead_nj256 * (i + 1)) / lead_wsum32;
      return (int)(t0 + rd < t1 ? t0 + rd : (uint32_t)P.n_j);          // P.n_j = one past the last tile: out of range
    }
    return jj < ns_loc ? (js0 + jj) * P.s_step : (jj - ns_loc < na ? (j0 + jj - ns_loc) * P.tile_step : (k0 + jj - ns_loc - na) * P.tile_step);
  };
  const int KB = KBT ? KBT : P.kb, NG = P.n_stages, TN = TNT ? TNT : P.tn;   // n_stages = number of TILE slots in the ring
  const int E = KB * KBLK;
  const uint32_t kb_bytes = (uint32_t)TN * KBLK * 2;     // one [TN x 64] fp16 SW128 sub-tile
  const uint32_t slot_bytes = kb_bytes * KB;             // one whole index tile [TN x E]

  uint8_t* b_smem = smem;                                // [NG] slots x [KB] sub-tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + (size_t)NG * slot_bytes);
  // bars: full[NG], empty[NG], a_full, acc_full[2], acc_empty[2], cl_empty[NG]
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + NG);
  const uint32_t bar_a = smem_u32(bars + 2 * NG);
  const uint32_t bar_accf = smem_u32(bars + 2 * NG + 1);
  const uint32_t bar_acce = smem_u32(bars + 2 * NG + 3);
  const uint32_t bar_cle = smem_u32(bars + 2 * NG + 5);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * NG + 5);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NG; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); mbar_init(bar_cle + 8 * s, (uint32_t)cs); }
    mbar_init(bar_a, mt_count * 4);
    for (int b = 0; b < 2; ++b) { mbar_init(bar_accf + 8 * b, 1); mbar_init(bar_acce + 8 * b, (ACC1 && mt_count > 1) ? 4 : mt_count * 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 10) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();     // every CTA's barriers exist before a peer arrives on them / multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_col0 = (uint32_t)P.a_cols;          // accumulators sit behind the query columns
  pdl_launch_dependents();                               // (PDL) the next kernel of the search may be scheduled behind this grid
  // accumulator scheme of THIS item: ACC1 pairs the two m-tiles (one accumulator each); an item with a single live m-tile
  // (the remainder group of a fully packed batch) would have nothing to overlap with, so it alternates between the two
  // accumulators over tiles instead (double buffer), exactly like the !ACC1 scheme with one m-tile
  const bool pair_acc = ACC1 && mt_count > 1;
  const int acc_mtg = ACC1 ? 1 : P.mtg;                  // m-tiles per double-buffer half in the !pair_acc addressing

  if (warp == 9) {
    // ===== TMA producer.  Every CTA arms its own full barrier for the slot once its consumers released it and tells the
    // cluster leader; the leader issues ONE multicast load per tile when all cs CTAs are ready (cs == 1: plain load).
    if (n_loc > 0) {     // whole warp, converged; one elected lane issues
      long long w_empty = 0, w_cl = 0, t_begin = clock64();
      const uint16_t mask = (uint16_t)((1u << cs) - 1u);
      for (int jj = 0; jj < n_loc; ++jj) {
        const uint32_t s = (uint32_t)jj % NG, ph = ((uint32_t)jj / NG) & 1;
        const int tile = tile_of(jj);
        mbar_wait_timed(bar_empty + 8 * s, ph ^ 1, w_empty, 64);
        if (elect_one_sync()) {
          if (P.dbg_flags & 1) {
            mbar_arrive(bar_full + 8 * s);
          } else {
            mbar_expect_tx(bar_full + 8 * s, slot_bytes);
            if (cs > 1) mbar_arrive_remote(bar_cle + 8 * s, 0);
          }
        }
        __syncwarp();
        if (!(P.dbg_flags & 1) && rank == 0) {
          if (cs > 1) { long long t = clock64(); mbar_wait_cluster(bar_cle + 8 * s, ph, 64); w_cl += clock64() - t; }
          if (elect_one_sync()) {
            const uint32_t dst = smem_u32(b_smem + (size_t)s * slot_bytes);
            if (P.use3d) {
              if (cs > 1) tma_load_3d_mc(dst, &tmap_idx, bar_full + 8 * s, 0, tile * TN, 0, mask);
              else tma_load_3d(dst, &tmap_idx, bar_full + 8 * s, 0, tile * TN, 0);
            } else {
              for (int kb = 0; kb < KB; ++kb) {
                if (cs > 1) tma_load_2d_mc(dst + (uint32_t)kb * kb_bytes, &tmap_idx, bar_full + 8 * s, kb * KBLK, tile * TN, mask);
                else tma_load_2d(dst + (uint32_t)kb * kb_bytes, &tmap_idx, bar_full + 8 * s, kb * KBLK, tile * TN);
              }
            }
          }
          __syncwarp();
        }
      }
      if (P.dbg && lane == 0) { P.dbg[item * DBG_N + 0] = w_empty; P.dbg[item * DBG_N + 1] = clock64() - t_begin; P.dbg[item * DBG_N + 11] = w_cl; }
    }
  } else if (warp == 8) {
    // ===== MMA issuer (one thread): all MMAs of an accumulator back to back, ONE tcgen05.commit per accumulator =====
    if (n_loc > 0) {     // whole warp, converged; one elected lane issues
      const uint32_t idesc = make_idesc_f16(TILE_M, TN);
      long long w_full = 0, w_acce = 0, w_a = 0, t_begin = clock64();
      mbar_wait_timed(bar_a, 0, w_a);
      tc_fence_after();
      for (int jj = 0; jj < n_loc; ++jj) {
        if (P.span && k1 > k0 && jj == na) {       // second segment: the epilogue has restaged the queries of group g2
          mbar_wait_timed(bar_a, 1, w_a);
          tc_fence_after();
        }
        const int buf = jj & 1;
        const uint32_t use = (uint32_t)(jj >> 1);
        const uint32_t s = (uint32_t)jj % NG, ph = ((uint32_t)jj / NG) & 1;
        if (!pair_acc) mbar_wait_timed(bar_acce + 8 * buf, (use & 1) ^ 1, w_acce);
        mbar_wait_timed(bar_full + 8 * s, ph, w_full);
        const uint64_t bd0 = make_sw128_desc(smem_u32(b_smem + (size_t)s * slot_bytes));
        const uint32_t bd_lo = (uint32_t)bd0, bd_hi = (uint32_t)(bd0 >> 32);
        const uint32_t kb_units = kb_bytes >> 4;           // descriptor address units (16 B) per k-block
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (mt < mt_count) {
            if (pair_acc) mbar_wait_timed(bar_acce + 8 * mt, ((uint32_t)jj & 1) ^ 1, w_acce);
            tc_fence_after();
            if (elect_one_sync()) {
              if (!(P.dbg_flags & 2)) {
                const uint32_t d = tmem_base + acc_col0 + (uint32_t)(pair_acc ? mt * TN : (buf * acc_mtg + mt) * TN);
                const uint32_t a0 = tmem_base + (uint32_t)(mt * (E / 2));
                if (KBT > 0) {
#pragma unroll
                  for (int kb = 0; kb < (KBT > 0 ? KBT : 1); ++kb)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)   // 4 x (K=16): A advances 8 columns, B 32 bytes (= 2 descriptor units)
                      tc_mma_f16_ts2(d, a0 + (uint32_t)(kb * (KBLK / 2) + 8 * k4), bd_lo + (uint32_t)kb * kb_units + 2 * k4, bd_hi, idesc,
                                     (kb | k4) ? 1u : 0u);
                } else {
                  for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                      tc_mma_f16_ts2(d, a0 + (uint32_t)(kb * (KBLK / 2) + 8 * k4), bd_lo + (uint32_t)kb * kb_units + 2 * k4, bd_hi, idesc,
                                     (kb | k4) ? 1u : 0u);
                }
              }
              if (pair_acc) tc_commit(bar_accf + 8 * mt);
            }
            __syncwarp();
          }
        }
        if (!pair_acc) {
          if (elect_one_sync()) tc_commit(bar_accf + 8 * buf);      // accumulators complete == this ring slot fully read
          __syncwarp();
        }
      }
      if (P.dbg && lane == 0) { P.dbg[item * DBG_N + 2] = w_full; P.dbg[item * DBG_N + 3] = w_acce; P.dbg[item * DBG_N + 4] = clock64() - t_begin; P.dbg[item * DBG_N + 5] = w_a; }
    }
  } else if (warp < 8) {
    // ===== epilogue: thread == query row =====
    const int e = warp;
    const int mt = e >> 2, quarter = e & 3;
    if (mt < mt_count && n_loc > 0) {
     const int nseg = (P.span && k1 > k0) ? 2 : 1;
     for (int seg = 0; seg < nseg; ++seg) {       // spanning items: second pass for the tiles that belong to the next m-group
      const int gs = seg ? g2 : g;
      const int slot_s = seg ? slot2 : slot;
      const int jj_lo = seg ? ns_loc + na : 0, jj_hi = seg ? n_loc : ns_loc + na;
      const int lrow = mt * TILE_M + quarter * 32 + lane;             // row within the group
      const int grow = gs * P.mtg * TILE_M + lrow;                    // padded global row
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
      float mg_row = 0.f;                                             // fused: 2*eps of this row
      bool live_row = true;
      // ---- stage this row's fp16 query into TMEM (A operand): 2 consecutive k per 32-bit column
      if (MODE == MODE_FUSED) {
        // straight from the fp32 queries: scale by a power of two so that max |q_i| lands in [0.5, 1) (see
        // prep_queries_kernel), convert, and keep the row's margin in a register
        const int qr = padded_row_to_query(grow, P.gstride, P.rpg, P.Q);
        live_row = qr >= 0;
        const int Et = P.E_true;
        const float* src = P.q32 + (size_t)(live_row ? qr : 0) * Et;
        const bool vec = (Et & 3) == 0;
        float mx = 0.f;
        if (live_row) {
          if (vec) for (int j = 0; j < Et; j += 4) { const float4 u = __ldg(reinterpret_cast<const float4*>(src + j)); mx = fmaxf(fmaxf(mx, fmaxf(fabsf(u.x), fabsf(u.y))), fmaxf(fabsf(u.z), fabsf(u.w))); }
          else for (int j = 0; j < Et; ++j) mx = fmaxf(mx, fabsf(__ldg(src + j)));
        }
        float scale = 1.f;
        if (mx > 0.f && mx < CUDART_INF_F) {
          int ex;
          frexpf(mx, &ex);
          ex = max(-100, min(100, ex));
          scale = ldexpf(1.f, -ex);
        }
        float ss = 0.f;
        const uint32_t a_t = lane_base + (uint32_t)(mt * (E / 2));
        for (int c = 0; c < E / 2; c += 32) {
          uint32_t v[32];
#pragma unroll
          for (int q4 = 0; q4 < 16; ++q4) {
            const int j = 2 * c + 4 * q4;
            float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live_row) {
              if (vec) { if (j < Et) u = __ldg(reinterpret_cast<const float4*>(src + j)); }
              else { if (j < Et) u.x = __ldg(src + j); if (j + 1 < Et) u.y = __ldg(src + j + 1); if (j + 2 < Et) u.z = __ldg(src + j + 2); if (j + 3 < Et) u.w = __ldg(src + j + 3); }
            }
            const __half2 h0 = __floats2half2_rn(u.x * scale, u.y * scale), h1 = __floats2half2_rn(u.z * scale, u.w * scale);
            ss = fmaf(u.x * scale, u.x * scale, ss); ss = fmaf(u.y * scale, u.y * scale, ss); ss = fmaf(u.z * scale, u.z * scale, ss); ss = fmaf(u.w * scale, u.w * scale, ss);
            v[q4 * 2 + 0] = *reinterpret_cast<const uint32_t*>(&h0);
            v[q4 * 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
          }
          TMEM_ST_32(a_t + c, v);
        }
        mg_row = live_row ? fmaxf(2.f * EPS_REL * sqrtf(ss) * __ldg(P.tnorm_max), 1e-20f) : 0.f;
        if (r_in_g == 0) P.margin_out[grow] = mg_row;
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a);
      } else {
        if (MODE == MODE_TILEMAX) pdl_wait();                         // qb comes from prep_queries, the kernel right before the sample pass
        const uint4* src = reinterpret_cast<const uint4*>(P.qb + (size_t)grow * E);
        const uint32_t a_t = lane_base + (uint32_t)(mt * (E / 2));
        for (int c = 0; c < E / 2; c += 32) {
          uint32_t v[32];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            uint4 u = __ldg(src + (c / 4) + q4);
            v[q4 * 4 + 0] = u.x; v[q4 * 4 + 1] = u.y; v[q4 * 4 + 2] = u.z; v[q4 * 4 + 3] = u.w;
          }
          TMEM_ST_32(a_t + c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a);
      }
      float thr = CUDART_INF_F;
      int cnt = 0;
      float* my_s = nullptr;
      int32_t* my_i = nullptr;
      if (MODE != MODE_TILEMAX) {
        if (MODE == MODE_FILTER) {
          // (PDL) everything above -- TMEM allocation, barrier set-up, query staging, the first index tiles in flight -- may
          // have run while select_tau was still executing; tau / margin are its outputs
          pdl_wait();
          thr = P.tau[grow];
        }
        size_t base = ((size_t)slot_s * (P.mtg * TILE_M) + lrow) * CAND_CAP;
        my_s = P.cand_s + base;
        my_i = P.cand_i + base;
      }
      const int n_chunks = TN / 32;
      // the warp that hands the ring slot back: all MMAs of the tile have retired once the LAST accumulator's commit fired
      const bool releases_slot = pair_acc ? (e == 4 * (mt_count - 1)) : (e == 0);
      long long w_accf = 0, w_ld = 0, w_cmp = 0, t_begin = clock64();
      for (int jj = jj_lo; jj < jj_hi; ++jj) {
        if (MODE == MODE_FUSED && jj == ns_loc) {
          // ---- all sample tiles of this item are done: per-group barrier (every epilogue warp of every item of the group
          // arrives once), then this row's threshold = k-th largest of ITS sampled tile maxima - 2 eps
          __threadfence();
          __syncwarp();
          if (lane == 0) {
            if (!late) atomicAdd(P.group_ctr + g, 1u);
            const unsigned target = (unsigned)(R * mt_count * 4);
            long long t0w = 0;
            for (unsigned spins = 0;; ++spins) {
              unsigned seen;
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(P.group_ctr + g) : "memory");
              if (seen >= target) break;
              __nanosleep(64);
              if ((spins & 0xfff) == 0xfff) {
                long long now = clock64();
                if (t0w == 0) t0w = now;
                else if (now - t0w > 4000000000LL) __trap();
              }
            }
          }
          __syncwarp();
          if (live_row) {
            float top[FUSED_MAX_K];
#pragma unroll
            for (int x = 0; x < FUSED_MAX_K; ++x) top[x] = -CUDART_INF_F;
            const float* tmr = P.tilemax + (size_t)grow * P.n_s;
            const int kk = P.k;
            auto offer = [&](float v) {
              if (v > top[FUSED_MAX_K - 1]) {
#pragma unroll
                for (int y = 0; y < FUSED_MAX_K; ++y) {          // sorted insertion (descending)
                  const float hi = fmaxf(top[y], v);
                  v = fminf(top[y], v);
                  top[y] = hi;
                }
              }
            };
            // the row's n_s sampled maxima, 64 values per batch of 16 independent 16-byte loads (the row was written by
            // other SMs: L2 reads; a one-value-at-a-time walk pays the L2 latency n_s times -- measured +70 us)
            const bool vec4 = (P.n_s & 3) == 0 && ((reinterpret_cast<uintptr_t>(tmr) & 15) == 0);
            int x = 0;
            if (vec4) {
              for (; x + 64 <= P.n_s; x += 64) {
                float4 b[16];
#pragma unroll
                for (int y = 0; y < 16; ++y) b[y] = __ldcg(reinterpret_cast<const float4*>(tmr + x) + y);
#pragma unroll
                for (int y = 0; y < 16; ++y) { offer(b[y].x); offer(b[y].y); offer(b[y].z); offer(b[y].w); }
              }
              for (; x + 4 <= P.n_s; x += 4) {
                const float4 b = __ldcg(reinterpret_cast<const float4*>(tmr + x));
                offer(b.x); offer(b.y); offer(b.z); offer(b.w);
              }
            }
            for (; x < P.n_s; ++x) offer(__ldcg(tmr + x));
            float kth = -CUDART_INF_F;
#pragma unroll
            for (int y = 0; y < FUSED_MAX_K; ++y)
              if (y == kk - 1) kth = top[y];
            thr = kth - mg_row;
            if (r_in_g == 0 && P.tau_out) P.tau_out[grow] = thr;
          } else {
            thr = CUDART_INF_F;
          }
        }
        const bool sampling = MODE == MODE_TILEMAX || (MODE == MODE_FUSED && jj < ns_loc);
        const int buf = pair_acc ? mt : (jj & 1);
        const uint32_t use = pair_acc ? (uint32_t)jj : (uint32_t)(jj >> 1);
        const int tile = tile_of(jj);
        const int64_t col0 = (int64_t)tile * TN;
        const bool ragged = col0 + TN > P.N;
        mbar_wait_timed(bar_accf + 8 * buf, use & 1, w_accf);
        tc_fence_after();
        // the MMAs of this tile have retired: hand its ring slot back to the TMA producer
        if (releases_slot && lane == 0) mbar_arrive(bar_empty + 8 * ((uint32_t)jj % NG));
        const uint32_t taddr = lane_base + acc_col0 + (uint32_t)(pair_acc ? mt * TN : (buf * acc_mtg + mt) * TN);
        float tmax = -CUDART_INF_F;
        // 64 accumulator columns per round: both TMEM loads in flight together; after the last round's data has
        // landed in registers the accumulator buffer goes straight back to the MMA warp, and the compares run
        // from registers while the next MMAs already overwrite it.
#pragma unroll 1
        for (int ch = 0; ch < n_chunks; ch += 2) {
          uint32_t va[32], vb[32];
          long long t_c0 = 0;
          if (P.dbg) t_c0 = clock64();
          TMEM_LD_32(taddr + ch * 32, va);
          TMEM_LD_32(taddr + ch * 32 + 32, vb);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (ch + 2 >= n_chunks) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce + 8 * buf);
          }
          if (P.dbg) { long long t = clock64(); w_ld += t - t_c0; t_c0 = t; }
          const int64_t cbase = col0 + ch * 32;
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (cbase + i >= P.N) va[i] = 0xff800000u;   // -inf
              if (cbase + 32 + i >= P.N) vb[i] = 0xff800000u;
            }
          }
          if (MODE == MODE_TILEMAX || (MODE == MODE_FUSED && sampling)) {
            tmax = fmaxf(tmax, fmaxf(chunk_max(va), chunk_max(vb)));
          } else {
            const int32_t id0 = (int32_t)(P.global_offset + cbase);
            cnt = chunk_filter(va, thr, my_s, my_i, cnt, id0);
            cnt = chunk_filter(vb, thr, my_s, my_i, cnt, id0 + 32);
            // keep >= 64 free slots for the next round
            if (cnt > CAND_CAP - 64) {
              Compacted cc = compact_candidates(my_s, my_i, cnt, P.k, MODE == MODE_FUSED ? mg_row : P.margin[grow], thr);
              cnt = cc.cnt;
              thr = cc.cnt > CAND_CAP ? CUDART_INF_F : cc.thr;   // overflow sentinel: stop collecting, finalize flags the row
            }
          }
          if (P.dbg) w_cmp += clock64() - t_c0;
        }
        if (MODE == MODE_TILEMAX) P.tilemax[(size_t)(tile_of(jj) / P.tile_step) * P.Qp + grow] = tmax;
        if (MODE == MODE_FUSED && sampling) P.tilemax[(size_t)grow * P.n_s + js0 + jj] = tmax;      // row-major: the selection reads one row
      }
      if (MODE == MODE_FUSED && n_loc == ns_loc && !late) {          // an item without filter tiles still owes the group its arrival
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(P.group_ctr + g, 1u);
      }
      if (MODE != MODE_TILEMAX) P.cand_cnt[(size_t)slot_s * (P.mtg * TILE_M) + lrow] = cnt;
      if (P.dbg && e == 0 && lane == 0) { P.dbg[item * DBG_N + 6] = w_accf; P.dbg[item * DBG_N + 7] = clock64() - t_begin; P.dbg[item * DBG_N + 8] = w_ld; P.dbg[item * DBG_N + 9] = w_cmp; P.dbg[item * DBG_N + 10] = cnt; }
     }
    } else if (mt < mt_count && MODE != MODE_TILEMAX) {
      P.cand_cnt[(size_t)item * (P.mtg * TILE_M) + mt * TILE_M + quarter * 32 + lane] = 0;
      if (MODE == MODE_FUSED && lane == 0 && !late) atomicAdd(P.group_ctr + g, 1u);      // no tiles at all: arrive, never wait
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();     // no CTA leaves while a peer may still arrive on its barriers
  if (warp == 10) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// --------------------------------------------------------------- small kernels
// q fp32 [Q,E] -> fp16 [Qp,E] (rows >= Q zero), qnorm[Qp]
// Padded row p = g * gstride + l  <->  query row g * rpg + l  (l < rpg): every m-group holds the same number
// of query rows, so all groups cost the same and sweep the index in lock-step (their re-reads hit L2).
__device__ __forceinline__ int padded_to_query_row(int p, int gstride, int rpg, int Q) {
  const int g = p / gstride, l = p - g * gstride;
  const int r = g * rpg + l;
  return (l < rpg && r < Q) ? r : -1;
}

// q fp32 [Q,E] -> fp16 [Qp,E] in padded row order (unused rows zero), qnorm[Qp].
// Every row is scaled by an exact power of two so that its largest element lands in [0.5, 1) before the fp16
// conversion: un-normalised sources (sse_demo.py:123) whose elements exceed the fp16 range (65504) or sit in its
// subnormal range keep full fp16 relative precision.  The scan works in the scaled units throughout (qnorm, tau,
// margin and the candidates' approximate scores all carry the same factor; finalize re-scores in fp32 from the
// original q), so the scaling never reaches the results.
__global__ void prep_queries_kernel(const float* __restrict__ q, int Q, int Qp, int E, int Ep, int gstride, int rpg,
                                    __half* __restrict__ qb, float* __restrict__ qnorm, int32_t* __restrict__ ov_count) {
  pdl_launch_dependents();
  if (blockIdx.x == 0 && threadIdx.x == 0) *ov_count = 0;
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= Qp) return;
  const int qr = padded_to_query_row(row, gstride, rpg, Q);
  float mx = 0.f;
  if (qr >= 0)
    for (int j = lane; j < E; j += 32) mx = fmaxf(mx, fabsf(q[(size_t)qr * E + j]));
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float scale = 1.f;
  if (mx > 0.f && mx < CUDART_INF_F) {
    int ex;
    frexpf(mx, &ex);                       // mx = m * 2^ex, m in [0.5, 1)
    ex = max(-100, min(100, ex));
    scale = ldexpf(1.f, -ex);
  }
  float ss = 0.f;
  for (int j = lane; j < Ep; j += 32) {           // columns [E, Ep): zero padding up to the next multiple of 64
    float v = (qr >= 0 && j < E) ? q[(size_t)qr * E + j] * scale : 0.f;
    ss = fmaf(v, v, ss);
    qb[(size_t)row * Ep + j] = __float2half_rn(v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) qnorm[row] = sqrtf(ss);
}

// max row norm of the index (one-time, at index_set): out[0] = max |t|
__global__ void max_row_norm_kernel(const float* __restrict__ x, int64_t N, int E, float* __restrict__ out) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  float best = 0.f;
  for (; row < N; row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    float ss = 0.f;
    for (int j = lane; j < E; j += 32) { float v = x[(size_t)row * E + j]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    best = fmaxf(best, ss);
  }
  if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(best)));   // non-negative floats order as ints
}

// warp per row: tau[r] = (k-th largest of tilemax[0..n_s)[r]) - 2*eps_r ; rows >= Q: +inf.
// Each lane keeps its <= 32 strided samples in registers; k rounds of warp arg-max with removal.
__global__ void select_tau_kernel(const float* __restrict__ tilemax, int n_s, int Qp, int Q, int k, int gstride, int rpg,
                                  const float* __restrict__ qnorm, const float* __restrict__ tnorm_max,
                                  float* __restrict__ tau, float* __restrict__ margin) {
  pdl_launch_dependents();                 // the filter scan may set itself up (TMEM, staging, first tiles) while this runs
  pdl_wait();                              // tilemax is the sample pass's output
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= Qp) return;
  if (padded_to_query_row(row, gstride, rpg, Q) < 0) { if (lane == 0) { tau[row] = CUDART_INF_F; margin[row] = 0.f; } return; }
  float v[32];
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    int j = lane + 32 * t;
    v[t] = j < n_s ? __ldg(tilemax + (size_t)j * Qp + row) : -CUDART_INF_F;
  }
  float kth = -CUDART_INF_F;
  for (int r = 0; r < k; ++r) {
    float bs = v[0];
    int bt = 0;
#pragma unroll
    for (int t = 1; t < 32; ++t)
      if (v[t] > bs) { bs = v[t]; bt = t; }
    float ws = bs;
#pragma unroll
    for (int o = 16; o; o >>= 1) ws = fmaxf(ws, __shfl_xor_sync(0xffffffffu, ws, o));
    kth = ws;
    // exactly one lane (the lowest holding the max) retires its element
    unsigned m = __ballot_sync(0xffffffffu, bs == ws);
    if (lane == __ffs(m) - 1) {
#pragma unroll
      for (int t = 0; t < 32; ++t)
        if (t == bt) v[t] = -CUDART_INF_F;
    }
  }
  if (lane == 0) {
    float mg = fmaxf(2.f * EPS_REL * qnorm[row] * tnorm_max[0], 1e-20f);
    margin[row] = mg;
    tau[row] = kth - mg;
  }
}

struct FinParams {
  int n_groups, mtg;
  int rpg;                 // query rows per m-group
  int cs, R;               // cluster scan layout: item of (group g, range it) = ((g / cs) * R + it) * cs + g % cs
  int group_first_item[MAX_GROUPS];   // cs == 1: items of group g are [first, first + items)
  int group_items[MAX_GROUPS];
  int group_first_late[MAX_GROUPS];   // late items of group g (see ScanParams)
  int group_late[MAX_GROUPS];
  int span;                           // spanning items: the candidate slots of group g are [span_first[g], span_first[g] + span_n[g])
  int span_first[MAX_GROUPS];
  int span_n[MAX_GROUPS];
  const float* cand_s;
  const int32_t* cand_i;
  const int32_t* cand_cnt;
  const float* margin;
  const float* q;          // fp32 [Q,E]
  const float* index;      // fp32 [N,E] local shard
  int64_t global_offset;
  int64_t N;
  int Q, E, k;
  int maxc;                // candidate capacity of the finalize kernel's shared-memory lists
  int out_stride;          // row stride of out_s / out_i in elements (k, or 2k when both live in one packed [Q,2k] block)
  float* out_s;
  int32_t* out_i;
  int32_t* overflow;       // [Q] flags
  int32_t* ov_list;        // [Q] rows that need the brute-force fallback, ov_list[-1].. see ov_count
  int32_t* ov_count;       // number of entries of ov_list (zeroed at the start of every search by prep_queries / the fused scan)
  unsigned* group_ctr;     // fused scan: per-group barrier counters, re-armed (zeroed) here for the next search
};

// bitonic sort of n (power of two) pairs in smem, order (score desc, idx asc)
__device__ __forceinline__ bool pair_before(float sa, int32_t ia, float sb, int32_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}
__device__ void bitonic_sort_pairs(float* s, int32_t* id, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool up = ((lo & size) == 0);     // first-half blocks sorted "best first"
        float sa = s[lo], sb = s[hi];
        int32_t ia = id[lo], ib = id[hi];
        bool swap = up ? pair_before(sb, ib, sa, ia) : pair_before(sa, ia, sb, ib);
        if (swap) { s[lo] = sb; s[hi] = sa; id[lo] = ib; id[hi] = ia; }
      }
    }
  }
  __syncthreads();
}

constexpr int FIN_MAX_ITEMS = 512;     // work items one m-group can have (regular + late)

__global__ void __launch_bounds__(128) finalize_kernel(const __grid_constant__ FinParams P) {
  extern __shared__ float fin_dyn[];                   // cs [maxc] | ci [maxc] | tmp [maxc] | sv [FIN_MAXR]: maxc = 2048 for k <= 32 (25 KB: 8 rows per SM in flight), 4096 above
  const int FIN_MAXC = P.maxc;
  float* cs = fin_dyn;
  int32_t* ci = reinterpret_cast<int32_t*>(fin_dyn + FIN_MAXC);
  __shared__ int s_off[FIN_MAX_ITEMS + 1];             // prefix sums of the per-item candidate counts of this row
  __shared__ int s_over, s_m;
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  if (P.group_ctr && row == 0 && tid < P.n_groups) P.group_ctr[tid] = 0u;
  const int rows_per_group = P.mtg * TILE_M;
  const int g = row / P.rpg, lrow = row % P.rpg;
  const int R = P.span ? P.span_n[g] : (P.cs > 1 ? P.R : P.group_items[g]);
  const int first = P.span ? P.span_first[g] : (P.cs > 1 ? (g / P.cs) * P.R * P.cs + g % P.cs : P.group_first_item[g]), istride = P.span ? 1 : P.cs;
  const int R_l = (P.cs > 1 || P.span) ? 0 : P.group_late[g];
  const int n_it = min(R + R_l, FIN_MAX_ITEMS);
  auto slot_of = [&](int it) { return (size_t)(it < R ? first + it * istride : P.group_first_late[g] + (it - R)) * rows_per_group + lrow; };
  if (tid == 0) { s_over = (R + R_l > FIN_MAX_ITEMS) ? 1 : 0; s_off[0] = 0; }
  __syncthreads();
  // ---- gather: counts of every item in parallel, prefix sum, then ONE candidate per thread-iteration (all loads of the
  // block in flight together; a thread-per-item copy loop serialised 2-3 dependent L2 round trips per item)
  for (int it = tid; it < n_it; it += blockDim.x) {
    int c = P.cand_cnt[slot_of(it)];
    if (c > CAND_CAP) { atomicExch(&s_over, 1); c = CAND_CAP; }
    s_off[it + 1] = c;
  }
  __syncthreads();
  if (tid < 32) {                                      // warp scan over the (<= 512) counts
    int carry = 0;
    for (int base = 0; base < n_it; base += 32) {
      const int i = base + tid;
      int v = i < n_it ? s_off[i + 1] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (tid >= o) v += u; }
      if (i < n_it) s_off[i + 1] = v + carry;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  const int total_all = s_off[n_it];
  if (total_all > FIN_MAXC) { if (tid == 0) s_over = 1; }
  __syncthreads();
  if (s_over) {
    if (tid == 0) { P.overflow[row] = 1; P.ov_list[atomicAdd(P.ov_count, 1)] = row; }
    return;
  }
  if (tid == 0) P.overflow[row] = 0;
  const int total = total_all;
  for (int e = tid; e < total; e += blockDim.x) {
    int lo = 0, hi = n_it;                             // item with s_off[it] <= e < s_off[it + 1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= e) lo = mid; else hi = mid; }
    const size_t src = slot_of(lo) * CAND_CAP + (e - s_off[lo]);
    cs[e] = P.cand_s[src];
    ci[e] = P.cand_i[src];
  }
  // ---- k-th best APPROXIMATE score: k rounds of block-wide max extraction on a scratch copy (no full sort: a bitonic sort
  // of ~256 pairs is 36 barrier-separated stages; measured 2.6 k instructions per warp at 14 cycles each)
  float* tmp = reinterpret_cast<float*>(ci + FIN_MAXC);      // [maxc] scratch scores, later the exact scores of the survivors
  int32_t* sv = reinterpret_cast<int32_t*>(tmp + FIN_MAXC);  // [FIN_MAXR] survivor ids
  __shared__ float s_wv[4];
  __shared__ int s_we[4];
  __shared__ float s_kth;
  const int k = P.k;
  const int warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < total; e += blockDim.x) tmp[e] = cs[e];
  if (tid == 0) { s_m = 0; s_kth = -CUDART_INF_F; }
  __syncthreads();
  if (total >= k) {
    for (int r = 0; r < k; ++r) {
      float bs = -CUDART_INF_F;
      int be = 0x7fffffff;
      for (int e = tid; e < total; e += blockDim.x) { const float v = tmp[e]; if (v > bs) { bs = v; be = e; } }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int oe = __shfl_xor_sync(0xffffffffu, be, o);
        if (os > bs || (os == bs && oe < be)) { bs = os; be = oe; }
      }
      if (lane == 0) { s_wv[warp] = bs; s_we[warp] = be; }
      __syncthreads();
      float gs = s_wv[0];
      int ge = s_we[0];
#pragma unroll
      for (int w2 = 1; w2 < 4; ++w2) { const float os = s_wv[w2]; const int oe = s_we[w2]; if (os > gs || (os == gs && oe < ge)) { gs = os; ge = oe; } }
      if (tid == 0) { s_kth = gs; if (ge < total) tmp[ge] = -CUDART_INF_F; }
      __syncthreads();
    }
    // survivors: approx >= A_k - margin (order is irrelevant: the exact scores decide)
    const float cut = s_kth - P.margin[g * rows_per_group + lrow];
    for (int e = tid; e < total; e += blockDim.x)
      if (cs[e] >= cut) { const int pos = atomicAdd(&s_m, 1); if (pos < FIN_MAXR) sv[pos] = ci[e]; }
  } else {
    for (int e = tid; e < total; e += blockDim.x) { const int pos = atomicAdd(&s_m, 1); if (pos < FIN_MAXR) sv[pos] = ci[e]; }
  }
  __syncthreads();
  const int m = s_m;
  if (m > FIN_MAXR) {
    if (tid == 0) { P.overflow[row] = 1; P.ov_list[atomicAdd(P.ov_count, 1)] = row; }
    return;
  }
  // ---- exact fp32 re-score of the m survivors: 8 lanes per candidate, 16 candidates per pass (m is ~k + a few: one pass)
  {
    const int sub = tid & 7, grp = tid >> 3;
    const float* qr = P.q + (size_t)row * P.E;
    for (int c0 = 0; c0 < m; c0 += 16) {
      const int c = c0 + grp;
      float acc = 0.f;
      if (c < m) {
        const float* tr = P.index + (size_t)(sv[c] - P.global_offset) * P.E;
        if ((P.E & 3) == 0) {
          for (int j = sub * 4; j < P.E; j += 32) {
            const float4 a4 = *reinterpret_cast<const float4*>(qr + j);
            const float4 b4 = *reinterpret_cast<const float4*>(tr + j);
            acc = fmaf(a4.x, b4.x, acc); acc = fmaf(a4.y, b4.y, acc); acc = fmaf(a4.z, b4.z, acc); acc = fmaf(a4.w, b4.w, acc);
          }
        } else {                    // odd encoding sizes (the reference recipes use E = 50): rows are not 16-byte aligned
          for (int j = sub; j < P.E; j += 8) acc = fmaf(qr[j], tr[j], acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (c < m && sub == 0) tmp[c] = acc;
    }
  }
  __syncthreads();
  // ---- final order (exact score desc, id asc) by rank counting among the m survivors; rank r < k goes to output slot r
  for (int c = tid; c < m; c += blockDim.x) {
    const float sc = tmp[c];
    const int32_t ic = sv[c];
    int rank = 0;
    for (int j = 0; j < m; ++j) rank += pair_before(tmp[j], sv[j], sc, ic) ? 1 : 0;
    if (rank < k) {
      P.out_s[(size_t)row * P.out_stride + rank] = sc;
      P.out_i[(size_t)row * P.out_stride + rank] = ic;
    }
  }
  for (int x = m + tid; x < k; x += blockDim.x) {
    P.out_s[(size_t)row * P.out_stride + x] = -CUDART_INF_F;
    P.out_i[(size_t)row * P.out_stride + x] = -1;
  }
}

// rows flagged by finalize: brute-force exact fp32 top-k (rare: pathological ties / clustered index)
__global__ void __launch_bounds__(256) fallback_kernel(const __grid_constant__ FinParams P) {
  extern __shared__ float dyn[];   // per warp k scores + k idx
  const int n_over = *P.ov_count;  // a fixed small grid walks the (usually empty) list: no per-row block launch
  for (int oi = blockIdx.x; oi < n_over; oi += gridDim.x) {
  const int row = P.ov_list[oi];
  __syncthreads();
  const int k = P.k, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float* ls = dyn + (size_t)warp * k * 2;
  int32_t* li = reinterpret_cast<int32_t*>(ls + k);
  for (int x = lane; x < k; x += 32) { ls[x] = -CUDART_INF_F; li[x] = -1; }
  __syncwarp();
  const float* qr = P.q + (size_t)row * P.E;
  for (int64_t n = warp; n < P.N; n += nw) {      // ascending n per warp keeps ties at the lower index
    const float* tr = P.index + (size_t)n * P.E;
    float acc = 0.f;
    if ((P.E & 3) == 0) {
      for (int j = lane * 4; j < P.E; j += 128) {
        float4 a = *reinterpret_cast<const float4*>(qr + j);
        float4 b = *reinterpret_cast<const float4*>(tr + j);
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
    } else {
      for (int j = lane; j < P.E; j += 32) acc = fmaf(qr[j], tr[j], acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && acc > ls[k - 1]) {
      int p = k - 1;
      while (p > 0 && ls[p - 1] < acc) { ls[p] = ls[p - 1]; li[p] = li[p - 1]; --p; }
      ls[p] = acc;
      li[p] = (int32_t)(P.global_offset + n);
    }
    __syncwarp();
  }
  __syncthreads();
  // merge the nw lists: thread 0, k rounds (tiny)
  if (threadIdx.x == 0) {
    int head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int x = 0; x < k; ++x) {
      int bw = -1;
      for (int w = 0; w < nw; ++w) {
        if (head[w] >= k) continue;
        float s = dyn[(size_t)w * k * 2 + head[w]];
        int32_t i = reinterpret_cast<int32_t*>(dyn + (size_t)w * k * 2 + k)[head[w]];
        if (i < 0) continue;
        if (bw < 0) { bw = w; continue; }
        float bs = dyn[(size_t)bw * k * 2 + head[bw]];
        int32_t bi = reinterpret_cast<int32_t*>(dyn + (size_t)bw * k * 2 + k)[head[bw]];
        if (pair_before(s, i, bs, bi)) bw = w;
      }
      if (bw < 0) { P.out_s[(size_t)row * P.out_stride + x] = -CUDART_INF_F; P.out_i[(size_t)row * P.out_stride + x] = -1; continue; }
      P.out_s[(size_t)row * P.out_stride + x] = dyn[(size_t)bw * k * 2 + head[bw]];
      P.out_i[(size_t)row * P.out_stride + x] = reinterpret_cast<int32_t*>(dyn + (size_t)bw * k * 2 + k)[head[bw]];
      ++head[bw];
    }
  }
  }
}

// ------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// fp16 [rows, E] row-major, box = 64 cols x box_rows rows, 128-byte swizzle
int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int E, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return SSE_ECUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)E, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)E * 2};
  cuuint32_t box[2] = {KBLK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld E=%d", (int)r, (long long)rows, E); return SSE_ECUDA; }
  return SSE_OK;
}

// 3-D view (k within a 64-wide block, index row, k-block), box = one whole [box_rows x E] tile
int make_tmap3d(CUtensorMap* tm, const void* base, int64_t rows, int E, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return SSE_ECUDA;
  cuuint64_t gdim[3] = {(cuuint64_t)KBLK, (cuuint64_t)rows, (cuuint64_t)(E / KBLK)};
  cuuint64_t gstr[2] = {(cuuint64_t)E * 2, (cuuint64_t)KBLK * 2};
  cuuint32_t box[3] = {KBLK, (cuuint32_t)box_rows, (cuuint32_t)(E / KBLK)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SSE_OK : SSE_ECUDA;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

// Any encoding size up to 512: the fp16 scan copy of the index (and of the queries) is zero-padded to the next multiple
// of 64 -- exact, the padding contributes 0 to every dot product; the fp32 re-score reads the unpadded rows.
// k up to SSE_MAX_TOPK: for k > 32 the in-scan threshold tightening is off (it needs 4k list slots), rows whose
// candidate lists overflow fall back to the exact brute-force kernel.
bool search_tc_supported(int E, int64_t N, int k) {
  return E >= 1 && E <= 512 && k >= 1 && k <= SSE_MAX_TOPK && N >= 8192 && N >= 4 * (int64_t)k && N < ((int64_t)1 << 31) - 256;
}
static inline int pad64(int e) { return (e + 63) / 64 * 64; }

int search_tc_prepare(TcIndex& ti, const float* index_f32, int64_t N, int E, cudaStream_t st, int64_t* launches) {
  search_tc_release(ti);
  ti.N = N; ti.E = E;
  const int Ep = pad64(E);
  size_t bytes = (size_t)N * Ep * 2 + 64;
  cudaError_t e = cudaMalloc(&ti.h16, bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(fp16 index, %zu) failed: %s", bytes, cudaGetErrorString(e)); return SSE_ENOMEM; }
  if (Ep == E) SSE_TRY(f32_to_f16(index_f32, ti.h16, N * E, st, launches));
  else SSE_TRY(convert_to_16(index_f32, N, E, E, reinterpret_cast<uint16_t*>(ti.h16), Ep, 0, st, launches));
  // slot for max |t| lives behind the matrix
  float* tnorm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ti.h16) + align_up((size_t)N * Ep * 2, 16));
  SSE_CUDA_OK(cudaMemsetAsync(tnorm, 0, 4, st));
  max_row_norm_kernel<<<148 * 4, 256, 0, st>>>(index_f32, N, E, tnorm);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  SSE_TRY(make_tmap(reinterpret_cast<CUtensorMap*>(ti.tmap), ti.h16, N, Ep, 128));
  SSE_TRY(make_tmap(reinterpret_cast<CUtensorMap*>(ti.tmap64), ti.h16, N, Ep, 64));
  ti.use3d = getenv("SSE_SCAN_NO3D") == nullptr &&
             make_tmap3d(reinterpret_cast<CUtensorMap*>(ti.tmap3d), ti.h16, N, Ep, 128) == SSE_OK &&
             make_tmap3d(reinterpret_cast<CUtensorMap*>(ti.tmap3d64), ti.h16, N, Ep, 64) == SSE_OK;
  if (!ti.group_ctr) {
    SSE_CUDA_OK(cudaMalloc(&ti.group_ctr, MAX_GROUPS * sizeof(unsigned)));
    SSE_CUDA_OK(cudaMemsetAsync(ti.group_ctr, 0, MAX_GROUPS * sizeof(unsigned), st));
  }
  ti.tmap_ok = true;
  return SSE_OK;
}

void search_tc_release(TcIndex& ti) {
  if (ti.h16) cudaFree(ti.h16);
  if (ti.group_ctr) cudaFree(ti.group_ctr);
  ti.group_ctr = nullptr;
  ti.h16 = nullptr; ti.N = 0; ti.E = 0; ti.tmap_ok = false;
}

int search_tc_max_rows(int E) { (void)E; return TILE_M * MAX_GROUPS; }   // one m-tile per group in the worst case

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

int search_tc(const float* q, int Q, int E, const float* index_f32, TcIndex& ti, int64_t global_offset, int k,
              float* out_scores, int32_t* out_idx, Scratch& ws, int num_sms, cudaStream_t st, int64_t* launches, int out_stride,
              int late_ctas, int late_share) {
  if (Q <= 0) return SSE_OK;
  if (out_stride <= 0) out_stride = k;
  if (!ti.tmap_ok || ti.E != E) { set_error("search_tc: index not prepared"); return SSE_ESTATE; }
  const int64_t N = ti.N;
  const int E_true = E;
  E = pad64(E_true);                               // the scan works on the zero-padded fp16 copies
  const int KB = E / KBLK;
  const int m_tiles = cdiv(Q, TILE_M);
  const size_t ring_cap = 232448 - 1024 - 512;
  // ---- work decomposition (knobs for experiments: SSE_SCAN_MTG / _ACC1 / _CLUSTER / _CS / _PACK)
  // m-tiles per CTA.  TMEM budget (512 columns): queries mtg*E/2 + accumulators.  Two m-tiles per CTA make one index
  // tile feed 256 query rows (half the shared-memory fill per flop); their accumulators are either 2 x 2 x 64 columns
  // (64-column tiles, double-buffered over tiles) or -- ACC1, the default -- ONE 128-column accumulator per m-tile,
  // the two m-tiles double-buffering each other.  Measured at 600 x 1M (whole search, round 2): 64-column tiles 0.379 ms,
  // ACC1 0.327 ms; one m-tile per CTA with double-buffered 128-column tiles 0.334 ms (0.453 vs 0.342 ms at 2400 x 250k).
  static const int env_mtg = env_int("SSE_SCAN_MTG", 0), env_cs = env_int("SSE_SCAN_CS", 0), env_acc1 = env_int("SSE_SCAN_ACC1", -1);
  // Thread-block clusters (SSE_SCAN_CLUSTER=1): cs CTAs holding consecutive m-groups sweep the same tile range and share
  // ONE multicast TMA load per index tile.  Correct and tested, but measured SLOWER than unicast loads with L2 sharing
  // (600 x 1M: 0.338 vs 0.327 ms with ACC1, 0.421 vs 0.334 ms with one m-tile per CTA and cs = 5): the per-tile handshake
  // (all CTAs release the slot -> leader issues -> data lands) adds latency a 3-slot ring cannot hide, the L2->SM fabric
  // was not the limiter (7.7 TB/s measured in the unicast one-m-tile configuration), and GPC granularity leaves SMs idle.
  static const bool env_cluster = env_int("SSE_SCAN_CLUSTER", 0) != 0;
  int mtg = (E > 256 || m_tiles == 1) ? 1 : (env_mtg ? std::min(env_mtg, 2) : ((m_tiles == 2 || m_tiles >= 4) ? 2 : 1));
  bool acc1 = false;
  int tn;
  if (mtg == 2) {
    acc1 = env_acc1 != 0 && E + 2 * 128 <= 512 && (size_t)3 * 128 * E * 2 <= ring_cap;
    tn = acc1 ? 128 : 64;
  } else {
    tn = (E / 2 + 2 * 128 <= 512 && (size_t)3 * 128 * E * 2 <= ring_cap) ? 128 : 64;
  }
  const int a_cols = mtg * E / 2;
  const int n_groups0 = cdiv(m_tiles, mtg);
  int cs = 1, nsg = n_groups0;
  if (env_cluster && n_groups0 > 1) {
    if (n_groups0 <= 8) { cs = n_groups0; nsg = 1; }
    else {
      int best_pad = 1 << 30;
      for (int c = 8; c >= 4; --c) {
        int pad = cdiv(n_groups0, c) * c - n_groups0;
        if (pad < best_pad) { best_pad = pad; cs = c; }
      }
      nsg = cdiv(n_groups0, cs);
    }
    if (env_cs > 0 && env_cs <= 8) { cs = env_cs; nsg = cdiv(n_groups0, cs); }
  }
  const int n_groups = nsg * cs;
  if (n_groups > MAX_GROUPS) { set_error("search_tc: Q=%d too large for one call (max %d rows)", Q, search_tc_max_rows(E)); return SSE_EINVAL; }
  const int Qp = n_groups * mtg * TILE_M;
  const int gstride = mtg * TILE_M;                 // padded rows per group
  // Row packing.  "full": groups of gstride rows, the last one holds the remainder and may need fewer m-tiles (600
  // rows -> 256 + 256 + 88: five m-tiles of MMA work instead of six); the groups then get work items in proportion to
  // their per-tile cost.  "equal": every group holds cdiv(Q, n_groups) rows, all groups cost the same and sweep
  // identical tile ranges in lock-step (their re-reads of a tile hit L2).  Full packing pays when it removes >= 10 % of
  // the m-tiles (round 1: Q=600 0.383 -> 0.359 ms; but Q=2400, where it saves 1 m-tile in 20 and the odd group breaks
  // the lock-step L2 sharing, 0.367 -> 0.426).  SSE_SCAN_PACK=0/1 forces.  Clusters always pack equally.
  static const int pack_env = env_int("SSE_SCAN_PACK", -1);
  const int mt_last_full = cdiv(Q - (n_groups - 1) * gstride, TILE_M);
  const bool pack_full = cs == 1 && (pack_env >= 0 ? pack_env != 0 : (mtg - mt_last_full) * 10 >= n_groups * mtg);
  const int rpg = pack_full ? gstride : cdiv(Q, n_groups);   // query rows per group (the last group may hold fewer)
  const int n_tiles = (int)cdiv64(N, tn);
  // sample = every 16th tile: tau = k-th largest sampled tile maximum - 2 eps keeps ~16 k candidates per query
  // (measured: 1/32 and 1/64 samples cost more in candidate handling than they save in the sample pass); rows whose
  // sampled threshold is loose tighten it inside the scan (compact_candidates)
  // The sample pass costs ~ Q N / div, the candidates it leaves ~ Q k div: the best divisor grows like sqrt(N / k).  Measured
  // (whole search, k = 10): N = 1M: div 16 best (0.314 ms vs 0.324 at 8); N = 250k: 8 (0.329 vs 0.341 at 16); N = 125k: 8 or lower
  // (0.373 vs 0.407 at 16).  div = 0.0506 sqrt(N / k) reproduces 16 at 1M / k = 10.
  static const int env_sample_div = getenv("SSE_SCAN_SAMPLE_DIV") ? std::max(1, atoi(getenv("SSE_SCAN_SAMPLE_DIV"))) : 0;
  const int sample_div = env_sample_div ? env_sample_div : std::max(4, std::min(32, (int)lrint(0.0506 * sqrt((double)N / std::max(k, 1)))));
  int n_s = (int)(N / sample_div / tn);
  if (n_s < 64) n_s = 64;
  if (n_s < 2 * k) n_s = 2 * k;
  if (n_s > 1024) n_s = 1024;
  if (n_s > n_tiles) n_s = n_tiles;
  const int s_step = n_tiles / n_s;

  // shared memory = the TMA ring only: NS slots of one whole [tn x E] fp16 index tile each
  const size_t stage_bytes = (size_t)tn * E * 2;
  int NS = (int)(ring_cap / stage_bytes);
  if (NS > 16) NS = 16;
  if (NS < 2) { set_error("search_tc: E=%d does not fit the shared-memory ring", E); return SSE_EINVAL; }
  const size_t smem = 1024 + (size_t)NS * stage_bytes + 512;

  // specialised instances for the common E = 256 (fully unrolled MMA issue), generic otherwise
  typedef void (*scan_fn)(const CUtensorMap, const ScanParams);
  scan_fn fn_tilemax, fn_filter, fn_fused;
  if (KB == 4 && tn == 64) { fn_tilemax = scan_kernel<MODE_TILEMAX, 4, 64, false>; fn_filter = scan_kernel<MODE_FILTER, 4, 64, false>; fn_fused = scan_kernel<MODE_FUSED, 4, 64, false>; }
  else if (KB == 4 && tn == 128 && !acc1) { fn_tilemax = scan_kernel<MODE_TILEMAX, 4, 128, false>; fn_filter = scan_kernel<MODE_FILTER, 4, 128, false>; fn_fused = scan_kernel<MODE_FUSED, 4, 128, false>; }
  else if (KB == 4 && tn == 128 && acc1) { fn_tilemax = scan_kernel<MODE_TILEMAX, 4, 128, true>; fn_filter = scan_kernel<MODE_FILTER, 4, 128, true>; fn_fused = scan_kernel<MODE_FUSED, 4, 128, true>; }
  else if (acc1) { fn_tilemax = scan_kernel<MODE_TILEMAX, 0, 0, true>; fn_filter = scan_kernel<MODE_FILTER, 0, 0, true>; fn_fused = scan_kernel<MODE_FUSED, 0, 0, true>; }
  else { fn_tilemax = scan_kernel<MODE_TILEMAX, 0, 0, false>; fn_filter = scan_kernel<MODE_FILTER, 0, 0, false>; fn_fused = scan_kernel<MODE_FUSED, 0, 0, false>; }
  // Fused scan (EXPERIMENT, SSE_SCAN_FUSED=1; k <= 16, no clusters): sample pass, threshold selection and filter pass in ONE launch --
  // the items of an m-group meet at a per-group barrier after their sample tiles (all items are co-resident: the grid
  // never exceeds the SM count and each CTA takes a whole SM), every epilogue thread then selects its own row's threshold
  // from the group's tile maxima.  Saves two launches, the second TMEM allocation / query staging and the fp16 query
  // pre-pass (the queries are scaled and converted while they are staged).  Measured SLOWER than the 3-kernel form in its
  // first version (600 x 1M: +0.07 ms): every epilogue thread walks its row's 488 sampled maxima with L2-latency-bound
  // loads, 59 items per group repeat the same selection, and the other roles idle meanwhile.  Off by default.
  static const bool env_fused = env_int("SSE_SCAN_FUSED", 0) != 0;
  // Programmatic dependent launch between the kernels of one search (SSE_SCAN_PDL=1): the filter scan's CTAs are
  // scheduled as soon as the sample pass's CTAs leave their SMs and overlap their whole set-up with select_tau.
  static const bool env_pdl = env_int("SSE_SCAN_PDL", 0) != 0;
  // spanning items (see ScanParams): whenever the groups are packed equally and the CTA budget is not a multiple of the group count
  static const int env_span = env_int("SSE_SCAN_SPAN", -1);
  // measured (whole search, auto vs SSE_SCAN_SPAN=0): 4800 x 125k (19 groups, 7 items each = 90 % of the SMs) 0.424 vs 0.439 ms, but
  // 2400 x 250k (10 groups, 95 %) 0.358 vs 0.349 and 1200 x 500k (5 groups, 98 %) 0.329 vs 0.305: the restaging and the second
  // candidate slot only pay when whole items would leave more than ~7 % of the SMs idle
  const int budget0 = std::max(num_sms, n_groups);
  const bool span_pays = (budget0 / n_groups) * n_groups * 100 < budget0 * 93;
  bool span = cs == 1 && !pack_full && n_groups > 1 && late_ctas == 0 && (env_span >= 0 ? env_span != 0 : span_pays);
  const bool fused = env_fused && k <= FUSED_MAX_K && cs == 1 && ti.group_ctr != nullptr && !span;
  SSE_CUDA_OK(cudaFuncSetAttribute(fn_tilemax, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  SSE_CUDA_OK(cudaFuncSetAttribute(fn_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  SSE_CUDA_OK(cudaFuncSetAttribute(fn_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));

  // clusters per super-group: fill the SM budget, but never more clusters than can be co-resident (one CTA per SM and
  // the CTAs of a cluster share a GPC, so fewer than num_sms / cs clusters may fit)
  cudaLaunchAttribute lattr[1];
  cudaLaunchConfig_t lcfg;
  memset(&lcfg, 0, sizeof(lcfg));
  lcfg.blockDim = dim3(SCAN_THREADS, 1, 1);
  lcfg.dynamicSmemBytes = smem;
  lcfg.stream = st;
  lcfg.numAttrs = 0;
  if (cs > 1) {
    lattr[0].id = cudaLaunchAttributeClusterDimension;
    lattr[0].val.clusterDim.x = cs; lattr[0].val.clusterDim.y = 1; lattr[0].val.clusterDim.z = 1;
    lcfg.attrs = lattr;
    lcfg.numAttrs = 1;
  }
  ScanParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.n_groups = n_groups; sp.mtg = mtg; sp.kb = KB; sp.n_stages = NS; sp.tn = tn; sp.a_cols = a_cols;
  sp.N = N; sp.Qp = Qp; sp.global_offset = global_offset;
  for (int g = 0; g < n_groups; ++g)
    sp.group_mt[g] = std::max(1, std::min(mtg, cdiv(std::max(0, std::min(rpg, Q - g * rpg)), TILE_M)));
  const int budget = std::max(num_sms, n_groups);
  int R = std::max(1, budget / n_groups);
  int items = 0;
  if (cs > 1) {
    static std::map<std::pair<const void*, int>, int> max_clusters_cache;
    auto key = std::make_pair((const void*)fn_filter, cs);
    auto it = max_clusters_cache.find(key);
    int max_cl = 0;
    if (it != max_clusters_cache.end()) max_cl = it->second;
    else {
      lcfg.gridDim = dim3(cs * 32, 1, 1);
      if (cudaOccupancyMaxActiveClusters(&max_cl, fn_filter, &lcfg) != cudaSuccess) { cudaGetLastError(); max_cl = 0; }
      max_clusters_cache[key] = max_cl;
    }
    if (max_cl > 0) R = std::max(1, std::min(R, max_cl / nsg));
    R = std::min(R, n_tiles);
    items = nsg * R * cs;
  } else {
    // groups with the same number of live m-tiles get the same number of work items with IDENTICAL tile ranges: the
    // CTAs of different groups that share a range run concurrently, so the index is fetched from HBM about once per
    // distinct item count and the other groups' reads of the same tiles hit L2.  A lighter last group (one live
    // m-tile) gets fewer items, in proportion to its per-tile cost (per m-tile MMA time + a small per-tile constant).
    // measured per-tile cost (in-kernel counters, 600 x 1M, ACC1): an item with two live m-tiles 2687 cycles, with one 1719
    // -- a single m-tile cannot amortise the 64 KB tile fill -- i.e. c_tile : c_mt = 0.78 : 1 (SSE_SCAN_COST="c_tile,c_mt")
    double c_tile = acc1 ? 780.0 : 400.0, c_mt = acc1 ? 1000.0 : 550.0;
    if (const char* ce = getenv("SSE_SCAN_COST")) { double a = 0, b = 0; if (sscanf(ce, "%lf,%lf", &a, &b) == 2 && a >= 0 && b > 0) { c_tile = a; c_mt = b; } }
    double cost_sum = 0;
    for (int g = 0; g < n_groups; ++g) cost_sum += c_tile + c_mt * sp.group_mt[g];
    int first = 0;
    for (int g = 0; g < n_groups; ++g) {
      int Rg = (int)((c_tile + c_mt * sp.group_mt[g]) / cost_sum * budget);
      Rg = std::max(1, std::min(Rg, n_tiles));
      sp.group_first_item[g] = first;
      sp.group_items[g] = Rg;
      first += Rg;
    }
    sp.n_early = first;
    // late items (see ScanParams): spread over the groups like the regular ones
    sp.late_share = std::max(0, std::min(256, late_share * 256 / 100));
    for (int g = 0; g < n_groups; ++g) {
      int Rl = late_ctas > 0 ? (int)((c_tile + c_mt * sp.group_mt[g]) / cost_sum * late_ctas + 0.5) : 0;
      sp.group_first_late[g] = first;
      sp.group_late[g] = Rl;
      first += Rl;
    }
    items = first;
    R = sp.group_items[0];
  }
  if (cs > 1) sp.n_early = items;
  int total_slots = items;
  int span_first[MAX_GROUPS], span_n[MAX_GROUPS];
  if (span) {
    // one flat list of n_groups * n_tiles units cut into `items` equal pieces (same cut for the sample pass with n_s units per group)
    items = std::max(n_groups, std::min(budget, (int)std::min<int64_t>((int64_t)n_groups * n_s, 1 << 20)));
    sp.span = 1;
    sp.n_early = items;
    for (int g = 0; g < n_groups; ++g) sp.group_mt[g] = sp.group_mt[0];
    // candidate slots of the FILTER pass: one per (item, segment), numbered in item order
    const int64_t U = (int64_t)n_groups * n_tiles;
    int slot = 0;
    for (int g = 0; g < n_groups; ++g) { span_first[g] = -1; span_n[g] = 0; }
    for (int i = 0; i < items; ++i) {
      const int64_t u0 = U * i / items, u1 = U * (i + 1) / items;
      const int ga = (int)(u0 / n_tiles);
      if (span_first[ga] < 0) span_first[ga] = slot;
      ++span_n[ga]; ++slot;
      if (u1 > (int64_t)(ga + 1) * n_tiles) {
        const int gb = ga + 1;
        if (span_first[gb] < 0) span_first[gb] = slot;
        ++span_n[gb]; ++slot;
      }
    }
    total_slots = slot;
  }
  sp.cs = cs; sp.R = R;
  // SSE_SCAN_FOLLOW=1 (experiment, off): measured at 600 x 1M -- DRAM reads of the filter scan 0.95 -> 0.67 GB, L2 hit 42 -> 51 %, but
  // the whole search 0.325 -> 0.342 ms: the strided order costs the remainder group's roles two integer divisions per tile and its
  // items still run ~18 % behind the heavy groups; HBM (3.8 TB/s of 6.5) was not what the heavy groups wait for
  static const int env_follow = env_int("SSE_SCAN_FOLLOW", 0);
  sp.follow = (env_follow && pack_full && cs == 1 && !span && n_groups >= 2 && sp.group_mt[n_groups - 1] < mtg && sp.group_mt[0] == mtg &&
               (sp.group_late[0] > 0) == (sp.group_late[n_groups - 1] > 0) && (int64_t)n_tiles * 256 * (items + 2) < (int64_t)1 << 31) ? 1 : 0;
  lcfg.gridDim = dim3(items, 1, 1);
  const int list_items = span ? total_slots : items;      // candidate lists to allocate / walk

  // workspace carve
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  size_t o_qb = carve((size_t)Qp * E * 2);
  size_t o_qn = carve((size_t)Qp * 4);
  size_t o_tau = carve((size_t)Qp * 4);
  size_t o_mg = carve((size_t)Qp * 4);
  size_t o_tm = carve((size_t)n_s * Qp * 4);
  size_t o_cs = carve((size_t)list_items * mtg * TILE_M * CAND_CAP * 4);
  size_t o_ci = carve((size_t)list_items * mtg * TILE_M * CAND_CAP * 4);
  size_t o_cc = carve((size_t)list_items * mtg * TILE_M * 4);
  size_t o_ov = carve((size_t)Qp * 4);
  size_t o_ovl = carve((size_t)Qp * 4);
  size_t o_ovc = carve(16);
  const bool want_dbg = getenv("SSE_SCAN_DEBUG") != nullptr;
  size_t o_dbg = carve(want_dbg ? (size_t)items * DBG_N * 8 : 8);
  SSE_TRY(ws.ensure(off));
  uint8_t* w = ws.as<uint8_t>();
  __half* qb = reinterpret_cast<__half*>(w + o_qb);
  float* qn = reinterpret_cast<float*>(w + o_qn);
  float* tau = reinterpret_cast<float*>(w + o_tau);
  float* mg = reinterpret_cast<float*>(w + o_mg);
  float* tm = reinterpret_cast<float*>(w + o_tm);
  float* tnorm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ti.h16) + align_up((size_t)N * E * 2, 16));

  const CUtensorMap& tmi = ti.use3d ? *reinterpret_cast<const CUtensorMap*>(tn == 128 ? ti.tmap3d : ti.tmap3d64)
                                    : *reinterpret_cast<const CUtensorMap*>(tn == 128 ? ti.tmap : ti.tmap64);
  sp.qb = qb;
  sp.use3d = ti.use3d ? 1 : 0;
  sp.k = k;
  sp.dbg = want_dbg ? reinterpret_cast<long long*>(w + o_dbg) : nullptr;
  sp.dbg_flags = (want_dbg && getenv("SSE_SCAN_FLAGS")) ? atoi(getenv("SSE_SCAN_FLAGS")) : 0;
  sp.cand_s = reinterpret_cast<float*>(w + o_cs);
  sp.cand_i = reinterpret_cast<int32_t*>(w + o_ci);
  sp.cand_cnt = reinterpret_cast<int32_t*>(w + o_cc);
  if (fused) {
    sp.q32 = q; sp.Q = Q; sp.E_true = E_true; sp.gstride = gstride; sp.rpg = rpg; sp.tnorm_max = tnorm;
    sp.margin_out = mg; sp.tau_out = tau; sp.group_ctr = ti.group_ctr; sp.n_s = n_s; sp.s_step = s_step;
    sp.n_j = n_tiles; sp.tile_step = 1; sp.tilemax = tm; sp.tau = nullptr; sp.margin = nullptr;
    SSE_CUDA_OK(cudaLaunchKernelEx(&lcfg, fn_fused, tmi, sp));
    if (launches) ++*launches;
  } else {
    prep_queries_kernel<<<cdiv(Qp, 8), 256, 0, st>>>(q, Q, Qp, E_true, E, gstride, rpg, qb, qn, reinterpret_cast<int32_t*>(w + o_ovc));
    if (launches) ++*launches;
    cudaLaunchAttribute pattr[2];
    int n_scan_attrs = lcfg.numAttrs;
    if (env_pdl) {
      for (int a = 0; a < n_scan_attrs; ++a) pattr[a] = lcfg.attrs[a];
      pattr[n_scan_attrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      pattr[n_scan_attrs].val.programmaticStreamSerializationAllowed = 1;
      lcfg.attrs = pattr;
      lcfg.numAttrs = n_scan_attrs + 1;
    }
    // pass A: tile maxima over the strided sample
    sp.n_j = n_s; sp.tile_step = s_step; sp.tilemax = tm; sp.tau = nullptr;
    {
      ScanParams spa = sp;
      spa.dbg = nullptr;
      SSE_CUDA_OK(cudaLaunchKernelEx(&lcfg, fn_tilemax, tmi, spa));
    }
    if (launches) ++*launches;
    {
      cudaLaunchConfig_t scfg;
      memset(&scfg, 0, sizeof(scfg));
      scfg.gridDim = dim3(cdiv(Qp, 8), 1, 1);
      scfg.blockDim = dim3(256, 1, 1);
      scfg.stream = st;
      cudaLaunchAttribute sattr[1];
      sattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      sattr[0].val.programmaticStreamSerializationAllowed = 1;
      scfg.attrs = sattr;
      scfg.numAttrs = env_pdl ? 1 : 0;
      SSE_CUDA_OK(cudaLaunchKernelEx(&scfg, select_tau_kernel, (const float*)tm, n_s, Qp, Q, k, gstride, rpg, (const float*)qn, (const float*)tnorm, tau, mg));
    }
    if (launches) ++*launches;
    // pass B: filter over all tiles
    sp.n_j = n_tiles; sp.tile_step = 1; sp.tilemax = nullptr; sp.tau = tau; sp.margin = mg;
    SSE_CUDA_OK(cudaLaunchKernelEx(&lcfg, fn_filter, tmi, sp));
    if (launches) ++*launches;
  }

  FinParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.n_groups = n_groups; fp.mtg = mtg; fp.rpg = rpg; fp.cs = cs; fp.R = R;
  for (int g = 0; g < n_groups; ++g) {
    fp.group_first_item[g] = sp.group_first_item[g]; fp.group_items[g] = sp.group_items[g];
    fp.group_first_late[g] = sp.group_first_late[g]; fp.group_late[g] = sp.group_late[g];
    if (span) { fp.span_first[g] = span_first[g]; fp.span_n[g] = span_n[g]; }
  }
  fp.span = span ? 1 : 0;
  fp.cand_s = sp.cand_s; fp.cand_i = sp.cand_i; fp.cand_cnt = sp.cand_cnt; fp.margin = mg;
  fp.q = q; fp.index = index_f32; fp.global_offset = global_offset; fp.N = N; fp.Q = Q; fp.E = E_true; fp.k = k;
  fp.out_s = out_scores; fp.out_i = out_idx; fp.out_stride = out_stride; fp.overflow = reinterpret_cast<int32_t*>(w + o_ov); fp.ov_list = reinterpret_cast<int32_t*>(w + o_ovl); fp.ov_count = reinterpret_cast<int32_t*>(w + o_ovc);
  fp.group_ctr = fused ? ti.group_ctr : nullptr;
  ti.last_cnt = sp.cand_cnt; ti.last_cnt_n = (int64_t)list_items * mtg * TILE_M; ti.last_overflow = fp.overflow; ti.last_Q = Q; ti.last_items = items;
  if (want_dbg) {
    std::vector<long long> hd((size_t)items * DBG_N);
    cudaStreamSynchronize(st);
    cudaMemcpy(hd.data(), w + o_dbg, hd.size() * 8, cudaMemcpyDeviceToHost);
    const char* nm[DBG_N] = {"prod_wait_empty", "prod_total", "mma_wait_full", "mma_wait_acce", "mma_total", "mma_wait_a", "epi_wait_accf", "epi_total",
                             "epi_tmem_ld", "epi_compare", "epi_cand_lane0", "lead_wait_peers"};
    for (int c = 0; c < DBG_N; ++c) {
      long long mn = 1LL << 62, mx = 0, sm = 0;
      for (int i = 0; i < items; ++i) { long long v = hd[(size_t)i * DBG_N + c]; mn = std::min(mn, v); mx = std::max(mx, v); sm += v; }
      fprintf(stderr, "[scan dbg] %-16s min %10lld avg %10lld max %10lld  (items %d = %d sg x %d clusters x cs %d, mtg %d acc1 %d, tiles/cluster ~%d, tn %d, NS %d)\n",
              nm[c], mn, sm / items, mx, items, nsg, R, cs, mtg, (int)acc1, n_tiles / R, tn, NS);
    }
    if (cs == 1 && !span) {      // per m-group view: which group's items set the kernel's duration
      for (int g = 0; g < n_groups; ++g) {
        const int i0 = sp.group_first_item[g], n = sp.group_items[g];
        long long tot = 0, tmx = 0, wf = 0, wa = 0, ec = 0;
        for (int i = i0; i < i0 + n; ++i) {
          const long long* d = &hd[(size_t)i * DBG_N];
          tot += d[4]; tmx = std::max(tmx, d[4]); wf += d[2]; wa += d[3]; ec += d[9];
        }
        fprintf(stderr, "[scan dbg] group %2d: %3d items x ~%d tiles, m-tiles %d | mma_total avg %lld max %lld | wait_full avg %lld | wait_acce avg %lld | epi_compare avg %lld\n",
                g, n, n_tiles / std::max(n, 1), sp.group_mt[g], tot / n, tmx, wf / n, wa / n, ec / n);
      }
    }
  }
  fp.maxc = k <= 32 ? FIN_MAXC_SMALL : FIN_MAXC_LARGE;
  {
    // a row's candidates = one list of <= CAND_CAP entries per candidate slot of its m-group: when few slots feed a row (many
    // m-groups share the SMs: thousands of queries against a small shard) the sort buffers shrink with them and twice as
    // many rows are in flight per SM (4800 rows: 4 waves of 8 blocks per SM -> 2 waves of 15)
    int slots_max = 1;
    for (int g = 0; g < n_groups; ++g)
      slots_max = std::max(slots_max, span ? span_n[g] : (cs > 1 ? R : sp.group_items[g] + sp.group_late[g]));
    const int need = (slots_max * CAND_CAP + 127) / 128 * 128;
    if (need < fp.maxc) fp.maxc = std::max(need, 256);
  }
  {
    static bool fin_attr = false;
    if (!fin_attr) { SSE_CUDA_OK(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FIN_MAXC_LARGE * 12 + FIN_MAXR * 4)); fin_attr = true; }
  }
  finalize_kernel<<<Q, 128, (size_t)fp.maxc * 12 + FIN_MAXR * 4, st>>>(fp);
  if (launches) ++*launches;
  if (fused) SSE_CUDA_OK(cudaMemsetAsync(fp.ov_count, 0, 4, st));      // (the 3-kernel form zeroes it in prep_queries)
  fallback_kernel<<<std::min(Q, 128), 256, (size_t)8 * k * 8, st>>>(fp);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

int search_tc_stats(const TcIndex& ti, SearchStats* out) {
  *out = SearchStats();
  if (!ti.last_cnt || ti.last_Q <= 0) return SSE_OK;
  std::vector<int32_t> cnt((size_t)ti.last_cnt_n), ov((size_t)ti.last_Q);
  SSE_CUDA_OK(cudaDeviceSynchronize());
  SSE_CUDA_OK(cudaMemcpy(cnt.data(), ti.last_cnt, cnt.size() * 4, cudaMemcpyDeviceToHost));
  SSE_CUDA_OK(cudaMemcpy(ov.data(), ti.last_overflow, ov.size() * 4, cudaMemcpyDeviceToHost));
  for (int32_t c : cnt) out->candidates += std::min<int32_t>(std::max<int32_t>(c, 0), CAND_CAP);
  for (int32_t o : ov) out->fallback_rows += o != 0;
  out->rows = ti.last_Q; out->items = ti.last_items;
  return SSE_OK;
}

}  // namespace sse
