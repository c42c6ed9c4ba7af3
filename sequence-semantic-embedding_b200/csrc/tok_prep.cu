// Token pre-pass of the LSTM encoders (device side, one or three tiny launches per batch):
//   1. range check -- ids outside [0, V) would index past the embedding / input-projection tables (and, in training,
//      scatter gradients out of bounds).  TensorFlow's gather raises InvalidArgument there (reference
//      sse_model.py:163-164); here such ids are replaced by PAD_ID and COUNTED in a sticky device counter that the
//      host entry points turn into SSE_EINVAL (device entry points: sse_token_errors()).
//   2. pad-prefix bucketing -- rows are left-padded with PAD_ID=0 (reference data_utils.py:149-155,
//      sse_index.py:79-85) and every row starts from the zero state, so the state after p leading PADs is the same
//      for every row.  Rows are counting-sorted by their number of leading PADs (descending), so that
//      the rows of one kernel tile share (almost) the same prefix length; a tile then starts at
//      t0 = lead_sorted[last row of the tile] from the tabulated state S[t0] instead of running the PAD steps.
// Outputs: stok [B,T] sanitised tokens in SORTED order, perm[pos] = original row, lead_sorted[pos] = leading PADs.
#include "sse_common.cuh"

namespace sse {

namespace {

constexpr int TP_THREADS = 1024;

// bin[key] += (number of active lanes holding `key`); returns this lane's rank among them + the old bin value: ONE
// shared-memory atomic per distinct key and warp instead of one per lane (same-address atomics serialise)
__device__ __forceinline__ int warp_agg_add(int* bins, int key) {
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, key);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&bins[key], __popc(peers));
  base = __shfl_sync(peers, base, leader);
  return base + __popc(peers & ((1u << lane) - 1u));
}
constexpr int TP_MAX_T = 2048;        // histogram bins held in shared memory

// warp per row: number of leading PADs (clamped to T-1) and number of out-of-range ids
__device__ __forceinline__ void scan_row(const int32_t* __restrict__ row, int T, int V, int lane, int& lead, int& bad) {
  lead = T;
  bad = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const int v = t < T ? __ldg(row + t) : 1;
    const bool oob = t < T && (v < 0 || v >= V);
    bad += __popc(__ballot_sync(0xffffffffu, oob));
    const unsigned nz = __ballot_sync(0xffffffffu, t < T && v != 0 && !oob);     // an out-of-range id becomes PAD
    if (lead == T && nz) lead = t0 + __ffs(nz) - 1;
  }
  if (lead > T - 1) lead = T - 1;
}

__device__ __forceinline__ void copy_row(const int32_t* __restrict__ src, int32_t* __restrict__ dst, int T, int V, int lane) {
  for (int t = lane; t < T; t += 32) {
    const int v = __ldg(src + t);
    dst[t] = (v < 0 || v >= V) ? 0 : v;
  }
}

// thread per row (small batches): every load of the row is independent of the others, so they are all in flight together
// (a warp-per-row loop over the batch serialises ~20 rows x 2 global-load latencies per warp: measured 30 us at 600 rows)
__device__ __forceinline__ void scan_row_thread(const int32_t* __restrict__ row, int T, int V, int& lead, int& bad) {
  lead = T;
  bad = 0;
  if ((T & 1) == 0) {
    const int2* r2 = reinterpret_cast<const int2*>(row);
#pragma unroll 8
    for (int t = T / 2 - 1; t >= 0; --t) {              // backwards: the last assignment that sticks is the FIRST live position
      const int2 v = __ldg(r2 + t);
      const bool o0 = v.x < 0 || v.x >= V, o1 = v.y < 0 || v.y >= V;
      bad += (int)o0 + (int)o1;
      if (v.y != 0 && !o1) lead = 2 * t + 1;
      if (v.x != 0 && !o0) lead = 2 * t;
    }
  } else {
#pragma unroll 8
    for (int t = T - 1; t >= 0; --t) {
      const int v = __ldg(row + t);
      const bool o = v < 0 || v >= V;
      bad += (int)o;
      if (v != 0 && !o) lead = t;
    }
  }
  if (lead > T - 1) lead = T - 1;
}
__device__ __forceinline__ void copy_row_thread(const int32_t* __restrict__ src, int32_t* __restrict__ dst, int T, int V) {
  if ((T & 1) == 0) {
    const int2* s2 = reinterpret_cast<const int2*>(src);
    int2* d2 = reinterpret_cast<int2*>(dst);
#pragma unroll 8
    for (int t = 0; t < T / 2; ++t) {
      int2 v = __ldg(s2 + t);
      if (v.x < 0 || v.x >= V) v.x = 0;
      if (v.y < 0 || v.y >= V) v.y = 0;
      d2[t] = v;
    }
  } else {
#pragma unroll 8
    for (int t = 0; t < T; ++t) {
      const int v = __ldg(src + t);
      dst[t] = (v < 0 || v >= V) ? 0 : v;
    }
  }
}

// B <= one block's worth of work: everything in ONE launch.  REGS = true (T even, T <= 64): a thread reads its row ONCE into
// registers (all loads in flight together) and writes it from there; otherwise the row is read twice in unrolled batches.
template <bool REGS>
__global__ void __launch_bounds__(TP_THREADS) tok_prep_fused_kernel(const int32_t* __restrict__ tok, int B, int T, int V, int sort,
                                                                    int32_t* __restrict__ stok, int32_t* __restrict__ perm,
                                                                    int32_t* __restrict__ lead_sorted, int* __restrict__ bad_total) {
  __shared__ int hist[TP_MAX_T];
  __shared__ int s_bad;
  extern __shared__ int16_t lead_s[];      // [B]
  for (int i = threadIdx.x; i < T; i += TP_THREADS) hist[i] = 0;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if (REGS) {
    // B <= TP_THREADS here: one row per thread
    const int r = threadIdx.x;
    int2 v[32];
    int lead = T, bad = 0;
    if (r < B) {
      const int2* r2 = reinterpret_cast<const int2*>(tok + (size_t)r * T);
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] = t < T / 2 ? __ldg(r2 + t) : make_int2(0, 0);
#pragma unroll
      for (int t = 31; t >= 0; --t) {
        if (t < T / 2) {
          const bool o0 = v[t].x < 0 || v[t].x >= V, o1 = v[t].y < 0 || v[t].y >= V;
          bad += (int)o0 + (int)o1;
          if (o0) v[t].x = 0;
          if (o1) v[t].y = 0;
          if (v[t].y != 0) lead = 2 * t + 1;
          if (v[t].x != 0) lead = 2 * t;
        }
      }
      if (lead > T - 1) lead = T - 1;
      if (sort) warp_agg_add(hist, lead);
      if (bad) atomicAdd(&s_bad, bad);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (s_bad) atomicAdd(bad_total, s_bad);
      if (sort) {
        int run = 0;
        for (int p = T - 1; p >= 0; --p) { const int c = hist[p]; hist[p] = run; run += c; }
      }
    }
    __syncthreads();
    if (r < B) {
      const int pos = sort ? warp_agg_add(hist, lead) : r;
      int2* d2 = reinterpret_cast<int2*>(stok + (size_t)pos * T);
#pragma unroll
      for (int t = 0; t < 32; ++t)
        if (t < T / 2) d2[t] = v[t];
      perm[pos] = r;
      lead_sorted[pos] = lead;
    }
    return;
  }
  int bad_t = 0;
  for (int r = threadIdx.x; r < B; r += TP_THREADS) {
    int lead, bad;
    scan_row_thread(tok + (size_t)r * T, T, V, lead, bad);
    bad_t += bad;
    lead_s[r] = (int16_t)lead;
    if (sort) atomicAdd(&hist[lead], 1);
  }
  if (bad_t) atomicAdd(&s_bad, bad_t);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_bad) atomicAdd(bad_total, s_bad);
    if (sort) {           // descending prefix length: bin T-1 first
      int run = 0;
      for (int p = T - 1; p >= 0; --p) { const int c = hist[p]; hist[p] = run; run += c; }
    }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < B; r += TP_THREADS) {
    const int lead = lead_s[r];
    const int pos = sort ? atomicAdd(&hist[lead], 1) : r;
    copy_row_thread(tok + (size_t)r * T, stok + (size_t)pos * T, T, V);
    perm[pos] = r;
    lead_sorted[pos] = lead;
  }
}

// large batches: three launches with global histogram / cursors (hist: [T] ints, zeroed by the caller)
__global__ void tok_count_kernel(const int32_t* __restrict__ tok, int B, int T, int V, int32_t* __restrict__ lead_of, int* __restrict__ hist,
                                 int* __restrict__ bad_total) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  int bad_w = 0;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < B; r += gridDim.x * wpb) {
    int lead, bad;
    scan_row(tok + (size_t)r * T, T, V, lane, lead, bad);
    bad_w += bad;
    if (lane == 0) { lead_of[r] = lead; atomicAdd(&hist[lead], 1); }
  }
  if (lane == 0 && bad_w) atomicAdd(bad_total, bad_w);
}
__global__ void tok_scan_kernel(int* __restrict__ hist, int T) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int run = 0;
    for (int p = T - 1; p >= 0; --p) { const int c = hist[p]; hist[p] = run; run += c; }
  }
}
__global__ void tok_scatter_kernel(const int32_t* __restrict__ tok, int B, int T, int V, const int32_t* __restrict__ lead_of, int* __restrict__ cursor,
                                   int sort, int32_t* __restrict__ stok, int32_t* __restrict__ perm, int32_t* __restrict__ lead_sorted) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < B; r += gridDim.x * wpb) {
    const int lead = lead_of[r];
    int pos = r;
    if (sort) {
      if (lane == 0) pos = atomicAdd(&cursor[lead], 1);
      pos = __shfl_sync(0xffffffffu, pos, 0);
    }
    copy_row(tok + (size_t)r * T, stok + (size_t)pos * T, T, V, lane);
    if (lane == 0) { perm[pos] = r; lead_sorted[pos] = lead; }
  }
}

// y[perm[r]] = x[r] * rsqrt(max(sum x[r]^2, 1e-12)) (normalize) or x[r]; warp per row
__global__ void unpermute_rows_kernel(const float* __restrict__ x, float* __restrict__ y, const int32_t* __restrict__ perm, int rows, int cols,
                                      int normalize) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * cols;
  float inv = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int j = lane; j < cols; j += 32) { float v = xr[j]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    inv = rsqrtf(fmaxf(ss, 1e-12f));
    inv = inv * (1.5f - 0.5f * fmaxf(ss, 1e-12f) * inv * inv);      // same Newton step as l2norm_rows_kernel: bit-identical results
  }
  float* yr = y + (size_t)(perm ? perm[row] : row) * cols;
  for (int j = lane; j < cols; j += 32) yr[j] = normalize ? xr[j] * inv : xr[j];
}

// Query-batch tail of an LSTM tower in ONE launch: u = h M (reference sse_model.py:228,245,254: the [H,E] projection), the
// optional l2-normalisation (sse_model.py:282-283) and the un-permutation of pad-prefix-sorted rows.  Block = PR_ROWS rows
// x all E columns (thread == column e, 8 row accumulators); M streams from L2 once per block, coalesced.  Replaces
// sgemm + l2norm_rows / unpermute_rows (two launches, ~16 us with their gaps at 600 rows) for batches up to a few thousand rows.
constexpr int PR_ROWS = 8, PR_THREADS = 256, PR_MAXNE = 4;
__global__ void __launch_bounds__(PR_THREADS) project_rows_kernel(const float* __restrict__ hmat, int ldh, const float* __restrict__ M, int H, int E,
                                                                  int rows, const int32_t* __restrict__ perm, int normalize, float* __restrict__ out) {
  extern __shared__ float4 hs4[];          // [H][2] float4: the 8 rows' h values of hidden unit k, side by side
  __shared__ float red[PR_THREADS / 32][PR_ROWS];
  __shared__ float inv_s[PR_ROWS];
  float* hs = reinterpret_cast<float*>(hs4);
  const int r0 = blockIdx.x * PR_ROWS;
  for (int i = threadIdx.x; i < H * PR_ROWS; i += PR_THREADS) {
    const int r = i / H, k = i - r * H;                    // coalesced over k
    hs[k * PR_ROWS + r] = r0 + r < rows ? hmat[(size_t)(r0 + r) * ldh + k] : 0.f;
  }
  __syncthreads();
  float acc[PR_MAXNE][PR_ROWS];
#pragma unroll
  for (int n = 0; n < PR_MAXNE; ++n)
#pragma unroll
    for (int r = 0; r < PR_ROWS; ++r) acc[n][r] = 0.f;
#pragma unroll
  for (int n = 0; n < PR_MAXNE; ++n) {
    const int e = threadIdx.x + n * PR_THREADS;
    if (e < E) {
      const float* mp = M + e;
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float m = __ldg(mp + (size_t)k * E);
        const float4 a = hs4[k * 2], b = hs4[k * 2 + 1];
        acc[n][0] = fmaf(a.x, m, acc[n][0]); acc[n][1] = fmaf(a.y, m, acc[n][1]); acc[n][2] = fmaf(a.z, m, acc[n][2]); acc[n][3] = fmaf(a.w, m, acc[n][3]);
        acc[n][4] = fmaf(b.x, m, acc[n][4]); acc[n][5] = fmaf(b.y, m, acc[n][5]); acc[n][6] = fmaf(b.z, m, acc[n][6]); acc[n][7] = fmaf(b.w, m, acc[n][7]);
      }
    }
  }
  if (normalize) {
    float ss[PR_ROWS];
#pragma unroll
    for (int r = 0; r < PR_ROWS; ++r) {
      ss[r] = 0.f;
#pragma unroll
      for (int n = 0; n < PR_MAXNE; ++n) ss[r] = fmaf(acc[n][r], acc[n][r], ss[r]);
#pragma unroll
      for (int o = 16; o; o >>= 1) ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], o);
    }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
      for (int r = 0; r < PR_ROWS; ++r) red[threadIdx.x >> 5][r] = ss[r];
    __syncthreads();
    if (threadIdx.x < PR_ROWS) {
      float t = 0.f;
      for (int w = 0; w < PR_THREADS / 32; ++w) t += red[w][threadIdx.x];
      t = fmaxf(t, 1e-12f);
      float inv = rsqrtf(t);
      inv = inv * (1.5f - 0.5f * t * inv * inv);            // one Newton step: rsqrt(max(sum x^2, 1e-12)) to fp32 accuracy
      inv_s[threadIdx.x] = inv;
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < PR_ROWS; ++r) {
    if (r0 + r >= rows) break;
    const float inv = normalize ? inv_s[r] : 1.f;
    float* o = out + (size_t)(perm ? perm[r0 + r] : r0 + r) * E;
#pragma unroll
    for (int n = 0; n < PR_MAXNE; ++n) {
      const int e = threadIdx.x + n * PR_THREADS;
      if (e < E) o[e] = acc[n][r] * inv;
    }
  }
}

// no bucketing wanted (pad-prefix start off): a flat, coalesced range check + copy over all SMs (the one-block fused
// kernel above is latency-bound: 17 us at 600 x 50)
__global__ void sanitize_copy_kernel(const int32_t* __restrict__ tok, int64_t n, int V, int32_t* __restrict__ stok, int* __restrict__ bad_total) {
  int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = __ldg(tok + i);
    const bool oob = v < 0 || v >= V;
    stok[i] = oob ? 0 : v;
    bad += (int)oob;
  }
  if (__any_sync(0xffffffffu, bad != 0)) {
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(bad_total, bad);
  }
}

__global__ void sanitize_inplace_kernel(int32_t* __restrict__ tok, int64_t n, int V, int* __restrict__ bad_total) {
  int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = tok[i];
    if (v < 0 || v >= V) { tok[i] = 0; ++bad; }
  }
  if (bad) atomicAdd(bad_total, bad);
}

}  // namespace

// training path: the step works on its own copy of the batch, so ids are checked / replaced in place
int sanitize_tokens_inplace(int32_t* tokens, int64_t n, int V, int* bad_total, cudaStream_t st, int64_t* launches) {
  if (n <= 0) return SSE_OK;
  sanitize_inplace_kernel<<<(int)std::min<int64_t>(cdiv64(n, 256), 148 * 4), 256, 0, st>>>(tokens, n, V, bad_total);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

size_t tok_prep_ws_bytes(int B, int T) {
  return ((size_t)B * T * 4 + 255) / 256 * 256 + 3 * (((size_t)B * 4 + 255) / 256 * 256) + (((size_t)T * 4 + 255) / 256 * 256);
}

// ws layout: stok [B,T] | perm [B] | lead_sorted [B] | lead_of [B] | hist [T]
int tok_prep(const int32_t* tokens, int B, int T, int V, bool sort, void* ws, TokPrep* out, int* bad_total, cudaStream_t st,
             int64_t* launches) {
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  uint8_t* w = static_cast<uint8_t*>(ws);
  out->stok = reinterpret_cast<int32_t*>(w);
  out->perm = reinterpret_cast<int32_t*>(w + al((size_t)B * T * 4));
  out->lead_sorted = out->perm + al((size_t)B * 4) / 4;
  int32_t* lead_of = out->lead_sorted + al((size_t)B * 4) / 4;
  int* hist = reinterpret_cast<int*>(lead_of + al((size_t)B * 4) / 4);
  out->sorted = sort;
  if (B <= 0) return SSE_OK;
  if (!sort) {
    const int64_t n = (int64_t)B * T;
    sanitize_copy_kernel<<<(int)std::min<int64_t>(cdiv64(n, 256), 148 * 8), 256, 0, st>>>(tokens, n, V, out->stok, bad_total);
    if (launches) ++*launches;
    SSE_CUDA_OK(cudaGetLastError());
    return SSE_OK;
  }
  if (B <= 8192 && T <= TP_MAX_T) {
    if (B <= TP_THREADS && (T & 1) == 0 && T <= 64)
      tok_prep_fused_kernel<true><<<1, TP_THREADS, (size_t)B * 2, st>>>(tokens, B, T, V, sort ? 1 : 0, out->stok, out->perm, out->lead_sorted, bad_total);
    else
      tok_prep_fused_kernel<false><<<1, TP_THREADS, (size_t)B * 2, st>>>(tokens, B, T, V, sort ? 1 : 0, out->stok, out->perm, out->lead_sorted, bad_total);
    if (launches) ++*launches;
  } else {
    SSE_CUDA_OK(cudaMemsetAsync(hist, 0, (size_t)T * 4, st));
    const int blocks = std::min(cdiv(B, 8), 148 * 8);
    tok_count_kernel<<<blocks, 256, 0, st>>>(tokens, B, T, V, lead_of, hist, bad_total);
    tok_scan_kernel<<<1, 32, 0, st>>>(hist, T);
    tok_scatter_kernel<<<blocks, 256, 0, st>>>(tokens, B, T, V, lead_of, hist, sort ? 1 : 0, out->stok, out->perm, out->lead_sorted);
    if (launches) *launches += 3;
  }
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

bool project_rows_supported(int rows, int H, int E) { return rows <= 4096 && E <= PR_MAXNE * PR_THREADS && (size_t)H * PR_ROWS * 4 <= 48 * 1024; }
int project_rows(const float* hmat, int ldh, const float* M, int H, int E, int rows, const int32_t* perm, int normalize, float* out,
                 cudaStream_t st, int64_t* launches) {
  if (rows <= 0) return SSE_OK;
  project_rows_kernel<<<cdiv(rows, PR_ROWS), PR_THREADS, (size_t)H * PR_ROWS * 4, st>>>(hmat, ldh, M, H, E, rows, perm, normalize, out);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

int unpermute_rows(const float* x, float* y, const int32_t* perm, int rows, int cols, int normalize, cudaStream_t st, int64_t* launches) {
  if (rows <= 0) return SSE_OK;
  unpermute_rows_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, y, perm, rows, cols, normalize);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

}  // namespace sse

// ======================================================================================================================
// Device-side train-batch sampler (reference data.py:95-115).  The corpus lives in HBM as arrays -- source
// rows [P,T], the verified-target lists in CSR form over target ROW numbers, target rows [N,T] -- and one launch writes the
// step's 2B pair rows straight into the buffers sse_train_step consumes: rows alternate (positive pair, 1.0), (random
// NON-verified target, 0.0) for the window of positives [start, start + B).  Same rule and layout as the reference's
// python loop; the random stream is a counter-based hash of (seed, step, row, draw) instead of numpy's (the reference is
// unseeded), so the distribution is the same and a (seed, step) pair is reproducible.
namespace sse {
namespace {
__device__ __forceinline__ uint64_t mix64(uint64_t x) {      // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// warp per positive: lane 0 draws, all lanes copy the token rows
__global__ void sample_train_batch_kernel(const int32_t* __restrict__ src_rows, const int64_t* __restrict__ ver_off, const int32_t* __restrict__ ver_rows,
                                          const int32_t* __restrict__ tgt_rows, int64_t N, int T, int64_t start, int B, uint64_t seed, uint64_t step,
                                          int32_t* __restrict__ src_out, int32_t* __restrict__ tgt_out, float* __restrict__ lab_out) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= B) return;
  const int64_t i = start + w;
  const int64_t o0 = ver_off[i], cnt = ver_off[i + 1] - o0;
  int pos = 0, neg = 0;
  if (lane == 0) {
    const uint64_t base = mix64(seed ^ mix64(step * 0x100000001B3ull + (uint64_t)i));
    pos = ver_rows[o0 + (int64_t)(mix64(base) % (uint64_t)cnt)];
    for (uint64_t d = 1;; ++d) {                      // rejection: a target that is NOT verified for this source
      neg = (int)(mix64(base + d) % (uint64_t)N);
      bool clash = false;
      for (int64_t x = 0; x < cnt; ++x) clash |= ver_rows[o0 + x] == neg;
      if (!clash || d > 64 || cnt >= N) break;
    }
  }
  pos = __shfl_sync(0xffffffffu, pos, 0);
  neg = __shfl_sync(0xffffffffu, neg, 0);
  for (int t = lane; t < T; t += 32) {
    const int32_t sv = src_rows[(size_t)i * T + t];
    src_out[(size_t)(2 * w) * T + t] = sv;
    src_out[(size_t)(2 * w + 1) * T + t] = sv;
    tgt_out[(size_t)(2 * w) * T + t] = tgt_rows[(size_t)pos * T + t];
    tgt_out[(size_t)(2 * w + 1) * T + t] = tgt_rows[(size_t)neg * T + t];
  }
  if (lane == 0) { lab_out[2 * w] = 1.f; lab_out[2 * w + 1] = 0.f; }
}
}  // namespace

int sample_train_batch(const int32_t* src_rows, const int64_t* ver_off, const int32_t* ver_rows, const int32_t* tgt_rows, int64_t N, int T, int64_t start,
                       int B, uint64_t seed, uint64_t step, int32_t* src_out, int32_t* tgt_out, float* lab_out, cudaStream_t st, int64_t* launches) {
  if (B <= 0) return SSE_OK;
  sample_train_batch_kernel<<<cdiv(B, 8), 256, 0, st>>>(src_rows, ver_off, ver_rows, tgt_rows, N, T, start, B, seed, step, src_out, tgt_out, lab_out);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}
}  // namespace sse
