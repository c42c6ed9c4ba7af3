// fp32 SIMT cosine scoring + fused row top-k (exact mode; any E, any k <= 128).
//
// Replaces  distances = np.dot(sourceEncodings, targetEncodings.T);
//           getSortedResults(distances)[:, :k]
// (reference sse_evaluator.py:110-111, data_utils.py:263-267) and
// tf.nn.top_k(similarity, TOP_N) (sse_model.py:348).  The [Q,N] matrix is never
// materialised: each block owns a (query tile, column range), keeps a sorted
// k-list per query row in shared memory, and a second kernel merges the
// per-range lists.  Order = (score descending, index ascending) -- the TF
// top_k tie rule; numpy's argsort leaves ties unspecified.
#include "sse_common.cuh"
#include <math_constants.h>

namespace sse {

namespace {

constexpr int BQ = 64, BN = 64, BKS = 16;

// dynamic smem: S[BQ][BN+1] | list_s[BQ][k] | list_i[BQ][k]
__global__ void __launch_bounds__(256) search_simt_kernel(const float* __restrict__ q, int Q, int E,
                                                          const float* __restrict__ index, int64_t N,
                                                          int64_t cols_per_chunk, int n_chunks, int k,
                                                          int64_t global_offset, float* __restrict__ part_s,
                                                          int32_t* __restrict__ part_i) {
  __shared__ __align__(16) float As[BKS][BQ + 4];
  __shared__ __align__(16) float Bs[BKS][BN + 4];
  extern __shared__ float dyn[];
  float* S = dyn;                                  // [BQ][BN+1]
  float* list_s = dyn + BQ * (BN + 1);             // [BQ][k]
  int32_t* list_i = reinterpret_cast<int32_t*>(list_s + BQ * k);

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int q0 = blockIdx.y * BQ;
  const int chunk = blockIdx.x;
  const int64_t c_begin = (int64_t)chunk * cols_per_chunk;
  const int64_t c_end = min(N, c_begin + cols_per_chunk);

  for (int i = tid; i < BQ * k; i += 256) { list_s[i] = -CUDART_INF_F; list_i[i] = -1; }
  __syncthreads();

  for (int64_t n0 = c_begin; n0 < c_end; n0 += BN) {
    float acc[4][4] = {};
    for (int k0 = 0; k0 < E; k0 += BKS) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int idx = tid + e * 256;
        int r = idx >> 4, kk = idx & 15;
        int gk = k0 + kk;
        int gq = q0 + r;
        As[kk][r] = (gq < Q && gk < E) ? __ldg(q + (size_t)gq * E + gk) : 0.f;
        int64_t gn = n0 + r;
        Bs[kk][r] = (gn < c_end && gk < E) ? __ldg(index + (size_t)gn * E + gk) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BKS; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) S[(ty * 4 + i) * (BN + 1) + tx * 4 + j] = acc[i][j];
    __syncthreads();
    if (tid < BQ && q0 + tid < Q) {
      float* ls = list_s + tid * k;
      int32_t* li = list_i + tid * k;
      float thr = ls[k - 1];
      int ncols = (int)min((int64_t)BN, c_end - n0);
      for (int j = 0; j < ncols; ++j) {
        float s = S[tid * (BN + 1) + j];
        if (s > thr) {          // strict: on ties the earlier (lower) index stays
          int p = k - 1;
          while (p > 0 && ls[p - 1] < s) { ls[p] = ls[p - 1]; li[p] = li[p - 1]; --p; }
          ls[p] = s;
          li[p] = (int32_t)(global_offset + n0 + j);
          thr = ls[k - 1];
        }
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < BQ * k; i += 256) {
    int r = i / k, j = i % k;
    int gq = q0 + r;
    if (gq < Q) {
      size_t o = ((size_t)gq * n_chunks + chunk) * k + j;
      part_s[o] = list_s[i];
      part_i[o] = list_i[i];
    }
  }
}

// warp per query row: k rounds of arg-best over C = n_groups * k_in candidates, order (score desc, idx asc).
// Candidate x of group g of row r sits at cand[g * group_stride + r * row_stride + x] (scores and ids in separate or
// in one bit-cast array): [Q, C] lists (n_groups = 1) and the all-gathered per-shard packed [G, Q, 2k] blocks alike.
// Entries with idx < 0 are padding.  dynamic smem: per warp C floats + C ints
__global__ void merge_topk_kernel(const float* __restrict__ cand_s, const int32_t* __restrict__ cand_i, int Q, int n_groups,
                                  long long group_stride, int row_stride, int k_in, int k, float* __restrict__ out_s,
                                  int32_t* __restrict__ out_i, int out_stride) {
  extern __shared__ float dyn[];
  const int wpb = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * wpb + w;
  const int C = n_groups * k_in;
  float* s = dyn + (size_t)w * C * 2;
  int32_t* id = reinterpret_cast<int32_t*>(s + C);
  if (row >= Q) return;
  for (int j = lane; j < C; j += 32) {
    const int g = j / k_in, x = j - g * k_in;
    const size_t at = (size_t)g * group_stride + (size_t)row * row_stride + x;
    s[j] = cand_s[at];
    id[j] = cand_i[at];
  }
  __syncwarp();
  for (int r = 0; r < k; ++r) {
    float bs = -CUDART_INF_F;
    int32_t bi = 0x7fffffff;
    int bp = -1;
    for (int j = lane; j < C; j += 32) {
      int32_t ij = id[j];
      if (ij < 0) continue;
      float sj = s[j];
      if (sj > bs || (sj == bs && ij < bi)) { bs = sj; bi = ij; bp = j; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      float os = __shfl_xor_sync(0xffffffffu, bs, o);
      int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (op >= 0 && (bp < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; bp = op; }
    }
    if (lane == 0) {
      out_s[(size_t)row * out_stride + r] = bp >= 0 ? bs : -CUDART_INF_F;
      out_i[(size_t)row * out_stride + r] = bp >= 0 ? bi : -1;
    }
    if (bp >= 0 && lane == (bp & 31)) id[bp] = -1;   // consumed (owner lane wrote it originally)
    __syncwarp();
  }
}

}  // namespace

int merge_topk_strided(const float* cand_s, const int32_t* cand_i, int Q, int n_groups, int64_t group_stride, int row_stride,
                       int k_in, int k, float* out_s, int32_t* out_i, int out_stride, cudaStream_t st, int64_t* launches) {
  if (Q <= 0) return SSE_OK;
  const int C = n_groups * k_in;
  int wpb = 4;
  size_t smem = (size_t)wpb * C * 8;
  while (wpb > 1 && smem > 96 * 1024) { wpb >>= 1; smem = (size_t)wpb * C * 8; }
  if (smem > 200 * 1024) { set_error("merge_topk: C=%d too large", C); return SSE_EINVAL; }
  if (smem > 48 * 1024)
    SSE_CUDA_OK(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_topk_kernel<<<cdiv(Q, wpb), wpb * 32, smem, st>>>(cand_s, cand_i, Q, n_groups, (long long)group_stride, row_stride, k_in, k,
                                                          out_s, out_i, out_stride);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

int merge_topk(const float* cand_s, const int32_t* cand_i, int Q, int C, int k, float* out_s, int32_t* out_i,
               cudaStream_t st, int64_t* launches) {
  return merge_topk_strided(cand_s, cand_i, Q, 1, 0, C, C, k, out_s, out_i, k, st, launches);
}

int search_simt(const float* q, int Q, int E, const float* index, int64_t N, int64_t global_offset, int k,
                float* out_scores, int32_t* out_idx, Scratch& ws, int num_sms, cudaStream_t st,
                int64_t* launches, int out_stride) {
  if (out_stride <= 0) out_stride = k;
  if (Q <= 0) return SSE_OK;
  const int q_tiles = cdiv(Q, BQ);
  // enough column ranges to fill the machine ~2x, each a multiple of BN columns
  int want = max(1, (2 * num_sms) / q_tiles);
  int64_t n_tiles = cdiv64(max((int64_t)1, N), BN);
  int n_chunks = (int)std::min<int64_t>(want, n_tiles);
  n_chunks = std::max(1, std::min(n_chunks, (96 * 1024) / (8 * k)));       // the merge holds n_chunks * k candidates per row in shared memory
  int64_t tiles_per_chunk = cdiv64(n_tiles, n_chunks);
  n_chunks = (int)cdiv64(n_tiles, tiles_per_chunk);
  int64_t cols_per_chunk = tiles_per_chunk * BN;
  size_t part = (size_t)Q * n_chunks * k;
  SSE_TRY(ws.ensure(part * 8));
  float* part_s = ws.as<float>();
  int32_t* part_i = reinterpret_cast<int32_t*>(part_s + part);
  size_t smem = (size_t)BQ * (BN + 1) * 4 + (size_t)BQ * k * 8;
  if (smem > 48 * 1024)
    SSE_CUDA_OK(cudaFuncSetAttribute(search_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_chunks, q_tiles);
  search_simt_kernel<<<grid, 256, smem, st>>>(q, Q, E, index, N, cols_per_chunk, n_chunks, k, global_offset,
                                              part_s, part_i);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return merge_topk_strided(part_s, part_i, Q, 1, 0, n_chunks * k, n_chunks * k, k, out_scores, out_idx, out_stride, st, launches);
}

}  // namespace sse
