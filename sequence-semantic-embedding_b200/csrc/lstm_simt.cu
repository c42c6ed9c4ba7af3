// fp32 SIMT LSTM tower: exact-mode encoder and the training forward.
//
// Restates BasicLSTMCell(forget_bias=1) + static_rnn over all T positions
// (reference sse_model.py:222-224, 240-242, 248-250, 262-264, 273-274): per
// step  z = [x_t, h_{t-1}] K + b ; (i,j,f,o) = split(z) ;
//       c = c*sigmoid(f+1) + sigmoid(i)*tanh(j) ; h = tanh(c)*sigmoid(o).
// One launch per step fuses the embedding gather, the [B,We+H]x[We+H,4H]
// contraction and the gate math; every thread owns RPT batch rows x the four
// gates of one hidden unit, so the cell update never leaves registers.
// Works for any We/H/E/B (the real reference models use 30/40/50/96/64).
#include "sse_common.cuh"

namespace sse {

namespace {

constexpr int BK = 16;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// grid: (ceil(B/BM), ceil(H/16)); block 256 threads = 16 (row groups) x 16 (hidden units)
template <int RPT>
__global__ void __launch_bounds__(256) lstm_step_kernel(
    const int32_t* __restrict__ tokens, int T, int t,
    const float* __restrict__ emb, int We,
    const float* __restrict__ Kw, const float* __restrict__ bias,
    const float* __restrict__ h_prev,   // nullptr => zero state (h = c = 0)
    float* __restrict__ h_next, float* __restrict__ c,
    const float* __restrict__ init_h, const float* __restrict__ init_c,  // optional [H] broadcast state used instead of h_prev/c
    const int32_t* __restrict__ lead,                                    // optional [B]: per-row pad-prefix start (row r joins at step lead[r]
    const float* __restrict__ pad_h, const float* __restrict__ pad_c,    //   from the tabulated state pad_*[lead[r]-1]; earlier steps skip it)
    int B, int H,
    float* __restrict__ save_h, float* __restrict__ save_c, float* __restrict__ save_g) {
  constexpr int BM = 16 * RPT;
  constexpr int APITCH = BM + 4;
  __shared__ __align__(16) float As[2][BK][APITCH];
  __shared__ __align__(16) float Ws[2][BK][64];

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int row0 = blockIdx.x * BM;
  const int u0 = blockIdx.y * 16;
  const int H4 = 4 * H;
  const bool has_state = (h_prev != nullptr) || (init_h != nullptr);
  const int Ktot = has_state ? We + H : We;
  const int nk = (Ktot + BK - 1) / BK;

  // A-tile loader: element e -> (row = idx/16, kk = idx%16)
  float areg[RPT];
  float wreg[4];
  int arow[RPT];
  int atok[RPT];
  bool afirst[RPT];             // this is the row's first live step: its h_{t-1} comes from the pad-prefix table
#pragma unroll
  for (int e = 0; e < RPT; ++e) {
    int idx = tid + e * 256;
    int r = row0 + (idx >> 4);
    arow[e] = r < B ? r : B - 1;
    atok[e] = tokens[(size_t)arow[e] * T + t];
    afirst[e] = lead != nullptr && t > 0 && min(lead[arow[e]], T - 1) == t;
  }
  const int akk = tid & 15;
  const int wkk = tid >> 4;       // W-tile loader: row kk = tid/16, unit = tid%16, 4 gates
  const int wu = u0 + (tid & 15);

  auto load_tiles = [&](int kt) {
    int k = kt * BK + akk;
#pragma unroll
    for (int e = 0; e < RPT; ++e) {
      float v = 0.f;
      if (k < We) v = __ldg(emb + (size_t)atok[e] * We + k);
      else if (k < Ktot) v = afirst[e] ? __ldg(pad_h + (size_t)(t - 1) * H + (k - We))
                                       : (init_h ? __ldg(init_h + (k - We)) : h_prev[(size_t)arow[e] * H + (k - We)]);
      areg[e] = v;
    }
    int kw = kt * BK + wkk;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v = 0.f;
      if (kw < Ktot && wu < H) v = __ldg(Kw + (size_t)kw * H4 + g * H + wu);
      wreg[g] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int e = 0; e < RPT; ++e) {
      int idx = tid + e * 256;
      As[buf][akk][idx >> 4] = areg[e];
    }
    *reinterpret_cast<float4*>(&Ws[buf][wkk][(tid & 15) * 4]) = make_float4(wreg[0], wreg[1], wreg[2], wreg[3]);
  };

  float acc[RPT][4];
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[i][g] = 0.f;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[RPT];
#pragma unroll
      for (int i4 = 0; i4 < RPT / 4; ++i4) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][ty * RPT + i4 * 4]);
        a[i4 * 4 + 0] = v.x; a[i4 * 4 + 1] = v.y; a[i4 * 4 + 2] = v.z; a[i4 * 4 + 3] = v.w;
      }
      float4 w = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        acc[i][0] = fmaf(a[i], w.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], w.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], w.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], w.w, acc[i][3]);
      }
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  const int u = u0 + tx;
  if (u >= H) return;
  const float bi = __ldg(bias + u), bj = __ldg(bias + H + u), bf = __ldg(bias + 2 * H + u), bo = __ldg(bias + 3 * H + u);
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    int r = row0 + ty * RPT + i;
    if (r >= B) continue;
    size_t off = (size_t)r * H + u;
    float c_old = 0.f;
    const int lr = lead ? min(lead[r], T - 1) : 0;
    if (lead && t < lr) continue;                       // still inside this row's pad prefix: it joins at step lr
    if (lead && t == lr && t > 0) c_old = __ldg(pad_c + (size_t)(t - 1) * H + u);
    else if (init_c) c_old = __ldg(init_c + u);
    else if (h_prev) c_old = c[off];
    float si = sigmoidf_(acc[i][0] + bi);
    float tj = tanhf(acc[i][1] + bj);
    float sf = sigmoidf_(acc[i][2] + bf + 1.0f);   // forget_bias = 1.0 added at run time
    float so = sigmoidf_(acc[i][3] + bo);
    float cn = c_old * sf + si * tj;
    float tc = tanhf(cn);
    float hn = tc * so;
    c[off] = cn;
    h_next[off] = hn;
    if (save_h) {
      size_t so_ = ((size_t)t * B + r) * H + u;
      save_h[so_] = hn;
      save_c[so_] = cn;
      size_t sg = ((size_t)t * B + r) * 5 * H + u;
      save_g[sg] = si; save_g[sg + H] = tj; save_g[sg + 2 * H] = sf; save_g[sg + 3 * H] = so; save_g[sg + 4 * H] = tc;
    }
  }
}

// ---- generic tiled SGEMM: C = alpha * op(A) op(B) + beta * C ------------------
// op(A) is [M,K], op(B) is [K,N]; row-major storage with leading dimensions.
// gridDim.z > 1: split-K -- slice z covers k in [z * kchunk, (z+1) * kchunk) and ADDS alpha * partial into C atomically
// (C must already hold beta * C_old; beta is ignored)
__global__ void __launch_bounds__(256) sgemm_kernel(bool ta, bool tb, int M, int N, int Kd, float alpha,
                                                    const float* __restrict__ A, int lda,
                                                    const float* __restrict__ Bm, int ldb, float beta,
                                                    float* __restrict__ C, int ldc, int kchunk) {
  __shared__ __align__(16) float As[BK][64 + 4];
  __shared__ __align__(16) float Bs[BK][64 + 4];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  const int kbeg = blockIdx.z * kchunk;
  const int Kend = min(Kd, kbeg + kchunk);
  for (int k0 = kbeg; k0 < Kend; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;
      // A tile: 64 rows x 16 k
      int r, kk;
      if (!ta) { r = idx >> 4; kk = idx & 15; } else { kk = idx >> 6; r = idx & 63; }
      int gm = m0 + r, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < Kend) v = ta ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      As[kk][r] = v;
      int c2, kk2;
      if (!tb) { kk2 = idx >> 6; c2 = idx & 63; } else { c2 = idx >> 4; kk2 = idx & 15; }
      int gn = n0 + c2, gk2 = k0 + kk2;
      float w = 0.f;
      if (gn < N && gk2 < Kend) w = tb ? Bm[(size_t)gn * ldb + gk2] : Bm[(size_t)gk2 * ldb + gn];
      Bs[kk2][c2] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
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
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (gridDim.z > 1) { atomicAdd(C + (size_t)gm * ldc + gn, v); continue; }
      if (beta != 0.f) v += beta * C[(size_t)gm * ldc + gn];
      C[(size_t)gm * ldc + gn] = v;
    }
  }
}

// one warp per row: y = x * rsqrt(max(sum x^2, 1e-12))   (tf.nn.l2_normalize)
__global__ void l2norm_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int cols) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * cols;
  float ss = 0.f;
  for (int j = lane; j < cols; j += 32) { float v = xr[j]; ss = fmaf(v, v, ss); }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  float inv = rsqrtf(fmaxf(ss, 1e-12f));
  // one Newton step so the result matches 1/sqrt() to fp32 rounding
  inv = inv * (1.5f - 0.5f * fmaxf(ss, 1e-12f) * inv * inv);
  float* yr = y + (size_t)row * cols;
  for (int j = lane; j < cols; j += 32) yr[j] = xr[j] * inv;
}

}  // namespace

int lstm_forward_simt(const int32_t* tokens, int B, int T, int t_start, const float* emb, int We,
                      const LstmTower& tw, float* h0, float* h1, float* c, const float* init_h,
                      const float* init_c, float* save_h, float* save_c, float* save_g, float** h_final,
                      cudaStream_t st, int64_t* launches, const int32_t* lead, const float* pad_h, const float* pad_c) {
  const int H = tw.H;
  float* hp = nullptr;   // h_{t-1}
  float* hn = h0;
  const bool big = B >= 4096;
  for (int t = t_start; t < T; ++t) {
    const float* ih = (t == t_start) ? init_h : nullptr;
    const float* ic = (t == t_start) ? init_c : nullptr;
    if (big) {
      dim3 grid(cdiv(B, 128), cdiv(H, 16));
      lstm_step_kernel<8><<<grid, 256, 0, st>>>(tokens, T, t, emb, We, tw.K, tw.b, hp, hn, c, ih, ic, lead, pad_h, pad_c, B, H,
                                                save_h, save_c, save_g);
    } else {
      dim3 grid(cdiv(B, 64), cdiv(H, 16));
      lstm_step_kernel<4><<<grid, 256, 0, st>>>(tokens, T, t, emb, We, tw.K, tw.b, hp, hn, c, ih, ic, lead, pad_h, pad_c, B, H,
                                                save_h, save_c, save_g);
    }
    if (launches) ++*launches;
    hp = hn;
    hn = (hn == h0) ? h1 : h0;
  }
  SSE_CUDA_OK(cudaGetLastError());
  *h_final = hp;
  return SSE_OK;
}

int sgemm(bool ta, bool tb, int M, int N, int Kd, float alpha, const float* A, int lda, const float* Bm, int ldb,
          float beta, float* C, int ldc, cudaStream_t st, int64_t* launches, bool allow_split) {
  if (M <= 0 || N <= 0) return SSE_OK;
  dim3 grid(cdiv(N, 64), cdiv(M, 64));
  // few output tiles and a long contraction (the small GEMMs around the towers: u = h M, dM = h^T du, dh = du M^T): split K
  // over blockIdx.z so the machine is busy; partial sums are added atomically (C = beta * C first).  Only on request
  // (the bf16 train step): the summation order is not deterministic, which the exact fp32 paths must not inherit.
  int splits = 1;
  if (allow_split && grid.x * grid.y < 48 && Kd >= 256 && (beta == 0.f || beta == 1.f) && (beta == 1.f || ldc == N)) splits = std::min(16, Kd / 64);
  int kchunk = Kd;
  if (splits > 1) {
    kchunk = cdiv(cdiv(Kd, splits), BK) * BK;
    splits = cdiv(Kd, kchunk);
    if (beta == 0.f) SSE_CUDA_OK(cudaMemsetAsync(C, 0, (size_t)M * N * 4, st));
    grid.z = splits;
  }
  sgemm_kernel<<<grid, 256, 0, st>>>(ta, tb, M, N, Kd, alpha, A, lda, Bm, ldb, beta, C, ldc, kchunk);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

int l2norm_rows_out(const float* x, float* y, int rows, int cols, cudaStream_t st, int64_t* launches) {
  if (rows <= 0) return SSE_OK;
  l2norm_rows_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, y, rows, cols);
  if (launches) ++*launches;
  SSE_CUDA_OK(cudaGetLastError());
  return SSE_OK;
}

int l2norm_rows(float* x, int rows, int cols, cudaStream_t st, int64_t* launches) {
  return l2norm_rows_out(x, x, rows, cols, st, launches);
}

}  // namespace sse
