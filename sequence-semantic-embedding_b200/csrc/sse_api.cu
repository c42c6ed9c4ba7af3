// C ABI of libsse_b200.so (see include/sse_b200.h for the contract and the reference
// call site each entry point replaces).
#include "sse_common.cuh"
#include "sse_handle.cuh"
#include <string.h>
#include <math.h>
#include <new>

using namespace sse;

namespace sse {

static int add_param(sse_handle* h, const std::string& name, std::vector<int64_t> shape, bool trainable, bool dense_grad) {
  Param p;
  p.name = name; p.shape = shape; p.trainable = trainable;
  p.numel = 1;
  for (auto d : shape) p.numel *= d;
  size_t bytes = (size_t)std::max<int64_t>(p.numel, 1) * 4;
  cudaError_t e = cudaMalloc(&p.dev, bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(%s, %zu) failed: %s", name.c_str(), bytes, cudaGetErrorString(e)); return -1; }
  cudaMemset(p.dev, 0, bytes);
  if (trainable && dense_grad) { p.grad_off = h->grad_floats; h->grad_floats += (p.numel + 63) / 64 * 64; }
  h->params.push_back(p);
  return (int)h->params.size() - 1;
}

int find_param(sse_handle* h, const char* name) {
  for (size_t i = 0; i < h->params.size(); ++i)
    if (h->params[i].name == name) return (int)i;
  return -1;
}

static int add_lstm(sse_handle* h, const std::string& scope, int H, LstmTower& tw) {
  const int We = h->cfg.embedding_size;
  tw.H = H;
  tw.kparam = add_param(h, scope + "/rnn/basic_lstm_cell/kernel", {We + H, 4 * H}, true, true);
  tw.bparam = add_param(h, scope + "/rnn/basic_lstm_cell/bias", {4 * H}, true, true);
  if (tw.kparam < 0 || tw.bparam < 0) return SSE_ENOMEM;
  return SSE_OK;
}
static int add_cnn(sse_handle* h, const std::string& scope, CnnTower& tw) {
  const int We = h->cfg.embedding_size;
  tw.nf = h->cfg.n_cnn_filters;
  tw.sumF = 0;
  for (int i = 0; i < tw.nf; ++i) {
    int k = h->cfg.cnn_filter_sizes[i], F = h->cfg.cnn_num_filters[i];
    tw.ksize[i] = k; tw.nfilt[i] = F; tw.sumF += F;
    char nm[64];
    snprintf(nm, sizeof(nm), "/conv-maxpool-%d/", k);
    tw.wparam[i] = add_param(h, scope + nm + "W", {k, We, 1, F}, true, true);
    tw.bparam[i] = add_param(h, scope + nm + "b", {F}, true, true);
    if (tw.wparam[i] < 0 || tw.bparam[i] < 0) return SSE_ENOMEM;
  }
  return SSE_OK;
}

void refresh_pointers(sse_handle* h) {
  auto fix_l = [&](LstmTower& t) {
    if (t.kparam >= 0) t.K = h->params[t.kparam].dev;
    if (t.bparam >= 0) t.b = h->params[t.bparam].dev;
    if (t.mparam >= 0) t.M = h->params[t.mparam].dev;
  };
  auto fix_c = [&](CnnTower& t) {
    for (int i = 0; i < t.nf; ++i) { t.W[i] = h->params[t.wparam[i]].dev; t.b[i] = h->params[t.bparam[i]].dev; }
    if (t.mparam >= 0) t.M = h->params[t.mparam].dev;
  };
  fix_l(h->lstm[0]); fix_l(h->lstm[1]); fix_c(h->cnn[0]); fix_c(h->cnn[1]);
}

void invalidate_derived(sse_handle* h) {
  h->pad[0].valid = h->pad[1].valid = false;
  h->pad_tc[0].valid = h->pad_tc[1].valid = false;
  h->tct[0].valid = h->tct[1].valid = false;
  h->gemm_tw[0].valid = h->gemm_tw[1].valid = false;
  h->tct[0].ptable_valid = h->tct[1].ptable_valid = false;
  h->emb_f16_valid = false;
  h->emb_pad_valid = false;
  h->padw[0].valid = h->padw[1].valid = false;
  h->cnn_tc[0].valid = h->cnn_tc[1].valid = false;
}

bool side_is_cnn(const sse_handle* h, int side) {
  (void)side;
  return h->cfg.network_mode == SSE_MODE_SOURCE_ONLY_CNN || h->cfg.network_mode == SSE_MODE_DUAL_CNN;
}

// pad-prefix state table of a tower: S[p] = state after p leading PAD tokens from the zero
// state, computed with the SAME step kernel on a single all-PAD row (bit-identical to running
// those steps inside any batch: per-row arithmetic does not depend on the tile composition).
static int ensure_pad_table(sse_handle* h, int side, cudaStream_t st) {
  PadTable& pt = h->pad[side];
  if (pt.valid) return SSE_OK;
  const LstmTower& tw = h->lstm[side];
  const int T = h->cfg.max_seq_length, H = tw.H;
  SSE_TRY(pt.buf.ensure((size_t)T * 2 * H * 4 + (size_t)T * 4 + 3 * H * 4 + (size_t)T * 5 * H * 4));
  float* tab_h = pt.buf.as<float>();                 // [T][H]  (index t = state after t+1 PADs)
  float* tab_c = tab_h + (size_t)T * H;              // [T][H]
  float* tmp = tab_c + (size_t)T * H;                // h0,h1,c [H] each
  float* sg = tmp + 3 * H;                           // [T][5H] (unused stash)
  int32_t* toks = reinterpret_cast<int32_t*>(sg + (size_t)T * 5 * H);
  SSE_CUDA_OK(cudaMemsetAsync(toks, 0, (size_t)T * 4, st));
  float* hf = nullptr;
  SSE_TRY(lstm_forward_simt(toks, 1, T, 0, h->params[h->emb_param].dev, h->cfg.embedding_size, tw, tmp, tmp + H,
                            tmp + 2 * H, nullptr, nullptr, tab_h, tab_c, sg, &hf, st, &h->launches));
  pt.h = tab_h; pt.c = tab_c; pt.valid = true;
  return SSE_OK;
}

// the same table in the arithmetic of the tabulated-projection tensor-core kernel (lstm_ptable_kernel run on one all-PAD
// row, dumping the state it carries past every step: h as the fp16 value of the recurrence, c in fp32), so that a tile
// started from S[t0] continues bit-identically to a tile that ran the t0 PAD steps itself
static int ensure_pad_table_tc(sse_handle* h, int side, int We, int H, cudaStream_t st) {
  PadTable& pt = h->pad_tc[side];
  if (pt.valid) return SSE_OK;
  const int T = h->cfg.max_seq_length;
  SSE_TRY(pt.buf.ensure((size_t)T * 2 * H * 4 + (size_t)T * 4 + (size_t)H * 4));
  float* tab_h = pt.buf.as<float>();
  float* tab_c = tab_h + (size_t)T * H;
  float* hout = tab_c + (size_t)T * H;
  int32_t* toks = reinterpret_cast<int32_t*>(hout + H);
  SSE_CUDA_OK(cudaMemsetAsync(toks, 0, (size_t)T * 4, st));
  PadSkip ps;
  ps.dump_h = tab_h; ps.dump_c = tab_c;
  SSE_TRY(lstm_forward_ptable(toks, 1, T, 0, We, H, h->tct[side], nullptr, nullptr, ps, hout, st, &h->launches));
  pt.h = tab_h; pt.c = tab_c; pt.valid = true;
  return SSE_OK;
}

// sticky count of out-of-range token ids seen by the device pre-pass: read + reset (synchronises `st`)
int take_token_errors(sse_handle* h, cudaStream_t st, int* count) {
  int bad = 0;
  SSE_CUDA_OK(cudaMemcpyAsync(&bad, h->tok_bad, 4, cudaMemcpyDeviceToHost, st));
  SSE_CUDA_OK(cudaStreamSynchronize(st));
  if (bad) SSE_CUDA_OK(cudaMemsetAsync(h->tok_bad, 0, 4, st));
  *count = bad;
  return SSE_OK;
}
static int fail_on_token_errors(sse_handle* h, cudaStream_t st, const char* who) {
  int bad = 0;
  SSE_TRY(take_token_errors(h, st, &bad));
  if (bad) {
    set_error("%s: %d token id(s) outside [0, vocab_size=%d) (TensorFlow's gather raises InvalidArgument here; they were read as PAD_ID)",
              who, bad, h->cfg.vocab_size);
    return SSE_EINVAL;
  }
  return SSE_OK;
}

// encode B rows of one side into out [B,E]
int encode_device(sse_handle* h, int side, const int32_t* tokens_in, int B, float* out, int normalize, cudaStream_t st) {
  const sse_config& c = h->cfg;
  const int T = c.max_seq_length, We = c.embedding_size, E = c.encoding_size;
  if (B <= 0) return SSE_OK;
  const float* emb = h->params[h->emb_param].dev;
  const bool is_cnn = side_is_cnn(h, side);
  if (is_cnn && c.network_mode == SSE_MODE_SOURCE_ONLY_CNN && side != SSE_SIDE_SRC) {
    set_error("target side of source_only_cnn is a table (target_embedding/tgt_seq_embedding), not an encoder");
    return SSE_ESTATE;
  }
  if (!is_cnn && c.network_mode == SSE_MODE_SOURCE_ENCODER_ONLY && side != SSE_SIDE_SRC) {
    set_error("target side of source-encoder-only is a table (target_embedding/tgt_seq_embedding), not an encoder");
    return SSE_ESTATE;
  }
  // device pre-pass: range check (+ pad-prefix bucketing for the LSTM towers)
  const bool sort_rows = !is_cnn && h->opt_pad_skip;
  TokPrep tp;
  SSE_TRY(h->tok_ws.ensure(tok_prep_ws_bytes(B, T)));
  SSE_TRY(tok_prep(tokens_in, B, T, c.vocab_size, sort_rows, h->tok_ws.p, &tp, h->tok_bad, st, &h->launches));
  const int32_t* tokens = tp.stok;
  // rows leave the tower in sorted order: project into a scratch [B,E], then un-permute (+ l2-normalise) into `out`
  float* proj = out;
  if (tp.sorted) {
    SSE_TRY(h->proj_ws.ensure((size_t)B * E * 4));
    proj = h->proj_ws.as<float>();
  }
  auto finish = [&]() -> int {
    if (tp.sorted) return unpermute_rows(proj, out, tp.perm, B, E, normalize, st, &h->launches);
    if (normalize) return l2norm_rows(out, B, E, st, &h->launches);
    return SSE_OK;
  };
  if (is_cnn) {
    const CnnTower& tw = h->cnn[side];
    const bool cnn_tc = (h->opt_encoder == 2 || (h->opt_encoder == 0 && c.precision == SSE_PRECISION_TC)) && cnn_tc_supported(We, T, tw);
    if (h->opt_encoder == 2 && !cnn_tc) { set_error("tcgen05 CNN tower needs We%%8==0, sum(filters)%%8==0, T-k+1 <= 128"); return SSE_EINVAL; }
    if (cnn_tc) {
      CnnTc& ct = h->cnn_tc[side];
      if (!ct.valid) SSE_TRY(cnn_tc_prepare(ct, tw, We, E, st, &h->launches));
      const int slab = std::min(B, 8192);
      SSE_TRY(h->enc_ws.ensure(cnn_tc_ws_bytes(slab, T, We, tw)));
      for (int b0 = 0; b0 < B; b0 += slab) {
        const int nb = std::min(slab, B - b0);
        SSE_TRY(cnn_forward_tc(tokens + (size_t)b0 * T, nb, T, emb, We, E, tw, ct, h->enc_ws.p, proj + (size_t)b0 * E, st, &h->launches));
      }
      return finish();
    }
    int maxF = 0;
    for (int i = 0; i < tw.nf; ++i) maxF = std::max(maxF, tw.nfilt[i]);
    // process in slabs so the conv scratch stays bounded
    const int slab = std::max(1, std::min(B, (int)((256u << 20) / ((size_t)T * std::max(maxF, We) * 4))));
    size_t need = (size_t)slab * T * We * 4 + (size_t)slab * T * maxF * 4 + (size_t)slab * tw.sumF * 4;
    SSE_TRY(h->enc_ws.ensure(need));
    float* xg = h->enc_ws.as<float>();
    float* conv = xg + (size_t)slab * T * We;
    float* pool = conv + (size_t)slab * T * maxF;
    for (int b0 = 0; b0 < B; b0 += slab) {
      int nb = std::min(slab, B - b0);
      SSE_TRY(cnn_forward_ws(tokens + (size_t)b0 * T, nb, T, emb, We, tw, xg, conv, pool, nullptr, st, &h->launches));
      SSE_TRY(sgemm(false, false, nb, E, tw.sumF, 1.f, pool, tw.sumF, tw.M, E, 0.f, proj + (size_t)b0 * E, E, st,
                    &h->launches));
    }
    return finish();
  }
  const LstmTower& tw = h->lstm[side];
  const int H = tw.H;
  const bool want_tc = h->opt_encoder == 2 || (h->opt_encoder == 0 && c.precision == SSE_PRECISION_TC);
  // The tensor-core kernels tile We and H in blocks of 64.  Other sizes (the reference recipes use We in {30,40,50},
  // H = 96) run on zero-padded copies of the embedding and the LSTM kernel/bias: a padded hidden unit has z = 0, so
  // j = tanh(0) = 0 and its c and h stay exactly 0, and padded inputs meet zero weights -- the live units compute the
  // same sums (sse_model.py:236-275 semantics unchanged).
  const int Wp = (We + 63) / 64 * 64, Hp = (H + 63) / 64 * 64;
  const bool use_gemm_tower = want_tc && (h->opt_lstm_kernel == 4 || !lstm_tc_supported(Wp, Hp)) && lstm_gemm_supported(Wp, Hp);
  if (want_tc && (lstm_tc_supported(Wp, Hp) || use_gemm_tower)) {
    const bool padded = Wp != We || Hp != H;
    const float* emb_x = emb;
    const float* K_x = tw.K;
    const float* b_x = tw.b;
    if (padded) {
      PaddedWeights& pw = h->padw[side];
      if (!pw.valid) {
        if (!pw.K) {
          SSE_CUDA_OK(cudaMalloc(&pw.K, (size_t)(Wp + Hp) * 4 * Hp * 4));
          SSE_CUDA_OK(cudaMalloc(&pw.b, (size_t)4 * Hp * 4));
        }
        SSE_TRY(pad_lstm_weights(tw.K, tw.b, We, H, Wp, Hp, pw.K, pw.b, st, &h->launches));
        pw.valid = true;
      }
      if (!h->emb_pad_valid) {
        if (!h->emb_pad) SSE_CUDA_OK(cudaMalloc(&h->emb_pad, (size_t)c.vocab_size * Wp * 4));
        SSE_TRY(pad_cols(emb, c.vocab_size, We, h->emb_pad, Wp, st, &h->launches));
        h->emb_pad_valid = true;
      }
      emb_x = h->emb_pad; K_x = pw.K; b_x = pw.b;
    }
    if (use_gemm_tower) {
      // cells wider than the resident-weight kernels hold (H = 512 of BASELINE config 5): one tensor-core GEMM per step
      GemmTower& gt = h->gemm_tw[side];
      if (!gt.valid) SSE_TRY(lstm_gemm_prepare(gt, K_x, Wp, Hp, st, &h->launches));
      const int slab = std::min(B, 8192);
      SSE_TRY(h->enc_ws.ensure(lstm_gemm_ws_bytes(slab, T, Wp, Hp) + (size_t)B * Hp * 4 + 256));
      float* hout = reinterpret_cast<float*>(h->enc_ws.as<uint8_t>() + (lstm_gemm_ws_bytes(slab, T, Wp, Hp) + 255) / 256 * 256);
      for (int b0 = 0; b0 < B; b0 += slab) {
        const int nb = std::min(slab, B - b0);
        SSE_TRY(lstm_forward_gemm(tokens + (size_t)b0 * T, nb, T, emb_x, Wp, Hp, gt, b_x, h->enc_ws.p, hout + (size_t)b0 * Hp, Hp, st, &h->launches));
      }
      SSE_TRY(sgemm(false, false, B, E, H, 1.f, hout, Hp, tw.M, E, 0.f, proj, E, st, &h->launches));
      return finish();
    }
    if (!h->emb_f16_valid) {
      if (!h->emb_f16) SSE_CUDA_OK(cudaMalloc(&h->emb_f16, (size_t)c.vocab_size * Wp * 2));
      SSE_TRY(f32_to_f16(emb_x, h->emb_f16, (int64_t)c.vocab_size * Wp, st, &h->launches));
      h->emb_f16_valid = true;
    }
    TcTower& tt = h->tct[side];
    if (!tt.valid) SSE_TRY(lstm_tc_prepare(tt, K_x, b_x, Wp, Hp, st, &h->launches));
    const size_t Bpad = (size_t)cdiv(B, 128) * 128;
    SSE_TRY(h->enc_ws.ensure((Bpad * Hp + (size_t)B * Hp) * 4));
    float* cs = h->enc_ws.as<float>();
    float* hout = cs + Bpad * Hp;
    // kernel choice (lstm_kernel option): 3 = cluster kernel with the tabulated input projection (default whenever the
    // table fits), 2 = cluster kernel with resident W_x and gathered x tiles, 1 = weight-streaming kernel
    const bool cluster_ok = lstm_cluster_supported(Wp, Hp);
    const bool ptable_ok = lstm_ptable_supported(c.vocab_size, Wp, Hp);
    int kern = h->opt_lstm_kernel;
    if (kern == 2 && !cluster_ok) { set_error("cluster LSTM kernel needs H <= 256 (padded to 64/128/256), We <= 256 (We=%d H=%d)", We, H); return SSE_EINVAL; }
    if (kern == 3 && !ptable_ok) { set_error("table LSTM kernel needs H <= 256 (padded to 64/128/256) and V*4H*4 <= 2 GiB (V=%d H=%d)", c.vocab_size, H); return SSE_EINVAL; }
    // measured (We=H=256, T=50): table kernel 0.22 ms at 600 rows, 0.67 ms at 4800; the weight-streaming kernel is
    // flat ~0.8 ms up to ~5k rows and wins once every SM holds a full 128-row tile (1.44 vs 2.2 ms at 18944 rows)
    if (kern == 0) kern = (ptable_ok && B <= 8192) ? 3 : ((cluster_ok && B <= 1024) ? 2 : 1);
    if (kern == 3 && !tt.ptable_valid) SSE_TRY(lstm_ptable_prepare(tt, emb_x, c.vocab_size, K_x, Wp, Hp, st, &h->launches));
    PadSkip ps;
    if (tp.sorted) {
      // kernel 3 continues bit-identically from a table in its own arithmetic; the other two start from the fp32 table
      // (equal within the tensor-core tolerance; not available for padded shapes: they run all T steps)
      if (kern == 3) { SSE_TRY(ensure_pad_table_tc(h, side, Wp, Hp, st)); ps.pad_h = h->pad_tc[side].h; ps.pad_c = h->pad_tc[side].c; ps.lead_sorted = tp.lead_sorted; }
      else if (!padded) { SSE_TRY(ensure_pad_table(h, side, st)); ps.pad_h = h->pad[side].h; ps.pad_c = h->pad[side].c; ps.lead_sorted = tp.lead_sorted; }
    }
    if (kern == 3) {
      SSE_TRY(lstm_forward_ptable(tokens, B, T, 0, Wp, Hp, tt, nullptr, nullptr, ps, hout, st, &h->launches, h->opt_cluster_rows, h->num_sms));
    } else if (kern == 2) {
      SSE_TRY(lstm_forward_cluster(tokens, B, T, 0, h->emb_f16, Wp, Hp, tt, nullptr, nullptr, ps, hout, st, &h->launches));
    } else {
      SSE_TRY(lstm_forward_tc(tokens, B, T, 0, h->emb_f16, Wp, Hp, tt, nullptr, nullptr, ps, cs, hout, st, &h->launches));
    }
    if (project_rows_supported(B, H, E))       // query batches: projection + l2-norm + un-permutation in one launch
      return project_rows(hout, Hp, tw.M, H, E, B, tp.sorted ? tp.perm : nullptr, normalize, out, st, &h->launches);
    SSE_TRY(sgemm(false, false, B, E, H, 1.f, hout, Hp, tw.M, E, 0.f, proj, E, st, &h->launches));
    return finish();
  } else if (h->opt_encoder == 2) {
    set_error("tcgen05 encoder needs We <= 256 and H <= 256 (We=%d H=%d)", We, H);
    return SSE_EINVAL;
  }
  SSE_TRY(h->enc_ws.ensure((size_t)B * H * 3 * 4));
  float* h0 = h->enc_ws.as<float>();
  float* h1 = h0 + (size_t)B * H;
  float* cc = h1 + (size_t)B * H;
  float* hf = nullptr;
  const int32_t* lead = nullptr;
  const float *ph = nullptr, *pc = nullptr;
  if (tp.sorted) {
    SSE_TRY(ensure_pad_table(h, side, st));
    lead = tp.lead_sorted; ph = h->pad[side].h; pc = h->pad[side].c;
  }
  SSE_TRY(lstm_forward_simt(tokens, B, T, 0, emb, We, tw, h0, h1, cc, nullptr, nullptr, nullptr, nullptr, nullptr, &hf, st,
                            &h->launches, lead, ph, pc));
  SSE_TRY(sgemm(false, false, B, E, H, 1.f, hf, H, tw.M, E, 0.f, proj, E, st, &h->launches));
  return finish();
}

int search_device(sse_handle* h, const float* q, int Q, int k, float* scores, int32_t* idx, cudaStream_t st, int out_stride = 0) {
  if (out_stride <= 0) out_stride = k;
  if (!h->index_f32) { set_error("no target index registered (call sse_index_set / sse_index_build first)"); return SSE_ESTATE; }
  if (k < 1 || k > SSE_MAX_TOPK) { set_error("k=%d out of range [1,%d]", k, SSE_MAX_TOPK); return SSE_EINVAL; }
  const int E = h->cfg.encoding_size;
  bool use_tc = h->opt_search == 2 || (h->opt_search == 0 && h->cfg.precision == SSE_PRECISION_TC);
  if (use_tc && !search_tc_supported(E, h->index_n, k)) {
    if (h->opt_search == 2) { set_error("tcgen05 search needs E<=512, k<=128, N>=8192 (E=%d k=%d N=%lld)", E, k, (long long)h->index_n); return SSE_EINVAL; }
    use_tc = false;
  }
  if (use_tc) {
    if (!h->tc.tmap_ok) SSE_TRY(search_tc_prepare(h->tc, h->index_f32, h->index_n, E, st, &h->launches));
    const int maxq = search_tc_max_rows(E);
    for (int q0 = 0; q0 < Q; q0 += maxq) {
      int nq = std::min(maxq, Q - q0);
      SSE_TRY(search_tc(q + (size_t)q0 * E, nq, E, h->index_f32, h->tc, h->index_off, k, scores + (size_t)q0 * out_stride,
                        idx + (size_t)q0 * out_stride, h->search_ws,
                        h->opt_search_ctas > 0 ? std::min(h->opt_search_ctas, h->num_sms) : h->num_sms, st, &h->launches, out_stride,
                        h->opt_search_ctas > 0 ? h->opt_search_late : 0, h->opt_search_late_share));
    }
    return SSE_OK;
  }
  return search_simt(q, Q, E, h->index_f32, h->index_n, h->index_off, k, scores, idx, h->search_ws, h->num_sms, st,
                     &h->launches, out_stride);
}

}  // namespace sse

// =============================================================================
extern "C" {

static void sampler_free(sse_handle* h);

const char* sse_last_error(void) { return get_error(); }
const char* sse_version(void) { return "sse_b200 0.1 sm_100a"; }

int sse_create(const sse_config* cfg_in, sse_handle** out) {
  if (!cfg_in || !out) { set_error("sse_create: null argument"); return SSE_EINVAL; }
  if (cfg_in->struct_size != (int32_t)sizeof(sse_config)) { set_error("sse_create: struct_size %d != %zu", cfg_in->struct_size, sizeof(sse_config)); return SSE_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device visible: libsse_b200 has no CPU fallback");
    return SSE_ENODEVICE;
  }
  sse_config c = *cfg_in;
  if (c.device < 0 || c.device >= ndev) { set_error("device %d out of range (%d visible)", c.device, ndev); return SSE_EINVAL; }
  if (c.vocab_size <= 0 || c.embedding_size <= 0 || c.encoding_size <= 0 || c.max_seq_length <= 1) { set_error("bad model dimensions"); return SSE_EINVAL; }
  if (c.network_mode < 0 || c.network_mode > SSE_MODE_DUAL_CNN) { set_error("Unsupported network mode: %d", c.network_mode); return SSE_EINVAL; }
  bool cnn_mode = c.network_mode == SSE_MODE_SOURCE_ONLY_CNN || c.network_mode == SSE_MODE_DUAL_CNN;
  if (cnn_mode && c.n_cnn_filters == 0) {   // reference layout, sse_model.py:183-184
    c.n_cnn_filters = 4;
    const int ks[4] = {2, 3, 4, 5}, fs[4] = {256, 128, 128, 64};
    for (int i = 0; i < 4; ++i) { c.cnn_filter_sizes[i] = ks[i]; c.cnn_num_filters[i] = fs[i]; }
  }
  if (c.n_cnn_filters < 0 || c.n_cnn_filters > SSE_MAX_CNN_FILTERS) { set_error("bad n_cnn_filters"); return SSE_EINVAL; }
  if (!cnn_mode && (c.src_cell_size <= 0 || (c.network_mode == SSE_MODE_DUAL_ENCODER && c.tgt_cell_size <= 0))) { set_error("bad cell size"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(c.device));
  sse_handle* h = new (std::nothrow) sse_handle();
  if (!h) return SSE_ENOMEM;
  h->cfg = c;
  cudaDeviceProp prop;
  SSE_CUDA_OK(cudaGetDeviceProperties(&prop, c.device));
  h->num_sms = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->learning_rate = c.learning_rate;
  h->global_step = 0;

  const int E = c.encoding_size;
  h->emb_param = add_param(h, "word_embedding", {c.vocab_size, c.embedding_size}, true, false);
  int rc = h->emb_param < 0 ? SSE_ENOMEM : SSE_OK;
  auto addM = [&](const std::string& name, int rows) { return add_param(h, name, {rows, E}, true, true); };
  if (rc == SSE_OK) switch (c.network_mode) {
    case SSE_MODE_DUAL_ENCODER:
      rc = add_lstm(h, "source_encoder", c.src_cell_size, h->lstm[0]);
      if (rc == SSE_OK) h->lstm[0].mparam = addM("source_encoder/src_M", c.src_cell_size);
      if (rc == SSE_OK) rc = add_lstm(h, "target_encoder", c.tgt_cell_size, h->lstm[1]);
      if (rc == SSE_OK) h->lstm[1].mparam = addM("target_encoder/tgt_M", c.tgt_cell_size);
      break;
    case SSE_MODE_SHARED_ENCODER:
      rc = add_lstm(h, "shared_encoder", c.src_cell_size, h->lstm[0]);
      if (rc == SSE_OK) {
        h->lstm[0].mparam = addM("shared_encoder/src_M", c.src_cell_size);
        h->lstm[1] = h->lstm[0];   // same cell variables, separate projection (sse_model.py:261-275)
        h->lstm[1].mparam = addM("shared_encoder/tgt_M", c.src_cell_size);
      }
      break;
    case SSE_MODE_SOURCE_ENCODER_ONLY:
      rc = add_lstm(h, "source_only_encoder", c.src_cell_size, h->lstm[0]);
      if (rc == SSE_OK) h->lstm[0].mparam = addM("source_only_encoder/src_M", c.src_cell_size);
      if (rc == SSE_OK) h->tgt_table_param = add_param(h, "target_embedding/tgt_seq_embedding", {c.target_space_size, E}, true, false);
      break;
    case SSE_MODE_SOURCE_ONLY_CNN:
      rc = add_cnn(h, "source_only_cnn", h->cnn[0]);
      if (rc == SSE_OK) h->cnn[0].mparam = addM("source_only_cnn/src_M", h->cnn[0].sumF);
      if (rc == SSE_OK) h->tgt_table_param = add_param(h, "target_embedding/tgt_seq_embedding", {c.target_space_size, E}, true, false);
      break;
    case SSE_MODE_DUAL_CNN:
      rc = add_cnn(h, "source_cnn", h->cnn[0]);
      if (rc == SSE_OK) h->cnn[0].mparam = addM("source_cnn/src_M", h->cnn[0].sumF);
      if (rc == SSE_OK) rc = add_cnn(h, "target_cnn", h->cnn[1]);
      if (rc == SSE_OK) h->cnn[1].mparam = addM("target_cnn/tgt_M", h->cnn[1].sumF);
      break;
  }
  if (rc == SSE_OK) {
    for (auto& p : h->params)
      if (!p.dev) rc = SSE_ENOMEM;
  }
  if (rc == SSE_OK) {
    // Adagrad slots (initial_accumulator_value = 0.1), same order as the variables
    size_t nvar = h->params.size();
    for (size_t i = 0; i < nvar && rc == SSE_OK; ++i) {
      int s = add_param(h, h->params[i].name + "/Adagrad", h->params[i].shape, false, false);
      if (s < 0) { rc = SSE_ENOMEM; break; }
      rc = fill_f32(h->params[s].dev, h->params[s].numel, 0.1f, 0, &h->launches);
    }
    h->n_vars = (int)nvar;
  }
  if (rc == SSE_OK && (cudaMalloc(&h->tok_bad, 4) != cudaSuccess || cudaMemset(h->tok_bad, 0, 4) != cudaSuccess)) rc = SSE_ENOMEM;
  if (rc != SSE_OK) { sse_destroy(h); return rc; }
  refresh_pointers(h);
  cudaDeviceSynchronize();
  *out = h;
  return SSE_OK;
}

int sse_destroy(sse_handle* h) {
  if (!h) return SSE_OK;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  if (h->train_graph) cudaGraphExecDestroy(h->train_graph);
  if (h->train_side) cudaStreamDestroy(h->train_side);
  if (h->train_main) cudaStreamDestroy(h->train_main);
  for (int i = 0; i < 2; ++i) { if (h->train_chain[i]) cudaStreamDestroy(h->train_chain[i]); for (int j = 0; j < 4; ++j) if (h->train_ev2[i][j]) cudaEventDestroy(h->train_ev2[i][j]); }
  for (int i = 0; i < 6; ++i) if (h->train_ev[i]) cudaEventDestroy(h->train_ev[i]);
  for (auto& p : h->params) if (p.dev) cudaFree(p.dev);
  h->enc_ws.release(); h->search_ws.release(); h->io_ws.release(); h->train_ws.release();
  h->pad[0].buf.release(); h->pad[1].buf.release();
  h->pad_tc[0].buf.release(); h->pad_tc[1].buf.release();
  h->tok_ws.release(); h->proj_ws.release(); h->train_tc_ws.release();
  if (h->tok_bad) cudaFree(h->tok_bad);
  sampler_free(h);
  if (h->index_f32 && h->index_owned) cudaFree(h->index_f32);
  if (h->grad_arena) cudaFree(h->grad_arena);
  search_tc_release(h->tc);
  lstm_tc_release(h->tct[0]); lstm_tc_release(h->tct[1]);
  lstm_ptable_release(h->tct[0]); lstm_ptable_release(h->tct[1]);
  cnn_tc_release(h->cnn_tc[0]); cnn_tc_release(h->cnn_tc[1]);
  lstm_gemm_release(h->gemm_tw[0]); lstm_gemm_release(h->gemm_tw[1]);
  if (h->emb_f16) cudaFree(h->emb_f16);
  if (h->emb_pad) cudaFree(h->emb_pad);
  for (int s2 = 0; s2 < 2; ++s2) { if (h->padw[s2].K) cudaFree(h->padw[s2].K); if (h->padw[s2].b) cudaFree(h->padw[s2].b); }
  delete h;
  return SSE_OK;
}

int sse_param_count(sse_handle* h) { return h ? (int)h->params.size() : 0; }

int sse_param_info(sse_handle* h, int i, char* name_out, int name_cap, int64_t* shape_out, int* ndim_out) {
  if (!h || i < 0 || i >= (int)h->params.size()) { set_error("sse_param_info: index out of range"); return SSE_EINVAL; }
  const Param& p = h->params[i];
  if (name_out && name_cap > 0) { strncpy(name_out, p.name.c_str(), name_cap - 1); name_out[name_cap - 1] = 0; }
  if (ndim_out) *ndim_out = (int)p.shape.size();
  if (shape_out) for (size_t d = 0; d < p.shape.size() && d < 4; ++d) shape_out[d] = p.shape[d];
  return SSE_OK;
}

int sse_set_param(sse_handle* h, const char* name, const void* ptr, const int64_t* shape, int ndim) {
  if (!h || !name || !ptr) { set_error("sse_set_param: null argument"); return SSE_EINVAL; }
  int i = find_param(h, name);
  if (i < 0) { set_error("sse_set_param: unknown variable '%s'", name); return SSE_EINVAL; }
  Param& p = h->params[i];
  if (shape) {
    bool ok = ndim == (int)p.shape.size();
    for (int d = 0; ok && d < ndim; ++d) ok = shape[d] == p.shape[d];
    if (!ok) { set_error("sse_set_param: shape mismatch for '%s'", name); return SSE_EINVAL; }
  }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  SSE_CUDA_OK(cudaMemcpy(p.dev, ptr, (size_t)p.numel * 4, cudaMemcpyDefault));
  invalidate_derived(h);   // weights changed: pad-prefix tables / fp16 copies are stale
  return SSE_OK;
}

int sse_get_param(sse_handle* h, const char* name, void* host_dst, int64_t nbytes) {
  if (!h || !name || !host_dst) { set_error("sse_get_param: null argument"); return SSE_EINVAL; }
  int i = find_param(h, name);
  if (i < 0) { set_error("sse_get_param: unknown variable '%s'", name); return SSE_EINVAL; }
  const Param& p = h->params[i];
  if (nbytes != p.numel * 4) { set_error("sse_get_param: '%s' is %lld bytes, buffer is %lld", name, (long long)p.numel * 4, (long long)nbytes); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  SSE_CUDA_OK(cudaMemcpy(host_dst, p.dev, (size_t)nbytes, cudaMemcpyDeviceToHost));
  return SSE_OK;
}

int sse_encode(sse_handle* h, int side, const int32_t* tokens_dev, int B, float* out_dev, int normalize, void* stream) {
  if (!h || !tokens_dev || !out_dev || B < 0 || (side != 0 && side != 1)) { set_error("sse_encode: bad argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  return encode_device(h, side, tokens_dev, B, out_dev, normalize, (cudaStream_t)stream);
}

int sse_encode_host(sse_handle* h, int side, const int32_t* tokens_host, int B, float* out_host, int normalize) {
  if (!h || !tokens_host || !out_host || B < 0 || (side != 0 && side != 1)) { set_error("sse_encode_host: bad argument"); return SSE_EINVAL; }
  if (B == 0) return SSE_OK;
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  const int T = h->cfg.max_seq_length, E = h->cfg.encoding_size;
  size_t tb = ((size_t)B * T * 4 + 255) / 256 * 256;
  SSE_TRY(h->io_ws.ensure(tb + (size_t)B * E * 4));
  int32_t* dt = h->io_ws.as<int32_t>();
  float* dout = reinterpret_cast<float*>(h->io_ws.as<uint8_t>() + tb);
  cudaStream_t st = 0;
  SSE_CUDA_OK(cudaMemcpyAsync(dt, tokens_host, (size_t)B * T * 4, cudaMemcpyHostToDevice, st));
  SSE_TRY(encode_device(h, side, dt, B, dout, normalize, st));
  SSE_CUDA_OK(cudaMemcpyAsync(out_host, dout, (size_t)B * E * 4, cudaMemcpyDeviceToHost, st));
  return fail_on_token_errors(h, st, "sse_encode_host");
}

int sse_index_set(sse_handle* h, const float* tgt, int64_t n_local, int64_t global_offset) {
  if (!h || (!tgt && n_local > 0) || n_local < 0) { set_error("sse_index_set: bad argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  const int E = h->cfg.encoding_size;
  if (h->index_f32 && h->index_owned) cudaFree(h->index_f32);
  h->index_f32 = nullptr; h->index_n = 0;
  search_tc_release(h->tc);
  if (n_local == 0) return SSE_OK;
  size_t bytes = (size_t)n_local * E * 4;
  cudaError_t e = cudaMalloc(&h->index_f32, bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(index, %zu) failed: %s", bytes, cudaGetErrorString(e)); return SSE_ENOMEM; }
  h->index_owned = true;
  SSE_CUDA_OK(cudaMemcpy(h->index_f32, tgt, bytes, cudaMemcpyDefault));
  h->index_n = n_local; h->index_off = global_offset;
  return SSE_OK;
}

int sse_index_build(sse_handle* h, const int32_t* tgt_tokens, int64_t n_local, int64_t global_offset, int batch) {
  if (!h || (!tgt_tokens && n_local > 0) || n_local < 0) { set_error("sse_index_build: bad argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  const int E = h->cfg.encoding_size, T = h->cfg.max_seq_length;
  if (batch <= 0) batch = 10000;   // sse_index.py:100 default batchsize
  if (h->index_f32 && h->index_owned) cudaFree(h->index_f32);
  h->index_f32 = nullptr; h->index_n = 0;
  search_tc_release(h->tc);
  if (n_local == 0) return SSE_OK;
  size_t bytes = (size_t)n_local * E * 4;
  cudaError_t e = cudaMalloc(&h->index_f32, bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(index, %zu) failed: %s", bytes, cudaGetErrorString(e)); return SSE_ENOMEM; }
  h->index_owned = true;
  SSE_TRY(h->io_ws.ensure((size_t)batch * T * 4));
  cudaStream_t st = 0;
  for (int64_t r0 = 0; r0 < n_local; r0 += batch) {
    int nb = (int)std::min<int64_t>(batch, n_local - r0);
    SSE_CUDA_OK(cudaMemcpyAsync(h->io_ws.p, tgt_tokens + (size_t)r0 * T, (size_t)nb * T * 4, cudaMemcpyDefault, st));
    SSE_TRY(encode_device(h, SSE_SIDE_TGT, h->io_ws.as<int32_t>(), nb, h->index_f32 + (size_t)r0 * E, 1, st));
  }
  h->index_n = n_local; h->index_off = global_offset;
  return fail_on_token_errors(h, st, "sse_index_build");
}

int sse_index_get(sse_handle* h, int64_t row0, int64_t n, float* host_dst) {
  if (!h || !host_dst || row0 < 0 || n < 0 || row0 + n > h->index_n) { set_error("sse_index_get: range out of bounds"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  SSE_CUDA_OK(cudaMemcpy(host_dst, h->index_f32 + (size_t)row0 * h->cfg.encoding_size, (size_t)n * h->cfg.encoding_size * 4, cudaMemcpyDeviceToHost));
  return SSE_OK;
}

int sse_search(sse_handle* h, const float* q_dev, int Q, int k, float* scores_dev, int32_t* idx_dev, void* stream) {
  if (!h || !q_dev || !scores_dev || !idx_dev || Q < 0) { set_error("sse_search: bad argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  return search_device(h, q_dev, Q, k, scores_dev, idx_dev, (cudaStream_t)stream);
}

int sse_search_packed(sse_handle* h, const float* q_dev, int Q, int k, float* packed_dev, void* stream) {
  if (!h || !q_dev || !packed_dev || Q < 0) { set_error("sse_search_packed: bad argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  return search_device(h, q_dev, Q, k, packed_dev, reinterpret_cast<int32_t*>(packed_dev) + k, (cudaStream_t)stream, 2 * k);
}

int sse_merge_packed(sse_handle* h, const float* gathered_dev, int G, int Q, int k, float* scores_dev, int32_t* idx_dev, void* stream) {
  if (!h || !gathered_dev || !scores_dev || !idx_dev || G < 1 || Q < 0 || k < 1) { set_error("sse_merge_packed: bad argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  return merge_topk_strided(gathered_dev, reinterpret_cast<const int32_t*>(gathered_dev) + k, Q, G, (int64_t)Q * 2 * k, 2 * k, k, k,
                            scores_dev, idx_dev, k, (cudaStream_t)stream, &h->launches);
}

int sse_search_stats(sse_handle* h, int64_t* candidates, int* rows, int* fallback_rows, int* items) {
  if (!h) return SSE_EINVAL;
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  SearchStats ss;
  SSE_TRY(search_tc_stats(h->tc, &ss));
  if (candidates) *candidates = ss.candidates;
  if (rows) *rows = ss.rows;
  if (fallback_rows) *fallback_rows = ss.fallback_rows;
  if (items) *items = ss.items;
  return SSE_OK;
}

int sse_merge_topk(sse_handle* h, const float* cand_scores_dev, const int32_t* cand_idx_dev, int Q, int C, int k,
                   float* scores_dev, int32_t* idx_dev, void* stream) {
  if (!h || !cand_scores_dev || !cand_idx_dev || !scores_dev || !idx_dev || Q < 0 || C < 1 || k < 1) { set_error("sse_merge_topk: bad argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  return merge_topk(cand_scores_dev, cand_idx_dev, Q, C, k, scores_dev, idx_dev, (cudaStream_t)stream, &h->launches);
}

int sse_query_host(sse_handle* h, const int32_t* tokens_host, int Q, int k, int normalize, float* scores_host,
                   int32_t* idx_host) {
  if (!h || !tokens_host || !scores_host || !idx_host || Q < 0) { set_error("sse_query_host: bad argument"); return SSE_EINVAL; }
  if (Q == 0) return SSE_OK;
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  const int T = h->cfg.max_seq_length, E = h->cfg.encoding_size;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o_tok = 0, o_enc = al((size_t)Q * T * 4), o_s = o_enc + al((size_t)Q * E * 4), o_i = o_s + al((size_t)Q * k * 4);
  SSE_TRY(h->io_ws.ensure(o_i + al((size_t)Q * k * 4)));
  uint8_t* w = h->io_ws.as<uint8_t>();
  int32_t* dt = reinterpret_cast<int32_t*>(w + o_tok);
  float* enc = reinterpret_cast<float*>(w + o_enc);
  float* ds = reinterpret_cast<float*>(w + o_s);
  int32_t* di = reinterpret_cast<int32_t*>(w + o_i);
  cudaStream_t st = 0;
  SSE_CUDA_OK(cudaMemcpyAsync(dt, tokens_host, (size_t)Q * T * 4, cudaMemcpyHostToDevice, st));
  SSE_TRY(encode_device(h, SSE_SIDE_SRC, dt, Q, enc, normalize, st));
  SSE_TRY(search_device(h, enc, Q, k, ds, di, st));
  SSE_CUDA_OK(cudaMemcpyAsync(scores_host, ds, (size_t)Q * k * 4, cudaMemcpyDeviceToHost, st));
  SSE_CUDA_OK(cudaMemcpyAsync(idx_host, di, (size_t)Q * k * 4, cudaMemcpyDeviceToHost, st));
  return fail_on_token_errors(h, st, "sse_query_host");
}

int sse_topk_batch(sse_handle* h, const float* q_dev, int Q, const float* tgt_dev, int64_t n_tgt, int k, int normalize_scores,
                   float* scores_dev, int32_t* idx_dev, void* stream) {
  if (!h || !q_dev || !tgt_dev || !scores_dev || !idx_dev || Q < 0 || n_tgt < 0) { set_error("sse_topk_batch: bad argument"); return SSE_EINVAL; }
  if (k < 1 || k > SSE_MAX_TOPK) { set_error("k=%d out of range [1,%d]", k, SSE_MAX_TOPK); return SSE_EINVAL; }
  if (n_tgt < k) { set_error("input must have at least k columns (k=%d, %lld targets)", k, (long long)n_tgt); return SSE_EINVAL; }   // TF top_k's error
  if (Q == 0) return SSE_OK;
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  SSE_TRY(search_simt(q_dev, Q, h->cfg.encoding_size, tgt_dev, n_tgt, 0, k, scores_dev, idx_dev, h->search_ws, h->num_sms, st, &h->launches));
  if (normalize_scores) SSE_TRY(l2norm_rows(scores_dev, Q, k, st, &h->launches));
  return SSE_OK;
}

int sse_token_errors(sse_handle* h, int64_t* count_out, void* stream) {
  if (!h || !count_out) { set_error("sse_token_errors: null argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  int bad = 0;
  SSE_TRY(take_token_errors(h, (cudaStream_t)stream, &bad));
  *count_out = bad;
  return SSE_OK;
}

int sse_debug_gemm_tc(sse_handle* h, const float* a_dev, const float* b_dev, int M, int N, int K, int fmt, int split_k, float alpha, float beta,
                      float* d_dev, void* stream) {
  if (!h || !a_dev || !b_dev || !d_dev || M < 1 || N < 1 || K < 1 || (fmt != 0 && fmt != 1)) { set_error("sse_debug_gemm_tc: bad argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ldk = (K + 7) / 8 * 8;
  SSE_TRY(h->io_ws.ensure(((size_t)M + N) * ldk * 2 + 512));
  uint16_t* a16 = h->io_ws.as<uint16_t>();
  uint16_t* b16 = a16 + ((size_t)M * ldk + 127) / 128 * 128;
  SSE_TRY(convert_to_16(a_dev, M, K, K, a16, ldk, fmt, st, &h->launches));
  SSE_TRY(convert_to_16(b_dev, N, K, K, b16, ldk, fmt, st, &h->launches));
  return gemm_tc(a16, ldk, b16, ldk, M, N, K, alpha, beta, d_dev, N, fmt, split_k, nullptr, 0, st, &h->launches);
}

static void sampler_free(sse_handle* h) {
  if (h->smp_src) cudaFree(h->smp_src);
  if (h->smp_ver) cudaFree(h->smp_ver);
  if (h->smp_tgt) cudaFree(h->smp_tgt);
  if (h->smp_off) cudaFree(h->smp_off);
  h->smp_src = h->smp_ver = h->smp_tgt = nullptr; h->smp_off = nullptr; h->smp_P = h->smp_N = 0;
}

int sse_sampler_set(sse_handle* h, const int32_t* src_rows, int64_t n_pos, const int64_t* ver_off, const int32_t* ver_rows, const int32_t* tgt_rows,
                    int64_t n_tgt) {
  if (!h || !src_rows || !ver_off || !ver_rows || !tgt_rows || n_pos < 1 || n_tgt < 2) { set_error("sse_sampler_set: bad argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  const int T = h->cfg.max_seq_length;
  const int64_t nnz = ver_off[n_pos];
  for (int64_t i = 0; i < n_pos; ++i)
    if (ver_off[i + 1] <= ver_off[i]) { set_error("sse_sampler_set: positive %lld has no verified target", (long long)i); return SSE_EINVAL; }
  for (int64_t x = 0; x < nnz; ++x)
    if (ver_rows[x] < 0 || ver_rows[x] >= n_tgt) { set_error("sse_sampler_set: verified target row out of range"); return SSE_EINVAL; }
  sampler_free(h);
  SSE_CUDA_OK(cudaMalloc(&h->smp_src, (size_t)n_pos * T * 4));
  SSE_CUDA_OK(cudaMalloc(&h->smp_off, (size_t)(n_pos + 1) * 8));
  SSE_CUDA_OK(cudaMalloc(&h->smp_ver, (size_t)nnz * 4));
  SSE_CUDA_OK(cudaMalloc(&h->smp_tgt, (size_t)n_tgt * T * 4));
  SSE_CUDA_OK(cudaMemcpy(h->smp_src, src_rows, (size_t)n_pos * T * 4, cudaMemcpyHostToDevice));
  SSE_CUDA_OK(cudaMemcpy(h->smp_off, ver_off, (size_t)(n_pos + 1) * 8, cudaMemcpyHostToDevice));
  SSE_CUDA_OK(cudaMemcpy(h->smp_ver, ver_rows, (size_t)nnz * 4, cudaMemcpyHostToDevice));
  SSE_CUDA_OK(cudaMemcpy(h->smp_tgt, tgt_rows, (size_t)n_tgt * T * 4, cudaMemcpyHostToDevice));
  h->smp_P = n_pos; h->smp_N = n_tgt;
  return SSE_OK;
}

int sse_sampler_batch(sse_handle* h, int64_t start, int batch_size, uint64_t seed, uint64_t step, int32_t* src_dev, int32_t* tgt_dev, float* labels_dev,
                      int* rows_out, void* stream) {
  if (!h || !src_dev || !tgt_dev || !labels_dev || batch_size < 1 || start < 0) { set_error("sse_sampler_batch: bad argument"); return SSE_EINVAL; }
  if (!h->smp_src) { set_error("sse_sampler_batch: no corpus registered (sse_sampler_set)"); return SSE_ESTATE; }
  if (start >= h->smp_P) { set_error("sse_sampler_batch: start %lld beyond the %lld positives", (long long)start, (long long)h->smp_P); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  const int B = (int)std::min<int64_t>(batch_size, h->smp_P - start);        // short window near the end (data.py:97-98)
  SSE_TRY(sample_train_batch(h->smp_src, h->smp_off, h->smp_ver, h->smp_tgt, h->smp_N, h->cfg.max_seq_length, start, B, seed, step, src_dev, tgt_dev,
                             labels_dev, (cudaStream_t)stream, &h->launches));
  if (rows_out) *rows_out = 2 * B;
  return SSE_OK;
}

int sse_l2_normalize_rows(sse_handle* h, float* x_dev, int rows, int cols, void* stream) {
  if (!h || !x_dev || rows < 0 || cols < 1) { set_error("sse_l2_normalize_rows: bad argument"); return SSE_EINVAL; }
  SSE_CUDA_OK(cudaSetDevice(h->cfg.device));
  return l2norm_rows(x_dev, rows, cols, (cudaStream_t)stream, &h->launches);
}

int sse_lr_decay(sse_handle* h) {
  if (!h) return SSE_EINVAL;
  // learning_rate.assign(max(learning_rate * decay_factor, 1e-3)), sse_model.py:123-124 (fp32 arithmetic)
  float v = h->learning_rate * h->cfg.learning_rate_decay_factor;
  h->learning_rate = v > 1e-3f ? v : 1e-3f;
  return SSE_OK;
}
int sse_get_scalars(sse_handle* h, float* learning_rate, int64_t* global_step) {
  if (!h) return SSE_EINVAL;
  if (learning_rate) *learning_rate = h->learning_rate;
  if (global_step) *global_step = h->global_step;
  return SSE_OK;
}
int sse_set_scalars(sse_handle* h, float learning_rate, int64_t global_step) {
  if (!h) return SSE_EINVAL;
  h->learning_rate = learning_rate; h->global_step = global_step;
  return SSE_OK;
}

int64_t sse_launch_count(sse_handle* h) { return h ? h->launches : 0; }

int sse_set_option(sse_handle* h, const char* key, int value) {
  if (!h || !key) return SSE_EINVAL;
  if (!strcmp(key, "search")) { if (value < 0 || value > 2) return SSE_EINVAL; h->opt_search = value; return SSE_OK; }
  if (!strcmp(key, "encoder")) { if (value < 0 || value > 2) return SSE_EINVAL; h->opt_encoder = value; return SSE_OK; }
  if (!strcmp(key, "train")) { if (value < 0 || value > 2) return SSE_EINVAL; h->opt_train = value; return SSE_OK; }
  if (!strcmp(key, "cluster_rows")) { if (value != 0 && value != 64 && value != 128) return SSE_EINVAL; h->opt_cluster_rows = value; return SSE_OK; }
  if (!strcmp(key, "pad_skip")) { h->opt_pad_skip = value != 0; return SSE_OK; }
  if (!strcmp(key, "lstm_kernel")) { if (value < 0 || value > 4) return SSE_EINVAL; h->opt_lstm_kernel = value; return SSE_OK; }
  if (!strcmp(key, "search_ctas")) { if (value < 0) return SSE_EINVAL; h->opt_search_ctas = value; return SSE_OK; }
  if (!strcmp(key, "search_late_ctas")) { if (value < 0 || value > 1024) return SSE_EINVAL; h->opt_search_late = value; return SSE_OK; }
  if (!strcmp(key, "search_late_share")) { if (value < 0 || value > 100) return SSE_EINVAL; h->opt_search_late_share = value; return SSE_OK; }
  set_error("sse_set_option: unknown key '%s'", key);
  return SSE_EINVAL;
}

}  // extern "C"
