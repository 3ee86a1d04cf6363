// Training step of the STLT path (BASELINE.json configs[3]): forward with
// saved activations, backward, losses, gradient-norm clipping and AdamW — the device side of the
// reference loop at src/train.py:117-135 (criterion: src/utils/train_inference_utils.py:64-76).
//
// Mixed precision: bf16 GEMM operands, fp32 accumulation, fp32 residual stream / LayerNorm / softmax,
// fp32 master weights, gradients and optimizer state. Every projection gradient runs on the same
// tcgen05 kernel as the forward pass (gemm_tcgen05.cu, layouts GEMM_NN / GEMM_TN_RED).
//
// What the forward keeps per encoder layer (everything else is recomputed):
//   xb  bf16 [M, 768]    layer input            (A of the in-projection; B of its weight gradient)
//   qkv bf16 [M, 2304]   in-projection output   (attention backward recomputes the probabilities)
//   att bf16 [M, 768]    attention context      (B of the out-projection weight gradient)
//   z1  f32  [Mt, 768]   x + attn branch        (LayerNorm-1 backward re-normalises it)
//   x1b bf16 [Mt, 768]   LN1 output             (A of linear1)
//   hid bf16 [2, Mt, 3072]  gelu(u) and u       (A of linear2; GELU backward)
//   z2  f32  [Mt, 768]   x1 + FFN branch
// Mt = M except in the last layer of each stack, whose row-wise tail only runs on the rows that are
// read (spatial CLS slot, models.py:79; extract frame, models.py:192) exactly as in inference.
#include <cmath>

#include "handle.h"

namespace {

using namespace stlt;

struct LayerSave {
  size_t xb, qkv, att, att_c, z1, x1b, hid, z2;
};

struct TrainPlan {
  long long n_sp, n_tm, m_sp, m_tm, m_hd;  // valid / padded row counts
  size_t off_err;
  std::vector<LayerSave> sp, tm;
  size_t cls_x;                  // f32 [m_tm, 768] output of the spatial stack (CLS rows)
  size_t pooled, h1, h2;         // f32 [B, 768] head activations
  size_t f[4];                   // f32 [m_sp, 768] transients (residual stream, GEMM outputs, gradients)
  size_t bz, batt, battc, bqkv, bh;  // bf16 transients: dz [m,768], dAtt [m,768], dAtt tail, dQKV, dH
  size_t dh1, dh2;               // f32 [B, 768] head gradients
  size_t total;
};

size_t take(size_t& off, size_t bytes) {
  const size_t at = off;
  off += align1k(bytes);
  return at;
}

TrainPlan plan_train(const StltDims& d, int B, int L, int S) {
  TrainPlan p{};
  p.n_sp = static_cast<long long>(B) * L * S;
  p.n_tm = static_cast<long long>(B) * L;
  p.m_sp = pad128(p.n_sp);
  p.m_tm = pad128(p.n_tm);
  p.m_hd = pad128(B);
  size_t off = 0;
  p.off_err = take(off, 1024);
  auto plan_stack = [&](std::vector<LayerSave>& v, int layers, long long m_full, long long m_last_tail) {
    v.resize(layers);
    for (int i = 0; i < layers; ++i) {
      const bool last = i == layers - 1;
      const size_t mf = static_cast<size_t>(m_full);
      const size_t mt = static_cast<size_t>(last ? m_last_tail : m_full);
      LayerSave& s = v[i];
      s.xb = take(off, mf * kHidden * 2);
      s.qkv = take(off, mf * kQkv * 2);
      s.att = take(off, mf * kHidden * 2);
      s.att_c = last ? take(off, mt * kHidden * 2) : s.att;
      s.z1 = take(off, mt * kHidden * 4);
      s.x1b = take(off, mt * kHidden * 2);
      s.hid = take(off, 2 * mt * kFfn * 2);
      s.z2 = take(off, mt * kHidden * 4);
    }
  };
  plan_stack(p.sp, d.num_spatial_layers, p.m_sp, p.m_tm);
  plan_stack(p.tm, d.num_temporal_layers, p.m_tm, p.m_hd);
  p.cls_x = take(off, static_cast<size_t>(p.m_tm) * kHidden * 4);
  p.pooled = take(off, static_cast<size_t>(p.m_hd) * kHidden * 4);
  p.h1 = take(off, static_cast<size_t>(B) * kHidden * 4);
  p.h2 = take(off, static_cast<size_t>(B) * kHidden * 4);
  p.dh1 = take(off, static_cast<size_t>(B) * kHidden * 4);
  p.dh2 = take(off, static_cast<size_t>(B) * kHidden * 4);
  const size_t m = static_cast<size_t>(p.m_sp > p.m_tm ? p.m_sp : p.m_tm);
  for (int i = 0; i < 4; ++i) p.f[i] = take(off, m * kHidden * 4);
  p.bz = take(off, m * kHidden * 2);
  p.batt = take(off, m * kHidden * 2);
  p.battc = take(off, m * kHidden * 2);
  p.bqkv = take(off, m * kQkv * 2);
  p.bh = take(off, m * kFfn * 2);
  p.total = off;
  return p;
}

template <typename T>
T* at(uint8_t* ws, size_t off) {
  return reinterpret_cast<T*>(ws + off);
}

float* grad_ptr(const float* g) { return const_cast<float*>(g); }

struct Ctx {
  Handle* h;
  cudaStream_t stream;
  uint8_t* ws;
  const TrainPlan* p;
  float dropout_p;
  uint64_t seed;
};

// Dropout sites (element indexing is documented at each kernel): 0 = category/box embedding output,
// 1 = frame embedding output, 16 + 4 * layer + {0: attention probabilities, 1: dropout1 (attention
// branch), 2: FFN inner, 3: dropout2 (FFN branch)} with the spatial layers numbered first.
DropCfg site_cfg(float p, uint64_t seed, int site) {
  DropCfg d{0u, 0u, 1.0f};
  if (p <= 0.0f) return d;
  uint32_t thr = static_cast<uint32_t>(std::lround(static_cast<double>(p) * 65536.0));
  if (thr < 1) thr = 1;
  if (thr > 65535) thr = 65535;
  d.thr16 = thr;
  d.scale = static_cast<float>(65536.0 / (65536.0 - thr));
  d.key = lowbias32(static_cast<uint32_t>(seed) ^
                    lowbias32(static_cast<uint32_t>(seed >> 32) ^ (static_cast<uint32_t>(site) * 0x9e3779b9u + 0x85ebca6bu)));
  return d;
}
DropCfg layer_cfg(const Ctx& c, int layer, int which) { return site_cfg(c.dropout_p, c.seed, 16 + 4 * layer + which); }

// ---- forward of one encoder layer, activations saved into `s` ------------------------------------
// x: fp32 residual stream of the full phase (in/out unless the tail is compacted, in which case the
// layer output is written to x_tail). next_xb: bf16 copy of the layer output (next layer's input).
int fwd_layer(const Ctx& c, int layer, const LayerWeights& lw, const LayerSave& s, long long m_full,
              long long n_full, const long long* mask_src, long long num_seqs, int T, bool causal,
              bool compact, int gather_stride, const long long* lengths, int L, long long m_tail,
              long long n_tail, float* x, float* x_tail, float* y, __nv_bfloat16* next_xb) {
  Handle* h = c.h;
  const float eps = h->dims.encoder_norm_eps;
  __nv_bfloat16* xb = at<__nv_bfloat16>(c.ws, s.xb);
  __nv_bfloat16* qkv = at<__nv_bfloat16>(c.ws, s.qkv);
  __nv_bfloat16* att = at<__nv_bfloat16>(c.ws, s.att);
  __nv_bfloat16* att_c = at<__nv_bfloat16>(c.ws, s.att_c);
  int rc = run_gemm(h, c.stream, xb, m_full, m_full, lw.in_p, kQkv, kHidden, lw.in_b, qkv, 1, GEMM_OUT_BF16, 0);
  if (rc) return rc;
  {
    ProfileScope prof(h, c.stream, STLT_PROF_ATTENTION);
    STLT_CUDA(h, launch_attention_mma(qkv, 1, m_full, mask_src, num_seqs, T, causal, att, m_full, c.stream,
                                      layer_cfg(c, layer, 0)));
  }
  h->launches++;
  float* xt = x;
  if (compact) {
    int* err_flag = at<int>(c.ws, c.p->off_err);
    ProfileScope prof(h, c.stream, STLT_PROF_OTHER);
    STLT_CUDA(h, launch_gather_rows(x, att, 1, m_full, gather_stride, lengths, L, n_tail, x_tail, att_c,
                                    m_tail, err_flag, c.stream));
    h->launches++;
    xt = x_tail;
  }
  (void)n_full;
  // the branch outputs (out-projection, linear2) travel as bf16, as in the bf16 inference path
  const __nv_bfloat16* y_b = reinterpret_cast<const __nv_bfloat16*>(y);
  rc = run_gemm(h, c.stream, att_c, m_tail, m_tail, lw.out_p, kHidden, kHidden, lw.out_b, y, 1, GEMM_OUT_BF16, 0);
  if (rc) return rc;
  {
    ProfileScope prof(h, c.stream, STLT_PROF_ADD_LN);
    ActOut o{xt, at<__nv_bfloat16>(c.ws, s.x1b), 1, m_tail};
    STLT_CUDA(h, launch_add_ln_bf16y(xt, y_b, lw.n1_g, lw.n1_b, eps, n_tail, o, c.stream, at<float>(c.ws, s.z1),
                                     layer_cfg(c, layer, 1)));
  }
  h->launches++;
  rc = run_gemm(h, c.stream, at<__nv_bfloat16>(c.ws, s.x1b), m_tail, m_tail, lw.l1_p, kFfn, kHidden, lw.l1_b,
                at<__nv_bfloat16>(c.ws, s.hid), 1, GEMM_OUT_BF16_DUAL, 2, layer_cfg(c, layer, 2));
  if (rc) return rc;
  rc = run_gemm(h, c.stream, at<__nv_bfloat16>(c.ws, s.hid), m_tail, m_tail, lw.l2_p, kHidden, kFfn, lw.l2_b, y,
                1, GEMM_OUT_BF16, 0);
  if (rc) return rc;
  {
    ProfileScope prof(h, c.stream, STLT_PROF_ADD_LN);
    ActOut o{xt, next_xb, 1, m_tail};
    STLT_CUDA(h, launch_add_ln_bf16y(xt, y_b, lw.n2_g, lw.n2_b, eps, n_tail, o, c.stream, at<float>(c.ws, s.z2),
                                     layer_cfg(c, layer, 3)));
  }
  h->launches++;
  return STLT_OK;
}

// ---- backward of one encoder layer --------------------------------------------------------------
// in : (d_a [+ d_b]) = gradient w.r.t. the layer output on its tail rows
// out: fa = gradient w.r.t. the layer input that flows through the in-projection (full rows, with
//      the residual part already added on the tail rows when the tail is compacted);
//      fb = residual part (tail rows == full rows), or null when it was folded into fa.
// Buffers: fa / fb may alias d_a / d_b (they are consumed first); fc, fd are scratch.
int bwd_layer(const Ctx& c, int layer, const LayerWeights& lw, const LayerWeights& gw, const LayerSave& s,
              long long m_full, long long n_full, const long long* mask_src, long long num_seqs, int T,
              bool causal, bool compact, int scatter_stride, const long long* lengths, int L,
              long long m_tail, long long n_tail, const float* d_a, const float* d_b, float* fa,
              float* fb, float* fc, float* fd, bool* fb_used) {
  Handle* h = c.h;
  const float eps = h->dims.encoder_norm_eps;
  uint8_t* ws = c.ws;
  const TrainPlan& p = *c.p;
  __nv_bfloat16* bz = at<__nv_bfloat16>(ws, p.bz);
  __nv_bfloat16* bh = at<__nv_bfloat16>(ws, p.bh);
  __nv_bfloat16* batt = at<__nv_bfloat16>(ws, p.batt);
  __nv_bfloat16* battc = at<__nv_bfloat16>(ws, p.battc);
  __nv_bfloat16* bqkv = at<__nv_bfloat16>(ws, p.bqkv);
  const __nv_bfloat16* hid = at<__nv_bfloat16>(ws, s.hid);
  const __nv_bfloat16* u = hid + static_cast<size_t>(m_tail) * kFfn;
  int rc;

  // LN2: dz2 -> fc (fp32, residual into LN1) and bz (bf16, GEMM operand); d b2 = colsum(dz2)
  {
    ProfileScope prof(h, c.stream, STLT_PROF_ADD_LN);
    STLT_CUDA(h, launch_ln_bwd(d_a, d_b, at<float>(ws, s.z2), lw.n2_g, eps, n_tail, fc, bz, grad_ptr(gw.n2_g),
                               grad_ptr(gw.n2_b), grad_ptr(gw.l2_b), c.stream, layer_cfg(c, layer, 3)));
  }
  h->launches++;
  // linear2 + GELU: dU = (dz2 W2) * [gelu'(u) * dropout mask] (the factor the forward epilogue stored in the second
  // plane of hid) and d b1 = colsum(dU) in the GEMM epilogue
  // (GEMM_EPI_ACT_BWD); dW2 += dz2^T h
  rc = run_gemm_grad(h, c.stream, GEMM_NN, bz, lw.l2_p, bh, m_tail, kFfn, kHidden, GEMM_OUT_BF16, u,
                     grad_ptr(gw.l1_b), n_tail);
  if (rc) return rc;
  if (gw.l2_w) {
    rc = run_gemm_grad(h, c.stream, GEMM_TN_RED, bz, hid, grad_ptr(gw.l2_w), kHidden, kFfn, n_tail, GEMM_OUT_F32);
    if (rc) return rc;
  }
  // linear1: dX1 = dU W1 -> fd ; dW1 += dU^T x1b
  rc = run_gemm_grad(h, c.stream, GEMM_NN, bh, lw.l1_p, fd, m_tail, kHidden, kFfn, GEMM_OUT_F32);
  if (rc) return rc;
  if (gw.l1_w) {
    rc = run_gemm_grad(h, c.stream, GEMM_TN_RED, bh, at<__nv_bfloat16>(ws, s.x1b), grad_ptr(gw.l1_w), kFfn,
                       kHidden, n_tail, GEMM_OUT_F32);
    if (rc) return rc;
  }
  // LN1: (dz2 + dX1) -> dz1 in fb (fp32 residual into the layer input) and bz; d b_out = colsum(dz1)
  {
    ProfileScope prof(h, c.stream, STLT_PROF_ADD_LN);
    STLT_CUDA(h, launch_ln_bwd(fc, fd, at<float>(ws, s.z1), lw.n1_g, eps, n_tail, fb, bz, grad_ptr(gw.n1_g),
                               grad_ptr(gw.n1_b), grad_ptr(gw.out_b), c.stream, layer_cfg(c, layer, 1)));
  }
  h->launches++;
  // out-projection: dAtt = dz1 Wo ; dWo += dz1^T att
  __nv_bfloat16* datt_tail = compact ? battc : batt;
  rc = run_gemm_grad(h, c.stream, GEMM_NN, bz, lw.out_p, datt_tail, m_tail, kHidden, kHidden, GEMM_OUT_BF16);
  if (rc) return rc;
  if (gw.out_w) {
    rc = run_gemm_grad(h, c.stream, GEMM_TN_RED, bz, at<__nv_bfloat16>(ws, s.att_c), grad_ptr(gw.out_w), kHidden,
                       kHidden, n_tail, GEMM_OUT_F32);
    if (rc) return rc;
  }
  if (compact) {  // only the gathered rows carry a context gradient
    ProfileScope prof(h, c.stream, STLT_PROF_OTHER);
    STLT_CUDA(h, cudaMemsetAsync(batt, 0, static_cast<size_t>(n_full) * kHidden * 2, c.stream));
    STLT_CUDA(h, launch_scatter_rows(nullptr, nullptr, battc, batt, scatter_stride, lengths, L, n_tail, c.stream));
    h->launches++;
  }
  // attention: dQKV; d b_in = colsum(dQKV)
  {
    ProfileScope prof(h, c.stream, STLT_PROF_ATTENTION);
    // also accumulates d b_in = colsum(dQKV)
    STLT_CUDA(h, launch_attention_bwd_mma(at<__nv_bfloat16>(ws, s.qkv), batt, mask_src, num_seqs, T, causal, bqkv,
                                          c.stream, layer_cfg(c, layer, 0), grad_ptr(gw.in_b)));
  }
  h->launches++;
  // in-projection: dX = dQKV Win -> fa ; dWin += dQKV^T xb
  rc = run_gemm_grad(h, c.stream, GEMM_NN, bqkv, lw.in_p, fa, m_full, kHidden, kQkv, GEMM_OUT_F32);
  if (rc) return rc;
  if (gw.in_w) {
    rc = run_gemm_grad(h, c.stream, GEMM_TN_RED, bqkv, at<__nv_bfloat16>(ws, s.xb), grad_ptr(gw.in_w), kQkv,
                       kHidden, n_full, GEMM_OUT_F32);
    if (rc) return rc;
  }
  if (compact) {  // fold the residual gradient of the gathered rows into the full-row gradient
    ProfileScope prof(h, c.stream, STLT_PROF_OTHER);
    STLT_CUDA(h, launch_scatter_rows(fb, fa, nullptr, nullptr, scatter_stride, lengths, L, n_tail, c.stream));
    h->launches++;
    *fb_used = false;
  } else {
    *fb_used = true;
  }
  return STLT_OK;
}

int check_shape(Handle* h, int B, int L, int S) {
  const StltDims& d = h->dims;
  if (B < 1) return fail(h, STLT_ERR_INVALID, "training needs a non-empty batch");
  if (L < 1 || L > d.max_positions || L > 32)
    return fail(h, STLT_ERR_INVALID, "frames=%d outside [1, min(%d, 32)]", L, d.max_positions);
  if (S < 1 || S > 32) return fail(h, STLT_ERR_INVALID, "slots=%d outside [1, 32]", S);
  if (d.num_spatial_layers < 1 || d.num_temporal_layers < 1)
    return fail(h, STLT_ERR_INVALID, "training needs at least one layer per stack");
  return STLT_OK;
}

}  // namespace

extern "C" {

int stlt_train_workspace_bytes(void* handle, int32_t B, int32_t L, int32_t S, size_t* bytes) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !bytes) return fail(h, STLT_ERR_INVALID, "null argument");
  int rc = check_shape(h, B, L, S);
  if (rc) return rc;
  *bytes = plan_train(h->dims, B, L, S).total;
  return STLT_OK;
}

int stlt_bind_grads(void* handle, const StltTensor* tensors, int32_t count) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || (!tensors && count > 0)) return fail(h, STLT_ERR_INVALID, "null argument");
  Weights g;
  int rc = bind_table(h, tensors, count, &g, false);
  if (rc) return rc;
  h->g = g;
  h->grads_bound = true;
  return STLT_OK;
}

int stlt_forward_train(void* handle, void* stream_, const int64_t* categories_, const float* boxes,
                       const float* scores, const int64_t* frame_types_, const int64_t* lengths_,
                       int32_t B, int32_t L, int32_t S, void* workspace, size_t workspace_bytes,
                       float dropout_p, uint64_t seed, float* logits) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (!h->bound) return fail(h, STLT_ERR_STATE, "stlt_bind_weights has not been called");
  if (h->packed_precision != STLT_PRECISION_BF16)
    return fail(h, STLT_ERR_STATE, "training runs in bf16 mixed precision: pack the weights for STLT_PRECISION_BF16");
  int rc = check_shape(h, B, L, S);
  if (rc) return rc;
  if (!(dropout_p >= 0.0f) || dropout_p >= 1.0f) return fail(h, STLT_ERR_INVALID, "dropout_p=%g outside [0, 1)", dropout_p);
  if (!categories_ || !boxes || !frame_types_ || !lengths_ || !workspace || !logits)
    return fail(h, STLT_ERR_INVALID, "null tensor pointer");
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0)
    return fail(h, STLT_ERR_INVALID, "workspace must be 1024-byte aligned");
  const StltDims& d = h->dims;
  const TrainPlan p = plan_train(d, B, L, S);
  if (workspace_bytes < p.total)
    return fail(h, STLT_ERR_INVALID, "workspace too small: %zu < %zu", workspace_bytes, p.total);
  if (p.m_sp > 0x7fffffffLL / 2) return fail(h, STLT_ERR_INVALID, "batch too large for one call");

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long* categories = reinterpret_cast<const long long*>(categories_);
  const long long* frame_types = reinterpret_cast<const long long*>(frame_types_);
  const long long* lengths = reinterpret_cast<const long long*>(lengths_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  Ctx c{h, stream, ws, &p, dropout_p, seed};
  int* err_flag = at<int>(ws, p.off_err);
  h->launches = 0;
  STLT_CUDA(h, cudaMemsetAsync(err_flag, 0, sizeof(int), stream));

  float* x = at<float>(ws, p.f[0]);
  float* y = at<float>(ws, p.f[1]);
  float* x_tail = at<float>(ws, p.f[2]);

  // ---- spatial stack ----
  {
    ProfileScope prof(h, stream, STLT_PROF_OTHER);
    ActOut emb{x, at<__nv_bfloat16>(ws, p.sp[0].xb), 1, p.m_sp};
    STLT_CUDA(h, launch_embed(categories, boxes, scores, h->w.cat_table, d.unique_categories, h->w.box_w,
                              h->w.box_b, h->w.score_w, h->w.score_b, h->w.emb_g, h->w.emb_b,
                              d.layer_norm_eps, p.n_sp, emb, err_flag, stream, y,
                              static_cast<size_t>(p.m_sp) * kHidden * 4,  // y is dead until the first out-projection
                              site_cfg(dropout_p, seed, 0)));
    h->launches += 2;  // embed_stats_kernel + embed_kernel
  }
  const int ns = d.num_spatial_layers, nt = d.num_temporal_layers;
  for (int i = 0; i < ns; ++i) {
    const bool last = i == ns - 1;
    rc = fwd_layer(c, i, h->w.spatial[i], p.sp[i], p.m_sp, p.n_sp, categories, p.n_tm, S, false, last, S, nullptr,
                   0, last ? p.m_tm : p.m_sp, last ? p.n_tm : p.n_sp, x, x_tail, y,
                   last ? nullptr : at<__nv_bfloat16>(ws, p.sp[i + 1].xb));
    if (rc) return rc;
  }
  // the spatial stack's output (CLS rows) is kept in fp32: frame-embedding backward re-normalises it
  STLT_CUDA(h, cudaMemcpyAsync(at<float>(ws, p.cls_x), x_tail, static_cast<size_t>(p.n_tm) * kHidden * 4,
                               cudaMemcpyDeviceToDevice, stream));

  // ---- temporal stack ----
  {
    ProfileScope prof(h, stream, STLT_PROF_OTHER);
    ActOut fr{x, at<__nv_bfloat16>(ws, p.tm[0].xb), 1, p.m_tm};
    STLT_CUDA(h, launch_frame_embed(at<float>(ws, p.cls_x), 1, frame_types, h->w.pos_table, h->w.ft_table,
                                    d.num_frame_types, h->w.fr_g, h->w.fr_b, d.layer_norm_eps, B, L, fr,
                                    err_flag, stream, site_cfg(dropout_p, seed, 1)));
    h->launches++;
  }
  float* pooled = at<float>(ws, p.pooled);
  for (int i = 0; i < nt; ++i) {
    const bool last = i == nt - 1;
    rc = fwd_layer(c, ns + i, h->w.temporal[i], p.tm[i], p.m_tm, p.n_tm, frame_types, B, L, true, last, 0, lengths, L,
                   last ? p.m_hd : p.m_tm, last ? B : p.n_tm, x, pooled, y,
                   last ? nullptr : at<__nv_bfloat16>(ws, p.tm[i + 1].xb));
    if (rc) return rc;
  }

  // ---- head (models.py:155-163): fc1 -> GELU -> LayerNorm -> fc2 ----
  {
    ProfileScope prof(h, stream, STLT_PROF_OTHER);
    float* h1 = at<float>(ws, p.h1);
    float* h2 = at<float>(ws, p.h2);
    STLT_CUDA(h, launch_gemm_simt(pooled, h->w.fc1_w, h->w.fc1_b, h1, B, kHidden, kHidden, false, stream));
    STLT_CUDA(h, launch_gelu_ln(h1, h->w.head_g, h->w.head_b, d.layer_norm_eps, B, h2, stream));
    STLT_CUDA(h, launch_gemm_simt(h2, h->w.fc2_w, h->w.fc2_b, logits, B, d.num_classes, kHidden, false, stream));
    h->launches += 3;
  }
  return STLT_OK;
}

int stlt_backward_stage_events(void* handle, int32_t enable, int32_t* num_stages_out) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  const int n = h->dims.num_temporal_layers + h->dims.num_spatial_layers + 3;
  if (num_stages_out) *num_stages_out = n;
  if (enable && h->bwd_stage_events.empty()) {
    for (int i = 0; i < n; ++i) {
      cudaEvent_t e;
      STLT_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->bwd_stage_events.push_back(e);
    }
  } else if (!enable) {
    for (cudaEvent_t e : h->bwd_stage_events) cudaEventDestroy(e);
    h->bwd_stage_events.clear();
  }
  return STLT_OK;
}

int stlt_stream_wait_backward_stage(void* handle, void* stream, int32_t stage) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (h->bwd_stage_events.empty()) return fail(h, STLT_ERR_STATE, "stlt_backward_stage_events has not been enabled");
  if (stage < 0 || stage >= static_cast<int>(h->bwd_stage_events.size()))
    return fail(h, STLT_ERR_INVALID, "stage %d out of range [0, %zu)", stage, h->bwd_stage_events.size());
  STLT_CUDA(h, cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->bwd_stage_events[stage], 0));
  return STLT_OK;
}

int stlt_backward(void* handle, void* stream_, const int64_t* categories_, const float* boxes,
                  const float* scores, const int64_t* frame_types_, const int64_t* lengths_, int32_t B,
                  int32_t L, int32_t S, void* workspace, size_t workspace_bytes, float dropout_p,
                  uint64_t seed, const float* d_logits, int32_t phases) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (!h->bound || !h->grads_bound)
    return fail(h, STLT_ERR_STATE, "stlt_bind_weights / stlt_bind_grads have not been called");
  if (h->packed_precision != STLT_PRECISION_BF16)
    return fail(h, STLT_ERR_STATE, "weights are not packed for STLT_PRECISION_BF16");
  int rc = check_shape(h, B, L, S);
  if (rc) return rc;
  if (!categories_ || !boxes || !frame_types_ || !lengths_ || !workspace)
    return fail(h, STLT_ERR_INVALID, "null tensor pointer");
  if ((phases & STLT_BWD_TEMPORAL) && !d_logits) return fail(h, STLT_ERR_INVALID, "null d_logits");
  const StltDims& d = h->dims;
  const TrainPlan p = plan_train(d, B, L, S);
  if (workspace_bytes < p.total)
    return fail(h, STLT_ERR_INVALID, "workspace too small: %zu < %zu", workspace_bytes, p.total);

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long* categories = reinterpret_cast<const long long*>(categories_);
  const long long* frame_types = reinterpret_cast<const long long*>(frame_types_);
  const long long* lengths = reinterpret_cast<const long long*>(lengths_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  Ctx c{h, stream, ws, &p, dropout_p, seed};
  const Weights& w = h->w;
  const Weights& g = h->g;
  const int ns = d.num_spatial_layers, nt = d.num_temporal_layers;
  const int C = d.num_classes;
  h->launches = 0;
  float* f0 = at<float>(ws, p.f[0]);
  float* f1 = at<float>(ws, p.f[1]);
  float* f2 = at<float>(ws, p.f[2]);
  float* f3 = at<float>(ws, p.f[3]);
  // gradient w.r.t. the spatial stack's output (CLS rows); survives between the two phases
  float* d_cls = at<float>(ws, p.cls_x);  // cls_x itself is consumed by frame_embed_bwd before being overwritten
  // the parameter gradients of `stage` are final: the communication stream may start their all-reduce (stlt_b200.h)
  auto stage_done = [&](int stage) -> cudaError_t {
    return h->bwd_stage_events.empty() ? cudaSuccess : cudaEventRecord(h->bwd_stage_events[stage], stream);
  };

  if (phases & STLT_BWD_TEMPORAL) {
    // ---- head ----
    float* dh1 = at<float>(ws, p.dh1);
    float* dh2 = at<float>(ws, p.dh2);
    const float* h1 = at<float>(ws, p.h1);
    const float* h2 = at<float>(ws, p.h2);
    const float* pooled = at<float>(ws, p.pooled);
    {
      ProfileScope prof(h, stream, STLT_PROF_OTHER);
      if (g.fc2_w)  // d fc2.weight [C, 768] += d_logits^T h2
        STLT_CUDA(h, launch_gemm_strided(d_logits, 1, C, h2, kHidden, 1, grad_ptr(g.fc2_w), C, kHidden, B, true, stream));
      if (g.fc2_b) STLT_CUDA(h, launch_colsum_f32(d_logits, B, C, grad_ptr(g.fc2_b), stream));
      STLT_CUDA(h, launch_gemm_strided(d_logits, C, 1, w.fc2_w, kHidden, 1, dh2, B, kHidden, C, false, stream));
      STLT_CUDA(h, launch_gelu_ln_bwd(dh2, h1, w.head_g, d.layer_norm_eps, B, dh1, grad_ptr(g.head_g),
                                      grad_ptr(g.head_b), stream));
      if (g.fc1_w)
        STLT_CUDA(h, launch_gemm_strided(dh1, 1, kHidden, pooled, kHidden, 1, grad_ptr(g.fc1_w), kHidden, kHidden,
                                         B, true, stream));
      if (g.fc1_b) STLT_CUDA(h, launch_colsum_f32(dh1, B, kHidden, grad_ptr(g.fc1_b), stream));
      // d pooled -> f0 (tail rows of the last temporal layer)
      STLT_CUDA(h, launch_gemm_strided(dh1, kHidden, 1, w.fc1_w, kHidden, 1, f0, B, kHidden, kHidden, false, stream));
      h->launches += 7;
    }
    STLT_CUDA(h, stage_done(0));
    // ---- temporal stack, last layer first ----
    const float* d_a = f0;
    const float* d_b = nullptr;
    for (int i = nt - 1; i >= 0; --i) {
      const bool last = i == nt - 1;
      bool fb_used = false;
      // the incoming gradient lives in f0 (+ f1); they are consumed by the first kernel, so the
      // layer's outputs go to the same pair and f2 / f3 are scratch
      rc = bwd_layer(c, ns + i, w.temporal[i], g.temporal[i], p.tm[i], p.m_tm, p.n_tm, frame_types, B, L, true, last, 0,
                     lengths, L, last ? p.m_hd : p.m_tm, last ? B : p.n_tm, d_a, d_b, f0, f1, f2, f3, &fb_used);
      if (rc) return rc;
      d_a = f0;
      d_b = fb_used ? f1 : nullptr;
      STLT_CUDA(h, stage_done(nt - i));
    }
    // ---- frame embedding ----
    {
      ProfileScope prof(h, stream, STLT_PROF_OTHER);
      STLT_CUDA(h, launch_frame_embed_bwd(d_a, d_b, at<float>(ws, p.cls_x), frame_types, w.pos_table, w.ft_table,
                                          d.num_frame_types, w.fr_g, d.layer_norm_eps, B, L, f2, grad_ptr(g.pos_table),
                                          grad_ptr(g.ft_table), grad_ptr(g.fr_g), grad_ptr(g.fr_b), stream,
                                          site_cfg(dropout_p, seed, 1)));
      STLT_CUDA(h, cudaMemcpyAsync(d_cls, f2, static_cast<size_t>(p.n_tm) * kHidden * 4, cudaMemcpyDeviceToDevice,
                                   stream));
      h->launches++;
    }
    STLT_CUDA(h, stage_done(nt + 1));
  }

  if (phases & STLT_BWD_SPATIAL) {
    STLT_CUDA(h, cudaMemcpyAsync(f0, d_cls, static_cast<size_t>(p.n_tm) * kHidden * 4, cudaMemcpyDeviceToDevice,
                                 stream));
    const float* d_a = f0;
    const float* d_b = nullptr;
    for (int i = ns - 1; i >= 0; --i) {
      const bool last = i == ns - 1;
      bool fb_used = false;
      rc = bwd_layer(c, i, w.spatial[i], g.spatial[i], p.sp[i], p.m_sp, p.n_sp, categories, p.n_tm, S, false, last, S,
                     nullptr, 0, last ? p.m_tm : p.m_sp, last ? p.n_tm : p.n_sp, d_a, d_b, f0, f1, f2, f3, &fb_used);
      if (rc) return rc;
      d_a = f0;
      d_b = fb_used ? f1 : nullptr;
      STLT_CUDA(h, stage_done(nt + 2 + (ns - 1 - i)));
    }
    {
      ProfileScope prof(h, stream, STLT_PROF_OTHER);
      const bool has_scores = scores != nullptr;
      STLT_CUDA(h, launch_embed_bwd(d_a, d_b, categories, boxes, scores, w.cat_table, d.unique_categories, w.box_w,
                                    w.box_b, w.score_w, w.score_b, w.emb_g, d.layer_norm_eps, p.n_sp, f2,
                                    grad_ptr(g.cat_table), grad_ptr(g.box_w), grad_ptr(g.box_b),
                                    has_scores ? grad_ptr(g.score_w) : nullptr,
                                    has_scores ? grad_ptr(g.score_b) : nullptr, grad_ptr(g.emb_g),
                                    grad_ptr(g.emb_b), stream, site_cfg(dropout_p, seed, 0)));
      h->launches += 2;
    }
    STLT_CUDA(h, stage_done(nt + ns + 2));
  }
  return STLT_OK;
}

int stlt_op_dropout_mask(void* handle, float dropout_p, uint64_t seed, int32_t site, int64_t first, int64_t n,
                         float* multipliers_host) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !multipliers_host || n < 0 || first < 0) return fail(h, STLT_ERR_INVALID, "invalid argument");
  const DropCfg d = site_cfg(dropout_p, seed, site);
  for (int64_t i = 0; i < n; ++i) {
    const unsigned long long e = static_cast<unsigned long long>(first + i);
    multipliers_host[i] = d.thr16 == 0 ? 1.0f : drop_mul(drop_bits(d.key, e >> 1), static_cast<int>(e & 1), d);
  }
  return STLT_OK;
}

int stlt_op_attention_bwd(void* handle, void* stream, const void* qkv, const void* d_ctx,
                          const int64_t* mask_src, int64_t num_seqs, int32_t seq_len, int32_t causal,
                          void* d_qkv, int32_t impl, void* d_bias_or_null) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !qkv || !d_ctx || !mask_src || !d_qkv) return fail(h, STLT_ERR_INVALID, "null argument");
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* d = static_cast<const __nv_bfloat16*>(d_ctx);
  const long long* m = reinterpret_cast<const long long*>(mask_src);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(d_qkv);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (impl == 0) STLT_CUDA(h, launch_attention_bwd(q, d, m, num_seqs, seq_len, causal != 0, o, s));
  else STLT_CUDA(h, launch_attention_bwd_mma(q, d, m, num_seqs, seq_len, causal != 0, o, s, DropCfg{0, 0, 1.f},
                                             static_cast<float*>(d_bias_or_null)));
  return STLT_OK;
}

int stlt_loss(void* handle, void* stream, int32_t kind, const float* logits, const void* labels,
              int32_t rows, int32_t classes, float grad_scale, float* loss_out, float* d_logits_out) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (rows < 1 || classes < 1) return fail(h, STLT_ERR_INVALID, "invalid shape");
  if (!logits || !labels) return fail(h, STLT_ERR_INVALID, "null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (loss_out) STLT_CUDA(h, cudaMemsetAsync(loss_out, 0, sizeof(float), s));
  if (kind == STLT_LOSS_CROSS_ENTROPY) {
    STLT_CUDA(h, launch_cross_entropy(logits, static_cast<const long long*>(labels), rows, classes, grad_scale,
                                      loss_out, d_logits_out, s));
  } else if (kind == STLT_LOSS_BCE_LOGITS) {
    STLT_CUDA(h, launch_bce_logits(logits, static_cast<const float*>(labels),
                                   static_cast<long long>(rows) * classes, grad_scale, loss_out, d_logits_out, s));
  } else {
    return fail(h, STLT_ERR_INVALID, "unknown loss kind %d", kind);
  }
  return STLT_OK;
}

int stlt_grad_sumsq(void* handle, void* stream, const float* grads, int64_t n, float* sumsq_out, float* scratch,
                    int32_t scratch_floats) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (n < 0 || !grads || !sumsq_out || !scratch || scratch_floats < 1) return fail(h, STLT_ERR_INVALID, "invalid argument");
  STLT_CUDA(h, launch_sumsq(grads, n, sumsq_out, scratch, scratch_floats, static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

int stlt_adamw_step(void* handle, void* stream, float* params, const float* grads, float* exp_avg,
                    float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int32_t step, const float* sumsq_or_null, float max_norm) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (n < 0 || !params || !grads || !exp_avg || !exp_avg_sq || step < 1)
    return fail(h, STLT_ERR_INVALID, "invalid argument");
  const double bc1 = 1.0 - std::pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - std::pow(static_cast<double>(beta2), step);
  STLT_CUDA(h, launch_adamw(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                            static_cast<float>(bc1), static_cast<float>(std::sqrt(bc2)), sumsq_or_null, max_norm,
                            static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

}  // extern "C"
