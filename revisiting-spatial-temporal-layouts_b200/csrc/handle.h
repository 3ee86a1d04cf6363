// Host-side state of a libstlt_b200 handle and the helpers shared by the inference (stlt_api.cu)
// and training (stlt_train.cu) translation units.
#pragma once

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/stlt_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace stlt {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp32 parameter tensors of one encoder layer (device pointers borrowed from the caller). The same
// struct describes the matching gradient buffers (written through const_cast by the backward pass).
struct LayerWeights {
  const float *in_w = nullptr, *in_b = nullptr, *out_w = nullptr, *out_b = nullptr;
  const float *l1_w = nullptr, *l1_b = nullptr, *l2_w = nullptr, *l2_b = nullptr;
  const float *n1_g = nullptr, *n1_b = nullptr, *n2_g = nullptr, *n2_b = nullptr;
  // packed bf16 planes (inside the caller-provided packed buffer)
  const __nv_bfloat16 *in_p = nullptr, *out_p = nullptr, *l1_p = nullptr, *l2_p = nullptr;
  // fused-LayerNorm path (bf16 inference): in-projection folded with the PREVIOUS layer's norm2 (null for
  // the first layer of a stack), linear1 folded with this layer's norm1; s / c epilogue vectors
  const __nv_bfloat16 *in_f = nullptr, *l1_f = nullptr;
  const float *in_s = nullptr, *in_c = nullptr, *l1_s = nullptr, *l1_c = nullptr;
  // in-projection + attention in one kernel (gemm_qkv_attn.cu): the same folded in-projection in HEAD-MAJOR row
  // order (every layer; identity fold for the first layer of a stack) and its s / c vectors in that order
  const __nv_bfloat16* in_h = nullptr;
  const float *in_hs = nullptr, *in_hc = nullptr;
};

struct Weights {
  const float *cat_table = nullptr, *box_w = nullptr, *box_b = nullptr, *score_w = nullptr,
              *score_b = nullptr, *emb_g = nullptr, *emb_b = nullptr;
  const float *pos_table = nullptr, *ft_table = nullptr, *fr_g = nullptr, *fr_b = nullptr;
  const float *fc1_w = nullptr, *fc1_b = nullptr, *head_g = nullptr, *head_b = nullptr,
              *fc2_w = nullptr, *fc2_b = nullptr;
  std::vector<LayerWeights> spatial, temporal;
};

// ---- CACNF fusion path (stlt_cacnf.cu) ----
struct MhaWeights {  // nn.MultiheadAttention: packed in-projection + out_proj
  const float *in_w = nullptr, *in_b = nullptr, *out_w = nullptr, *out_b = nullptr;
  const __nv_bfloat16 *in_p = nullptr, *out_p = nullptr;
};
struct AttnLayerWeights {  // CrossAttentionLayer / SelfAttentionLayer (models.py:328-373)
  MhaWeights attn;
  const float *ln_g = nullptr, *ln_b = nullptr;
};
struct FfnWeights {  // FeedforwardModule (models.py:313-325)
  const float *l1_w = nullptr, *l1_b = nullptr, *l2_w = nullptr, *l2_b = nullptr, *ln_g = nullptr, *ln_b = nullptr;
  const __nv_bfloat16 *l1_p = nullptr, *l2_p = nullptr;
};
struct FusionLayerWeights {  // CrossModalModule (models.py:376-431)
  AttnLayerWeights cross, layout_attn, app_attn, app_ffn;
  FfnWeights layout_ffn;
};
struct HeadWeights {  // ClassificationHead / FusionHead
  const float *fc1_w = nullptr, *fc1_b = nullptr, *ln_g = nullptr, *ln_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr;
};
struct CacnfWeights {
  int app_layers = 0, fusion_layers = 0, app_tokens = 0, feat_channels = 0;
  const float *proj_w = nullptr, *proj_b = nullptr, *cls_token = nullptr, *pos_embed = nullptr;
  const __nv_bfloat16* proj_p = nullptr;
  std::vector<LayerWeights> app;  // TransformerResnet.transformer (ReLU, eps 1e-5)
  std::vector<FusionLayerWeights> fusion;
  HeadWeights app_head, fusion_head;
  bool bound = false;
  int packed_precision = -1;
};

struct Handle {
  StltDims dims{};
  Weights w;
  Weights g;  // fp32 gradient buffers bound by stlt_bind_grads (same names; null = not wanted)
  bool bound = false;
  bool grads_bound = false;
  std::vector<cudaEvent_t> bwd_stage_events;  // stlt_backward_stage_events: recorded as each stage's gradients are final
  int packed_precision = -1;
  const void* packed_ptr = nullptr;
  int num_sms = 0;
  int launches = 0;
  EncodeTiledFn encode = nullptr;
  StltTaps taps{};
  CacnfWeights cacnf;
  // internal capture of the temporal stack's full output (all frames) for the fusion path
  float* cap_tm_x = nullptr;
  __nv_bfloat16* cap_tm_xb = nullptr;
  // optional per-category timing (CUDA events on the launching stream)
  bool fused_ln = true;     // bf16 mode: LayerNorm folded into the GEMM epilogues (stlt_set_fused_ln)
  bool fused_ln_fp32 = true;  // fp32-parity mode: the same on split operands (stlt_set_fused_ln_fp32)
  int fused_attn_max_t = 32;  // longest sequence that takes the fused kernel (STLT_FUSED_ATTENTION_MAX_T, experiments)
  int qkv_attn_debug = 0;   // QkvAttnArgs::debug (STLT_QKV_ATTN_DEBUG environment variable, read at stlt_create)
  // ... with the residual stream of the full phases as two bf16 planes (stlt_set_hilo_residual). Off by default: the
  // register-direct epilogue that reads / writes the planes moves 20 % fewer bytes but in 16-byte pieces per thread and
  // row, and measured 9 % SLOWER per step than the TMA-staged fp32 tiles (DESIGN.md "Measured and rejected")
  bool hilo = false;
  bool compaction = true;   // ... on the pad-skipping row layout of the spatial phase (stlt_set_compaction)
  bool fused_attn = true;   // ... and the attention into the in-projection's epilogue (stlt_set_fused_attention)
  bool bf16_branch = true;  // bf16 mode: out-projection / linear2 outputs travel as bf16 (see run_tail_part)
  bool pruning = true;  // run the row-wise tail of the last layer of each stack on the rows that are read
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  // dyn != null: the launch covered *dyn units of dyn_flops each (pad-skipping layout: counts live on the device and
  // are read back by stlt_get_profile, which synchronises anyway)
  struct Span { int cat; cudaEvent_t a, b; double flops; const int* dyn; double dyn_flops; int role; };
  std::vector<Span> spans;
  StltRoleProfile last_roles{};  // STLT_PROF_GEMM of the most recent stlt_get_profile, by role of the launch
  std::map<std::tuple<const void*, int, long long, long long, int, int>, CUtensorMap> tm_cache;
  char err[512] = {0};
};

extern thread_local char g_err[512];

inline int fail(Handle* h, int code, const char* fmt, ...) {
  char* dst = h ? h->err : g_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

inline cudaEvent_t next_event(Handle* h) {
  if (h->ev_used == h->ev_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    h->ev_pool.push_back(e);
  }
  return h->ev_pool[h->ev_used++];
}

// RAII span: records an event before and after the launches issued in its scope.
struct ProfileScope {
  Handle* h;
  cudaStream_t s;
  cudaEvent_t a = nullptr, b = nullptr;
  int cat;
  double flops;
  const int* dyn = nullptr;
  double dyn_flops = 0.0;
  int role = -1;  // STLT_PROF_ROLE_* of a STLT_PROF_GEMM span
  ProfileScope(Handle* h_, cudaStream_t s_, int cat_, double flops_ = 0.0, const int* dyn_ = nullptr,
               double dyn_flops_ = 0.0, int role_ = -1)
      : h(h_), s(s_), cat(cat_), flops(flops_), dyn(dyn_), dyn_flops(dyn_flops_), role(role_) {
    if (!h->profiling) return;
    a = next_event(h);
    b = next_event(h);
    if (a && b) cudaEventRecord(a, s);
  }
  ~ProfileScope() {
    if (!h->profiling || !a || !b) return;
    cudaEventRecord(b, s);
    h->spans.push_back({cat, a, b, flops, dyn, dyn_flops, role});
  }
};

#define STLT_CUDA(h, expr)                                                                  \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(h, STLT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                  __FILE__, __LINE__);                                                      \
  } while (0)

inline long long pad128(long long v) { return (v + 127) / 128 * 128; }

// Role of a forward projection GEMM, from its shape (stlt_get_profile_by_role).
inline int gemm_role(int n, int k) {
  if (n == kQkv && k == kHidden) return STLT_PROF_ROLE_IN_PROJ;
  if (n == kHidden && k == kHidden) return STLT_PROF_ROLE_OUT_PROJ;
  if (n == kFfn && k == kHidden) return STLT_PROF_ROLE_LINEAR1;
  if (n == kHidden && k == kFfn) return STLT_PROF_ROLE_LINEAR2;
  return STLT_PROF_ROLE_OTHER_GEMM;
}
inline size_t align1k(size_t v) { return (v + 1023) / 1024 * 1024; }

// 2-D row-major tensor map with 128-byte swizzle. dtype: 0 = f32, 1 = bf16.
inline int make_tm(Handle* h, CUtensorMap* tm, const void* ptr, int dtype, long long rows, long long cols,
            int box_cols, int box_rows) {
  auto key = std::make_tuple(ptr, dtype, rows, cols, box_cols, box_rows);
  auto it = h->tm_cache.find(key);
  if (it != h->tm_cache.end()) {
    *tm = it->second;
    return STLT_OK;
  }
  const size_t es = dtype == 0 ? 4 : 2;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(tm, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                         2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(h, STLT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box=%dx%d",
                static_cast<int>(r), rows, cols, box_cols, box_rows);
  if (h->tm_cache.size() > 4096) h->tm_cache.clear();
  h->tm_cache[key] = *tm;
  return STLT_OK;
}

// out = epilogue(A W^T + bias) on tcgen05. a/w point at plane 0; planes are a_plane_rows / n rows apart.
inline int run_gemm(Handle* h, cudaStream_t stream, const void* a, long long m_rows, long long a_plane_rows,
             const void* w, int n, int k, const float* bias, void* out, int terms, int out_kind,
             int gelu, DropCfg drop = DropCfg{0, 0, 1.f}, const int* m_tiles_dyn = nullptr) {
  GemmArgs g{};
  const int planes = terms == 3 ? 2 : 1;
  int rc = make_tm(h, &g.tm_a, a, 1, a_plane_rows * (planes - 1) + m_rows, k, 64, 128);
  if (rc) return rc;
  rc = make_tm(h, &g.tm_b, w, 1, static_cast<long long>(n) * planes, k, 64, 128);  // half tile per CTA
  if (rc) return rc;
  if (out_kind == GEMM_OUT_F32)
    rc = make_tm(h, &g.tm_out, out, 0, m_rows, n, 32, 32);
  else if (out_kind == GEMM_OUT_BF16)
    rc = make_tm(h, &g.tm_out, out, 1, m_rows, n, 64, 32);
  else
    rc = make_tm(h, &g.tm_out, out, 1, a_plane_rows + m_rows, n, 64, 32);
  if (rc) return rc;
  g.bias = bias;
  g.m_rows = static_cast<int>(m_rows);
  g.n = n;
  g.k = k;
  g.terms = terms;
  g.out_kind = out_kind;
  g.gelu = gelu;
  g.a_plane_rows = static_cast<int>(a_plane_rows);
  g.b_plane_rows = n;
  g.out_plane_rows = static_cast<int>(a_plane_rows);
  g.layout = GEMM_NT;
  g.drop = drop;
  g.epilogue = GEMM_EPI_PLAIN;
  g.epi = EpiArgs{};
  g.epi.m_tiles_dyn = m_tiles_dyn;  // pad-skipping layout: live 128-row tiles decided on the device
  g.tm_out2 = g.tm_out;
  ProfileScope prof(h, stream, STLT_PROF_GEMM, m_tiles_dyn ? 0.0 : 2.0 * static_cast<double>(m_rows) * n * k, m_tiles_dyn,
                    2.0 * 128 * n * k, gemm_role(n, k));
  STLT_CUDA(h, launch_gemm_tcgen05(g, stream, h->num_sms));
  h->launches++;
  return STLT_OK;
}

// Projection GEMM with a fused-LayerNorm epilogue (bf16 inference path):
//   GEMM_EPI_NORM_A: out bf16 [m_rows, n] = act(LN(z) W^T + b) from A = bf16(z) and gamma-folded weights
//   GEMM_EPI_RESID : z_out fp32 [m_rows, 768] (+ bf16 copy zb_out) = (prev_norm ? LN(z_prev) : z_prev) + A W^T + bias
// terms == 3 (fp32-parity mode): A, W and the bf16 outputs are hi / lo plane pairs (planes m_rows resp. n rows apart);
// the RESID epilogue then writes the split of the new z as the next GEMM's operand (out_bf16 = hi plane, lo plane behind it).
inline int run_gemm_fused(Handle* h, cudaStream_t stream, int epilogue, const void* a, long long m_rows,
                          const void* w, int n, int k, const float* bias, void* out, void* out_bf16, int gelu,
                          const EpiArgs& epi_in, const int* m_tiles_dyn = nullptr, int terms = 1) {
  GemmArgs g{};
  EpiArgs epi = epi_in;
  const int planes = terms == 3 ? 2 : 1;
  int rc = make_tm(h, &g.tm_a, a, 1, m_rows * planes, k, 64, 128);
  if (rc) return rc;
  rc = make_tm(h, &g.tm_b, w, 1, static_cast<long long>(n) * planes, k, 64, 128);
  if (rc) return rc;
  if (terms == 3 && epilogue == GEMM_EPI_RESID) {
    rc = make_tm(h, &g.tm_out, out, 0, m_rows, n, 32, 32);
    g.tm_out2 = g.tm_out;  // unused: both bf16 planes are written from registers
    g.out_kind = GEMM_OUT_F32_BF16_DIRECT;
    epi.zb_lo_out = static_cast<__nv_bfloat16*>(out_bf16) + static_cast<size_t>(m_rows) * n;
  } else if (terms == 3) {
    rc = make_tm(h, &g.tm_out, out, 1, m_rows * 2, n, 64, 32);
    g.tm_out2 = g.tm_out;
    g.out_kind = GEMM_OUT_BF16_SPLIT;
  } else if (epilogue == GEMM_EPI_RESID && epi.z_lo != nullptr) {
    g.tm_out = g.tm_a;  // unused: the residual stream is read and written as two bf16 planes from registers
    g.tm_out2 = g.tm_a;
    g.out_kind = GEMM_OUT_HILO;
  } else if (epilogue == GEMM_EPI_RESID) {
    rc = make_tm(h, &g.tm_out, out, 0, m_rows, n, 32, 32);
    if (rc) return rc;
    rc = make_tm(h, &g.tm_out2, out_bf16, 1, m_rows, n, 64, 32);
    // short reductions (out-projection) are paced by the epilogue's HBM traffic: TMA-staged bf16 copy;
    // long ones (linear2) by the mainloop: spend the shared memory on a fifth stage instead
    g.out_kind = k <= kHidden ? GEMM_OUT_F32_BF16 : GEMM_OUT_F32_BF16_DIRECT;
  } else {
    rc = make_tm(h, &g.tm_out, out, 1, m_rows, n, 64, 32);
    g.tm_out2 = g.tm_out;
    g.out_kind = GEMM_OUT_BF16;
  }
  if (rc) return rc;
  g.bias = bias;
  g.m_rows = static_cast<int>(m_rows);
  g.n = n;
  g.k = k;
  g.terms = terms;
  g.gelu = gelu;
  g.a_plane_rows = static_cast<int>(m_rows);
  g.b_plane_rows = n;
  g.out_plane_rows = static_cast<int>(m_rows);
  g.layout = GEMM_NT;
  g.drop = DropCfg{0, 0, 1.f};
  g.epilogue = epilogue;
  g.epi = epi;
  g.epi.m_tiles_dyn = m_tiles_dyn;
  ProfileScope prof(h, stream, STLT_PROF_GEMM, m_tiles_dyn ? 0.0 : 2.0 * static_cast<double>(m_rows) * n * k,
                    m_tiles_dyn, 2.0 * 128 * n * k, gemm_role(n, k));
  STLT_CUDA(h, launch_gemm_tcgen05(g, stream, h->num_sms));
  h->launches++;
  return STLT_OK;
}

// ctx bf16 [m_rows, 768] = attention(act_in W_in^T + b_in) in one kernel (gemm_qkv_attn.cu). `a` = bf16 [m_rows, 768]
// (un-normalised residual stream when stats != null: its LayerNorm is folded into w_h / vec_s / vec_c).
// dyn != null: pad-skipping layout (compact.cu) — rows and block counts are read from the device header, num_seqs is
// the static upper bound (all frames) used to size the grid, the last counter of `dyn` is the executed block count.
inline int run_qkv_attention(Handle* h, cudaStream_t stream, const void* a, long long m_rows, long long valid_rows,
                             const __nv_bfloat16* w_h, const float* vec_s, const float* vec_c, const float2* stats,
                             float eps, const long long* mask_src, long long num_seqs, int T, bool causal, void* ctx,
                             const int* dyn = nullptr) {
  const int R = qkv_attention_rows_per_block(T);
  if (R == 0) return fail(h, STLT_ERR_INVALID, "fused attention: sequence length %d outside [1, 32]", T);
  const long long per_block = R / T;
  CUtensorMap tm_a, tm_b, tm_out, tm_out1;
  int rc = make_tm(h, &tm_a, a, 1, m_rows, kHidden, 64, 128);
  if (rc) return rc;
  rc = make_tm(h, &tm_b, w_h, 1, kQkv, kHidden, 64, 96);
  if (rc) return rc;
  rc = make_tm(h, &tm_out, ctx, 1, m_rows, kHidden, 64, R);
  if (rc) return rc;
  tm_out1 = tm_out;
  if (dyn != nullptr) {
    rc = make_tm(h, &tm_out1, ctx, 1, m_rows, kHidden, 64, 128);
    if (rc) return rc;
  }
  QkvAttnArgs p{};
  p.vec_s = vec_s;
  p.vec_c = vec_c;
  p.stats_in = stats;
  p.mask_src = mask_src;
  p.m_rows = m_rows;
  p.valid_rows = valid_rows;
  p.eps = eps;
  p.prev_norm = stats != nullptr ? 1 : 0;
  p.seq_len = T;
  p.rows_per_block = R;
  p.row_blocks = static_cast<int>((num_seqs + per_block - 1) / per_block);
  if (dyn != nullptr) p.row_blocks += static_cast<int>((num_seqs + 127) / 128);  // + blocks of single-token sequences
  p.causal = causal ? 1 : 0;
  p.debug = h->qkv_attn_debug;
  p.dyn = dyn;
  // the executed MMAs cover 128-row blocks of which R rows are kept
  ProfileScope prof(h, stream, STLT_PROF_GEMM, dyn ? 0.0 : 2.0 * static_cast<double>(p.row_blocks) * 128 * kQkv * kHidden,
                    dyn ? dyn + kDynAttnBlocks : nullptr, 2.0 * 128 * kQkv * kHidden, STLT_PROF_ROLE_QKV_ATTENTION);
  STLT_CUDA(h, launch_qkv_attention(tm_a, tm_b, tm_out, tm_out1, p, stream, h->num_sms));
  h->launches++;
  return STLT_OK;
}

// Gradient GEMMs of the training step (bf16 operands, fp32 accumulation):
//   GEMM_NN     out[m_rows, n]  = A[m_rows, k] * B[k, n]       (out fp32 or bf16; rows padded to 128)
//   GEMM_TN_RED out[m_rows, n] += A[k, m_rows]^T * B[k, n]      (out fp32; k = token count, any value)
// act_in != null (GEMM_NN, bf16 output): GEMM_EPI_ACT_BWD, i.e. out = (A B) * gelu'(act_in) * dropout mask and
// colsum += column sums of the first valid_rows rows of out.
inline int run_gemm_grad(Handle* h, cudaStream_t stream, int layout, const void* a, const void* b, void* out,
                  long long m_rows, int n, long long k, int out_kind, const __nv_bfloat16* act_in = nullptr,
                  float* colsum = nullptr, long long valid_rows = 0, DropCfg drop = DropCfg{0, 0, 1.f}) {
  GemmArgs g{};
  int rc;
  if (layout == GEMM_NN) {
    rc = make_tm(h, &g.tm_a, a, 1, m_rows, k, 64, 128);
    if (rc) return rc;
    rc = make_tm(h, &g.tm_b, b, 1, k, n, 64, 64);
  } else if (layout == GEMM_TN_RED) {
    rc = make_tm(h, &g.tm_a, a, 1, k, m_rows, 64, 64);
    if (rc) return rc;
    rc = make_tm(h, &g.tm_b, b, 1, k, n, 64, 64);
  } else {
    return fail(h, STLT_ERR_INVALID, "run_gemm_grad: bad layout %d", layout);
  }
  if (rc) return rc;
  if (out_kind == GEMM_OUT_F32) rc = make_tm(h, &g.tm_out, out, 0, m_rows, n, 32, 32);
  else if (out_kind == GEMM_OUT_BF16) rc = make_tm(h, &g.tm_out, out, 1, m_rows, n, 64, 32);
  else return fail(h, STLT_ERR_INVALID, "run_gemm_grad: bad output kind %d", out_kind);
  if (rc) return rc;
  g.bias = nullptr;
  g.m_rows = static_cast<int>(m_rows);
  g.n = n;
  g.k = static_cast<int>(k);
  g.terms = 1;
  g.out_kind = out_kind;
  g.gelu = 0;
  g.layout = layout;
  g.drop = drop;
  g.epilogue = GEMM_EPI_PLAIN;
  g.epi = EpiArgs{};
  if (act_in != nullptr) {
    if (layout != GEMM_NN || out_kind != GEMM_OUT_BF16) return fail(h, STLT_ERR_INVALID, "run_gemm_grad: ACT_BWD needs NN / bf16");
    g.epilogue = GEMM_EPI_ACT_BWD;
    g.epi.act_in = act_in;
    g.epi.colsum_out = colsum;
    g.epi.valid_rows = static_cast<int>(valid_rows);
  }
  g.tm_out2 = g.tm_out;
  ProfileScope prof(h, stream, STLT_PROF_GEMM, 2.0 * static_cast<double>(m_rows) * n * static_cast<double>(k), nullptr, 0.0,
                    STLT_PROF_ROLE_GRADIENT);
  STLT_CUDA(h, launch_gemm_tcgen05(g, stream, h->num_sms));
  h->launches++;
  return STLT_OK;
}


int bind_table(Handle* h, const StltTensor* tensors, int32_t count, Weights* dst, bool require_all);

}  // namespace stlt
