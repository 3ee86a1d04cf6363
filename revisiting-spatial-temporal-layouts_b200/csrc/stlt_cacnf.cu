// CACNF (CrossAttentionCentralNetFusion, reference src/modelling/models.py:504-549) on precomputed
// per-clip ResNet3D features — SURVEY.md 8(f) rank 2, BASELINE.json configs[4]. Inference forward:
//   layout branch      StltBackbone (the STLT path of this library; all frame tokens are kept)
//   appearance branch  TransformerResnet.forward_features minus the 3D-ResNet trunk (models.py:256-276):
//                      1x1x1 Conv3d projector (a 2048 -> 768 linear per position) + CLS token + position
//                      embedding + 4 post-norm encoder layers (ReLU, LayerNorm eps 1e-5, no masks)
//   fusion             4 x CrossModalModule (models.py:376-431): ONE cross-attention layer applied in both
//                      directions, per-stream self-attention, a GELU feed-forward on the layout stream and
//                      (as in the reference) a second self-attention layer as the appearance "ffn";
//                      LayerNorm eps = config.layer_norm_eps (1e-12)
//   heads              layout / appearance ClassificationHead, FusionHead on the concatenated states,
//                      ensemble = mean of the three (models.py:526-548)
// bf16 GEMM operands on the tcgen05 kernel, fp32 residual streams / LayerNorm / softmax.
#include "handle.h"

namespace {

using namespace stlt;

inline size_t take(size_t& off, size_t bytes) {
  const size_t at = off;
  off += align1k(bytes);
  return at;
}

struct CacnfPlan {
  long long n_l, n_a, m_l, m_a;  // layout / appearance tokens (valid, padded to 128)
  size_t stlt_ws, stlt_bytes;
  size_t feat_tok, proj;                  // bf16 [B*P, C]; f32 [m_p, 768]
  size_t xl, xlb, xa, xab;                // residual streams (f32) + GEMM operands (bf16)
  size_t qkv_l, qkv_a, ctx_l, ctx_a, y_l, y_a, hid;
  size_t pooled, h1, h2, cat;             // head scratch
  size_t total;
};

CacnfPlan plan_cacnf(Handle* h, int B, int L, int S, size_t stlt_bytes, int precision) {
  const CacnfWeights& w = h->cacnf;
  const size_t pl = precision == STLT_PRECISION_FP32 ? 2 : 1;  // bf16 planes (hi / lo in fp32-parity mode)
  CacnfPlan p{};
  const int T = w.app_tokens + 1;
  p.n_l = static_cast<long long>(B) * L;
  p.n_a = static_cast<long long>(B) * T;
  p.m_l = pad128(p.n_l);
  p.m_a = pad128(p.n_a);
  const long long m_p = pad128(static_cast<long long>(B) * w.app_tokens);
  size_t off = 0;
  p.stlt_ws = take(off, stlt_bytes);
  p.stlt_bytes = stlt_bytes;
  p.feat_tok = take(off, static_cast<size_t>(m_p) * w.feat_channels * 2 * pl);
  p.proj = take(off, static_cast<size_t>(m_p) * kHidden * 4);
  p.xl = take(off, static_cast<size_t>(p.m_l) * kHidden * 4);
  p.xlb = take(off, static_cast<size_t>(p.m_l) * kHidden * 2 * pl);
  p.xa = take(off, static_cast<size_t>(p.m_a) * kHidden * 4);
  p.xab = take(off, static_cast<size_t>(p.m_a) * kHidden * 2 * pl);
  p.qkv_l = take(off, static_cast<size_t>(p.m_l) * kQkv * 2 * pl);
  p.qkv_a = take(off, static_cast<size_t>(p.m_a) * kQkv * 2 * pl);
  p.ctx_l = take(off, static_cast<size_t>(p.m_l) * kHidden * 2 * pl);
  p.ctx_a = take(off, static_cast<size_t>(p.m_a) * kHidden * 2 * pl);
  p.y_l = take(off, static_cast<size_t>(p.m_l) * kHidden * 4);
  p.y_a = take(off, static_cast<size_t>(p.m_a) * kHidden * 4);
  p.hid = take(off, static_cast<size_t>(p.m_l > p.m_a ? p.m_l : p.m_a) * kFfn * 2 * pl);
  p.pooled = take(off, static_cast<size_t>(B) * kHidden * 4);
  p.h1 = take(off, static_cast<size_t>(B) * kHidden * 4);
  p.h2 = take(off, static_cast<size_t>(B) * kHidden * 4);
  p.cat = take(off, static_cast<size_t>(B) * 2 * kHidden * 4);
  p.total = off;
  (void)S;
  return p;
}

template <typename T>
T* at(uint8_t* ws, size_t off) {
  return reinterpret_cast<T*>(ws + off);
}

struct Stream {  // one token stream of the fusion stage
  float* x;
  __nv_bfloat16* xb;   // [planes][m][768]
  __nv_bfloat16* qkv;  // [planes][m][2304]
  __nv_bfloat16* ctx;  // [planes][m][768]
  float* y;
  long long m, n;  // padded / valid rows
  int planes;      // 1 = bf16 mode, 2 = fp32-parity mode (hi / lo planes, 3-term GEMMs)
};

int in_proj(Handle* h, cudaStream_t s, const MhaWeights& w, const Stream& st) {
  const bool f = st.planes == 2;
  return run_gemm(h, s, st.xb, st.m, st.m, w.in_p, kQkv, kHidden, w.in_b, st.qkv, f ? 3 : 1,
                  f ? GEMM_OUT_BF16_SPLIT : GEMM_OUT_BF16, 0);
}

// x <- LN(out_proj(ctx) + x)
int out_proj_ln(Handle* h, cudaStream_t s, const MhaWeights& w, const float* g, const float* b, float eps,
                const Stream& st) {
  // bf16 mode: branch outputs travel as bf16 (as in the bf16 STLT path); the residual stream stays fp32
  const bool f = st.planes == 2;
  int rc = run_gemm(h, s, st.ctx, st.m, st.m, w.out_p, kHidden, kHidden, w.out_b, st.y, f ? 3 : 1,
                    f ? GEMM_OUT_F32 : GEMM_OUT_BF16, 0);
  if (rc) return rc;
  ProfileScope prof(h, s, STLT_PROF_ADD_LN);
  ActOut o{st.x, st.xb, st.planes, st.m};
  if (f) STLT_CUDA(h, launch_add_ln(st.x, st.y, g, b, eps, st.n, o, s));
  else STLT_CUDA(h, launch_add_ln_bf16y(st.x, reinterpret_cast<const __nv_bfloat16*>(st.y), g, b, eps, st.n, o, s));
  h->launches++;
  return STLT_OK;
}

int self_attention(Handle* h, cudaStream_t s, const Stream& st, long long num_seqs, int T, const long long* mask_src,
                   bool causal) {
  ProfileScope prof(h, s, STLT_PROF_ATTENTION);
  if (T <= 32)
    STLT_CUDA(h, launch_attention_mma(st.qkv, st.planes, st.m, mask_src, num_seqs, T, causal, st.ctx, st.m, s));
  else
    STLT_CUDA(h, launch_attention_cross(st.qkv, kQkv, 0, st.qkv, kQkv, kHidden, 2 * kHidden, mask_src, num_seqs, T, T,
                                        causal, st.ctx, s, st.planes, st.m, st.m, st.m));
  h->launches++;
  return STLT_OK;
}

// FFN block: x <- LN(l2(act(l1(x))) + x)
int ffn_ln(Handle* h, cudaStream_t s, const __nv_bfloat16* l1_p, const float* l1_b, const __nv_bfloat16* l2_p,
           const float* l2_b, const float* g, const float* b, float eps, int act, const Stream& st,
           __nv_bfloat16* hid) {
  const bool f = st.planes == 2;
  if (f && act == 2) act = 1;  // fp32-parity mode: exact erf GELU
  int rc = run_gemm(h, s, st.xb, st.m, st.m, l1_p, kFfn, kHidden, l1_b, hid, f ? 3 : 1,
                    f ? GEMM_OUT_BF16_SPLIT : GEMM_OUT_BF16, act);
  if (rc) return rc;
  rc = run_gemm(h, s, hid, st.m, st.m, l2_p, kHidden, kFfn, l2_b, st.y, f ? 3 : 1, f ? GEMM_OUT_F32 : GEMM_OUT_BF16, 0);
  if (rc) return rc;
  ProfileScope prof(h, s, STLT_PROF_ADD_LN);
  ActOut o{st.x, st.xb, st.planes, st.m};
  if (f) STLT_CUDA(h, launch_add_ln(st.x, st.y, g, b, eps, st.n, o, s));
  else STLT_CUDA(h, launch_add_ln_bf16y(st.x, reinterpret_cast<const __nv_bfloat16*>(st.y), g, b, eps, st.n, o, s));
  h->launches++;
  return STLT_OK;
}

// ClassificationHead / FusionHead: fc2(LN(gelu(fc1(x))))
int run_head(Handle* h, cudaStream_t s, const HeadWeights& w, const float* x, int in_features, int B, float* h1,
             float* h2, float* logits) {
  ProfileScope prof(h, s, STLT_PROF_OTHER);
  STLT_CUDA(h, launch_gemm_simt(x, w.fc1_w, w.fc1_b, h1, B, kHidden, in_features, true, s));
  ActOut ho{h2, nullptr, 1, 0};
  STLT_CUDA(h, launch_add_ln(h1, nullptr, w.ln_g, w.ln_b, h->dims.layer_norm_eps, B, ho, s));
  STLT_CUDA(h, launch_gemm_simt(h2, w.fc2_w, w.fc2_b, logits, B, h->dims.num_classes, kHidden, false, s));
  h->launches += 3;
  return STLT_OK;
}

}  // namespace

extern "C" {

int stlt_cacnf_bind_weights(void* handle, const StltTensor* tensors, int32_t count, int32_t num_appearance_layers,
                            int32_t num_fusion_layers, int32_t appearance_tokens, int32_t feature_channels) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !tensors) return fail(h, STLT_ERR_INVALID, "null argument");
  if (num_appearance_layers < 0 || num_fusion_layers < 0 || appearance_tokens < 1 || appearance_tokens > 32 ||
      feature_channels < 64 || feature_channels % 64 != 0)
    return fail(h, STLT_ERR_INVALID, "invalid CACNF dimensions (tokens <= 32, channels a multiple of 64)");
  const long long H = kHidden, F = kFfn, C = h->dims.num_classes;

  // 1. layout branch + layout classifier = the STLT path under different prefixes
  std::vector<std::string> renamed(count);
  std::vector<StltTensor> stlt_tensors;
  const std::string lb = "backbone.layout_branch.", lc = "layout_classifier.";
  for (int i = 0; i < count; ++i) {
    if (!tensors[i].name) return fail(h, STLT_ERR_INVALID, "tensor %d has no name", i);
    const std::string name = tensors[i].name;
    if (name.compare(0, lb.size(), lb) == 0) renamed[i] = "backbone." + name.substr(lb.size());
    else if (name.compare(0, lc.size(), lc) == 0) renamed[i] = "prediction_head." + name.substr(lc.size());
    else continue;
    StltTensor t = tensors[i];
    t.name = renamed[i].c_str();
    stlt_tensors.push_back(t);
  }
  Weights w;
  int rc = bind_table(h, stlt_tensors.data(), static_cast<int32_t>(stlt_tensors.size()), &w, true);
  if (rc) return rc;
  h->w = w;
  h->bound = true;
  h->packed_precision = -1;
  h->packed_ptr = nullptr;

  // 2. appearance branch, fusion layers, heads
  CacnfWeights cw;
  cw.app_layers = num_appearance_layers;
  cw.fusion_layers = num_fusion_layers;
  cw.app_tokens = appearance_tokens;
  cw.feat_channels = feature_channels;
  cw.app.resize(num_appearance_layers);
  cw.fusion.resize(num_fusion_layers);
  struct Slot {
    const float** dst;
    std::vector<long long> shape;
  };
  std::map<std::string, Slot> slots;
  auto add = [&](const std::string& name, const float** dst, std::vector<long long> shape) {
    slots[name] = Slot{dst, std::move(shape)};
  };
  const std::string ab = "backbone.appearance_branch.";
  add(ab + "projector.weight", &cw.proj_w, {H, feature_channels, 1, 1});  // [768, C, 1, 1, 1]; ndim 5 checked below
  add(ab + "projector.bias", &cw.proj_b, {H});
  add(ab + "cls_token", &cw.cls_token, {1, 1, H});
  add(ab + "pos_embed", &cw.pos_embed, {appearance_tokens + 1, 1, H});
  for (int i = 0; i < num_appearance_layers; ++i) {
    const std::string p = ab + "transformer.layers." + std::to_string(i) + ".";
    LayerWeights& lw = cw.app[i];
    add(p + "self_attn.in_proj_weight", &lw.in_w, {3 * H, H});
    add(p + "self_attn.in_proj_bias", &lw.in_b, {3 * H});
    add(p + "self_attn.out_proj.weight", &lw.out_w, {H, H});
    add(p + "self_attn.out_proj.bias", &lw.out_b, {H});
    add(p + "linear1.weight", &lw.l1_w, {F, H});
    add(p + "linear1.bias", &lw.l1_b, {F});
    add(p + "linear2.weight", &lw.l2_w, {H, F});
    add(p + "linear2.bias", &lw.l2_b, {H});
    add(p + "norm1.weight", &lw.n1_g, {H});
    add(p + "norm1.bias", &lw.n1_b, {H});
    add(p + "norm2.weight", &lw.n2_g, {H});
    add(p + "norm2.bias", &lw.n2_b, {H});
  }
  auto add_attn = [&](const std::string& p, AttnLayerWeights& a) {
    add(p + "attn.in_proj_weight", &a.attn.in_w, {3 * H, H});
    add(p + "attn.in_proj_bias", &a.attn.in_b, {3 * H});
    add(p + "attn.out_proj.weight", &a.attn.out_w, {H, H});
    add(p + "attn.out_proj.bias", &a.attn.out_b, {H});
    add(p + "ln.weight", &a.ln_g, {H});
    add(p + "ln.bias", &a.ln_b, {H});
  };
  for (int i = 0; i < num_fusion_layers; ++i) {
    const std::string p = "backbone.mm_fusion." + std::to_string(i) + ".";
    FusionLayerWeights& f = cw.fusion[i];
    add_attn(p + "cross_attn.", f.cross);
    add_attn(p + "layout_attn.", f.layout_attn);
    add_attn(p + "appearance_attn.", f.app_attn);
    add_attn(p + "appearance_ffn.", f.app_ffn);
    add(p + "layout_ffn.linear1.weight", &f.layout_ffn.l1_w, {F, H});
    add(p + "layout_ffn.linear1.bias", &f.layout_ffn.l1_b, {F});
    add(p + "layout_ffn.linear2.weight", &f.layout_ffn.l2_w, {H, F});
    add(p + "layout_ffn.linear2.bias", &f.layout_ffn.l2_b, {H});
    add(p + "layout_ffn.ln.weight", &f.layout_ffn.ln_g, {H});
    add(p + "layout_ffn.ln.bias", &f.layout_ffn.ln_b, {H});
  }
  auto add_head = [&](const std::string& p, HeadWeights& hw, long long in_features) {
    add(p + "fc1.weight", &hw.fc1_w, {H, in_features});
    add(p + "fc1.bias", &hw.fc1_b, {H});
    add(p + "layer_norm.weight", &hw.ln_g, {H});
    add(p + "layer_norm.bias", &hw.ln_b, {H});
    add(p + "fc2.weight", &hw.fc2_w, {C, H});
    add(p + "fc2.bias", &hw.fc2_b, {C});
  };
  add_head("appearance_classifier.", cw.app_head, H);
  add_head("fusion_classifier.", cw.fusion_head, 2 * H);

  for (int i = 0; i < count; ++i) {
    const StltTensor& t = tensors[i];
    auto it = slots.find(t.name);
    if (it == slots.end()) continue;  // ResNet3D trunk, unused classifiers, the STLT tensors bound above
    if (t.dtype != STLT_DTYPE_F32) return fail(h, STLT_ERR_INVALID, "%s: expected float32", t.name);
    if (!t.data || (reinterpret_cast<uintptr_t>(t.data) & 15) != 0)
      return fail(h, STLT_ERR_INVALID, "%s: data pointer must be non-null and 16-byte aligned", t.name);
    // compare the element count and the leading dimensions (the Conv3d weight arrives as a 5-D tensor,
    // which StltTensor cannot describe beyond 4 dimensions: callers flatten the trailing 1x1x1)
    const auto& want = it->second.shape;
    long long n_want = 1, n_got = 1;
    for (long long v : want) n_want *= v;
    for (int k = 0; k < t.ndim && k < 4; ++k) n_got *= t.shape[k];
    if (n_want != n_got || t.shape[0] != want[0]) return fail(h, STLT_ERR_INVALID, "%s: unexpected shape", t.name);
    *it->second.dst = static_cast<const float*>(t.data);
  }
  for (auto& kv : slots)
    if (*kv.second.dst == nullptr) return fail(h, STLT_ERR_INVALID, "missing weight: %s", kv.first.c_str());
  cw.bound = true;
  h->cacnf = cw;
  return STLT_OK;
}

int stlt_op_attention_cross(void* handle, void* stream, const void* q_qkv, const void* kv_qkv,
                            const int64_t* mask_src, int64_t num_seqs, int32_t q_len, int32_t kv_len, int32_t causal,
                            void* out_bf16) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !q_qkv || !kv_qkv || !out_bf16) return fail(h, STLT_ERR_INVALID, "null argument");
  STLT_CUDA(h, launch_attention_cross(static_cast<const __nv_bfloat16*>(q_qkv), kQkv, 0,
                                      static_cast<const __nv_bfloat16*>(kv_qkv), kQkv, kHidden, 2 * kHidden,
                                      reinterpret_cast<const long long*>(mask_src), num_seqs, q_len, kv_len, causal != 0,
                                      static_cast<__nv_bfloat16*>(out_bf16), static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

int stlt_cacnf_packed_weights_bytes(void* handle, int32_t precision, size_t* bytes) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !bytes) return fail(h, STLT_ERR_INVALID, "null argument");
  if (precision != STLT_PRECISION_FP32 && precision != STLT_PRECISION_BF16)
    return fail(h, STLT_ERR_INVALID, "unknown precision %d", precision);
  const CacnfWeights& w = h->cacnf;
  if (!w.bound) return fail(h, STLT_ERR_STATE, "stlt_cacnf_bind_weights has not been called");
  const size_t mha = static_cast<size_t>(kHidden) * (kQkv + kHidden);
  const size_t ffn = static_cast<size_t>(kHidden) * 2 * kFfn;
  size_t elems = static_cast<size_t>(kHidden) * w.feat_channels;
  elems += static_cast<size_t>(w.app_layers) * (mha + ffn);
  elems += static_cast<size_t>(w.fusion_layers) * (4 * mha + ffn);
  *bytes = elems * 2 * (precision == STLT_PRECISION_FP32 ? 2 : 1);
  return STLT_OK;
}

int stlt_cacnf_pack_weights(void* handle, void* stream_, int32_t precision, void* packed, size_t bytes) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !packed) return fail(h, STLT_ERR_INVALID, "null argument");
  size_t need = 0;
  int rc = stlt_cacnf_packed_weights_bytes(handle, precision, &need);
  if (rc) return rc;
  if (bytes < need) return fail(h, STLT_ERR_INVALID, "packed buffer too small: %zu < %zu", bytes, need);
  if ((reinterpret_cast<uintptr_t>(packed) & 127) != 0) return fail(h, STLT_ERR_INVALID, "packed buffer must be 128-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CacnfWeights& w = h->cacnf;
  __nv_bfloat16* cur = static_cast<__nv_bfloat16*>(packed);
  const int planes = precision == STLT_PRECISION_FP32 ? 2 : 1;
  auto pack = [&](const float* src, long long n, const __nv_bfloat16** dst) -> cudaError_t {
    *dst = cur;
    cudaError_t e = launch_pack_bf16(src, cur, n, planes, stream);
    cur += n * planes;
    return e;
  };
  const long long H = kHidden, F = kFfn;
  STLT_CUDA(h, pack(w.proj_w, H * w.feat_channels, &w.proj_p));
  for (auto& lw : w.app) {
    STLT_CUDA(h, pack(lw.in_w, 3 * H * H, &lw.in_p));
    STLT_CUDA(h, pack(lw.out_w, H * H, &lw.out_p));
    STLT_CUDA(h, pack(lw.l1_w, F * H, &lw.l1_p));
    STLT_CUDA(h, pack(lw.l2_w, H * F, &lw.l2_p));
  }
  for (auto& f : w.fusion) {
    for (AttnLayerWeights* a : {&f.cross, &f.layout_attn, &f.app_attn, &f.app_ffn}) {
      STLT_CUDA(h, pack(a->attn.in_w, 3 * H * H, &a->attn.in_p));
      STLT_CUDA(h, pack(a->attn.out_w, H * H, &a->attn.out_p));
    }
    STLT_CUDA(h, pack(f.layout_ffn.l1_w, F * H, &f.layout_ffn.l1_p));
    STLT_CUDA(h, pack(f.layout_ffn.l2_w, H * F, &f.layout_ffn.l2_p));
  }
  w.packed_precision = precision;
  return STLT_OK;
}

int stlt_cacnf_workspace_bytes(void* handle, int32_t B, int32_t L, int32_t S, int32_t precision, size_t* bytes) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !bytes) return fail(h, STLT_ERR_INVALID, "null argument");
  if (!h->cacnf.bound) return fail(h, STLT_ERR_STATE, "stlt_cacnf_bind_weights has not been called");
  size_t stlt_bytes = 0;
  int rc = stlt_workspace_bytes(handle, B, L, S, precision, &stlt_bytes);
  if (rc) return rc;
  *bytes = plan_cacnf(h, B, L, S, stlt_bytes, precision).total;
  return STLT_OK;
}

int stlt_cacnf_forward(void* handle, void* stream_, int32_t precision, const int64_t* categories, const float* boxes,
                       const float* scores, const int64_t* frame_types_, const int64_t* lengths_,
                       const float* features, int32_t B, int32_t L, int32_t S, void* workspace,
                       size_t workspace_bytes, float* logits_stlt, float* logits_resnet3d, float* logits_caf,
                       float* logits_ensemble) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  CacnfWeights& w = h->cacnf;
  if (precision != STLT_PRECISION_FP32 && precision != STLT_PRECISION_BF16)
    return fail(h, STLT_ERR_INVALID, "unknown precision %d", precision);
  if (!w.bound || w.packed_precision != precision)
    return fail(h, STLT_ERR_STATE, "CACNF weights are not bound / packed for precision %d", precision);
  if (B < 0) return fail(h, STLT_ERR_INVALID, "negative batch size");
  if (B == 0) return STLT_OK;
  if (!features || !workspace || !logits_stlt || !logits_resnet3d || !logits_caf || !logits_ensemble)
    return fail(h, STLT_ERR_INVALID, "null tensor pointer");
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) return fail(h, STLT_ERR_INVALID, "workspace must be 1024-byte aligned");
  size_t stlt_bytes = 0;
  int rc = stlt_workspace_bytes(handle, B, L, S, precision, &stlt_bytes);
  if (rc) return rc;
  const CacnfPlan p = plan_cacnf(h, B, L, S, stlt_bytes, precision);
  const bool fp32 = precision == STLT_PRECISION_FP32;
  const int planes = fp32 ? 2 : 1;
  if (workspace_bytes < p.total) return fail(h, STLT_ERR_INVALID, "workspace too small: %zu < %zu", workspace_bytes, p.total);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const long long* frame_types = reinterpret_cast<const long long*>(frame_types_);
  const long long* lengths = reinterpret_cast<const long long*>(lengths_);
  const int P = w.app_tokens, T = P + 1;
  const float eps_enc = h->dims.encoder_norm_eps, eps_fus = h->dims.layer_norm_eps;

  Stream sl{at<float>(ws, p.xl), at<__nv_bfloat16>(ws, p.xlb), at<__nv_bfloat16>(ws, p.qkv_l),
            at<__nv_bfloat16>(ws, p.ctx_l), at<float>(ws, p.y_l), p.m_l, p.n_l, planes};
  Stream sa{at<float>(ws, p.xa), at<__nv_bfloat16>(ws, p.xab), at<__nv_bfloat16>(ws, p.qkv_a),
            at<__nv_bfloat16>(ws, p.ctx_a), at<float>(ws, p.y_a), p.m_a, p.n_a, planes};
  __nv_bfloat16* hid = at<__nv_bfloat16>(ws, p.hid);
  float* pooled = at<float>(ws, p.pooled);
  float* h1 = at<float>(ws, p.h1);
  float* h2 = at<float>(ws, p.h2);

  // ---- layout branch: the STLT path; keeps every frame token, emits the layout logits ("stlt") ----
  h->cap_tm_x = sl.x;
  h->cap_tm_xb = sl.xb;
  rc = stlt_forward(handle, stream_, precision, categories, boxes, scores, frame_types_, lengths_, B, L, S,
                    ws + p.stlt_ws, p.stlt_bytes, logits_stlt, nullptr, nullptr);
  h->cap_tm_x = nullptr;
  h->cap_tm_xb = nullptr;
  if (rc) return rc;

  // ---- appearance branch (models.py:256-276) ----
  {
    ProfileScope prof(h, stream, STLT_PROF_OTHER);
    STLT_CUDA(h, launch_features_to_tokens(features, at<__nv_bfloat16>(ws, p.feat_tok), B, w.feat_channels, P, stream,
                                           fp32 ? pad128(static_cast<long long>(B) * P) : 0));
    h->launches++;
  }
  const long long m_p = pad128(static_cast<long long>(B) * P);
  rc = run_gemm(h, stream, at<__nv_bfloat16>(ws, p.feat_tok), m_p, m_p, w.proj_p, kHidden, w.feat_channels, w.proj_b,
                at<float>(ws, p.proj), fp32 ? 3 : 1, GEMM_OUT_F32, 0);
  if (rc) return rc;
  {
    ProfileScope prof(h, stream, STLT_PROF_OTHER);
    ActOut o{sa.x, sa.xb, planes, sa.m};
    STLT_CUDA(h, launch_app_embed(at<float>(ws, p.proj), w.cls_token, w.pos_embed, B, P, o, stream));
  }
  for (const LayerWeights& lw : w.app) {
    MhaWeights mha{lw.in_w, lw.in_b, lw.out_w, lw.out_b, lw.in_p, lw.out_p};
    if ((rc = in_proj(h, stream, mha, sa))) return rc;
    if ((rc = self_attention(h, stream, sa, B, T, nullptr, false))) return rc;
    if ((rc = out_proj_ln(h, stream, mha, lw.n1_g, lw.n1_b, eps_enc, sa))) return rc;
    if ((rc = ffn_ln(h, stream, lw.l1_p, lw.l1_b, lw.l2_p, lw.l2_b, lw.n2_g, lw.n2_b, eps_enc, 3 /*ReLU*/, sa, hid)))
      return rc;
  }
  // appearance_hidden_state = CLS token (models.py:463) -> appearance classifier ("resnet3d")
  {
    int* err_flag = reinterpret_cast<int*>(ws + p.stlt_ws);
    STLT_CUDA(h, launch_gather_rows(sa.x, nullptr, 0, 0, T, nullptr, 0, B, pooled, nullptr, 0, err_flag, stream));
  }
  if ((rc = run_head(h, stream, w.app_head, pooled, kHidden, B, h1, h2, logits_resnet3d))) return rc;

  // ---- multimodal fusion (models.py:464-470, CrossModalModule.forward :395-431) ----
  for (const FusionLayerWeights& f : w.fusion) {
    // one cross-attention layer, both directions; both use the *incoming* states as context
    if ((rc = in_proj(h, stream, f.cross.attn, sl))) return rc;
    if ((rc = in_proj(h, stream, f.cross.attn, sa))) return rc;
    {
      ProfileScope prof(h, stream, STLT_PROF_ATTENTION);
      STLT_CUDA(h, launch_attention_cross(sl.qkv, kQkv, 0, sa.qkv, kQkv, kHidden, 2 * kHidden, nullptr, B, L, T, false,
                                          sl.ctx, stream, planes, sl.m, sa.m, sl.m));
      STLT_CUDA(h, launch_attention_cross(sa.qkv, kQkv, 0, sl.qkv, kQkv, kHidden, 2 * kHidden, frame_types, B, T, L,
                                          false, sa.ctx, stream, planes, sa.m, sl.m, sa.m));
      h->launches += 2;
    }
    if ((rc = out_proj_ln(h, stream, f.cross.attn, f.cross.ln_g, f.cross.ln_b, eps_fus, sl))) return rc;
    if ((rc = out_proj_ln(h, stream, f.cross.attn, f.cross.ln_g, f.cross.ln_b, eps_fus, sa))) return rc;
    // per-stream self-attention
    if ((rc = in_proj(h, stream, f.layout_attn.attn, sl))) return rc;
    if ((rc = self_attention(h, stream, sl, B, L, frame_types, true))) return rc;
    if ((rc = out_proj_ln(h, stream, f.layout_attn.attn, f.layout_attn.ln_g, f.layout_attn.ln_b, eps_fus, sl))) return rc;
    if ((rc = in_proj(h, stream, f.app_attn.attn, sa))) return rc;
    if ((rc = self_attention(h, stream, sa, B, T, nullptr, false))) return rc;
    if ((rc = out_proj_ln(h, stream, f.app_attn.attn, f.app_attn.ln_g, f.app_attn.ln_b, eps_fus, sa))) return rc;
    // layout feed-forward (GELU); the appearance "ffn" is a second self-attention layer (models.py:386)
    const FfnWeights& ff = f.layout_ffn;
    if ((rc = ffn_ln(h, stream, ff.l1_p, ff.l1_b, ff.l2_p, ff.l2_b, ff.ln_g, ff.ln_b, eps_fus, 2 /*GELU*/, sl, hid))) return rc;
    if ((rc = in_proj(h, stream, f.app_ffn.attn, sa))) return rc;
    if ((rc = self_attention(h, stream, sa, B, T, nullptr, false))) return rc;
    if ((rc = out_proj_ln(h, stream, f.app_ffn.attn, f.app_ffn.ln_g, f.app_ffn.ln_b, eps_fus, sa))) return rc;
  }

  // ---- fused state -> fusion classifier ("caf"), ensemble ----
  {
    ProfileScope prof(h, stream, STLT_PROF_OTHER);
    STLT_CUDA(h, launch_gather_concat(sl.x, L, lengths, sa.x, T, B, at<float>(ws, p.cat), stream));
  }
  if ((rc = run_head(h, stream, w.fusion_head, at<float>(ws, p.cat), 2 * kHidden, B, h1, h2, logits_caf))) return rc;
  STLT_CUDA(h, launch_mean3(logits_stlt, logits_resnet3d, logits_caf, static_cast<long long>(B) * h->dims.num_classes,
                            logits_ensemble, stream));
  h->launches += 4;  // app_embed, gather_rows, gather_concat, mean3
  return STLT_OK;
}

}  // extern "C"
