// C ABI of libstlt_b200.so (declared in include/stlt_b200.h): handle management, weight binding
// and packing, workspace planning and the orchestration of the forward pass
// (reference: Stlt.forward, src/modelling/models.py:185-195 and everything it calls).
#include "handle.h"

namespace stlt {
thread_local char g_err[512] = "invalid handle";
}

namespace stlt {

// Resolves reference state_dict names to the slots of a Weights table (parameters or gradients).
int bind_table(Handle* h, const StltTensor* tensors, int32_t count, Weights* dst, bool require_all) {
  const StltDims& d = h->dims;
  Weights w;
  w.spatial.resize(d.num_spatial_layers);
  w.temporal.resize(d.num_temporal_layers);
  const std::string bfe = "backbone.frames_embeddings.";
  const std::string cbe = bfe + "layout_embedding.category_box_embeddings.";
  const std::string sp = bfe + "layout_embedding.transformer.layers.";
  const std::string tp = "backbone.transformer.layers.";
  const std::string hd = "prediction_head.";

  struct Slot {
    const float** dst;
    std::vector<long long> shape;
  };
  std::map<std::string, Slot> slots;
  auto add = [&](const std::string& name, const float** dst, std::vector<long long> shape) {
    slots[name] = Slot{dst, std::move(shape)};
  };
  const long long H = kHidden, F = kFfn;
  add(cbe + "category_embeddings.weight", &w.cat_table, {d.unique_categories, H});
  add(cbe + "box_embedding.weight", &w.box_w, {H, 4});
  add(cbe + "box_embedding.bias", &w.box_b, {H});
  add(cbe + "score_embeddings.weight", &w.score_w, {H, 1});
  add(cbe + "score_embeddings.bias", &w.score_b, {H});
  add(cbe + "layer_norm.weight", &w.emb_g, {H});
  add(cbe + "layer_norm.bias", &w.emb_b, {H});
  add(bfe + "position_embeddings.weight", &w.pos_table, {d.max_positions, H});
  add(bfe + "frame_type_embedding.weight", &w.ft_table, {d.num_frame_types, H});
  add(bfe + "layer_norm.weight", &w.fr_g, {H});
  add(bfe + "layer_norm.bias", &w.fr_b, {H});
  add(hd + "fc1.weight", &w.fc1_w, {H, H});
  add(hd + "fc1.bias", &w.fc1_b, {H});
  add(hd + "layer_norm.weight", &w.head_g, {H});
  add(hd + "layer_norm.bias", &w.head_b, {H});
  add(hd + "fc2.weight", &w.fc2_w, {d.num_classes, H});
  add(hd + "fc2.bias", &w.fc2_b, {d.num_classes});
  auto add_layer = [&](const std::string& p, LayerWeights& lw) {
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
  };
  for (int i = 0; i < d.num_spatial_layers; ++i) add_layer(sp + std::to_string(i) + ".", w.spatial[i]);
  for (int i = 0; i < d.num_temporal_layers; ++i) add_layer(tp + std::to_string(i) + ".", w.temporal[i]);

  for (int i = 0; i < count; ++i) {
    const StltTensor& t = tensors[i];
    if (!t.name) return fail(h, STLT_ERR_INVALID, "tensor %d has no name", i);
    auto it = slots.find(t.name);
    if (it == slots.end()) continue;  // orphan encoder_layer.*, position_ids, ...
    if (t.dtype != STLT_DTYPE_F32) return fail(h, STLT_ERR_INVALID, "%s: expected float32", t.name);
    if (!t.data) return fail(h, STLT_ERR_INVALID, "%s: null data pointer", t.name);
    if ((reinterpret_cast<uintptr_t>(t.data) & 15) != 0)
      return fail(h, STLT_ERR_INVALID, "%s: data pointer must be 16-byte aligned", t.name);
    const auto& want = it->second.shape;
    bool ok = t.ndim == static_cast<int>(want.size());
    for (size_t k = 0; ok && k < want.size(); ++k) ok = t.shape[k] == want[k];
    if (!ok) return fail(h, STLT_ERR_INVALID, "%s: unexpected shape", t.name);
    *it->second.dst = static_cast<const float*>(t.data);
  }
  if (require_all)
    for (auto& kv : slots)
      if (*kv.second.dst == nullptr) return fail(h, STLT_ERR_INVALID, "missing weight: %s", kv.first.c_str());
  *dst = w;
  return STLT_OK;
}


}  // namespace stlt

namespace {

using namespace stlt;

// Activation buffers of one phase of the forward (rows padded to the 128-row GEMM tile).
struct BufSet {
  size_t x, y, xb, att, qkv, hid;  // byte offsets into the workspace
  long long m_pad;
};

struct WorkspacePlan {
  size_t off_plan = 0;  // pad-skipping layout: header int[8] | frame_row int[B*L] | mask i64[rows] | plan scratch
  size_t off_frame_row = 0, off_mask = 0, off_plan_scratch = 0;
  BufSet sp;  // spatial phase : B*L*S object tokens
  BufSet tm;  // temporal phase: B*L frame tokens (also the CLS-only tail of the last spatial layer)
  BufSet hd;  // B extract-frame tokens (tail of the last temporal layer)
  size_t off_err, off_head, total;
  // fused-LayerNorm path: per-row (sum, sum of squares) accumulators of every LayerNorm site
  size_t off_stats, stats_bytes;
};

size_t plan_bufset(BufSet* b, size_t off, long long rows, int precision, bool with_qkv) {
  const size_t planes = precision == STLT_PRECISION_FP32 ? 2 : 1;
  const size_t m = static_cast<size_t>(pad128(rows));
  b->m_pad = static_cast<long long>(m);
  b->x = off;    off += align1k(m * kHidden * 4);
  b->y = off;    off += align1k(m * kHidden * 4);
  b->xb = off;   off += align1k(m * kHidden * 2 * 2);  // bf16 mode: second plane = lo plane of the hi/lo residual stream
  b->att = off;  off += align1k(m * kHidden * 2 * planes);
  b->qkv = off;
  if (with_qkv) off += align1k(m * kQkv * (precision == STLT_PRECISION_FP32 ? 4 : 2));
  b->hid = off;  off += align1k(m * kFfn * 2 * planes);
  return off;
}

WorkspacePlan plan_workspace(int B, int L, int S, int precision, int ns = 0, int nt = 0) {
  WorkspacePlan p{};
  size_t off = 0;
  p.off_err = off;
  off += 1024;
  // the spatial phase may run on the pad-skipping layout, whose static row bound is a little larger
  long long sp_rows = static_cast<long long>(B) * L * S;
  const long long compact_rows = compact_rows_bound(static_cast<long long>(B) * L, S);  // 0 when S > 32
  if (compact_rows > sp_rows) sp_rows = compact_rows;
  off = plan_bufset(&p.sp, off, sp_rows, precision, true);
  off = plan_bufset(&p.tm, off, static_cast<long long>(B) * L, precision, true);
  off = plan_bufset(&p.hd, off, B, precision, false);
  p.off_head = off;
  off += align1k(static_cast<size_t>(B) * kHidden * 4 * 3);
  // fused-LayerNorm statistics (bf16 mode): 2 sites per spatial layer over the object tokens, 2 per temporal
  // layer + 6 compaction sites over the frame tokens
  p.off_stats = off;
  p.stats_bytes = 0;
  p.stats_bytes = (static_cast<size_t>(2 * ns) * p.sp.m_pad + static_cast<size_t>(2 * nt + 6) * p.tm.m_pad) *
                  kStatSlots * sizeof(float2);
  off += align1k(p.stats_bytes);
  if (compact_rows > 0) {
    const long long frames = static_cast<long long>(B) * L;
    p.off_plan = off;
    off += 1024;
    p.off_frame_row = off;
    off += align1k(static_cast<size_t>(frames) * sizeof(int));
    p.off_mask = off;
    off += align1k(static_cast<size_t>(p.sp.m_pad) * sizeof(long long));
    p.off_plan_scratch = off;
    off += align1k(compact_plan_scratch_bytes(frames));
  }
  p.total = off;
  return p;
}

struct Phase {
  float* x;              // fp32 residual stream [m, 768]
  float* y;              // fp32 GEMM output [m, 768]
  __nv_bfloat16* xb;     // bf16 plane(s) of x
  __nv_bfloat16* xlo = nullptr;  // hi/lo residual stream (GEMM_OUT_HILO): lo plane, xb is the hi plane; null = fp32 x
  __nv_bfloat16* att;    // bf16 plane(s) of the attention context
  void* qkv;             // bf16 [P, m, 2304]
  __nv_bfloat16* hid;    // bf16 plane(s) of the FFN hidden [m, 3072]
  long long m_pad;       // rows per plane
  long long m_valid;     // real tokens
};

Phase make_phase(uint8_t* ws, const BufSet& b, long long m_valid) {
  Phase ph{};
  ph.x = reinterpret_cast<float*>(ws + b.x);
  ph.y = reinterpret_cast<float*>(ws + b.y);
  ph.xb = reinterpret_cast<__nv_bfloat16*>(ws + b.xb);
  ph.att = reinterpret_cast<__nv_bfloat16*>(ws + b.att);
  ph.qkv = ws + b.qkv;
  ph.hid = reinterpret_cast<__nv_bfloat16*>(ws + b.hid);
  ph.m_pad = b.m_pad;
  ph.m_valid = m_valid;
  return ph;
}

// The FFN hidden buffer is dead until the first linear1: the embedding kernel's statistics tables live there.
size_t embed_scratch(const Phase& ph) { return static_cast<size_t>(ph.m_pad) * kFfn * 2; }

// First half of a post-norm nn.TransformerEncoderLayer (eval mode): in-projection + attention.
// Needs every token of the sequences (keys / values), so it always runs on the full phase.
// dyn != null: the phase is on the pad-skipping layout (compact.cu): row / sequence counts come from the device header,
// `num_seqs` is the static bound; the one-token sequences of the second row region get their own attention launch.
// The attention of the separate-kernel paths on the packed q | k | v of ph.qkv (context -> ph.att).
int run_attention_only(Handle* h, cudaStream_t stream, bool fp32, const Phase& ph, const long long* mask_src,
                       long long num_seqs, int T, bool causal, const int* dyn) {
  // fp32 mode: q/k/v and the context travel as bf16 hi/lo planes; products are 3-term splits
  ActOut att{nullptr, ph.att, fp32 ? 2 : 1, ph.m_pad};
  {
    ProfileScope prof(h, stream, STLT_PROF_ATTENTION);
    if (dyn != nullptr) {
      const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(ph.qkv);
      STLT_CUDA(h, launch_attention_mma(q, att.planes, ph.m_pad, mask_src, num_seqs, T, causal, ph.att, ph.m_pad, stream,
                                        DropCfg{0, 0, 1.f}, dyn, 0));
      STLT_CUDA(h, launch_attention_mma(q, att.planes, ph.m_pad, mask_src, num_seqs, 1, false, ph.att, ph.m_pad, stream,
                                        DropCfg{0, 0, 1.f}, dyn, 1));
      h->launches++;
    } else {
      STLT_CUDA(h, launch_attention(ph.qkv, true, mask_src, num_seqs, T, causal, att, stream));
    }
  }
  h->launches++;
  return STLT_OK;
}

int run_attention_part(Handle* h, cudaStream_t stream, int precision, const LayerWeights& lw,
                       const Phase& ph, const long long* mask_src, long long num_seqs, int T,
                       bool causal, const int* dyn = nullptr) {
  const bool fp32 = precision == STLT_PRECISION_FP32;
  int rc = run_gemm(h, stream, ph.xb, ph.m_pad, ph.m_pad, lw.in_p, kQkv, kHidden, lw.in_b, ph.qkv,
                    fp32 ? 3 : 1, fp32 ? GEMM_OUT_BF16_SPLIT : GEMM_OUT_BF16, 0, DropCfg{0, 0, 1.f},
                    dyn != nullptr ? dyn + kDynTiles : nullptr);
  if (rc) return rc;
  return run_attention_only(h, stream, fp32, ph, mask_src, num_seqs, T, causal, dyn);
}

// Second half: out-projection -> +residual -> LN -> FFN -> +residual -> LN. Row-wise, so it may
// run on a compacted subset of the rows (the pruned last layer of each stack).
int run_tail_part(Handle* h, cudaStream_t stream, int precision, const LayerWeights& lw,
                  const Phase& ph, const int* dyn = nullptr) {
  const int* tiles_dyn = dyn != nullptr ? dyn + kDynTiles : nullptr;
  const int* rows_dyn = dyn != nullptr ? dyn + kDynRows : nullptr;
  const DropCfg nodrop{0, 0, 1.f};
  const bool fp32 = precision == STLT_PRECISION_FP32;
  const int terms = fp32 ? 3 : 1;
  const int planes = fp32 ? 2 : 1;
  const float eps = h->dims.encoder_norm_eps;
  // bf16 mode: the branch outputs (out-projection, linear2) are stored as bf16 and widened again in the
  // residual add; the residual stream itself stays fp32
  const bool y16 = !fp32 && h->bf16_branch;
  __nv_bfloat16* y_b = reinterpret_cast<__nv_bfloat16*>(ph.y);
  int rc = run_gemm(h, stream, ph.att, ph.m_pad, ph.m_pad, lw.out_p, kHidden, kHidden, lw.out_b, ph.y,
                    terms, y16 ? GEMM_OUT_BF16 : GEMM_OUT_F32, 0, nodrop, tiles_dyn);
  if (rc) return rc;
  ActOut xo{ph.x, ph.xb, planes, ph.m_pad};
  // on the pad-skipping layout the row bound is the allocation and the live count comes from the device
  const long long ln_rows = dyn != nullptr ? ph.m_pad : ph.m_valid;
  {
    ProfileScope prof(h, stream, STLT_PROF_ADD_LN);
    if (y16) STLT_CUDA(h, launch_add_ln_bf16y(ph.x, y_b, lw.n1_g, lw.n1_b, eps, ln_rows, xo, stream, nullptr, nodrop, rows_dyn));
    else STLT_CUDA(h, launch_add_ln(ph.x, ph.y, lw.n1_g, lw.n1_b, eps, ln_rows, xo, stream, nullptr, nodrop, rows_dyn));
  }
  h->launches++;
  rc = run_gemm(h, stream, ph.xb, ph.m_pad, ph.m_pad, lw.l1_p, kFfn, kHidden, lw.l1_b, ph.hid, terms,
                fp32 ? GEMM_OUT_BF16_SPLIT : GEMM_OUT_BF16, fp32 ? 1 : 2, nodrop, tiles_dyn);
  if (rc) return rc;
  rc = run_gemm(h, stream, ph.hid, ph.m_pad, ph.m_pad, lw.l2_p, kHidden, kFfn, lw.l2_b, ph.y, terms,
                y16 ? GEMM_OUT_BF16 : GEMM_OUT_F32, 0, nodrop, tiles_dyn);
  if (rc) return rc;
  {
    ProfileScope prof(h, stream, STLT_PROF_ADD_LN);
    if (y16) STLT_CUDA(h, launch_add_ln_bf16y(ph.x, y_b, lw.n2_g, lw.n2_b, eps, ln_rows, xo, stream, nullptr, nodrop, rows_dyn));
    else STLT_CUDA(h, launch_add_ln(ph.x, ph.y, lw.n2_g, lw.n2_b, eps, ln_rows, xo, stream, nullptr, nodrop, rows_dyn));
  }
  h->launches++;
  return STLT_OK;
}


// ---- fused-LayerNorm forward (bf16 mode; DESIGN.md "LayerNorm fused into the GEMM epilogues") ----------
// The residual stream of a stack is kept PRE-norm: ph.x holds z (fp32), ph.xb its bf16 copy, and a float2 per
// row holds (sum, sum of squares) of z. `pending` describes the LayerNorm that still has to be applied to z
// (null gamma: z is already a normalised activation, i.e. the embedding output).
struct PendingNorm {
  const float2* stats = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
};

// in-projection (+ deferred LayerNorm of its input) and attention over the full phase
int fused_attention_part(Handle* h, cudaStream_t stream, const LayerWeights& lw, const Phase& ph,
                         const PendingNorm& in, const long long* mask_src, long long num_seqs, int T, bool causal,
                         const int* dyn = nullptr, bool fp32 = false) {
  int rc;
  if (!fp32 && (dyn != nullptr || (h->fused_attn && T <= h->fused_attn_max_t && lw.in_h != nullptr)))  // one kernel: the packed QKV activations never reach HBM
    return run_qkv_attention(h, stream, ph.xb, ph.m_pad, ph.m_valid, lw.in_h, lw.in_hs, lw.in_hc,
                             in.gamma != nullptr ? in.stats : nullptr, h->dims.encoder_norm_eps, mask_src, num_seqs, T,
                             causal, ph.att, dyn);
  const int terms = fp32 ? 3 : 1;
  const int* tiles_dyn = dyn != nullptr ? dyn + kDynTiles : nullptr;
  if (in.gamma == nullptr) {
    rc = run_gemm(h, stream, ph.xb, ph.m_pad, ph.m_pad, lw.in_p, kQkv, kHidden, lw.in_b, ph.qkv, terms,
                  fp32 ? GEMM_OUT_BF16_SPLIT : GEMM_OUT_BF16, 0, DropCfg{0, 0, 1.f}, tiles_dyn);
  } else {
    EpiArgs e{in.stats, lw.in_s, lw.in_c, nullptr, nullptr, nullptr, h->dims.encoder_norm_eps, 1};
    rc = run_gemm_fused(h, stream, GEMM_EPI_NORM_A, ph.xb, ph.m_pad, lw.in_f, kQkv, kHidden, nullptr, ph.qkv, nullptr, 0, e,
                        tiles_dyn, terms);
  }
  if (rc) return rc;
  // fp32-parity mode: the attention stays the separate split-plane kernel (on the pad-skipping layout: one launch per row region)
  if (fp32) return run_attention_only(h, stream, true, ph, mask_src, num_seqs, T, causal, dyn);
  ActOut att{nullptr, ph.att, 1, ph.m_pad};
  {
    ProfileScope prof(h, stream, STLT_PROF_ATTENTION);
    STLT_CUDA(h, launch_attention(ph.qkv, true, mask_src, num_seqs, T, causal, att, stream));
  }
  h->launches++;
  return STLT_OK;
}

// out-projection + residual, linear1 (+ LN1, GELU), linear2 + residual; LN2 stays pending (stats in s2)
int fused_tail_part(Handle* h, cudaStream_t stream, const LayerWeights& lw, const Phase& ph, const PendingNorm& in,
                    float2* s1, float2* s2, const int* tiles_dyn = nullptr, bool fp32 = false) {
  const float eps = h->dims.encoder_norm_eps;
  const int terms = fp32 ? 3 : 1;  // fp32-parity mode: split operands, exact erf GELU, hi / lo copy of the new z
  EpiArgs e1{in.stats, in.gamma, in.beta, ph.x, s1, ph.xb, eps, in.gamma != nullptr ? 1 : 0};
  e1.z_lo = ph.xlo;
  int rc = run_gemm_fused(h, stream, GEMM_EPI_RESID, ph.att, ph.m_pad, lw.out_p, kHidden, kHidden, lw.out_b, ph.x, ph.xb, 0, e1,
                          tiles_dyn, terms);
  if (rc) return rc;
  EpiArgs e2{s1, lw.l1_s, lw.l1_c, nullptr, nullptr, nullptr, eps, 1};
  rc = run_gemm_fused(h, stream, GEMM_EPI_NORM_A, ph.xb, ph.m_pad, lw.l1_f, kFfn, kHidden, nullptr, ph.hid, nullptr,
                      fp32 ? 1 : 2, e2, tiles_dyn, terms);
  if (rc) return rc;
  EpiArgs e3{s1, lw.n1_g, lw.n1_b, ph.x, s2, ph.xb, eps, 1};
  e3.z_lo = ph.xlo;
  return run_gemm_fused(h, stream, GEMM_EPI_RESID, ph.hid, ph.m_pad, lw.l2_p, kHidden, kFfn, lw.l2_b, ph.x, ph.xb, 0, e3,
                        tiles_dyn, terms);
}

// One stack of post-norm encoder layers; the last layer's row-wise tail runs on the compacted rows of `tail`
// (gathered with `stride` / `lengths`). On return tail.x holds the PRE-norm output of the stack on those rows
// and *out the LayerNorm still to be applied to it.
// dyn / frame_row != null: `full` is on the pad-skipping layout (compact.cu) and the tail rows are the frames in padded order.
int fused_stack(Handle* h, cudaStream_t stream, const std::vector<LayerWeights>& layers, const Phase& full,
                const Phase& tail, const long long* mask_src, long long num_seqs, int T, bool causal, int stride,
                const long long* lengths, int L, float2* stats_full, float2* stats_tail, int* err_flag,
                PendingNorm* out, const int* dyn = nullptr, const int* frame_row = nullptr, bool prune_last = true,
                bool fp32 = false) {
  const int n = static_cast<int>(layers.size());
  const int planes = fp32 ? 2 : 1;  // bf16 planes of the context rows the gather compacts
  PendingNorm pending;  // layer 0 reads the (already normalised) embedding output
  for (int i = 0; i < n; ++i) {
    const LayerWeights& lw = layers[i];
    int rc = fused_attention_part(h, stream, lw, full, pending, mask_src, num_seqs, T, causal, dyn, fp32);
    if (rc) return rc;
    if (i < n - 1 || !prune_last) {  // !prune_last (CACNF: every frame token is consumed): full.x keeps the PRE-norm output
      float2* s1 = stats_full + static_cast<size_t>(2 * i) * full.m_pad * kStatSlots;
      float2* s2 = s1 + full.m_pad * kStatSlots;
      rc = fused_tail_part(h, stream, lw, full, pending, s1, s2, dyn != nullptr ? dyn + kDynTiles : nullptr, fp32);
      if (rc) return rc;
      pending = PendingNorm{s2, lw.n2_g, lw.n2_b};
      if (i == n - 1) *out = pending;
    } else {
      float2* sc_in = stats_tail;
      float2* sc1 = stats_tail + tail.m_pad * kStatSlots;
      float2* sc2 = sc1 + tail.m_pad * kStatSlots;
      {
        ProfileScope prof(h, stream, STLT_PROF_OTHER);
        const __nv_bfloat16* hi = full.xlo != nullptr ? full.xb : nullptr;
        if (frame_row != nullptr)
          STLT_CUDA(h, launch_gather_frames(full.x, full.att, frame_row, tail.m_valid, tail.x, tail.att, pending.stats,
                                            sc_in, stream, hi, full.xlo, planes, full.m_pad, tail.m_pad));
        else
          STLT_CUDA(h, launch_gather_rows(full.x, full.att, planes, full.m_pad, stride, lengths, L, tail.m_valid, tail.x,
                                          tail.att, tail.m_pad, err_flag, stream, pending.stats, sc_in, hi, full.xlo));
      }
      h->launches++;
      PendingNorm tail_in{pending.gamma ? sc_in : nullptr, pending.gamma, pending.beta};
      rc = fused_tail_part(h, stream, lw, tail, tail_in, sc1, sc2, nullptr, fp32);
      if (rc) return rc;
      *out = PendingNorm{sc2, lw.n2_g, lw.n2_b};
    }
  }
  return STLT_OK;
}

}  // namespace

extern "C" {

int stlt_create(const StltDims* dims, void** handle) {
  if (!dims || !handle) return fail(nullptr, STLT_ERR_INVALID, "null argument");
  if (dims->hidden_size != kHidden || dims->num_heads != kHeads)
    return fail(nullptr, STLT_ERR_INVALID,
                "kernels are specialised for hidden_size=768, num_heads=12 (got %d, %d)",
                dims->hidden_size, dims->num_heads);
  if (dims->num_spatial_layers < 0 || dims->num_temporal_layers < 0 || dims->unique_categories < 1 ||
      dims->num_classes < 1 || dims->max_positions < 1 || dims->num_frame_types < 1)
    return fail(nullptr, STLT_ERR_INVALID, "invalid model dimensions");
  Handle* h = new Handle();
  h->dims = *dims;
  h->w.spatial.resize(dims->num_spatial_layers);
  h->w.temporal.resize(dims->num_temporal_layers);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  cudaDeviceProp prop{};
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    fail(nullptr, STLT_ERR_CUDA, "no usable CUDA device: %s", cudaGetErrorString(e));
    delete h;
    return STLT_ERR_CUDA;
  }
  if (prop.major != 10) {
    fail(nullptr, STLT_ERR_CUDA, "libstlt_b200 requires an sm_100 GPU (found sm_%d%d); no fallback",
         prop.major, prop.minor);
    delete h;
    return STLT_ERR_CUDA;
  }
  h->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    fail(nullptr, STLT_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    delete h;
    return STLT_ERR_CUDA;
  }
  h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  if (const char* dbg = getenv("STLT_QKV_ATTN_DEBUG")) h->qkv_attn_debug = atoi(dbg);
  if (const char* mt = getenv("STLT_FUSED_ATTENTION_MAX_T")) {
    const int v = atoi(mt);
    h->fused_attn_max_t = v < 0 ? 0 : (v > 32 ? 32 : v);
  }
  *handle = h;
  return STLT_OK;
}

int stlt_destroy(void* handle) {
  Handle* h = static_cast<Handle*>(handle);
  if (h)
  {
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : h->bwd_stage_events) cudaEventDestroy(e);
  }
  delete h;
  return STLT_OK;
}

const char* stlt_last_error(void* handle) {
  return handle ? static_cast<Handle*>(handle)->err : g_err;
}

int stlt_bind_weights(void* handle, const StltTensor* tensors, int32_t count) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !tensors) return fail(h, STLT_ERR_INVALID, "null argument");
  Weights w;
  int rc = bind_table(h, tensors, count, &w, true);
  if (rc) return rc;
  h->w = w;
  h->bound = true;
  h->packed_precision = -1;
  h->packed_ptr = nullptr;
  return STLT_OK;
}

int stlt_packed_weights_bytes(void* handle, int32_t precision, size_t* bytes) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !bytes) return fail(h, STLT_ERR_INVALID, "null argument");
  if (precision != STLT_PRECISION_FP32 && precision != STLT_PRECISION_BF16)
    return fail(h, STLT_ERR_INVALID, "unknown precision %d", precision);
  const size_t planes = precision == STLT_PRECISION_FP32 ? 2 : 1;
  const size_t per_layer = static_cast<size_t>(kHidden) * (kQkv + kHidden + 2 * kFfn);
  const size_t layers = h->dims.num_spatial_layers + h->dims.num_temporal_layers;
  *bytes = layers * per_layer * planes * 2;
  if (precision == STLT_PRECISION_BF16)  // fused-LayerNorm copies: folded in-proj (row-major and head-major) / linear1 + their s, c vectors
    *bytes += layers * (static_cast<size_t>(kHidden) * (2 * kQkv + kFfn) * 2 + static_cast<size_t>(2 * kQkv + kFfn) * 2 * 4);
  else  // fp32-parity mode: folded in-proj / linear1 as hi / lo plane pairs + their s, c vectors
    *bytes += layers * (static_cast<size_t>(kHidden) * (kQkv + kFfn) * 2 * 2 + static_cast<size_t>(kQkv + kFfn) * 2 * 4);
  return STLT_OK;
}

int stlt_pack_weights(void* handle, void* stream_, int32_t precision, void* packed, size_t bytes) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !packed) return fail(h, STLT_ERR_INVALID, "null argument");
  if (!h->bound) return fail(h, STLT_ERR_STATE, "stlt_bind_weights has not been called");
  size_t need = 0;
  int rc = stlt_packed_weights_bytes(handle, precision, &need);
  if (rc) return rc;
  if (bytes < need) return fail(h, STLT_ERR_INVALID, "packed buffer too small: %zu < %zu", bytes, need);
  if ((reinterpret_cast<uintptr_t>(packed) & 127) != 0)
    return fail(h, STLT_ERR_INVALID, "packed buffer must be 128-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int planes = precision == STLT_PRECISION_FP32 ? 2 : 1;
  __nv_bfloat16* cur = static_cast<__nv_bfloat16*>(packed);
  auto pack = [&](const float* src, long long n, const __nv_bfloat16** dst) -> cudaError_t {
    *dst = cur;
    cudaError_t e = launch_pack_bf16(src, cur, n, planes, stream);
    cur += n * planes;
    return e;
  };
  auto pack_layer = [&](LayerWeights& lw) -> cudaError_t {
    cudaError_t e;
    if ((e = pack(lw.in_w, static_cast<long long>(kQkv) * kHidden, &lw.in_p)) != cudaSuccess) return e;
    if ((e = pack(lw.out_w, static_cast<long long>(kHidden) * kHidden, &lw.out_p)) != cudaSuccess) return e;
    if ((e = pack(lw.l1_w, static_cast<long long>(kFfn) * kHidden, &lw.l1_p)) != cudaSuccess) return e;
    if ((e = pack(lw.l2_w, static_cast<long long>(kHidden) * kFfn, &lw.l2_p)) != cudaSuccess) return e;
    return cudaSuccess;
  };
  for (auto& lw : h->w.spatial) STLT_CUDA(h, pack_layer(lw));
  for (auto& lw : h->w.temporal) STLT_CUDA(h, pack_layer(lw));
  if (precision == STLT_PRECISION_BF16) {
    // fused-LayerNorm section (see GemmEpilogue): per layer [in_f | l1_f | in_h] bf16, then
    // [in_s in_c l1_s l1_c in_hs in_hc] f32
    auto fold_stack = [&](std::vector<LayerWeights>& stack) -> cudaError_t {
      for (size_t i = 0; i < stack.size(); ++i) {
        LayerWeights& lw = stack[i];
        __nv_bfloat16* in_f = cur;
        __nv_bfloat16* l1_f = in_f + static_cast<size_t>(kQkv) * kHidden;
        __nv_bfloat16* in_h = l1_f + static_cast<size_t>(kFfn) * kHidden;
        float* vec = reinterpret_cast<float*>(in_h + static_cast<size_t>(kQkv) * kHidden);
        cur = reinterpret_cast<__nv_bfloat16*>(vec + 2 * (2 * kQkv + kFfn));
        lw.in_f = nullptr;
        lw.in_s = lw.in_c = nullptr;
        cudaError_t e;
        {  // head-major copy for the attention-fused in-projection; the first layer of a stack reads a normalised input
          const float* g2 = i > 0 ? stack[i - 1].n2_g : nullptr;
          const float* b2 = i > 0 ? stack[i - 1].n2_b : nullptr;
          float* vh = vec + 2 * (kQkv + kFfn);
          e = launch_pack_folded(lw.in_w, g2, b2, lw.in_b, kQkv, kHidden, in_h, vh, vh + kQkv, stream, true);
          if (e != cudaSuccess) return e;
          lw.in_h = in_h;
          lw.in_hs = vh;
          lw.in_hc = vh + kQkv;
        }
        if (i > 0) {  // the in-projection reads LN2 of the previous layer
          const LayerWeights& prev = stack[i - 1];
          e = launch_pack_folded(lw.in_w, prev.n2_g, prev.n2_b, lw.in_b, kQkv, kHidden, in_f, vec, vec + kQkv, stream);
          if (e != cudaSuccess) return e;
          lw.in_f = in_f;
          lw.in_s = vec;
          lw.in_c = vec + kQkv;
        }
        float* v1 = vec + 2 * kQkv;
        e = launch_pack_folded(lw.l1_w, lw.n1_g, lw.n1_b, lw.l1_b, kFfn, kHidden, l1_f, v1, v1 + kFfn, stream);
        if (e != cudaSuccess) return e;
        lw.l1_f = l1_f;
        lw.l1_s = v1;
        lw.l1_c = v1 + kFfn;
      }
      return cudaSuccess;
    };
    STLT_CUDA(h, fold_stack(h->w.spatial));
    STLT_CUDA(h, fold_stack(h->w.temporal));
  }
  if (precision == STLT_PRECISION_FP32) {
    // fused-LayerNorm section of the fp32-parity mode: per layer [in_f hi | in_f lo | l1_f hi | l1_f lo] bf16, then
    // [in_s in_c l1_s l1_c] f32 (no head-major copy: the attention stays a separate kernel in this mode)
    auto fold_stack = [&](std::vector<LayerWeights>& stack) -> cudaError_t {
      for (size_t i = 0; i < stack.size(); ++i) {
        LayerWeights& lw = stack[i];
        __nv_bfloat16* in_f = cur;
        __nv_bfloat16* l1_f = in_f + static_cast<size_t>(kQkv) * kHidden * 2;
        float* vec = reinterpret_cast<float*>(l1_f + static_cast<size_t>(kFfn) * kHidden * 2);
        cur = reinterpret_cast<__nv_bfloat16*>(vec + 2 * (kQkv + kFfn));
        lw.in_f = lw.in_h = nullptr;
        lw.in_s = lw.in_c = lw.in_hs = lw.in_hc = nullptr;
        cudaError_t e;
        if (i > 0) {  // the in-projection reads LN2 of the previous layer
          const LayerWeights& prev = stack[i - 1];
          e = launch_pack_folded(lw.in_w, prev.n2_g, prev.n2_b, lw.in_b, kQkv, kHidden, in_f, vec, vec + kQkv, stream, false,
                                 true);
          if (e != cudaSuccess) return e;
          lw.in_f = in_f;
          lw.in_s = vec;
          lw.in_c = vec + kQkv;
        }
        float* v1 = vec + 2 * kQkv;
        e = launch_pack_folded(lw.l1_w, lw.n1_g, lw.n1_b, lw.l1_b, kFfn, kHidden, l1_f, v1, v1 + kFfn, stream, false, true);
        if (e != cudaSuccess) return e;
        lw.l1_f = l1_f;
        lw.l1_s = v1;
        lw.l1_c = v1 + kFfn;
      }
      return cudaSuccess;
    };
    STLT_CUDA(h, fold_stack(h->w.spatial));
    STLT_CUDA(h, fold_stack(h->w.temporal));
  }
  h->packed_precision = precision;
  h->packed_ptr = packed;
  return STLT_OK;
}

int stlt_workspace_bytes(void* handle, int32_t B, int32_t L, int32_t S, int32_t precision,
                         size_t* bytes) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !bytes) return fail(h, STLT_ERR_INVALID, "null argument");
  if (B < 0 || L < 1 || S < 1) return fail(h, STLT_ERR_INVALID, "invalid batch shape");
  *bytes = plan_workspace(B, L, S, precision, h->dims.num_spatial_layers, h->dims.num_temporal_layers).total;
  return STLT_OK;
}

int stlt_prepare(void* handle, void* stream, const double* raw_boxes, const int64_t* video_sizes,
                 const int64_t* categories, const int64_t* frame_types, int32_t B, int32_t L,
                 int32_t S, float* boxes_out, uint8_t* mask_boxes_out, uint8_t* mask_frames_out) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (B < 0 || L < 1 || S < 1) return fail(h, STLT_ERR_INVALID, "invalid batch shape");
  if (B == 0) return STLT_OK;
  if (!raw_boxes || !video_sizes || !categories || !frame_types || !boxes_out || !mask_boxes_out ||
      !mask_frames_out)
    return fail(h, STLT_ERR_INVALID, "null tensor pointer");
  STLT_CUDA(h, launch_prepare(raw_boxes, reinterpret_cast<const long long*>(video_sizes),
                              reinterpret_cast<const long long*>(categories),
                              reinterpret_cast<const long long*>(frame_types), B, L, S, boxes_out,
                              mask_boxes_out, mask_frames_out, static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

int stlt_build_batch(void* handle, void* stream_, const StltLayoutStore* store, const StltLayoutIds* ids,
                     const int64_t* video_index, const int64_t* frame_indices, const int64_t* num_sampled,
                     int32_t B, int32_t T, int32_t L, int32_t S, double score_threshold,
                     int64_t* categories, float* boxes, float* scores, int64_t* frame_types,
                     int64_t* lengths, uint8_t* mask_boxes, uint8_t* mask_frames, int32_t* status) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (B < 0 || T < 1 || L < 1 || S < 1) return fail(h, STLT_ERR_INVALID, "invalid batch shape");
  if (B == 0) return STLT_OK;
  if (!store || !ids || !video_index || !categories || !boxes || !frame_types || !lengths ||
      !mask_boxes || !mask_frames || !status)
    return fail(h, STLT_ERR_INVALID, "null pointer");
  if ((frame_indices == nullptr) != (num_sampled == nullptr))
    return fail(h, STLT_ERR_INVALID, "frame_indices and num_sampled must be given together");
  if (!store->video_frame_offsets || !store->frame_object_offsets || !store->obj_boxes ||
      !store->obj_categories || !store->obj_scores || !store->video_sizes)
    return fail(h, STLT_ERR_INVALID, "incomplete layout store");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BatchStore st{reinterpret_cast<const long long*>(store->video_frame_offsets),
                reinterpret_cast<const long long*>(store->frame_object_offsets), store->obj_boxes,
                reinterpret_cast<const long long*>(store->obj_categories), store->obj_scores,
                reinterpret_cast<const long long*>(store->video_sizes)};
  BatchIds bi{ids->cls, ids->ft_pad, ids->ft_regular, ids->ft_empty, ids->ft_extract};
  STLT_CUDA(h, cudaMemsetAsync(status, 0, sizeof(int32_t), stream));
  STLT_CUDA(h, launch_build_batch(st, reinterpret_cast<const long long*>(video_index),
                                  reinterpret_cast<const long long*>(frame_indices),
                                  reinterpret_cast<const long long*>(num_sampled), B, T, L, S,
                                  score_threshold, bi, reinterpret_cast<long long*>(categories), boxes,
                                  scores, reinterpret_cast<long long*>(frame_types),
                                  reinterpret_cast<long long*>(lengths), mask_boxes, mask_frames, status,
                                  stream));
  return STLT_OK;
}

int stlt_topk_count(void* handle, void* stream, const float* logits, const int64_t* labels, int32_t rows,
                    int32_t classes, uint64_t* counters) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (rows < 0 || classes < 1) return fail(h, STLT_ERR_INVALID, "invalid shape");
  if (rows == 0) return STLT_OK;
  if (!logits || !labels || !counters) return fail(h, STLT_ERR_INVALID, "null pointer");
  STLT_CUDA(h, launch_topk_count(logits, reinterpret_cast<const long long*>(labels), rows, classes,
                                 reinterpret_cast<unsigned long long*>(counters),
                                 static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

int stlt_map_accumulate(void* handle, void* stream, const float* logits, const float* labels, int32_t rows,
                        int32_t classes, float* predictions_out, float* ground_truths_out) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (rows < 0 || classes < 1) return fail(h, STLT_ERR_INVALID, "invalid shape");
  if (rows == 0) return STLT_OK;
  if (!logits || !labels || !predictions_out || !ground_truths_out) return fail(h, STLT_ERR_INVALID, "null pointer");
  STLT_CUDA(h, launch_map_accumulate(logits, labels, static_cast<long long>(rows) * classes, predictions_out,
                                     ground_truths_out, static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

int stlt_charades_map(void* handle, void* stream, const float* predictions, const float* ground_truths,
                      int32_t instances, int32_t classes, double* ap_out, double* map_out) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (!predictions || !ground_truths || !ap_out || !map_out) return fail(h, STLT_ERR_INVALID, "null pointer");
  if (instances < 1 || instances > 32768 || classes < 1)
    return fail(h, STLT_ERR_INVALID, "instances must be in [1, 32768] (one shared-memory sort per class)");
  STLT_CUDA(h, launch_charades_map(predictions, ground_truths, instances, classes, ap_out, map_out,
                                   static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

int stlt_forward(void* handle, void* stream_, int32_t precision, const int64_t* categories_,
                 const float* boxes, const float* scores, const int64_t* frame_types_,
                 const int64_t* lengths_, int32_t B, int32_t L, int32_t S, void* workspace,
                 size_t workspace_bytes, float* logits, uint8_t* mask_boxes, uint8_t* mask_frames) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  const StltDims& d = h->dims;
  if (!h->bound) return fail(h, STLT_ERR_STATE, "stlt_bind_weights has not been called");
  if (precision != STLT_PRECISION_FP32 && precision != STLT_PRECISION_BF16)
    return fail(h, STLT_ERR_INVALID, "unknown precision %d", precision);
  if (h->packed_precision != precision)
    return fail(h, STLT_ERR_STATE, "weights are not packed for precision %d (call stlt_pack_weights)",
                precision);
  if (B < 0) return fail(h, STLT_ERR_INVALID, "negative batch size");
  if (L < 1 || L > d.max_positions)
    return fail(h, STLT_ERR_INVALID, "frames=%d outside [1, %d] (position table)", L, d.max_positions);
  if (L > 256) return fail(h, STLT_ERR_INVALID, "frames=%d > 256 is not supported by the attention kernels", L);
  if (S < 1 || S > 64) return fail(h, STLT_ERR_INVALID, "slots=%d outside [1, 64]", S);
  h->launches = 0;
  if (B == 0) return STLT_OK;
  if (!categories_ || !boxes || !frame_types_ || !lengths_ || !workspace || !logits)
    return fail(h, STLT_ERR_INVALID, "null tensor pointer");
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0)
    return fail(h, STLT_ERR_INVALID, "workspace must be 1024-byte aligned");
  const WorkspacePlan p = plan_workspace(B, L, S, precision, d.num_spatial_layers, d.num_temporal_layers);
  if (workspace_bytes < p.total)
    return fail(h, STLT_ERR_INVALID, "workspace too small: %zu < %zu", workspace_bytes, p.total);
  if (p.sp.m_pad > 0x7fffffffLL / 2) return fail(h, STLT_ERR_INVALID, "batch too large for one call");

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long* categories = reinterpret_cast<const long long*>(categories_);
  const long long* frame_types = reinterpret_cast<const long long*>(frame_types_);
  const long long* lengths = reinterpret_cast<const long long*>(lengths_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const bool fp32 = precision == STLT_PRECISION_FP32;
  const int planes = fp32 ? 2 : 1;
  int* err_flag = reinterpret_cast<int*>(ws + p.off_err);
  const long long n_sp = static_cast<long long>(B) * L * S;
  const long long n_tm = static_cast<long long>(B) * L;

  STLT_CUDA(h, cudaMemsetAsync(err_flag, 0, sizeof(int), stream));
  if (mask_boxes || mask_frames) {
    STLT_CUDA(h, launch_masks(categories, frame_types, n_sp, n_tm, mask_boxes, mask_frames, stream));
    h->launches++;
  }

  // ---- spatial phase: B*L sequences of S object tokens (models.py:57-81) ----
  const Phase sp = make_phase(ws, p.sp, n_sp);
  const Phase tm = make_phase(ws, p.tm, n_tm);
  const Phase hd = make_phase(ws, p.hd, B);
  // Only slot 0 of the spatial output (models.py:79) and only frame lengths-1 of the temporal output
  // (models.py:192) are ever read, so the row-wise tail of the LAST layer of each stack runs on
  // those rows only. Taps that expose the full stack output switch the pruning off.
  const bool prune_sp = h->pruning && h->taps.spatial == nullptr && d.num_spatial_layers > 0;
  const bool prune_tm = h->pruning && h->taps.temporal == nullptr && h->cap_tm_x == nullptr &&
                        d.num_temporal_layers > 0;

  // fp32-parity mode: the same epilogue fusions on split operands (stlt_set_fused_ln_fp32); CACNF's capture of every frame
  // token keeps the separate kernels there
  const bool fused = (fp32 ? h->fused_ln_fp32 && h->cap_tm_x == nullptr : h->fused_ln) && h->pruning &&
                     d.num_spatial_layers > 0 && d.num_temporal_layers > 0 &&
                     h->taps.embed == nullptr && h->taps.spatial == nullptr && h->taps.frames == nullptr &&
                     h->taps.temporal == nullptr && h->taps.pooled == nullptr && h->w.spatial[0].l1_f != nullptr;
  if (fused) {
    // ---- LayerNorm folded into the GEMM epilogues (no add_ln launches) ----
    float2* stats = reinterpret_cast<float2*>(ws + p.off_stats);
    float2* st_sp = stats;
    float2* st_tm = st_sp + static_cast<size_t>(2 * d.num_spatial_layers) * p.sp.m_pad * kStatSlots;
    float2* st_c_sp = st_tm + static_cast<size_t>(2 * d.num_temporal_layers) * p.tm.m_pad * kStatSlots;  // 3 compaction sites
    float2* st_c_tm = st_c_sp + 3 * p.tm.m_pad * kStatSlots;                                               // 3 more (B rows)
    // the residual stream of the two full phases lives as two bf16 planes (hi = GEMM operand, lo = remainder)
    Phase sp_f = sp, tm_f = tm;
    const bool hilo = h->hilo && !fp32;
    if (hilo) {
      sp_f.xlo = sp.xb + static_cast<size_t>(sp.m_pad) * kHidden;
      tm_f.xlo = tm.xb + static_cast<size_t>(tm.m_pad) * kHidden;
    }
    ActOut emb{hilo ? nullptr : sp.x, sp.xb, hilo || fp32 ? 2 : 1, sp.m_pad};
    // Pad-skipping layout of the spatial phase (compact.cu): padding frames and the padded slots of one-token frames
    // are not computed. bf16: needs the attention-fused in-projection (it understands the two row regions); fp32-parity
    // mode: the separate attention kernel runs once per row region.
    const bool compact = h->compaction && p.off_plan != 0 &&
                         (fp32 ? S <= 32 : h->fused_attn && S <= h->fused_attn_max_t && h->w.spatial[0].in_h != nullptr);
    const int* dyn = nullptr;
    const int* frame_row = nullptr;
    const long long* sp_mask = categories;
    if (compact) {
      int* hdr = reinterpret_cast<int*>(ws + p.off_plan);
      int* fr = reinterpret_cast<int*>(ws + p.off_frame_row);
      long long* mask_c = reinterpret_cast<long long*>(ws + p.off_mask);
      {
        ProfileScope prof(h, stream, STLT_PROF_OTHER);
        STLT_CUDA(h, launch_compact_plan(categories, lengths, B, L, S, fr, hdr, ws + p.off_plan_scratch, err_flag, stream));
      }
      h->launches += 3;
      dyn = hdr;
      frame_row = fr;
      sp_mask = mask_c;
    }
    {
      ProfileScope prof(h, stream, STLT_PROF_OTHER);
      STLT_CUDA(h, launch_embed(categories, boxes, scores, h->w.cat_table, d.unique_categories, h->w.box_w, h->w.box_b,
                                h->w.score_w, h->w.score_b, h->w.emb_g, h->w.emb_b, d.layer_norm_eps, n_sp, emb,
                                err_flag, stream, reinterpret_cast<float*>(sp.hid), embed_scratch(sp), DropCfg{0, 0, 1.f},
                                frame_row, S, compact ? reinterpret_cast<long long*>(ws + p.off_mask) : nullptr));
    }
    h->launches += 2;  // embed_stats_kernel + embed_kernel
    PendingNorm sp_out;
    int rc = fused_stack(h, stream, h->w.spatial, sp_f, tm, sp_mask, n_tm, S, false, S, nullptr, 0, st_sp, st_c_sp,
                         err_flag, &sp_out, dyn, frame_row, true, fp32);
    if (rc) return rc;
    {
      ProfileScope prof(h, stream, STLT_PROF_OTHER);
      ActOut fr{hilo ? nullptr : tm.x, tm.xb, hilo || fp32 ? 2 : 1, tm.m_pad};
      // the spatial stack's LayerNorm-2 is applied on the fly (row statistics recomputed in registers)
      STLT_CUDA(h, launch_frame_embed(tm.x, 1, frame_types, h->w.pos_table, h->w.ft_table, d.num_frame_types,
                                      h->w.fr_g, h->w.fr_b, d.layer_norm_eps, B, L, fr, err_flag, stream,
                                      DropCfg{0, 0, 1.f}, sp_out.gamma, sp_out.beta, d.encoder_norm_eps));
    }
    h->launches++;
    PendingNorm tm_out;
    const bool capture = h->cap_tm_x != nullptr;  // CACNF: the fusion layers consume every frame token of the stack
    if (capture && hilo) return fail(h, STLT_ERR_STATE, "the two-plane residual stream is not available under CACNF");
    rc = fused_stack(h, stream, h->w.temporal, tm_f, hd, frame_types, B, L, true, 0, lengths, L, st_tm, st_c_tm, err_flag,
                     &tm_out, nullptr, nullptr, !capture, fp32);
    if (rc) return rc;
    float* h1f = reinterpret_cast<float*>(ws + p.off_head);
    float* h2f = h1f + static_cast<size_t>(B) * kHidden;
    float* pooledf = h2f + static_cast<size_t>(B) * kHidden;
    {
      ProfileScope prof(h, stream, STLT_PROF_OTHER);
      ActOut po{pooledf, nullptr, 1, 0};
      if (capture) {
        // the stack's last LayerNorm over all frames, straight into the caller's buffers (fp32 stream + bf16 operand)
        ActOut cap{h->cap_tm_x, h->cap_tm_xb, 1, tm.m_pad};
        STLT_CUDA(h, launch_add_ln(tm.x, nullptr, tm_out.gamma, tm_out.beta, d.encoder_norm_eps, n_tm, cap, stream));
        STLT_CUDA(h, launch_gather_last(h->cap_tm_x, lengths, B, L, pooledf, err_flag, stream));
        h->launches++;
      } else
      STLT_CUDA(h, launch_add_ln(hd.x, nullptr, tm_out.gamma, tm_out.beta, d.encoder_norm_eps, B, po, stream));
      STLT_CUDA(h, launch_gemm_simt(pooledf, h->w.fc1_w, h->w.fc1_b, h1f, B, kHidden, kHidden, true, stream));
      ActOut ho{h2f, nullptr, 1, 0};
      STLT_CUDA(h, launch_add_ln(h1f, nullptr, h->w.head_g, h->w.head_b, d.layer_norm_eps, B, ho, stream));
      STLT_CUDA(h, launch_gemm_simt(h2f, h->w.fc2_w, h->w.fc2_b, logits, B, d.num_classes, kHidden, false, stream));
      h->launches += 4;
    }
    return STLT_OK;
  }

  ActOut emb{sp.x, sp.xb, planes, sp.m_pad};
  // the pad-skipping layout of the spatial phase (compact.cu) on the separate kernels: fp32-parity mode, or bf16 with the
  // epilogue fusions switched off. Needs the pruned last layer (its tail is what maps back to the padded [B, L] grid).
  const bool compact_u = h->compaction && prune_sp && p.off_plan != 0 && S <= 32 && h->taps.embed == nullptr &&
                         h->taps.frames == nullptr;
  const int* dyn_u = nullptr;
  const int* frame_row_u = nullptr;
  const long long* sp_mask_u = categories;
  if (compact_u) {
    int* hdr = reinterpret_cast<int*>(ws + p.off_plan);
    int* fr = reinterpret_cast<int*>(ws + p.off_frame_row);
    {
      ProfileScope prof(h, stream, STLT_PROF_OTHER);
      STLT_CUDA(h, launch_compact_plan(categories, lengths, B, L, S, fr, hdr, ws + p.off_plan_scratch, err_flag, stream));
    }
    h->launches += 3;
    dyn_u = hdr;
    frame_row_u = fr;
    sp_mask_u = reinterpret_cast<long long*>(ws + p.off_mask);
  }
  {
    ProfileScope prof(h, stream, STLT_PROF_OTHER);
    STLT_CUDA(h, launch_embed(categories, boxes, scores, h->w.cat_table, d.unique_categories,
                              h->w.box_w, h->w.box_b, h->w.score_w, h->w.score_b, h->w.emb_g,
                              h->w.emb_b, d.layer_norm_eps, n_sp, emb, err_flag, stream,
                              reinterpret_cast<float*>(sp.hid), embed_scratch(sp), DropCfg{0, 0, 1.f}, frame_row_u, S,
                              compact_u ? reinterpret_cast<long long*>(ws + p.off_mask) : nullptr));
  }
  h->launches += 2;  // embed_stats_kernel + embed_kernel
  if (h->taps.embed)
    STLT_CUDA(h, cudaMemcpyAsync(h->taps.embed, sp.x, n_sp * kHidden * 4, cudaMemcpyDeviceToDevice, stream));

  bool cls_compact = false;  // spatial CLS rows already compacted into tm.x
  for (int i = 0; i < d.num_spatial_layers; ++i) {
    const LayerWeights& lw = h->w.spatial[i];
    int rc = run_attention_part(h, stream, precision, lw, sp, sp_mask_u, n_tm, S, false, dyn_u);
    if (rc) return rc;
    if (prune_sp && i == d.num_spatial_layers - 1) {
      {
        ProfileScope prof(h, stream, STLT_PROF_OTHER);
        if (compact_u)
          STLT_CUDA(h, launch_gather_frames(sp.x, sp.att, frame_row_u, n_tm, tm.x, tm.att, nullptr, nullptr, stream, nullptr,
                                            nullptr, planes, sp.m_pad, tm.m_pad));
        else
          STLT_CUDA(h, launch_gather_rows(sp.x, sp.att, planes, sp.m_pad, S, nullptr, 0, n_tm, tm.x, tm.att,
                                          tm.m_pad, err_flag, stream));
      }
      h->launches++;
      rc = run_tail_part(h, stream, precision, lw, tm);
      cls_compact = true;
    } else {
      rc = run_tail_part(h, stream, precision, lw, sp, dyn_u);
    }
    if (rc) return rc;
  }
  if (h->taps.spatial)
    STLT_CUDA(h, cudaMemcpyAsync(h->taps.spatial, sp.x, n_sp * kHidden * 4, cudaMemcpyDeviceToDevice, stream));

  // ---- temporal phase: B sequences of L frame tokens (models.py:98-111,136-152) ----
  ActOut fr{tm.x, tm.xb, planes, tm.m_pad};
  {
    ProfileScope prof(h, stream, STLT_PROF_OTHER);
    // in-place when the CLS rows were compacted (row f -> row f, one warp per row)
    STLT_CUDA(h, launch_frame_embed(cls_compact ? tm.x : sp.x, cls_compact ? 1 : S, frame_types,
                                    h->w.pos_table, h->w.ft_table, d.num_frame_types, h->w.fr_g,
                                    h->w.fr_b, d.layer_norm_eps, B, L, fr, err_flag, stream));
  }
  h->launches++;
  if (h->taps.frames)
    STLT_CUDA(h, cudaMemcpyAsync(h->taps.frames, tm.x, n_tm * kHidden * 4, cudaMemcpyDeviceToDevice, stream));

  float* pooled = reinterpret_cast<float*>(ws + p.off_head);
  float* h1 = pooled + static_cast<size_t>(B) * kHidden;
  float* h2 = h1 + static_cast<size_t>(B) * kHidden;
  bool pooled_done = false;
  for (int i = 0; i < d.num_temporal_layers; ++i) {
    const LayerWeights& lw = h->w.temporal[i];
    int rc = run_attention_part(h, stream, precision, lw, tm, frame_types, B, L, true);
    if (rc) return rc;
    if (prune_tm && i == d.num_temporal_layers - 1) {
      {
        ProfileScope prof(h, stream, STLT_PROF_OTHER);
        STLT_CUDA(h, launch_gather_rows(tm.x, tm.att, planes, tm.m_pad, 0, lengths, L, B, hd.x, hd.att,
                                        hd.m_pad, err_flag, stream));
      }
      h->launches++;
      rc = run_tail_part(h, stream, precision, lw, hd);
      pooled = hd.x;
      pooled_done = true;
    } else {
      rc = run_tail_part(h, stream, precision, lw, tm);
    }
    if (rc) return rc;
  }
  if (h->taps.temporal)
    STLT_CUDA(h, cudaMemcpyAsync(h->taps.temporal, tm.x, n_tm * kHidden * 4, cudaMemcpyDeviceToDevice, stream));
  if (h->cap_tm_x) {  // CACNF: the fusion layers consume every frame token (fp32 stream + bf16 GEMM operand)
    STLT_CUDA(h, cudaMemcpyAsync(h->cap_tm_x, tm.x, n_tm * kHidden * 4, cudaMemcpyDeviceToDevice, stream));
    for (int pl = 0; pl < planes; ++pl)  // same padded row count on both sides (pad128(B * L))
      STLT_CUDA(h, cudaMemcpyAsync(h->cap_tm_xb + static_cast<size_t>(pl) * tm.m_pad * kHidden,
                                   tm.xb + static_cast<size_t>(pl) * tm.m_pad * kHidden, n_tm * kHidden * 2,
                                   cudaMemcpyDeviceToDevice, stream));
  }

  // ---- head (models.py:155-163,189-193) ----
  {
    ProfileScope prof(h, stream, STLT_PROF_OTHER);
    if (!pooled_done) {
      STLT_CUDA(h, launch_gather_last(tm.x, lengths, B, L, pooled, err_flag, stream));
      h->launches++;
    }
    if (h->taps.pooled)
      STLT_CUDA(h, cudaMemcpyAsync(h->taps.pooled, pooled, static_cast<size_t>(B) * kHidden * 4,
                                   cudaMemcpyDeviceToDevice, stream));
    STLT_CUDA(h, launch_gemm_simt(pooled, h->w.fc1_w, h->w.fc1_b, h1, B, kHidden, kHidden, true, stream));
    h->launches++;
    ActOut ho{h2, nullptr, 1, 0};
    STLT_CUDA(h, launch_add_ln(h1, nullptr, h->w.head_g, h->w.head_b, d.layer_norm_eps, B, ho, stream));
    h->launches++;
    STLT_CUDA(h, launch_gemm_simt(h2, h->w.fc2_w, h->w.fc2_b, logits, B, d.num_classes, kHidden, false, stream));
    h->launches++;
  }
  return STLT_OK;
}

int stlt_set_fused_ln_fp32(void* handle, int32_t enable) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  h->fused_ln_fp32 = enable != 0;
  return STLT_OK;
}

int stlt_set_fused_ln(void* handle, int32_t enable) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  h->fused_ln = enable != 0;
  return STLT_OK;
}

int stlt_set_hilo_residual(void* handle, int32_t enable) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  h->hilo = enable != 0;
  return STLT_OK;
}

int stlt_set_compaction(void* handle, int32_t enable) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  h->compaction = enable != 0;
  return STLT_OK;
}

int stlt_set_fused_attention(void* handle, int32_t enable) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  h->fused_attn = enable != 0;
  return STLT_OK;
}

int stlt_set_pruning(void* handle, int32_t enable) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  h->pruning = enable != 0;
  return STLT_OK;
}

int stlt_check_errors(void* handle, void* stream, const void* workspace) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !workspace) return fail(h, STLT_ERR_INVALID, "null argument");
  int flag = 0;
  STLT_CUDA(h, cudaMemcpyAsync(&flag, workspace, sizeof(int), cudaMemcpyDeviceToHost,
                               static_cast<cudaStream_t>(stream)));
  STLT_CUDA(h, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  if (flag == 1) return fail(h, STLT_ERR_INPUT, "categories index out of range");
  if (flag == 2) return fail(h, STLT_ERR_INPUT, "frame_types index out of range");
  if (flag == 3) return fail(h, STLT_ERR_INPUT, "lengths outside [1, frames]");
  return STLT_OK;
}

int stlt_last_launch_count(void* handle) {
  Handle* h = static_cast<Handle*>(handle);
  return h ? h->launches : -1;
}

int stlt_set_taps(void* handle, const StltTaps* taps) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  if (taps)
    h->taps = *taps;
  else
    h->taps = StltTaps{};
  return STLT_OK;
}

int stlt_set_profiling(void* handle, int32_t enable) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  h->profiling = enable != 0;
  h->spans.clear();
  h->ev_used = 0;
  return STLT_OK;
}

int stlt_get_profile(void* handle, StltProfile* out) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !out) return fail(h, STLT_ERR_INVALID, "null argument");
  std::memset(out, 0, sizeof(*out));
  h->last_roles = StltRoleProfile{};
  for (const auto& sp : h->spans) {
    STLT_CUDA(h, cudaEventSynchronize(sp.b));
    float ms = 0.f;
    STLT_CUDA(h, cudaEventElapsedTime(&ms, sp.a, sp.b));
    if (sp.cat < 0 || sp.cat >= STLT_PROF_CATEGORIES) continue;
    out->ms[sp.cat] += ms;
    out->launches[sp.cat] += 1;
    double flops = sp.flops;
    if (sp.dyn != nullptr) {  // pad-skipping layout: units executed by this launch, from the device header
      int units = 0;
      STLT_CUDA(h, cudaMemcpy(&units, sp.dyn, sizeof(int), cudaMemcpyDeviceToHost));
      flops += sp.dyn_flops * units;
    }
    out->flops[sp.cat] += flops;
    if (sp.cat == STLT_PROF_GEMM) {
      const int r = sp.role >= 0 && sp.role < STLT_PROF_ROLES ? sp.role : STLT_PROF_ROLE_OTHER_GEMM;
      h->last_roles.ms[r] += ms;
      h->last_roles.flops[r] += flops;
      h->last_roles.launches[r] += 1;
    }
  }
  h->spans.clear();
  h->ev_used = 0;
  return STLT_OK;
}

int stlt_get_profile_by_role(void* handle, StltRoleProfile* out) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !out) return fail(h, STLT_ERR_INVALID, "null argument");
  *out = h->last_roles;
  return STLT_OK;
}

// ---- single-operator entry points -------------------------------------------------------------

int stlt_op_gemm(void* handle, void* stream, const void* a_planes, const void* w_planes,
                 const float* bias, void* out, int32_t m_rows, int32_t n, int32_t k, int32_t terms,
                 int32_t out_kind, int32_t gelu) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !a_planes || !w_planes || !bias || !out) return fail(h, STLT_ERR_INVALID, "null argument");
  if (terms != 1 && terms != 3) return fail(h, STLT_ERR_INVALID, "terms must be 1 or 3");
  if (m_rows % 128 || n % 256 || k % 64) return fail(h, STLT_ERR_INVALID, "shape not tile aligned");
  return run_gemm(h, static_cast<cudaStream_t>(stream), a_planes, m_rows, m_rows, w_planes, n, k, bias,
                  out, terms, out_kind, gelu);
}

namespace {
int op_gemm_fused(void* handle, void* stream, int32_t epilogue, const void* a, int32_t m_rows, const void* w,
                  int32_t n, int32_t k, const float* bias, void* out, void* out_bf16, int32_t gelu,
                  const float* stats_in, const float* vec_a, const float* vec_b, float* stats_out, float eps,
                  int32_t prev_norm, int terms);
}  // namespace

int stlt_op_gemm_fused(void* handle, void* stream, int32_t epilogue, const void* a, int32_t m_rows, const void* w,
                       int32_t n, int32_t k, const float* bias, void* out, void* out_bf16, int32_t gelu,
                       const float* stats_in, const float* vec_a, const float* vec_b, float* stats_out, float eps,
                       int32_t prev_norm) {
  return op_gemm_fused(handle, stream, epilogue, a, m_rows, w, n, k, bias, out, out_bf16, gelu, stats_in, vec_a, vec_b,
                       stats_out, eps, prev_norm, 1);
}

int stlt_op_gemm_fused_split(void* handle, void* stream, int32_t epilogue, const void* a_planes, int32_t m_rows,
                             const void* w_planes, int32_t n, int32_t k, const float* bias, void* out,
                             void* out_bf16_planes, int32_t gelu, const float* stats_in, const float* vec_a,
                             const float* vec_b, float* stats_out, float eps, int32_t prev_norm) {
  return op_gemm_fused(handle, stream, epilogue, a_planes, m_rows, w_planes, n, k, bias, out, out_bf16_planes, gelu, stats_in,
                       vec_a, vec_b, stats_out, eps, prev_norm, 3);
}

namespace {
int op_gemm_fused(void* handle, void* stream, int32_t epilogue, const void* a, int32_t m_rows, const void* w,
                  int32_t n, int32_t k, const float* bias, void* out, void* out_bf16, int32_t gelu,
                  const float* stats_in, const float* vec_a, const float* vec_b, float* stats_out, float eps,
                  int32_t prev_norm, int terms) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !a || !w || !out) return fail(h, STLT_ERR_INVALID, "null argument");
  if (m_rows % 128 || n % 256 || k % 64 || m_rows < 128) return fail(h, STLT_ERR_INVALID, "shape not tile aligned");
  EpiArgs e{};
  e.stats_in = reinterpret_cast<const float2*>(stats_in);
  e.vec_a = vec_a;
  e.vec_b = vec_b;
  e.eps = eps;
  if (epilogue == GEMM_EPI_NORM_A) {
    if (!stats_in || !vec_a || !vec_b || (gelu != 0 && gelu != (terms == 3 ? 1 : 2)))
      return fail(h, STLT_ERR_INVALID, "NORM_A: stats_in, vec_a (s), vec_b (c) required; gelu 0 or 2 (split operands: 0 or 1)");
    e.prev_norm = 1;
    return run_gemm_fused(h, static_cast<cudaStream_t>(stream), GEMM_EPI_NORM_A, a, m_rows, w, n, k, nullptr, out,
                          nullptr, gelu, e, nullptr, terms);
  }
  if (epilogue == GEMM_EPI_RESID) {
    if (n != kHidden || !bias || !out_bf16 || !stats_out || gelu != 0 || (prev_norm && (!stats_in || !vec_a || !vec_b)))
      return fail(h, STLT_ERR_INVALID, "RESID: n = 768, bias, out_bf16, stats_out (and LayerNorm inputs when prev_norm)");
    e.z_prev = static_cast<const float*>(out);
    e.stats_out = reinterpret_cast<float2*>(stats_out);
    e.zb_out = static_cast<__nv_bfloat16*>(out_bf16);
    e.prev_norm = prev_norm != 0 ? 1 : 0;
    return run_gemm_fused(h, static_cast<cudaStream_t>(stream), GEMM_EPI_RESID, a, m_rows, w, n, k, bias, out, out_bf16,
                          0, e, nullptr, terms);
  }
  return fail(h, STLT_ERR_INVALID, "epilogue must be 1 (NORM_A) or 2 (RESID)");
}
}  // namespace

int stlt_op_gemm_resid_hilo(void* handle, void* stream, const void* a, int32_t m_rows, const void* w, int32_t k,
                            const float* bias, void* z_hi, void* z_lo, const float* stats_in, const float* gamma,
                            const float* beta, float* stats_out, float eps, int32_t prev_norm) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !a || !w || !bias || !z_hi || !z_lo || !stats_out) return fail(h, STLT_ERR_INVALID, "null argument");
  if (m_rows % 128 || k % 64 || m_rows < 128) return fail(h, STLT_ERR_INVALID, "shape not tile aligned");
  if (prev_norm && (!stats_in || !gamma || !beta)) return fail(h, STLT_ERR_INVALID, "prev_norm needs stats_in, gamma, beta");
  EpiArgs e{};
  e.stats_in = reinterpret_cast<const float2*>(stats_in);
  e.vec_a = gamma;
  e.vec_b = beta;
  e.stats_out = reinterpret_cast<float2*>(stats_out);
  e.zb_out = static_cast<__nv_bfloat16*>(z_hi);
  e.z_lo = static_cast<__nv_bfloat16*>(z_lo);
  e.eps = eps;
  e.prev_norm = prev_norm != 0 ? 1 : 0;
  return run_gemm_fused(h, static_cast<cudaStream_t>(stream), GEMM_EPI_RESID, a, m_rows, w, kHidden, k, bias, nullptr,
                        z_hi, 0, e);
}

int stlt_op_pack_folded(void* handle, void* stream, const float* w, const float* gamma, const float* beta,
                        const float* bias, int32_t n, int32_t k, void* w_folded, float* s_out, float* c_out,
                        int32_t head_major) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !w || !bias || !w_folded || !s_out || !c_out) return fail(h, STLT_ERR_INVALID, "null argument");
  if ((gamma == nullptr) != (beta == nullptr)) return fail(h, STLT_ERR_INVALID, "gamma and beta go together");
  // head_major: bit 0 = head-major row order, bit 1 = hi / lo split planes (fp32-parity mode)
  STLT_CUDA(h, launch_pack_folded(w, gamma, beta, bias, n, k, static_cast<__nv_bfloat16*>(w_folded), s_out, c_out,
                                  static_cast<cudaStream_t>(stream), (head_major & 1) != 0, (head_major & 2) != 0));
  return STLT_OK;
}

int stlt_op_qkv_attention(void* handle, void* stream, const void* a, int64_t m_rows, int64_t valid_rows,
                          const void* w_head_major, const float* vec_s, const float* vec_c, const float* stats,
                          float eps, const int64_t* mask_src, int64_t num_seqs, int32_t seq_len, int32_t causal,
                          void* ctx) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !a || !w_head_major || !vec_s || !vec_c || !mask_src || !ctx) return fail(h, STLT_ERR_INVALID, "null argument");
  if (m_rows % 128 || m_rows < 128 || valid_rows > m_rows || num_seqs * seq_len > valid_rows)
    return fail(h, STLT_ERR_INVALID, "shape: m_rows must be a positive multiple of 128 covering the sequences");
  return run_qkv_attention(h, static_cast<cudaStream_t>(stream), a, m_rows, valid_rows,
                           static_cast<const __nv_bfloat16*>(w_head_major), vec_s, vec_c,
                           reinterpret_cast<const float2*>(stats), eps, reinterpret_cast<const long long*>(mask_src),
                           num_seqs, seq_len, causal != 0, ctx);
}

int stlt_op_gemm_grad(void* handle, void* stream, int32_t layout, const void* a, const void* b, void* out,
                      int32_t m_rows, int32_t n, int64_t k, int32_t out_kind) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h || !a || !b || !out) return fail(h, STLT_ERR_INVALID, "null argument");
  if (m_rows % 128 || n % 256 || k < 1) return fail(h, STLT_ERR_INVALID, "shape not tile aligned");
  if (layout == GEMM_NN && k % 64) return fail(h, STLT_ERR_INVALID, "k must be a multiple of 64");
  return run_gemm_grad(h, static_cast<cudaStream_t>(stream), layout, a, b, out, m_rows, n, k, out_kind);
}

int stlt_op_gemm_simt(void* handle, void* stream, const float* a, const float* w, const float* bias,
                      float* out, int32_t m, int32_t n, int32_t k, int32_t gelu) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  STLT_CUDA(h, launch_gemm_simt(a, w, bias, out, m, n, k, gelu != 0, static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

int stlt_op_attention(void* handle, void* stream, const void* qkv, int32_t qkv_is_bf16,
                      const int64_t* mask_src, int64_t num_seqs, int32_t seq_len, int32_t causal,
                      void* out_bf16, int32_t planes, int64_t plane_rows) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  ActOut o{nullptr, static_cast<__nv_bfloat16*>(out_bf16), planes, plane_rows};
  STLT_CUDA(h, launch_attention(qkv, qkv_is_bf16 != 0, reinterpret_cast<const long long*>(mask_src),
                                num_seqs, seq_len, causal != 0, o, static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

int stlt_op_add_ln(void* handle, void* stream, const float* x, const float* y, const float* gamma,
                   const float* beta, float eps, int64_t rows, float* out_f32, void* out_bf16,
                   int32_t planes, int64_t plane_rows) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  ActOut o{out_f32, static_cast<__nv_bfloat16*>(out_bf16), planes, plane_rows};
  STLT_CUDA(h, launch_add_ln(x, y, gamma, beta, eps, rows, o, static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

int stlt_op_pack_bf16(void* handle, void* stream, const float* src, void* dst, int64_t n,
                      int32_t planes) {
  Handle* h = static_cast<Handle*>(handle);
  if (!h) return fail(h, STLT_ERR_INVALID, "null handle");
  STLT_CUDA(h, launch_pack_bf16(src, static_cast<__nv_bfloat16*>(dst), n, planes,
                                static_cast<cudaStream_t>(stream)));
  return STLT_OK;
}

}  // extern "C"
