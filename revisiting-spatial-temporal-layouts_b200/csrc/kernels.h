// Internal launcher declarations shared by the translation units of libstlt_b200.so.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace stlt {

enum GemmOutKind : int {
  GEMM_OUT_F32 = 0,         // fp32 [M, N]
  GEMM_OUT_BF16 = 1,        // bf16 [M, N]
  GEMM_OUT_BF16_SPLIT = 2,  // bf16 hi plane [M, N] followed by lo plane [M, N]
  GEMM_OUT_BF16_DUAL = 3,   // bf16 act(z) [M, N] followed by bf16 act'(z) [M, N], both under the dropout mask (training)
  GEMM_OUT_F32_BF16 = 4,    // fp32 [M, N] (tm_out) and a bf16 copy [M, N] (tm_out2): fused-LayerNorm residual epilogue
  // same outputs, the bf16 copy written straight from registers (ep.zb_out): one staging tile less per warp buys
  // a fifth pipeline stage, which pays for long reductions (linear2, K = 3072)
  GEMM_OUT_F32_BF16_DIRECT = 5,
  // fused-LayerNorm residual epilogue on a residual stream stored ONCE as two bf16 planes, z = hi + lo (hi = bf16(z) is
  // also the next GEMM's A operand, lo = bf16(z - hi); ~2^-18 relative): the epilogue reads and writes 2 x 2 bytes per
  // element instead of 4 + 4 + 2 (fp32 in, fp32 out, bf16 copy), straight from / to registers (no staging tiles)
  GEMM_OUT_HILO = 6,
};

// LayerNorm fused into the projection GEMMs (bf16 inference path, see DESIGN.md "fused LayerNorm"): the
// residual stream is kept PRE-norm (z, fp32 + a bf16 copy) together with per-row (sum, sum of squares).
enum GemmEpilogue : int {
  GEMM_EPI_PLAIN = 0,
  // A = bf16(z) un-normalised, W pre-multiplied by gamma: out = act(rstd * (acc - mean * s[n]) + c[n]),
  // s[n] = sum_k (W gamma)[n,k], c[n] = (W beta)[n] + bias[n]  ==  act(LN(z) W^T + bias)
  GEMM_EPI_NORM_A = 1,
  // z_new = (prev_norm ? LN(z_prev) : z_prev) + acc + bias[n]; writes z_new (fp32 + bf16) and adds the row's
  // (sum, sum of squares) over this CTA's columns to stats_out. N must be 768 (the LayerNorm width).
  GEMM_EPI_RESID = 2,
  // training, data gradient of linear2 (GEMM_NN, bf16 output): out = acc * act_in, act_in = gelu'(u) * dropout mask
  // as stored by the forward FFN1 epilogue (second plane of GEMM_OUT_BF16_DUAL), i.e. the gradient w.r.t. the GELU
  // input, and colsum_out[n] += sum over the valid rows of the rounded output (the bias gradient of linear1).
  // Replaces a separate elementwise pass over the [M, 3072] gradient.
  GEMM_EPI_ACT_BWD = 3,
};

constexpr int kStatSlots = 6;  // 768 columns / 128-column slabs: partial row statistics per slab

struct EpiArgs {
  const float2* stats_in;  // [M][kStatSlots]; NORM_A: stats of the A rows; RESID: of the z_prev rows (prev_norm only)
  const float* vec_a;      // NORM_A: s[N]; RESID: gamma of the LayerNorm applied to z_prev
  const float* vec_b;      // NORM_A: c[N]; RESID: beta of that LayerNorm
  const float* z_prev;     // RESID: fp32 [M, 768]; may alias the fp32 output (in place)
  float2* stats_out;       // RESID: [M][kStatSlots], every slot is written (no zeroing needed)
  __nv_bfloat16* zb_out;   // RESID: bf16 copy of the new z [M, 768]
  float eps;
  int prev_norm;
  const __nv_bfloat16* act_in;  // ACT_BWD: gelu'(u) * dropout mask, bf16 [M, N]
  float* colsum_out;            // ACT_BWD: [N] fp32, accumulated with atomics (may be null)
  int valid_rows;               // ACT_BWD: rows >= valid_rows are tile padding and stay out of the column sums
  const int* m_tiles_dyn;       // device int (may be null): live 128-row tiles of A / out when the row count is dynamic
  __nv_bfloat16* z_lo;          // GEMM_OUT_HILO: lo plane of the residual stream [M, 768], updated in place (hi = zb_out)
  __nv_bfloat16* zb_lo_out;     // RESID, GEMM_OUT_F32_BF16_DIRECT: lo plane of the bf16 copy (fp32-parity mode: the next GEMM's
                                // A operand is the hi / lo split of z); null = one plane
};

enum GemmLayout : int {
  GEMM_NT = 0,      // out[M,N]  = A[M,K] * B[N,K]^T + bias   (forward projections)
  GEMM_NN = 1,      // out[M,N]  = A[M,K] * B[K,N]            (data gradients)
  GEMM_TN_RED = 2,  // out[M,N] += A[K,M]^T * B[K,N]          (weight gradients, stream-K + reduce-add)
};

struct GemmArgs {
  CUtensorMap tm_a;    // bf16 [planes * a_plane_rows, K], box {64, 128}, SWIZZLE_128B
  CUtensorMap tm_b;    // bf16 [planes * b_plane_rows, K], box {64, 128} (half a W tile), SWIZZLE_128B
  CUtensorMap tm_out;  // fp32 box {32, 32} or bf16 box {64, 32}, SWIZZLE_128B
  const float* bias;   // [N]
  int m_rows;          // multiple of 128
  int n;               // multiple of 256
  int k;               // multiple of 64
  int terms;           // 1 (bf16) or 3 (fp32-parity split)
  int out_kind;        // GemmOutKind
  int gelu;            // 0 none, 1 exact erf GELU (erff), 2 fast erf GELU (|erf err| < 7e-7), 3 ReLU
  int a_plane_rows;    // row offset of the lo plane of A
  int b_plane_rows;    // row offset of the lo plane of W
  int out_plane_rows;  // row offset of the lo plane of the output (split only)
  int layout;          // GemmLayout; MN-major operands use box {64, 64} tensor maps
  DropCfg drop;        // dropout on the activated output (GEMM_OUT_BF16_DUAL only; thr16 = 0: off)
  int epilogue;        // GemmEpilogue
  EpiArgs epi;
  CUtensorMap tm_out2; // bf16 [M, N] box {64, 32} (GEMM_OUT_F32_BF16 only)
};

int gemm_smem_bytes();
cudaError_t launch_gemm_tcgen05(const GemmArgs& g, cudaStream_t stream, int num_sms);

// In-projection + masked self-attention in one kernel (gemm_qkv_attn.cu; bf16 inference path, sequences of at
// most 32 tokens). The weights are the HEAD-MAJOR gamma-folded pack of launch_pack_folded(head_major = true).
// ---- pad-skipping row layout of the spatial phase (compact.cu) ----
// header of dynamic counts written by launch_compact_plan (device int[8])
enum CompactHeader : int {
  kDynFull = 0,          // frames with at least one non-padding object slot (sequences of S tokens)
  kDynSingle = 1,        // live frames whose slots 1.. are all padding (one-token sequences)
  kDynFullBlocks = 2,    // row blocks of the full frames (floor(128 / S) frames each)
  kDynSingleRow0 = 3,    // first row of the single-token frames = kDynFullBlocks * rows_per_block
  kDynRows = 4,          // kDynSingleRow0 + kDynSingle
  kDynTiles = 5,         // ceil(kDynRows / 128): live tiles of the row-wise GEMMs
  kDynSingleBlocks = 6,  // ceil(kDynSingle / 128)
  kDynAttnBlocks = 7,    // kDynFullBlocks + kDynSingleBlocks: row blocks the attention-fused in-projection executes
};
constexpr int kSingleFrameFlag = 0x40000000;  // frame_row[f]: row of slot 0 | flag when the frame is a single token; -1 dead
size_t compact_plan_scratch_bytes(long long frames);
long long compact_rows_bound(long long frames, int S);  // static upper bound of kDynRows (allocation)
cudaError_t launch_compact_plan(const long long* categories, const long long* lengths, int B, int L, int S,
                                int* frame_row, int* hdr, void* scratch, int* err_flag, cudaStream_t stream);
cudaError_t launch_gather_frames(const float* src_x, const __nv_bfloat16* src_att, const int* frame_row,
                                 long long frames, float* dst_x, __nv_bfloat16* dst_att, const float2* src_stats,
                                 float2* dst_stats, cudaStream_t stream, const __nv_bfloat16* src_hi = nullptr,
                                 const __nv_bfloat16* src_lo = nullptr,  // src_hi / src_lo: the source stream as bf16 planes
                                 int att_planes = 1, long long src_plane_rows = 0, long long dst_plane_rows = 0);

struct QkvAttnArgs {
  const float* vec_s;         // [2304] head-major: row sums of the folded weights (read when prev_norm)
  const float* vec_c;         // [2304] head-major: (W beta)[n] + bias[n]
  const float2* stats_in;     // [m_rows][kStatSlots] partial (sum, sum of squares) of the input rows (prev_norm)
  const long long* mask_src;  // [valid_rows]: key j is masked when mask_src[j] == 0
  long long m_rows;           // allocated rows of the activation / statistics / context tensors
  long long valid_rows;       // real tokens
  float eps;
  int prev_norm;       // 1: the input rows still need LayerNorm (applied algebraically in the epilogue)
  int seq_len;         // T <= 32
  int rows_per_block;  // floor(128 / T) * T (qkv_attention_rows_per_block)
  int row_blocks;      // ceil(sequences / floor(128 / T)); with `dyn`: static upper bound of full + single blocks
  int causal;
  const int* dyn;      // CompactHeader (device) or null: rows [0, dyn[kDynFull] * T) hold sequences of T tokens in blocks
                       // of rows_per_block, rows from dyn[kDynSingleRow0] hold dyn[kDynSingle] one-token sequences
  int debug;           // timing decomposition only (tools/bench_qkv_attention.py): 1 skip the attention math, 2 skip the
                       // Q/K/V tile stores, 4 skip the statistics loads, 8 sleep 2 us per unit in the epilogue; results are
                       // garbage when bits 1 / 2 / 4 are set
};
int qkv_attention_rows_per_block(int seq_len);
// tm_a: bf16 [m_rows, 768] box {64, 128}; tm_b: bf16 [2304, 768] box {64, 96}; tm_out: bf16 [m_rows, 768] box
// {64, rows_per_block}; tm_out1: same tensor, box {64, 128} (blocks of single-token sequences; = tm_out without
// `dyn`); all SWIZZLE_128B.
cudaError_t launch_qkv_attention(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const CUtensorMap& tm_out,
                                 const CUtensorMap& tm_out1, const QkvAttnArgs& p, cudaStream_t stream, int num_sms);

// fp32 SIMT GEMM: out[M, N] = act(A[M, K] * W[N, K]^T + bias). Used by the classifier head
// (src/modelling/models.py:155-163) and as an independent cross-check of the tcgen05 path in tests.
cudaError_t launch_gemm_simt(const float* a, const float* w, const float* bias, float* out, int m,
                             int n, int k, bool gelu, cudaStream_t stream);

// Activation outputs: fp32 stream + bf16 plane(s) used as the next GEMM's A operand.
struct ActOut {
  float* x;               // fp32 [rows, 768] (may be null)
  __nv_bfloat16* xb;      // bf16 [planes, plane_rows, 768]
  int planes;             // 1 or 2
  long long plane_rows;   // rows per plane (M padded to 128)
};

// K0: box repair + normalisation + padding masks (src/utils/data_utils.py:205-231,
// src/modelling/datasets.py:78-83,247-286).
cudaError_t launch_prepare(const double* raw_boxes, const long long* video_sizes,
                           const long long* categories, const long long* frame_types, int B, int L,
                           int S, float* boxes_out, uint8_t* mask_boxes, uint8_t* mask_frames,
                           cudaStream_t stream);

cudaError_t launch_masks(const long long* categories, const long long* frame_types,
                         long long n_slots, long long n_frames, uint8_t* mask_boxes,
                         uint8_t* mask_frames, cudaStream_t stream);

// CSR layout store of a dataset on the device (see include/stlt_b200.h: StltLayoutStore).
struct BatchStore {
  const long long* video_frame_offsets;   // [V + 1]
  const long long* frame_object_offsets;  // [F + 1]
  const double* obj_boxes;                // [O, 4] raw pixel boxes
  const long long* obj_categories;        // [O] ids (category2id already applied)
  const double* obj_scores;               // [O]
  const long long* video_sizes;           // [V, 2] (width, height)
};
struct BatchIds {
  long long cls, ft_pad, ft_regular, ft_empty, ft_extract;
};
cudaError_t launch_build_batch(const BatchStore& st, const long long* video_index,
                               const long long* frame_indices, const long long* num_sampled, int B,
                               int T, int L, int S, double score_threshold, const BatchIds& ids,
                               long long* categories, float* boxes, float* scores,
                               long long* frame_types, long long* lengths, uint8_t* mask_boxes,
                               uint8_t* mask_frames, int* err_flag, cudaStream_t stream);
cudaError_t launch_topk_count(const float* logits, const long long* labels, int rows, int classes,
                              unsigned long long* counters, cudaStream_t stream);

// Multi-label evaluator (src/utils/evaluation.py:61-132): sigmoid + label accumulation, Charades mAP.
cudaError_t launch_map_accumulate(const float* logits, const float* labels, long long n, float* pred, float* gt,
                                  cudaStream_t stream);
cudaError_t launch_charades_map(const float* pred, const float* gt, int n, int classes, double* ap_out,
                                double* map_out, cudaStream_t stream);

// K1: category + box (+score) embedding and LayerNorm (src/modelling/models.py:29-39).
cudaError_t launch_embed(const long long* categories, const float* boxes, const float* scores,
                         const float* cat_table, int unique_categories, const float* box_w,
                         const float* box_b, const float* score_w, const float* score_b,
                         const float* ln_g, const float* ln_b, float eps, long long tokens,
                         ActOut out, int* err_flag, cudaStream_t stream, float* scratch, size_t scratch_bytes,
                         DropCfg drop = DropCfg{0, 0, 1.f}, const int* frame_row = nullptr, int S = 1,
                         long long* mask_out = nullptr);
// frame_row != null (pad-skipping layout): token (f, s) of the padded grid goes to row frame_row[f] + s (full frame), to
// row frame_row[f] & ~kSingleFrameFlag if s == 0 (single-token frame) or nowhere; mask_out[row] receives its category.
// `scratch` (>= embed_scratch_bytes, any buffer that is dead until the kernel after the embedding) receives the
// per-category LayerNorm statistics tables that embed_stats_kernel derives from the live parameters on every call.
size_t embed_scratch_bytes(int unique_categories);

// x <- LayerNorm(x + y) (post-norm residual of nn.TransformerEncoderLayer); y may be null.
// z_out (optional, training): receives the pre-LayerNorm sum x + y.
cudaError_t launch_add_ln(const float* x_in, const float* y, const float* g, const float* b,
                          float eps, long long rows, ActOut out, cudaStream_t stream,
                          float* z_out = nullptr, DropCfg drop = DropCfg{0, 0, 1.f}, const int* rows_dyn = nullptr);
// rows_dyn (device int, may be null): live row count when it is decided on the device (rows is then the static bound)

// Same with the branch output y in bf16 (bf16 inference mode: the out-projection / linear2 GEMMs store
// bf16, which halves their HBM write and this kernel's y read).
cudaError_t launch_add_ln_bf16y(const float* x_in, const __nv_bfloat16* y, const float* g, const float* b,
                                float eps, long long rows, ActOut out, cudaStream_t stream,
                                float* z_out = nullptr, DropCfg drop = DropCfg{0, 0, 1.f}, const int* rows_dyn = nullptr);

// K7: frame tokens = LN(spatial CLS slot + position + frame type) (src/modelling/models.py:98-111).
cudaError_t launch_frame_embed(const float* spatial_x, int S, const long long* frame_types,
                               const float* pos_table, const float* ft_table, int n_frame_types,
                               const float* ln_g, const float* ln_b, float eps, int B, int L,
                               ActOut out, int* err_flag, cudaStream_t stream,
                               DropCfg drop = DropCfg{0, 0, 1.f}, const float* pre_g = nullptr,
                               const float* pre_b = nullptr, float pre_eps = 0.f);

// K9: h[b] = x[b * L + lengths[b] - 1] (src/modelling/models.py:189-192).
cudaError_t launch_gather_last(const float* x, const long long* lengths, int B, int L, float* out,
                               int* err_flag, cudaStream_t stream);

// Row compaction for the pruned last layer of each stack (see elementwise.cu).
cudaError_t launch_gather_rows(const float* src_x, const __nv_bfloat16* src_att, int planes,
                               long long src_plane_rows, int stride, const long long* lengths, int L,
                               long long rows, float* dst_x, __nv_bfloat16* dst_att,
                               long long dst_plane_rows, int* err_flag, cudaStream_t stream,
                               const float2* src_stats = nullptr, float2* dst_stats = nullptr,
                               const __nv_bfloat16* src_hi = nullptr, const __nv_bfloat16* src_lo = nullptr);

// K3: masked multi-head attention over short sequences held in shared memory.
//   qkv: [tokens, 2304] fp32 (fp32-parity mode; bf16 input is handled by launch_attention_mma); key j of a sequence is masked when mask_src[token_j] == 0,
//   and (causal) when j > i. Output: bf16 plane(s) [tokens, 768].
cudaError_t launch_attention(const void* qkv, bool qkv_is_bf16, const long long* mask_src,
                             long long num_seqs, int T, bool causal, ActOut out,
                             cudaStream_t stream);

// K3 on warp-level tensor-core tiles (attention_mma.cu). planes = 1: bf16 QKV [tokens, 2304] ->
// bf16 context; planes = 2: hi/lo bf16 planes in and out (fp32-parity mode, 3-term split products).
// drop (training, planes = 1 only): dropout on the attention probabilities; element index =
// ((query token * 12 + head) * 32 + key position within the sequence).
cudaError_t launch_attention_mma(const __nv_bfloat16* qkv, int planes, long long qkv_plane_rows,
                                 const long long* mask_src, long long num_seqs, int T, bool causal,
                                 __nv_bfloat16* out, long long out_plane_rows, cudaStream_t stream,
                                 DropCfg drop = DropCfg{0, 0, 1.f}, const int* dyn = nullptr, int dyn_region = 0);
// dyn (CompactHeader, pad-skipping layout): region 0 = dyn[kDynFull] sequences of T tokens from row 0 (num_seqs is then only the
// static bound that sizes the grid), region 1 = dyn[kDynSingle] sequences from row dyn[kDynSingleRow0] (call with T = 1).

// K3 for 65..256-token sequences (attention_long.cu): one CTA per (sequence, head), Q / K / V staged once in shared
// memory, online softmax over 64-key blocks. planes as in launch_attention_mma; inference only.
cudaError_t launch_attention_long(const __nv_bfloat16* qkv, int planes, long long qkv_plane_rows,
                                  const long long* mask_src, long long num_seqs, int T, bool causal,
                                  __nv_bfloat16* out, long long out_plane_rows, cudaStream_t stream);

// gamma-folded bf16 weights + the two epilogue vectors of GEMM_EPI_NORM_A (see GemmEpilogue).
// gamma / beta may be null (identity LayerNorm: plain bf16 weights, s = row sums, c = bias). head_major (n must be
// 2304): output row h*192 + t*64 + j holds row t*768 + h*64 + j of the packed in-projection (t = Q, K, V).
// split (fp32-parity mode): wf holds two planes n rows apart, hi = bf16(w'), lo = bf16(w' - hi); s sums hi + lo.
cudaError_t launch_pack_folded(const float* w, const float* gamma, const float* beta, const float* bias, int n,
                               int k, __nv_bfloat16* wf, float* s_out, float* c_out, cudaStream_t stream,
                               bool head_major = false, bool split = false);

// fp32 -> bf16 plane(s) for weights.
cudaError_t launch_pack_bf16(const float* src, __nv_bfloat16* dst, long long n, int planes,
                             cudaStream_t stream);

// ---- training step; see train_kernels.cu / attention_bwd.cu -------------

// Backward of x_out = LN(z): dz (fp32 and/or bf16), d gamma, d beta and colsum(dz) (bias gradient of
// the linear that produced the residual branch). d_b (second addend of the incoming gradient) and
// every output may be null.
cudaError_t launch_ln_bwd(const float* d_a, const float* d_b, const float* z, const float* gamma,
                          float eps, long long rows, float* dz_out, __nv_bfloat16* dzb_out,
                          float* d_gamma, float* d_beta, float* d_bias, cudaStream_t stream,
                          DropCfg drop = DropCfg{0, 0, 1.f});
cudaError_t launch_colsum_f32(const float* x, int rows, int n, float* out, cudaStream_t stream);
// Adjoint of launch_gather_rows: dst_f[map(r)] += src_f[r]; dst_b[map(r)] = src_b[r].
cudaError_t launch_scatter_rows(const float* src_f, float* dst_f, const __nv_bfloat16* src_b,
                                __nv_bfloat16* dst_b, int stride, const long long* lengths, int L,
                                long long rows, cudaStream_t stream);
cudaError_t launch_frame_embed_bwd(const float* d_a, const float* d_b, const float* cls_x,
                                   const long long* frame_types, const float* pos_table,
                                   const float* ft_table, int n_frame_types, const float* gamma,
                                   float eps, int B, int L, float* d_cls, float* d_pos, float* d_ft,
                                   float* d_gamma, float* d_beta, cudaStream_t stream,
                                   DropCfg drop = DropCfg{0, 0, 1.f});
cudaError_t launch_embed_bwd(const float* d_a, const float* d_b, const long long* categories,
                             const float* boxes, const float* scores, const float* cat_table,
                             int unique_categories, const float* box_w, const float* box_b,
                             const float* score_w, const float* score_b, const float* gamma, float eps,
                             long long tokens, float* d_pre, float* d_cat, float* d_box_w,
                             float* d_box_b, float* d_score_w, float* d_score_b, float* d_gamma,
                             float* d_beta, cudaStream_t stream, DropCfg drop = DropCfg{0, 0, 1.f});
cudaError_t launch_gelu_ln(const float* h1, const float* g, const float* b, float eps, long long rows,
                           float* out, cudaStream_t stream);
cudaError_t launch_gelu_ln_bwd(const float* d_h2, const float* h1, const float* gamma, float eps,
                               long long rows, float* d_h1, float* d_gamma, float* d_beta,
                               cudaStream_t stream);
// out[i, j] (+)= sum_k A[i*sai + k*sak] * B[k*sbk + j*sbj], fp32 CUDA cores (classifier-head gradients; gemm_simt.cu).
// B must be row-major (sbj == 1) and A contiguous along i or k; accumulate = true adds into `out` with atomics and
// splits the reduction over the grid (weight gradients: few output tiles, the batch is the reduction).
cudaError_t launch_gemm_strided(const float* a, long long sai, long long sak, const float* b,
                                long long sbk, long long sbj, float* out, int m, int n, int k,
                                bool accumulate, cudaStream_t stream);
cudaError_t launch_cross_entropy(const float* logits, const long long* labels, int rows, int classes,
                                 float grad_scale, float* loss, float* d_logits, cudaStream_t stream);
cudaError_t launch_bce_logits(const float* logits, const float* targets, long long n, float grad_scale,
                              float* loss, float* d_logits, cudaStream_t stream);
// out[0] = sum(g^2), deterministic (fixed reduction order); scratch: >= 1 floats, up to 1184 are used
cudaError_t launch_sumsq(const float* g, long long n, float* out, float* scratch, int scratch_len,
                         cudaStream_t stream);
cudaError_t launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float bias_c1,
                         float bias_c2_sqrt, const float* sumsq, float max_norm, cudaStream_t stream);
cudaError_t launch_attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* d_ctx,
                                 const long long* mask_src, long long num_seqs, int T, bool causal,
                                 __nv_bfloat16* d_qkv, cudaStream_t stream, DropCfg drop = DropCfg{0, 0, 1.f});

// ---- CACNF fusion path; attention_cross.cu / cacnf_kernels.cu ------------

// Attention between two streams (<= 64 queries / keys per sequence). q: [*, ldq] bf16, Q at column
// q_off + 64*head; kv: [*, ldkv] bf16 with K / V at k_off / v_off + 64*head. mask_src (i64 per key
// token, may be null): key masked when 0. out: bf16 [num_seqs * Tq, 768].
cudaError_t launch_attention_cross(const __nv_bfloat16* q, int ldq, int q_off, const __nv_bfloat16* kv,
                                   int ldkv, int k_off, int v_off, const long long* mask_src,
                                   long long num_seqs, int Tq, int Tk, bool causal, __nv_bfloat16* out,
                                   cudaStream_t stream, int planes = 1, long long q_plane_rows = 0,
                                   long long kv_plane_rows = 0, long long out_plane_rows = 0);
// features f32 [B, C, P] (channel-major, as the 3D ResNet emits them) -> bf16 tokens [B * P, C]
// lo_plane_rows != 0 (fp32-parity mode): also writes the bf16 residual plane that many rows further
cudaError_t launch_features_to_tokens(const float* feat, __nv_bfloat16* out, int B, int C, int P,
                                      cudaStream_t stream, long long lo_plane_rows = 0);
// x[b, 0] = cls + pos[0]; x[b, 1 + s] = proj[b, s] + pos[1 + s]   (models.py:262-270)
cudaError_t launch_app_embed(const float* proj, const float* cls, const float* pos, int B, int P, ActOut out,
                             cudaStream_t stream);
// dst[b, 0:768] = a[b * a_stride (+ lengths[b] - 1 if lengths)], dst[b, 768:1536] = c[b * c_stride]
cudaError_t launch_gather_concat(const float* a, int a_stride, const long long* lengths, const float* c,
                                 int c_stride, int B, float* dst, cudaStream_t stream);
cudaError_t launch_mean3(const float* a, const float* b, const float* c, long long n, float* out,
                         cudaStream_t stream);

// Same on mma.sync tensor-core tiles (attention_bwd_mma.cu); the CUDA-core version above is kept as a
// cross-check for the tests. d_bias (optional, f32 [2304]): += column sums of d_qkv (in-projection bias gradient).
cudaError_t launch_attention_bwd_mma(const __nv_bfloat16* qkv, const __nv_bfloat16* d_ctx,
                                     const long long* mask_src, long long num_seqs, int T, bool causal,
                                     __nv_bfloat16* d_qkv, cudaStream_t stream,
                                     DropCfg drop = DropCfg{0, 0, 1.f}, float* d_bias = nullptr);

}  // namespace stlt
