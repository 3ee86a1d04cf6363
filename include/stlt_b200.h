/*
 * stlt_b200.h — C ABI of libstlt_b200.so: the B200 (sm_100a) implementation of the STLT
 * layout-encoding forward path of gorjanradevski/revisiting-spatial-temporal-layouts.
 *
 * The reference has no FFI layer: its boundary is the Python nn.Module protocol of `Stlt`
 * (reference src/modelling/models.py:166-195) as called from src/train.py:125,142 and
 * src/inference.py:77. This header is what a binding for that path calls; the Python shim in
 * revisiting-spatial-temporal-layouts_b200/module.py binds it with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative STLT_ERR_* code otherwise; no exceptions
 *     cross the ABI. stlt_last_error() returns a human-readable message for the last failure.
 *   - all tensor pointers are DEVICE pointers, contiguous, row-major. The caller owns every byte
 *     of device memory (inputs, weights, packed weights, workspace, outputs); the library never
 *     allocates device memory and never synchronises, except stlt_check_errors().
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it and
 *     is CUDA-graph capturable.
 *   - a handle is not thread-safe; distinct handles are independent. One handle per device.
 *   - there is no CPU fallback: a machine without an sm_100 GPU gets STLT_ERR_CUDA.
 */
#ifndef STLT_B200_H_
#define STLT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STLT_OK 0
#define STLT_ERR_INVALID (-1) /* bad argument / shape / missing weight */
#define STLT_ERR_CUDA (-2)    /* CUDA runtime or driver error */
#define STLT_ERR_STATE (-3)   /* call order (weights not bound / packed) */
#define STLT_ERR_INPUT (-4)   /* out-of-range categories / frame_types / lengths (check_errors) */

/* precision of the projection GEMMs */
#define STLT_PRECISION_FP32 0 /* 3-term bf16 split on tcgen05; logits within 1e-4 of fp32 */
#define STLT_PRECISION_BF16 1 /* bf16 operands, fp32 accumulate/residual/LN/softmax */

/* dtype codes for StltTensor */
#define STLT_DTYPE_F32 0
#define STLT_DTYPE_I64 1

/* Mirrors the fields of the reference StltModelConfig (src/modelling/configs.py:92-111) that
 * reach the forward path. hidden_size/num_heads must be 768/12 (kernels are specialised). */
typedef struct StltDims {
  int32_t hidden_size;         /* configs.py:96  (768) */
  int32_t num_heads;           /* configs.py:99  (12) */
  int32_t num_spatial_layers;  /* configs.py:107 (4) */
  int32_t num_temporal_layers; /* configs.py:108 (8) */
  int32_t unique_categories;   /* configs.py:105 */
  int32_t num_classes;         /* configs.py:94 */
  int32_t max_positions;       /* configs.py:109 layout_num_frames (256) */
  int32_t num_frame_types;     /* models.py:91 (5) */
  float layer_norm_eps;        /* configs.py:98 (1e-12): embeddings, frames, head LayerNorms */
  float encoder_norm_eps;      /* 1e-5: LayerNorms inside nn.TransformerEncoderLayer */
} StltDims;

/* One named parameter tensor; `name` is the reference state_dict key. */
typedef struct StltTensor {
  const char* name;
  const void* data; /* device pointer */
  int32_t dtype;    /* STLT_DTYPE_* */
  int32_t ndim;
  int64_t shape[4];
} StltTensor;

/* Optional device buffers that receive the fp32 activations after each stage (parity tests).
 * Any pointer may be NULL. Shapes: embed/spatial [B*L*S, 768]; frames/temporal [B*L, 768];
 * pooled [B, 768]. */
typedef struct StltTaps {
  float* embed;    /* CategoryBoxEmbeddings output       (models.py:29-39) */
  float* spatial;  /* spatial encoder output, all slots  (models.py:68-71) */
  float* frames;   /* FramesEmbeddings output            (models.py:98-111) */
  float* temporal; /* temporal encoder output            (models.py:146-150) */
  float* pooled;   /* gathered extract-frame token       (models.py:189-192) */
} StltTaps;

/* Replaces Stlt.__init__ (models.py:167-178): validates the dimensions, creates host state. */
int stlt_create(const StltDims* dims, void** handle);
int stlt_destroy(void* handle);
const char* stlt_last_error(void* handle);

/* Replaces module.load_state_dict / .to(device) on the library side (inference.py:58-69): binds
 * device pointers of the fp32 parameters by their reference state_dict names. Pointers are
 * borrowed, not copied. Unknown names (the orphan encoder_layer.*, position_ids) are ignored. */
int stlt_bind_weights(void* handle, const StltTensor* tensors, int32_t count);

/* The tcgen05 GEMMs read bf16 copies of the 4 projection matrices of every encoder layer
 * (2 planes hi/lo for STLT_PRECISION_FP32, 1 plane for BF16). The caller provides the buffer and
 * re-runs stlt_pack_weights whenever the fp32 parameters change. */
int stlt_packed_weights_bytes(void* handle, int32_t precision, size_t* bytes);
int stlt_pack_weights(void* handle, void* stream, int32_t precision, void* packed, size_t bytes);

int stlt_workspace_bytes(void* handle, int32_t batch, int32_t frames, int32_t slots,
                         int32_t precision, size_t* bytes);

/* K0 — replaces fix_box (src/utils/data_utils.py:205-231), the division by the video size
 * (src/modelling/datasets.py:54,82) and the two padding masks of StltCollater
 * (src/modelling/datasets.py:274-286) for an already padded layout:
 *   raw_boxes   f64 [B, L, S, 4] pixel boxes (x1, y1, x2, y2) as read from the dataset JSON
 *   video_sizes i64 [B, 2] (width, height)
 *   categories  i64 [B, L, S], frame_types i64 [B, L]
 * Slot 0 of every frame receives the CLS box [0,0,1,1], slots with category 0 receive zeros.
 *   boxes_out f32 [B, L, S, 4]; mask_boxes_out u8 [B, L, S]; mask_frames_out u8 [B, L]. */
int stlt_prepare(void* handle, void* stream, const double* raw_boxes, const int64_t* video_sizes,
                 const int64_t* categories, const int64_t* frame_types, int32_t batch,
                 int32_t frames, int32_t slots, float* boxes_out, uint8_t* mask_boxes_out,
                 uint8_t* mask_frames_out);

/* A dataset's layouts in CSR form on the device (the reference keeps them as nested JSON lists,
 * src/modelling/datasets.py:35-37; schema written by src/create_something_datasets.py:18-34). */
typedef struct StltLayoutStore {
  const int64_t* video_frame_offsets;  /* [V + 1] frame range of every video */
  const int64_t* frame_object_offsets; /* [F + 1] object range of every frame */
  const double* obj_boxes;             /* [O, 4] raw pixel (x1, y1, x2, y2) */
  const int64_t* obj_categories;       /* [O] category ids (category2id applied, configs.py:30-78) */
  const double* obj_scores;            /* [O] detector scores */
  const int64_t* video_sizes;          /* [V, 2] (width, height) */
} StltLayoutStore;

/* Category / frame-type ids of the dataset (configs.py:30-88). */
typedef struct StltLayoutIds {
  int64_t cls, ft_pad, ft_regular, ft_empty, ft_extract;
} StltLayoutIds;

/* §8(f) rank 3 — replaces StltDataset.__getitem__ (datasets.py:52-125) + StltCollater.__call__
 * (datasets.py:243-288) for a batch of `batch` videos `video_index[b]` of the store: frame sampling
 * (test-time get_test_layout_indices, data_utils.py:47-56, unless explicit `frame_indices` i64
 * [batch, layout_num_frames] + `num_sampled` i64 [batch] are given), score filter, CLS slot, fix_box +
 * normalisation, extract frame, object / frame padding and both masks, in one kernel.
 * `frames` must be max_b(min(n_frames_b, layout_num_frames)) + 1 (what pad_sequence produces),
 * `slots` = max_num_objects + 1. Outputs as the reference collater: categories i64 [B,L,S], boxes f32
 * [B,L,S,4], scores f32 [B,L,S] (or NULL), frame_types i64 [B,L], lengths i64 [B], masks u8.
 * `status_out` (device int32) receives 0, or 4/5/6 for inconsistent frames / indices / slots. */
int stlt_build_batch(void* handle, void* stream, const StltLayoutStore* store, const StltLayoutIds* ids,
                     const int64_t* video_index, const int64_t* frame_indices_or_null,
                     const int64_t* num_sampled_or_null, int32_t batch, int32_t layout_num_frames,
                     int32_t frames, int32_t slots, double score_threshold, int64_t* categories_out,
                     float* boxes_out, float* scores_out_or_null, int64_t* frame_types_out,
                     int64_t* lengths_out, uint8_t* mask_boxes_out, uint8_t* mask_frames_out,
                     int32_t* status_out);

/* §8(f) rank 4 — replaces the per-batch `.cpu()` bookkeeping of EvaluatorSomething.process
 * (src/utils/evaluation.py:21-34): adds the number of top-1 and top-5 hits of logits f32 [rows,
 * classes] against labels i64 [rows] to counters[0] / counters[1] (device uint64[2]), no sync. */
int stlt_topk_count(void* handle, void* stream, const float* logits, const int64_t* labels, int32_t rows,
                    int32_t classes, uint64_t* counters);

/* §8(f) rank 4, multi-label head — replaces EvaluatorActionGenome.process (src/utils/evaluation.py:76-83):
 * predictions_out[r] = sigmoid(logits[r]), ground_truths_out[r] = labels[r] for `rows` rows (the caller passes
 * the write position inside its [total_instances, classes] device buffers), no synchronisation. */
int stlt_map_accumulate(void* handle, void* stream, const float* logits, const float* labels, int32_t rows,
                        int32_t classes, float* predictions_out, float* ground_truths_out);
/* Replaces charades_map / map (src/utils/evaluation.py:100-132) on the device: per-class average precision
 * (f64 [classes]; nan for a class without positives) and their mean (f64 [1], np.mean semantics). Rows whose
 * ground truth is all zero rank last (-inf). At most 32768 instances. */
int stlt_charades_map(void* handle, void* stream, const float* predictions, const float* ground_truths,
                      int32_t instances, int32_t classes, double* ap_out, double* map_out);

/* Replaces Stlt.forward (models.py:185-195) in eval mode.
 *   categories i64 [B, L, S]; boxes f32 [B, L, S, 4]; scores f32 [B, L, S] or NULL (presence
 *   toggles the score embedding, models.py:33-35); frame_types i64 [B, L]; lengths i64 [B].
 *   logits_out f32 [B, num_classes]. mask_*_out (u8, optional) receive the padding masks the
 *   reference collater would have produced (categories == 0, frame_types == 0).
 *   Shapes: B >= 0 (0 is a no-op), 1 <= L <= min(256, max_positions) (the position table, models.py:88-96),
 *   1 <= S <= 64; anything else returns STLT_ERR_INVALID. */
int stlt_forward(void* handle, void* stream, int32_t precision, const int64_t* categories,
                 const float* boxes, const float* scores_or_null, const int64_t* frame_types,
                 const int64_t* lengths, int32_t batch, int32_t frames, int32_t slots,
                 void* workspace, size_t workspace_bytes, float* logits_out,
                 uint8_t* mask_boxes_out_or_null, uint8_t* mask_frames_out_or_null);

/* Synchronises `stream` and reports STLT_ERR_INPUT if the last forward on `workspace` saw an
 * index outside its table (PyTorch raises IndexError for those on CPU). */
int stlt_check_errors(void* handle, void* stream, const void* workspace);

/* Number of kernels enqueued by the most recent stlt_forward on this handle. */
int stlt_last_launch_count(void* handle);

/* Per-category device timing of the launches issued by stlt_forward, measured with CUDA events on
 * the launching stream. stlt_set_profiling(1) starts collecting (and clears previous spans);
 * stlt_get_profile() synchronises on the recorded events, sums them per category and clears. */
#define STLT_PROF_GEMM 0      /* tcgen05 projection GEMMs */
#define STLT_PROF_ATTENTION 1 /* shared-memory attention */
#define STLT_PROF_ADD_LN 2    /* residual + LayerNorm */
#define STLT_PROF_OTHER 3     /* embeddings, gather, classifier head */
#define STLT_PROF_CATEGORIES 4
typedef struct StltProfile {
  double ms[STLT_PROF_CATEGORIES];
  double flops[STLT_PROF_CATEGORIES]; /* executed multiply-add FLOPs (2*M*N*K), GEMM only */
  int64_t launches[STLT_PROF_CATEGORIES];
} StltProfile;
int stlt_set_profiling(void* handle, int32_t enable);
int stlt_get_profile(void* handle, StltProfile* out);

/* The STLT_PROF_GEMM category of the most recent stlt_get_profile call, split by the role of the launch in the encoder layer
 * (models.py:46-55,118-128): the measurement behind a per-kernel roofline. QKV_ATTENTION is the in-projection with the
 * attention in its epilogue (bf16 path; its FLOPs are the in-projection's), GRADIENT the backward GEMMs of the training step. */
#define STLT_PROF_ROLE_IN_PROJ 0
#define STLT_PROF_ROLE_QKV_ATTENTION 1
#define STLT_PROF_ROLE_OUT_PROJ 2
#define STLT_PROF_ROLE_LINEAR1 3
#define STLT_PROF_ROLE_LINEAR2 4
#define STLT_PROF_ROLE_GRADIENT 5
#define STLT_PROF_ROLE_OTHER_GEMM 6
#define STLT_PROF_ROLES 7
typedef struct StltRoleProfile {
  double ms[STLT_PROF_ROLES];
  double flops[STLT_PROF_ROLES];
  int64_t launches[STLT_PROF_ROLES];
} StltRoleProfile;
int stlt_get_profile_by_role(void* handle, StltRoleProfile* out);

/* Last-layer pruning (default on): only slot 0 of the spatial stack's output (models.py:79) and
 * only frame lengths-1 of the temporal stack's output (models.py:192) are read, so the row-wise
 * half of the last layer of each stack (out-proj, FFN, LayerNorms) runs on those rows only. The
 * logits are bit-identical with pruning off; taps of the full stack outputs disable it. */
int stlt_set_pruning(void* handle, int32_t enable);

/* bf16 mode folds every encoder LayerNorm into the epilogues of the GEMMs around it (default on): the
 * residual stream is kept pre-norm with per-row statistics, the in-projection / linear1 GEMMs read it with
 * gamma-folded weights and normalise in their epilogue, the out-projection / linear2 GEMMs add the residual
 * and accumulate the next statistics. 0 restores the separate residual + LayerNorm kernels (same results
 * up to bf16 rounding; used by the tests as a cross-check). Has no effect with taps set.
 * stlt_set_fused_ln_fp32 is the same switch for the fp32-parity mode (default on): there the GEMMs run on hi / lo
 * split operands (3 bf16 MMAs per product), the folded weights are split the same way, the residual epilogues write
 * the split of the new stream as the next operand, and the attention stays a separate kernel. */
int stlt_set_fused_ln(void* handle, int32_t enable);
int stlt_set_fused_ln_fp32(void* handle, int32_t enable);

/* bf16 mode with fused LayerNorms also folds the attention into the epilogue of the in-projection GEMM (default
 * on; sequences of at most 32 tokens): F.linear(x, in_proj_weight, in_proj_bias) and the scaled-dot-product
 * attention of nn.MultiheadAttention (models.py:46-55,118-128) run as ONE kernel whose work unit is a
 * (row block, head) and whose weights are packed head-major, so the packed Q | K | V activations never reach
 * HBM. 0 restores the separate in-projection GEMM + attention kernel (cross-check for the tests). */
int stlt_set_fused_attention(void* handle, int32_t enable);

/* Optional (default OFF): keep the (pre-norm) residual stream of the two encoder stacks as two bf16 planes,
 * z = hi + lo with hi = bf16(z) — which is also the A operand of the next projection — and lo = bf16(z - hi)
 * (~2^-18 relative, far below the bf16 operand rounding): the residual epilogues then move 8 instead of 10 bytes
 * per element, but straight from registers in 16-byte pieces, which measured 9 % slower per step than the
 * TMA-staged fp32 stream + bf16 copy of the default path. Kept as a tested switch for that comparison. */
int stlt_set_hilo_residual(void* handle, int32_t enable);

/* ... and runs the spatial stack on a pad-skipping row layout (default on): frames at or after lengths[b] and
 * the padded slots of frames whose slots 1.. are all padding (the "extract" frame, datasets.py:97-113) can never
 * reach the logits (key-padding / causal masks, models.py:66-71,79,142-150,192) and are not computed. Which rows
 * exist is decided on the device from `lengths` / `categories`; allocation sizes do not change. 0 computes the
 * whole padded [B, L, S] grid as the reference does (cross-check for the tests). */
int stlt_set_compaction(void* handle, int32_t enable);

/* Test taps (NULL disables). The struct is copied. */
int stlt_set_taps(void* handle, const StltTaps* taps);

/* ---- training step (BASELINE.json configs[3]) -------------------------
 * Device side of the reference loop src/train.py:117-135: forward with saved activations,
 * criterion (src/utils/train_inference_utils.py:64-76), backward, clip_grad_norm_ (train.py:129)
 * and torch.optim.AdamW (train.py:102-104,130). bf16 GEMM operands, everything else fp32; the
 * weights must be packed for STLT_PRECISION_BF16. Gradients are ACCUMULATED into the buffers bound
 * with stlt_bind_grads (zero them first, like optimizer.zero_grad()); the weight-gradient GEMMs and
 * the column reductions add with atomics, so results are not bit-reproducible run to run. */

/* Binds fp32 gradient buffers by reference state_dict name (same shapes as the parameters). A
 * parameter without an entry gets no gradient (frozen backbone, models.py:172-176). */
int stlt_bind_grads(void* handle, const StltTensor* tensors, int32_t count);

int stlt_train_workspace_bytes(void* handle, int32_t batch, int32_t frames, int32_t slots, size_t* bytes);

/* Stlt.forward in train mode. The workspace keeps the activations stlt_backward needs and must not
 * be touched in between. dropout_p: the reference's hidden_dropout_prob (configs.py:97). */
int stlt_forward_train(void* handle, void* stream, const int64_t* categories, const float* boxes,
                       const float* scores_or_null, const int64_t* frame_types, const int64_t* lengths,
                       int32_t batch, int32_t frames, int32_t slots, void* workspace,
                       size_t workspace_bytes, float dropout_p, uint64_t seed, float* logits_out);

/* loss.backward() from d loss / d logits (f32 [batch, num_classes]). `phases` selects which half
 * runs, so a data-parallel caller can start the all-reduce of the first half's gradients while the
 * second half computes: STLT_BWD_TEMPORAL = classifier head + temporal stack + frame embedding,
 * STLT_BWD_SPATIAL = spatial stack + category/box embedding (must follow the first half). */
#define STLT_BWD_TEMPORAL 1
#define STLT_BWD_SPATIAL 2
#define STLT_BWD_ALL 3
int stlt_backward(void* handle, void* stream, const int64_t* categories, const float* boxes,
                  const float* scores_or_null, const int64_t* frame_types, const int64_t* lengths,
                  int32_t batch, int32_t frames, int32_t slots, void* workspace, size_t workspace_bytes,
                  float dropout_p, uint64_t seed, const float* d_logits, int32_t phases);

/* Dropout masks are not stored: every site regenerates its mask from (seed, site, element index)
 * with a counter-based hash, in the forward and again in the backward pass (same dropout_p / seed).
 * This host-side helper returns the multipliers (0 or 1/(1-p_eff), p_eff = round(p * 65536) / 65536)
 * of elements [first, first + n) of a site so a test can replay the exact masks in the oracle. Sites:
 * 0 category/box embedding output [token][768]; 1 frame embedding output [frame][768];
 * 16 + 4*layer + {0: attention probabilities [(query token * 12 + head) * 32 + key position],
 * 1: attention branch before the residual [row][768], 2: FFN inner [row][3072], 3: FFN branch
 * [row][768]}, spatial layers numbered first; rows of the last layer of each stack are the compacted
 * rows (frame index b*L + l for the spatial stack, video index b for the temporal stack). */
int stlt_op_dropout_mask(void* handle, float dropout_p, uint64_t seed, int32_t site, int64_t first,
                         int64_t n, float* multipliers_host);

/* Criterion: mean cross-entropy (labels i64 [rows], Something-Else) or mean BCE-with-logits (targets
 * f32 [rows, classes], Action Genome). loss_out (device float, may be NULL) receives the loss,
 * d_logits_out (may be NULL) its gradient times grad_scale. */
#define STLT_LOSS_CROSS_ENTROPY 0
#define STLT_LOSS_BCE_LOGITS 1
int stlt_loss(void* handle, void* stream, int32_t kind, const float* logits, const void* labels,
              int32_t rows, int32_t classes, float grad_scale, float* loss_out, float* d_logits_out);

/* Data-parallel training overlaps the gradient all-reduce with the rest of the backward pass (what DistributedDataParallel's
 * bucketed hooks do around the loop of src/train.py:117-135). With stage events enabled, stlt_backward records one event on
 * its stream as soon as the parameter gradients of a STAGE are final; stages in completion order:
 *   0 classifier head | 1 .. nt temporal layers nt-1 .. 0 | nt+1 frame embedding (position / frame-type tables, LayerNorm)
 *   | nt+2 .. nt+1+ns spatial layers ns-1 .. 0 | nt+ns+2 category / box / score embedding      (count = nt + ns + 3)
 * stlt_stream_wait_backward_stage makes `stream` (the communication stream) wait for the event of `stage` of the most
 * recent stlt_backward call, so that the all-reduce of that stage's gradient bucket starts while later stages still run. */
int stlt_backward_stage_events(void* handle, int32_t enable, int32_t* num_stages_out);
int stlt_stream_wait_backward_stage(void* handle, void* stream, int32_t stage);

/* sumsq_out[0] = sum(grads^2) over a flat fp32 buffer (the squared total norm of clip_grad_norm_), reduced in
 * a fixed order so that data-parallel ranks derive bit-identical clip coefficients from their all-reduced
 * gradients. scratch: caller-owned device floats (up to 1184 are used). */
int stlt_grad_sumsq(void* handle, void* stream, const float* grads, int64_t n, float* sumsq_out, float* scratch,
                    int32_t scratch_floats);

/* One AdamW step on flat fp32 buffers (torch.optim.AdamW semantics, decoupled weight decay, bias
 * correction for `step` >= 1). If sumsq is given, gradients are first scaled by
 * min(1, max_norm / (sqrt(sumsq) + 1e-6)) as clip_grad_norm_ does — read on the device, no sync. */
int stlt_adamw_step(void* handle, void* stream, float* params, const float* grads, float* exp_avg,
                    float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int32_t step, const float* sumsq_or_null, float max_norm);

/* ---- CACNF on precomputed appearance features (BASELINE.json configs[4]) ----
 * Replaces CrossAttentionCentralNetFusion.forward (models.py:526-549) with the 3D-ResNet trunk factored
 * out: `features` is what Resnet3D.forward_features returns (models.py:219-220), f32
 * [batch, feature_channels, appearance_tokens] (e.g. [B, 2048, 2*4*4]). Inference only; both precision modes
 * of the STLT path (STLT_PRECISION_BF16, or STLT_PRECISION_FP32 = 3-term split, logits within 1e-4). Weights are bound by their reference state_dict names ("backbone.layout_branch.*",
 * "backbone.appearance_branch.{projector,cls_token,pos_embed,transformer}.*", "backbone.mm_fusion.*",
 * "{layout,appearance,fusion}_classifier.*"); the ResNet trunk's and the two unused classifiers' entries
 * are ignored. The Conv3d projector weight [768, C, 1, 1, 1] is passed with its trailing 1x1x1 flattened.
 * The STLT part is packed with stlt_pack_weights, the rest with stlt_cacnf_pack_weights (same precision). Outputs: the four logits of `logit_names` (models.py:524), f32 [batch, classes]. */
int stlt_cacnf_bind_weights(void* handle, const StltTensor* tensors, int32_t count,
                            int32_t num_appearance_layers, int32_t num_fusion_layers,
                            int32_t appearance_tokens, int32_t feature_channels);
int stlt_cacnf_packed_weights_bytes(void* handle, int32_t precision, size_t* bytes);
int stlt_cacnf_pack_weights(void* handle, void* stream, int32_t precision, void* packed, size_t bytes);
int stlt_cacnf_workspace_bytes(void* handle, int32_t batch, int32_t frames, int32_t slots, int32_t precision,
                               size_t* bytes);
int stlt_cacnf_forward(void* handle, void* stream, int32_t precision, const int64_t* categories, const float* boxes,
                       const float* scores_or_null, const int64_t* frame_types, const int64_t* lengths,
                       const float* features, int32_t batch, int32_t frames, int32_t slots, void* workspace,
                       size_t workspace_bytes, float* logits_stlt, float* logits_resnet3d, float* logits_caf,
                       float* logits_ensemble);

/* ---- single-operator entry points used by the parity tests -------------------------------- */

/* out = epilogue(sum_terms A_t W_t^T + bias): A bf16 [terms>1 ? 2 : 1][m_rows][k], W bf16
 * [terms>1 ? 2 : 1][n][k]; out_kind 0 = f32 [m_rows][n], 1 = bf16, 2 = bf16 hi/lo planes. */
int stlt_op_gemm(void* handle, void* stream, const void* a_planes, const void* w_planes,
                 const float* bias, void* out, int32_t m_rows, int32_t n, int32_t k, int32_t terms,
                 int32_t out_kind, int32_t gelu /* 0 none, 1 erf GELU (erff), 2 fast erf GELU */);
/* Gradient GEMMs of the training step on the same tcgen05 kernel (bf16 operands, no bias):
 *   layout 1: out[m_rows][n]  = A[m_rows][k] * B[k][n]     (data gradient; out_kind 0 f32 / 1 bf16)
 *   layout 2: out[m_rows][n] += A[k][m_rows]^T * B[k][n]   (weight gradient; out f32, accumulated with
 *             TMA reduce-add stores, work split stream-K over the token axis k; any k >= 1). */
int stlt_op_gemm_grad(void* handle, void* stream, int32_t layout, const void* a, const void* b,
                      void* out, int32_t m_rows, int32_t n, int64_t k, int32_t out_kind);
/* Attention backward: qkv bf16 [tokens][2304] (saved by the forward), d_ctx bf16 [tokens][768] ->
 * d_qkv bf16 [tokens][2304]. impl 0 = CUDA cores, 1 = mma.sync tiles (the one the training step uses; it
 * can also add the column sums of d_qkv — the in-projection bias gradient — to d_bias f32 [2304]). */
int stlt_op_attention_bwd(void* handle, void* stream, const void* qkv, const void* d_ctx,
                          const int64_t* mask_src, int64_t num_seqs, int32_t seq_len, int32_t causal,
                          void* d_qkv, int32_t impl, void* d_bias_or_null);
/* Attention between two token streams (<= 64 queries / keys per sequence): Q from columns [0, 768) of
 * q_qkv bf16 [num_seqs*q_len][2304], K / V from columns [768, 1536) / [1536, 2304) of kv_qkv bf16
 * [num_seqs*kv_len][2304]; mask_src i64 per key token (0 = masked) or NULL. out bf16 [num_seqs*q_len][768]. */
int stlt_op_attention_cross(void* handle, void* stream, const void* q_qkv, const void* kv_qkv,
                            const int64_t* mask_src_or_null, int64_t num_seqs, int32_t q_len, int32_t kv_len,
                            int32_t causal, void* out_bf16);
/* The LayerNorm-fused projection GEMMs of the bf16 inference path (what stlt_forward runs by default):
 *   epilogue 1 (NORM_A): out bf16 [m_rows][n] = act(LN(z) W^T + b) computed from a = bf16(z) [m_rows][k],
 *     w = gamma-folded bf16 weights [n][k] and vec_a / vec_b = s / c of stlt_op_pack_folded, stats_in = per-row
 *     (sum, sum of squares) of z in 6 partial slots f32 [m_rows][6][2]; gelu 0 or 2.
 *   epilogue 2 (RESID, n must be 768): z f32 [m_rows][768] is updated IN PLACE to
 *     (prev_norm ? LN(z; vec_a = gamma, vec_b = beta, stats_in) : z) + a W^T + bias, out_bf16 receives its bf16
 *     copy and stats_out f32 [m_rows][6][2] the partial row statistics of the new z. */
int stlt_op_gemm_fused(void* handle, void* stream, int32_t epilogue, const void* a, int32_t m_rows, const void* w,
                       int32_t n, int32_t k, const float* bias_or_null, void* out, void* out_bf16_or_null,
                       int32_t gelu, const float* stats_in_or_null, const float* vec_a, const float* vec_b,
                       float* stats_out_or_null, float eps, int32_t prev_norm);
/* The same RESID epilogue on a residual stream stored as two bf16 planes (stlt_set_hilo_residual): z = z_hi + z_lo
 * (bf16 [m_rows][768] each) is updated in place to (prev_norm ? LN(z) : z) + a W^T + bias, re-split into the planes. */
/* The same two epilogues on the split operands of the fp32-parity mode (3 bf16 MMAs per product): a_planes bf16
 * [2][m_rows][k] and w_planes bf16 [2][n][k] are hi / lo plane pairs (stlt_op_pack_folded with flag 2 writes the
 * folded pair), NORM_A writes out bf16 [2][m_rows][n] (gelu 0 or 1 = exact erf), RESID updates z f32 in place and
 * writes its split to out_bf16_planes bf16 [2][m_rows][768]. */
int stlt_op_gemm_fused_split(void* handle, void* stream, int32_t epilogue, const void* a_planes, int32_t m_rows,
                             const void* w_planes, int32_t n, int32_t k, const float* bias, void* out,
                             void* out_bf16_planes, int32_t gelu, const float* stats_in, const float* vec_a,
                             const float* vec_b, float* stats_out, float eps, int32_t prev_norm);
int stlt_op_gemm_resid_hilo(void* handle, void* stream, const void* a, int32_t m_rows, const void* w, int32_t k,
                            const float* bias, void* z_hi, void* z_lo, const float* stats_in_or_null,
                            const float* gamma_or_null, const float* beta_or_null, float* stats_out, float eps,
                            int32_t prev_norm);
/* gamma-folded bf16 copy of a projection matrix w f32 [n][k] for the NORM_A epilogue: w_folded[n][k] =
 * bf16(w * gamma), s_out[n] = sum_k w_folded[n][k], c_out[n] = (w beta)[n] + bias[n]; gamma = beta = NULL
 * folds the identity. head_major != 0 (n = 2304 only) writes the rows of a packed in-projection in the order
 * h*192 + t*64 + j <- t*768 + h*64 + j (t = Q, K, V), the layout stlt_op_qkv_attention reads. Flag 2 in
 * head_major (fp32-parity mode): w_folded holds two planes [2][n][k], hi = bf16(w gamma), lo = bf16(w gamma - hi),
 * and s sums hi + lo. */
int stlt_op_pack_folded(void* handle, void* stream, const float* w, const float* gamma_or_null,
                        const float* beta_or_null, const float* bias, int32_t n, int32_t k, void* w_folded,
                        float* s_out, float* c_out, int32_t head_major);
/* In-projection + masked self-attention in one kernel (see stlt_set_fused_attention): a bf16 [m_rows][768],
 * w_head_major / vec_s / vec_c from stlt_op_pack_folded(head_major = 1), stats_or_null f32 [m_rows][6][2]
 * (non-NULL: the rows of `a` are an un-normalised residual stream whose LayerNorm is folded into the weights),
 * mask_src i64 [valid_rows] (key masked when 0), num_seqs sequences of seq_len <= 32 consecutive rows.
 * ctx bf16 [m_rows][768]. */
int stlt_op_qkv_attention(void* handle, void* stream, const void* a, int64_t m_rows, int64_t valid_rows,
                          const void* w_head_major, const float* vec_s, const float* vec_c,
                          const float* stats_or_null, float eps, const int64_t* mask_src, int64_t num_seqs,
                          int32_t seq_len, int32_t causal, void* ctx);
int stlt_op_gemm_simt(void* handle, void* stream, const float* a, const float* w,
                      const float* bias, float* out, int32_t m, int32_t n, int32_t k, int32_t gelu);
int stlt_op_attention(void* handle, void* stream, const void* qkv, int32_t qkv_is_bf16,
                      const int64_t* mask_src, int64_t num_seqs, int32_t seq_len, int32_t causal,
                      void* out_bf16, int32_t planes, int64_t plane_rows);
int stlt_op_add_ln(void* handle, void* stream, const float* x, const float* y_or_null,
                   const float* gamma, const float* beta, float eps, int64_t rows, float* out_f32,
                   void* out_bf16, int32_t planes, int64_t plane_rows);
int stlt_op_pack_bf16(void* handle, void* stream, const float* src, void* dst, int64_t n,
                      int32_t planes);

#ifdef __cplusplus
}
#endif

#endif /* STLT_B200_H_ */
