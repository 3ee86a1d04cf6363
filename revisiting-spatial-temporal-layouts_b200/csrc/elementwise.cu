// HBM-bound stages of the STLT forward: box preparation + masks (K0), category/box embedding + LN
// (K1), residual add + LayerNorm, frame embedding (K7), last-frame gather (K9), weight packing.
// One warp owns one 768-wide row; lane l holds columns 4*l + 128*k + {0..3}, k = 0..5, so every
// global access is a coalesced 16-byte vector.
#include "rowops.cuh"

namespace stlt {

namespace {

__device__ __forceinline__ float4 fix_and_normalize_box(const double* rb, long long W, long long H);

// ------------------------------------------------------------------------------------------------
// K0: fix_box (src/utils/data_utils.py:205-231) + division by the video size
// (src/modelling/datasets.py:54,82) + the two padding masks (src/modelling/datasets.py:274-286).
// One thread per (video, frame, slot).
// ------------------------------------------------------------------------------------------------
__global__ void prepare_kernel(const double* __restrict__ raw_boxes,
                               const long long* __restrict__ video_sizes,
                               const long long* __restrict__ categories,
                               const long long* __restrict__ frame_types, int L, int S,
                               long long total, float4* __restrict__ boxes_out,
                               uint8_t* __restrict__ mask_boxes, uint8_t* __restrict__ mask_frames) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int s = static_cast<int>(idx % S);
  const long long frame = idx / S;
  const long long b = frame / L;
  const long long cat = categories[idx];
  mask_boxes[idx] = (cat == 0) ? 1 : 0;  // src/modelling/datasets.py:277
  if (s == 0) mask_frames[frame] = (frame_types[frame] == 0) ? 1 : 0;  // :283-285

  float4 out;
  if (s == 0) {
    // CLS object / extract frame / padded frame: box [0, 0, 1, 1] (datasets.py:70,100,263)
    out = make_float4(0.f, 0.f, 1.f, 1.f);
  } else if (cat == 0) {
    out = make_float4(0.f, 0.f, 0.f, 0.f);  // padded object slot (datasets.py:91)
  } else {
    out = fix_and_normalize_box(raw_boxes + idx * 4, video_sizes[2 * b + 0], video_sizes[2 * b + 1]);
  }
  boxes_out[idx] = out;
}

__device__ __forceinline__ float4 fix_and_normalize_box(const double* rb, long long W, long long H) {
  // fix_box (src/utils/data_utils.py:205-231) then int64 / int64 true-divide (datasets.py:82)
  long long x1 = static_cast<long long>(rb[0]);  // int(): truncation toward zero
  long long y1 = static_cast<long long>(rb[1]);
  long long x2 = static_cast<long long>(rb[2]);
  long long y2 = static_cast<long long>(rb[3]);
  x1 = x1 < 0 ? 0 : x1;
  y1 = y1 < 0 ? 0 : y1;
  x2 = x2 < 0 ? 0 : x2;
  y2 = y2 < 0 ? 0 : y2;
  if (x1 > x2) { const long long t = x1; x1 = x2; x2 = t; }
  if (y1 > y2) { const long long t = y1; y1 = y2; y2 = t; }
  if (x1 >= W) x1 = W - 1;
  if (y1 >= H) y1 = H - 1;
  if (x2 >= W) x2 = W - 1;
  if (y2 >= H) y2 = H - 1;
  if (x1 == x2 && x1 == 0) x2 = 1;
  if (y1 == y2 && y1 == 0) y2 = 1;
  if (x1 == x2) x1 -= 1;
  if (y1 == y2) y1 -= 1;
  const float fw = __ll2float_rn(W), fh = __ll2float_rn(H);
  float4 out;
  out.x = __fdiv_rn(__ll2float_rn(x1), fw);
  out.y = __fdiv_rn(__ll2float_rn(y1), fh);
  out.z = __fdiv_rn(__ll2float_rn(x2), fw);
  out.w = __fdiv_rn(__ll2float_rn(y2), fh);
  return out;
}

// ------------------------------------------------------------------------------------------------
// Batch builder: StltDataset.__getitem__ (src/modelling/datasets.py:52-125) + StltCollater.__call__
// (:243-288) for B videos of a CSR layout store, one thread per output slot (b, l, s):
//   frames [0, n_b)   : sampled frames; slot 0 = CLS object, slot s >= 1 = the (s-1)-th object with
//                       score >= threshold (order preserved), fix_box + normalisation; rest padding
//   frame n_b         : the "extract" frame (CLS only)
//   frames (n_b, L)   : collater padding (categories [cls, 0..], slot-0 box [0,0,1,1], type 0)
// Test-time frame sampling (get_test_layout_indices, src/utils/data_utils.py:47-56) is evaluated
// here when no explicit indices are given: int(tick / 2 + tick * x), tick = n / T, in fp64.
// ------------------------------------------------------------------------------------------------
__global__ void build_batch_kernel(BatchStore st, const long long* __restrict__ video_index,
                                   const long long* __restrict__ frame_indices,
                                   const long long* __restrict__ num_sampled, int T, int L, int S,
                                   double score_threshold, BatchIds ids, long long total,
                                   long long* __restrict__ categories, float4* __restrict__ boxes,
                                   float* __restrict__ scores, long long* __restrict__ frame_types,
                                   long long* __restrict__ lengths, uint8_t* __restrict__ mask_boxes,
                                   uint8_t* __restrict__ mask_frames, int* __restrict__ err_flag) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int s = static_cast<int>(idx % S);
  const long long frame_slot = idx / S;
  const int l = static_cast<int>(frame_slot % L);
  const long long b = frame_slot / L;
  const long long vid = video_index[b];
  const long long f0 = st.video_frame_offsets[vid];
  const long long n_frames = st.video_frame_offsets[vid + 1] - f0;
  long long n_s;  // sampled frames of this video
  if (num_sampled != nullptr) n_s = num_sampled[b];
  else n_s = n_frames > T ? T : n_frames;
  if (n_s > L - 1) {
    if (s == 0 && l == 0) atomicExch(err_flag, 4);
    n_s = L - 1;
  }

  long long cat = 0, ftype = 0;
  float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
  float score = 0.f;
  if (s == 0) {  // CLS object / extract / collater padding all carry the same slot-0 payload
    cat = ids.cls;
    box = make_float4(0.f, 0.f, 1.f, 1.f);
    score = 1.f;
  }
  if (l < n_s) {
    long long fi;
    if (frame_indices != nullptr) {
      fi = frame_indices[b * T + l];
    } else if (n_frames > T) {
      const double tick = static_cast<double>(n_frames) * 1.0 / static_cast<double>(T);
      fi = static_cast<long long>(tick / 2.0 + tick * static_cast<double>(l));
    } else {
      fi = l;
    }
    if (fi < 0 || fi >= n_frames) {
      if (s == 0) atomicExch(err_flag, 5);
      fi = 0;
    }
    const long long o0 = st.frame_object_offsets[f0 + fi];
    const long long o1 = st.frame_object_offsets[f0 + fi + 1];
    ftype = (o1 == o0) ? ids.ft_empty : ids.ft_regular;  // decided before the score filter (:65-69)
    if (s > 0) {
      int seen = 0;
      for (long long o = o0; o < o1; ++o) {
        if (st.obj_scores[o] < score_threshold) continue;  // :75
        if (++seen == s) {
          const long long W = st.video_sizes[2 * vid], H = st.video_sizes[2 * vid + 1];
          box = fix_and_normalize_box(st.obj_boxes + o * 4, W, H);
          cat = st.obj_categories[o];
          score = __double2float_rn(st.obj_scores[o]);
          break;
        }
      }
    } else {
      // more kept objects than slots cannot happen when S-1 is the dataset maximum (:38-47)
      int kept = 0;
      for (long long o = o0; o < o1; ++o) kept += st.obj_scores[o] >= score_threshold ? 1 : 0;
      if (kept > S - 1) atomicExch(err_flag, 6);
    }
  } else if (l == n_s) {
    ftype = ids.ft_extract;
  } else {
    ftype = ids.ft_pad;
  }
  categories[idx] = cat;
  boxes[idx] = box;
  if (scores != nullptr) scores[idx] = score;
  mask_boxes[idx] = cat == 0 ? 1 : 0;
  if (s == 0) {
    frame_types[frame_slot] = ftype;
    mask_frames[frame_slot] = ftype == ids.ft_pad ? 1 : 0;
    if (l == 0) lengths[b] = n_s + 1;
  }
}

// ------------------------------------------------------------------------------------------------
// Evaluators (src/utils/evaluation.py:21-34): top-1 / top-5 hit counters, one warp per row.
// argmax / topk semantics: the label is a top-k hit when fewer than k logits rank before it
// (greater, or equal with a smaller index).
// ------------------------------------------------------------------------------------------------
__global__ void topk_count_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                  int rows, int classes, unsigned long long* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long top1 = 0, top5 = 0;
  for (int r = warp0; r < rows; r += nwarps) {
    const long long lab = labels[r];
    if (lab < 0 || lab >= classes) continue;
    const float* row = logits + static_cast<long long>(r) * classes;
    const float ref = row[lab];
    int before = 0;
    for (int c = lane; c < classes; c += 32) {
      const float v = row[c];
      before += (v > ref || (v == ref && c < lab)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    top1 += before == 0 ? 1 : 0;
    top5 += before < 5 ? 1 : 0;
  }
  if (lane == 0 && (top1 | top5)) {
    atomicAdd(counters + 0, top1);
    atomicAdd(counters + 1, top5);
  }
}

// ------------------------------------------------------------------------------------------------
// K1: CategoryBoxEmbeddings.forward (src/modelling/models.py:29-39)
// ------------------------------------------------------------------------------------------------
// The pre-LayerNorm row is affine in u = (box, score, 1):  x[col] = E[cat][col] + sum_q u_q V_q[col], so its
// LayerNorm statistics need no reduction over the 768 columns at run time:
//   mean = mean(E[cat]) + sum_q u_q mean(V_q)                      -> removed by centring E, V_q once
//   var  = u^T G_cat u / 768, G_cat = Gram matrix of the centred vectors (w0..w3, score_w, E[cat] + bias)
// `embed_stats_kernel` (one CTA per category, fp64 accumulation) writes mean(E[cat]) and the 21 Gram entries;
// `embed_kernel` then has thread t of 192 own columns 4t..4t+3 with their centred parameters resident in
// registers, evaluates the 6x6 quadratic form for 8 tokens in lanes 0-7 of every warp, and streams the 4.6 KB
// per token out as 512-byte warp stores. No shared memory, no barrier, no cross-lane reduction.
constexpr int kEmbedTok = 8;
constexpr int kEmbedThreads = kHidden / 4;  // 192
constexpr int kEmbedCatStride = 24;         // doubles per category: mean(E[cat]), 21 quadratic-form coefficients
constexpr int kEmbedGlobal = 8;             // after the categories: mean of w0..w3, score_w, bias

__global__ void __launch_bounds__(256)
embed_stats_kernel(const float* __restrict__ cat_table, int unique_categories,
                   const float* __restrict__ box_w, const float* __restrict__ box_b,
                   const float* __restrict__ score_w, const float* __restrict__ score_b,
                   double* __restrict__ stats) {
  // One sweep over the 768 columns in fp64: raw first and second moments of v = (w0, w1, w2, w3, score_w,
  // E[cat] + bias) plus the sums of E[cat] and of the bias alone; the centred Gram matrix is
  // (sum v_i v_j - sum v_i sum v_j / n) / n (exact to ~1e-16 relative in fp64).
  constexpr int kAcc = 6 + 21 + 2;
  __shared__ double red[8][kAcc];
  const int cat = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* erow = cat_table + static_cast<long long>(cat) * kHidden;
  double a[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) a[i] = 0;
#pragma unroll
  for (int r = 0; r < kHidden / 256; ++r) {
    const int col = tid + 256 * r;
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(box_w) + col);
    const double bias = static_cast<double>(box_b[col]) + (score_b != nullptr ? score_b[col] : 0.f);
    const double e = erow[col];
    const double v[6] = {w4.x, w4.y, w4.z, w4.w, score_w != nullptr ? score_w[col] : 0.f, e + bias};
    int k = 6;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      a[i] += v[i];
#pragma unroll
      for (int j = i; j < 6; ++j) a[k++] += v[i] * v[j];
    }
    a[27] += e;
    a[28] += bias;
  }
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
    if (lane == 0) red[warp][i] = a[i];
  }
  __syncthreads();
  if (tid < kAcc) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][tid];
    red[0][tid] = t;  // column tid is only touched by thread tid
  }
  __syncthreads();
  const double* tot = red[0];
  double* out = stats + static_cast<long long>(cat) * kEmbedCatStride;
  if (tid < 21) {
    // entry tid of the upper triangle -> (i, j); off-diagonal terms appear twice in u^T G u
    int i = 0, rem = tid;
    while (rem >= 6 - i) { rem -= 6 - i; ++i; }
    const int j = i + rem;
    const double g = (tot[6 + tid] - tot[i] * tot[j] / kHidden) / kHidden;
    out[1 + tid] = rem == 0 ? g : 2.0 * g;
  }
  if (tid == 21) out[0] = tot[27] / kHidden;
  if (tid == 22 || tid == 23) out[tid] = 0.0;
  if (cat == 0 && tid < kEmbedGlobal) {
    const double m = tid < 5 ? tot[tid] : (tid == 5 ? tot[28] : 0.0);
    stats[static_cast<long long>(unique_categories) * kEmbedCatStride + tid] = m / kHidden;
  }
}

__global__ void __launch_bounds__(kEmbedThreads, 4)
embed_kernel(const long long* __restrict__ categories, const float4* __restrict__ boxes,
             const float* __restrict__ scores, const float* __restrict__ cat_table,
             int unique_categories, const float* __restrict__ box_w,
             const float* __restrict__ box_b, const float* __restrict__ score_w,
             const float* __restrict__ score_b, const float* __restrict__ ln_g,
             const float* __restrict__ ln_b, float eps, long long tokens, ActOut out,
             int* __restrict__ err_flag, DropCfg drop, const double* __restrict__ stats,
             const int* __restrict__ frame_row, int S, long long* __restrict__ mask_out) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int c = 4 * tid;
  // resident, centred parameters of this thread's four columns
  const double* glob = stats + static_cast<long long>(unique_categories) * kEmbedCatStride;
  const float4 mw = make_float4(static_cast<float>(glob[0]), static_cast<float>(glob[1]), static_cast<float>(glob[2]),
                                static_cast<float>(glob[3]));
  const float msw = static_cast<float>(glob[4]), mb = static_cast<float>(glob[5]);
  float4 w[4];  // box_w is [768][4]: one float4 per feature
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    w[q] = __ldg(reinterpret_cast<const float4*>(box_w) + c + q);
    w[q].x -= mw.x; w[q].y -= mw.y; w[q].z -= mw.z; w[q].w -= mw.w;
  }
  float4 bias = __ldg(reinterpret_cast<const float4*>(box_b + c));
  float4 sw = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scores != nullptr) {
    sw = __ldg(reinterpret_cast<const float4*>(score_w + c));  // [768][1]
    sw.x -= msw; sw.y -= msw; sw.z -= msw; sw.w -= msw;
    const float4 sb = __ldg(reinterpret_cast<const float4*>(score_b + c));
    bias.x += sb.x; bias.y += sb.y; bias.z += sb.z; bias.w += sb.w;
  }
  bias.x -= mb; bias.y -= mb; bias.z -= mb; bias.w -= mb;
  const float4 gam = __ldg(reinterpret_cast<const float4*>(ln_g + c));
  const float4 bet = __ldg(reinterpret_cast<const float4*>(ln_b + c));

  const long long groups = (tokens + kEmbedTok - 1) / kEmbedTok;
  for (long long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const long long t0 = grp * kEmbedTok;
    // ---- row statistics of token t0 + (lane & 7) from the category's quadratic form ----
    long long tl = t0 + (lane & 7);
    if (tl >= tokens) tl = tokens - 1;  // tail: recompute the last token
    long long cat_l = __ldg(categories + tl);
    if (cat_l < 0 || cat_l >= unique_categories) {
      if (tid < kEmbedTok) atomicExch(err_flag, 1);  // lanes 0-7 of warp 0 cover the group's eight tokens
      cat_l = 0;
    }
    const float4 box_l = __ldg(boxes + tl);
    const float sc_l = scores != nullptr ? __ldg(scores + tl) : 0.f;
    float mean_l, rstd_l;
    {
      // the quadratic form is evaluated in fp64 (22 loads and 27 DFMA per token, in one lane): its terms may cancel,
      // and this way the variance is as exact as a two-pass reduction
      const double* gq = stats + cat_l * kEmbedCatStride;
      mean_l = static_cast<float>(gq[0]);
      const double u[6] = {box_l.x, box_l.y, box_l.z, box_l.w, sc_l, 1.0};
      double var = 0.0;
      int k = 1;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double inner = 0.0;
#pragma unroll
        for (int j = i; j < 6; ++j) inner += gq[k++] * u[j];
        var += inner * u[i];
      }
      rstd_l = 1.0f / sqrtf(fmaxf(static_cast<float>(var), 0.f) + eps);
    }
    const int cat_i = static_cast<int>(cat_l);
    // pad-skipping layout: the compact row of token t0 + (lane & 7) (-1: dead), looked up next to its statistics so that the
    // stores below do not wait for a dependent load (32-bit division: the token count is below 2^31, checked by the caller)
    int dst_l = static_cast<int>(tl);
    if (frame_row != nullptr) {
      const unsigned f = static_cast<unsigned>(tl) / static_cast<unsigned>(S);
      const int slot = static_cast<int>(static_cast<unsigned>(tl) - f * static_cast<unsigned>(S));
      const int fr = __ldg(frame_row + f);
      dst_l = (fr < 0 || ((fr & kSingleFrameFlag) && slot != 0)) ? -1 : (fr & ~kSingleFrameFlag) + slot;
    }
#pragma unroll
    for (int i = 0; i < kEmbedTok; ++i) {
      const long long t = t0 + i;
      if (t >= tokens) break;
      long long dst = t;  // output row
      if (frame_row != nullptr) {  // scatter to the compact row, skip dead tokens (block-uniform)
        const int d = __shfl_sync(0xffffffffu, dst_l, i);
        if (d < 0) continue;
        dst = d;
      }
      const int cat = __shfl_sync(0xffffffffu, cat_i, i);
      if (mask_out != nullptr && tid == 0) mask_out[dst] = cat;
      const float mean = __shfl_sync(0xffffffffu, mean_l, i);
      const float rstd = __shfl_sync(0xffffffffu, rstd_l, i);
      const float4 box = __ldg(boxes + t);
      const float4 row = __ldg(reinterpret_cast<const float4*>(cat_table + static_cast<long long>(cat) * kHidden + c));
      float v[4] = {bias.x - mean, bias.y - mean, bias.z - mean, bias.w - mean};
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] += box.x * w[q].x + box.y * w[q].y + box.z * w[q].z + box.w * w[q].w;
      if (scores != nullptr) {
        const float sc = __ldg(scores + t);
        v[0] += sc * sw.x; v[1] += sc * sw.y; v[2] += sc * sw.z; v[3] += sc * sw.w;
      }
      float4 y;
      y.x = (row.x + v[0]) * rstd * gam.x + bet.x;
      y.y = (row.y + v[1]) * rstd * gam.y + bet.y;
      y.z = (row.z + v[2]) * rstd * gam.z + bet.z;
      y.w = (row.w + v[3]) * rstd * gam.w + bet.w;
      if (drop.thr16 != 0) {  // models.py:27,38 (training only); element index = token * 768 + column
        const unsigned long long pair = (static_cast<unsigned long long>(t) * kHidden + c) >> 1;
        const uint32_t b0 = drop_bits(drop.key, pair), b1 = drop_bits(drop.key, pair + 1);
        y.x *= drop_mul(b0, 0, drop);
        y.y *= drop_mul(b0, 1, drop);
        y.z *= drop_mul(b1, 0, drop);
        y.w *= drop_mul(b1, 1, drop);
      }
      if (out.x != nullptr) *reinterpret_cast<float4*>(out.x + dst * kHidden + c) = y;
      if (out.xb != nullptr) {
        uint2 h;
        h.x = pack_bf16x2(y.x, y.y);
        h.y = pack_bf16x2(y.z, y.w);
        *reinterpret_cast<uint2*>(out.xb + dst * kHidden + c) = h;
        if (out.planes == 2) {
          uint2 l;
          l.x = pack_bf16x2(bf16_residual(y.x), bf16_residual(y.y));
          l.y = pack_bf16x2(bf16_residual(y.z), bf16_residual(y.w));
          *reinterpret_cast<uint2*>(out.xb + (out.plane_rows + dst) * kHidden + c) = l;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// x <- LN(x + y): the two post-norm residual sites of every encoder layer.
// ------------------------------------------------------------------------------------------------
// bf16 branch output (bf16 inference mode): 96 x 16-byte loads per row instead of 192
__device__ __forceinline__ RowRegs load_row_bf16(const __nv_bfloat16* base, long long row, int lane) {
  RowRegs r;
  const uint2* p = reinterpret_cast<const uint2*>(base + row * kHidden);
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const uint2 v = __ldg(p + lane + 32 * k);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
    r.v[k] = make_float4(a.x, a.y, b.x, b.y);
  }
  return r;
}

template <typename TY>
__global__ void __launch_bounds__(256)
add_ln_kernel(const float* __restrict__ x_in, const TY* __restrict__ y,
              const float* __restrict__ g, const float* __restrict__ b, float eps, long long rows,
              ActOut out, float* __restrict__ z_out, DropCfg drop, const int* __restrict__ rows_dyn) {
  if (rows_dyn != nullptr) {  // pad-skipping layout: the live row count is decided on the device
    const long long r = __ldg(rows_dyn);
    rows = r < rows ? r : rows;
  }
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long row = warp0; row < rows; row += nwarps) {
    RowRegs r = load_row(x_in, row, lane);
    if (y != nullptr) {
      RowRegs a;
      if constexpr (sizeof(TY) == 2) a = load_row_bf16(y, row, lane);
      else a = load_row(y, row, lane);
      drop_row(a, row, lane, drop);  // dropout1 / dropout2 of nn.TransformerEncoderLayer (training only)
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        r.v[k].x += a.v[k].x;
        r.v[k].y += a.v[k].y;
        r.v[k].z += a.v[k].z;
        r.v[k].w += a.v[k].w;
      }
    }
    if (z_out != nullptr) {  // training: the pre-LayerNorm sum is what the backward pass re-normalises
      float4* pz = reinterpret_cast<float4*>(z_out + row * kHidden);
#pragma unroll
      for (int k = 0; k < kVec; ++k) pz[lane + 32 * k] = r.v[k];
    }
    layer_norm_row(r, g, b, eps, lane);
    store_act(out, row, r, lane);
  }
}

// ------------------------------------------------------------------------------------------------
// K7: FramesEmbeddings.forward (src/modelling/models.py:98-111); position_ids[:, :L] = 0..L-1.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
frame_embed_kernel(const float* __restrict__ spatial_x, int S,
                   const long long* __restrict__ frame_types, const float* __restrict__ pos_table,
                   const float* __restrict__ ft_table, int n_frame_types,
                   const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps, int L,
                   long long frames, ActOut out, int* __restrict__ err_flag, DropCfg drop,
                   const float* __restrict__ pre_g, const float* __restrict__ pre_b, float pre_eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long f = warp0; f < frames; f += nwarps) {
    const int l = static_cast<int>(f % L);
    long long ft = frame_types[f];
    if (ft < 0 || ft >= n_frame_types) {
      if (lane == 0) atomicExch(err_flag, 2);
      ft = 0;
    }
    RowRegs r = load_row(spatial_x, f * S, lane);  // slot 0 = CLS object (models.py:79)
    // fused-LayerNorm path: the spatial stack hands over its PRE-norm output, normalise it here
    if (pre_g != nullptr) layer_norm_row(r, pre_g, pre_b, pre_eps, lane);
    const RowRegs p = load_row(pos_table, l, lane);
    const RowRegs t = load_row(ft_table, ft, lane);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      r.v[k].x = (r.v[k].x + p.v[k].x) + t.v[k].x;
      r.v[k].y = (r.v[k].y + p.v[k].y) + t.v[k].y;
      r.v[k].z = (r.v[k].z + p.v[k].z) + t.v[k].z;
      r.v[k].w = (r.v[k].w + p.v[k].w) + t.v[k].w;
    }
    layer_norm_row(r, ln_g, ln_b, eps, lane);
    drop_row(r, f, lane, drop);  // models.py:93,110 (training only)
    store_act(out, f, r, lane);
  }
}

// K9: stlt_output[lengths - 1, arange(B)] (src/modelling/models.py:189-192)
__global__ void __launch_bounds__(256)
gather_last_kernel(const float* __restrict__ x, const long long* __restrict__ lengths, int B, int L,
                   float* __restrict__ out, int* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long b = warp0; b < B; b += nwarps) {
    long long len = lengths[b];
    if (len < 1 || len > L) {
      if (lane == 0) atomicExch(err_flag, 3);
      len = 1;
    }
    const RowRegs r = load_row(x, b * L + (len - 1), lane);
    float4* p = reinterpret_cast<float4*>(out + b * kHidden);
#pragma unroll
    for (int k = 0; k < kVec; ++k) p[lane + 32 * k] = r.v[k];
  }
}

// Row compaction for the pruned last layer of each stack: copies the fp32 residual row and the
// bf16 attention-context row(s) of the tokens whose output is actually consumed.
//   stride > 0 : source row = r * stride            (spatial CLS slot, models.py:79)
//   stride == 0: source row = r * L + lengths[r] - 1 (extract frame, models.py:189-192)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src_x, const __nv_bfloat16* __restrict__ src_att,
                   int planes, long long src_plane_rows, int stride,
                   const long long* __restrict__ lengths, int L, long long rows,
                   float* __restrict__ dst_x, __nv_bfloat16* __restrict__ dst_att,
                   long long dst_plane_rows, int* __restrict__ err_flag,
                   const float2* __restrict__ src_stats, float2* __restrict__ dst_stats,
                   const __nv_bfloat16* __restrict__ src_hi, const __nv_bfloat16* __restrict__ src_lo) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    long long src;
    if (stride > 0) {
      src = r * stride;
    } else {
      long long len = lengths[r];
      if (len < 1 || len > L) {
        if (lane == 0) atomicExch(err_flag, 3);
        len = 1;
      }
      src = r * L + (len - 1);
    }
    if (src_stats != nullptr && lane < kStatSlots)  // fused-LN path: the row's partial statistics travel along
      dst_stats[r * kStatSlots + lane] = src_stats[src * kStatSlots + lane];
    const RowRegs x = src_hi != nullptr ? load_row_hilo(src_hi, src_lo, src, lane) : load_row(src_x, src, lane);
    float4* px = reinterpret_cast<float4*>(dst_x + r * kHidden);
#pragma unroll
    for (int k = 0; k < kVec; ++k) px[lane + 32 * k] = x.v[k];
    for (int pl = 0; pl < planes; ++pl) {
      const uint4* sa = reinterpret_cast<const uint4*>(src_att + (pl * src_plane_rows + src) * kHidden);
      uint4* da = reinterpret_cast<uint4*>(dst_att + (pl * dst_plane_rows + r) * kHidden);
#pragma unroll
      for (int k = 0; k < 3; ++k) da[lane + 32 * k] = __ldg(sa + lane + 32 * k);  // 96 x 16 B
    }
  }
}

__global__ void pack_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ hi,
                                 uint2* __restrict__ lo, long long n4) {
  const long long stride = gridDim.x * static_cast<long long>(blockDim.x);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += stride) {
    const float4 v = __ldg(src + i);
    uint2 h;
    h.x = pack_bf16x2(v.x, v.y);
    h.y = pack_bf16x2(v.z, v.w);
    hi[i] = h;
    if (lo != nullptr) {
      uint2 l;
      l.x = pack_bf16x2(bf16_residual(v.x), bf16_residual(v.y));
      l.y = pack_bf16x2(bf16_residual(v.z), bf16_residual(v.w));
      lo[i] = l;
    }
  }
}

// Fused-LayerNorm weights: Wf[n, k] = bf16(W[n, k] * gamma[k]), s[n] = sum_k Wf[n, k] (of the ROUNDED
// values, so that mean * s cancels exactly what the MMA accumulates), c[n] = sum_k W[n, k] * beta[k] + bias[n].
// One block per output row n.
__global__ void __launch_bounds__(256)
pack_folded_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ bias, int k, __nv_bfloat16* __restrict__ wf, float* __restrict__ s_out,
                   float* __restrict__ c_out, int head_major, long long lo_offset) {
  const int n = blockIdx.x;  // output row
  // head-major order of the packed in-projection: output row h*192 + t*64 + j <- source row t*768 + h*64 + j
  const int src = head_major ? ((n % 192) / 64) * kHidden + (n / 192) * kHeadDim + (n % 64) : n;
  const float* row = w + static_cast<long long>(src) * k;
  float s = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const float v = row[i];
    const float folded = gamma != nullptr ? v * gamma[i] : v;
    const __nv_bfloat16 f = __float2bfloat16_rn(folded);
    wf[static_cast<long long>(n) * k + i] = f;
    s += __bfloat162float(f);
    if (lo_offset != 0) {  // fp32-parity mode: the remainder plane; s sums what the 3-term product multiplies
      const __nv_bfloat16 lo = __float2bfloat16_rn(folded - __bfloat162float(f));
      wf[lo_offset + static_cast<long long>(n) * k + i] = lo;
      s += __bfloat162float(lo);
    }
    if (beta != nullptr) c = fmaf(v, beta[i], c);
  }
  s = warp_sum(s);
  c = warp_sum(c);
  __shared__ float ps[8], pc[8];
  if ((threadIdx.x & 31) == 0) {
    ps[threadIdx.x >> 5] = s;
    pc[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tc = 0.f;
    for (int i = 0; i < 8; ++i) {
      ts += ps[i];
      tc += pc[i];
    }
    s_out[n] = ts;
    c_out[n] = tc + bias[src];
  }
}

// Padding masks of StltCollater (src/modelling/datasets.py:274-286) from the already padded ids.
__global__ void masks_kernel(const long long* __restrict__ categories,
                             const long long* __restrict__ frame_types, long long n_slots,
                             long long n_frames, uint8_t* __restrict__ mask_boxes,
                             uint8_t* __restrict__ mask_frames) {
  const long long stride = gridDim.x * static_cast<long long>(blockDim.x);
  const long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (mask_boxes != nullptr)
    for (long long i = i0; i < n_slots; i += stride) mask_boxes[i] = categories[i] == 0 ? 1 : 0;
  if (mask_frames != nullptr)
    for (long long i = i0; i < n_frames; i += stride) mask_frames[i] = frame_types[i] == 0 ? 1 : 0;
}

}  // namespace

cudaError_t launch_prepare(const double* raw_boxes, const long long* video_sizes,
                           const long long* categories, const long long* frame_types, int B, int L,
                           int S, float* boxes_out, uint8_t* mask_boxes, uint8_t* mask_frames,
                           cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * L * S;
  if (total == 0) return cudaSuccess;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  prepare_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
      raw_boxes, video_sizes, categories, frame_types, L, S, total,
      reinterpret_cast<float4*>(boxes_out), mask_boxes, mask_frames);
  return cudaGetLastError();
}

cudaError_t launch_masks(const long long* categories, const long long* frame_types,
                         long long n_slots, long long n_frames, uint8_t* mask_boxes,
                         uint8_t* mask_frames, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  long long blocks = (n_slots + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  masks_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(categories, frame_types, n_slots,
                                                                  n_frames, mask_boxes, mask_frames);
  return cudaGetLastError();
}

size_t embed_scratch_bytes(int unique_categories) {
  return (static_cast<size_t>(unique_categories) * kEmbedCatStride + kEmbedGlobal) * sizeof(double);
}

cudaError_t launch_embed(const long long* categories, const float* boxes, const float* scores,
                         const float* cat_table, int unique_categories, const float* box_w,
                         const float* box_b, const float* score_w, const float* score_b,
                         const float* ln_g, const float* ln_b, float eps, long long tokens,
                         ActOut out, int* err_flag, cudaStream_t stream, float* scratch,
                         size_t scratch_bytes, DropCfg drop, const int* frame_row, int S, long long* mask_out) {
  if (tokens == 0) return cudaSuccess;
  if (scratch == nullptr || scratch_bytes < embed_scratch_bytes(unique_categories) || unique_categories < 1)
    return cudaErrorInvalidValue;
  embed_stats_kernel<<<unique_categories, 256, 0, stream>>>(
      cat_table, unique_categories, box_w, box_b, scores != nullptr ? score_w : nullptr,
      scores != nullptr ? score_b : nullptr, reinterpret_cast<double*>(scratch));
  long long blocks = (tokens + kEmbedTok - 1) / kEmbedTok;
  if (blocks > 148LL * 4 * 8) blocks = 148LL * 4 * 8;  // 4 resident CTAs per SM, grid-stride inside
  embed_kernel<<<static_cast<unsigned>(blocks), kEmbedThreads, 0, stream>>>(
      categories, reinterpret_cast<const float4*>(boxes), scores, cat_table, unique_categories,
      box_w, box_b, score_w, score_b, ln_g, ln_b, eps, tokens, out, err_flag, drop,
      reinterpret_cast<const double*>(scratch), frame_row, S, mask_out);
  return cudaGetLastError();
}

cudaError_t launch_add_ln(const float* x_in, const float* y, const float* g, const float* b,
                          float eps, long long rows, ActOut out, cudaStream_t stream, float* z_out,
                          DropCfg drop, const int* rows_dyn) {
  if (rows == 0) return cudaSuccess;
  add_ln_kernel<float><<<row_grid(rows, 8), 256, 0, stream>>>(x_in, y, g, b, eps, rows, out, z_out, drop, rows_dyn);
  return cudaGetLastError();
}

cudaError_t launch_add_ln_bf16y(const float* x_in, const __nv_bfloat16* y, const float* g, const float* b,
                                float eps, long long rows, ActOut out, cudaStream_t stream, float* z_out,
                                DropCfg drop, const int* rows_dyn) {
  if (rows == 0) return cudaSuccess;
  add_ln_kernel<__nv_bfloat16><<<row_grid(rows, 8), 256, 0, stream>>>(x_in, y, g, b, eps, rows, out, z_out, drop, rows_dyn);
  return cudaGetLastError();
}

cudaError_t launch_frame_embed(const float* spatial_x, int S, const long long* frame_types,
                               const float* pos_table, const float* ft_table, int n_frame_types,
                               const float* ln_g, const float* ln_b, float eps, int B, int L,
                               ActOut out, int* err_flag, cudaStream_t stream, DropCfg drop,
                               const float* pre_g, const float* pre_b, float pre_eps) {
  const long long frames = static_cast<long long>(B) * L;
  if (frames == 0) return cudaSuccess;
  frame_embed_kernel<<<row_grid(frames, 8), 256, 0, stream>>>(spatial_x, S, frame_types, pos_table,
                                                              ft_table, n_frame_types, ln_g, ln_b,
                                                              eps, L, frames, out, err_flag, drop, pre_g,
                                                              pre_b, pre_eps);
  return cudaGetLastError();
}

cudaError_t launch_gather_last(const float* x, const long long* lengths, int B, int L, float* out,
                               int* err_flag, cudaStream_t stream) {
  if (B == 0) return cudaSuccess;
  gather_last_kernel<<<row_grid(B, 8), 256, 0, stream>>>(x, lengths, B, L, out, err_flag);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* src_x, const __nv_bfloat16* src_att, int planes,
                               long long src_plane_rows, int stride, const long long* lengths, int L,
                               long long rows, float* dst_x, __nv_bfloat16* dst_att,
                               long long dst_plane_rows, int* err_flag, cudaStream_t stream,
                               const float2* src_stats, float2* dst_stats, const __nv_bfloat16* src_hi,
                               const __nv_bfloat16* src_lo) {
  if (rows == 0) return cudaSuccess;
  gather_rows_kernel<<<row_grid(rows, 8), 256, 0, stream>>>(src_x, src_att, planes, src_plane_rows,
                                                            stride, lengths, L, rows, dst_x, dst_att,
                                                            dst_plane_rows, err_flag, src_stats, dst_stats, src_hi,
                                                            src_lo);
  return cudaGetLastError();
}

cudaError_t launch_build_batch(const BatchStore& st, const long long* video_index,
                               const long long* frame_indices, const long long* num_sampled, int B,
                               int T, int L, int S, double score_threshold, const BatchIds& ids,
                               long long* categories, float* boxes, float* scores,
                               long long* frame_types, long long* lengths, uint8_t* mask_boxes,
                               uint8_t* mask_frames, int* err_flag, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * L * S;
  if (total == 0) return cudaSuccess;
  const long long blocks = (total + 255) / 256;
  build_batch_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      st, video_index, frame_indices, num_sampled, T, L, S, score_threshold, ids, total, categories,
      reinterpret_cast<float4*>(boxes), scores, frame_types, lengths, mask_boxes, mask_frames, err_flag);
  return cudaGetLastError();
}

cudaError_t launch_topk_count(const float* logits, const long long* labels, int rows, int classes,
                              unsigned long long* counters, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  int blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  topk_count_kernel<<<blocks, 256, 0, stream>>>(logits, labels, rows, classes, counters);
  return cudaGetLastError();
}

cudaError_t launch_pack_folded(const float* w, const float* gamma, const float* beta, const float* bias, int n,
                               int k, __nv_bfloat16* wf, float* s_out, float* c_out, cudaStream_t stream,
                               bool head_major, bool split) {
  if (head_major && n != kQkv) return cudaErrorInvalidValue;
  pack_folded_kernel<<<n, 256, 0, stream>>>(w, gamma, beta, bias, k, wf, s_out, c_out, head_major ? 1 : 0,
                                            split ? static_cast<long long>(n) * k : 0);
  return cudaGetLastError();
}

cudaError_t launch_pack_bf16(const float* src, __nv_bfloat16* dst, long long n, int planes,
                             cudaStream_t stream) {
  if (n % 4 != 0) return cudaErrorInvalidValue;
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst),
      planes == 2 ? reinterpret_cast<uint2*>(dst + n) : nullptr, n4);
  return cudaGetLastError();
}

}  // namespace stlt
