// HBM-bound kernels of the STLT training step: the backward passes of the
// row-wise stages (LayerNorm sites, GELU, embeddings, classifier head), the losses of the reference
// Criterion (src/utils/train_inference_utils.py:64-76), the global gradient norm of
// clip_grad_norm_ (src/train.py:129) and the AdamW update (src/train.py:102-104,130).
//
// Two thread mappings are used:
//   * "warp owns a row" (rowops.cuh) wherever a row reduction is needed (LayerNorm backward); the
//     per-column sums that fall out of it (d gamma, d beta, bias gradients) are accumulated in
//     registers over all rows a warp visits, combined through shared memory per block and added to
//     the fp32 gradient with one atomic per column per block;
//   * "thread owns columns, block walks rows" for pure column reductions (bias gradients of the wide
//     bf16 tensors, embedding-table gradients): no shuffles, no atomics until the final flush.
#include "rowops.cuh"

namespace stlt {

namespace {

constexpr int kBwdWarps = 8;

__device__ __forceinline__ void acc_add(RowRegs& a, const RowRegs& b) {
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    a.v[k].x += b.v[k].x;
    a.v[k].y += b.v[k].y;
    a.v[k].z += b.v[k].z;
    a.v[k].w += b.v[k].w;
  }
}

__device__ __forceinline__ RowRegs zero_row() {
  RowRegs r;
#pragma unroll
  for (int k = 0; k < kVec; ++k) r.v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  return r;
}

// mean / rstd of a row exactly as layer_norm_row computes them in the forward pass.
__device__ __forceinline__ void row_stats(const RowRegs& r, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kVec; ++k) s += (r.v[k].x + r.v[k].y) + (r.v[k].z + r.v[k].w);
  mean = warp_sum(s) * (1.0f / kHidden);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const float a = r.v[k].x - mean, b = r.v[k].y - mean, c = r.v[k].z - mean, d = r.v[k].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kHidden) + eps);
}

// LayerNorm backward of one row held by a warp.
//   in : z (pre-norm input), dy (gradient of the output); out: z <- x_hat, dy <- dz
//   dz = rstd * (g*dy - mean(g*dy) - x_hat * mean(g*dy*x_hat));  d gamma += dy * x_hat;  d beta += dy
__device__ __forceinline__ void ln_bwd_row(RowRegs& z, RowRegs& dy, const float* __restrict__ gamma,
                                           float eps, int lane, RowRegs& acc_g, RowRegs& acc_b) {
  float mean, rstd;
  row_stats(z, eps, mean, rstd);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float s1 = 0.f, s2 = 0.f;
  RowRegs gd;
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const float4 g = __ldg(g4 + lane + 32 * k);
    float4& zz = z.v[k];
    const float4 d = dy.v[k];
    zz.x = (zz.x - mean) * rstd;
    zz.y = (zz.y - mean) * rstd;
    zz.z = (zz.z - mean) * rstd;
    zz.w = (zz.w - mean) * rstd;
    acc_g.v[k].x += d.x * zz.x;
    acc_g.v[k].y += d.y * zz.y;
    acc_g.v[k].z += d.z * zz.z;
    acc_g.v[k].w += d.w * zz.w;
    acc_b.v[k].x += d.x;
    acc_b.v[k].y += d.y;
    acc_b.v[k].z += d.z;
    acc_b.v[k].w += d.w;
    gd.v[k] = make_float4(g.x * d.x, g.y * d.y, g.z * d.z, g.w * d.w);
    s1 += (gd.v[k].x + gd.v[k].y) + (gd.v[k].z + gd.v[k].w);
    s2 += (gd.v[k].x * zz.x + gd.v[k].y * zz.y) + (gd.v[k].z * zz.z + gd.v[k].w * zz.w);
  }
  const float m1 = warp_sum(s1) * (1.0f / kHidden);
  const float m2 = warp_sum(s2) * (1.0f / kHidden);
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    dy.v[k].x = rstd * (gd.v[k].x - m1 - z.v[k].x * m2);
    dy.v[k].y = rstd * (gd.v[k].y - m1 - z.v[k].y * m2);
    dy.v[k].z = rstd * (gd.v[k].z - m1 - z.v[k].z * m2);
    dy.v[k].w = rstd * (gd.v[k].w - m1 - z.v[k].w * m2);
  }
}

// Adds the per-warp column accumulators of a block to dst[768] (one atomic per column per block).
// `scratch` is kBwdWarps x 768 floats of shared memory; every thread of the block must call this.
__device__ __forceinline__ void flush_columns(const RowRegs& acc, float* __restrict__ dst,
                                              float* scratch) {
  if (dst == nullptr) return;  // uniform across the block
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kVec; ++k)
    *reinterpret_cast<float4*>(scratch + warp * kHidden + 4 * (lane + 32 * k)) = acc.v[k];
  __syncthreads();
  for (int c = threadIdx.x; c < kHidden; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kBwdWarps; ++w) s += scratch[w * kHidden + c];
    atomicAdd(dst + c, s);
  }
}

__device__ __forceinline__ void store_row_f32(float* base, long long row, const RowRegs& r, int lane) {
  float4* p = reinterpret_cast<float4*>(base + row * kHidden);
#pragma unroll
  for (int k = 0; k < kVec; ++k) p[lane + 32 * k] = r.v[k];
}

__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* base, long long row, const RowRegs& r,
                                               int lane) {
  uint2* p = reinterpret_cast<uint2*>(base + row * kHidden);
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    uint2 h;
    h.x = pack_bf16x2(r.v[k].x, r.v[k].y);
    h.y = pack_bf16x2(r.v[k].z, r.v[k].w);
    p[lane + 32 * k] = h;
  }
}

__device__ __forceinline__ float gelu_erf_grad(float u) {
  // d/du [0.5 u (1 + erf(u / sqrt 2))] = 0.5 (1 + erf(u / sqrt 2)) + u exp(-u^2 / 2) / sqrt(2 pi)
  return 0.5f * (1.0f + erff(u * 0.70710678118654752440f)) +
         u * 0.3989422804014327f * __expf(-0.5f * u * u);
}

// ------------------------------------------------------------------------------------------------
// Backward of x_out = LN(z), z = x + y (the two post-norm residual sites of an encoder layer,
// src/modelling/models.py:46-52 -> nn.TransformerEncoderLayer): dz feeds both the residual branch
// (fp32) and, as bf16, the data/weight-gradient GEMMs of the linear that produced y; its column sum
// is that linear's bias gradient.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBwdWarps * 32, 1)
ln_bwd_kernel(const float* __restrict__ d_a, const float* __restrict__ d_b,
              const float* __restrict__ z, const float* __restrict__ gamma, float eps, long long rows,
              float* __restrict__ dz_out, __nv_bfloat16* __restrict__ dzb_out,
              float* __restrict__ d_gamma, float* __restrict__ d_beta, float* __restrict__ d_bias,
              DropCfg drop) {
  __shared__ __align__(16) float scratch[kBwdWarps * kHidden];
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  RowRegs acc_g = zero_row(), acc_b = zero_row(), acc_z = zero_row();
  for (long long row = warp0; row < rows; row += nwarps) {
    RowRegs dy = load_row(d_a, row, lane);
    if (d_b != nullptr) acc_add(dy, load_row(d_b, row, lane));
    RowRegs zz = load_row(z, row, lane);
    ln_bwd_row(zz, dy, gamma, eps, lane, acc_g, acc_b);
    if (dz_out != nullptr) store_row_f32(dz_out, row, dy, lane);  // residual branch: no dropout
    drop_row(dy, row, lane, drop);  // branch through dropout1 / dropout2 into the producing linear
    acc_add(acc_z, dy);
    if (dzb_out != nullptr) store_row_bf16(dzb_out, row, dy, lane);
  }
  flush_columns(acc_g, d_gamma, scratch);
  flush_columns(acc_b, d_beta, scratch);
  flush_columns(acc_z, d_bias, scratch);
}

// ------------------------------------------------------------------------------------------------
// fp32 column sums of a [rows, n] fp32 tensor (bias gradients of the classifier head).
__global__ void colsum_f32_kernel(const float* __restrict__ x, int rows, int n, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) s += x[static_cast<long long>(r) * n + c];
  atomicAdd(out + c, s);
}

// ------------------------------------------------------------------------------------------------
// Row scatter, the adjoint of gather_rows_kernel (pruned last layer of each stack):
//   dst row = r * stride (stride > 0) or r * L + lengths[r] - 1 (stride == 0)
//   bf16: dst (pre-zeroed) <- src;  f32: dst += src
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const float* __restrict__ src_f, float* __restrict__ dst_f,
                    const __nv_bfloat16* __restrict__ src_b, __nv_bfloat16* __restrict__ dst_b,
                    int stride, const long long* __restrict__ lengths, int L, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    long long dst;
    if (stride > 0) {
      dst = r * stride;
    } else {
      long long len = lengths[r];
      if (len < 1 || len > L) len = 1;
      dst = r * L + (len - 1);
    }
    if (src_f != nullptr) {
      RowRegs a = load_row(dst_f, dst, lane);
      acc_add(a, load_row(src_f, r, lane));
      store_row_f32(dst_f, dst, a, lane);
    }
    if (src_b != nullptr) {
      const uint4* s = reinterpret_cast<const uint4*>(src_b + r * kHidden);
      uint4* d = reinterpret_cast<uint4*>(dst_b + dst * kHidden);
#pragma unroll
      for (int k = 0; k < 3; ++k) d[lane + 32 * k] = __ldg(s + lane + 32 * k);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of FramesEmbeddings.forward (src/modelling/models.py:98-111):
//   f = LN(y_cls + E_pos[l] + E_ft[t]);  d_pre -> d y_cls (written), d E_pos[l] += , d E_ft[t] +=
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBwdWarps * 32, 1)
frame_embed_bwd_kernel(const float* __restrict__ d_a, const float* __restrict__ d_b,
                       const float* __restrict__ cls_x, const long long* __restrict__ frame_types,
                       const float* __restrict__ pos_table, const float* __restrict__ ft_table,
                       int n_frame_types, const float* __restrict__ gamma, float eps, int L,
                       long long frames, float* __restrict__ d_cls, float* __restrict__ d_pos,
                       float* __restrict__ d_ft, float* __restrict__ d_gamma,
                       float* __restrict__ d_beta, DropCfg drop) {
  __shared__ __align__(16) float scratch[kBwdWarps * kHidden];
  // per-block accumulators of the two table gradients: [L + n_frame_types][768]; shared-memory atomics
  // from the 8 warps, one global atomic per touched element per block at the end (the direct global
  // version serialised ~35k adds on each of the few hot rows)
  extern __shared__ __align__(16) float table_acc[];
  const int table_rows = L + n_frame_types;
  for (int i = threadIdx.x; i < table_rows * kHidden; i += blockDim.x) table_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  RowRegs acc_g = zero_row(), acc_b = zero_row();
  for (long long f = warp0; f < frames; f += nwarps) {
    const int l = static_cast<int>(f % L);
    long long ft = frame_types[f];
    if (ft < 0 || ft >= n_frame_types) ft = 0;
    RowRegs zz = load_row(cls_x, f, lane);
    const RowRegs p = load_row(pos_table, l, lane);
    const RowRegs t = load_row(ft_table, ft, lane);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      zz.v[k].x = (zz.v[k].x + p.v[k].x) + t.v[k].x;
      zz.v[k].y = (zz.v[k].y + p.v[k].y) + t.v[k].y;
      zz.v[k].z = (zz.v[k].z + p.v[k].z) + t.v[k].z;
      zz.v[k].w = (zz.v[k].w + p.v[k].w) + t.v[k].w;
    }
    RowRegs dy = load_row(d_a, f, lane);
    if (d_b != nullptr) acc_add(dy, load_row(d_b, f, lane));
    drop_row(dy, f, lane, drop);  // dropout follows the LayerNorm (models.py:110)
    ln_bwd_row(zz, dy, gamma, eps, lane, acc_g, acc_b);
    store_row_f32(d_cls, f, dy, lane);
    float* ap = table_acc + l * kHidden;
    float* at = table_acc + (L + static_cast<int>(ft)) * kHidden;
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      const int c = 4 * (lane + 32 * k);
      atomicAdd(ap + c + 0, dy.v[k].x);
      atomicAdd(ap + c + 1, dy.v[k].y);
      atomicAdd(ap + c + 2, dy.v[k].z);
      atomicAdd(ap + c + 3, dy.v[k].w);
      atomicAdd(at + c + 0, dy.v[k].x);
      atomicAdd(at + c + 1, dy.v[k].y);
      atomicAdd(at + c + 2, dy.v[k].z);
      atomicAdd(at + c + 3, dy.v[k].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < table_rows * kHidden; i += blockDim.x) {
    const float v = table_acc[i];
    if (v == 0.f) continue;
    const int r = i / kHidden, c = i - r * kHidden;
    if (r < L) {
      if (d_pos != nullptr) atomicAdd(d_pos + static_cast<long long>(r) * kHidden + c, v);
    } else if (d_ft != nullptr && r != L) {  // row 0 is nn.Embedding's padding_idx (models.py:91): its gradient stays zero
      atomicAdd(d_ft + static_cast<long long>(r - L) * kHidden + c, v);
    }
  }
  flush_columns(acc_g, d_gamma, scratch);
  flush_columns(acc_b, d_beta, scratch);
}

// ------------------------------------------------------------------------------------------------
// Backward of CategoryBoxEmbeddings.forward (src/modelling/models.py:29-39), part 1: recompute the
// pre-LayerNorm embedding of every token, LayerNorm backward -> d_pre (fp32 [tokens, 768]).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBwdWarps * 32, 1)
embed_ln_bwd_kernel(const float* __restrict__ d_a, const float* __restrict__ d_b,
                    const long long* __restrict__ categories, const float4* __restrict__ boxes,
                    const float* __restrict__ scores, const float* __restrict__ cat_table,
                    int unique_categories, const float* __restrict__ box_w,
                    const float* __restrict__ box_b, const float* __restrict__ score_w,
                    const float* __restrict__ score_b, const float* __restrict__ gamma, float eps,
                    long long tokens, float* __restrict__ d_pre, float* __restrict__ d_gamma,
                    float* __restrict__ d_beta, DropCfg drop) {
  __shared__ __align__(16) float scratch[kBwdWarps * kHidden];
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  const float4* bw4 = reinterpret_cast<const float4*>(box_w);
  RowRegs acc_g = zero_row(), acc_b = zero_row();
  for (long long t = warp0; t < tokens; t += nwarps) {
    long long cat = categories[t];
    if (cat < 0 || cat >= unique_categories) cat = 0;
    const float4 box = __ldg(boxes + t);
    const float score = scores != nullptr ? __ldg(scores + t) : 0.f;
    RowRegs zz = load_row(cat_table, cat, lane);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      const int c = 4 * (lane + 32 * k);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(box_b + c));
      float e[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w = __ldg(bw4 + c + q);
        e[q] += box.x * w.x + box.y * w.y + box.z * w.z + box.w * w.w;
      }
      if (scores != nullptr) {
        const float4 sw = __ldg(reinterpret_cast<const float4*>(score_w + c));
        const float4 sb = __ldg(reinterpret_cast<const float4*>(score_b + c));
        e[0] += score * sw.x + sb.x;
        e[1] += score * sw.y + sb.y;
        e[2] += score * sw.z + sb.z;
        e[3] += score * sw.w + sb.w;
      }
      zz.v[k].x += e[0];
      zz.v[k].y += e[1];
      zz.v[k].z += e[2];
      zz.v[k].w += e[3];
    }
    RowRegs dy = load_row(d_a, t, lane);
    if (d_b != nullptr) acc_add(dy, load_row(d_b, t, lane));
    drop_row(dy, t, lane, drop);  // dropout follows the LayerNorm (models.py:38)
    ln_bwd_row(zz, dy, gamma, eps, lane, acc_g, acc_b);
    store_row_f32(d_pre, t, dy, lane);
  }
  flush_columns(acc_g, d_gamma, scratch);
  flush_columns(acc_b, d_beta, scratch);
}

// Part 2: parameter gradients from d_pre. Thread t of a 192-thread block owns columns 4t..4t+3 and
// walks the block's rows: d E_cat[cat] (shared-memory table, flushed once), d W_box [768, 4],
// d b_box (= d b_score), d W_score [768, 1] accumulate without any intra-block synchronisation.
__global__ void __launch_bounds__(192)
embed_param_grad_kernel(const float* __restrict__ d_pre, const long long* __restrict__ categories,
                        const float4* __restrict__ boxes, const float* __restrict__ scores,
                        int unique_categories, long long tokens, float* __restrict__ d_cat,
                        float* __restrict__ d_box_w, float* __restrict__ d_box_b,
                        float* __restrict__ d_score_w, float* __restrict__ d_score_b) {
  extern __shared__ __align__(16) float cat_acc[];  // [unique_categories][768]
  const int t = threadIdx.x;
  for (int u = 0; u < unique_categories; ++u)
    *reinterpret_cast<float4*>(cat_acc + u * kHidden + 4 * t) = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 wacc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) wacc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f), sacc = bacc;
  const float4* dp4 = reinterpret_cast<const float4*>(d_pre);
  constexpr int kUnroll = 4;
  for (long long r0 = static_cast<long long>(blockIdx.x) * kUnroll; r0 < tokens;
       r0 += static_cast<long long>(gridDim.x) * kUnroll) {
    float4 g[kUnroll];
    float4 box[kUnroll];
    float score[kUnroll];
    long long cat[kUnroll];
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const bool ok = r0 + i < tokens;
      const long long r = ok ? r0 + i : tokens - 1;
      g[i] = ok ? __ldg(dp4 + r * (kHidden / 4) + t) : make_float4(0.f, 0.f, 0.f, 0.f);
      box[i] = __ldg(boxes + r);
      score[i] = scores != nullptr ? __ldg(scores + r) : 0.f;
      cat[i] = categories[r];
      if (cat[i] < 0 || cat[i] >= unique_categories) cat[i] = 0;
    }
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      float4* ca = reinterpret_cast<float4*>(cat_acc + cat[i] * kHidden + 4 * t);
      float4 c = *ca;
      c.x += g[i].x;
      c.y += g[i].y;
      c.z += g[i].z;
      c.w += g[i].w;
      *ca = c;
      const float gq[4] = {g[i].x, g[i].y, g[i].z, g[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // d W_box[4t + q][0..3] += g[q] * box
        wacc[q].x += gq[q] * box[i].x;
        wacc[q].y += gq[q] * box[i].y;
        wacc[q].z += gq[q] * box[i].z;
        wacc[q].w += gq[q] * box[i].w;
      }
      bacc.x += g[i].x;
      bacc.y += g[i].y;
      bacc.z += g[i].z;
      bacc.w += g[i].w;
      sacc.x += g[i].x * score[i];
      sacc.y += g[i].y * score[i];
      sacc.z += g[i].z * score[i];
      sacc.w += g[i].w * score[i];
    }
  }
  const int c = 4 * t;
  if (d_cat != nullptr)
    for (int u = 1; u < unique_categories; ++u) {  // row 0 is nn.Embedding's padding_idx (models.py:19-23): no gradient
      const float4 v = *reinterpret_cast<const float4*>(cat_acc + u * kHidden + c);
      if (v.x != 0.f) atomicAdd(d_cat + u * kHidden + c + 0, v.x);
      if (v.y != 0.f) atomicAdd(d_cat + u * kHidden + c + 1, v.y);
      if (v.z != 0.f) atomicAdd(d_cat + u * kHidden + c + 2, v.z);
      if (v.w != 0.f) atomicAdd(d_cat + u * kHidden + c + 3, v.w);
    }
  if (d_box_w != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      atomicAdd(d_box_w + (c + q) * 4 + 0, wacc[q].x);
      atomicAdd(d_box_w + (c + q) * 4 + 1, wacc[q].y);
      atomicAdd(d_box_w + (c + q) * 4 + 2, wacc[q].z);
      atomicAdd(d_box_w + (c + q) * 4 + 3, wacc[q].w);
    }
  }
  const float bq[4] = {bacc.x, bacc.y, bacc.z, bacc.w};
  const float sq[4] = {sacc.x, sacc.y, sacc.z, sacc.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (d_box_b != nullptr) atomicAdd(d_box_b + c + q, bq[q]);
    if (scores != nullptr) {
      if (d_score_b != nullptr) atomicAdd(d_score_b + c + q, bq[q]);
      if (d_score_w != nullptr) atomicAdd(d_score_w + c + q, sq[q]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Classifier head (src/modelling/models.py:155-163): h2 = LN(gelu(h1)) and its backward.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gelu_ln_kernel(const float* __restrict__ h1, const float* __restrict__ g, const float* __restrict__ b,
               float eps, long long rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long row = warp0; row < rows; row += nwarps) {
    RowRegs r = load_row(h1, row, lane);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      r.v[k].x = gelu_erf(r.v[k].x);
      r.v[k].y = gelu_erf(r.v[k].y);
      r.v[k].z = gelu_erf(r.v[k].z);
      r.v[k].w = gelu_erf(r.v[k].w);
    }
    layer_norm_row(r, g, b, eps, lane);
    store_row_f32(out, row, r, lane);
  }
}

__global__ void __launch_bounds__(kBwdWarps * 32, 1)
gelu_ln_bwd_kernel(const float* __restrict__ d_h2, const float* __restrict__ h1,
                   const float* __restrict__ gamma, float eps, long long rows,
                   float* __restrict__ d_h1, float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  __shared__ __align__(16) float scratch[kBwdWarps * kHidden];
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  RowRegs acc_g = zero_row(), acc_b = zero_row();
  for (long long row = warp0; row < rows; row += nwarps) {
    const RowRegs u = load_row(h1, row, lane);
    RowRegs zz;
#pragma unroll
    for (int k = 0; k < kVec; ++k)
      zz.v[k] = make_float4(gelu_erf(u.v[k].x), gelu_erf(u.v[k].y), gelu_erf(u.v[k].z), gelu_erf(u.v[k].w));
    RowRegs dy = load_row(d_h2, row, lane);
    ln_bwd_row(zz, dy, gamma, eps, lane, acc_g, acc_b);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      dy.v[k].x *= gelu_erf_grad(u.v[k].x);
      dy.v[k].y *= gelu_erf_grad(u.v[k].y);
      dy.v[k].z *= gelu_erf_grad(u.v[k].z);
      dy.v[k].w *= gelu_erf_grad(u.v[k].w);
    }
    store_row_f32(d_h1, row, dy, lane);
  }
  flush_columns(acc_g, d_gamma, scratch);
  flush_columns(acc_b, d_beta, scratch);
}

// ------------------------------------------------------------------------------------------------
// Losses of the reference Criterion (src/utils/train_inference_utils.py:64-76), mean reduction, and
// their gradients w.r.t. the logits. One warp per row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cross_entropy_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int rows,
                     int classes, float grad_scale, float* __restrict__ loss, float* __restrict__ d_logits) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float inv_rows = 1.0f / static_cast<float>(rows);
  float local = 0.f;
  for (int r = warp0; r < rows; r += nwarps) {
    const float* row = logits + static_cast<long long>(r) * classes;
    float m = -INFINITY;
    for (int c = lane; c < classes; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < classes; c += 32) s += expf(row[c] - m);
    s = warp_sum(s);
    const long long lab = labels[r];
    const float lse = m + logf(s);
    if (lab >= 0 && lab < classes) local += lse - row[lab];
    if (d_logits != nullptr) {
      float* drow = d_logits + static_cast<long long>(r) * classes;
      const float inv = 1.0f / s;
      for (int c = lane; c < classes; c += 32) {
        float p = expf(row[c] - m) * inv;
        if (c == lab) p -= 1.0f;
        drow[c] = p * inv_rows * grad_scale;
      }
    }
  }
  if (lane == 0 && loss != nullptr) atomicAdd(loss, local * inv_rows);
}

__global__ void __launch_bounds__(256)
bce_logits_kernel(const float* __restrict__ logits, const float* __restrict__ targets, long long n,
                  float grad_scale, float* __restrict__ loss, float* __restrict__ d_logits) {
  const float inv_n = 1.0f / static_cast<float>(n);
  float local = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += gridDim.x * static_cast<long long>(blockDim.x)) {
    const float x = logits[i], y = targets[i];
    // max(x, 0) - x*y + log(1 + exp(-|x|))  (nn.BCEWithLogitsLoss)
    local += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
    if (d_logits != nullptr) d_logits[i] = (1.0f / (1.0f + expf(-x)) - y) * inv_n * grad_scale;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0 && loss != nullptr) atomicAdd(loss, local * inv_n);
}

// ------------------------------------------------------------------------------------------------
// clip_grad_norm_ (src/train.py:129) + AdamW (torch.optim.AdamW as configured at src/train.py:102-104)
// on flat fp32 buffers.
// ------------------------------------------------------------------------------------------------
// Deterministic two-stage reduction: data-parallel ranks must derive bit-identical clip coefficients from
// their (identical, all-reduced) gradients, otherwise their parameters drift apart.
__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float4* __restrict__ g, long long n4, const float* __restrict__ tail, int n_tail,
                     float* __restrict__ partials) {
  float s = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += gridDim.x * static_cast<long long>(blockDim.x)) {
    const float4 v = __ldg(g + i);
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < n_tail) s += tail[threadIdx.x] * tail[threadIdx.x];
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256)
sumsq_final_kernel(const float* __restrict__ partials, int count, float* __restrict__ out) {
  __shared__ double part[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < count; i += 256) s += static_cast<double>(partials[i]);
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 256; ++i) t += part[i];
    out[0] = static_cast<float>(t);
  }
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
             float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
             float weight_decay, float bias_c1, float bias_c2_sqrt, const float* __restrict__ sumsq,
             float max_norm) {
  float coef = 1.0f;
  if (sumsq != nullptr) {  // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float c = max_norm / (sqrtf(__ldg(sumsq)) + 1e-6f);
    coef = c < 1.0f ? c : 1.0f;
  }
  const float step_size = lr / bias_c1;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += gridDim.x * static_cast<long long>(blockDim.x)) {
    const float grad = g[i] * coef;
    float w = p[i];
    w *= 1.0f - lr * weight_decay;
    const float mi = beta1 * m[i] + (1.0f - beta1) * grad;
    const float vi = beta2 * v[i] + (1.0f - beta2) * grad * grad;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bias_c2_sqrt + eps;
    p[i] = w - step_size * (mi / denom);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_ln_bwd(const float* d_a, const float* d_b, const float* z, const float* gamma,
                          float eps, long long rows, float* dz_out, __nv_bfloat16* dzb_out,
                          float* d_gamma, float* d_beta, float* d_bias, cudaStream_t stream,
                          DropCfg drop) {
  if (rows == 0) return cudaSuccess;
  ln_bwd_kernel<<<row_grid(rows, kBwdWarps, 2), kBwdWarps * 32, 0, stream>>>(
      d_a, d_b, z, gamma, eps, rows, dz_out, dzb_out, d_gamma, d_beta, d_bias, drop);
  return cudaGetLastError();
}

cudaError_t launch_colsum_f32(const float* x, int rows, int n, float* out, cudaStream_t stream) {
  if (rows == 0 || n == 0) return cudaSuccess;
  dim3 grid((n + 127) / 128, rows < 64 ? rows : 64);
  colsum_f32_kernel<<<grid, 128, 0, stream>>>(x, rows, n, out);
  return cudaGetLastError();
}

cudaError_t launch_scatter_rows(const float* src_f, float* dst_f, const __nv_bfloat16* src_b,
                                __nv_bfloat16* dst_b, int stride, const long long* lengths, int L,
                                long long rows, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  scatter_rows_kernel<<<row_grid(rows, 8), 256, 0, stream>>>(src_f, dst_f, src_b, dst_b, stride, lengths,
                                                             L, rows);
  return cudaGetLastError();
}

cudaError_t launch_frame_embed_bwd(const float* d_a, const float* d_b, const float* cls_x,
                                   const long long* frame_types, const float* pos_table,
                                   const float* ft_table, int n_frame_types, const float* gamma,
                                   float eps, int B, int L, float* d_cls, float* d_pos, float* d_ft,
                                   float* d_gamma, float* d_beta, cudaStream_t stream, DropCfg drop) {
  const long long frames = static_cast<long long>(B) * L;
  if (frames == 0) return cudaSuccess;
  const int smem = (L + n_frame_types) * kHidden * static_cast<int>(sizeof(float));
  cudaError_t e = cudaFuncSetAttribute(frame_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  frame_embed_bwd_kernel<<<row_grid(frames, kBwdWarps, 1), kBwdWarps * 32, smem, stream>>>(
      d_a, d_b, cls_x, frame_types, pos_table, ft_table, n_frame_types, gamma, eps, L, frames, d_cls,
      d_pos, d_ft, d_gamma, d_beta, drop);
  return cudaGetLastError();
}

cudaError_t launch_embed_bwd(const float* d_a, const float* d_b, const long long* categories,
                             const float* boxes, const float* scores, const float* cat_table,
                             int unique_categories, const float* box_w, const float* box_b,
                             const float* score_w, const float* score_b, const float* gamma, float eps,
                             long long tokens, float* d_pre, float* d_cat, float* d_box_w,
                             float* d_box_b, float* d_score_w, float* d_score_b, float* d_gamma,
                             float* d_beta, cudaStream_t stream, DropCfg drop) {
  if (tokens == 0) return cudaSuccess;
  const float4* boxes4 = reinterpret_cast<const float4*>(boxes);
  embed_ln_bwd_kernel<<<row_grid(tokens, kBwdWarps, 2), kBwdWarps * 32, 0, stream>>>(
      d_a, d_b, categories, boxes4, scores, cat_table, unique_categories, box_w, box_b, score_w,
      score_b, gamma, eps, tokens, d_pre, d_gamma, d_beta, drop);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int smem = unique_categories * kHidden * static_cast<int>(sizeof(float));
  if (smem > 200 * 1024) return cudaErrorInvalidValue;  // > 66 categories: table does not fit
  e = cudaFuncSetAttribute(embed_param_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  long long blocks = (tokens + 3) / 4;
  if (blocks > 148 * 2) blocks = 148 * 2;
  embed_param_grad_kernel<<<static_cast<unsigned>(blocks), 192, smem, stream>>>(
      d_pre, categories, boxes4, scores, unique_categories, tokens, d_cat, d_box_w, d_box_b, d_score_w,
      d_score_b);
  return cudaGetLastError();
}

cudaError_t launch_gelu_ln(const float* h1, const float* g, const float* b, float eps, long long rows,
                           float* out, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  gelu_ln_kernel<<<row_grid(rows, 8), 256, 0, stream>>>(h1, g, b, eps, rows, out);
  return cudaGetLastError();
}

cudaError_t launch_gelu_ln_bwd(const float* d_h2, const float* h1, const float* gamma, float eps,
                               long long rows, float* d_h1, float* d_gamma, float* d_beta,
                               cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  gelu_ln_bwd_kernel<<<row_grid(rows, kBwdWarps, 2), kBwdWarps * 32, 0, stream>>>(
      d_h2, h1, gamma, eps, rows, d_h1, d_gamma, d_beta);
  return cudaGetLastError();
}

cudaError_t launch_cross_entropy(const float* logits, const long long* labels, int rows, int classes,
                                 float grad_scale, float* loss, float* d_logits, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  int blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cross_entropy_kernel<<<blocks, 256, 0, stream>>>(logits, labels, rows, classes, grad_scale, loss, d_logits);
  return cudaGetLastError();
}

cudaError_t launch_bce_logits(const float* logits, const float* targets, long long n, float grad_scale,
                              float* loss, float* d_logits, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  bce_logits_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(logits, targets, n, grad_scale, loss,
                                                                        d_logits);
  return cudaGetLastError();
}

cudaError_t launch_sumsq(const float* g, long long n, float* out, float* scratch, int scratch_len,
                         cudaStream_t stream) {
  if (n == 0) return cudaMemsetAsync(out, 0, sizeof(float), stream);
  if ((reinterpret_cast<uintptr_t>(g) & 15) != 0 || scratch_len < 1) return cudaErrorInvalidValue;
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks > scratch_len) blocks = scratch_len;
  if (blocks < 1) blocks = 1;
  sumsq_partial_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(g), n4, g + n4 * 4, static_cast<int>(n - n4 * 4), scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sumsq_final_kernel<<<1, 256, 0, stream>>>(scratch, static_cast<int>(blocks), out);
  return cudaGetLastError();
}

cudaError_t launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float bias_c1,
                         float bias_c2_sqrt, const float* sumsq, float max_norm, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adamw_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps,
                                                                   weight_decay, bias_c1, bias_c2_sqrt,
                                                                   sumsq, max_norm);
  return cudaGetLastError();
}

}  // namespace stlt
