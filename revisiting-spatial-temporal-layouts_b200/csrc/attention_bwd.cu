// Backward of K3 (masked multi-head self-attention over the short STLT sequences; forward in
// attention_mma.cu / attention.cu, reference: the SDPA inside nn.MultiheadAttention configured at
// src/modelling/models.py:46-55,118-128). Training step, SURVEY.md 8(f) rank 1.
//
//   S = Q K^T / 8 (+ masks),  P = softmax(S),  O = P V
//   dV = P^T dO,  dP = dO V^T,  dS = P * (dP - rowsum(dP * P)),  dQ = dS K / 8,  dK = dS^T Q / 8
//
// The probabilities are recomputed from the saved bf16 QKV (nothing but QKV and the context is kept
// by the forward pass). One warp owns one (group of G = floor(32 / T) whole sequences, head): lane r
// owns token r of the group during the score phase (its K and V rows live in registers, Q / K / dO
// rows in warp-private shared memory), and a pair of features during the three output products.
// Masked keys have P = 0, hence dS = 0: key-padding and causal masks need no special handling.
#include "common.cuh"
#include "kernels.h"

namespace stlt {

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kPitch = 68;  // floats per staged row (16 B aligned, sequences spread over banks)

__device__ __forceinline__ void load_bf16_row64(const __nv_bfloat16* p, float (&dst)[64]) {
  const uint4* p4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 v = __ldg(p4 + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      dst[8 * i + 2 * j] = f.x;
      dst[8 * i + 2 * j + 1] = f.y;
    }
  }
}

__device__ __forceinline__ void stage_row(float* dst, const float (&src)[64]) {
#pragma unroll
  for (int i = 0; i < 16; ++i)
    *reinterpret_cast<float4*>(dst + 4 * i) =
        make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_ctx,
                     const long long* __restrict__ mask_src, long long num_seqs, int T, int G,
                     int causal, __nv_bfloat16* __restrict__ d_qkv, long long num_items, DropCfg drop) {
  extern __shared__ __align__(16) float smem_f[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int R = G * T;
  const int PT = T + 1;
  const int per_warp = (3 * R * kPitch + 2 * R * PT + 3) & ~3;
  float* Qs = smem_f + warp * per_warp;
  float* Ks = Qs + R * kPitch;
  float* Ds = Ks + R * kPitch;   // dO rows
  float* Ps = Ds + R * kPitch;   // probabilities [R][PT]
  float* Ss = Ps + R * PT;       // dS (already scaled by 1/8) [R][PT]

  const long long total_tokens = num_seqs * T;
  const long long gwarp = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + warp;
  const long long nwarps = gridDim.x * static_cast<long long>(kWarpsPerBlock);

  for (long long item = gwarp; item < num_items; item += nwarps) {
    const long long grp = item / kHeads;
    const int head = static_cast<int>(item - grp * kHeads);
    const long long base = grp * R;
    const long long remaining = total_tokens - base;
    const int nrows = remaining < R ? static_cast<int>(remaining) : R;
    const bool active = lane < nrows;

    float kreg[64], vreg[64];
    bool key_masked = true;
    if (active) {
      const long long tok = base + lane;
      const __nv_bfloat16* row = qkv + tok * kQkv + head * kHeadDim;
      load_bf16_row64(row, vreg);  // Q (staged, then the registers are reused)
      stage_row(Qs + lane * kPitch, vreg);
      load_bf16_row64(d_ctx + tok * kHidden + head * kHeadDim, vreg);
      stage_row(Ds + lane * kPitch, vreg);
      load_bf16_row64(row + kHidden, kreg);
      stage_row(Ks + lane * kPitch, kreg);
      load_bf16_row64(row + 2 * kHidden, vreg);
      key_masked = (mask_src[tok] == 0);
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) kreg[i] = vreg[i] = 0.f;
    }
    __syncwarp();

    const int g = active ? lane / T : 0;
    const int j = lane - g * T;
    const int seq_lane0 = g * T;
    for (int i = 0; i < T; ++i) {
      const float4* q4 = reinterpret_cast<const float4*>(Qs + (seq_lane0 + i) * kPitch);
      const float4* d4 = reinterpret_cast<const float4*>(Ds + (seq_lane0 + i) * kPitch);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 q = q4[d];
        const float4 o = d4[d];
        s0 = fmaf(q.x, kreg[4 * d + 0], s0);
        s1 = fmaf(q.y, kreg[4 * d + 1], s1);
        s2 = fmaf(q.z, kreg[4 * d + 2], s2);
        s3 = fmaf(q.w, kreg[4 * d + 3], s3);
        p0 = fmaf(o.x, vreg[4 * d + 0], p0);
        p1 = fmaf(o.y, vreg[4 * d + 1], p1);
        p2 = fmaf(o.z, vreg[4 * d + 2], p2);
        p3 = fmaf(o.w, vreg[4 * d + 3], p3);
      }
      const bool masked = !active || key_masked || (causal && j > i);
      const float s = masked ? -INFINITY : ((s0 + s1) + (s2 + s3)) * 0.125f;
      float m = -INFINITY;
      for (int jj = 0; jj < T; ++jj) m = fmaxf(m, __shfl_sync(0xffffffffu, s, seq_lane0 + jj));
      const float e = masked ? 0.f : expf(s - m);
      float sum = 0.f;
      for (int jj = 0; jj < T; ++jj) sum += __shfl_sync(0xffffffffu, e, seq_lane0 + jj);
      const float p = e / sum;
      // dropout on the probabilities: O = (mask * P / (1-p)) V, so dV uses the dropped P and the
      // gradient w.r.t. P is the masked, scaled dO V^T
      float keep = 1.0f;
      if (drop.thr16 != 0 && !masked) {
        const unsigned long long el =
            (static_cast<unsigned long long>(base + seq_lane0 + i) * kHeads + head) * 32ull + static_cast<unsigned>(j);
        keep = drop_mul(drop_bits(drop.key, el >> 1), static_cast<int>(el & 1), drop);
      }
      const float dp = masked ? 0.f : ((p0 + p1) + (p2 + p3)) * keep;
      const float pd = p * dp;
      float dsum = 0.f;
      for (int jj = 0; jj < T; ++jj) dsum += __shfl_sync(0xffffffffu, pd, seq_lane0 + jj);
      if (active) {
        Ps[(seq_lane0 + i) * PT + j] = p * keep;
        Ss[(seq_lane0 + i) * PT + j] = p * (dp - dsum) * 0.125f;
      }
    }
    __syncwarp();

    // outputs: lane owns features 2*lane, 2*lane+1 of every row
    for (int r = 0; r < nrows; ++r) {
      const int s0row = (r / T) * T;
      const int jr = r - s0row;
      float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
      for (int t = 0; t < T; ++t) {
        const float ds_q = Ss[r * PT + t];             // dS[r][t]   -> dQ_r += dS * K_t
        const float ds_k = Ss[(s0row + t) * PT + jr];  // dS[t][r]   -> dK_r += dS * Q_t
        const float p_v = Ps[(s0row + t) * PT + jr];   // P[t][r]    -> dV_r += P * dO_t
        const float2 kk = *reinterpret_cast<const float2*>(Ks + (s0row + t) * kPitch + 2 * lane);
        const float2 qq = *reinterpret_cast<const float2*>(Qs + (s0row + t) * kPitch + 2 * lane);
        const float2 oo = *reinterpret_cast<const float2*>(Ds + (s0row + t) * kPitch + 2 * lane);
        q0 = fmaf(ds_q, kk.x, q0);
        q1 = fmaf(ds_q, kk.y, q1);
        k0 = fmaf(ds_k, qq.x, k0);
        k1 = fmaf(ds_k, qq.y, k1);
        v0 = fmaf(p_v, oo.x, v0);
        v1 = fmaf(p_v, oo.y, v1);
      }
      const long long off = (base + r) * kQkv + head * kHeadDim + 2 * lane;
      *reinterpret_cast<uint32_t*>(d_qkv + off) = pack_bf16x2(q0, q1);
      *reinterpret_cast<uint32_t*>(d_qkv + off + kHidden) = pack_bf16x2(k0, k1);
      *reinterpret_cast<uint32_t*>(d_qkv + off + 2 * kHidden) = pack_bf16x2(v0, v1);
    }
    __syncwarp();
  }
}

}  // namespace

cudaError_t launch_attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* d_ctx,
                                 const long long* mask_src, long long num_seqs, int T, bool causal,
                                 __nv_bfloat16* d_qkv, cudaStream_t stream, DropCfg drop) {
  if (T < 1 || T > 32) return cudaErrorInvalidValue;
  if (num_seqs == 0) return cudaSuccess;
  const int G = 32 / T;
  const int R = G * T;
  const long long groups = (num_seqs + G - 1) / G;
  const long long items = groups * kHeads;
  const int per_warp = (3 * R * kPitch + 2 * R * (T + 1) + 3) & ~3;
  const int smem = per_warp * kWarpsPerBlock * static_cast<int>(sizeof(float));
  long long blocks = (items + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = 148LL * 16;
  if (blocks > cap) blocks = cap;
  cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  attention_bwd_kernel<<<static_cast<unsigned>(blocks), kWarpsPerBlock * 32, smem, stream>>>(
      qkv, d_ctx, mask_src, num_seqs, T, G, causal ? 1 : 0, d_qkv, items, drop);
  return cudaGetLastError();
}

}  // namespace stlt
