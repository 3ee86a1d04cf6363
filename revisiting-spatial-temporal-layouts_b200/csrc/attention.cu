// K3: masked multi-head self-attention over the short STLT sequences (S object slots per frame in
// the spatial encoder, L frames per video in the temporal encoder). Replaces the SDPA inside
// nn.MultiheadAttention as configured at src/modelling/models.py:46-55,118-128 with
//   spatial : key_padding_mask = (categories == 0)                  (models.py:66-71)
//   temporal: causal mask (src/utils/model_utils.py:4-7) OR key_padding_mask = (frame_types == 0)
//             (models.py:142-150)
//
// One warp handles one (group of G = 32 / T consecutive sequences, head). Lane r owns token r of
// the group: its K row (64 values) stays in registers, its Q and V rows go to warp-private shared
// memory. Scores for query i are computed with lane = key, the softmax max / sum are warp-shuffle
// reductions over the T lanes of a sequence, and P*V is evaluated with lane = feature pair.
// Nothing but the final bf16 context rows leaves the SM.
#include "common.cuh"
#include "kernels.h"

namespace stlt {

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kQPitch = 68;  // floats; 272 B rows keep 16 B alignment and spread sequences over banks
constexpr int kVPitch = 64;

__device__ __forceinline__ void load_slice(const float* p, float (&dst)[64]) {
  const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 v = __ldg(p4 + i);
    dst[4 * i + 0] = v.x;
    dst[4 * i + 1] = v.y;
    dst[4 * i + 2] = v.z;
    dst[4 * i + 3] = v.w;
  }
}

template <typename TIn>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
attention_kernel(const TIn* __restrict__ qkv, const long long* __restrict__ mask_src,
                 long long num_seqs, int T, int G, int causal, ActOut out, long long num_items) {
  extern __shared__ __align__(16) float smem_f[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int R = G * T;       // token rows per group (<= 32)
  const int PT = T + 1;      // pitch of the probability rows
  const int per_warp = (R * kQPitch + R * kVPitch + R * PT + 3) & ~3;  // keep float4 alignment
  float* Qs = smem_f + warp * per_warp;
  float* Vs = Qs + R * kQPitch;
  float* Ps = Vs + R * kVPitch;

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

    float kreg[64];
    bool key_masked = true;
    if (active) {
      const long long tok = base + lane;
      const TIn* row = qkv + tok * kQkv + head * kHeadDim;
      float tmp[64];
      load_slice(row, tmp);  // Q
#pragma unroll
      for (int i = 0; i < 16; ++i)
        *reinterpret_cast<float4*>(Qs + lane * kQPitch + 4 * i) =
            make_float4(tmp[4 * i], tmp[4 * i + 1], tmp[4 * i + 2], tmp[4 * i + 3]);
      load_slice(row + 2 * kHidden, tmp);  // V
#pragma unroll
      for (int i = 0; i < 16; ++i)
        *reinterpret_cast<float4*>(Vs + lane * kVPitch + 4 * i) =
            make_float4(tmp[4 * i], tmp[4 * i + 1], tmp[4 * i + 2], tmp[4 * i + 3]);
      load_slice(row + kHidden, kreg);  // K
      key_masked = (mask_src[tok] == 0);
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) kreg[i] = 0.f;
    }
    __syncwarp();

    const int g = active ? lane / T : 0;
    const int j = lane - g * T;
    const int seq_lane0 = g * T;  // first lane of this lane's sequence
    for (int i = 0; i < T; ++i) {
      // score(i, j) = q_i . k_j / sqrt(64)
      const float4* q4 = reinterpret_cast<const float4*>(Qs + (seq_lane0 + i) * kQPitch);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 q = q4[d];
        s0 = fmaf(q.x, kreg[4 * d + 0], s0);
        s1 = fmaf(q.y, kreg[4 * d + 1], s1);
        s2 = fmaf(q.z, kreg[4 * d + 2], s2);
        s3 = fmaf(q.w, kreg[4 * d + 3], s3);
      }
      const bool masked = !active || key_masked || (causal && j > i);
      const float s = masked ? -INFINITY : ((s0 + s1) + (s2 + s3)) * 0.125f;
      // warp-shuffle softmax over the T lanes of this sequence
      float m = -INFINITY;
      for (int jj = 0; jj < T; ++jj) m = fmaxf(m, __shfl_sync(0xffffffffu, s, seq_lane0 + jj));
      const float e = masked ? 0.f : expf(s - m);
      float sum = 0.f;
      for (int jj = 0; jj < T; ++jj) sum += __shfl_sync(0xffffffffu, e, seq_lane0 + jj);
      if (active) Ps[(seq_lane0 + i) * PT + j] = e / sum;
    }
    __syncwarp();

    // context rows: O[r, :] = sum_j P[r, j] * V[seq(r) + j, :], lane owns features 2*lane, 2*lane+1
    for (int r = 0; r < nrows; ++r) {
      const int s0row = (r / T) * T;
      const float* prow = Ps + r * PT;
      float o0 = 0.f, o1 = 0.f;
      for (int jj = 0; jj < T; ++jj) {
        const float p = prow[jj];
        const float2 v = *reinterpret_cast<const float2*>(Vs + (s0row + jj) * kVPitch + 2 * lane);
        o0 = fmaf(p, v.x, o0);
        o1 = fmaf(p, v.y, o1);
      }
      const long long tok = base + r;
      const long long off = tok * kHidden + head * kHeadDim + 2 * lane;
      *reinterpret_cast<uint32_t*>(out.xb + off) = pack_bf16x2(o0, o1);
      if (out.planes == 2)
        *reinterpret_cast<uint32_t*>(out.xb + out.plane_rows * kHidden + off) =
            pack_bf16x2(bf16_residual(o0), bf16_residual(o1));
    }
    __syncwarp();
  }
}

}  // namespace

cudaError_t launch_attention(const void* qkv, bool qkv_is_bf16, const long long* mask_src,
                             long long num_seqs, int T, bool causal, ActOut out,
                             cudaStream_t stream) {
  if (T < 1 || T > 256) return cudaErrorInvalidValue;
  if (num_seqs == 0) return cudaSuccess;
  if (T > 64) {
    // 65..256 tokens per sequence: one CTA per (sequence, head), online softmax over 64-key blocks (attention_long.cu)
    if (!qkv_is_bf16) return cudaErrorInvalidValue;
    return launch_attention_long(static_cast<const __nv_bfloat16*>(qkv), out.planes, out.plane_rows, mask_src, num_seqs,
                                 T, causal, out.xb, out.plane_rows, stream);
  }
  if (T > 32) {
    // 33..64 tokens per sequence: one warp per (sequence, head) on the 48 / 64-row tiles of attention_cross.cu
    if (!qkv_is_bf16) return cudaErrorInvalidValue;
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(qkv);
    return launch_attention_cross(p, kQkv, 0, p, kQkv, kHidden, 2 * kHidden, mask_src, num_seqs, T, T, causal, out.xb,
                                  stream, out.planes, out.plane_rows, out.plane_rows, out.plane_rows);
  }
  const int G = 32 / T;
  const int R = G * T;
  const long long groups = (num_seqs + G - 1) / G;
  const long long items = groups * kHeads;
  const int per_warp = (R * kQPitch + R * kVPitch + R * (T + 1) + 3) & ~3;
  const int smem = per_warp * kWarpsPerBlock * static_cast<int>(sizeof(float));
  long long blocks = (items + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = 148LL * 32;
  if (blocks > cap) blocks = cap;
  cudaError_t e;
  if (qkv_is_bf16) {
    // bf16 QKV (1 plane) or hi/lo bf16 planes (2 planes, rows out.plane_rows apart): tensor-core tiles
    return launch_attention_mma(static_cast<const __nv_bfloat16*>(qkv), out.planes, out.plane_rows,
                                mask_src, num_seqs, T, causal, out.xb, out.plane_rows, stream);
  } else {
    auto kern = attention_kernel<float>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kern<<<static_cast<unsigned>(blocks), kWarpsPerBlock * 32, smem, stream>>>(
        static_cast<const float*>(qkv), mask_src, num_seqs, T, G, causal ? 1 : 0, out, items);
  }
  return cudaGetLastError();
}

}  // namespace stlt
