// Row-wise helpers shared by the HBM-bound kernels: one warp owns one 768-wide row; lane l holds
// columns 4*l + 128*k + {0..3}, k = 0..5, so every global access is a coalesced 16-byte vector.
#pragma once

#include "common.cuh"
#include "kernels.h"

namespace stlt {

constexpr int kVec = kHidden / 128;  // 6 float4 per lane

struct RowRegs {
  float4 v[kVec];
};

__device__ __forceinline__ const float4* row4(const float* base, long long row) {
  return reinterpret_cast<const float4*>(base + row * kHidden);
}

__device__ __forceinline__ RowRegs load_row(const float* base, long long row, int lane) {
  RowRegs r;
  const float4* p = row4(base, row);
#pragma unroll
  for (int k = 0; k < kVec; ++k) r.v[k] = __ldg(p + lane + 32 * k);
  return r;
}

// Row of a residual stream stored as two bf16 planes (z = hi + lo, see GEMM_OUT_HILO).
__device__ __forceinline__ RowRegs load_row_hilo(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long row, int lane) {
  RowRegs r;
  const uint2* ph = reinterpret_cast<const uint2*>(hi + row * kHidden);
  const uint2* pl = reinterpret_cast<const uint2*>(lo + row * kHidden);
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const uint2 h = __ldg(ph + lane + 32 * k), l = __ldg(pl + lane + 32 * k);
    const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.x));
    const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.y));
    const float2 l0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.x));
    const float2 l1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.y));
    r.v[k] = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
  }
  return r;
}

// LayerNorm over the 768 features held by one warp (biased variance, two-pass in registers).
__device__ __forceinline__ void layer_norm_row(RowRegs& r, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps, int lane) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kVec; ++k) s += (r.v[k].x + r.v[k].y) + (r.v[k].z + r.v[k].w);
  const float mean = warp_sum(s) * (1.0f / kHidden);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const float a = r.v[k].x - mean, b = r.v[k].y - mean, c = r.v[k].z - mean, d = r.v[k].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float var = warp_sum(q) * (1.0f / kHidden);
  const float rstd = 1.0f / sqrtf(var + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const float4 g = __ldg(g4 + lane + 32 * k);
    const float4 b = __ldg(b4 + lane + 32 * k);
    r.v[k].x = (r.v[k].x - mean) * rstd * g.x + b.x;
    r.v[k].y = (r.v[k].y - mean) * rstd * g.y + b.y;
    r.v[k].z = (r.v[k].z - mean) * rstd * g.z + b.z;
    r.v[k].w = (r.v[k].w - mean) * rstd * g.w + b.w;
  }
}

__device__ __forceinline__ void store_act(const ActOut& out, long long row, const RowRegs& r,
                                          int lane) {
  if (out.x != nullptr) {
    float4* p = reinterpret_cast<float4*>(out.x + row * kHidden);
#pragma unroll
    for (int k = 0; k < kVec; ++k) p[lane + 32 * k] = r.v[k];
  }
  if (out.xb != nullptr) {
    uint2* hi = reinterpret_cast<uint2*>(out.xb + row * kHidden);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      uint2 h;
      h.x = pack_bf16x2(r.v[k].x, r.v[k].y);
      h.y = pack_bf16x2(r.v[k].z, r.v[k].w);
      hi[lane + 32 * k] = h;
    }
    if (out.planes == 2) {
      uint2* lo = reinterpret_cast<uint2*>(out.xb + (out.plane_rows + row) * kHidden);
#pragma unroll
      for (int k = 0; k < kVec; ++k) {
        uint2 l;
        l.x = pack_bf16x2(bf16_residual(r.v[k].x), bf16_residual(r.v[k].y));
        l.y = pack_bf16x2(bf16_residual(r.v[k].z), bf16_residual(r.v[k].w));
        lo[lane + 32 * k] = l;
      }
    }
  }
}

// Dropout of one row held by a warp (element index = row * 768 + column).
__device__ __forceinline__ void drop_row(RowRegs& r, long long row, int lane, const DropCfg& d) {
  if (d.thr16 == 0) return;
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const unsigned long long pair =
        (static_cast<unsigned long long>(row) * kHidden + 4 * (lane + 32 * k)) >> 1;
    const uint32_t b0 = drop_bits(d.key, pair), b1 = drop_bits(d.key, pair + 1);
    r.v[k].x *= drop_mul(b0, 0, d);
    r.v[k].y *= drop_mul(b0, 1, d);
    r.v[k].z *= drop_mul(b1, 0, d);
    r.v[k].w *= drop_mul(b1, 1, d);
  }
}

inline int row_grid(long long rows, int warps_per_block, int blocks_per_sm = 16) {
  long long blocks = (rows + warps_per_block - 1) / warps_per_block;
  const long long cap = 148LL * blocks_per_sm;  // kernels are grid-stride
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace stlt
