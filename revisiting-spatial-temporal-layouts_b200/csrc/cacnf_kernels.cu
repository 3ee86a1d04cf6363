// Small HBM-bound kernels of the CACNF fusion path on precomputed ResNet3D features (SURVEY.md 8(f)
// rank 2; reference src/modelling/models.py:232-283 TransformerResnet, :434-483
// CrossAttentionFusionBackbone, :504-549 CrossAttentionCentralNetFusion).
#include "rowops.cuh"

namespace stlt {

namespace {

// [B][C][P] f32 -> [B * P][C] bf16 through a padded shared-memory tile (64 channels x P positions).
// Reads are 128-byte rows of P = 32 floats, writes are 128-byte runs of 64 bf16.
__global__ void __launch_bounds__(256)
features_to_tokens_kernel(const float* __restrict__ feat, __nv_bfloat16* __restrict__ out, int C, int P,
                          long long lo_plane_elems) {
  __shared__ float tile[64][33];
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * 64;
  const float* src = feat + (static_cast<long long>(b) * C + c0) * P;
  for (int e = threadIdx.x; e < 64 * P; e += 256) {
    const int c = e / P, s = e - c * P;
    tile[c][s] = (c0 + c < C) ? __ldg(src + static_cast<long long>(c) * P + s) : 0.f;
  }
  __syncthreads();
  // thread -> (position s, channel pair)
  for (int e = threadIdx.x; e < P * 32; e += 256) {
    const int s = e >> 5, cp = e & 31;
    if (c0 + 2 * cp + 1 < C + 1) {
      const float a = tile[2 * cp][s], d = tile[2 * cp + 1][s];
      __nv_bfloat16* dst = out + (static_cast<long long>(b) * P + s) * C + c0 + 2 * cp;
      *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(a, d);
      if (lo_plane_elems != 0)  // fp32-parity mode: residual plane
        *reinterpret_cast<uint32_t*>(dst + lo_plane_elems) = pack_bf16x2(bf16_residual(a), bf16_residual(d));
    }
  }
}

__global__ void __launch_bounds__(256)
app_embed_kernel(const float* __restrict__ proj, const float* __restrict__ cls,
                 const float* __restrict__ pos, int P, long long rows, ActOut out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  const int T = P + 1;
  for (long long row = warp0; row < rows; row += nwarps) {
    const long long b = row / T;
    const int tpos = static_cast<int>(row - b * T);
    RowRegs r = tpos == 0 ? load_row(cls, 0, lane) : load_row(proj, b * P + (tpos - 1), lane);
    const RowRegs pe = load_row(pos, tpos, lane);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      r.v[k].x += pe.v[k].x;
      r.v[k].y += pe.v[k].y;
      r.v[k].z += pe.v[k].z;
      r.v[k].w += pe.v[k].w;
    }
    store_act(out, row, r, lane);
  }
}

__global__ void __launch_bounds__(256)
gather_concat_kernel(const float* __restrict__ a, int a_stride, const long long* __restrict__ lengths,
                     const float* __restrict__ c, int c_stride, int B, float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long b = warp0; b < B; b += nwarps) {
    long long ra = b * a_stride;
    if (lengths != nullptr) {
      long long len = lengths[b];
      if (len < 1 || len > a_stride) len = 1;
      ra += len - 1;
    }
    const RowRegs x = load_row(a, ra, lane);
    const RowRegs y = load_row(c, b * c_stride, lane);
    float4* d = reinterpret_cast<float4*>(dst + b * 2 * kHidden);
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      d[lane + 32 * k] = x.v[k];
      d[kHidden / 4 + lane + 32 * k] = y.v[k];
    }
  }
}

__global__ void mean3_kernel(const float* __restrict__ a, const float* __restrict__ b,
                             const float* __restrict__ c, long long n, float* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  // sum(logits) / 3 with Python's left-to-right sum starting from 0 (models.py:546)
  if (i < n) out[i] = (((0.f + a[i]) + b[i]) + c[i]) / 3.0f;
}

}  // namespace

cudaError_t launch_features_to_tokens(const float* feat, __nv_bfloat16* out, int B, int C, int P,
                                      cudaStream_t stream, long long lo_plane_rows) {
  if (B == 0) return cudaSuccess;
  if (P < 1 || P > 32 || C % 2 != 0) return cudaErrorInvalidValue;
  dim3 grid((C + 63) / 64, B);
  features_to_tokens_kernel<<<grid, 256, 0, stream>>>(feat, out, C, P, lo_plane_rows * C);
  return cudaGetLastError();
}

cudaError_t launch_app_embed(const float* proj, const float* cls, const float* pos, int B, int P, ActOut out,
                             cudaStream_t stream) {
  const long long rows = static_cast<long long>(B) * (P + 1);
  if (rows == 0) return cudaSuccess;
  app_embed_kernel<<<row_grid(rows, 8), 256, 0, stream>>>(proj, cls, pos, P, rows, out);
  return cudaGetLastError();
}

cudaError_t launch_gather_concat(const float* a, int a_stride, const long long* lengths, const float* c,
                                 int c_stride, int B, float* dst, cudaStream_t stream) {
  if (B == 0) return cudaSuccess;
  gather_concat_kernel<<<row_grid(B, 8), 256, 0, stream>>>(a, a_stride, lengths, c, c_stride, B, dst);
  return cudaGetLastError();
}

cudaError_t launch_mean3(const float* a, const float* b, const float* c, long long n, float* out,
                         cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  mean3_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(a, b, c, n, out);
  return cudaGetLastError();
}

}  // namespace stlt
