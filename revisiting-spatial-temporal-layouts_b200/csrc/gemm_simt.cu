// fp32 CUDA-core GEMM: out[M, N] = act(A[M, K] * W[N, K]^T + bias[N]).
// Used for the classifier head fc1 / fc2 (src/modelling/models.py:155-163; 0.02 % of the FLOPs,
// kept off the tensor cores per the north star) and as an independent numerical cross-check of the
// tcgen05 GEMM in the parity tests. 64x64 tile, 16-wide K steps, 4x4 outputs per thread.
#include "common.cuh"
#include "kernels.h"

namespace stlt {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <bool kGelu>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ a, const float* __restrict__ w,
                 const float* __restrict__ bias, float* __restrict__ out, int m, int n, int k) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15;  // column group
  const int ty = tid >> 4;  // row group
  const int m0 = blockIdx.y * TM;
  const int n0 = blockIdx.x * TN;
  const int lr = tid >> 2;        // 0..63: tile row loaded by this thread
  const int lk = (tid & 3) * 4;   // 0,4,8,12: k offset loaded by this thread

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < k; k0 += TK) {
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vw = va;
    if (m0 + lr < m) va = __ldg(reinterpret_cast<const float4*>(a + (long long)(m0 + lr) * k + k0 + lk));
    if (n0 + lr < n) vw = __ldg(reinterpret_cast<const float4*>(w + (long long)(n0 + lr) * k + k0 + lk));
    As[lk + 0][lr] = va.x; As[lk + 1][lr] = va.y; As[lk + 2][lr] = va.z; As[lk + 3][lr] = va.w;
    Ws[lk + 0][lr] = vw.x; Ws[lk + 1][lr] = vw.y; Ws[lk + 2][lr] = vw.z; Ws[lk + 3][lr] = vw.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= n) continue;
      float v = acc[i][j] + (bias != nullptr ? __ldg(bias + col) : 0.f);
      if (kGelu) v = gelu_erf(v);
      out[(long long)row * n + col] = v;
    }
  }
}

}  // namespace

cudaError_t launch_gemm_simt(const float* a, const float* w, const float* bias, float* out, int m,
                             int n, int k, bool gelu, cudaStream_t stream) {
  if (m == 0 || n == 0) return cudaSuccess;
  if (k % TK != 0) return cudaErrorInvalidValue;
  dim3 grid((n + TN - 1) / TN, (m + TM - 1) / TM);
  if (gelu)
    gemm_simt_kernel<true><<<grid, 256, 0, stream>>>(a, w, bias, out, m, n, k);
  else
    gemm_simt_kernel<false><<<grid, 256, 0, stream>>>(a, w, bias, out, m, n, k);
  return cudaGetLastError();
}

}  // namespace stlt
