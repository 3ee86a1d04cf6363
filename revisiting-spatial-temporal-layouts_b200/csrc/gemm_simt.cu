// fp32 CUDA-core GEMM: out[M, N] = act(A[M, K] * W[N, K]^T + bias[N]).
// Used for the classifier head fc1 / fc2 (src/modelling/models.py:155-163; 0.02 % of the FLOPs,
// kept off the tensor cores per the north star) and as an independent numerical cross-check of the
// tcgen05 GEMM in the parity tests. 128x64 tile, 32-wide K steps prefetched into registers, 8x8 outputs per thread.
#include "common.cuh"
#include "kernels.h"

namespace stlt {

namespace {

constexpr int TM = 128, TN = 64, TK = 32, kThreads = 128;

// Thread (ty, tx) of a 16 x 8 grid owns rows {4ty..4ty+3, 64+4ty..} x columns {4tx..4tx+3, 32+4tx..}: an 8x8
// register tile (64 FFMA per four 16-byte shared-memory loads) whose loads are bank-conflict free.
template <bool kGelu>
__global__ void __launch_bounds__(kThreads, 3)
gemm_simt_kernel(const float* __restrict__ a, const float* __restrict__ w,
                 const float* __restrict__ bias, float* __restrict__ out, int m, int n, int k) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 7;
  const int ty = tid >> 3;
  const int m0 = blockIdx.y * TM;
  const int n0 = blockIdx.x * TN;
  const int lr = tid >> 3;        // 0..15 (+16 per vector): tile row loaded by this thread
  const int lk = (tid & 7) * 4;   // k offset loaded by this thread: a row's 32 k-values are one 128-byte line
  const int sw_store = 4 * (((tid & 7) >> 1) & 3);  // = 4 * ((k >> 3) & 3) for k = lk .. lk + 3
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // the next k-slab travels from global memory to registers while the current one is multiplied
  float4 va[TM / 16], vw[TN / 16];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int h = 0; h < TM / 16; ++h) {
      const int row = m0 + lr + 16 * h;
      va[h] = row < m ? __ldg(reinterpret_cast<const float4*>(a + (long long)row * k + k0 + lk)) : zero;
    }
#pragma unroll
    for (int h = 0; h < TN / 16; ++h) {
      const int row = n0 + lr + 16 * h;
      vw[h] = row < n ? __ldg(reinterpret_cast<const float4*>(w + (long long)row * k + k0 + lk)) : zero;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < k; k0 += TK) {
    // transposed stores: the 4-row groups are XOR-permuted per 8 k-values so that the eight k-offsets of a warp
    // hit different banks (the row pitch of 132 / 68 floats alone separates only two of them)
#pragma unroll
    for (int h = 0; h < TM / 16; ++h) {
      const int r = (lr + 16 * h) ^ sw_store;
      As[lk + 0][r] = va[h].x; As[lk + 1][r] = va[h].y; As[lk + 2][r] = va[h].z; As[lk + 3][r] = va[h].w;
    }
#pragma unroll
    for (int h = 0; h < TN / 16; ++h) {
      const int r = (lr + 16 * h) ^ sw_store;
      Ws[lk + 0][r] = vw[h].x; Ws[lk + 1][r] = vw[h].y; Ws[lk + 2][r] = vw[h].z; Ws[lk + 3][r] = vw[h].w;
    }
    __syncthreads();
    if (k0 + TK < k) fetch(k0 + TK);
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const int sw = 4 * ((kk >> 3) & 3);
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][(ty * 4) ^ sw]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ((ty * 4) ^ sw)]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[kk][(tx * 4) ^ sw]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[kk][32 + ((tx * 4) ^ sw)]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wr[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (row >= m) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = n0 + jh * 32 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = acc[i][jh * 4 + j] + ((bias != nullptr && col + j < n) ? __ldg(bias + col + j) : 0.f);
        if (kGelu) v[j] = gelu_erf(v[j]);
      }
      float* dst = out + (long long)row * n + col;
      if (col + 3 < n && (n & 3) == 0) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < n) dst[j] = v[j];
      }
    }
  }
}

// Head gradients of the training step (fc1 / fc2 of ClassificationHead, models.py:155-163):
//   out[i, j] (+)= sum_k A(i, k) * B[k * ldb + j],  A(i, k) = a[i * lda + k] (kAKContig) or a[k * lda + i].
// Same 128x64 tile, 8x8 register tile and swizzled [k][m] shared-memory layout as gemm_simt_kernel; operands are
// fetched element-wise (coalesced along their contiguous axis; the 174-column logits gradient is not 16-byte aligned)
// into registers one k-slab ahead. Weight gradients (few output tiles, reduction over the batch) are split over
// blockIdx.z and accumulated with atomics.
template <bool kAKContig>
__global__ void __launch_bounds__(kThreads, 2)
gemm_grad_simt_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b, long long ldb,
                      float* __restrict__ out, int m, int n, int k, int k_per_split, int atomic) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 7;
  const int ty = tid >> 3;
  const int m0 = blockIdx.y * TM;
  const int n0 = blockIdx.x * TN;
  const int kz0 = blockIdx.z * k_per_split;
  const int kz1 = kz0 + k_per_split < k ? kz0 + k_per_split : k;
  constexpr int kAPer = TM * TK / kThreads;  // 32
  constexpr int kBPer = TN * TK / kThreads;  // 16

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float va[kAPer], vb[kBPer];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < kAPer; ++r) {
      const int e = tid + kThreads * r;
      const int kk = kAKContig ? (e & (TK - 1)) : (e / TM);
      const int ii = kAKContig ? (e / TK) : (e & (TM - 1));
      const int gi = m0 + ii, gk = k0 + kk;
      const bool ok = gi < m && gk < kz1;
      va[r] = ok ? __ldg(kAKContig ? a + gi * lda + gk : a + gk * lda + gi) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < kBPer; ++r) {
      const int e = tid + kThreads * r;
      const int kk = e / TN, jj = e & (TN - 1);
      const int gj = n0 + jj, gk = k0 + kk;
      vb[r] = (gj < n && gk < kz1) ? __ldg(b + gk * ldb + gj) : 0.f;
    }
  };
  fetch(kz0);
  for (int k0 = kz0; k0 < kz1; k0 += TK) {
#pragma unroll
    for (int r = 0; r < kAPer; ++r) {
      const int e = tid + kThreads * r;
      const int kk = kAKContig ? (e & (TK - 1)) : (e / TM);
      const int ii = kAKContig ? (e / TK) : (e & (TM - 1));
      As[kk][ii ^ (4 * ((kk >> 3) & 3))] = va[r];
    }
#pragma unroll
    for (int r = 0; r < kBPer; ++r) {
      const int e = tid + kThreads * r;
      const int kk = e / TN, jj = e & (TN - 1);
      Ws[kk][jj ^ (4 * ((kk >> 3) & 3))] = vb[r];
    }
    __syncthreads();
    if (k0 + TK < kz1) fetch(k0 + TK);
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const int sw = 4 * ((kk >> 3) & 3);
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][(ty * 4) ^ sw]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ((ty * 4) ^ sw)]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[kk][(tx * 4) ^ sw]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[kk][32 + ((tx * 4) ^ sw)]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wr[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = n0 + (j >> 2) * 32 + tx * 4 + (j & 3);
      if (col >= n) continue;
      float* o = out + static_cast<long long>(row) * n + col;
      if (atomic) atomicAdd(o, acc[i][j]);
      else *o = acc[i][j];
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
    gemm_simt_kernel<true><<<grid, kThreads, 0, stream>>>(a, w, bias, out, m, n, k);
  else
    gemm_simt_kernel<false><<<grid, kThreads, 0, stream>>>(a, w, bias, out, m, n, k);
  return cudaGetLastError();
}

cudaError_t launch_gemm_strided(const float* a, long long sai, long long sak, const float* b,
                                long long sbk, long long sbj, float* out, int m, int n, int k,
                                bool accumulate, cudaStream_t stream) {
  if (m == 0 || n == 0) return cudaSuccess;
  if (sbj != 1 || (sai != 1 && sak != 1)) return cudaErrorInvalidValue;  // B row-major [K, N]; A contiguous along M or K
  const int tiles = ((n + TN - 1) / TN) * ((m + TM - 1) / TM);
  // accumulating calls (weight gradients: few tiles, long reduction) are split along k until the grid is ~one wave
  int splits = 1;
  if (accumulate) {
    splits = (148 * 2 + tiles - 1) / tiles;
    const int max_splits = (k + 4 * TK - 1) / (4 * TK);  // at least four k-slabs per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int k_per_split = (k + splits - 1) / splits;
  k_per_split = (k_per_split + TK - 1) / TK * TK;
  splits = (k + k_per_split - 1) / k_per_split;
  dim3 grid((n + TN - 1) / TN, (m + TM - 1) / TM, splits);
  if (sak == 1)
    gemm_grad_simt_kernel<true><<<grid, kThreads, 0, stream>>>(a, sai, b, sbk, out, m, n, k, k_per_split, accumulate ? 1 : 0);
  else
    gemm_grad_simt_kernel<false><<<grid, kThreads, 0, stream>>>(a, sak, b, sbk, out, m, n, k, k_per_split, accumulate ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace stlt
