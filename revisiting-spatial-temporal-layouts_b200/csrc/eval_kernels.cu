// On-device evaluator for the multi-label (Action Genome / Charades) head — SURVEY.md 8(f) rank 4:
// EvaluatorActionGenome.process (src/utils/evaluation.py:76-83: sigmoid of the logits and the labels are
// appended to two [total_instances, classes] arrays — here device buffers, no per-batch .cpu()) and
// charades_map / map (:100-132: videos without any positive label get -inf scores, then per class the
// average precision over the score-sorted list).
#include "common.cuh"
#include "kernels.h"

namespace stlt {

namespace {

__global__ void map_accumulate_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                      long long n, float* __restrict__ pred, float* __restrict__ gt) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  pred[i] = 1.0f / (1.0f + expf(-logits[i]));  // torch.sigmoid in fp32 (evaluation.py:79-80)
  gt[i] = labels[i];
}

// One block per class. Scores (with the -inf fix of charades_map) and the true-positive flags are sorted
// together by descending score with a bitonic network in shared memory; ties keep no particular order, as
// with np.argsort(-x). AP = sum over positives of (true positives so far / rank) / number of positives.
__global__ void __launch_bounds__(1024)
charades_ap_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int n, int classes, int n_pow2,
                   double* __restrict__ ap_out) {
  extern __shared__ __align__(8) unsigned char smem[];
  float* key = reinterpret_cast<float*>(smem);                          // [n_pow2] negated scores (ascending sort)
  unsigned char* tp = reinterpret_cast<unsigned char*>(key + n_pow2);   // [n_pow2]
  int* scan = reinterpret_cast<int*>(tp + ((n_pow2 + 3) & ~3));         // [blockDim.x] running positives
  const int c = blockIdx.x;
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
    float k = INFINITY;  // padding sorts last
    unsigned char t = 0;
    if (i < n) {
      const float* grow = gt + static_cast<long long>(i) * classes;
      float any = 0.f;
      for (int j = 0; j < classes; ++j) any += grow[j];                // charades_map: empty = sum(gt, axis=1) == 0
      const float s = any == 0.f ? -INFINITY : pred[static_cast<long long>(i) * classes + c];
      k = -s;
      t = grow[c] == 1.0f ? 1 : 0;
    }
    key[i] = k;
    tp[i] = t;
  }
  __syncthreads();
  for (int size = 2; size <= n_pow2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < n_pow2 / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const float a = key[lo], b = key[hi];
        if ((a > b) == up) {
          key[lo] = b;
          key[hi] = a;
          const unsigned char ta = tp[lo];
          tp[lo] = tp[hi];
          tp[hi] = ta;
        }
      }
      __syncthreads();
    }
  // chunked scan: thread t owns ranks [t * per, (t + 1) * per)
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int r0 = threadIdx.x * per, r1 = min(n, r0 + per);
  int local = 0;
  for (int r = r0; r < r1; ++r) local += tp[r];
  scan[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < blockDim.x; ++i) {
      const int v = scan[i];
      scan[i] = run;
      run += v;
    }
  }
  __syncthreads();
  double acc = 0.0;
  int seen = scan[threadIdx.x];
  for (int r = r0; r < r1; ++r)
    if (tp[r]) {
      ++seen;
      acc += static_cast<double>(seen) / static_cast<double>(r + 1);  // prec[i] = t_pcs / (f_pcs + t_pcs)
    }
  __shared__ double partial[1024];
  __shared__ int total_pos;
  partial[threadIdx.x] = acc;
  if (threadIdx.x == blockDim.x - 1) total_pos = seen;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < blockDim.x; ++i) s += partial[i];
    ap_out[c] = total_pos < 1 ? NAN : s / static_cast<double>(total_pos);  // n_pos < 0.1 -> nan (evaluation.py:108-110)
  }
}

__global__ void mean_ap_kernel(const double* __restrict__ ap, int classes, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < classes; ++i) s += ap[i];  // np.mean: a class without positives makes the mean nan
    out[0] = s / classes;
  }
}

}  // namespace

cudaError_t launch_map_accumulate(const float* logits, const float* labels, long long n, float* pred, float* gt,
                                  cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  map_accumulate_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(logits, labels, n, pred, gt);
  return cudaGetLastError();
}

cudaError_t launch_charades_map(const float* pred, const float* gt, int n, int classes, double* ap_out,
                                double* map_out, cudaStream_t stream) {
  if (n < 1 || classes < 1) return cudaErrorInvalidValue;
  int n_pow2 = 1;
  while (n_pow2 < n) n_pow2 <<= 1;
  const int threads = 1024;
  const size_t smem = static_cast<size_t>(n_pow2) * 4 + ((n_pow2 + 3) & ~3) + threads * sizeof(int);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;  // > 32768 instances: does not fit one block's shared memory
  cudaError_t e = cudaFuncSetAttribute(charades_ap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  charades_ap_kernel<<<classes, threads, smem, stream>>>(pred, gt, n, classes, n_pow2, ap_out);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  mean_ap_kernel<<<1, 32, 0, stream>>>(ap_out, classes, map_out);
  return cudaGetLastError();
}

}  // namespace stlt
