// Shared device/host helpers for the STLT sm_100a library.
// Inline-PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor) and tcgen05 (UMMA + TMEM).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace stlt {

constexpr int kHidden = 768;   // src/modelling/configs.py:96
constexpr int kHeads = 12;     // src/modelling/configs.py:99
constexpr int kHeadDim = 64;
constexpr int kFfn = 3072;     // src/modelling/models.py:49 (hidden_size * 4)
constexpr int kQkv = 3 * kHidden;

// ---------------------------------------------------------------------------------------------
// host: opt a kernel into > 48 KB of dynamic shared memory, once per (kernel instantiation, device)
// ---------------------------------------------------------------------------------------------
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, unsigned long long* done_mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (*done_mask & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) *done_mask |= bit;  // the attribute is per device: a second GPU in the process needs its own call
  return e;
}

// ---------------------------------------------------------------------------------------------
// small math helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {
  // F.gelu default ("gelu" activation string, src/modelling/models.py:51,123,163)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// 2^x on both halves: ex2.approx.f16x2 is two MUFU.EX2.F16 and a PRMT; cuda_fp16's h2exp2() widens to fp32 and back
// (two conversions, two MUFU.EX2, one pack).
__device__ __forceinline__ __half2 ex2_h2(__half2 x) {
  uint32_t r;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(*reinterpret_cast<const uint32_t*>(&x)));
  return *reinterpret_cast<__half2*>(&r);
}

// Single-branch erf-GELU for the bf16 epilogue: erf(z) = 1 - 2^(-a*q(a)), a = |x| (degree-4 minimax
// fit of -log2(erfc(z))/z, z = a/sqrt(2); |erf error| < 7e-7 in fp32, far below bf16 rounding).
// ~11 instructions and one MUFU.EX2 per element instead of erff()'s two divergent branches.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float ax = fabsf(x);
  const float a = fminf(ax, 5.9f);        // erf(5.9/sqrt(2)) == 1 in fp32
  float q = 5.204604041e-04f;
  q = fmaf(q, a, -7.397519993e-03f);
  q = fmaf(q, a, 5.256125276e-02f);
  q = fmaf(q, a, 4.592546886e-01f);
  q = fmaf(q, a, 1.151091390e+00f);
  float e;                                // erfc(|x|/sqrt(2)) = 2^(-q*a)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-q * a));
  const float h = fmaf(-ax, e, ax);       // |x| * erf(|x|/sqrt(2))
  return 0.5f * (x + h);                  // 0.5*x*(1 + erf(x/sqrt(2)))
}

// Two elements at a time in packed fp16 (HFMA2 / MUFU.EX2 f16x2): half the instructions of the fp32 version.
// The bf16 epilogue rounds the result to 8 mantissa bits anyway; fp16's 11 bits keep the extra error below a
// quarter of that rounding step (|x| <= 65504, far above any pre-activation of this model).
__device__ __forceinline__ __half2 gelu_erf_fast_h2(__half2 x) {
  const __half2 ax = __habs2(x);
  const __half2 a = __hmin2(ax, __float2half2_rn(5.9f));
  __half2 q = __float2half2_rn(5.204604041e-04f);
  q = __hfma2(q, a, __float2half2_rn(-7.397519993e-03f));
  q = __hfma2(q, a, __float2half2_rn(5.256125276e-02f));
  q = __hfma2(q, a, __float2half2_rn(4.592546886e-01f));
  q = __hfma2(q, a, __float2half2_rn(1.151091390e+00f));
  const __half2 e = ex2_h2(__hneg2(__hmul2(q, a)));       // erfc(|x| / sqrt 2)
  const __half2 h = __hfma2(__hneg2(ax), e, ax);          // |x| * erf(|x| / sqrt 2)
  return __hmul2(__float2half2_rn(0.5f), __hadd2(x, h));  // 0.5 * x * (1 + erf(x / sqrt 2))
}

// GELU and its derivative together (training forward: the FFN1 epilogue stores act(u) and gelu'(u), both already
// multiplied by the FFN-inner dropout mask, so the backward pass never re-evaluates the activation): the erfc term is
// shared, the derivative costs one more ex2 and a few HFMA2. cdf = 0.5 + sign(u) * 0.5 * erf(|u| / sqrt 2).
__device__ __forceinline__ void gelu_erf_and_grad_fast_h2(__half2 x, __half2& act, __half2& grad) {
  const __half2 ax = __habs2(x);
  const __half2 a = __hmin2(ax, __float2half2_rn(5.9f));
  __half2 q = __float2half2_rn(5.204604041e-04f);
  q = __hfma2(q, a, __float2half2_rn(-7.397519993e-03f));
  q = __hfma2(q, a, __float2half2_rn(5.256125276e-02f));
  q = __hfma2(q, a, __float2half2_rn(4.592546886e-01f));
  q = __hfma2(q, a, __float2half2_rn(1.151091390e+00f));
  const __half2 e = ex2_h2(__hneg2(__hmul2(q, a)));       // erfc(|x| / sqrt 2)
  const __half2 h = __hfma2(__hneg2(ax), e, ax);          // |x| * erf(|x| / sqrt 2)
  act = __hmul2(__float2half2_rn(0.5f), __hadd2(x, h));
  const __half2 pdf = ex2_h2(__hmul2(__float2half2_rn(-0.72134752044f), __hmul2(a, a)));  // exp(-x^2 / 2)
  const __half2 half_erf = __hfma2(__float2half2_rn(-0.5f), e, __float2half2_rn(0.5f));   // >= 0
  // copy the sign of x onto half_erf (both halves)
  const uint32_t xb = *reinterpret_cast<const uint32_t*>(&x);
  uint32_t hb = *reinterpret_cast<const uint32_t*>(&half_erf);
  hb ^= xb & 0x80008000u;
  const __half2 signed_half_erf = *reinterpret_cast<const __half2*>(&hb);
  const __half2 cdf = __hadd2(signed_half_erf, __float2half2_rn(0.5f));
  grad = __hfma2(__hmul2(x, __float2half2_rn(0.3989422804014327f)), pdf, cdf);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
  return *reinterpret_cast<uint32_t*>(&v);
}

// Residual of a float after bf16 rounding (the "lo" plane of the 2-term split).
__device__ __forceinline__ float bf16_residual(float x) {
  return x - __bfloat162float(__float2bfloat16_rn(x));
}

// ---------------------------------------------------------------------------------------------
// Dropout (training step): stateless, counter-based masks so the backward pass regenerates the mask
// of every site instead of storing it. Element e of a site is kept iff the 16-bit field (e & 1) of
// hash(site key, e >> 1) is >= thr16 (drop probability thr16 / 65536); kept values are scaled by
// 65536 / (65536 - thr16). thr16 == 0 disables the site. The hash is "lowbias32" (two xorshift-
// multiply rounds), keyed per (seed, site) on the host (stlt_train.cu: site_key()).
// ---------------------------------------------------------------------------------------------
struct DropCfg {
  uint32_t key;
  uint32_t thr16;
  float scale;
};

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

// 32 mask bits for the element pair (2*pair, 2*pair + 1) of a site.
__host__ __device__ __forceinline__ uint32_t drop_bits(uint32_t key, unsigned long long pair) {
  uint32_t x = lowbias32(static_cast<uint32_t>(pair) ^ key);
  const uint32_t hi = static_cast<uint32_t>(pair >> 32);
  if (hi != 0) x = lowbias32(x ^ (hi * 0x9e3779b9u));  // sites with more than 2^33 elements only
  return x;
}

// multiplier (0 or scale) of element `which` (0 / 1) of the pair
__host__ __device__ __forceinline__ float drop_mul(uint32_t bits, int which, const DropCfg& d) {
  return ((bits >> (16 * which)) & 0xffffu) >= d.thr16 ? d.scale : 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// PTX: shared-memory addresses, election
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// PTX: mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// PTX: TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* d) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(d)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* d, uint64_t* bar, void* smem_dst,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(d)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Multicast variant: the box lands at the same CTA-relative smem offset of every CTA in cta_mask and
// completes bytes on the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_multicast(const CUtensorMap* d, uint64_t* bar,
                                                      void* smem_dst, int32_t c0, int32_t c1,
                                                      uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(d)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// CTA-pair variant (cta_group::2): the box lands in THIS CTA's smem, the bytes are signalled on an
// mbarrier that may live in the peer CTA (`bar_cluster_addr` is a shared::cluster address, see mapa).
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* d, uint32_t bar_cluster_addr,
                                                 void* smem_dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(d)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* d, const void* smem_src, int32_t c0,
                                             int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(d)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
      : "memory");
}

// TMA reduce-add store: global[tile] += smem tile (fp32 adds performed by the memory system).
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* d, const void* smem_src,
                                                  int32_t c0, int32_t c1) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(d)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// Wait until the smem source of all committed bulk stores has been read.
__device__ __forceinline__ void tma_store_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// Make generic-proxy smem writes visible to the async proxy (TMA).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// PTX: tcgen05 (TMEM allocation, UMMA, commit, TMEM loads)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (bf16 inputs, fp32 accumulate), issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mbarrier arrive once all previously issued tcgen05 ops of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// Same, arriving on the barrier at this offset in every CTA of cta_mask (cluster-shared stages).
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ---- CTA-pair (cta_group::2) flavours: TMEM is allocated in both CTAs, ONE thread of the leader
// CTA issues MMAs that span both SMs (M = 256: 128 rows per CTA), commits signal both CTAs. ----
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// shared::cluster address of `p` (a smem object of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  // default (.release at CTA scope): the only ordering needed is against this warp's tcgen05.ld,
  // which tcgen05.fence::before_thread_sync provides. A cluster-scope release costs a MEMBAR.ALL.GPU.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}

// 32 lanes x 32 columns of fp32: thread t of the warp receives row (lane base + t), cols c..c+31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Same, but the destination registers of the pending load are routed through the statement so the
// compiler cannot schedule their first use above the wait (needed when loads are issued ahead).
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]),
                 "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]),
                 "+r"(v[31])
               :
               : "memory");
}

// UMMA shared-memory matrix descriptor for a K-major, 128-byte-swizzled tile whose rows are 128 B
// (64 bf16) wide: 8-row groups are 1024 B apart (SBO), LBO unused for swizzled K-major (=1),
// descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // LBO (16 B units) [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO (16 B units) [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // version         [46,48)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B    [61,64)
  return d;
}

// Same for an MN-major operand (the MN index is the contiguous one): the tile is a row of 64-element
// (128 B) wide, 128-byte-swizzled chunks; `chunk_bytes` apart along MN (LBO), 8-row groups of the
// reduction axis 1024 B apart (SBO).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t chunk_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(chunk_bytes >> 4) << 16;  // LBO
  d |= static_cast<uint64_t>(1024 >> 4) << 32;         // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4)                                 // c_format = F32
         | (1u << 7)                               // a_format = BF16
         | (1u << 10)                              // b_format = BF16
         | (static_cast<uint32_t>(n >> 3) << 17)   // n_dim
         | (static_cast<uint32_t>(m >> 4) << 24);  // m_dim
}

}  // namespace stlt
