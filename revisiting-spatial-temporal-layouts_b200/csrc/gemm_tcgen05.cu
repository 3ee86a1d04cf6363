// tcgen05 projection GEMM for the STLT encoder layers (QKV / out-proj / FFN1 / FFN2).
//
//   out[M, N] = epilogue( sum_terms A_t[M, K] * W_t[N, K]^T + bias[N] )
//
// Replaces the F.linear calls inside nn.TransformerEncoderLayer / nn.MultiheadAttention that the
// reference instantiates at src/modelling/models.py:46-55 and :118-128.
//
// Operands are bf16, K-major, staged by TMA into 128B-swizzled shared memory; accumulators are fp32
// in TMEM (2 x 256 columns, double buffered); one thread issues tcgen05.mma (M=128, N=256, K=16).
//
// Precision modes:
//   kTerms == 1 : plain bf16 operands.
//   kTerms == 3 : fp32-parity mode. Activations and weights are each stored as two bf16 planes
//                 (hi = bf16(x), lo = bf16(x - hi)); the kernel accumulates hi*hi + lo*hi + hi*lo
//                 into the same fp32 TMEM accumulator (the dropped lo*lo term is ~2^-16 relative).
//                 Planes are stacked along the row axis of the same tensor map: plane p of A starts
//                 at row p * a_plane_rows, plane p of W at row p * b_plane_rows.
//
// CTAs run as pairs (cluster of 2, tcgen05 cta_group::2): the pair computes a 256 x 256 output tile
// with ONE M=256 MMA stream issued by the leader CTA. Each CTA stages its own 128 A rows and HALF
// (128 rows) of the W tile, so a k-block costs 32 KB of L2->SM traffic and shared-memory fill per CTA
// instead of 48 KB, the tensor core reads each W half from shared memory once for both SMs, and the
// smaller stage buys a 6-deep pipeline. (The 1-CTA version measured 71 % tensor-pipe utilisation,
// paced by shared-memory bandwidth: 96 B/clk of operand reads + 96 B/clk of TMA fill per SM.)
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only), warp 2 =
// TMEM allocator, warp 3 = idle, warps 4..11 = epilogue (warp w reads TMEM lane quarter w % 4, column
// half (w-4)/4 of this CTA's 128 x 256 accumulator).
//
// Operand layouts (kLayout), used by the training step:
//   GEMM_NT    (forward)        out[M,N]  = A[M,K] * B[N,K]^T + bias      A, B K-major
//   GEMM_NN    (data gradient)  out[M,N]  = A[M,K] * B[K,N]                B MN-major (no transposed copy of W)
//   GEMM_TN_RED(weight gradient) out[M,N] += A[K,M]^T * B[K,N]             A, B MN-major; the reduction runs
//              over the token axis and is split into aligned k-slices, one (tile, slice) unit per CTA pair;
//              every partial tile is added to the fp32 output with a TMA reduce-add store (no fix-up pass).
// MN-major tiles are staged as 64-element (128 B) wide TMA boxes of 64 reduction rows: a 128-wide
// operand is two boxes 8 KiB apart (descriptor LBO), 8-row groups are 1 KiB apart (SBO).
//
// Barriers: full[s] lives in the leader (both CTAs' TMA loads signal it); empty[s] and tmem_full[a]
// exist in both CTAs and are signalled by the leader's multicast tcgen05.commit; tmem_empty[a] lives
// in the leader and collects the epilogue warps of both CTAs.
#include "common.cuh"
#include "kernels.h"

namespace stlt {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int kABytes = BM * BK * 2;  // 16 KiB
constexpr int kBHalfBytes = (BN / 2) * BK * 2;  // 16 KiB: this CTA's half of the W tile
constexpr int kStageBytes = kABytes + kBHalfBytes;  // per CTA
constexpr int kEpiWarps = 8;
constexpr int kEpiBufBytes = 32 * 128;  // 32 rows x 128 B staging tile for one TMA store
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 6;
constexpr int kCluster = 2;
constexpr uint16_t kClusterMask = (1u << kCluster) - 1;

// The split epilogue needs two staging tiles per warp (hi and lo plane), paid for with one stage.
template <int kOut>
struct Cfg {
  static constexpr bool kTwoPlanes =
      kOut == GEMM_OUT_BF16_SPLIT || kOut == GEMM_OUT_BF16_DUAL || kOut == GEMM_OUT_F32_BF16_DIRECT;
  // fused-LN residual epilogue: outgoing fp32 tile, incoming z tile (+ outgoing bf16 tile unless DIRECT)
  static constexpr int kStages = kOut == GEMM_OUT_F32_BF16 ? 4 : (kTwoPlanes ? 5 : 6);
  static constexpr int kBufsPerWarp = kOut == GEMM_OUT_F32_BF16 ? 3 : (kTwoPlanes ? 2 : (kOut == GEMM_OUT_HILO ? 0 : 1));
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kEpiWarps * kBufsPerWarp * kEpiBufBytes + 320 /*barriers*/;
};
static_assert(Cfg<GEMM_OUT_F32>::kSmemBytes <= 232448, "smem budget");
static_assert(Cfg<GEMM_OUT_BF16_SPLIT>::kSmemBytes <= 232448, "smem budget");
static_assert(Cfg<GEMM_OUT_BF16_DUAL>::kSmemBytes <= 232448, "smem budget");
static_assert(Cfg<GEMM_OUT_F32_BF16>::kSmemBytes <= 232448, "smem budget");
static_assert(Cfg<GEMM_OUT_F32_BF16_DIRECT>::kSmemBytes <= 232448, "smem budget");
static_assert(Cfg<GEMM_OUT_HILO>::kSmemBytes <= 232448, "smem budget");

constexpr int kMnChunkBytes = 64 * BK * 2;  // one MN-major TMA box: 64 reduction rows x 128 B

struct Work {
  int tile, kb0, kb1;
};

// Work distribution of one CTA pair. Tile mode: whole tiles, round-robin. Split-K mode (weight
// gradients, few output tiles and a very long reduction over tokens): the reduction is cut into
// S = floor(pairs / tiles) ALIGNED slices and every pair owns one (tile, slice) unit, accumulated in
// TMEM over the whole slice and reduce-added once. All pairs of a slice sweep the same token rows at
// the same time, so the 256-column operand slabs are read from HBM once and shared by the output tiles
// through L2. (An equal-share stream-K split gave every pair a different k offset: ncu showed 4.8-6.3
// GB of DRAM reads per launch against 1.3 GB of operands; chunking the reduction instead multiplied
// the reduce-add traffic.)
template <bool kSplitK>
struct Sched {
  int cur, end, total_kb, stride, num_tiles, slices;
  __device__ Sched(int cluster_id, int num_clusters, int num_tiles_, int total_kb_)
      : total_kb(total_kb_), num_tiles(num_tiles_) {
    slices = 1;
    if (kSplitK) {
      slices = num_clusters / num_tiles_;
      if (slices < 1) slices = 1;
      if (slices > total_kb_) slices = total_kb_;
    }
    cur = cluster_id;
    end = num_tiles_ * slices;
    stride = num_clusters;
  }
  __device__ bool next(Work& w) {
    if (cur >= end) return false;
    if (kSplitK) {
      w.tile = cur % num_tiles;
      const int slice = cur / num_tiles;
      // even split; never empty because slices <= total_kb
      w.kb0 = static_cast<int>(static_cast<long long>(total_kb) * slice / slices);
      w.kb1 = static_cast<int>(static_cast<long long>(total_kb) * (slice + 1) / slices);
    } else {
      w.tile = cur;
      w.kb0 = 0;
      w.kb1 = total_kb;
    }
    cur += stride;
    return true;
  }
};

struct __align__(8) Barriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t zin[kEpiWarps];  // RESID epilogue: per-warp "incoming z tile has landed"
  uint32_t tmem_base;
  uint32_t pad;
};

}  // namespace

template <int kTerms, int kOut, int kGelu, int kLayout, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_out2,
                    const float* __restrict__ bias, int m_tiles_arg, int n_tiles, int k_blocks, int a_plane_rows,
                    int b_plane_rows, int out_plane_rows, DropCfg drop, EpiArgs ep) {
  // pad-skipping row layout (compact.cu): the number of live 128-row tiles is decided on the device
  const int m_tiles = ep.m_tiles_dyn != nullptr ? min(m_tiles_arg, __ldg(ep.m_tiles_dyn)) : m_tiles_arg;
  static_assert(kEpi == GEMM_EPI_PLAIN || kEpi == GEMM_EPI_ACT_BWD || kLayout == GEMM_NT, "fused-LN epilogues: forward only");
  static_assert(kEpi != GEMM_EPI_ACT_BWD || (kLayout == GEMM_NN && kOut == GEMM_OUT_BF16 && kGelu == 0),
                "ACT_BWD: bf16 data gradient");
  constexpr bool kResidOut = kOut == GEMM_OUT_F32_BF16 || kOut == GEMM_OUT_F32_BF16_DIRECT || kOut == GEMM_OUT_HILO;
  constexpr bool kHiLo = kOut == GEMM_OUT_HILO;
  static_assert((kEpi == GEMM_EPI_RESID) == kResidOut, "RESID epilogue <-> fp32 + bf16 output");
  constexpr int kStages = Cfg<kOut>::kStages;
  constexpr int kBufsPerWarp = Cfg<kOut>::kBufsPerWarp;
  constexpr bool kAMn = kLayout == GEMM_TN_RED;
  constexpr bool kBMn = kLayout != GEMM_NT;
  constexpr bool kStreamK = kLayout == GEMM_TN_RED;  // aligned split-K + reduce-add epilogue
  constexpr bool kBias = kLayout == GEMM_NT;
  static_assert(kLayout == GEMM_NT || kTerms == 1, "gradient GEMMs are single-term");
  static_assert(!kStreamK || kOut == GEMM_OUT_F32, "stream-K partials are reduce-added in fp32");

  // SWIZZLE_128B tiles need 1024 B alignment; the dynamic smem window starts 1024-aligned when
  // the kernel has no static shared memory (checked below).
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_ab = smem;
  uint8_t* smem_epi = smem + kStages * kStageBytes;
  Barriers* bars =
      reinterpret_cast<Barriers*>(smem_epi + kEpiWarps * kBufsPerWarp * kEpiBufBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x / kCluster;
  const int num_clusters = gridDim.x / kCluster;
  // A "pair tile" is two vertically adjacent 128-row tiles (one per CTA of the cluster).
  const int num_tiles = ((m_tiles + kCluster - 1) / kCluster) * n_tiles;
  const int total_kb = kTerms * k_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars->full[i], 1);   // used in the leader: its producer arms the bytes of both CTAs
      mbar_init(&bars->empty[i], 1);  // leader's tcgen05.commit, multicast to both CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);                     // leader's commit, multicast
      mbar_init(&bars->tmem_empty[i], kCluster * kEpiWarps);  // used in the leader: both epilogues
    }
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(&bars->zin[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(&bars->tmem_base, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      Sched<kStreamK> sched(cluster_id, num_clusters, num_tiles, total_kb);
      Work w;
      while (sched.next(w)) {
        const int m_blk = (w.tile / n_tiles) * kCluster + cta_rank;
        const int n_blk = w.tile % n_tiles;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          const int term = kb / k_blocks;
          const int kk = kb - term * k_blocks;
          // term 0: A_hi * W_hi, term 1: A_lo * W_hi, term 2: A_hi * W_lo
          const int a_row = (term == 1 ? a_plane_rows : 0) + m_blk * BM;
          const int b_row = (term == 2 ? b_plane_rows : 0) + n_blk * BN;
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          if (cta_rank == 0) mbar_expect_tx(&bars->full[stage], kCluster * kStageBytes);
          uint8_t* sa = smem_ab + stage * kStageBytes;
          const uint32_t full_leader = mapa_u32(&bars->full[stage], 0);
          if (!kAMn) {
            tma_load_2d_pair(&tm_a, full_leader, sa, kk * BK, a_row);
          } else {  // A^T operand: tensor [reduction, M], two 64-wide boxes
            tma_load_2d_pair(&tm_a, full_leader, sa, m_blk * BM, kk * BK);
            tma_load_2d_pair(&tm_a, full_leader, sa + kMnChunkBytes, m_blk * BM + 64, kk * BK);
          }
          // this CTA's half (128 rows) of the 256-row W tile
          if (!kBMn) {
            tma_load_2d_pair(&tm_b, full_leader, sa + kABytes, kk * BK, b_row + cta_rank * (BN / kCluster));
          } else {  // B operand: tensor [reduction, N]
            const int col = n_blk * BN + cta_rank * (BN / kCluster);
            tma_load_2d_pair(&tm_b, full_leader, sa + kABytes, col, kk * BK);
            tma_load_2d_pair(&tm_b, full_leader, sa + kABytes + kMnChunkBytes, col + 64, kk * BK);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0 && cta_rank == 0) {
      // M = 256 across the pair
      constexpr uint32_t idesc = umma_idesc_bf16(kCluster * BM, BN) | (kAMn ? (1u << 15) : 0u) |
                                 (kBMn ? (1u << 16) : 0u);
      // K advance of 16 elements inside a stage, in 16-byte descriptor units: 32 B along a K-major
      // row, 16 rows x 128 B for MN-major
      constexpr uint32_t kStepA = kAMn ? 128 : 2;
      constexpr uint32_t kStepB = kBMn ? 128 : 2;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      Sched<kStreamK> sched(cluster_id, num_clusters, num_tiles, total_kb);
      Work w;
      for (; sched.next(w); ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1u);  // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&bars->full[stage], phase);  // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_ab + stage * kStageBytes);
          const uint64_t da = kAMn ? umma_desc_mn_sw128(sa, kMnChunkBytes) : umma_desc_k_sw128(sa);
          const uint64_t db = kBMn ? umma_desc_mn_sw128(sa + kABytes, kMnChunkBytes)
                                   : umma_desc_k_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16_ss_pair(tmem_d, da + kStepA * k, db + kStepB * k, idesc,
                              (kb != w.kb0 || k != 0) ? 1u : 0u);
          }
          // frees this smem stage in BOTH CTAs once these MMAs have read it
          umma_commit_pair(&bars->empty[stage], kClusterMask);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_pair(&bars->tmem_full[acc], kClusterMask);  // accumulators complete -> epilogues
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int ew = warp - 4;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = ew >> 2;      // which 128 columns of the 256-wide tile
    uint8_t* ebuf = smem_epi + ew * kBufsPerWarp * kEpiBufBytes;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const uint32_t row_smem = smem_u32(ebuf) + lane * 128;  // this thread's 128 B staging row
    uint8_t* zbuf = ebuf + (kOut == GEMM_OUT_F32_BF16 ? 2 : 1) * kEpiBufBytes;  // RESID: incoming z tile (TMA load)
    uint32_t zphase = 0;
    int it = 0;
    Sched<kStreamK> sched(cluster_id, num_clusters, num_tiles, total_kb);
    Work w;
    for (; sched.next(w); ++it) {
      const int m_blk = (w.tile / n_tiles) * kCluster + cta_rank;
      const int n_blk = w.tile % n_tiles;
      const bool store_ok = m_blk < m_tiles;  // odd tile counts: the last pair has a dummy half
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int col0 = n_blk * BN + half * 128;
      const int row0 = m_blk * BM + quarter * 32;
      const float4* bias4 = reinterpret_cast<const float4*>(bias + col0);  // warp-uniform reads
      // fused LayerNorm: statistics of this thread's row (NORM_A: of the A row; RESID: of the z_prev row)
      float mu = 0.f, rstd = 1.f;
      float psum = 0.f, psq = 0.f;
      if constexpr (kEpi != GEMM_EPI_PLAIN) {
        if (store_ok && (kEpi == GEMM_EPI_NORM_A || ep.prev_norm)) {
          // kStatSlots partial (sum, sum of squares) per row, one per 128-column slab of the producing GEMM,
          // added in a fixed order: no atomics, bit-reproducible
          const float4* st4 = reinterpret_cast<const float4*>(ep.stats_in + static_cast<size_t>(row0 + lane) * kStatSlots);
          float sx = 0.f, sy = 0.f;
#pragma unroll
          for (int i = 0; i < kStatSlots / 2; ++i) {
            const float4 v4 = __ldg(st4 + i);
            sx += v4.x;
            sy += v4.y;
            sx += v4.z;
            sy += v4.w;
          }
          mu = sx * (1.0f / kHidden);
          rstd = rsqrtf(fmaxf(sy * (1.0f / kHidden) - mu * mu, 0.f) + ep.eps);
        }
      }
      const float4* va4 = reinterpret_cast<const float4*>(ep.vec_a + col0);
      const float4* vb4 = reinterpret_cast<const float4*>(ep.vec_b + col0);
      if constexpr (kEpi == GEMM_EPI_RESID && !kHiLo) {
        // the residual input tile (32 rows x 32 fp32 columns) arrives by TMA, one chunk ahead of its use;
        // z is updated in place, so the output tensor map also describes the input
        if (store_ok && lane == 0) {
          mbar_expect_tx(&bars->zin[ew], kEpiBufBytes);
          tma_load_2d(&tm_out, &bars->zin[ew], zbuf, col0, row0);
        }
      }
      // HILO: this thread's 32 residual values of a chunk = 64 B of its row in each of the two bf16 planes
      uint4 zh[kHiLo ? 4 : 1], zl[kHiLo ? 4 : 1];
      __nv_bfloat16* hi_row = nullptr;
      __nv_bfloat16* lo_row = nullptr;
      if constexpr (kHiLo) {
        const size_t off = static_cast<size_t>(row0 + lane) * (static_cast<unsigned>(n_tiles) * BN) + col0;
        hi_row = ep.zb_out + off;
        lo_row = ep.z_lo + off;
      }
      // ACT_BWD: this thread's 32 activation-gradient factors of a chunk (64 B of its row), fetched one chunk ahead
      uint4 ubuf[kEpi == GEMM_EPI_ACT_BWD ? 2 : 1][4];
      const uint4* urow = nullptr;
      if constexpr (kEpi == GEMM_EPI_ACT_BWD) {
        urow = reinterpret_cast<const uint4*>(ep.act_in + static_cast<size_t>(row0 + lane) * (static_cast<unsigned>(n_tiles) * BN) + col0);
        if (store_ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j) ubuf[0][j] = __ldg(urow + j);
        }
      }
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + acc * BN + half * 128 + (static_cast<uint32_t>(quarter * 32) << 16);

      // 4 chunks of 32 accumulator columns; the TMEM load of chunk c+1 is in flight while chunk c is
      // converted and stored (register double buffer).
      uint32_t vbuf[2][32];
      tmem_ld_32x32(taddr, vbuf[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t(&v)[32] = vbuf[c & 1];
        if constexpr (kHiLo) {  // issued before the TMEM wait: in flight while the bias is added
          if (store_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              zh[j] = *reinterpret_cast<const uint4*>(hi_row + c * 32 + 8 * j);
              zl[j] = *reinterpret_cast<const uint4*>(lo_row + c * 32 + 8 * j);
            }
          }
        }
        tmem_ld_wait_regs(v);
        if (c + 1 < 4) tmem_ld_32x32(taddr + (c + 1) * 32, vbuf[(c + 1) & 1]);
        if constexpr (kEpi == GEMM_EPI_ACT_BWD) {
          if (c + 1 < 4 && store_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ubuf[(c + 1) & 1][j] = __ldg(urow + (c + 1) * 4 + j);
          }
        }
        float f[32];
        // NORM_A with a bf16 output: evaluated eight columns at a time inside the staging-store loop below
        constexpr bool kNormInStore = kEpi == GEMM_EPI_NORM_A && kOut == GEMM_OUT_BF16 && (kGelu == 0 || kGelu == 2);
        if constexpr (kNormInStore) {
        } else if constexpr (kEpi == GEMM_EPI_NORM_A) {
          // act(LN(z) W^T + b) from the raw product: rstd * (acc - mean * s[n]) + c[n]
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 s4 = __ldg(va4 + c * 8 + (j >> 2));
            const float4 c4 = __ldg(vb4 + c * 8 + (j >> 2));
            f[j + 0] = fmaf(rstd, fmaf(-mu, s4.x, __uint_as_float(v[j + 0])), c4.x);
            f[j + 1] = fmaf(rstd, fmaf(-mu, s4.y, __uint_as_float(v[j + 1])), c4.y);
            f[j + 2] = fmaf(rstd, fmaf(-mu, s4.z, __uint_as_float(v[j + 2])), c4.z);
            f[j + 3] = fmaf(rstd, fmaf(-mu, s4.w, __uint_as_float(v[j + 3])), c4.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = kBias ? __ldg(bias4 + c * 8 + (j >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
            f[j + 0] = __uint_as_float(v[j + 0]) + b4.x;
            f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
            f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
            f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
          }
        }
        if constexpr (kEpi == GEMM_EPI_RESID) {
          // z_new = x + branch, x = LN(z_prev) (or z_prev itself for the first layer of a stack); the row's
          // partial (sum, sum of squares) over these 32 columns feeds the next LayerNorm
          if (store_ok) {
            float4 zrow[8];
            if constexpr (kHiLo) {
              const uint32_t* hw = reinterpret_cast<const uint32_t*>(zh);
              const uint32_t* lw = reinterpret_cast<const uint32_t*>(zl);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[2 * j]));
                const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[2 * j + 1]));
                const float2 l0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[2 * j]));
                const float2 l1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[2 * j + 1]));
                zrow[j] = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
              }
            } else {
            mbar_wait(&bars->zin[ew], zphase);
            zphase ^= 1u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t addr = smem_u32(zbuf) + lane * 128 + ((static_cast<uint32_t>(j) ^ sw) << 4);
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(zrow[j].x), "=f"(zrow[j].y), "=f"(zrow[j].z), "=f"(zrow[j].w)
                           : "r"(addr)
                           : "memory");
            }
            __syncwarp();  // every lane has read its row: the buffer may be refilled
            if (c + 1 < 4 && lane == 0) {
              mbar_expect_tx(&bars->zin[ew], kEpiBufBytes);
              tma_load_2d(&tm_out, &bars->zin[ew], zbuf, col0 + (c + 1) * 32, row0);
            }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 x = zrow[j >> 2];
              if (ep.prev_norm) {
                const float4 g4 = __ldg(va4 + c * 8 + (j >> 2));
                const float4 b4 = __ldg(vb4 + c * 8 + (j >> 2));
                x.x = fmaf((x.x - mu) * rstd, g4.x, b4.x);
                x.y = fmaf((x.y - mu) * rstd, g4.y, b4.y);
                x.z = fmaf((x.z - mu) * rstd, g4.z, b4.z);
                x.w = fmaf((x.w - mu) * rstd, g4.w, b4.w);
              }
              f[j + 0] += x.x;
              f[j + 1] += x.y;
              f[j + 2] += x.z;
              f[j + 3] += x.w;
              psum += (f[j + 0] + f[j + 1]) + (f[j + 2] + f[j + 3]);
              psq += (f[j + 0] * f[j + 0] + f[j + 1] * f[j + 1]) + (f[j + 2] * f[j + 2] + f[j + 3] * f[j + 3]);
            }
          }
        }
        if constexpr (kEpi == GEMM_EPI_ACT_BWD) {
          // d u = d h * (gelu'(u) * dropout mask): the factor was stored by the forward FFN1 epilogue
          const uint32_t* gw = reinterpret_cast<const uint32_t*>(ubuf[c & 1]);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 gg = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[j >> 1]));
            f[j] *= gg.x;
            f[j + 1] *= gg.y;
          }
        }
        // GEMM_OUT_BF16_DUAL (training forward of linear1): act(u) and gelu'(u) in one evaluation, both under the
        // FFN-inner dropout mask of nn.TransformerEncoderLayer (element = row * N + column); the backward GEMM
        // (GEMM_EPI_ACT_BWD) multiplies by the second plane and never sees u. Evaluated eight columns at a time right
        // before their two 16-byte staging stores (below), so that only one group of results is live at a time.
        const unsigned long long dual_e0 =
            static_cast<unsigned long long>(row0 + lane) * (static_cast<unsigned>(n_tiles) * BN) + col0 + c * 32;
        // ONE code path (the epilogue is instruction-fetch sensitive: ncu showed 4 specialised copies of this block and
        // `no_instruction` stalls): pair indices are 32-bit — launch_gemm_tcgen05 rejects DUAL outputs of 2^33 elements
        // or more, where drop_bits would need its second hash round — and the mask is applied unconditionally
        // (thr16 = 0 keeps everything, scale = 1)
        const uint32_t dual_p0 = static_cast<uint32_t>(dual_e0 >> 1);
        auto dual_group = [&](int g8, uint32_t (&ap)[4], uint32_t (&gp)[4]) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = 8 * g8 + 2 * q;
            __half2 act, grad;
            gelu_erf_and_grad_fast_h2(__floats2half2_rn(f[j], f[j + 1]), act, grad);
            // drop_mul on both 16-bit fields at once: a per-halfword compare gives the keep mask, dropped halves are
            // zeroed in the fp16 domain, the (exact, fp32) scale follows the widening
            const uint32_t keep = __vcmpgeu2(lowbias32((dual_p0 + (j >> 1)) ^ drop.key), drop.thr16 * 0x00010001u);
            *reinterpret_cast<uint32_t*>(&act) &= keep;
            *reinterpret_cast<uint32_t*>(&grad) &= keep;
            const float scale = drop.scale;
            const float2 a2 = __half22float2(act), g2 = __half22float2(grad);
            ap[q] = pack_bf16x2(a2.x * scale, a2.y * scale);
            gp[q] = pack_bf16x2(g2.x * scale, g2.y * scale);
          }
        };
        if constexpr (kOut == GEMM_OUT_BF16_DUAL) {
          static_assert(kOut != GEMM_OUT_BF16_DUAL || kGelu == 2, "DUAL output: fast erf GELU");
        } else if (kGelu == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
        } else if (kGelu == 2) {
          if constexpr (kOut == GEMM_OUT_BF16) {
            // packed fp16 math (the output is bf16 anyway), evaluated eight columns at a time right before their
            // staging store below: fewer live registers, more independent work between the stores
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf_fast(f[j]);
          }
        } else if (kGelu == 3) {  // ReLU (nn.TransformerEncoderLayer default, appearance branch of CACNF)
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }

        if constexpr (kHiLo) {
          // the new residual value as two bf16 planes, 64 contiguous bytes per thread and plane, from registers
          if (store_ok) {
            uint4* dh = reinterpret_cast<uint4*>(hi_row + c * 32);
            uint4* dl = reinterpret_cast<uint4*>(lo_row + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              dh[j] = make_uint4(pack_bf16x2(f[8 * j + 0], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                 pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
              dl[j] = make_uint4(pack_bf16x2(bf16_residual(f[8 * j + 0]), bf16_residual(f[8 * j + 1])),
                                 pack_bf16x2(bf16_residual(f[8 * j + 2]), bf16_residual(f[8 * j + 3])),
                                 pack_bf16x2(bf16_residual(f[8 * j + 4]), bf16_residual(f[8 * j + 5])),
                                 pack_bf16x2(bf16_residual(f[8 * j + 6]), bf16_residual(f[8 * j + 7])));
            }
          }
        } else if (kOut == GEMM_OUT_F32 || kResidOut) {
          // 32 fp32 columns = 128 B per row -> one TMA store per chunk
          if (lane == 0) tma_store_wait_read0();  // previous store has finished reading ebuf
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t addr = row_smem + ((static_cast<uint32_t>(j) ^ sw) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(f[4 * j]),
                         "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                         : "memory");
          }
          if constexpr (kOut == GEMM_OUT_F32_BF16) {
            // bf16 copy (the next GEMM's A operand): two chunks fill one 128 B row of the third buffer; the
            // wait_read0 above (before the fp32 tile was overwritten) also covers its previous store
            const int hc = c & 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t addr = row_smem + kEpiBufBytes + ((static_cast<uint32_t>(hc * 4 + j) ^ sw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                           "r"(pack_bf16x2(f[8 * j + 0], f[8 * j + 1])), "r"(pack_bf16x2(f[8 * j + 2], f[8 * j + 3])),
                           "r"(pack_bf16x2(f[8 * j + 4], f[8 * j + 5])), "r"(pack_bf16x2(f[8 * j + 6], f[8 * j + 7]))
                           : "memory");
            }
          }
          if constexpr (kOut == GEMM_OUT_F32_BF16_DIRECT) {
            // bf16 copy of the row: 64 contiguous bytes per thread, written straight from registers
            if (store_ok) {
              uint4* dst = reinterpret_cast<uint4*>(
                  ep.zb_out + static_cast<size_t>(row0 + lane) * (static_cast<unsigned>(n_tiles) * BN) + col0 + c * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j] = make_uint4(pack_bf16x2(f[8 * j + 0], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                    pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
              if constexpr (kTerms == 3) {  // fp32-parity mode: the remainder plane of the split
                uint4* dlo = reinterpret_cast<uint4*>(
                    ep.zb_lo_out + static_cast<size_t>(row0 + lane) * (static_cast<unsigned>(n_tiles) * BN) + col0 + c * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  dlo[j] = make_uint4(pack_bf16x2(bf16_residual(f[8 * j + 0]), bf16_residual(f[8 * j + 1])),
                                      pack_bf16x2(bf16_residual(f[8 * j + 2]), bf16_residual(f[8 * j + 3])),
                                      pack_bf16x2(bf16_residual(f[8 * j + 4]), bf16_residual(f[8 * j + 5])),
                                      pack_bf16x2(bf16_residual(f[8 * j + 6]), bf16_residual(f[8 * j + 7])));
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && store_ok) {
            if (kStreamK) tma_reduce_add_2d(&tm_out, ebuf, col0 + c * 32, row0);
            else tma_store_2d(&tm_out, ebuf, col0 + c * 32, row0);
            if constexpr (kOut == GEMM_OUT_F32_BF16) {
              if (c & 1) tma_store_2d(&tm_out2, ebuf + kEpiBufBytes, col0 + (c - 1) * 32, row0);
            }
            tma_store_commit();
          }
        } else {
          // bf16 output: two 32-column chunks fill one 128 B row (64 bf16) -> store every 2nd chunk
          const int hc = c & 1;
          if (hc == 0) {
            if (lane == 0) tma_store_wait_read0();
            __syncwarp();
          }
          if constexpr (kOut == GEMM_OUT_BF16_DUAL) {
            // two groups of eight columns are evaluated before their four staging stores: eight independent
            // polynomial / MUFU chains in flight per thread (the stores are compiler barriers)
#pragma unroll
            for (int jj = 0; jj < 4; jj += 2) {
              uint32_t ap[2][4], gp[2][4];
#pragma unroll
              for (int u = 0; u < 2; ++u) dual_group(jj + u, ap[u], gp[u]);
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const uint32_t addr = row_smem + ((static_cast<uint32_t>(hc * 4 + jj + u) ^ sw) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(ap[u][0]), "r"(ap[u][1]),
                             "r"(ap[u][2]), "r"(ap[u][3])
                             : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + kEpiBufBytes), "r"(gp[u][0]),
                             "r"(gp[u][1]), "r"(gp[u][2]), "r"(gp[u][3])
                             : "memory");
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t addr = row_smem + ((static_cast<uint32_t>(hc * 4 + j) ^ sw) << 4);
            if constexpr (kOut == GEMM_OUT_BF16_DUAL) {
            } else if constexpr (kNormInStore || (kOut == GEMM_OUT_BF16 && kGelu == 2)) {
              float gv[8];
              if constexpr (kNormInStore) {
                // act(LN(z) W^T + b) from the raw product: rstd * (acc - mean * s[n]) + c[n]
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  const float4 s4 = __ldg(va4 + c * 8 + 2 * j + q);
                  const float4 c4 = __ldg(vb4 + c * 8 + 2 * j + q);
                  const int e = 8 * j + 4 * q;
                  gv[4 * q + 0] = fmaf(rstd, fmaf(-mu, s4.x, __uint_as_float(v[e + 0])), c4.x);
                  gv[4 * q + 1] = fmaf(rstd, fmaf(-mu, s4.y, __uint_as_float(v[e + 1])), c4.y);
                  gv[4 * q + 2] = fmaf(rstd, fmaf(-mu, s4.z, __uint_as_float(v[e + 2])), c4.z);
                  gv[4 * q + 3] = fmaf(rstd, fmaf(-mu, s4.w, __uint_as_float(v[e + 3])), c4.w);
                }
              } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) gv[q] = f[8 * j + q];
              }
              uint32_t ap[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if constexpr (kGelu == 2) {
                  const float2 r = __half22float2(gelu_erf_fast_h2(__floats2half2_rn(gv[2 * q], gv[2 * q + 1])));
                  ap[q] = pack_bf16x2(r.x, r.y);
                } else {
                  ap[q] = pack_bf16x2(gv[2 * q], gv[2 * q + 1]);
                }
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(ap[0]), "r"(ap[1]), "r"(ap[2]),
                           "r"(ap[3])
                           : "memory");
            } else {
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                           "r"(pack_bf16x2(f[8 * j + 0], f[8 * j + 1])),
                           "r"(pack_bf16x2(f[8 * j + 2], f[8 * j + 3])),
                           "r"(pack_bf16x2(f[8 * j + 4], f[8 * j + 5])),
                           "r"(pack_bf16x2(f[8 * j + 6], f[8 * j + 7]))
                           : "memory");
            }
            if (kOut == GEMM_OUT_BF16_SPLIT) {
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + kEpiBufBytes),
                           "r"(pack_bf16x2(bf16_residual(f[8 * j + 0]), bf16_residual(f[8 * j + 1]))),
                           "r"(pack_bf16x2(bf16_residual(f[8 * j + 2]), bf16_residual(f[8 * j + 3]))),
                           "r"(pack_bf16x2(bf16_residual(f[8 * j + 4]), bf16_residual(f[8 * j + 5]))),
                           "r"(pack_bf16x2(bf16_residual(f[8 * j + 6]), bf16_residual(f[8 * j + 7])))
                           : "memory");
            }
          }
          if (hc == 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if constexpr (kEpi == GEMM_EPI_ACT_BWD) {
              // bias gradient of linear1: column sums of the ROUNDED values the weight-gradient GEMM will read, taken
              // from the staging tile (lane l owns columns 2l, 2l + 1 of this 64-column slab; a row is one
              // conflict-free 128-byte read)
              int nvalid = ep.valid_rows - row0;
              nvalid = nvalid < 0 ? 0 : (nvalid > 32 ? 32 : nvalid);
              if (store_ok && nvalid > 0 && ep.colsum_out != nullptr) {
                float sx = 0.f, sy = 0.f;
                const uint32_t cbase = smem_u32(ebuf) + (lane & 3) * 4;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                  if (r < nvalid) {
                    uint32_t wv;
                    asm volatile("ld.shared.b32 %0, [%1];"
                                 : "=r"(wv)
                                 : "r"(cbase + r * 128 + (((static_cast<uint32_t>(lane) >> 2) ^ (r & 7)) << 4))
                                 : "memory");
                    const float2 v2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv));
                    sx += v2.x;
                    sy += v2.y;
                  }
                }
                float* dst = ep.colsum_out + col0 + (c - 1) * 32 + 2 * lane;
                atomicAdd(dst, sx);
                atomicAdd(dst + 1, sy);
              }
            }
            if (lane == 0 && store_ok) {
              tma_store_2d(&tm_out, ebuf, col0 + (c - 1) * 32, row0);
              if (kOut == GEMM_OUT_BF16_SPLIT || kOut == GEMM_OUT_BF16_DUAL)
                tma_store_2d(&tm_out, ebuf + kEpiBufBytes, col0 + (c - 1) * 32,
                             out_plane_rows + row0);
              tma_store_commit();
            }
          }
        }
      }
      if constexpr (kEpi == GEMM_EPI_RESID) {
        if (store_ok)  // this thread's row, this warp's 128-column slab -> its own slot
          ep.stats_out[static_cast<size_t>(row0 + lane) * kStatSlots + n_blk * 2 + half] = make_float2(psum, psq);
      }
      // all TMEM reads of this accumulator have completed (last tmem_ld_wait_regs) -> release it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&bars->tmem_empty[acc], 0));
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA exits while its peer may still signal its barriers / write its smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
int gemm_smem_bytes() { return Cfg<GEMM_OUT_F32>::kSmemBytes; }

template <int kTerms, int kOut, int kGelu, int kLayout = GEMM_NT, int kEpi = GEMM_EPI_PLAIN>
static cudaError_t launch_one(const GemmArgs& g, cudaStream_t stream, int num_sms) {
  auto kern = gemm_tcgen05_kernel<kTerms, kOut, kGelu, kLayout, kEpi>;
  constexpr int smem = Cfg<kOut>::kSmemBytes;
  static unsigned long long smem_done = 0;  // per instantiation, one bit per device
  {
    cudaError_t e = ensure_dynamic_smem(kern, smem, &smem_done);
    if (e != cudaSuccess) return e;
  }
  const int m_tiles = g.m_rows / BM;
  const int n_tiles = g.n / BN;
  const int pair_tiles = ((m_tiles + kCluster - 1) / kCluster) * n_tiles;
  const int max_clusters = num_sms / kCluster;
  const int k_blocks = (g.k + BK - 1) / BK;
  int clusters = pair_tiles < max_clusters ? pair_tiles : max_clusters;
  if (kLayout == GEMM_TN_RED) {  // split-K: one (tile, k-slice) unit per pair, see Sched
    int slices = max_clusters / pair_tiles;
    if (slices < 1) slices = 1;
    if (slices > k_blocks) slices = k_blocks;
    const int units = pair_tiles * slices;
    clusters = units < max_clusters ? units : max_clusters;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * kCluster);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, g.tm_a, g.tm_b, g.tm_out, g.tm_out2, g.bias, m_tiles, n_tiles,
                            k_blocks, g.a_plane_rows, g.b_plane_rows, g.out_plane_rows, g.drop, g.epi);
}

cudaError_t launch_gemm_tcgen05(const GemmArgs& g, cudaStream_t stream, int num_sms) {
  if (g.m_rows <= 0 || g.m_rows % BM != 0 || g.n % BN != 0 || g.k <= 0) return cudaErrorInvalidValue;
  // the reduction of the weight-gradient layout may be ragged (TMA zero-fills rows past the end)
  if (g.layout != GEMM_TN_RED && g.k % BK != 0) return cudaErrorInvalidValue;
  if (g.layout == GEMM_NN) {
    if (g.terms == 1 && g.gelu == 0 && g.out_kind == GEMM_OUT_F32)
      return launch_one<1, GEMM_OUT_F32, 0, GEMM_NN>(g, stream, num_sms);
    if (g.epilogue == GEMM_EPI_ACT_BWD) {
      if (g.terms == 1 && g.gelu == 0 && g.out_kind == GEMM_OUT_BF16 && g.epi.act_in != nullptr)
        return launch_one<1, GEMM_OUT_BF16, 0, GEMM_NN, GEMM_EPI_ACT_BWD>(g, stream, num_sms);
      return cudaErrorInvalidValue;
    }
    if (g.terms == 1 && g.gelu == 0 && g.out_kind == GEMM_OUT_BF16)
      return launch_one<1, GEMM_OUT_BF16, 0, GEMM_NN>(g, stream, num_sms);
    return cudaErrorInvalidValue;
  }
  if (g.layout == GEMM_TN_RED) {
    if (g.terms == 1 && g.gelu == 0 && g.out_kind == GEMM_OUT_F32)
      return launch_one<1, GEMM_OUT_F32, 0, GEMM_TN_RED>(g, stream, num_sms);
    return cudaErrorInvalidValue;
  }
  if (g.layout != GEMM_NT) return cudaErrorInvalidValue;
  if (g.epilogue == GEMM_EPI_NORM_A) {
    if (g.terms == 1 && g.out_kind == GEMM_OUT_BF16 && g.gelu == 0)
      return launch_one<1, GEMM_OUT_BF16, 0, GEMM_NT, GEMM_EPI_NORM_A>(g, stream, num_sms);
    if (g.terms == 1 && g.out_kind == GEMM_OUT_BF16 && g.gelu == 2)
      return launch_one<1, GEMM_OUT_BF16, 2, GEMM_NT, GEMM_EPI_NORM_A>(g, stream, num_sms);
    // fp32-parity mode: split operands (the NORM_A algebra is linear in the accumulator), split output, exact erf GELU
    if (g.terms == 3 && g.out_kind == GEMM_OUT_BF16_SPLIT && g.gelu == 0)
      return launch_one<3, GEMM_OUT_BF16_SPLIT, 0, GEMM_NT, GEMM_EPI_NORM_A>(g, stream, num_sms);
    if (g.terms == 3 && g.out_kind == GEMM_OUT_BF16_SPLIT && g.gelu == 1)
      return launch_one<3, GEMM_OUT_BF16_SPLIT, 1, GEMM_NT, GEMM_EPI_NORM_A>(g, stream, num_sms);
    return cudaErrorInvalidValue;
  }
  if (g.epilogue == GEMM_EPI_RESID) {
    if (g.terms == 1 && g.out_kind == GEMM_OUT_F32_BF16 && g.gelu == 0 && g.n == kHidden)
      return launch_one<1, GEMM_OUT_F32_BF16, 0, GEMM_NT, GEMM_EPI_RESID>(g, stream, num_sms);
    if (g.terms == 1 && g.out_kind == GEMM_OUT_F32_BF16_DIRECT && g.gelu == 0 && g.n == kHidden)
      return launch_one<1, GEMM_OUT_F32_BF16_DIRECT, 0, GEMM_NT, GEMM_EPI_RESID>(g, stream, num_sms);
    if (g.terms == 3 && g.out_kind == GEMM_OUT_F32_BF16_DIRECT && g.gelu == 0 && g.n == kHidden && g.epi.zb_lo_out)
      return launch_one<3, GEMM_OUT_F32_BF16_DIRECT, 0, GEMM_NT, GEMM_EPI_RESID>(g, stream, num_sms);
    if (g.terms == 1 && g.out_kind == GEMM_OUT_HILO && g.gelu == 0 && g.n == kHidden && g.epi.zb_out && g.epi.z_lo)
      return launch_one<1, GEMM_OUT_HILO, 0, GEMM_NT, GEMM_EPI_RESID>(g, stream, num_sms);
    return cudaErrorInvalidValue;
  }
  if (g.epilogue != GEMM_EPI_PLAIN) return cudaErrorInvalidValue;
#define STLT_GEMM_CASE(T, O, G) \
  if (g.terms == T && g.out_kind == O && g.gelu == G) return launch_one<T, O, G>(g, stream, num_sms);
  STLT_GEMM_CASE(1, GEMM_OUT_F32, 0)
  STLT_GEMM_CASE(1, GEMM_OUT_BF16, 0)
  STLT_GEMM_CASE(1, GEMM_OUT_BF16, 1)
  STLT_GEMM_CASE(1, GEMM_OUT_BF16, 2)
  STLT_GEMM_CASE(3, GEMM_OUT_F32, 0)
  STLT_GEMM_CASE(3, GEMM_OUT_BF16_SPLIT, 1)
  STLT_GEMM_CASE(3, GEMM_OUT_BF16_SPLIT, 0)
  // the dropout mask of the DUAL epilogue hashes 32-bit pair indices (drop_bits without its second round)
  if (g.out_kind == GEMM_OUT_BF16_DUAL && (static_cast<unsigned long long>(g.m_rows) * g.n >> 33) != 0) return cudaErrorInvalidValue;
  STLT_GEMM_CASE(1, GEMM_OUT_BF16_DUAL, 2)
  STLT_GEMM_CASE(1, GEMM_OUT_BF16, 3)
  STLT_GEMM_CASE(3, GEMM_OUT_BF16_SPLIT, 3)
#undef STLT_GEMM_CASE
  return cudaErrorInvalidValue;
}

}  // namespace stlt
