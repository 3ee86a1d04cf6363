// In-projection GEMM with the masked self-attention of the short STLT sequences fused into its epilogue
// (bf16 inference path): the packed Q | K | V activations never reach HBM.
//
//   ctx[M, 768] = concat_h softmax(Q_h K_h^T / 8 + mask) V_h,   [Q | K | V] = act_in W_in^T + b_in
//
// Replaces F.linear(x, in_proj_weight, in_proj_bias) + scaled-dot-product attention inside the
// nn.MultiheadAttention of every encoder layer the reference instantiates at src/modelling/models.py:46-55
// (spatial: sequences of S object slots, key-padding mask = categories == 0, datasets.py:274-278) and
// :118-128 (temporal: sequences of L frames, causal mask src/utils/model_utils.py:4-7 or key padding
// frame_types == 0, datasets.py:280-286). SURVEY.md K2 + K3.
//
// Unfused, one spatial layer at B = 4096 writes 1.6 GB of QKV, reads it back in the attention kernel and
// writes 0.53 GB of context; here only the context leaves the SM.
//
// Work unit: one (block of rows, head). The in-projection weights are packed HEAD-MAJOR (row h*192 + t*64 + j
// = row t*768 + h*64 + j of in_proj_weight, t = Q, K, V), so the 192 output columns of a unit are exactly
// [Q_h | K_h | V_h] of its rows. CTA pairs (cluster of 2, tcgen05 cta_group::2) run ONE M=256 x N=192 x K=16
// MMA stream per unit; each CTA stages its own 128 activation rows and 96 of the 192 weight rows by TMA
// (28 KB per stage, 5 stages). The accumulators are double buffered in TMEM (2 x 192 columns), so the epilogue
// of unit i overlaps the MMAs of unit i+1. A cluster sweeps the 12 heads of a pair of row blocks back to back
// (the activation rows are re-read from L2, the per-row epilogue state is loaded once per pair).
//
// Rows: sequences must not straddle row blocks, so a CTA's block holds R = floor(128 / T) * T rows (whole
// sequences; 125 of 128 at T = 5, 119 at T = 17) and starts at row block * R. The TMA box still loads 128
// rows; the surplus rows belong to the next block, are computed and never stored (the TMA store box has R rows).
//
// Epilogue (8 warps per CTA, 128 rows x 192 columns of this CTA's accumulator):
//   1. tcgen05.ld -> bias (or the deferred LayerNorm of the input, GEMM_EPI_NORM_A algebra: the A operand is
//      the un-normalised bf16 residual stream and W is pre-multiplied by gamma) -> bf16 -> three XOR-swizzled
//      [128 x 64] shared-memory tiles Q, K, V (rows past the real tokens as zeros). The TMEM buffer is released
//      as soon as it has been read.
//   2. attention in 16-query tiles on mma.sync.m16n8k16 (attend_tile): for T <= 16 a tile holds whole sequences, so its
//      keys are its own 16 rows; for longer sequences warp w owns rows 16w .. 16w+15 and attends a band of 8 * kNT key
//      rows that covers every sequence touching them. Same-sequence, key-padding and causal predicates on the
//      accumulator fragments, fp32 softmax with quad shuffles, O = P V with P from registers.
//   3. the normalised context goes to a fourth staging tile and one thread issues the TMA store of its first R
//      rows to ctx[:, 64 h : 64 h + 64]; two CTA-wide named barriers per unit (tiles complete / tiles free).
#include "kernels.h"
#include "mma_tiles.cuh"

namespace stlt {

namespace {

constexpr int BM = 128;
constexpr int BN = 192;  // Q_h | K_h | V_h
constexpr int BK = 64;
constexpr int kABytes = BM * BK * 2;            // 16 KiB
constexpr int kBHalfBytes = (BN / 2) * BK * 2;  // 12 KiB: this CTA's 96 weight rows
constexpr int kStageBytes = kABytes + kBHalfBytes;
constexpr int kStages = 5;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kTmemCols = 512;
constexpr int kCluster = 2;
constexpr uint16_t kClusterMask = (1u << kCluster) - 1;
constexpr int kTileBytes = BM * 128;  // one of the Q / K / V / context tiles: 128 rows x 64 bf16
constexpr int kVecBytes = 2 * kQkv * 4;  // the epilogue vectors s[2304], c[2304] (head-major), staged once per CTA
constexpr int kSmemBytes = kStages * kStageBytes + 4 * kTileBytes + kVecBytes + 256 /*barriers*/;
static_assert(kSmemBytes <= 232448, "smem budget");
static_assert(kStageBytes % 1024 == 0, "SWIZZLE_128B tiles need 1024 B alignment");

struct __align__(8) Barriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Masked attention of one tile of 16 query rows [q_row0, q_row0 + 16) of the Q / K / V tiles against the 8 * kNT key rows
// from band0: S = Q K^T on mma.sync.m16n8k16, predicates ok[h] (bit nt*2+e = key band0 + nt*8 + 2t + e may be attended by
// query row q_row0 + g + 8h) on the accumulator fragments, fp32 softmax with quad shuffles, O = P V with P from registers;
// the normalised context rows below row_limit (tile-local) go to the staging tile. Row indices are clamped to the 128-row
// tiles (keys past the end are masked by the caller).
template <int kNT>
__device__ __forceinline__ void attend_tile(uint32_t q_base, uint32_t k_base, uint32_t v_base, uint32_t c_base, int q_row0,
                                            int band0, uint32_t ok0, uint32_t ok1, int row_limit, int lane) {
  static_assert(kNT % 2 == 0, "P V consumes the keys 16 at a time");
  const int g = lane >> 2, t = lane & 3;
  const float kScale = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  auto clampr = [](int r) { return r > BM - 1 ? BM - 1 : r; };
  float s[kNT][4];
#pragma unroll
  for (int nt = 0; nt < kNT; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t a[4];
    ldmatrix_x4(tile_addr(q_base, clampr(q_row0 + (lane & 7) + ((lane >> 3) & 1) * 8), kt * 2 + (lane >> 4)), a);
#pragma unroll
    for (int np = 0; np < kNT / 2; ++np) {
      uint32_t b[4];
      ldmatrix_x4(tile_addr(k_base, clampr(band0 + (np * 2 + (lane >> 4)) * 8 + (lane & 7)), kt * 2 + ((lane >> 3) & 1)), b);
      mma_bf16(s[np * 2 + 0], a, b[0], b[1]);
      mma_bf16(s[np * 2 + 1], a, b[2], b[3]);
    }
  }
  // ---- masked softmax on the accumulator fragments (fp32) ----
  float inv_sum[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t ok = h == 0 ? ok0 : ok1;
    float m = -INFINITY;  // maximum of the raw scores (the scale 1/8 * log2 e > 0 is folded into the exponent below)
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float v = ((ok >> (nt * 2 + e)) & 1u) ? s[nt][2 * h + e] : -INFINITY;
        s[nt][2 * h + e] = v;
        m = fmaxf(m, v);
      }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    const float mm = (m == -INFINITY) ? 0.f : m * kScale;  // fully masked row -> all-zero probabilities
    float sum = 0.f;
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float pv;  // 2^(-inf) = 0 for masked keys
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv) : "f"(fmaf(s[nt][2 * h + e], kScale, -mm)));
        s[nt][2 * h + e] = pv;
        sum += pv;
      }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    inv_sum[h] = sum > 0.f ? __frcp_rn(sum) : 0.f;
  }
  // ---- O = P V ----
  float o[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[dt][i] = 0.f;
#pragma unroll
  for (int j = 0; j < kNT / 2; ++j) {
    uint32_t pa[4];
    pa[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
    pa[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
    pa[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
    pa[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldmatrix_x4_trans(tile_addr(v_base, clampr(band0 + j * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)), dp * 2 + (lane >> 4)), b);
      mma_bf16(o[dp * 2 + 0], pa, b[0], b[1]);
      mma_bf16(o[dp * 2 + 1], pa, b[2], b[3]);
    }
  }
  // ---- normalised context -> the tile's rows of the staging tile ----
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int lr = g + 8 * h;
    const int row = q_row0 + lr;
    if (lr < row_limit && row < BM) {
      const float is = inv_sum[h];
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        const uint32_t v = pack_bf16x2(o[dt][2 * h] * is, o[dt][2 * h + 1] * is);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(c_base, row, dt) + 4 * t), "r"(v) : "memory");
      }
    }
  }
}

template <int kNT, int kBandOff, bool kAligned>
__global__ void __launch_bounds__(kThreads, 1)
qkv_attention_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                     const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_out1,
                     QkvAttnArgs p) {
  // Two ways to cut the 128 rows of a block into 16-query tiles:
  //  kAligned (T <= 16): a tile holds floor(16 / T) WHOLE sequences (15 rows at T = 5, 11 at T = 11), so its keys are its own
  //    rows (kNT = 2); a block has ceil(sequences / floor(16 / T)) tiles, dealt round-robin to the 8 warps (at most 2 each);
  //  band (17 <= T <= 32): warp w owns query rows 16w .. 16w+15 and attends a band of 8 * kNT key rows from 16w - kBandOff
  //    that covers every sequence touching them (causal launches need no keys after the tile: narrower band).
  static_assert(kAligned ? (kNT == 2 && kBandOff == 0) : (kNT == 4 || kNT == 6 || kNT == 10), "tiling");
  constexpr int kBand = 8 * kNT;
  constexpr int kMaskWords = (kBand + 31) / 32;

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_ab = smem;
  uint8_t* smem_qkv = smem + kStages * kStageBytes;
  uint8_t* smem_ctx = smem_qkv + 3 * kTileBytes;  // staging tile of the context store
  float* smem_vec = reinterpret_cast<float*>(smem_ctx + kTileBytes);  // s[2304] then c[2304]
  Barriers* bars = reinterpret_cast<Barriers*>(smem_ctx + kTileBytes + kVecBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x / kCluster;
  const int num_clusters = gridDim.x / kCluster;
  // a cluster sweeps the 12 heads of one pair of row blocks before it moves on: the activation rows are re-read from
  // L2 by the same SM, and the epilogue's per-row state (LayerNorm statistics, key-padding bits) is loaded once per pair
  // Row blocks: [0, nb_full) hold whole sequences of T tokens (R rows each, from row 0); with the pad-skipping layout
  // (p.dyn, compact.cu) the counts come from the device and blocks [nb_full, nb_full + nb_single) hold one-token
  // sequences, 128 per block, from row single_row0.
  int nb_full = p.row_blocks, nb_single = 0;
  long long n_full_seq = p.valid_rows / p.seq_len, n_single = 0, single_row0 = 0;
  if (p.dyn != nullptr) {
    n_full_seq = __ldg(p.dyn + kDynFull);
    n_single = __ldg(p.dyn + kDynSingle);
    nb_full = __ldg(p.dyn + kDynFullBlocks);
    single_row0 = __ldg(p.dyn + kDynSingleRow0);
    nb_single = __ldg(p.dyn + kDynSingleBlocks);
  }
  const int total_blocks = nb_full + nb_single;
  const int pair_blocks = (total_blocks + kCluster - 1) / kCluster;
  constexpr int k_blocks = kHidden / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_out);
    tma_prefetch_desc(&tm_out1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], kCluster * kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(&bars->tmem_base, kTmemCols);
    tmem_relinquish_pair();
  }
  if (warp >= 4) {  // every lane of an epilogue warp reads the same s / c entries: from shared memory, not through L1
    const float4* gs = reinterpret_cast<const float4*>(p.vec_s);
    const float4* gc = reinterpret_cast<const float4*>(p.vec_c);
    float4* ds = reinterpret_cast<float4*>(smem_vec);
    for (int i = threadIdx.x - 128; i < kQkv / 4; i += kEpiWarps * 32) {
      ds[i] = __ldg(gs + i);
      ds[kQkv / 4 + i] = __ldg(gc + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = cluster_id; pb < pair_blocks; pb += num_clusters) {
        const int pblk = pb * kCluster + cta_rank;  // rows past the end of the tensor are zero-filled
        const int a_row = pblk < nb_full ? pblk * p.rows_per_block : static_cast<int>(single_row0) + (pblk - nb_full) * BM;
        for (int head = 0; head < kHeads; ++head) {
          const int b_row = head * BN + cta_rank * (BN / kCluster);
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&bars->empty[stage], phase ^ 1u);
            if (cta_rank == 0) mbar_expect_tx(&bars->full[stage], kCluster * kStageBytes);
            uint8_t* sa = smem_ab + stage * kStageBytes;
            const uint32_t full_leader = mapa_u32(&bars->full[stage], 0);
            tma_load_2d_pair(&tm_a, full_leader, sa, kb * BK, a_row);
            tma_load_2d_pair(&tm_b, full_leader, sa + kABytes, kb * BK, b_row);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kCluster * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int pb = cluster_id; pb < pair_blocks; pb += num_clusters) {
        for (int head = 0; head < kHeads; ++head, ++it) {
          const int acc = it & 1;
          const uint32_t acc_phase = (it >> 1) & 1;
          mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BN;
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem_ab + stage * kStageBytes);
            const uint64_t da = umma_desc_k_sw128(sa);
            const uint64_t db = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb != 0 || k != 0) ? 1u : 0u);
            umma_commit_pair(&bars->empty[stage], kClusterMask);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit_pair(&bars->tmem_full[acc], kClusterMask);
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue: QKV tile -> attention ===============
    const int ew = warp - 4;       // also the 16-row query tile this warp attends
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = ew >> 2;      // accumulator columns [96 * half, 96 * half + 96)
    const int g = lane >> 2, t = lane & 3;
    const int T = p.seq_len;
    const int R = p.rows_per_block;
    const uint32_t q_base = smem_u32(smem_qkv);
    const uint32_t k_base = q_base + kTileBytes;
    const uint32_t v_base = k_base + kTileBytes;
    const uint32_t c_base = smem_u32(smem_ctx);
    const int arow = quarter * 32 + lane;  // accumulator row (block-local) of this thread
    int band0 = 16 * ew - kBandOff;        // band scheme: first key row of this warp's band
    band0 = band0 < 0 ? 0 : (band0 > BM - kBand ? BM - kBand : band0);
    // aligned scheme: whole sequences per 16-row tile
    const int seqs_per_tile = kAligned ? 16 / T : 0;
    const int tile_rows = kAligned ? seqs_per_tile * T : 16;

    // static part of the mask (block and tile origins are multiples of T): bit nt*2+e of allow[h] = the key in this thread's
    // column nt*8 + 2t + e belongs to the sequence of the query in fragment row g + 8h (and is not in its future when causal)
    uint32_t allow[2], allow1[2];  // sequences of T tokens / one-token sequences (a query attends to its own row only)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = (kAligned ? 0 : 16 * ew) + g + 8 * h;  // tile-local (aligned) or block-local (band) row
      const int lim = kAligned ? tile_rows : R;
      uint32_t bits = 0;
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = (kAligned ? 0 : band0) + nt * 8 + 2 * t + e;
          const bool ok = q < lim && key < lim && key / T == q / T && (!p.causal || key <= q);
          bits |= (ok ? 1u : 0u) << (nt * 2 + e);
        }
      allow[h] = bits;
      const int d = q - (kAligned ? 0 : band0) - 2 * t;  // column index of the query's own row
      allow1[h] = (d >= 0 && d < kBand && (d & 7) < 2) ? 1u << ((d >> 3) * 2 + (d & 7)) : 0u;
    }
    const int seqs_per_block = R / T;
    const int tiles_full = kAligned ? (seqs_per_block + seqs_per_tile - 1) / seqs_per_tile : kEpiWarps;

    int it = 0;
    for (int pb = cluster_id; pb < pair_blocks; pb += num_clusters) {
      const int blk = pb * kCluster + cta_rank;
      const bool live = blk < total_blocks;  // odd block counts: the last pair has a dummy half
      const bool single = blk >= nb_full;
      const long long row0 = single ? single_row0 + static_cast<long long>(blk - nb_full) * BM : static_cast<long long>(blk) * R;
      // is block-local row `r` a real token? (the rest of the 128-row box belongs to the next block or is padding)
      auto row_live = [&](int r) -> bool {
        if (single) return static_cast<long long>(blk - nb_full) * BM + r < n_single;
        return r < R && static_cast<long long>(blk) * seqs_per_block + r / T < n_full_seq;
      };
      // rows that are no real token hold whatever the workspace held: written as zeros (0 * NaN would poison P V)
      const bool real_row = live && row_live(arow);

      // ---- per row block: deferred LayerNorm statistics of this thread's row, key-padding bits of the band ----
      float mu = 0.f, rstd = 1.f;
      if (p.prev_norm && real_row && !(p.debug & 4)) {
        const float4* st4 = reinterpret_cast<const float4*>(p.stats_in + (row0 + arow) * kStatSlots);
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
        rstd = rsqrtf(fmaxf(sy * (1.0f / kHidden) - mu * mu, 0.f) + p.eps);
      }
      // key-padding bits of this warp's tiles (aligned: tiles ew and ew + 8; band: the band of rows 16 ew ..), lane = key row
      const int tile_rows_b = kAligned ? (single ? 16 : tile_rows) : 16;
      const int tiles_b = kAligned ? (single ? kEpiWarps : tiles_full) : kEpiWarps;
      uint32_t okm[kAligned ? 2 : 1][2];
#pragma unroll
      for (int ti = 0; ti < (kAligned ? 2 : 1); ++ti) {
        const int key0 = kAligned ? (ew + ti * kEpiWarps) * tile_rows_b : band0;
        uint32_t kw[kMaskWords];
#pragma unroll
        for (int i = 0; i < kMaskWords; ++i) {
          const int key = key0 + i * 32 + lane;
          const bool in = live && i * 32 + lane < kBand && key < BM && row_live(key);
          const bool ok = in && __ldg(p.mask_src + (in ? row0 + key : 0)) != 0;
          kw[i] = __ballot_sync(0xffffffffu, ok);
        }
        uint32_t kbits = 0;  // this thread's 2 * kNT key columns
#pragma unroll
        for (int nt = 0; nt < kNT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int idx = nt * 8 + 2 * t + e;
            kbits |= ((kw[idx >> 5] >> (idx & 31)) & 1u) << (nt * 2 + e);
          }
        okm[ti][0] = (single ? allow1[0] : allow[0]) & kbits;
        okm[ti][1] = (single ? allow1[1] : allow[1]) & kbits;
      }

      for (int head = 0; head < kHeads; ++head, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * BN + half * 96 + (static_cast<uint32_t>(quarter * 32) << 16);
        const float4* vs4 = reinterpret_cast<const float4*>(smem_vec + head * BN + half * 96);
        const float4* vc4 = reinterpret_cast<const float4*>(smem_vec + kQkv + head * BN + half * 96);

        // ---- 1. accumulator -> bias / deferred LayerNorm -> bf16 -> Q / K / V tiles (every warp has left the
        //         attention phase of the previous unit: barrier B below) ----
        uint32_t vbuf[2][32];
        tmem_ld_32x32(taddr, vbuf[0]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t(&v)[32] = vbuf[c & 1];
          tmem_ld_wait_regs(v);
          if (c + 1 < 3) tmem_ld_32x32(taddr + (c + 1) * 32, vbuf[(c + 1) & 1]);
          const int col = half * 96 + c * 32;  // a 32-column chunk never straddles two of the 64-wide tiles
          const uint32_t base = q_base + (col >> 6) * kTileBytes;
          const int chunk0 = (col & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // 8 columns = one 16-byte store
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int e = 8 * j + 4 * q;
              // one code path: without a pending LayerNorm mu = 0, rstd = 1 (exact: acc * 1 - 0 * s + c)
              const float4 c4 = vc4[c * 8 + 2 * j + q];
              const float4 s4 = vs4[c * 8 + 2 * j + q];
              const float f0 = fmaf(rstd, fmaf(-mu, s4.x, __uint_as_float(v[e + 0])), c4.x);
              const float f1 = fmaf(rstd, fmaf(-mu, s4.y, __uint_as_float(v[e + 1])), c4.y);
              const float f2 = fmaf(rstd, fmaf(-mu, s4.z, __uint_as_float(v[e + 2])), c4.z);
              const float f3 = fmaf(rstd, fmaf(-mu, s4.w, __uint_as_float(v[e + 3])), c4.w);
              pk[2 * q] = real_row ? pack_bf16x2(f0, f1) : 0u;
              pk[2 * q + 1] = real_row ? pack_bf16x2(f2, f3) : 0u;
            }
            if (live && !(p.debug & 2))
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_addr(base, arow, chunk0 + j)),
                           "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                           : "memory");
          }
        }
        // every TMEM read of this accumulator has completed -> the MMAs of the unit after next may overwrite it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(&bars->tmem_empty[acc], 0));
        if (!live) continue;  // CTA-uniform: the dummy half only takes part in the TMEM handshake

        // the context store of the previous unit was issued a whole phase ago: it has long finished reading the
        // staging tile, which every warp rewrites after barrier A
        if (ew == 0 && lane == 0) tma_store_wait_read0();
        named_bar_sync(1, kEpiWarps * 32);  // A: Q / K / V tiles complete

        if (p.debug & 8) __nanosleep(2000);  // decomposition: a 2 us stall here does not change the kernel time
        if (!(p.debug & 1)) {
          // ---- 2. / 3. attention of this warp's tile(s); the normalised context goes to the staging tile ----
          if constexpr (kAligned) {
#pragma unroll
            for (int ti = 0; ti < 2; ++ti) {
              const int tile = ew + ti * kEpiWarps;
              if (tile < tiles_b)
                attend_tile<kNT>(q_base, k_base, v_base, c_base, tile * tile_rows_b, tile * tile_rows_b, okm[ti][0],
                                 okm[ti][1], tile_rows_b, lane);
            }
          } else {
            attend_tile<kNT>(q_base, k_base, v_base, c_base, 16 * ew, band0, okm[0][0], okm[0][1], 16, lane);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2, kEpiWarps * 32);  // B: context tile complete, Q / K / V tiles free
        if (ew == 0 && lane == 0) {
          tma_store_2d(single ? &tm_out1 : &tm_out, smem_ctx, head * kHeadDim, static_cast<int>(row0));
          tma_store_commit();
        }
      }
    }
    if (ew == 0 && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

template <int kNT, int kBandOff, bool kAligned>
cudaError_t launch_nt(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const CUtensorMap& tm_out,
                      const CUtensorMap& tm_out1, const QkvAttnArgs& p, cudaStream_t stream, int num_sms) {
  auto kern = qkv_attention_kernel<kNT, kBandOff, kAligned>;
  static unsigned long long smem_done = 0;
  cudaError_t e = ensure_dynamic_smem(kern, kSmemBytes, &smem_done);
  if (e != cudaSuccess) return e;
  const int pair_blocks = (p.row_blocks + kCluster - 1) / kCluster;
  const int max_clusters = num_sms / kCluster;
  const int clusters = pair_blocks < max_clusters ? pair_blocks : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * kCluster);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tm_a, tm_b, tm_out, tm_out1, p);
}

}  // namespace

int qkv_attention_rows_per_block(int seq_len) {
  if (seq_len < 1 || seq_len > 32) return 0;
  const int full = BM / seq_len;  // whole sequences that fit a 128-row block
  if (seq_len <= 16) {
    // sequence-aligned 16-row tiles: 8 tiles (one per epilogue warp, one round) hold 8 * floor(16 / T) sequences. Taken when
    // that gives up at most 7 % of the block (T = 5: 120 of 125 rows); otherwise some warps attend a second tile per unit.
    const int one_round = kEpiWarps * (16 / seq_len);
    if (one_round * 100 >= full * 93) return (one_round < full ? one_round : full) * seq_len;
  }
  return full * seq_len;
}

cudaError_t launch_qkv_attention(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const CUtensorMap& tm_out,
                                 const CUtensorMap& tm_out1, const QkvAttnArgs& p, cudaStream_t stream, int num_sms) {
  if (p.seq_len < 1 || p.seq_len > 32 || p.rows_per_block != qkv_attention_rows_per_block(p.seq_len) ||
      p.row_blocks < 1 || p.vec_c == nullptr || p.vec_s == nullptr || p.mask_src == nullptr || (p.prev_norm && !p.stats_in))
    return cudaErrorInvalidValue;
  if (p.seq_len <= 16) return launch_nt<2, 0, true>(tm_a, tm_b, tm_out, tm_out1, p, stream, num_sms);
  if (p.causal) {  // no keys after the query tile
    if (p.seq_len <= 17) return launch_nt<4, 16, false>(tm_a, tm_b, tm_out, tm_out1, p, stream, num_sms);
    return launch_nt<6, 32, false>(tm_a, tm_b, tm_out, tm_out1, p, stream, num_sms);
  }
  if (p.seq_len <= 17) return launch_nt<6, 16, false>(tm_a, tm_b, tm_out, tm_out1, p, stream, num_sms);
  return launch_nt<10, 32, false>(tm_a, tm_b, tm_out, tm_out1, p, stream, num_sms);
}

}  // namespace stlt
