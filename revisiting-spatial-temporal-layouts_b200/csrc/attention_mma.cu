// K3: masked multi-head self-attention over the short STLT sequences, evaluated on
// warp-level tensor-core tiles. Same semantics as attention.cu (reference: the SDPA inside
// nn.MultiheadAttention configured at src/modelling/models.py:46-55,118-128, key-padding masks from
// src/modelling/datasets.py:274-286, causal mask src/utils/model_utils.py:4-7).
//
// One warp owns one (tile of R = floor(32/T)*T consecutive tokens = floor(32/T) whole sequences,
// head). The tile's Q, K, V head slices (R x 64 bf16 each) are fetched with coalesced 16-byte
// cp.async into warp-private, XOR-swizzled shared memory and never leave the SM again:
//   S = Q K^T      32x32 scores, mma.sync m16n8k16 (bf16 in, fp32 accumulate)
//   P = softmax    block-diagonal (same sequence) + key-padding + causal predicates applied on the
//                  accumulator fragments; row max / sum are quad shuffles; fp32 throughout
//   O = P V        P re-used from registers as the A operand (bf16), V via ldmatrix.trans
// The normalised context rows are staged through the (dead) Q tile and written as 16-byte vectors.
//
// kSplit = true is the fp32-parity flavour: Q, K, V arrive as bf16 hi/lo planes (the split output
// of the in-projection GEMM) and every product is evaluated as hi*hi + lo*hi + hi*lo on the tensor
// cores (P is split in registers), which keeps ~2^-16 relative accuracy; the context is written as
// hi/lo planes again.
//
// The scores of sequences sharing a tile are computed and then masked away; at T = 5 that is
// 30x30 computed for 6x(5x5) used, which is still ~4x fewer issued instructions per token than a
// CUDA-core formulation, and the kernel is HBM-bound (reads 4.6 KB, writes 1.5 KB per token).
#include "kernels.h"
#include "mma_tiles.cuh"

namespace stlt {

namespace {

constexpr int kTileBytes = 32 * 128;  // 32 rows x 64 bf16

template <bool kSplit>
__global__ void __launch_bounds__(kSplit ? 128 : 256, 2)
attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv, long long qkv_plane_elems,
                     const long long* __restrict__ mask_src, long long num_seqs, int T, int G,
                     int causal, __nv_bfloat16* __restrict__ out, long long out_plane_elems,
                     long long num_items, DropCfg drop, const int* __restrict__ dyn, int dyn_region) {
  constexpr int kWarps = kSplit ? 4 : 8;
  constexpr int kTiles = kSplit ? 6 : 3;  // Q, K, V (+ their lo planes)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2;  // fragment row within an 8-row group
  const int t = lane & 3;   // fragment column pair
  const int R = G * T;
  const uint32_t q_base = smem_u32(smem_raw) + warp * kTiles * kTileBytes;
  const uint32_t k_base = q_base + kTileBytes;
  const uint32_t v_base = k_base + kTileBytes;
  constexpr uint32_t kLo = 3 * kTileBytes;  // lo-plane tiles follow the three hi tiles

  // Static part of the mask for this thread's fragment positions: query rows mt*16 + g + 8h,
  // key columns nt*8 + 2t + e. Bit (nt*2 + e) of allow[mt][h] = same sequence (and causal order).
  uint32_t allow[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = mt * 16 + g + 8 * h;
      const int rseq = row / T, rpos = row - rseq * T;
      uint32_t bits = 0;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = nt * 8 + 2 * t + e;
          const int kseq = key / T, kpos = key - kseq * T;
          const bool ok = row < R && key < R && kseq == rseq && (!causal || kpos <= rpos);
          bits |= (ok ? 1u : 0u) << (nt * 2 + e);
        }
      allow[mt][h] = bits;
    }

  // pad-skipping layout (compact.cu): the number of sequences and the first row of the region come from the device header
  long long row_base = 0;
  if (dyn != nullptr) {
    num_seqs = __ldg(dyn + (dyn_region == 0 ? kDynFull : kDynSingle));
    row_base = dyn_region == 0 ? 0 : __ldg(dyn + kDynSingleRow0);
    num_items = ((num_seqs + G - 1) / G) * kHeads;
  }
  const long long total_tokens = num_seqs * T;
  const long long gwarp = blockIdx.x * static_cast<long long>(kWarps) + warp;
  const long long nwarps = gridDim.x * static_cast<long long>(kWarps);
  const float kScale = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)

  for (long long item = gwarp; item < num_items; item += nwarps) {
    const long long grp = item / kHeads;
    const int head = static_cast<int>(item - grp * kHeads);
    const long long remaining = total_tokens - grp * R;
    const long long base = row_base + grp * R;
    const int nrows = remaining < R ? static_cast<int>(remaining) : R;

    // ---- stage Q, K, V (coalesced: 8 lanes x 16 B per row, 4 rows per instruction) ----
    {
      const int chunk = lane & 7;
      const int r0 = lane >> 3;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + r0;
        const __nv_bfloat16* src = qkv + (base + row) * kQkv + head * kHeadDim + chunk * 8;
        if (row < nrows) {
          cp_async16(tile_addr(q_base, row, chunk), src);
          cp_async16(tile_addr(k_base, row, chunk), src + kHidden);
          cp_async16(tile_addr(v_base, row, chunk), src + 2 * kHidden);
          if (kSplit) {
            const __nv_bfloat16* lo = src + qkv_plane_elems;
            cp_async16(tile_addr(q_base + kLo, row, chunk), lo);
            cp_async16(tile_addr(k_base + kLo, row, chunk), lo + kHidden);
            cp_async16(tile_addr(v_base + kLo, row, chunk), lo + 2 * kHidden);
          }
        } else {
          const uint32_t z = 0;
#pragma unroll
          for (int tl = 0; tl < kTiles; ++tl)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(
                             tile_addr(q_base + tl * kTileBytes, row, chunk)),
                         "r"(z)
                         : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // key-padding bits of the tile (lane = token); overlaps with the copies in flight
    const bool key_ok = lane < nrows && mask_src[base + (lane < nrows ? lane : 0)] != 0;
    const uint32_t keybits = __ballot_sync(0xffffffffu, key_ok);
    uint32_t kb = 0;  // this thread's 8 key columns
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) kb |= ((keybits >> (nt * 8 + 2 * t + e)) & 1u) << (nt * 2 + e);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    // ---- S = Q K^T ----
    float s[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[mt][nt][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      uint32_t a[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk = kt * 2 + (lane >> 4);
        ldmatrix_x4(tile_addr(q_base, row, chunk), a[mt]);
        if (kSplit) ldmatrix_x4(tile_addr(q_base + kLo, row, chunk), al[mt]);
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {  // pairs of key tiles
        uint32_t b[4], bl[4];
        const int row = (np * 2 + (lane >> 4)) * 8 + (lane & 7), chunk = kt * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(tile_addr(k_base, row, chunk), b);
        if (kSplit) ldmatrix_x4(tile_addr(k_base + kLo, row, chunk), bl);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16(s[mt][np * 2 + 0], a[mt], b[0], b[1]);
          mma_bf16(s[mt][np * 2 + 1], a[mt], b[2], b[3]);
          if (kSplit) {
            mma_bf16(s[mt][np * 2 + 0], al[mt], b[0], b[1]);
            mma_bf16(s[mt][np * 2 + 1], al[mt], b[2], b[3]);
            mma_bf16(s[mt][np * 2 + 0], a[mt], bl[0], bl[1]);
            mma_bf16(s[mt][np * 2 + 1], a[mt], bl[2], bl[3]);
          }
        }
      }
    }

    // ---- masked softmax on the accumulator fragments (fp32) ----
    uint32_t p[2][2][4];   // P as bf16 A fragments: [m tile][key k-step][4 regs]
    uint32_t pl[2][2][4];  // lo plane of P (kSplit)
    float inv_sum[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t ok = allow[mt][h] & kb;
        float m = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float v = s[mt][nt][2 * h + e] * kScale;
            v = ((ok >> (nt * 2 + e)) & 1u) ? v : -INFINITY;
            s[mt][nt][2 * h + e] = v;
            m = fmaxf(m, v);
          }
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        const float mm = (m == -INFINITY) ? 0.f : m;  // fully masked row -> all-zero probabilities
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = exp2f(s[mt][nt][2 * h + e] - mm);
            s[mt][nt][2 * h + e] = pv;
            sum += pv;
          }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        inv_sum[mt][h] = sum > 0.f ? 1.0f / sum : 0.f;
        if (!kSplit && drop.thr16 != 0) {
          // dropout on the probabilities (nn.MultiheadAttention dropout, training only); the softmax
          // denominator above is taken before the mask, as F.dropout(softmax(...)) does
          const int row = mt * 16 + g + 8 * h;
          const int seq0 = (row / T) * T;
          const unsigned long long ebase =
              (static_cast<unsigned long long>(base + row) * kHeads + head) * 32ull;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int kpos = nt * 8 + 2 * t + e - seq0;
              if ((ok >> (nt * 2 + e)) & 1u) {
                const unsigned long long el = ebase + static_cast<unsigned>(kpos);
                s[mt][nt][2 * h + e] *= drop_mul(drop_bits(drop.key, el >> 1), static_cast<int>(el & 1), drop);
              }
            }
        }
      }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        p[mt][j][0] = pack_bf16x2(s[mt][2 * j][0], s[mt][2 * j][1]);
        p[mt][j][1] = pack_bf16x2(s[mt][2 * j][2], s[mt][2 * j][3]);
        p[mt][j][2] = pack_bf16x2(s[mt][2 * j + 1][0], s[mt][2 * j + 1][1]);
        p[mt][j][3] = pack_bf16x2(s[mt][2 * j + 1][2], s[mt][2 * j + 1][3]);
        if (kSplit) {
          pl[mt][j][0] = pack_bf16x2(bf16_residual(s[mt][2 * j][0]), bf16_residual(s[mt][2 * j][1]));
          pl[mt][j][1] = pack_bf16x2(bf16_residual(s[mt][2 * j][2]), bf16_residual(s[mt][2 * j][3]));
          pl[mt][j][2] = pack_bf16x2(bf16_residual(s[mt][2 * j + 1][0]), bf16_residual(s[mt][2 * j + 1][1]));
          pl[mt][j][3] = pack_bf16x2(bf16_residual(s[mt][2 * j + 1][2]), bf16_residual(s[mt][2 * j + 1][3]));
        }
      }

    // ---- O = P V ----
    float o[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int dt = 0; dt < 8; ++dt)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[mt][dt][i] = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-wide feature tiles
        uint32_t b[4], bl[4];
        const int row = j * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), chunk = dp * 2 + (lane >> 4);
        ldmatrix_x4_trans(tile_addr(v_base, row, chunk), b);
        if (kSplit) ldmatrix_x4_trans(tile_addr(v_base + kLo, row, chunk), bl);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16(o[mt][dp * 2 + 0], p[mt][j], b[0], b[1]);
          mma_bf16(o[mt][dp * 2 + 1], p[mt][j], b[2], b[3]);
          if (kSplit) {
            mma_bf16(o[mt][dp * 2 + 0], pl[mt][j], b[0], b[1]);
            mma_bf16(o[mt][dp * 2 + 1], pl[mt][j], b[2], b[3]);
            mma_bf16(o[mt][dp * 2 + 0], p[mt][j], bl[0], bl[1]);
            mma_bf16(o[mt][dp * 2 + 1], p[mt][j], bl[2], bl[3]);
          }
        }
      }

    // ---- normalise, stage through the Q tile, write 16-byte vectors ----
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = mt * 16 + g + 8 * h;
        const float is = inv_sum[mt][h];
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          const float x0 = o[mt][dt][2 * h] * is, x1 = o[mt][dt][2 * h + 1] * is;
          const uint32_t v = pack_bf16x2(x0, x1);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(q_base, row, dt) + 4 * t), "r"(v) : "memory");
          if (kSplit) {
            const uint32_t vl = pack_bf16x2(bf16_residual(x0), bf16_residual(x1));
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(q_base + kLo, row, dt) + 4 * t), "r"(vl) : "memory");
          }
        }
      }
    __syncwarp();
    {
      const int chunk = lane & 7;
      const int r0 = lane >> 3;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + r0;
        if (row < nrows) {
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(tile_addr(q_base, row, chunk))
                       : "memory");
          __nv_bfloat16* dst = out + (base + row) * kHidden + head * kHeadDim + chunk * 8;
          *reinterpret_cast<uint4*>(dst) = v;
          if (kSplit) {
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(tile_addr(q_base + kLo, row, chunk))
                         : "memory");
            *reinterpret_cast<uint4*>(dst + out_plane_elems) = v;
          }
        }
      }
    }
    __syncwarp();
  }
}

template <bool kSplit>
static cudaError_t launch_mma(const __nv_bfloat16* qkv, long long qkv_plane_rows,
                              const long long* mask_src, long long num_seqs, int T, bool causal,
                              __nv_bfloat16* out, long long out_plane_rows, cudaStream_t stream,
                              DropCfg drop, const int* dyn, int dyn_region) {
  constexpr int kWarps = kSplit ? 4 : 8;
  constexpr int kTiles = kSplit ? 6 : 3;
  const int G = 32 / T;
  const long long groups = (num_seqs + G - 1) / G;
  const long long items = groups * kHeads;
  const int smem = kWarps * kTiles * kTileBytes;
  static unsigned long long smem_done = 0;  // per instantiation, one bit per device
  {
    cudaError_t e = ensure_dynamic_smem(attention_mma_kernel<kSplit>, smem, &smem_done);
    if (e != cudaSuccess) return e;
  }
  long long blocks = (items + kWarps - 1) / kWarps;
  const long long cap = 148LL * 2 * 8;  // 2 CTAs resident per SM, several waves; grid-stride inside
  if (blocks > cap) blocks = cap;
  attention_mma_kernel<kSplit><<<static_cast<unsigned>(blocks), kWarps * 32, smem, stream>>>(
      qkv, qkv_plane_rows * kQkv, mask_src, num_seqs, T, G, causal ? 1 : 0, out,
      out_plane_rows * kHidden, items, drop, dyn, dyn_region);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_attention_mma(const __nv_bfloat16* qkv, int planes, long long qkv_plane_rows,
                                 const long long* mask_src, long long num_seqs, int T, bool causal,
                                 __nv_bfloat16* out, long long out_plane_rows, cudaStream_t stream,
                                 DropCfg drop, const int* dyn, int dyn_region) {
  if (T < 1 || T > 32) return cudaErrorInvalidValue;
  if (num_seqs == 0) return cudaSuccess;
  if (planes == 2) {
    if (drop.thr16 != 0) return cudaErrorInvalidValue;  // training runs the single-plane flavour
    return launch_mma<true>(qkv, qkv_plane_rows, mask_src, num_seqs, T, causal, out, out_plane_rows, stream, drop, dyn,
                            dyn_region);
  }
  return launch_mma<false>(qkv, 0, mask_src, num_seqs, T, causal, out, 0, stream, drop, dyn, dyn_region);
}

}  // namespace stlt
