// Masked multi-head self-attention for sequences of 65..256 tokens: the temporal stack when a caller samples more
// frames than the 32- / 64-token tiles of attention_mma.cu / attention_cross.cu hold. The reference's position table
// allows 256 frames (src/modelling/models.py:88-96, `max_position_embeddings`; `--layout_num_frames` in
// src/utils/parser.py:61-66 sets how many are sampled). Same semantics as the short kernels: the SDPA inside
// nn.MultiheadAttention (src/modelling/models.py:46-55,118-128) with the key-padding mask (`mask_src == 0`,
// src/modelling/datasets.py:274-286) and the causal mask (src/utils/model_utils.py:4-7) as predicates.
//
// One CTA of four warps owns one (sequence, head): the Q, K and V head slices (T x 64 bf16 each) are staged once with
// coalesced 16-byte cp.async into XOR-swizzled shared memory; each warp then takes 16-query tiles round-robin and
// sweeps the keys in blocks of 64 with an online softmax (running row maximum and sum, fp32):
//   S = Q K^T (16 x 64, mma.sync m16n8k16) -> predicates -> rescale O and the row sum -> O += P V.
// Under the causal mask the key blocks past a query tile are skipped. The bf16 context rows leave through the tile's
// own (dead) Q rows as 16-byte vectors. kSplit = fp32-parity flavour: Q, K, V arrive as bf16 hi/lo planes, every
// product is hi*hi + lo*hi + hi*lo with P split in registers, the context leaves as hi/lo planes.
#include "kernels.h"
#include "mma_tiles.cuh"

namespace stlt {

namespace {

constexpr int kLongWarps = 4;
constexpr int kLongMaxT = 256;
constexpr int kLongTileBytes = kLongMaxT * 128;  // 256 rows x 64 bf16

template <bool kSplit>
__global__ void __launch_bounds__(kLongWarps * 32, 1)
attention_long_kernel(const __nv_bfloat16* __restrict__ qkv, long long qkv_plane_elems,
                      const long long* __restrict__ mask_src, long long num_seqs, int T, int causal,
                      __nv_bfloat16* __restrict__ out, long long out_plane_elems) {
  constexpr int kTiles = kSplit ? 6 : 3;  // Q, K, V (+ their lo planes)
  constexpr uint32_t kLo = 3 * kLongTileBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t key_bits[kLongMaxT / 32];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int g = lane >> 2;
  const int t = lane & 3;
  const uint32_t q_base = smem_u32(smem_raw);
  const uint32_t k_base = q_base + kLongTileBytes;
  const uint32_t v_base = k_base + kLongTileBytes;
  const float kScale = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int m_tiles = (T + 15) / 16;
  const int rows_pad = m_tiles * 16;
  const long long num_items = num_seqs * kHeads;

  for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
    const long long seq = item / kHeads;
    const int head = static_cast<int>(item - seq * kHeads);
    const long long tok0 = seq * T;
    __syncthreads();  // the previous item's tiles and key bits are no longer read
    // ---- stage Q, K, V: 8 lanes x 16 B per row, 16 rows per pass of the CTA ----
    {
      const int chunk = tid & 7;
      for (int row = tid >> 3; row < rows_pad; row += kLongWarps * 4) {
        if (row < T) {
          const __nv_bfloat16* src = qkv + (tok0 + row) * kQkv + head * kHeadDim + chunk * 8;
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
#pragma unroll
          for (int tl = 0; tl < kTiles; ++tl)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(
                             tile_addr(q_base + tl * kLongTileBytes, row, chunk)),
                         "r"(0u)
                         : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // key validity bits (overlaps with the copies in flight): word w covers keys 32w .. 32w + 31
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int key = half * 128 + tid;
      const bool ok = key < T && mask_src[tok0 + (key < T ? key : 0)] != 0;
      const uint32_t bits = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) key_bits[half * 4 + warp] = bits;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    for (int mt = warp; mt < m_tiles; mt += kLongWarps) {
      // Q fragments of this tile stay in registers for the whole key sweep
      uint32_t a[4][4], al[kSplit ? 4 : 1][4];
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const int arow = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, achunk = kt * 2 + (lane >> 4);
        ldmatrix_x4(tile_addr(q_base, arow, achunk), a[kt]);
        if constexpr (kSplit) ldmatrix_x4(tile_addr(q_base + kLo, arow, achunk), al[kt]);
      }
      float o[8][4];
#pragma unroll
      for (int dt = 0; dt < 8; ++dt)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[dt][i] = 0.f;
      float run_max[2] = {-INFINITY, -INFINITY}, run_sum[2] = {0.f, 0.f};
      const int last_row = mt * 16 + 15;
      const int key_end = causal ? (last_row + 1 < T ? last_row + 1 : T) : T;  // keys >= key_end are masked for every row
      const int k_blocks = (key_end + 63) / 64;

      for (int kb = 0; kb < k_blocks; ++kb) {
        const int key0 = kb * 64;
        // ---- S = Q K^T for 16 queries x 64 keys ----
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            if (key0 + np * 16 < key_end) {  // warp-uniform
              uint32_t b[4], bl[4];
              const int brow = key0 + (np * 2 + (lane >> 4)) * 8 + (lane & 7), bchunk = kt * 2 + ((lane >> 3) & 1);
              ldmatrix_x4(tile_addr(k_base, brow, bchunk), b);
              mma_bf16(s[np * 2 + 0], a[kt], b[0], b[1]);
              mma_bf16(s[np * 2 + 1], a[kt], b[2], b[3]);
              if constexpr (kSplit) {
                ldmatrix_x4(tile_addr(k_base + kLo, brow, bchunk), bl);
                mma_bf16(s[np * 2 + 0], al[kt], b[0], b[1]);
                mma_bf16(s[np * 2 + 1], al[kt], b[2], b[3]);
                mma_bf16(s[np * 2 + 0], a[kt], bl[0], bl[1]);
                mma_bf16(s[np * 2 + 1], a[kt], bl[2], bl[3]);
              }
            }
          }
        }
        // ---- predicates + online softmax (fp32) ----
        const uint32_t bits_lo = key_bits[kb * 2], bits_hi = key_bits[kb * 2 + 1];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = mt * 16 + g + 8 * h;
          float m = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int kl = nt * 8 + 2 * t + e;  // key within the block
              const uint32_t word = kl < 32 ? bits_lo : bits_hi;
              const bool ok = ((word >> (kl & 31)) & 1u) && (!causal || key0 + kl <= row);
              const float v = ok ? s[nt][2 * h + e] * kScale : -INFINITY;
              s[nt][2 * h + e] = v;
              m = fmaxf(m, v);
            }
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          const float new_max = fmaxf(run_max[h], m);
          const float mm = (new_max == -INFINITY) ? 0.f : new_max;  // nothing but masked keys so far
          const float alpha = exp2f(run_max[h] - mm);               // 0 while run_max is -inf
          run_max[h] = new_max;
          float sum = 0.f;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float pv = exp2f(s[nt][2 * h + e] - mm);
              s[nt][2 * h + e] = pv;
              sum += pv;
            }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          run_sum[h] = run_sum[h] * alpha + sum;
#pragma unroll
          for (int dt = 0; dt < 8; ++dt) {
            o[dt][2 * h] *= alpha;
            o[dt][2 * h + 1] *= alpha;
          }
        }
        // ---- O += P V ----
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (key0 + j * 16 < key_end) {  // warp-uniform
            uint32_t p[4], pl[4];
            p[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
            p[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
            p[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
            p[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
            if constexpr (kSplit) {
              pl[0] = pack_bf16x2(bf16_residual(s[2 * j][0]), bf16_residual(s[2 * j][1]));
              pl[1] = pack_bf16x2(bf16_residual(s[2 * j][2]), bf16_residual(s[2 * j][3]));
              pl[2] = pack_bf16x2(bf16_residual(s[2 * j + 1][0]), bf16_residual(s[2 * j + 1][1]));
              pl[3] = pack_bf16x2(bf16_residual(s[2 * j + 1][2]), bf16_residual(s[2 * j + 1][3]));
            }
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
              uint32_t b[4], bl[4];
              const int vrow = key0 + j * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), vchunk = dp * 2 + (lane >> 4);
              ldmatrix_x4_trans(tile_addr(v_base, vrow, vchunk), b);
              mma_bf16(o[dp * 2 + 0], p, b[0], b[1]);
              mma_bf16(o[dp * 2 + 1], p, b[2], b[3]);
              if constexpr (kSplit) {
                ldmatrix_x4_trans(tile_addr(v_base + kLo, vrow, vchunk), bl);
                mma_bf16(o[dp * 2 + 0], pl, b[0], b[1]);
                mma_bf16(o[dp * 2 + 1], pl, b[2], b[3]);
                mma_bf16(o[dp * 2 + 0], p, bl[0], bl[1]);
                mma_bf16(o[dp * 2 + 1], p, bl[2], bl[3]);
              }
            }
          }
        }
      }

      // ---- normalise, stage through this tile's (dead) Q rows, write 16-byte vectors ----
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = mt * 16 + g + 8 * h;
        const float is = run_sum[h] > 0.f ? 1.0f / run_sum[h] : 0.f;  // fully masked row -> zeros
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          const float x0 = o[dt][2 * h] * is, x1 = o[dt][2 * h + 1] * is;
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(q_base, row, dt) + 4 * t), "r"(pack_bf16x2(x0, x1)) : "memory");
          if (kSplit) {
            const uint32_t vl = pack_bf16x2(bf16_residual(x0), bf16_residual(x1));
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(q_base + kLo, row, dt) + 4 * t), "r"(vl) : "memory");
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int row = mt * 16 + it * 4 + (lane >> 3);
        const int chunk = lane & 7;
        if (row < T) {
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(tile_addr(q_base, row, chunk))
                       : "memory");
          __nv_bfloat16* dst = out + (tok0 + row) * kHidden + head * kHeadDim + chunk * 8;
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
  }
}

template <bool kSplit>
cudaError_t launch_long(const __nv_bfloat16* qkv, long long qkv_plane_rows, const long long* mask_src,
                        long long num_seqs, int T, bool causal, __nv_bfloat16* out, long long out_plane_rows,
                        cudaStream_t stream) {
  const int smem = (kSplit ? 6 : 3) * kLongTileBytes;
  static unsigned long long smem_done = 0;  // per instantiation, one bit per device
  {
    cudaError_t e = ensure_dynamic_smem(attention_long_kernel<kSplit>, smem, &smem_done);
    if (e != cudaSuccess) return e;
  }
  long long blocks = num_seqs * kHeads;
  const long long cap = 148LL * (kSplit ? 1 : 2) * 8;  // grid-stride inside
  if (blocks > cap) blocks = cap;
  attention_long_kernel<kSplit><<<static_cast<unsigned>(blocks), kLongWarps * 32, smem, stream>>>(
      qkv, qkv_plane_rows * kQkv, mask_src, num_seqs, T, causal ? 1 : 0, out, out_plane_rows * kHidden);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_attention_long(const __nv_bfloat16* qkv, int planes, long long qkv_plane_rows,
                                  const long long* mask_src, long long num_seqs, int T, bool causal,
                                  __nv_bfloat16* out, long long out_plane_rows, cudaStream_t stream) {
  if (T < 1 || T > kLongMaxT || mask_src == nullptr) return cudaErrorInvalidValue;
  if (num_seqs == 0) return cudaSuccess;
  if (planes == 2)
    return launch_long<true>(qkv, qkv_plane_rows, mask_src, num_seqs, T, causal, out, out_plane_rows, stream);
  return launch_long<false>(qkv, 0, mask_src, num_seqs, T, causal, out, 0, stream);
}

}  // namespace stlt
