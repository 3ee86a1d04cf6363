// Multi-head attention between two token streams with up to 64 queries and 64 keys per sequence, on
// mma.sync tiles. Used by the CACNF fusion path for the sequences that do
// not fit attention_mma.cu's 32-token tile: the 33-token appearance stream (self-attention inside
// TransformerResnet, reference src/modelling/models.py:236-276, and inside the fusion layers,
// :328-431) and the shared cross-attention of CrossModalModule in both directions (17 x 33 and
// 33 x 17, :395-405).
//
// One warp owns one (sequence, head): the Q rows come from `q` ([*, ldq] bf16, column q_off + 64 h),
// K / V rows from `kv` ([*, ldkv] bf16, columns k_off / v_off + 64 h); all three tiles (64 x 64 bf16)
// live in warp-private swizzled shared memory. Queries are processed 16 at a time:
// S = Q K^T (16 x 64, fp32 fragments) -> masks (key padding from mask_src == 0, optional causal) ->
// softmax with quad shuffles -> O = P V with P re-used from registers. Nothing but the bf16 context
// leaves the SM.
#include "kernels.h"
#include "mma_tiles.cuh"

namespace stlt {

namespace {

constexpr int kWarps = 4;

// kRows = rows of the Q / K / V tiles (48 or 64): 48-row tiles hold the 33-token appearance stream in
// 18 KB per warp, so three 4-warp blocks fit on an SM instead of two.
// kSplit = fp32-parity flavour (as in attention_mma.cu): Q, K, V arrive as bf16 hi/lo planes (`*_plane` elements
// apart), every product is hi*hi + lo*hi + hi*lo with P split in registers, the context leaves as hi/lo planes.
template <int kRows, bool kSplit>
__global__ void __launch_bounds__(kWarps * 32, kSplit ? 1 : (kRows <= 48 ? 3 : 2))
attention_cross_kernel(const __nv_bfloat16* __restrict__ q, int ldq, int q_off,
                       const __nv_bfloat16* __restrict__ kv, int ldkv, int k_off, int v_off,
                       const long long* __restrict__ mask_src, long long num_seqs, int Tq, int Tk,
                       int causal, __nv_bfloat16* __restrict__ out, long long q_plane, long long kv_plane,
                       long long out_plane) {
  constexpr int kTileBytes = kRows * 128;  // kRows x 64 bf16
  constexpr int kKeyTiles = kRows / 8;     // 8-key accumulator tiles
  constexpr int kTiles = kSplit ? 6 : 3;   // Q, K, V (+ their lo planes)
  constexpr uint32_t kLo = 3 * kTileBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2;
  const int t = lane & 3;
  const uint32_t q_base = smem_u32(smem_raw) + warp * kTiles * kTileBytes;
  const uint32_t k_base = q_base + kTileBytes;
  const uint32_t v_base = k_base + kTileBytes;
  const long long num_items = num_seqs * kHeads;
  const long long gwarp = blockIdx.x * static_cast<long long>(kWarps) + warp;
  const long long nwarps = gridDim.x * static_cast<long long>(kWarps);
  const float kScale = 0.125f * 1.4426950408889634f;
  const int m_tiles = (Tq + 15) / 16;
  const int k_steps = (Tk + 15) / 16;  // 16-key steps that hold at least one real key

  for (long long item = gwarp; item < num_items; item += nwarps) {
    const long long seq = item / kHeads;
    const int head = static_cast<int>(item - seq * kHeads);
    {
      const int chunk = lane & 7;
      const int r0 = lane >> 3;
#pragma unroll 4
      for (int it = 0; it < kRows / 4; ++it) {
        const int row = it * 4 + r0;
        if (row < Tq) {
          const __nv_bfloat16* src = q + (seq * Tq + row) * ldq + q_off + head * kHeadDim + chunk * 8;
          cp_async16(tile_addr(q_base, row, chunk), src);
          if (kSplit) cp_async16(tile_addr(q_base + kLo, row, chunk), src + q_plane);
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile_addr(q_base, row, chunk)), "r"(0u) : "memory");
          if (kSplit)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile_addr(q_base + kLo, row, chunk)), "r"(0u) : "memory");
        }
        if (row < Tk) {
          const __nv_bfloat16* src = kv + (seq * Tk + row) * ldkv + head * kHeadDim + chunk * 8;
          cp_async16(tile_addr(k_base, row, chunk), src + k_off);
          cp_async16(tile_addr(v_base, row, chunk), src + v_off);
          if (kSplit) {
            cp_async16(tile_addr(k_base + kLo, row, chunk), src + kv_plane + k_off);
            cp_async16(tile_addr(v_base + kLo, row, chunk), src + kv_plane + v_off);
          }
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile_addr(k_base, row, chunk)), "r"(0u) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile_addr(v_base, row, chunk)), "r"(0u) : "memory");
          if (kSplit) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile_addr(k_base + kLo, row, chunk)), "r"(0u) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile_addr(v_base + kLo, row, chunk)), "r"(0u) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // key validity bits: key j = lane (low word) and lane + 32 (high word)
    bool ok_lo = lane < Tk, ok_hi = lane + 32 < Tk;
    if (mask_src != nullptr) {
      if (ok_lo) ok_lo = mask_src[seq * Tk + lane] != 0;
      if (ok_hi) ok_hi = mask_src[seq * Tk + lane + 32] != 0;
    }
    const uint32_t bits_lo = __ballot_sync(0xffffffffu, ok_lo);
    const uint32_t bits_hi = __ballot_sync(0xffffffffu, ok_hi);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    for (int mt = 0; mt < m_tiles; ++mt) {
      // ---- S = Q K^T for 16 queries x 64 keys ----
      float s[kKeyTiles][4];
#pragma unroll
      for (int nt = 0; nt < kKeyTiles; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {  // 64 features
        uint32_t a[4], al[4];
        const int arow = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, achunk = kt * 2 + (lane >> 4);
        ldmatrix_x4(tile_addr(q_base, arow, achunk), a);
        if (kSplit) ldmatrix_x4(tile_addr(q_base + kLo, arow, achunk), al);
#pragma unroll
        for (int np = 0; np < kKeyTiles / 2; ++np) {
          if (np * 16 < Tk) {  // warp-uniform
            uint32_t b[4], bl[4];
            const int brow = (np * 2 + (lane >> 4)) * 8 + (lane & 7), bchunk = kt * 2 + ((lane >> 3) & 1);
            ldmatrix_x4(tile_addr(k_base, brow, bchunk), b);
            mma_bf16(s[np * 2 + 0], a, b[0], b[1]);
            mma_bf16(s[np * 2 + 1], a, b[2], b[3]);
            if (kSplit) {
              ldmatrix_x4(tile_addr(k_base + kLo, brow, bchunk), bl);
              mma_bf16(s[np * 2 + 0], al, b[0], b[1]);
              mma_bf16(s[np * 2 + 1], al, b[2], b[3]);
              mma_bf16(s[np * 2 + 0], a, bl[0], bl[1]);
              mma_bf16(s[np * 2 + 1], a, bl[2], bl[3]);
            }
          }
        }
      }
      // ---- masked softmax (fp32) ----
      uint32_t p[kKeyTiles / 2][4];
      uint32_t pl[kSplit ? kKeyTiles / 2 : 1][4];  // lo plane of P
      float inv_sum[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = mt * 16 + g + 8 * h;
        float m = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < kKeyTiles; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int key = nt * 8 + 2 * t + e;
            const uint32_t word = key < 32 ? bits_lo : bits_hi;
            const bool ok = ((word >> (key & 31)) & 1u) && (!causal || key <= row);
            const float v = ok ? s[nt][2 * h + e] * kScale : -INFINITY;
            s[nt][2 * h + e] = v;
            m = fmaxf(m, v);
          }
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        const float mm = (m == -INFINITY) ? 0.f : m;
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < kKeyTiles; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = exp2f(s[nt][2 * h + e] - mm);
            s[nt][2 * h + e] = pv;
            sum += pv;
          }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        inv_sum[h] = sum > 0.f ? 1.0f / sum : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kKeyTiles / 2; ++j) {
        p[j][0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
        p[j][1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
        p[j][2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
        p[j][3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
        if constexpr (kSplit) {
          pl[j][0] = pack_bf16x2(bf16_residual(s[2 * j][0]), bf16_residual(s[2 * j][1]));
          pl[j][1] = pack_bf16x2(bf16_residual(s[2 * j][2]), bf16_residual(s[2 * j][3]));
          pl[j][2] = pack_bf16x2(bf16_residual(s[2 * j + 1][0]), bf16_residual(s[2 * j + 1][1]));
          pl[j][3] = pack_bf16x2(bf16_residual(s[2 * j + 1][2]), bf16_residual(s[2 * j + 1][3]));
        }
      }
      // ---- O = P V ----
      float o[8][4];
#pragma unroll
      for (int dt = 0; dt < 8; ++dt)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[dt][i] = 0.f;
#pragma unroll
      for (int j = 0; j < kKeyTiles / 2; ++j) {
        if (j < k_steps) {  // warp-uniform
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t b[4], bl[4];
            const int vrow = j * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), vchunk = dp * 2 + (lane >> 4);
            ldmatrix_x4_trans(tile_addr(v_base, vrow, vchunk), b);
            mma_bf16(o[dp * 2 + 0], p[j], b[0], b[1]);
            mma_bf16(o[dp * 2 + 1], p[j], b[2], b[3]);
            if constexpr (kSplit) {
              ldmatrix_x4_trans(tile_addr(v_base + kLo, vrow, vchunk), bl);
              mma_bf16(o[dp * 2 + 0], pl[j], b[0], b[1]);
              mma_bf16(o[dp * 2 + 1], pl[j], b[2], b[3]);
              mma_bf16(o[dp * 2 + 0], p[j], bl[0], bl[1]);
              mma_bf16(o[dp * 2 + 1], p[j], bl[2], bl[3]);
            }
          }
        }
      }
      // ---- normalise, stage through this m-tile's (dead) Q rows, write 16-byte vectors ----
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = mt * 16 + g + 8 * h;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          const float x0 = o[dt][2 * h] * inv_sum[h], x1 = o[dt][2 * h + 1] * inv_sum[h];
          const uint32_t v = pack_bf16x2(x0, x1);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(q_base, row, dt) + 4 * t), "r"(v) : "memory");
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
        if (row < Tq) {
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(tile_addr(q_base, row, chunk))
                       : "memory");
          __nv_bfloat16* dst = out + (seq * Tq + row) * kHidden + head * kHeadDim + chunk * 8;
          *reinterpret_cast<uint4*>(dst) = v;
          if (kSplit) {
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(tile_addr(q_base + kLo, row, chunk))
                         : "memory");
            *reinterpret_cast<uint4*>(dst + out_plane) = v;
          }
        }
      }
    }
    __syncwarp();
  }
}

template <int kRows, bool kSplit>
static cudaError_t launch_rows(const __nv_bfloat16* q, int ldq, int q_off, const __nv_bfloat16* kv, int ldkv,
                               int k_off, int v_off, const long long* mask_src, long long num_seqs, int Tq,
                               int Tk, bool causal, __nv_bfloat16* out, long long q_plane, long long kv_plane,
                               long long out_plane, cudaStream_t stream) {
  const int smem = kWarps * (kSplit ? 6 : 3) * kRows * 128;
  static unsigned long long smem_done = 0;  // per instantiation, one bit per device
  {
    cudaError_t e = ensure_dynamic_smem(attention_cross_kernel<kRows, kSplit>, smem, &smem_done);
    if (e != cudaSuccess) return e;
  }
  const long long items = num_seqs * kHeads;
  long long blocks = (items + kWarps - 1) / kWarps;
  const long long cap = 148LL * 3 * 8;
  if (blocks > cap) blocks = cap;
  attention_cross_kernel<kRows, kSplit><<<static_cast<unsigned>(blocks), kWarps * 32, smem, stream>>>(
      q, ldq, q_off, kv, ldkv, k_off, v_off, mask_src, num_seqs, Tq, Tk, causal ? 1 : 0, out, q_plane, kv_plane,
      out_plane);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_attention_cross(const __nv_bfloat16* q, int ldq, int q_off, const __nv_bfloat16* kv,
                                   int ldkv, int k_off, int v_off, const long long* mask_src,
                                   long long num_seqs, int Tq, int Tk, bool causal, __nv_bfloat16* out,
                                   cudaStream_t stream, int planes, long long q_plane_rows,
                                   long long kv_plane_rows, long long out_plane_rows) {
  if (Tq < 1 || Tq > 64 || Tk < 1 || Tk > 64) return cudaErrorInvalidValue;
  if (causal && Tq != Tk) return cudaErrorInvalidValue;
  if (num_seqs == 0) return cudaSuccess;
  const bool small = Tq <= 48 && Tk <= 48;
  if (planes == 2) {  // fp32-parity mode: hi / lo planes in and out
    const long long qp = q_plane_rows * ldq, kp = kv_plane_rows * ldkv, op = out_plane_rows * kHidden;
    if (small)
      return launch_rows<48, true>(q, ldq, q_off, kv, ldkv, k_off, v_off, mask_src, num_seqs, Tq, Tk, causal, out, qp, kp, op, stream);
    return launch_rows<64, true>(q, ldq, q_off, kv, ldkv, k_off, v_off, mask_src, num_seqs, Tq, Tk, causal, out, qp, kp, op, stream);
  }
  if (small)
    return launch_rows<48, false>(q, ldq, q_off, kv, ldkv, k_off, v_off, mask_src, num_seqs, Tq, Tk, causal, out, 0, 0, 0, stream);
  return launch_rows<64, false>(q, ldq, q_off, kv, ldkv, k_off, v_off, mask_src, num_seqs, Tq, Tk, causal, out, 0, 0, 0, stream);
}

}  // namespace stlt
