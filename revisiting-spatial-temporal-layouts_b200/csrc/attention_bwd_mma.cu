// Backward of K3 on warp-level tensor-core tiles (training step). Same
// semantics as attention_bwd.cu (the CUDA-core version, kept as a cross-check): reference = the
// SDPA inside nn.MultiheadAttention configured at src/modelling/models.py:46-55,118-128.
//
// One warp owns one (tile of R = floor(32/T)*T tokens = floor(32/T) whole sequences, head). Q, K, V
// (from the saved bf16 QKV) and dO (gradient of the context) are staged with 16-byte cp.async into
// warp-private, XOR-swizzled 32 x 64 bf16 tiles. All five products run on mma.sync m16n8k16:
//   S  = Q K^T, dP = dO V^T            accumulators in the row = query layout
//   P  = softmax(S) (masks as predicates, fp32), P~ = dropout(P), dS = P * (dP~ - rowsum(P dP~)) / 8
//   dQ = dS K                          dS re-used from registers as the A operand
//   dV = P~^T dO, dK = dS^T Q          P~ and dS go through two 32 x 32 bf16 tiles in shared memory and
//                                      come back transposed with ldmatrix.trans
// The three gradient tiles are staged through dead input tiles and written as 16-byte vectors.
#include "kernels.h"
#include "mma_tiles.cuh"

namespace stlt {

namespace {

constexpr int kTileBytes = 32 * 128;   // 32 rows x 64 bf16
constexpr int kSqTileBytes = 32 * 64;  // 32 rows x 32 bf16
constexpr int kWarps = 4;
constexpr int kWarpBytes = 4 * kTileBytes + 2 * kSqTileBytes;  // Q K V dO + P~ dS

// 32 x 32 bf16 tile, 64-byte rows, 16-byte chunks swizzled so that 8 consecutive rows of one chunk
// column hit 8 different 16-byte bank groups
__device__ __forceinline__ uint32_t sq_addr(uint32_t base, int row, int chunk) {
  return base + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}

// acc[32 x 32] = X[32 x 64] * Y[32 x 64]^T for two row-major 64-wide tiles (scores-shaped product)
__device__ __forceinline__ void product_nt(uint32_t x_base, uint32_t y_base, int lane, float (&acc)[2][4][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t a[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk = kt * 2 + (lane >> 4);
      ldmatrix_x4(tile_addr(x_base, row, chunk), a[mt]);
    }
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b[4];
      const int row = (np * 2 + (lane >> 4)) * 8 + (lane & 7), chunk = kt * 2 + ((lane >> 3) & 1);
      ldmatrix_x4(tile_addr(y_base, row, chunk), b);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma_bf16(acc[mt][np * 2 + 0], a[mt], b[0], b[1]);
        mma_bf16(acc[mt][np * 2 + 1], a[mt], b[2], b[3]);
      }
    }
  }
}

// out[32 x 64] = A[32 x 32] * Y[32 x 64] with the A fragments given per (m tile, k step)
__device__ __forceinline__ void product_nn(const uint32_t (&a)[2][2][4], uint32_t y_base, int lane,
                                           float (&o)[2][8][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[mt][dt][i] = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      const int row = j * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), chunk = dp * 2 + (lane >> 4);
      ldmatrix_x4_trans(tile_addr(y_base, row, chunk), b);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma_bf16(o[mt][dp * 2 + 0], a[mt][j], b[0], b[1]);
        mma_bf16(o[mt][dp * 2 + 1], a[mt][j], b[2], b[3]);
      }
    }
}

// A fragments of M^T for a 32 x 32 bf16 tile M stored row-major in shared memory (sq_addr layout):
// a[mt][ks] covers rows mt*16.. of M^T (= columns of M) and k = ks*16.. (= rows of M).
__device__ __forceinline__ void load_transposed(uint32_t m_base, int lane, uint32_t (&a)[2][2][4]) {
  const int q = lane >> 3, r = lane & 7;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      ldmatrix_x4_trans(sq_addr(m_base, ks * 16 + (q >> 1) * 8 + r, mt * 2 + (q & 1)), a[mt][ks]);
}

// per-thread column sums of an accumulator tile (rows mt*16 + g + 8h summed; columns dt*8 + 2t + e)
__device__ __forceinline__ void add_columns(float (&acc)[16], const float (&o)[2][8][4]) {
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
#pragma unroll
    for (int e = 0; e < 2; ++e)
      acc[dt * 2 + e] += (o[0][dt][e] + o[0][dt][2 + e]) + (o[1][dt][e] + o[1][dt][2 + e]);
}

// accumulator tile -> bf16 rows of a 64-wide staging tile
__device__ __forceinline__ void stage_out(uint32_t base, int g, int t, const float (&o)[2][8][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = mt * 16 + g + 8 * h;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        const uint32_t v = pack_bf16x2(o[mt][dt][2 * h], o[mt][dt][2 * h + 1]);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(base, row, dt) + 4 * t), "r"(v) : "memory");
      }
    }
}

__global__ void __launch_bounds__(kWarps * 32, 2)
attention_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_ctx,
                         const long long* __restrict__ mask_src, long long num_seqs, int T, int G,
                         int causal, __nv_bfloat16* __restrict__ d_qkv, long long num_items,
                         DropCfg drop, float* __restrict__ d_bias) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2;
  const int t = lane & 3;
  const int R = G * T;
  const uint32_t q_base = smem_u32(smem_raw) + warp * kWarpBytes;
  const uint32_t k_base = q_base + kTileBytes;
  const uint32_t v_base = k_base + kTileBytes;
  const uint32_t o_base = v_base + kTileBytes;   // dO
  const uint32_t p_base = o_base + kTileBytes;   // P~  [32][32]
  const uint32_t s_base = p_base + kSqTileBytes; // dS  [32][32]

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

  const long long total_tokens = num_seqs * T;
  const long long gwarp = blockIdx.x * static_cast<long long>(kWarps) + warp;
  const long long nwarps = gridDim.x * static_cast<long long>(kWarps);
  const float kScale = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  // Bias gradient of the packed in-projection = column sums of dQ | dK | dV. The launcher makes the
  // total warp count a multiple of 12, so every item of this warp has the same head and the partial
  // sums of its 3 x 64 columns stay in registers until the end (column dt*8 + 2t + e of each part).
  float colacc[3][16];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int i = 0; i < 16; ++i) colacc[a][i] = 0.f;
  const int my_head = static_cast<int>(gwarp % kHeads);

  for (long long item = gwarp; item < num_items; item += nwarps) {
    const long long grp = item / kHeads;
    const int head = static_cast<int>(item - grp * kHeads);
    const long long base = grp * R;
    const long long remaining = total_tokens - base;
    const int nrows = remaining < R ? static_cast<int>(remaining) : R;

    {
      const int chunk = lane & 7;
      const int r0 = lane >> 3;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + r0;
        if (row < nrows) {
          const __nv_bfloat16* src = qkv + (base + row) * kQkv + head * kHeadDim + chunk * 8;
          cp_async16(tile_addr(q_base, row, chunk), src);
          cp_async16(tile_addr(k_base, row, chunk), src + kHidden);
          cp_async16(tile_addr(v_base, row, chunk), src + 2 * kHidden);
          cp_async16(tile_addr(o_base, row, chunk), d_ctx + (base + row) * kHidden + head * kHeadDim + chunk * 8);
        } else {
          const uint32_t z = 0;
#pragma unroll
          for (int tl = 0; tl < 4; ++tl)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(
                             tile_addr(q_base + tl * kTileBytes, row, chunk)),
                         "r"(z)
                         : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const bool key_ok = lane < nrows && mask_src[base + (lane < nrows ? lane : 0)] != 0;
    const uint32_t keybits = __ballot_sync(0xffffffffu, key_ok);
    uint32_t kb = 0;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) kb |= ((keybits >> (nt * 8 + 2 * t + e)) & 1u) << (nt * 2 + e);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    float s[2][4][4], dp[2][4][4];
    product_nt(q_base, k_base, lane, s);    // S  = Q K^T
    product_nt(o_base, v_base, lane, dp);   // dP = dO V^T

    // ---- softmax, dropout, dS on the accumulator fragments ----
    uint32_t ds_a[2][2][4];  // dS as bf16 A fragments [m tile][k step over keys]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t ok = allow[mt][h] & kb;
        const int row = mt * 16 + g + 8 * h;
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
        const float mm = (m == -INFINITY) ? 0.f : m;
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
        const float inv = sum > 0.f ? 1.0f / sum : 0.f;
        const int seq0 = (row / T) * T;
        const unsigned long long ebase = (static_cast<unsigned long long>(base + row) * kHeads + head) * 32ull;
        float dsum = 0.f;
        float keepv[8];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const bool on = (ok >> (nt * 2 + e)) & 1u;
            const float p = s[mt][nt][2 * h + e] * inv;
            float keep = 1.0f;
            if (drop.thr16 != 0 && on) {
              const unsigned long long el = ebase + static_cast<unsigned>(nt * 8 + 2 * t + e - seq0);
              keep = drop_mul(drop_bits(drop.key, el >> 1), static_cast<int>(el & 1), drop);
            }
            keepv[nt * 2 + e] = keep;
            const float d = on ? dp[mt][nt][2 * h + e] * keep : 0.f;  // dP through the dropout
            s[mt][nt][2 * h + e] = p;
            dp[mt][nt][2 * h + e] = d;
            dsum += p * d;
          }
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float pt[2], dsv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float p = s[mt][nt][2 * h + e];
            pt[e] = p * keepv[nt * 2 + e];                               // P~ (for dV)
            dsv[e] = p * (dp[mt][nt][2 * h + e] - dsum) * 0.125f;        // dS, scores scale folded in
            dp[mt][nt][2 * h + e] = dsv[e];
          }
          // P~ and dS tiles in shared memory (row = query, column = key): chunk nt, pair at byte 4t
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(sq_addr(p_base, row, nt) + 4 * t),
                       "r"(pack_bf16x2(pt[0], pt[1]))
                       : "memory");
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(sq_addr(s_base, row, nt) + 4 * t),
                       "r"(pack_bf16x2(dsv[0], dsv[1]))
                       : "memory");
        }
      }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        ds_a[mt][j][0] = pack_bf16x2(dp[mt][2 * j][0], dp[mt][2 * j][1]);
        ds_a[mt][j][1] = pack_bf16x2(dp[mt][2 * j][2], dp[mt][2 * j][3]);
        ds_a[mt][j][2] = pack_bf16x2(dp[mt][2 * j + 1][0], dp[mt][2 * j + 1][1]);
        ds_a[mt][j][3] = pack_bf16x2(dp[mt][2 * j + 1][2], dp[mt][2 * j + 1][3]);
      }
    __syncwarp();

    float o[2][8][4];
    uint32_t at[2][2][4];
    // dV = P~^T dO  -> staged in the V tile (V is dead after dP)
    load_transposed(p_base, lane, at);
    product_nn(at, o_base, lane, o);
    add_columns(colacc[2], o);
    __syncwarp();
    stage_out(v_base, g, t, o);
    // dK = dS^T Q   -> staged in the dO tile (dead after dV)
    load_transposed(s_base, lane, at);
    product_nn(at, q_base, lane, o);
    add_columns(colacc[1], o);
    __syncwarp();
    stage_out(o_base, g, t, o);
    // dQ = dS K     -> staged in the Q tile (dead after dK)
    product_nn(ds_a, k_base, lane, o);
    add_columns(colacc[0], o);
    __syncwarp();
    stage_out(q_base, g, t, o);
    __syncwarp();

    {
      const int chunk = lane & 7;
      const int r0 = lane >> 3;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + r0;
        if (row < nrows) {
          __nv_bfloat16* dst = d_qkv + (base + row) * kQkv + head * kHeadDim + chunk * 8;
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(tile_addr(q_base, row, chunk))
                       : "memory");
          *reinterpret_cast<uint4*>(dst) = v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(tile_addr(o_base, row, chunk))
                       : "memory");
          *reinterpret_cast<uint4*>(dst + kHidden) = v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(tile_addr(v_base, row, chunk))
                       : "memory");
          *reinterpret_cast<uint4*>(dst + 2 * kHidden) = v;
        }
      }
    }
    __syncwarp();
  }

  if (d_bias != nullptr) {
    // sum over the 8 row groups (lanes that share t), then lanes 0..3 add their 16 columns of each part
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float v = colacc[a][i];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0 && v != 0.f)
          atomicAdd(d_bias + a * kHidden + my_head * kHeadDim + (i >> 1) * 8 + 2 * t + (i & 1), v);
      }
  }
}

}  // namespace

cudaError_t launch_attention_bwd_mma(const __nv_bfloat16* qkv, const __nv_bfloat16* d_ctx,
                                     const long long* mask_src, long long num_seqs, int T, bool causal,
                                     __nv_bfloat16* d_qkv, cudaStream_t stream, DropCfg drop, float* d_bias) {
  if (T < 1 || T > 32) return cudaErrorInvalidValue;
  if (num_seqs == 0) return cudaSuccess;
  const int G = 32 / T;
  const long long groups = (num_seqs + G - 1) / G;
  const long long items = groups * kHeads;
  const int smem = kWarps * kWarpBytes;
  static unsigned long long smem_done = 0;  // per instantiation, one bit per device
  {
    cudaError_t e = ensure_dynamic_smem(attention_bwd_mma_kernel, smem, &smem_done);
    if (e != cudaSuccess) return e;
  }
  long long blocks = (items + kWarps - 1) / kWarps;
  const long long cap = 148LL * 2 * 8;
  if (blocks > cap) blocks = cap;
  blocks = (blocks + 2) / 3 * 3;  // total warps a multiple of 12: a warp keeps one head (bias-gradient sums)
  attention_bwd_mma_kernel<<<static_cast<unsigned>(blocks), kWarps * 32, smem, stream>>>(
      qkv, d_ctx, mask_src, num_seqs, T, G, causal ? 1 : 0, d_qkv, items, drop, d_bias);
  return cudaGetLastError();
}

}  // namespace stlt
