// Pad-skipping row layout of the spatial phase (bf16 inference path; SURVEY.md 7.3, verified exact against the
// reference's masks: src/modelling/models.py:66-71,79,142-150,192).
//
// The reference encodes every slot of every frame of the padded [B, L, S] grid. Three kinds of rows can never reach
// the logits and are not computed here:
//   * frames at or after lengths[b] (padding frames, datasets.py:262-269): the temporal mask is causal and only frame
//     lengths[b] - 1 is read (models.py:142-150,192), so their spatial encoding is dead;
//   * padded slots (category 0) of a frame whose slots 1.. are ALL padding — the "extract" frame (datasets.py:97-113),
//     frames without confident detections: they are masked as keys (models.py:66-71) and only slot 0 is read
//     (models.py:79), so such a frame is a one-token sequence (its attention output is its own value row).
// Rows of the compact layout:  [ full frames: n_full * S rows, blocks of R = floor(128 / S) * S rows ]
//                              [ single-token frames: n_single rows, from row F1 = nb_full * R ]
// Everything is decided on the device from `lengths` and `categories` (no host synchronisation): three tiny kernels
// classify the frames, scan the counts and write frame_row[b * L + l] (row of the frame's slot 0, or -1) plus a header
// of dynamic counts that the embedding, attention and GEMM kernels read. Allocation sizes use the static upper bounds.
#include "kernels.h"
#include "rowops.cuh"

namespace stlt {

namespace {

constexpr int kPlanThreads = 1024;

// class of frame f: 0 dead, 1 single token, 2 full
__global__ void __launch_bounds__(kPlanThreads)
plan_count_kernel(const long long* __restrict__ categories, const long long* __restrict__ lengths, int L, int S,
                  long long frames, uint8_t* __restrict__ cls, int2* __restrict__ block_counts, int* __restrict__ err_flag) {
  const long long f = blockIdx.x * static_cast<long long>(kPlanThreads) + threadIdx.x;
  int c = 0;
  if (f < frames) {
    const long long b = f / L;
    const int l = static_cast<int>(f - b * L);
    long long len = lengths[b];
    if (len < 1 || len > L) {
      atomicExch(err_flag, 3);
      len = len < 1 ? 1 : L;
    }
    if (l < len) {
      c = 1;
      const long long* row = categories + f * S;
      for (int s = 1; s < S; ++s)
        if (__ldg(row + s) != 0) {
          c = 2;
          break;
        }
    }
    cls[f] = static_cast<uint8_t>(c);
  }
  const int full = __syncthreads_count(c == 2);
  const int single = __syncthreads_count(c == 1);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = make_int2(full, single);
}

// exclusive scan of the per-block counts (in place) + the header of dynamic sizes
__global__ void __launch_bounds__(kPlanThreads)
plan_scan_kernel(int2* __restrict__ block_counts, int num_blocks, int S, int rows_per_block, int* __restrict__ hdr) {
  __shared__ int2 warp_tot[32];
  __shared__ int2 carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = make_int2(0, 0);
  __syncthreads();
  for (int base = 0; base < num_blocks; base += kPlanThreads) {
    const int i = base + threadIdx.x;
    const int2 v = i < num_blocks ? block_counts[i] : make_int2(0, 0);
    int2 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, inc.x, o), y = __shfl_up_sync(0xffffffffu, inc.y, o);
      if (lane >= o) {
        inc.x += x;
        inc.y += y;
      }
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int2 w = warp_tot[lane];
      int2 winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, winc.x, o), y = __shfl_up_sync(0xffffffffu, winc.y, o);
        if (lane >= o) {
          winc.x += x;
          winc.y += y;
        }
      }
      warp_tot[lane] = make_int2(winc.x - w.x, winc.y - w.y);  // exclusive prefix of the warps
    }
    __syncthreads();
    const int2 carry = carry_s;
    const int2 wp = warp_tot[warp];
    if (i < num_blocks) block_counts[i] = make_int2(carry.x + wp.x + inc.x - v.x, carry.y + wp.y + inc.y - v.y);
    __syncthreads();
    if (threadIdx.x == kPlanThreads - 1) carry_s = make_int2(carry.x + wp.x + inc.x, carry.y + wp.y + inc.y);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int n_full = carry_s.x, n_single = carry_s.y;
    const int per_block = rows_per_block / S;
    const int nb_full = (n_full + per_block - 1) / per_block;
    const int f1 = nb_full * rows_per_block;
    const int total = f1 + n_single;
    hdr[kDynFull] = n_full;
    hdr[kDynSingle] = n_single;
    hdr[kDynFullBlocks] = nb_full;
    hdr[kDynSingleRow0] = f1;
    hdr[kDynRows] = total;
    hdr[kDynTiles] = (total + 127) / 128;
    hdr[kDynSingleBlocks] = (n_single + 127) / 128;
    hdr[kDynAttnBlocks] = nb_full + (n_single + 127) / 128;
  }
}

__global__ void __launch_bounds__(kPlanThreads)
plan_assign_kernel(const uint8_t* __restrict__ cls, const int2* __restrict__ block_offsets, const int* __restrict__ hdr,
                   int S, long long frames, int* __restrict__ frame_row) {
  __shared__ int2 warp_tot[32];
  const long long f = blockIdx.x * static_cast<long long>(kPlanThreads) + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = f < frames ? cls[f] : 0;
  const unsigned full_m = __ballot_sync(0xffffffffu, c == 2), single_m = __ballot_sync(0xffffffffu, c == 1);
  const unsigned below = (1u << lane) - 1u;
  if (lane == 0) warp_tot[warp] = make_int2(__popc(full_m), __popc(single_m));
  __syncthreads();
  if (warp == 0) {
    const int2 w = warp_tot[lane];
    int2 winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, winc.x, o), y = __shfl_up_sync(0xffffffffu, winc.y, o);
      if (lane >= o) {
        winc.x += x;
        winc.y += y;
      }
    }
    warp_tot[lane] = make_int2(winc.x - w.x, winc.y - w.y);
  }
  __syncthreads();
  if (f >= frames) return;
  const int2 off = block_offsets[blockIdx.x];
  const int2 wp = warp_tot[warp];
  int row = -1;
  if (c == 2) row = (off.x + wp.x + __popc(full_m & below)) * S;
  else if (c == 1) row = (hdr[kDynSingleRow0] + off.y + wp.y + __popc(single_m & below)) | kSingleFrameFlag;
  frame_row[f] = row;
}

// Rows of the last spatial layer that the temporal stack consumes: slot 0 of every live frame, in padded [B, L] order;
// dead frames (padding) become zero rows (finite, masked as keys downstream).
__global__ void __launch_bounds__(256)
gather_frames_kernel(const float* __restrict__ src_x, const __nv_bfloat16* __restrict__ src_att,
                     const int* __restrict__ frame_row, long long frames, float* __restrict__ dst_x,
                     __nv_bfloat16* __restrict__ dst_att, const float2* __restrict__ src_stats,
                     float2* __restrict__ dst_stats, const __nv_bfloat16* __restrict__ src_hi,
                     const __nv_bfloat16* __restrict__ src_lo, int att_planes, long long src_plane_rows,
                     long long dst_plane_rows) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long r = warp0; r < frames; r += nwarps) {
    const int fr = frame_row[r];
    float4* px = reinterpret_cast<float4*>(dst_x + r * kHidden);
    if (fr < 0) {
      if (src_stats != nullptr && lane < kStatSlots) dst_stats[r * kStatSlots + lane] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < kVec; ++k) px[lane + 32 * k] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int pl = 0; pl < att_planes; ++pl) {
        uint4* da = reinterpret_cast<uint4*>(dst_att + (pl * dst_plane_rows + r) * kHidden);
#pragma unroll
        for (int k = 0; k < 3; ++k) da[lane + 32 * k] = make_uint4(0u, 0u, 0u, 0u);
      }
      continue;
    }
    const long long src = fr & ~kSingleFrameFlag;
    if (src_stats != nullptr && lane < kStatSlots) dst_stats[r * kStatSlots + lane] = src_stats[src * kStatSlots + lane];
    const RowRegs x = src_hi != nullptr ? load_row_hilo(src_hi, src_lo, src, lane) : load_row(src_x, src, lane);
#pragma unroll
    for (int k = 0; k < kVec; ++k) px[lane + 32 * k] = x.v[k];
    for (int pl = 0; pl < att_planes; ++pl) {  // 1 plane (bf16 mode) or hi / lo planes (fp32-parity mode)
      const uint4* sa = reinterpret_cast<const uint4*>(src_att + (pl * src_plane_rows + src) * kHidden);
      uint4* da = reinterpret_cast<uint4*>(dst_att + (pl * dst_plane_rows + r) * kHidden);
#pragma unroll
      for (int k = 0; k < 3; ++k) da[lane + 32 * k] = __ldg(sa + lane + 32 * k);  // 96 x 16 B
    }
  }
}

}  // namespace

size_t compact_plan_scratch_bytes(long long frames) {
  const size_t blocks = static_cast<size_t>((frames + kPlanThreads - 1) / kPlanThreads);
  return (static_cast<size_t>(frames) + 127) / 128 * 128 + blocks * sizeof(int2);
}

long long compact_rows_bound(long long frames, int S) {
  const int rows_per_block = qkv_attention_rows_per_block(S);
  if (rows_per_block == 0) return 0;
  const long long per_block = rows_per_block / S;
  const long long nb_full = (frames + per_block - 1) / per_block;
  return nb_full * rows_per_block + (frames + 127) / 128 * 128 + 128;
}

cudaError_t launch_compact_plan(const long long* categories, const long long* lengths, int B, int L, int S,
                                int* frame_row, int* hdr, void* scratch, int* err_flag, cudaStream_t stream) {
  const long long frames = static_cast<long long>(B) * L;
  const int rows_per_block = qkv_attention_rows_per_block(S);
  if (frames == 0 || rows_per_block == 0) return cudaErrorInvalidValue;
  const int blocks = static_cast<int>((frames + kPlanThreads - 1) / kPlanThreads);
  uint8_t* cls = static_cast<uint8_t*>(scratch);
  int2* counts = reinterpret_cast<int2*>(cls + (static_cast<size_t>(frames) + 127) / 128 * 128);
  plan_count_kernel<<<blocks, kPlanThreads, 0, stream>>>(categories, lengths, L, S, frames, cls, counts, err_flag);
  plan_scan_kernel<<<1, kPlanThreads, 0, stream>>>(counts, blocks, S, rows_per_block, hdr);
  plan_assign_kernel<<<blocks, kPlanThreads, 0, stream>>>(cls, counts, hdr, S, frames, frame_row);
  return cudaGetLastError();
}

cudaError_t launch_gather_frames(const float* src_x, const __nv_bfloat16* src_att, const int* frame_row,
                                 long long frames, float* dst_x, __nv_bfloat16* dst_att, const float2* src_stats,
                                 float2* dst_stats, cudaStream_t stream, const __nv_bfloat16* src_hi,
                                 const __nv_bfloat16* src_lo, int att_planes, long long src_plane_rows,
                                 long long dst_plane_rows) {
  if (frames == 0) return cudaSuccess;
  gather_frames_kernel<<<row_grid(frames, 8), 256, 0, stream>>>(src_x, src_att, frame_row, frames, dst_x, dst_att,
                                                                src_stats, dst_stats, src_hi, src_lo, att_planes,
                                                                src_plane_rows, dst_plane_rows);
  return cudaGetLastError();
}

}  // namespace stlt
