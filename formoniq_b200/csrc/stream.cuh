// stream.cuh — the segmented-reduction engine shared by K3 (assembly gather)
// and K4 (CSR SpMV): "CSR-stream".
//
// A CTA owns a block of consecutive segments (CSR rows / structural non-zeros)
// whose items fit a shared-memory chunk.
//   phase 1  every thread issues UNROLL independent (index, value) loads and
//            then UNROLL independent gathers  src[index]  before touching
//            shared memory — enough bytes in flight per SM to cover HBM latency
//            (the v1 kernel had one dependent load chain per thread and sat at
//            ~45 % of DRAM bandwidth with full occupancy);
//   phase 2  one thread per segment adds the staged items left to right, i.e.
//            in the reference's serial order (no FMA), so results are
//            bit-identical to the CPU path.
// Blocks are cut on the device: block b starts at the first segment whose
// offset is >= b * kChunk (binary search), so a block holds < kChunk + maxlen
// items; blocks that do not fit the staging buffer (a segment longer than the
// slack) take a block-wide strided fallback.
#pragma once
#include "common.cuh"

namespace fq {

constexpr int kStreamThreads = 256;
// A block starts at the first segment at or behind a multiple of kStreamChunk, so it holds up to kStreamChunk + (longest
// segment) items.  The target sits 32 items below one pass of phase 1 (kStreamThreads * kStreamUnroll = 2048): with a
// target of 2048 itself 61 % of the blocks of a FEM matrix (rows of ~16) ran a second, almost empty pass — a full round of
// load latency for a handful of items (SpMV on M1 at N = 128: 0.754 -> 0.699 ms).
constexpr int kStreamChunk = 2016;              // target items per block
constexpr int kStreamCap = 2048 + 512;          // staging capacity (doubles, 20 KB -> 8+ CTAs/SM)
constexpr int kStreamUnroll = 8;                // gathers in flight per thread
constexpr int kStreamSegRegs = 4;               // segment bounds preloaded per thread
// Staged items are padded by one slot every 16: in phase 2 consecutive threads walk consecutive segments (~16 items
// each for the FEM blocks), which without padding start in the same shared-memory bank.
__device__ __forceinline__ uint32_t stream_slot(uint32_t k) { return k + (k >> 4); }
constexpr int kStreamStage = kStreamCap + kStreamCap / 16 + 1;

static __global__ void stream_blocks_kernel(const uint32_t* __restrict__ seg_ptr, uint32_t nseg, uint32_t nblocks,
                                     uint32_t* __restrict__ blocks) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nblocks) return;
  if (b == nblocks) {
    blocks[b] = nseg;
    return;
  }
  const uint32_t target = b * uint32_t(kStreamChunk);
  uint32_t lo = 0, hi = nseg;  // first segment with seg_ptr[s] >= target
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (seg_ptr[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  blocks[b] = lo;
}

// Builds the block table for `nseg` segments with `nitems` items in total.
inline void stream_build_blocks(fq_ctx* ctx, const uint32_t* seg_ptr, size_t nseg, size_t nitems, DevBuf<uint32_t>& blocks,
                                size_t& nblocks) {
  nblocks = nseg == 0 ? 0 : (nitems + kStreamChunk - 1) / kStreamChunk;
  if (nblocks == 0 && nseg > 0) nblocks = 1;  // only empty segments
  blocks.alloc(nblocks + 1);
  if (nseg == 0) {
    FQ_CUDA(cudaMemsetAsync(blocks.p, 0, sizeof(uint32_t), ctx->stream));
    return;
  }
  stream_blocks_kernel<<<unsigned((nblocks + 1 + 255) / 256), 256, 0, ctx->stream>>>(seg_ptr, uint32_t(nseg),
                                                                                    uint32_t(nblocks), blocks.p);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

// Policy interface:
//   static constexpr bool kHasValues;
//   static constexpr bool kCustomSrc;   // true: the gathered operand comes from policy.load(index) instead of src[index]
//   static constexpr bool kGated;       // true: policy.prologue() runs first (e.g. a halo copy by a few CTAs); the
//                                       // blocks with policy.is_late(b) depend on it and wait in policy.gate_wait()
//   __device__ double load(uint32_t index, bool late) const;   // late: the block was gated (CTA-uniform)
//   __device__ void store(uint32_t seg, double sum, bool any_nonzero) const;
template <class Policy>
__global__ void __launch_bounds__(kStreamThreads) stream_reduce_kernel(const uint32_t* __restrict__ blocks,
                                                                        uint32_t nblocks,
                                                                        const uint32_t* __restrict__ seg_ptr,
                                                                        const uint32_t* __restrict__ index,
                                                                        const double* __restrict__ values,
                                                                        const double* __restrict__ src, Policy policy) {
  __shared__ double stage[kStreamStage];
  __shared__ double red[kStreamThreads / 32];
  __shared__ int red_any[kStreamThreads / 32];
  if constexpr (Policy::kGated) policy.prologue();
  for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    bool late = false;  // CTA-uniform: this block reads what the prologue wrote
    if constexpr (Policy::kGated) {
      late = policy.is_late(b);
      if (late) policy.gate_wait();
    }
    const uint32_t s0 = blocks[b], s1 = blocks[b + 1];
    if (s0 >= s1) continue;
    const uint32_t p0 = seg_ptr[s0], p1 = seg_ptr[s1];
    const uint32_t cnt = p1 - p0;
    if (cnt <= uint32_t(kStreamCap)) {
      // segment bounds for phase 2, requested now so that their latency overlaps phase 1
      uint32_t sb[kStreamSegRegs], se[kStreamSegRegs];
#pragma unroll
      for (int j = 0; j < kStreamSegRegs; ++j) {
        const uint32_t s = s0 + j * kStreamThreads + threadIdx.x;
        sb[j] = s < s1 ? __ldg(seg_ptr + s) : 0u;
        se[j] = s < s1 ? __ldg(seg_ptr + s + 1) : 0u;
      }
      // ---- phase 1: coalesced, unrolled loads; all gathers of a batch in flight together
      for (uint32_t base = 0; base < cnt; base += kStreamThreads * kStreamUnroll) {
        uint32_t idx[kStreamUnroll];
        double val[kStreamUnroll];
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
          const uint32_t k = base + u * kStreamThreads + threadIdx.x;
          idx[u] = k < cnt ? __ldg(index + p0 + k) : 0u;
          if (Policy::kHasValues) val[u] = k < cnt ? __ldg(values + p0 + k) : 0.0;
        }
        double g[kStreamUnroll];
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
          const uint32_t k = base + u * kStreamThreads + threadIdx.x;
          g[u] = k < cnt ? (Policy::kCustomSrc ? policy.load(idx[u], late) : __ldg(src + idx[u])) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
          const uint32_t k = base + u * kStreamThreads + threadIdx.x;
          if (k < cnt) stage[stream_slot(k)] = Policy::kHasValues ? __dmul_rn(val[u], g[u]) : g[u];
        }
      }
      __syncthreads();
      // ---- phase 2: one thread per segment, left-to-right sum
#pragma unroll
      for (int j = 0; j < kStreamSegRegs; ++j) {
        const uint32_t s = s0 + j * kStreamThreads + threadIdx.x;
        if (s < s1) {
          double acc = 0.0;
          bool any = false;
          for (uint32_t q = sb[j] - p0; q < se[j] - p0; ++q) {
            const double v = stage[stream_slot(q)];
            any = any || (v != 0.0);
            acc = __dadd_rn(acc, v);
          }
          policy.store(s, acc, any);
        }
      }
      for (uint32_t s = s0 + kStreamSegRegs * kStreamThreads + threadIdx.x; s < s1; s += kStreamThreads) {
        const uint32_t b0 = seg_ptr[s] - p0, b1 = seg_ptr[s + 1] - p0;
        double acc = 0.0;
        bool any = false;
        for (uint32_t q = b0; q < b1; ++q) {
          const double v = stage[stream_slot(q)];
          any = any || (v != 0.0);
          acc = __dadd_rn(acc, v);
        }
        policy.store(s, acc, any);
      }
      __syncthreads();
    } else {
      // ---- fallback: segments too long for the staging buffer (tree order)
      for (uint32_t s = s0; s < s1; ++s) {
        const uint32_t b0 = seg_ptr[s], b1 = seg_ptr[s + 1];
        double acc = 0.0;
        int any = 0;
        for (uint32_t p = b0 + threadIdx.x; p < b1; p += kStreamThreads) {
          const uint32_t ip = __ldg(index + p);
          const double g = Policy::kCustomSrc ? policy.load(ip, late) : __ldg(src + ip);
          const double v = Policy::kHasValues ? __dmul_rn(__ldg(values + p), g) : g;
          any |= (v != 0.0);
          acc = __dadd_rn(acc, v);
        }
        for (int o = 16; o > 0; o >>= 1) {
          acc = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, o));
          any |= __shfl_down_sync(0xffffffffu, any, o);
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc, red_any[threadIdx.x >> 5] = any;
        __syncthreads();
        if (threadIdx.x == 0) {
          double t = 0.0;
          int a = 0;
          for (int w = 0; w < kStreamThreads / 32; ++w) t = __dadd_rn(t, red[w]), a |= red_any[w];
          policy.store(s, t, a != 0);
        }
        __syncthreads();
      }
    }
  }
}

template <class Policy>
inline void stream_reduce(fq_ctx* ctx, const uint32_t* blocks, size_t nblocks, const uint32_t* seg_ptr,
                          const uint32_t* index, const double* values, const double* src, const Policy& policy) {
  if (nblocks == 0) return;
  const size_t cap = size_t(ctx->sm_count) * 8;  // 8 CTAs of 20 KB staging per SM
  const int grid = int(nblocks < cap ? nblocks : cap);
  stream_reduce_kernel<Policy><<<grid, kStreamThreads, 0, ctx->stream>>>(blocks, uint32_t(nblocks), seg_ptr, index, values,
                                                                         src, policy);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

}  // namespace fq
