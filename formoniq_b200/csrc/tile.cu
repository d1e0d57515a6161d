// tile.cu — tile-fused numeric assembly: K1 (element matrices) and K3 (segmented reduction into CSR) in ONE
// persistent kernel, with the element data of a tile living only in shared memory, plus the device builder of its
// plan (which is at the same time the symbolic phase: it produces the structural CSR pattern of every block).
//
// Decomposition = the multi-GPU one, repeated at CTA level: *owner computes*.
//   * vertices are clustered into tiles (closed-form bricks on Kuhn grids, breadth-first clusters on generic meshes);
//   * a tile owns the rows (simplices) whose top vertex it contains, hence whole CSR rows, hence every structural
//     non-zero of those rows;
//   * it visits every cell touching one of its vertices ("cell visit"; halo cells are evaluated by several tiles —
//     FP64 work is cheap here, HBM traffic is not) and stores, for every OWNED local row of the cell, the distinct
//     values of that row of every block's element matrix (sandwiches d*M*D included) into a shared slab;
//   * every owned non-zero is then one lane of a warp-sized record that adds its contributions — plain 16-bit slab
//     indices — left to right in ascending cell order and stores the sum once.
//
// tile_fused_kernel: 16 producer warps (one cell visit per thread) evaluate the generated staged tapes (elmat_gen.cuh: stage A = geometry, then one
// stage group per mass grade), 16 consumer warps stream the tile's records through private TMA double buffers
// (cp.async.bulk + mbarrier).  Every stage group has its own region of the slab and a full/empty mbarrier pair, so
// the producers refill the region of group g for tile i+1 as soon as every consumer warp is past the group-g records
// of tile i: the FP64 pipe works in the shadow of the gather without a second slab.
//
// tile_build_kernel: one CTA per tile builds the plan from the mesh's face tables alone (no global sort of the
// element entries): local face numbering, owned-row slots, block-wide radix sort of the tile's (row, col) keys,
// run/record layout — first a counting pass (row lengths, chunks per tile), then the emitting pass.  The same bytes
// as tile_plan.hpp's host reference builder (FQ_TILE_BUILD=host), which the CPU tests interpret against the oracle.
//
// Reference path replaced: formoniq/src/galerkin.rs:138-188 (assemble_matrix) + hodge.rs:62-72 (HodgeBlocks).
#include <cub/cub.cuh>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <type_traits>

#include "elmat_gen.cuh"
#include "internal.hpp"
#include "kuhn.hpp"
#include "tile_plan.hpp"

namespace fq {
using namespace tp;

struct FusedParams {
  const TileHdr* tiles;
  uint32_t ntiles;
  const uint32_t* cv_rec;
  int cv_words;
  const uint16_t* gbase;          // group records: first row slot of every (class, local row) of 32 cell visits
  const double* lengths;
  uint32_t edge_lo;
  const unsigned char* stream;
  double* values[kMaxBlocks];
  uint8_t* keep[kMaxBlocks];      // structural mode: "some contribution != 0.0" flags (galerkin.rs:173); may be null
  int* changed;
  uint32_t chunk_rotation;
  uint32_t ring_off, mbar_off;
  uint32_t group_pack;            // stage group of block b in byte b
  int debug;                      // FQ_TILE_DEBUG: 1 skip K1, 2 skip records
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (bytes % 16 == 0, 16-byte aligned)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Producer side: one cell visit's stores.  masks = owned-row masks (8 bits per row class), slot = the row slots of its
// owned rows; the value of column slot CS of local row R of block B goes to Fn::sb(B) + CS * Fn::plane(C) + slot — the
// region bases and plane strides are compile-time constants of the generated set, so a store is one predicated STS with
// an immediate offset.  Lanes of a warp that own row R hold consecutive slots: the store is conflict-free.
template <class Fn>
struct FusedSink {
  double* __restrict__ slab;
  uint32_t masks;
  uint32_t slot[kMaxClasses][kMaxLocal];
  template <int B, int R, int CS, int D, int C>
  __device__ __forceinline__ void put(double v) const {
    constexpr uint32_t off = Fn::sb(B) + uint32_t(CS) * Fn::plane(C);
    if ((masks >> (8 * C + R)) & 1u) slab[slot[C][R] + off] = v;
  }
};

// non-zero bits of a double (x != 0.0 for finite and non-finite values alike; -0.0 counts as zero)
__device__ __forceinline__ uint32_t nz_bits(double x) {
  return (uint32_t(__double2hiint(x)) << 1) | uint32_t(__double2loint(x));
}

// What a record lane does with its sum: store it (and the classification) as the stream's destination says.
template <bool COMPACT>
__device__ __forceinline__ void finish_lane(uint32_t dest, double acc, uint32_t any, double* __restrict__ vals,
                                            uint8_t* __restrict__ keep, int* __restrict__ changed) {
  if (COMPACT) {
    // dest = 0: padding lane (it read slab[0]); dest = 1: a dropped non-zero, every contribution must still be an exact
    // zero; dest >= 2: a kept one, some contribution must be non-zero (galerkin.rs:173)
    if (dest != kPadDest && ((dest != kNoDest) != (any != 0u))) *changed = 1;
    if (dest > kNoDest) vals[dest - 2u] = acc;
  } else if (dest > kNoDest) {
    vals[dest - 2u] = acc;
    if (keep) keep[dest - 2u] = any != 0u ? 1 : 0;
  }
}

// L contributions of the two chains of a lane, every load issued before the first add (memory-level parallelism is
// what the consumer warps live on: they are bound by shared-memory latency, not by issue slots).
template <int L>
__device__ __forceinline__ void gather_fixed(const uint16_t* __restrict__ e0, const uint16_t* __restrict__ e1, uint32_t stride,
                                             const double* __restrict__ slab, double& acc0, double& acc1, uint32_t& any0,
                                             uint32_t& any1) {
  uint32_t c0[L], c1[L];
#pragma unroll
  for (int j = 0; j < L; ++j) {
    c0[j] = e0[j * stride];
    c1[j] = e1[j * stride];
  }
  double x0[L], x1[L];
#pragma unroll
  for (int j = 0; j < L; ++j) {
    x0[j] = slab[c0[j]];
    x1[j] = slab[c1[j]];
  }
#pragma unroll
  for (int j = 0; j < L; ++j) {
    any0 |= nz_bits(x0[j]);
    any1 |= nz_bits(x1[j]);
    acc0 = __dadd_rn(acc0, x0[j]);
    acc1 = __dadd_rn(acc1, x1[j]);
  }
}
// Any record: two non-zeros per lane (columns lane and lane + 32 of a wide record; a narrow record — 32 or 16 lanes,
// for non-zeros with many contributions — runs the second chain on the first column and discards it; lanes beyond the
// record shadow the first ones and never store), left-to-right sums.
template <bool COMPACT>
__device__ __forceinline__ void record_any(const unsigned char* __restrict__ rp, uint32_t L, uint32_t stride, int lane,
                                           const double* __restrict__ slab, double* __restrict__ vals, uint8_t* __restrict__ keep,
                                           int* __restrict__ changed) {
  const uint32_t* destp = reinterpret_cast<const uint32_t*>(rp + kRecHdr);
  const uint32_t l0 = uint32_t(lane) & (stride - 1u);
  const uint32_t dest0 = uint32_t(lane) < stride ? destp[l0] : kPadDest;
  const uint32_t dest1 = stride == 64u ? destp[lane + 32] : kPadDest;
  const uint16_t* __restrict__ e0 = reinterpret_cast<const uint16_t*>(rp + kRecHdr + 4 * stride) + l0;
  const uint16_t* __restrict__ e1 = e0 + (stride == 64u ? 32 : 0);
  double acc0 = 0.0, acc1 = 0.0;
  uint32_t any0 = 0, any1 = 0;
  switch (L) {
    case 1: gather_fixed<1>(e0, e1, stride, slab, acc0, acc1, any0, any1); break;
    case 2: gather_fixed<2>(e0, e1, stride, slab, acc0, acc1, any0, any1); break;
    case 3: gather_fixed<3>(e0, e1, stride, slab, acc0, acc1, any0, any1); break;
    case 4: gather_fixed<4>(e0, e1, stride, slab, acc0, acc1, any0, any1); break;
    default: {
#pragma unroll 1
      for (uint32_t j = 0; j + 4 <= L; j += 4) {
        gather_fixed<4>(e0, e1, stride, slab, acc0, acc1, any0, any1);
        e0 += 4 * stride;
        e1 += 4 * stride;
      }
      switch (L & 3u) {
        case 1: gather_fixed<1>(e0, e1, stride, slab, acc0, acc1, any0, any1); break;
        case 2: gather_fixed<2>(e0, e1, stride, slab, acc0, acc1, any0, any1); break;
        case 3: gather_fixed<3>(e0, e1, stride, slab, acc0, acc1, any0, any1); break;
        default: break;
      }
    }
  }
  finish_lane<COMPACT>(dest0, acc0, any0, vals, keep, changed);
  finish_lane<COMPACT>(dest1, acc1, any1, vals, keep, changed);
}

// NP producer / NC consumer warps; PR / CR registers per producer / consumer thread after setmaxnreg (the launch
// allocates LR = 65536 / threads rounded down to 8 per thread; what the consumers release must cover what the producers
// acquire).  COMPACT: the stream's destinations target the value-dependent pattern and the classification is verified.
template <class Fn, int NE, int NP, int NC, int PR, int CR, bool COMPACT>
__global__ void __launch_bounds__(32 * (NP + NC), 1) tile_fused_kernel(const __grid_constant__ FusedParams P) {
  constexpr int kThreads = 32 * (NP + NC);
  constexpr int LR = 65536 / kThreads / 8 * 8;
  static_assert(PR % 8 == 0 && CR % 8 == 0 && CR <= LR && PR >= LR && 32 * NC * (LR - CR) >= 32 * NP * (PR - LR), "register split");
  static_assert(NC <= kMaxConsumerWarps, "ring size");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NEE = NE > 0 ? NE : 1;
  constexpr int NM = Fn::kMid;
  constexpr int NG = Fn::kGroups;
  constexpr int NPT = 32 * NP;                       // producer threads
  constexpr int CPT = (kMaxCv + NPT - 1) / NPT;      // cell visits per producer thread
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* slab = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + P.mbar_off);
  uint64_t* tma_bar = bars;                                    // [consumer warp][2]
  uint64_t* g_full = bars + kMaxConsumerWarps * kSlotsPerWarp; // [group]: the producers stored the group's rows of the tile
  uint64_t* g_empty = g_full + kMaxGroups;                     // [group]: every consumer warp is past the group's records
  if (tid == 0) {
    for (int g = 0; g < kMaxGroups; ++g) {
      mbar_init(&g_full[g], NP);
      mbar_init(&g_empty[g], NC);
    }
    for (int i = 0; i < kSlotsPerWarp * NC; ++i) mbar_init(&tma_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < kZeroSlots) slab[tid] = 0.0;
  __syncthreads();
  const uint32_t G = gridDim.x, t0 = blockIdx.x;
  if (warp < NP) {
    // ------------------------------------------------------------------ producers: K1, CPT cell visits per thread
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PR));
    const uint32_t W = uint32_t(P.cv_words);
    uint32_t cvb = 0, ncv = 0, cvb_n = 0, ncv_n = 0;
    auto load_range = [&](uint64_t t, uint32_t& b, uint32_t& n) {
      b = 0, n = 0;
      if (t < uint64_t(P.ntiles) && !(P.debug & 1)) {
        b = __ldg(&P.tiles[t].cv_begin);
        n = __ldg(&P.tiles[t].ncv);
      }
    };
    auto load_ids = [&](uint32_t b, uint32_t n, uint32_t (*eid)[NEE]) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint32_t c = uint32_t(tid) + uint32_t(j) * NPT;
        if (c < n) {
          const uint32_t* rec = P.cv_rec + size_t(b + c) * W;
#pragma unroll
          for (int e = 0; e < NE; ++e) eid[j][e] = __ldg(rec + e);
        }
      }
    };
    load_range(t0, cvb, ncv);
    load_range(uint64_t(t0) + G, cvb_n, ncv_n);
    uint32_t eid[CPT][NEE];
    load_ids(cvb, ncv, eid);
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t it = 0;; ++it) {
      const uint64_t t = uint64_t(t0) + uint64_t(it) * G;
      if (t >= uint64_t(P.ntiles)) break;
      // edge lengths and stage A (metric, inverse, volume) of this tile's cell visits: the consumers are still busy
      // with the previous tile, so this latency is off the critical path
      double mid[CPT][NM];
      FusedSink<Fn> sink[CPT];
      const uint32_t gb_slot = __ldg(&P.tiles[t].gb_slot);
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint32_t c = uint32_t(tid) + uint32_t(j) * NPT;
        sink[j].slab = slab;
        sink[j].masks = c < ncv ? __ldg(P.cv_rec + size_t(cvb + c) * W + NE) : 0u;
        // row slots: group record + rank among the lanes of the warp (= group of 32 cell visits) that own the row
        const uint16_t* grec = P.gbase + (size_t(gb_slot) + (c >> 5)) * kGroupWords;
#pragma unroll
        for (int cl = 0; cl < kMaxClasses; ++cl) {
          const int nrows = cl == 0 ? Fn::kRows0 : Fn::kRows1;
#pragma unroll
          for (int R = 0; R < kMaxLocal; ++R) {
            sink[j].slot[cl][R] = 0u;
            if (R >= nrows) continue;
            const uint32_t owners = __ballot_sync(0xFFFFFFFFu, (sink[j].masks >> (8 * cl + R)) & 1u);
            // (lanes beyond the tile's cell visits take part in the ballot only: their group record may not exist)
            const uint32_t first = c < ncv ? uint32_t(__ldg(grec + cl * kMaxLocal + R)) : 0u;
            sink[j].slot[cl][R] = first + uint32_t(__popc(owners & lt_mask));
          }
        }
        if (c < ncv) {
          double sl[NEE];
#pragma unroll
          for (int e = 0; e < NE; ++e) sl[e] = __ldg(P.lengths + (eid[j][e] - P.edge_lo));
          Fn::a(sl, mid[j]);
        }
      }
      load_ids(cvb_n, ncv_n, eid);  // ids of the next tile: in flight while the groups are evaluated
      uint32_t cvb_nn, ncv_nn;
      load_range(t + 2 * uint64_t(G), cvb_nn, ncv_nn);
      const uint32_t par_prev = (it - 1u) & 1u;
      auto stage = [&](auto gc) {
        constexpr int g = decltype(gc)::value;
        if (g >= NG) return;
        if (it >= 1) mbar_wait(&g_empty[g], par_prev);  // every consumer warp is past this group's records of the previous tile
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const uint32_t c = uint32_t(tid) + uint32_t(j) * NPT;
          if (c < ncv) Fn::template g<g>(mid[j], sink[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&g_full[g]);
      };
      stage(std::integral_constant<int, 0>());
      stage(std::integral_constant<int, 1>());
      stage(std::integral_constant<int, 2>());
      cvb = cvb_n, ncv = ncv_n;
      cvb_n = cvb_nn, ncv_n = ncv_nn;
    }
  } else {
    // ------------------------------------------------------------------ consumers: K3
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CR));
    const int cw = warp - NP;
    unsigned char* myring = smem_raw + P.ring_off + size_t(cw) * kSlotsPerWarp * kChunkBytes;
    uint64_t* mybar = tma_bar + cw * kSlotsPerWarp;
    uint32_t n_issued = 0, n_consumed = 0;
    uint32_t cur_it = 0xFFFFFFFFu, cur_chunk = 0, cur_end = 0;
    auto load_chunks = [&](uint64_t t, uint32_t& a, uint32_t& e) {
      a = 0, e = 0;
      if (t < uint64_t(P.ntiles)) {
        a = __ldg(&P.tiles[t].chunk_begin);
        e = a + ((P.debug & 2) ? 0u : __ldg(&P.tiles[t].nchunks));
      }
    };
    uint32_t c0, c1, c0n = 0, c1n = 0;
    load_chunks(t0, c0, c1);
    // Chunk c of a tile goes to the warp (c - rotation) mod NC, the rotation advancing with every tile: without it the
    // same warps would get the extra chunk of every tile.
    const uint32_t rot_step = P.chunk_rotation;
    auto lane_of = [&](uint32_t iter) { return (uint32_t(cw) + iter * rot_step) % uint32_t(NC); };
    for (uint32_t it = 0;; ++it) {
      const uint64_t t = uint64_t(t0) + uint64_t(it) * G;
      if (t >= uint64_t(P.ntiles)) break;
      const bool has_next = t + G < uint64_t(P.ntiles);
      load_chunks(t + G, c0n, c1n);
      auto issue_more = [&]() {  // keep this warp's double buffer full, crossing into the next tile when this one is done
        while (n_issued - n_consumed < uint32_t(kSlotsPerWarp)) {
          if (cur_chunk >= cur_end) {
            if (cur_it == it && has_next) {
              cur_it = it + 1;
              cur_chunk = c0n + lane_of(it + 1);
              cur_end = c1n;
              if (cur_chunk >= cur_end) break;
            } else {
              break;
            }
          }
          if (lane == 0) {
            uint64_t* bar = &mybar[n_issued & 1u];
            mbar_expect_tx(bar, kChunkBytes);
            tma_load_1d(myring + (n_issued & 1u) * kChunkBytes, P.stream + size_t(cur_chunk) * kChunkBytes, kChunkBytes, bar);
          }
          cur_chunk += NC;
          ++n_issued;
        }
      };
      if (cur_it != it) {
        cur_it = it;
        cur_chunk = c0 + lane_of(it);
        cur_end = c1;
      }
      issue_more();
      const uint32_t par = it & 1u;
      uint32_t g_cur = 0;
      mbar_wait(&g_full[0], par);
      for (uint32_t c = c0 + lane_of(it); c < c1; c += NC) {
        mbar_wait(&mybar[n_consumed & 1u], (n_consumed >> 1) & 1u);
        const unsigned char* chunk = myring + (n_consumed & 1u) * kChunkBytes;
        const uint32_t nrec = *reinterpret_cast<const uint32_t*>(chunk);
        const unsigned char* rp = chunk + kChunkHdr;
        uint32_t h = *reinterpret_cast<const uint32_t*>(rp);  // header of the first record (a chunk holds at least one)
        for (uint32_t r = 0; r < nrec; ++r) {
          const uint32_t L = h & 0xFFu, b = (h >> 8) & 3u, stride = h >> 16;  // stride = lanes of the record: 64, 32 or 16
          const unsigned char* rec = rp;
          rp += kRecHdr + stride * (4u + 2u * L);
          // header of the next record: off the critical path
          if (r + 1 < nrec) h = *reinterpret_cast<const uint32_t*>(rp);
          const uint32_t g = (P.group_pack >> (8 * b)) & 0xFFu;
          while (g_cur < g) {  // done with a stage group: its slab region goes back to the producers; wait for the next one
            __syncwarp();
            if (lane == 0) mbar_arrive(&g_empty[g_cur]);
            ++g_cur;
            mbar_wait(&g_full[g_cur], par);
          }
          double* __restrict__ vals = P.values[b];
          uint8_t* __restrict__ keep = COMPACT ? nullptr : P.keep[b];
          record_any<COMPACT>(rec, L, stride, lane, slab, vals, keep, P.changed);
        }
        __syncwarp();  // every lane is done reading the slot before it is refilled
        ++n_consumed;
        issue_more();
      }
      // Release the remaining groups.  A warp always observes full(g) before it arrives on empty(g): otherwise it could
      // run a tile ahead of a slow warp and complete a phase of empty(g) while that warp still reads the region.
      __syncwarp();
      for (;;) {
        if (lane == 0) mbar_arrive(&g_empty[g_cur]);
        if (++g_cur >= uint32_t(NG)) break;
        mbar_wait(&g_full[g_cur], par);
      }
      c0 = c0n;
      c1 = c1n;
    }
  }
}

// After the first (structural) pass of a dropping assembly: dest -> position in the compacted pattern, or kNoDest.
struct RetargetArgs {
  const uint8_t* keep[kMaxBlocks];
  const uint32_t* pos[kMaxBlocks];
};
__global__ void retarget_kernel(unsigned char* __restrict__ stream, uint32_t nchunks, RetargetArgs A) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t wstride = gridDim.x * (blockDim.x >> 5);
  for (uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nchunks; c += wstride) {
    unsigned char* chunk = stream + size_t(c) * kChunkBytes;
    const uint32_t nrec = *reinterpret_cast<const uint32_t*>(chunk);
    unsigned char* rp = chunk + kChunkHdr;
    for (uint32_t r = 0; r < nrec; ++r) {
      const uint32_t h = *reinterpret_cast<const uint32_t*>(rp);
      const uint32_t L = h & 0xFFu, b = (h >> 8) & 3u, lanes = h >> 16;
      uint32_t* dest = reinterpret_cast<uint32_t*>(rp + kRecHdr);
      for (uint32_t l = lane; l < lanes; l += 32u) {
        const uint32_t d = dest[l];
        if (d > kNoDest) dest[l] = A.keep[b][d - 2u] ? A.pos[b][d - 2u] + 2u : kNoDest;
      }
      rp += kRecHdr + lanes * (4u + 2u * L);
    }
  }
}

// ------------------------------------------------------------------ device plan builder
constexpr int kBT = 512;                      // threads of a builder CTA
constexpr int kIPT = kMaxEntries / kBT;       // entries per thread of the block-wide sorts
constexpr int kIdIPT = kMaxCv * kMaxLocal / kBT;
static_assert(kIPT * kBT == kMaxEntries && kIdIPT * kBT == kMaxCv * kMaxLocal, "sort tiling");
typedef cub::BlockRadixSort<uint32_t, kBT, kIPT, uint16_t> EntSort;
typedef cub::BlockRadixSort<uint32_t, kBT, kIdIPT, uint16_t> IdSort;
typedef cub::BlockScan<uint32_t, kBT> BScan;

struct BuildParams {
  SetDesc S;
  int gtab[4];                 // grade -> local-id table (0/1), -1 unused
  int id_bits[4];              // significant bits of the global face ids per grade
  const uint32_t* faces[4];
  const uint32_t* vertex_tile;
  uint32_t v_lo;
  uint32_t ntiles;
  const uint32_t* tile_cv_ptr;
  const uint32_t* tile_cv_cells;
  uint32_t slab_base[kMaxBlocks];    // emitting pass in: region of every block (the same for all tiles)
  uint32_t plane[kMaxClasses];       // emitting pass in: plane strides
  uint32_t* rs_max;                  // counting pass out: largest row-slot count per row class
  uint16_t* gbase;                   // emitting pass out: group records
  uint32_t* row_ptr[kMaxBlocks];     // counting pass: row lengths out; emitting pass: structural row_ptr in
  uint32_t* col_idx[kMaxBlocks];
  uint32_t* tile_nchunks;            // counting pass out
  unsigned long long* ncontrib;      // counting pass out: owned element entries per block
  const uint32_t* tile_chunk_ptr;    // emitting pass in
  TileHdr* tiles;
  uint32_t* cv_rec;
  unsigned char* stream;
  int* err;
};

struct RunInfo {
  uint32_t p0, lanes, size, fits, m, off0, fresh0;
};
struct BuildSmem {
  union {
    typename EntSort::TempStorage ent;
    typename IdSort::TempStorage id;
    typename BScan::TempStorage scan;
  } tmp;
  uint32_t skey[kMaxEntries];
  uint16_t sval[kMaxEntries];
  uint16_t nz_first[kMaxEntries + 2];
  uint16_t ord[kMaxEntries];
  uint8_t Lp[kMaxEntries];
  uint16_t lid[2][kMaxCv * kMaxLocal];
  uint32_t glob[2][kMaxCv * kMaxLocal];
  uint32_t cells[kMaxCv];
  uint8_t cmask[kMaxClasses][kMaxCv];
  uint16_t rs[kMaxClasses][kMaxCv * kMaxLocal];  // row slot of (cell visit, local row)
  uint16_t gcnt[kMaxClasses][(kMaxCv / 32) * kMaxLocal];  // owners of (group, local row), then their first row slot
  uint16_t ebase[kMaxCv];
  uint16_t rowstart[kMaxCv * kMaxLocal];
  uint16_t runstart[64];
  RunInfo run[64];
  uint32_t RS[kMaxClasses];
  uint32_t E, Q, bad;
  // placement cursor of the tile's stream (tile_plan.hpp: Cursor)
  uint32_t cur_off, cur_chunk_start, cur_in_chunk, cur_open, cur_next_start;
};

template <bool EMIT>
__global__ void __launch_bounds__(kBT, 1) tile_build_kernel(const __grid_constant__ BuildParams P) {
  extern __shared__ __align__(16) unsigned char build_smem_raw[];
  BuildSmem& sm = *reinterpret_cast<BuildSmem*>(build_smem_raw);
  const SetDesc& S = P.S;
  const uint32_t tid = threadIdx.x;
  for (uint32_t t = blockIdx.x; t < P.ntiles; t += gridDim.x) {
    __syncthreads();
    const uint32_t cv0 = P.tile_cv_ptr[t], ncv = P.tile_cv_ptr[t + 1] - cv0;
    if (tid == 0) {
      sm.bad = ncv > uint32_t(kMaxCv) ? 1u : 0u;
      sm.cur_off = 0, sm.cur_chunk_start = 0, sm.cur_in_chunk = 0, sm.cur_open = 0, sm.cur_next_start = 0;
    }
    if (tid < ncv && tid < uint32_t(kMaxCv)) sm.cells[tid] = P.tile_cv_cells[cv0 + tid];
    __syncthreads();
    if (sm.bad) {
      if (tid == 0) atomicExch(P.err, 1);
      if (!EMIT && tid == 0) P.tile_nchunks[t] = 0;
      continue;
    }
    // ---- row classes: owned-row masks; row slots numbered (group of 32 cell visits, local row, cell visit)
    for (int c = 0; c < S.nclasses; ++c) {
      const int g = S.class_grade[c], nl = S.nl[g];
      uint32_t m = 0;
      if (tid < ncv) {
        const size_t cell = sm.cells[tid];
        for (int r = 0; r < nl; ++r) {
          const uint32_t row = P.faces[g][cell * nl + r];
          const uint32_t topv = P.faces[0][cell * S.nv + S.top[g][r]];
          if (P.vertex_tile[topv - P.v_lo] == t && row >= S.class_lo[c] && row < S.class_hi[c]) m |= 1u << r;
        }
        sm.cmask[c][tid] = uint8_t(m);
      }
      const uint32_t grp = tid >> 5, lane = tid & 31u;
      uint32_t rank[kMaxLocal];
#pragma unroll
      for (int r = 0; r < kMaxLocal; ++r) {
        const uint32_t owners = __ballot_sync(0xFFFFFFFFu, (m >> r) & 1u);
        rank[r] = uint32_t(__popc(owners & ((1u << lane) - 1u)));
        if (lane == 0) sm.gcnt[c][grp * kMaxLocal + r] = uint16_t(__popc(owners));
      }
      __syncthreads();
      if (tid == 0) {  // exclusive scan over (group, row): 16 x 6 entries
        uint32_t run = 0;
        const uint32_t ngroups = (ncv + 31u) / 32u;
        for (uint32_t i = 0; i < ngroups * kMaxLocal; ++i) {
          const uint32_t cnt = sm.gcnt[c][i];
          sm.gcnt[c][i] = uint16_t(run);
          run += cnt;
        }
        sm.RS[c] = run;
        if (run > 0xFFFFu) sm.bad = 1;
      }
      __syncthreads();
      if (tid < ncv) {
#pragma unroll
        for (int r = 0; r < kMaxLocal; ++r) sm.rs[c][tid * kMaxLocal + r] = uint16_t(uint32_t(sm.gcnt[c][grp * kMaxLocal + r]) + rank[r]);
      }
      if (EMIT) {
        const uint32_t gb_slot = (cv0 >> 5) + t;
        const uint32_t ngroups = (ncv + 31u) / 32u;
        for (uint32_t i = tid; i < ngroups * kMaxLocal; i += kBT)
          P.gbase[(size_t(gb_slot) + i / kMaxLocal) * kGroupWords + size_t(c) * kMaxLocal + i % kMaxLocal] = sm.gcnt[c][i];
      }
      __syncthreads();
    }
    if (!EMIT && tid < uint32_t(S.nclasses)) atomicMax(&P.rs_max[tid], sm.RS[tid]);
    if (sm.bad) {
      if (tid == 0) atomicExch(P.err, 2);
      if (!EMIT && tid == 0) P.tile_nchunks[t] = 0;
      continue;
    }
    if (EMIT) {
      if (tid < ncv) {
        const size_t cell = sm.cells[tid];
        uint32_t* rec = P.cv_rec + size_t(cv0 + tid) * S.cv_words;
        for (int e = 0; e < S.ne; ++e) rec[e] = P.faces[1][cell * S.ne + e];
        uint32_t word = 0;
        for (int c = 0; c < S.nclasses; ++c) word |= uint32_t(sm.cmask[c][tid]) << (8 * c);
        rec[S.ne] = word;
      }
      if (tid == 0) {
        TileHdr H;
        H.cv_begin = cv0, H.ncv = ncv;
        H.chunk_begin = P.tile_chunk_ptr[t];
        H.nchunks = P.tile_chunk_ptr[t + 1] - P.tile_chunk_ptr[t];
        H.gb_slot = (cv0 >> 5) + t;
        H.pad[0] = H.pad[1] = H.pad[2] = 0;
        P.tiles[t] = H;
      }
    }
    // ---- local face numbering (monotone in the global id) of the grades the blocks use
    for (int g = 0; g <= S.n; ++g) {
      const int x = P.gtab[g];
      if (x < 0) continue;
      const uint32_t nl = uint32_t(S.nl[g]), N = ncv * nl;
      uint32_t keys[kIdIPT];
      uint16_t vals[kIdIPT];
#pragma unroll
      for (int k = 0; k < kIdIPT; ++k) {
        const uint32_t idx = tid * kIdIPT + k;
        keys[k] = 0xFFFFFFFFu, vals[k] = 0xFFFFu;
        if (idx < N) {
          keys[k] = P.faces[g][size_t(sm.cells[idx / nl]) * nl + idx % nl];
          vals[k] = uint16_t(idx);
        }
      }
      IdSort(sm.tmp.id).Sort(keys, vals, 0, P.id_bits[g]);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kIdIPT; ++k) {
        sm.skey[tid * kIdIPT + k] = keys[k];
        sm.sval[tid * kIdIPT + k] = vals[k];
      }
      __syncthreads();
      uint32_t heads = 0;
#pragma unroll
      for (int k = 0; k < kIdIPT; ++k) {
        const uint32_t idx = tid * kIdIPT + k;
        if (idx < N && (idx == 0 || sm.skey[idx] != sm.skey[idx - 1])) heads |= 1u << k;
      }
      uint32_t prefix, total;
      BScan(sm.tmp.scan).ExclusiveSum(uint32_t(__popc(heads)), prefix, total);
#pragma unroll
      for (int k = 0; k < kIdIPT; ++k) {
        const uint32_t idx = tid * kIdIPT + k;
        if (idx >= N) continue;
        if (heads >> k & 1u) {
          sm.glob[x][prefix] = sm.skey[idx];
          ++prefix;
        }
        sm.lid[x][sm.sval[idx]] = uint16_t(prefix - 1u);
      }
      __syncthreads();
    }
    // ---- blocks
    unsigned char* sbase = EMIT ? P.stream + size_t(P.tile_chunk_ptr[t]) * kChunkBytes : nullptr;
    for (int b = 0; b < S.nblocks; ++b) {
      const BlockDesc& B = S.blk[b];
      if (B.empty) continue;
      const int c = B.rclass, xt = P.gtab[B.tg], xr = P.gtab[B.rg];
      const uint32_t nt = uint32_t(B.nt), nr = uint32_t(B.nr);
      // entries of the owned rows, in (cell visit, row, column) order
      {
        const uint32_t m = tid < ncv ? uint32_t(sm.cmask[c][tid]) : 0u;
        uint32_t ebase, E;
        BScan(sm.tmp.scan).ExclusiveSum(uint32_t(__popc(m)) * nr, ebase, E);
        if (tid == 0) {
          sm.E = E;
          if (E > uint32_t(kMaxEntries)) sm.bad = 1;
          if (!EMIT) atomicAdd(&P.ncontrib[b], (unsigned long long)E);
        }
        __syncthreads();
        if (sm.bad) break;
        if (tid < ncv) {
          uint32_t e = ebase;
          for (uint32_t r = 0; r < nt; ++r) {
            if (!(m >> r & 1u)) continue;
            const uint32_t rl = sm.lid[xt][tid * nt + r];
            const uint32_t slot = sm.rs[c][tid * kMaxLocal + r];
            for (uint32_t j = 0; j < nr; ++j) {
              const uint32_t cs = B.cs[r * nr + j];
              sm.skey[e] = (rl << 12) | uint32_t(sm.lid[xr][tid * nr + j]);
              sm.sval[e] = uint16_t(cs == 0xFFu ? 0u : P.slab_base[b] + cs * P.plane[c] + slot);
              ++e;
            }
          }
        }
        for (uint32_t e = E + tid; e < uint32_t(kMaxEntries); e += kBT) sm.skey[e] = 0xFFFFFFFFu;
        __syncthreads();
      }
      const uint32_t E = sm.E;
      {
        uint32_t keys[kIPT];
        uint16_t vals[kIPT];
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
          keys[k] = sm.skey[tid * kIPT + k];
          vals[k] = sm.sval[tid * kIPT + k];
        }
        __syncthreads();
        EntSort(sm.tmp.ent).Sort(keys, vals, 0, 24);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
          sm.skey[tid * kIPT + k] = keys[k];
          sm.sval[tid * kIPT + k] = vals[k];
        }
        __syncthreads();
      }
      // non-zeros = runs of equal keys (CSR order: local row, column)
      {
        uint32_t heads = 0;
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
          const uint32_t e = tid * kIPT + k;
          if (e < E && (e == 0 || sm.skey[e] != sm.skey[e - 1])) heads |= 1u << k;
        }
        uint32_t prefix, Q;
        BScan(sm.tmp.scan).ExclusiveSum(uint32_t(__popc(heads)), prefix, Q);
#pragma unroll
        for (int k = 0; k < kIPT; ++k)
          if (heads >> k & 1u) sm.nz_first[prefix++] = uint16_t(tid * kIPT + k);
        if (tid == 0) {
          sm.Q = Q;
          sm.nz_first[Q] = uint16_t(E);
        }
        __syncthreads();
      }
      const uint32_t Q = sm.Q;
      // rows of the tile: first non-zero, length (counting pass), column ids (emitting pass)
      for (uint32_t q = tid; q < Q; q += kBT) {
        const uint32_t rl = sm.skey[sm.nz_first[q]] >> 12;
        if (q == 0 || (sm.skey[sm.nz_first[q - 1]] >> 12) != rl) sm.rowstart[rl] = uint16_t(q);
      }
      __syncthreads();
      for (uint32_t q = tid; q < Q; q += kBT) {
        const uint32_t key = sm.skey[sm.nz_first[q]];
        const uint32_t rl = key >> 12;
        const uint32_t grow = sm.glob[xt][rl] - B.row_begin;
        if (!EMIT) {
          if (q + 1 == Q || (sm.skey[sm.nz_first[q + 1]] >> 12) != rl) P.row_ptr[b][grow] = q + 1u - uint32_t(sm.rowstart[rl]);
        } else {
          P.col_idx[b][P.row_ptr[b][grow] + (q - uint32_t(sm.rowstart[rl]))] = sm.glob[xr][key & 0xFFFu];
        }
      }
      // non-zeros grouped by their number of contributions (stable: CSR order inside a run)
      {
        uint32_t keys[kIPT];
        uint16_t vals[kIPT];
        bool too_long = false;
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
          const uint32_t q = tid * kIPT + k;
          keys[k] = 0xFFFFFFFFu, vals[k] = 0xFFFFu;
          if (q < Q) {
            keys[k] = uint32_t(sm.nz_first[q + 1]) - uint32_t(sm.nz_first[q]);
            vals[k] = uint16_t(q);
            too_long = too_long || keys[k] > uint32_t(kMaxLen);
          }
        }
        if (too_long) sm.bad = 1;
        __syncthreads();
        if (sm.bad) break;
        EntSort(sm.tmp.ent).Sort(keys, vals, 0, 6);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
          const uint32_t p = tid * kIPT + k;
          if (p < Q) {
            sm.ord[p] = vals[k];
            sm.Lp[p] = uint8_t(keys[k]);
          }
        }
        if (tid < 64) sm.runstart[tid] = 0xFFFFu;
        __syncthreads();
      }
      for (uint32_t p = tid; p < Q; p += kBT)
        if (p == 0 || sm.Lp[p] != sm.Lp[p - 1]) sm.runstart[sm.Lp[p]] = uint16_t(p);
      __syncthreads();
      // placement of the runs' records in the tile's chunks (closed form of Cursor::place per run)
      if (tid == 0) {
        uint32_t off = sm.cur_off, chunk_start = sm.cur_chunk_start, in_chunk = sm.cur_in_chunk, open = sm.cur_open,
                 next_start = sm.cur_next_start;
        int prevL = -1;
        for (int L = 1; L <= kMaxLen + 1; ++L) {
          const bool present = L <= kMaxLen && sm.runstart[L] != 0xFFFFu;
          if (!present && L <= kMaxLen) continue;
          if (prevL >= 0) {
            RunInfo& R = sm.run[prevL];
            const uint32_t p1 = L <= kMaxLen ? uint32_t(sm.runstart[L]) : Q;
            const uint32_t cnt = p1 - R.p0;
            R.lanes = rec_lanes(uint32_t(prevL));
            R.size = rec_bytes(uint32_t(prevL));
            const uint32_t nrec = (cnt + R.lanes - 1u) / R.lanes;
            R.m = (uint32_t(kChunkBytes) - uint32_t(kChunkHdr)) / R.size;
            R.fits = open ? (chunk_start + uint32_t(kChunkBytes) - off) / R.size : 0u;
            R.off0 = off;
            R.fresh0 = next_start;
            if (nrec <= R.fits) {
              off += nrec * R.size;
              in_chunk += nrec;
            } else {
              if (open && EMIT) *reinterpret_cast<uint32_t*>(sbase + chunk_start) = in_chunk + R.fits;
              const uint32_t rem = nrec - R.fits, nfull = rem / R.m, last = rem % R.m;
              const uint32_t used = nfull + (last ? 1u : 0u);
              if (EMIT)
                for (uint32_t f = 0; f + 1 < used; ++f) *reinterpret_cast<uint32_t*>(sbase + R.fresh0 + f * kChunkBytes) = R.m;
              in_chunk = last ? last : R.m;
              chunk_start = R.fresh0 + (used - 1u) * uint32_t(kChunkBytes);
              off = chunk_start + uint32_t(kChunkHdr) + in_chunk * R.size;
              next_start = chunk_start + uint32_t(kChunkBytes);
              open = 1;
            }
          }
          if (L <= kMaxLen) {
            sm.run[L].p0 = sm.runstart[L];
            prevL = L;
          }
        }
        sm.cur_off = off, sm.cur_chunk_start = chunk_start, sm.cur_in_chunk = in_chunk, sm.cur_open = open,
        sm.cur_next_start = next_start;
      }
      __syncthreads();
      if (EMIT) {
        for (uint32_t p = tid; p < Q; p += kBT) {
          const uint32_t L = sm.Lp[p], q = sm.ord[p];
          const RunInfo& R = sm.run[L];
          const uint32_t idx = p - R.p0, k = idx / R.lanes, lane = idx % R.lanes;
          uint32_t at;
          if (k < R.fits) {
            at = R.off0 + k * R.size;
          } else {
            const uint32_t kk = k - R.fits;
            at = R.fresh0 + (kk / R.m) * uint32_t(kChunkBytes) + uint32_t(kChunkHdr) + (kk % R.m) * R.size;
          }
          unsigned char* rp = sbase + at;
          if (lane == 0) *reinterpret_cast<uint32_t*>(rp) = L | (uint32_t(b) << 8) | (R.lanes << 16);
          const uint32_t first = sm.nz_first[q];
          const uint32_t rl = sm.skey[first] >> 12;
          const uint32_t grow = sm.glob[xt][rl] - B.row_begin;
          reinterpret_cast<uint32_t*>(rp + kRecHdr)[lane] = P.row_ptr[b][grow] + (q - uint32_t(sm.rowstart[rl])) + 2u;
          uint16_t* ent = reinterpret_cast<uint16_t*>(rp + kRecHdr + 4u * R.lanes) + lane;
          for (uint32_t j = 0; j < L; ++j) ent[j * R.lanes] = sm.sval[first + j];
        }
      }
      __syncthreads();
    }
    if (sm.bad) {
      if (tid == 0) atomicExch(P.err, 3);
      if (!EMIT && tid == 0) P.tile_nchunks[t] = 0;
      continue;
    }
    if (tid == 0) {
      if (sm.cur_open && EMIT) *reinterpret_cast<uint32_t*>(sbase + sm.cur_chunk_start) = sm.cur_in_chunk;
      if (!EMIT) P.tile_nchunks[t] = sm.cur_next_start / uint32_t(kChunkBytes);
    }
  }
}

// ------------------------------------------------------------------ generated block sets
#define FQ_DECLARE_SET(fn, n, fk, kind, grade, nin)                                                      \
  struct Set_##fn {                                                                                      \
    static constexpr int kMid = fn##_nmid;                                                               \
    static constexpr int kGroups = fn##_ngroups;                                                         \
    static constexpr int kRows0 = fn##_crows0;                                                           \
    static constexpr int kRows1 = fn##_crows1;                                                           \
    static __host__ __device__ constexpr uint32_t plane(int c) { return c == 0 ? fn##_plane0 : fn##_plane1; }            \
    static __host__ __device__ constexpr uint32_t sb(int b) {                                                            \
      return b == 0 ? fn##_sb0 : (b == 1 ? fn##_sb1 : (b == 2 ? fn##_sb2 : fn##_sb3));                                   \
    }                                                                                                    \
    static __device__ __forceinline__ void a(const double* __restrict__ s, double* __restrict__ mid) {   \
      fn##_a(s, mid);                                                                                    \
    }                                                                                                    \
    template <int G, class S>                                                                            \
    static __device__ __forceinline__ void g(const double* __restrict__ mid, S& sink) {                  \
      if (G == 0) fn##_g0(mid, sink);                                                                    \
      if (G == 1) fn##_g1(mid, sink);                                                                    \
      if (G == 2) fn##_g2(mid, sink);                                                                    \
    }                                                                                                    \
  };
FQ_GEN_SET_LIST(FQ_DECLARE_SET)
#undef FQ_DECLARE_SET

struct TilePlan;
struct FusedLaunch {
  int grid;
  size_t slab_bytes;   // aligned size of the slab: ring and barriers follow
  bool compact;
};
template <class Fn, int NE, int NP, int NC, int PR, int CR>
static void launch_fused_v(fq_ctx* ctx, const FusedLaunch& L, FusedParams params) {
  static bool attr_set = false;
  if (!attr_set) {
    FQ_CUDA(cudaFuncSetAttribute(tile_fused_kernel<Fn, NE, NP, NC, PR, CR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemCta)));
    FQ_CUDA(cudaFuncSetAttribute(tile_fused_kernel<Fn, NE, NP, NC, PR, CR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemCta)));
    attr_set = true;
  }
  params.ring_off = uint32_t(L.slab_bytes);
  params.mbar_off = uint32_t(L.slab_bytes + size_t(NC) * kSlotsPerWarp * kChunkBytes);
  const size_t smem = size_t(params.mbar_off) + kBarBytes;
  if (L.compact)
    tile_fused_kernel<Fn, NE, NP, NC, PR, CR, true><<<L.grid, 32 * (NP + NC), smem, ctx->stream>>>(params);
  else
    tile_fused_kernel<Fn, NE, NP, NC, PR, CR, false><<<L.grid, 32 * (NP + NC), smem, ctx->stream>>>(params);
}
template <class Fn, int NE>
static void launch_fused(fq_ctx* ctx, const FusedLaunch& L, const FusedParams& params) {
  static int variant = -1;
  if (variant < 0) {
    // tuning: producer + consumer warps and registers per producer / consumer thread.  Default 16 + 16 warps (one
    // cell visit per producer thread: one copy of the straight-line tape in the instruction cache), 80 / 48 registers:
    // 3.95 ms at N = 128 against 4.42 ms for 1: 16 + 12 warps at 88 / 48 and 5.5 ms for 2: 8 + 16 warps at 144 / 48
    // (two cell visits per producer thread) — the consumers are bound by shared-memory and TMA latency, so they want warps
    const char* e = std::getenv("FQ_TILE_WARPS");
    variant = e ? std::atoi(e) : 0;
    if (variant < 0 || variant > 2) variant = 0;
  }
  switch (variant) {
    case 1: launch_fused_v<Fn, NE, 16, 12, 88, 48>(ctx, L, params); break;
    case 2: launch_fused_v<Fn, NE, 8, 16, 144, 48>(ctx, L, params); break;
    default: launch_fused_v<Fn, NE, 16, 16, 80, 48>(ctx, L, params); break;
  }
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

struct SetEntryRt {
  int n, fused_k, kind, grade, ninputs, ngroups;
  void (*launch)(fq_ctx*, const FusedLaunch&, const FusedParams&);
};
#define FQ_SET_ENTRY(fn, n, fk, kind, grade, nin) SetEntryRt{n, fk, kind, grade, nin, fn##_ngroups, &launch_fused<Set_##fn, nin>},
static const SetEntryRt g_sets[] = {FQ_GEN_SET_LIST(FQ_SET_ENTRY)};
#undef FQ_SET_ENTRY

// The generated set serving exactly these blocks: hodge_blocks(k) or one single block.
static const SetEntryRt* find_set(int dim, fq_csr* const* csrs, int nblocks, std::vector<BlockSpec>& specs) {
  specs.clear();
  for (int b = 0; b < nblocks; ++b) specs.push_back(BlockSpec{csrs[b]->kind, csrs[b]->grade});
  if (nblocks == 4) {
    for (const SetEntryRt& e : g_sets) {
      if (e.n != dim || e.fused_k < 0) continue;
      const auto hb = hodge_blocks(e.fused_k);
      bool same = true;
      for (int b = 0; b < 4; ++b) same = same && hb[size_t(b)].kind == specs[size_t(b)].kind && hb[size_t(b)].grade == specs[size_t(b)].grade;
      if (same) return &e;
    }
    return nullptr;
  }
  if (nblocks == 1)
    for (const SetEntryRt& e : g_sets)
      if (e.n == dim && e.fused_k < 0 && e.kind == specs[0].kind && e.grade == specs[0].grade) return &e;
  return nullptr;
}

// ------------------------------------------------------------------ plan
struct TilePlan {
  const fq_mesh* mesh = nullptr;
  int dim = 0, nblocks = 0;
  const SetEntryRt* set = nullptr;
  SetDesc desc;
  int gtab[4] = {-1, -1, -1, -1};
  uint32_t ntiles = 0, nchunks = 0;
  uint32_t max_slab = 0;
  size_t slab_bytes = 0;
  DevBuf<uint32_t> tile_cv_ptr, tile_cv_cells;
  DevBuf<TileHdr> tiles;
  DevBuf<uint32_t> cv_rec;
  DevBuf<uint16_t> gbase;
  DevBuf<unsigned char> stream;
  uint32_t plane[kMaxClasses] = {1, 1};
  uint32_t slab_base[kMaxBlocks] = {0, 0, 0, 0};
  DevBuf<int> changed;
  fq_csr* csr[kMaxBlocks] = {nullptr, nullptr, nullptr, nullptr};
  bool compact = false;     // the stream's dests target the value-dependent (dropped) pattern
  bool drop = false;        // semantics the active pattern was produced under
  bool fresh = true;        // no numeric pass yet: dests are structural
  double build_ms = 0.0;
  int grid = 0;
};

static const DimBudget& dim_cost(int dim) {
  static std::mutex mu;
  static std::map<int, DimBudget> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(dim);
  if (it != cache.end()) return it->second;
  return cache.emplace(dim, dim_budget(dim)).first->second;
}

// ---- vertex clustering -------------------------------------------------------
__global__ void vtile_kuhn_kernel(int n, const uint32_t* __restrict__ nv /*[n] vertices per axis*/,
                                  const uint32_t* __restrict__ brick, const uint32_t* __restrict__ nb, uint64_t v_lo,
                                  uint64_t v_hi, uint32_t z_lo, uint32_t* __restrict__ vtile) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t v = v_lo + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; v < v_hi; v += stride) {
    uint64_t rem = v;
    uint32_t t = 0, mul = 1;
    for (int a = 0; a < n; ++a) {
      uint32_t c = uint32_t(rem % nv[a]);
      rem /= nv[a];
      if (a == n - 1) c -= z_lo;
      t += (c / brick[a]) * mul;
      mul *= nb[a];
    }
    vtile[v - v_lo] = t;
  }
}

template <class T>
static void upload_vec(DevBuf<T>& d, const std::vector<T>& h) {
  d.alloc(h.size() ? h.size() : 1);
  // blocks may be recycled by the caching allocator: order this blocking copy after everything in flight
  FQ_CUDA(cudaDeviceSynchronize());
  if (!h.empty()) FQ_CUDA(cudaMemcpy(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
}
static uint32_t read_u32(fq_ctx* ctx, const uint32_t* p) {
  uint32_t v = 0;
  FQ_CUDA(cudaMemcpyAsync(&v, p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return v;
}
// in-place exclusive scan of n + 1 values (the last one is a zero pad); returns the total
static uint32_t exclusive_scan_inplace(fq_ctx* ctx, uint32_t* a, size_t n_plus_1) {
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, a, a, int64_t(n_plus_1), ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, a, a, int64_t(n_plus_1), ctx->stream));
  fq_count_launch(ctx, 2);
  return read_u32(ctx, a + (n_plus_1 - 1));
}

// Closed-form vertex bricks on a Kuhn grid: the largest brick (most owned vertices, then fewest cell visits) whose
// slab, cell visits and per-block entries fit the kernel's limits for every Hodge set of this dimension.
void tile_cluster_kuhn(fq_ctx* ctx, fq_mesh* mesh, int dim, const size_t* shape, size_t slab_begin, size_t slab_end_held) {
  if (dim > 3 || std::getenv("FQ_NO_TILE")) return;
  const DimBudget& dc = dim_cost(dim);
  const double vmax = double(dc.vstar);  // owned vertices of a brick
  std::vector<uint32_t> brick(size_t(dim), 1);
  // exact count for a brick in the interior of the grid: boxes with origin in prod [-1, b_a - 1], dim! chains each
  auto cells_touching = [&](const std::vector<uint32_t>& b) {
    std::vector<int> perm(size_t(dim), 0);
    uint64_t count = 0;
    std::vector<int> o(size_t(dim), -1);
    for (;;) {
      for (int a = 0; a < dim; ++a) perm[size_t(a)] = a;
      do {
        std::vector<int> v = o;
        bool hit = true;
        for (int a = 0; a < dim; ++a) hit = hit && v[size_t(a)] >= 0 && v[size_t(a)] < int(b[size_t(a)]);
        for (int step = 0; step < dim && !hit; ++step) {
          v[size_t(perm[size_t(step)])] += 1;
          bool in = true;
          for (int a = 0; a < dim; ++a) in = in && v[size_t(a)] >= 0 && v[size_t(a)] < int(b[size_t(a)]);
          hit = in;
        }
        if (hit) ++count;
      } while (std::next_permutation(perm.begin(), perm.end()));
      int a = 0;
      while (a < dim && ++o[size_t(a)] >= int(b[size_t(a)])) o[size_t(a++)] = -1;
      if (a == dim) break;
    }
    return count;
  };
  auto fits = [&](const std::vector<uint32_t>& b) {
    double owned = 1;
    for (int a = 0; a < dim; ++a) owned *= b[size_t(a)];
    return owned <= vmax && cells_touching(b) <= uint64_t(kMaxCv);
  };
  if (const char* env = std::getenv("FQ_TILE_BRICK")) {
    int a = 0;
    const char* p = env;
    while (*p && a < dim) {
      brick[size_t(a++)] = uint32_t(std::max(1l, std::strtol(p, const_cast<char**>(&p), 10)));
      if (*p == ',') ++p;
    }
  } else {
    // greedy: grow the axis that keeps the halo ratio smallest while the tile fits
    for (;;) {
      int best = -1;
      double best_ratio = 1e300;
      for (int a = 0; a < dim; ++a) {
        if (brick[size_t(a)] >= shape[a] + 1) continue;
        uint64_t owned = 1;
        std::vector<uint32_t> cand = brick;
        cand[size_t(a)] += 1;
        for (int b = 0; b < dim; ++b) owned *= cand[size_t(b)];
        if (!fits(cand)) continue;
        const double ratio = double(cells_touching(cand)) / double(owned);
        // prefer the lower axis on ties (longer contiguous CSR runs)
        if (ratio < best_ratio - 1e-12) best_ratio = ratio, best = a;
      }
      if (best < 0) break;
      brick[size_t(best)] += 1;
    }
  }
  std::vector<uint32_t> nv(static_cast<size_t>(dim), 0u), nb(static_cast<size_t>(dim), 0u);
  const uint32_t z_lo = uint32_t(slab_begin);
  uint64_t ntiles = 1, layer = 1;
  for (int a = 0; a < dim; ++a) {
    nv[size_t(a)] = uint32_t(shape[a] + 1);
    const uint64_t ext = (a == dim - 1) ? (slab_end_held - slab_begin + 1) : (shape[a] + 1);
    nb[size_t(a)] = uint32_t((ext + brick[size_t(a)] - 1) / brick[size_t(a)]);
    ntiles *= nb[size_t(a)];
    if (a < dim - 1) layer *= shape[a] + 1;
  }
  FQ_REQUIRE(ntiles < (1ull << 31), "too many tiles");
  const uint64_t v_lo = layer * slab_begin, v_hi = layer * (slab_end_held + 1);
  DevBuf<uint32_t> d_nv, d_brick, d_nb;
  upload_vec(d_nv, nv);
  upload_vec(d_brick, brick);
  upload_vec(d_nb, nb);
  mesh->vertex_tile.alloc(size_t(v_hi - v_lo));
  mesh->vtile_lo = size_t(v_lo);
  mesh->ntiles = size_t(ntiles);
  vtile_kuhn_kernel<<<grid_for(v_hi - v_lo, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
      dim, d_nv.p, d_brick.p, d_nb.p, v_lo, v_hi, z_lo, mesh->vertex_tile.p);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
}

// Generic (uploaded) meshes: vertices are cut into tiles along their numbering, on the device.  Every vertex gets a
// weight = its share of the tightest tile budget (row slots per grade are additive over vertices; the cell visits are
// estimated as half the incident cells, neighbouring vertices sharing about that many); a tile is a run of
// consecutive vertices of total weight ~1/scale.  Mesh generators number vertices with locality (Kuhn grids: x-lines),
// so the runs are compact; the plan builder verifies the budgets exactly and the caller retries with a larger scale
// (smaller tiles) when one is exceeded.
__global__ void vertex_weight_kernel(const uint32_t* __restrict__ cell_verts, int nv, size_t ncells, size_t V,
                                     uint32_t* __restrict__ rs /*[nv][V]*/, uint32_t* __restrict__ deg, int* __restrict__ err) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t c = size_t(blockIdx.x) * blockDim.x + threadIdx.x; c < ncells; c += stride)
    for (int j = 0; j < nv; ++j) {
      const uint32_t v = cell_verts[c * nv + j];
      if (v >= V) {
        *err = 1;
        continue;
      }
      atomicAdd(&deg[v], 1u);
      // faces of grade g with this vertex on top: C(j, g)
      uint32_t bin = 1;
      for (int g = 0; g < nv; ++g) {
        if (bin) atomicAdd(&rs[size_t(g) * V + v], bin);
        bin = bin * uint32_t(j - g) / uint32_t(g + 1);
      }
    }
}
struct VertexCostArgs {
  float inv_budget[4];
  float inv_cv;
  float scale;
  int ngrades;
};
__global__ void vertex_cost_kernel(size_t V, const uint32_t* __restrict__ rs, const uint32_t* __restrict__ deg, VertexCostArgs A,
                                   unsigned long long* __restrict__ w) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t v = size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += stride) {
    float cost = float(deg[v]) * A.inv_cv;
    for (int g = 0; g < A.ngrades; ++g) cost = fmaxf(cost, float(rs[size_t(g) * V + v]) * A.inv_budget[g]);
    w[v] = (unsigned long long)(double(cost * A.scale) * 1048576.0) + 1ull;  // fixed point: the scan is exact
  }
}
__global__ void vertex_tile_from_prefix_kernel(size_t V, const unsigned long long* __restrict__ prefix /*exclusive*/,
                                               uint32_t* __restrict__ vtile) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t v = size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += stride) vtile[v] = uint32_t(prefix[v] >> 20);
}
void tile_cluster_generic(fq_ctx* ctx, fq_mesh* mesh) {
  const int dim = mesh->dim;
  mesh->cluster_tried = true;
  if (dim > 3 || std::getenv("FQ_NO_TILE") || !mesh->cell_faces[0].p || mesh->edge_lo != 0 || mesh->cell_offset != 0) return;
  ScopedSpan span(ctx, "tp_cluster");
  const DimBudget& dc = dim_cost(dim);
  const int nv = dim + 1, block = 256;
  const size_t ncells = mesh->ncells, V = mesh->nsimplices[0];
  if (V == 0 || ncells == 0 || V >= (size_t(1) << 32)) return;
  if (mesh->cluster_scale <= 0.0f) mesh->cluster_scale = 1.1f;
  DevBuf<uint32_t> rs(size_t(nv) * V), deg(V);
  DevBuf<unsigned long long> w(V + 1), prefix(V + 1);
  DevBuf<int> d_err(1);
  FQ_CUDA(cudaMemsetAsync(rs.p, 0, rs.bytes(), ctx->stream));
  FQ_CUDA(cudaMemsetAsync(deg.p, 0, deg.bytes(), ctx->stream));
  FQ_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), ctx->stream));
  vertex_weight_kernel<<<grid_for(ncells, block, ctx->sm_count), block, 0, ctx->stream>>>(mesh->cell_faces[0].p, nv, ncells, V, rs.p,
                                                                                         deg.p, d_err.p);
  VertexCostArgs A{};
  A.ngrades = nv;
  for (int g = 0; g < nv; ++g) A.inv_budget[g] = dc.budget[g] ? 1.0f / float(dc.budget[g]) : 0.0f;
  A.inv_cv = 1.0f / (2.0f * float(kMaxCv));
  A.scale = mesh->cluster_scale;
  vertex_cost_kernel<<<grid_for(V, block, ctx->sm_count), block, 0, ctx->stream>>>(V, rs.p, deg.p, A, w.p);
  FQ_CUDA(cudaMemsetAsync(w.p + V, 0, sizeof(unsigned long long), ctx->stream));
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, w.p, prefix.p, int64_t(V + 1), ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, w.p, prefix.p, int64_t(V + 1), ctx->stream));
  mesh->vertex_tile.alloc(V);
  vertex_tile_from_prefix_kernel<<<grid_for(V, block, ctx->sm_count), block, 0, ctx->stream>>>(V, prefix.p, mesh->vertex_tile.p);
  fq_count_launch(ctx, 5);
  FQ_CUDA(cudaGetLastError());
  unsigned long long total = 0;
  int h_err = 0;
  FQ_CUDA(cudaMemcpyAsync(&total, prefix.p + V, sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaMemcpyAsync(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h_err || (total >> 20) >= (1ull << 31)) {  // malformed table: leave the mesh unclustered (slab path)
    mesh->vertex_tile.release();
    mesh->ntiles = 0;
    return;
  }
  mesh->vtile_lo = 0;
  mesh->ntiles = size_t(total >> 20) + 1;
  mesh->cluster_generic = true;
}

// ---- cell visits of every tile ------------------------------------------------
// one key per (cell, distinct vertex tile): (tile << 32) | cell, ~0 for duplicates
__global__ void tile_cell_keys_kernel(const uint32_t* __restrict__ cell_verts, int nv, size_t ncells,
                                      const uint32_t* __restrict__ vtile, uint32_t v_lo, uint64_t* __restrict__ keys) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t c = size_t(blockIdx.x) * blockDim.x + threadIdx.x; c < ncells; c += stride) {
    uint32_t seen[16];
    for (int j = 0; j < nv; ++j) {
      const uint32_t t = vtile[cell_verts[c * nv + j] - v_lo];
      bool dup = false;
      for (int i = 0; i < j; ++i) dup = dup || (seen[i] == t);
      seen[j] = t;
      keys[c * nv + j] = dup ? ~0ull : ((uint64_t(t) << 32) | uint64_t(c));
    }
  }
}
struct IsValidKey64 {
  __device__ __forceinline__ uint32_t operator()(const uint64_t& k) const { return k != ~0ull ? 1u : 0u; }
};
__global__ void split_keys_kernel(const uint64_t* __restrict__ keys, size_t n, uint32_t* __restrict__ tile,
                                  uint32_t* __restrict__ cell) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    tile[i] = uint32_t(keys[i] >> 32);
    cell[i] = uint32_t(keys[i]);
  }
}
// ptr[t] = first i with key[i] >= t, for sorted keys; ptr[nseg] = n
__global__ void seg_ptr_kernel(const uint32_t* __restrict__ key, size_t n, uint32_t nseg, uint32_t* __restrict__ ptr) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i <= n; i += stride) {
    const uint32_t hi = (i == n) ? nseg : key[i];
    const uint32_t lo = (i == 0) ? 0u : key[i - 1] + 1;
    for (uint32_t t = lo; t <= hi && t <= nseg; ++t) ptr[t] = uint32_t(i);
  }
}
static int bits_for32(uint64_t n) {
  int b = 1;
  while ((1ull << b) < n) ++b;
  return b;
}
static void build_cell_visits(fq_ctx* ctx, const fq_mesh* mesh, TilePlan& plan) {
  ScopedSpan span(ctx, "tp_cell_visits");
  const int nv = mesh->dim + 1, block = 256;
  const size_t ncells = mesh->ncells, nkeys = ncells * size_t(nv);
  DevBuf<uint64_t> keys(nkeys ? nkeys : 1), keys_alt(nkeys ? nkeys : 1);
  tile_cell_keys_kernel<<<grid_for(ncells, block, ctx->sm_count), block, 0, ctx->stream>>>(
      mesh->cell_faces[0].p, nv, ncells, mesh->vertex_tile.p, uint32_t(mesh->vtile_lo), keys.p);
  fq_count_launch(ctx);
  cub::DoubleBuffer<uint64_t> dk(keys.p, keys_alt.p);
  size_t tmp_bytes = 0;
  // the all-ones keys of duplicates sort last on any bit range that covers the valid keys
  const int end_bit = std::min(64, 32 + bits_for32(uint64_t(plan.ntiles) + 1));
  FQ_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, dk, int64_t(nkeys), 0, end_bit, ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, dk, int64_t(nkeys), 0, end_bit, ctx->stream));
  fq_count_launch(ctx, 9);
  cub::TransformInputIterator<uint32_t, IsValidKey64, const uint64_t*> it(dk.Current(), IsValidKey64());
  DevBuf<uint32_t> d_count(1);
  size_t tmp3 = 0;
  FQ_CUDA(cub::DeviceReduce::Sum(nullptr, tmp3, it, d_count.p, int64_t(nkeys), ctx->stream));
  if (tmp.n < tmp3) tmp.alloc(tmp3);
  FQ_CUDA(cub::DeviceReduce::Sum(tmp.p, tmp3, it, d_count.p, int64_t(nkeys), ctx->stream));
  fq_count_launch(ctx);
  const uint32_t nvalid = read_u32(ctx, d_count.p);
  DevBuf<uint32_t> tile_of(nvalid ? nvalid : 1);
  plan.tile_cv_cells.alloc(nvalid ? nvalid : 1);
  split_keys_kernel<<<grid_for(nvalid, block, ctx->sm_count), block, 0, ctx->stream>>>(dk.Current(), nvalid, tile_of.p,
                                                                                     plan.tile_cv_cells.p);
  plan.tile_cv_ptr.alloc(size_t(plan.ntiles) + 1);
  seg_ptr_kernel<<<grid_for(size_t(nvalid) + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(tile_of.p, nvalid, plan.ntiles,
                                                                                              plan.tile_cv_ptr.p);
  fq_count_launch(ctx, 2);
  FQ_CUDA(cudaGetLastError());
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
}

// Host reference builder on downloaded tables (FQ_TILE_BUILD=host; development / validation of the device builder).
static bool build_streams_host(fq_ctx* ctx, const fq_mesh* mesh, TilePlan& plan) {
  const SetDesc& S = plan.desc;
  HostMesh M;
  std::vector<std::vector<uint32_t>> faces(4);
  for (int g = 0; g <= mesh->dim; ++g) {
    if (!mesh->cell_faces[size_t(g)].p) continue;
    faces[size_t(g)].resize(mesh->cell_faces[size_t(g)].n);
    FQ_CUDA(cudaMemcpy(faces[size_t(g)].data(), mesh->cell_faces[size_t(g)].p, faces[size_t(g)].size() * 4, cudaMemcpyDeviceToHost));
    M.faces[g] = faces[size_t(g)].data();
  }
  std::vector<uint32_t> vtile(mesh->vertex_tile.n), cvp(plan.tile_cv_ptr.n), cvc(plan.tile_cv_cells.n);
  FQ_CUDA(cudaMemcpy(vtile.data(), mesh->vertex_tile.p, vtile.size() * 4, cudaMemcpyDeviceToHost));
  FQ_CUDA(cudaMemcpy(cvp.data(), plan.tile_cv_ptr.p, cvp.size() * 4, cudaMemcpyDeviceToHost));
  FQ_CUDA(cudaMemcpy(cvc.data(), plan.tile_cv_cells.p, cvc.size() * 4, cudaMemcpyDeviceToHost));
  M.ncells = mesh->ncells;
  M.vertex_tile = vtile.data();
  M.v_lo = uint32_t(mesh->vtile_lo);
  M.ntiles = plan.ntiles;
  M.tile_cv_ptr = cvp.data();
  M.tile_cv_cells = cvc.data();
  HostPlan H;
  try {
    HostBuilder builder(S, M);
    H = builder.build(kSlabCapacity);
  } catch (const std::runtime_error&) {
    return false;
  }
  plan.max_slab = H.max_slab;
  plan.nchunks = uint32_t(H.stream.size() / kChunkBytes);
  upload_vec(plan.tiles, H.tiles);
  upload_vec(plan.cv_rec, H.cv_rec);
  upload_vec(plan.gbase, H.gbase);
  upload_vec(plan.stream, H.stream);
  for (int b = 0; b < plan.nblocks; ++b) {
    fq_csr* csr = plan.csr[b];
    if (S.blk[b].empty) continue;
    upload_vec(csr->s_row_ptr, H.row_ptr[b]);
    upload_vec(csr->s_col_idx, H.col_idx[b]);
    csr->s_nnz = H.col_idx[b].size();
    csr->ncontrib = size_t(H.nentries[b]);
  }
  return true;
}

static bool build_streams_device(fq_ctx* ctx, const fq_mesh* mesh, TilePlan& plan) {
  const SetDesc& S = plan.desc;
  BuildParams P{};
  P.S = S;
  for (int g = 0; g < 4; ++g) {
    P.gtab[g] = plan.gtab[g];
    P.id_bits[g] = g <= mesh->dim ? std::min(32, bits_for32(uint64_t(mesh->nsimplices[size_t(g)]) + 1)) : 1;
    P.faces[g] = (g <= mesh->dim) ? mesh->cell_faces[size_t(g)].p : nullptr;
  }
  P.vertex_tile = mesh->vertex_tile.p;
  P.v_lo = uint32_t(mesh->vtile_lo);
  P.ntiles = plan.ntiles;
  P.tile_cv_ptr = plan.tile_cv_ptr.p;
  P.tile_cv_cells = plan.tile_cv_cells.p;
  for (int c = 0; c < kMaxClasses; ++c) P.plane[c] = plan.plane[c];
  for (int b = 0; b < kMaxBlocks; ++b) P.slab_base[b] = plan.slab_base[b];
  DevBuf<int> d_err(1);
  DevBuf<uint32_t> d_max(kMaxClasses), tile_chunks(size_t(plan.ntiles) + 1);
  FQ_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), ctx->stream));
  FQ_CUDA(cudaMemsetAsync(d_max.p, 0, d_max.bytes(), ctx->stream));
  FQ_CUDA(cudaMemsetAsync(tile_chunks.p, 0, tile_chunks.bytes(), ctx->stream));
  for (int b = 0; b < plan.nblocks; ++b) {
    fq_csr* csr = plan.csr[b];
    if (S.blk[b].empty) continue;
    const size_t nrows_local = csr->row_end - csr->row_begin;
    csr->s_row_ptr.alloc(nrows_local + 1);
    FQ_CUDA(cudaMemsetAsync(csr->s_row_ptr.p, 0, csr->s_row_ptr.bytes(), ctx->stream));
    P.row_ptr[b] = csr->s_row_ptr.p;
  }
  P.tile_nchunks = tile_chunks.p;
  DevBuf<unsigned long long> d_ncontrib(kMaxBlocks);
  FQ_CUDA(cudaMemsetAsync(d_ncontrib.p, 0, d_ncontrib.bytes(), ctx->stream));
  P.ncontrib = d_ncontrib.p;
  P.err = d_err.p;
  P.rs_max = d_max.p;
  static bool attr_set = false;
  if (!attr_set) {
    FQ_CUDA(cudaFuncSetAttribute(tile_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(BuildSmem))));
    FQ_CUDA(cudaFuncSetAttribute(tile_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(BuildSmem))));
    attr_set = true;
  }
  const int grid = int(std::min<uint64_t>(plan.ntiles, uint64_t(ctx->sm_count)));
  {
    ScopedSpan span(ctx, "tp_count");
    tile_build_kernel<false><<<grid, kBT, sizeof(BuildSmem), ctx->stream>>>(P);
  }
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
  int h_err = 0;
  FQ_CUDA(cudaMemcpyAsync(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h_err) return false;
  {
    uint32_t rs_max[kMaxClasses];
    FQ_CUDA(cudaMemcpy(rs_max, d_max.p, sizeof rs_max, cudaMemcpyDeviceToHost));
    for (int c = 0; c < S.nclasses; ++c)
      if (rs_max[c] > plan.plane[c]) return false;  // a tile exceeds its row-slot budget: slab path
  }
  unsigned long long h_ncontrib[kMaxBlocks];
  FQ_CUDA(cudaMemcpy(h_ncontrib, d_ncontrib.p, sizeof h_ncontrib, cudaMemcpyDeviceToHost));
  for (int b = 0; b < plan.nblocks; ++b) {
    fq_csr* csr = plan.csr[b];
    if (S.blk[b].empty) continue;
    csr->ncontrib = size_t(h_ncontrib[b]);
    const size_t nrows_local = csr->row_end - csr->row_begin;
    csr->s_nnz = exclusive_scan_inplace(ctx, csr->s_row_ptr.p, nrows_local + 1);
    csr->s_col_idx.alloc(csr->s_nnz ? csr->s_nnz : 1);
    P.col_idx[b] = csr->s_col_idx.p;
  }
  plan.nchunks = exclusive_scan_inplace(ctx, tile_chunks.p, size_t(plan.ntiles) + 1);
  P.tile_chunk_ptr = tile_chunks.p;
  plan.tiles.alloc(plan.ntiles ? plan.ntiles : 1);
  plan.cv_rec.alloc(std::max<size_t>(1, plan.tile_cv_cells.n * size_t(S.cv_words)));
  plan.gbase.alloc(((plan.tile_cv_cells.n >> 5) + size_t(plan.ntiles) + 1) * kGroupWords);
  P.gbase = plan.gbase.p;
  plan.stream.alloc(size_t(plan.nchunks ? plan.nchunks : 1) * kChunkBytes);
  FQ_CUDA(cudaMemsetAsync(plan.stream.p, 0, plan.stream.bytes(), ctx->stream));  // padding lanes: dest 0, codes 0
  P.tiles = plan.tiles.p;
  P.cv_rec = plan.cv_rec.p;
  P.stream = plan.stream.p;
  {
    ScopedSpan span(ctx, "tp_emit");
    tile_build_kernel<true><<<grid, kBT, sizeof(BuildSmem), ctx->stream>>>(P);
  }
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
  FQ_CUDA(cudaMemcpyAsync(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return h_err == 0;
}

// Builds the plan of the fused blocks `csrs` and, with it, their structural patterns (s_row_ptr, s_col_idx, s_nnz).
// Returns nullptr when the tile path does not apply (the caller falls back to the slab path).
std::shared_ptr<TilePlan> tile_plan_build(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* const* csrs, int nblocks) {
  if (std::getenv("FQ_NO_TILE")) return nullptr;
  if (nblocks < 1 || nblocks > kMaxBlocks) return nullptr;
  if (!mesh->vertex_tile.p || mesh->ntiles == 0 || !mesh->cell_faces[0].p) return nullptr;
  const int dim = mesh->dim;
  if (dim > 3 || dim < 1) return nullptr;
  std::vector<BlockSpec> specs;
  const SetEntryRt* set = find_set(dim, csrs, nblocks, specs);
  if (!set) return nullptr;
  cudaEvent_t ev0, ev1;
  FQ_CUDA(cudaEventCreate(&ev0));
  FQ_CUDA(cudaEventCreate(&ev1));
  FQ_CUDA(cudaEventRecord(ev0, ctx->stream));
  auto plan = std::make_shared<TilePlan>();
  plan->mesh = mesh;
  plan->dim = dim;
  plan->nblocks = nblocks;
  plan->set = set;
  plan->ntiles = uint32_t(mesh->ntiles);
  plan->desc = make_set(dim, specs);
  SetDesc& S = plan->desc;
  int ntab = 0;
  for (int b = 0; b < nblocks; ++b) {
    plan->csr[b] = csrs[b];
    BlockDesc& B = S.blk[b];
    if (B.empty) continue;
    if (csrs[b]->row_end > 0xFFFFFFFEull) return nullptr;
    B.row_begin = uint32_t(csrs[b]->row_begin);
    B.row_end = uint32_t(csrs[b]->row_end);
    for (int g : {B.tg, B.rg}) {
      if (!mesh->cell_faces[size_t(g)].p) return nullptr;
      if (plan->gtab[g] < 0) {
        if (ntab == 2) return nullptr;
        plan->gtab[g] = ntab++;
      }
    }
  }
  if (!mesh->cell_faces[1].p) return nullptr;
  if (!finish_classes(S)) return nullptr;  // blocks of one test grade with different row ranges: slab path
  set_slab_layout(S, dim_cost(dim), plan->plane, plan->slab_base, plan->max_slab);
  if (plan->max_slab > kSlabCapacity || plan->max_slab > 0x10000u) return nullptr;
  build_cell_visits(ctx, mesh, *plan);
  const char* how = std::getenv("FQ_TILE_BUILD");
  const bool ok = (how && how[0] == 'h') ? build_streams_host(ctx, mesh, *plan) : build_streams_device(ctx, mesh, *plan);
  if (!ok) return nullptr;
  // dynamic shared memory of the fused kernel: slab | TMA ring | mbarriers (laid out at launch)
  plan->slab_bytes = (size_t(plan->max_slab) * sizeof(double) + 127) / 128 * 128;
  if (plan->slab_bytes + kRingBytesMax + kBarBytes > kSmemCta) return nullptr;
  plan->changed.alloc(1);
  plan->grid = int(std::min<uint64_t>(uint64_t(ctx->sm_count), std::max<uint32_t>(plan->ntiles, 1u)));
  FQ_CUDA(cudaEventRecord(ev1, ctx->stream));
  FQ_CUDA(cudaEventSynchronize(ev1));
  float ms = 0;
  FQ_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  plan->build_ms = ms;
  FQ_CUDA(cudaEventDestroy(ev0));
  FQ_CUDA(cudaEventDestroy(ev1));
  return plan;
}

// One launch of the fused kernel over the plan's blocks.  structural: dest = structural position, `keep` flags written
// (when the block has a keep array); compact: dest = position in the value-dependent pattern, classification verified.
// Returns false when the classification differs from the plan's.
bool tile_assemble(fq_ctx* ctx, const fq_mesh* mesh, TilePlan& plan, double* const* values, uint8_t* const* keep) {
  FusedParams P{};
  P.tiles = plan.tiles.p;
  P.ntiles = plan.ntiles;
  P.cv_rec = plan.cv_rec.p;
  P.cv_words = plan.desc.cv_words;
  P.gbase = plan.gbase.p;
  P.lengths = mesh->lengths.p;
  P.edge_lo = uint32_t(mesh->edge_lo);
  P.stream = plan.stream.p;
  P.changed = plan.changed.p;
  P.chunk_rotation = std::getenv("FQ_TILE_ROTATE") ? uint32_t(std::atoi(std::getenv("FQ_TILE_ROTATE"))) : 5u;
  P.debug = std::getenv("FQ_TILE_DEBUG") ? std::atoi(std::getenv("FQ_TILE_DEBUG")) : 0;
  for (int b = 0; b < kMaxBlocks; ++b) {
    P.values[b] = b < plan.nblocks ? values[b] : nullptr;
    P.keep[b] = (b < plan.nblocks && keep) ? keep[b] : nullptr;
    const bool live = b < plan.nblocks && !plan.desc.blk[b].empty;
    P.group_pack |= uint32_t(live ? plan.desc.blk[b].group : 0) << (8 * b);
  }
  {
    ScopedSpan span(ctx, "k13_tile_fused");
    FQ_CUDA(cudaMemsetAsync(plan.changed.p, 0, sizeof(int), ctx->stream));
    plan.set->launch(ctx, FusedLaunch{plan.grid, plan.slab_bytes, plan.compact}, P);
  }
  if (!plan.compact) return true;
  int changed = 0;
  FQ_CUDA(cudaMemcpyAsync(&changed, plan.changed.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return changed == 0;
}

// dest of every record lane: structural position -> position in the compacted pattern (keep / pos of every block)
void tile_retarget(fq_ctx* ctx, TilePlan& plan) {
  RetargetArgs A{};
  for (int b = 0; b < plan.nblocks; ++b) {
    A.keep[b] = plan.csr[b]->keep.p;
    A.pos[b] = plan.csr[b]->pos.p;
  }
  ScopedSpan span(ctx, "tile_retarget");
  retarget_kernel<<<grid_for(size_t(plan.nchunks) * 32, 256, ctx->sm_count), 256, 0, ctx->stream>>>(plan.stream.p, plan.nchunks, A);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
  plan.compact = true;
}

bool tile_plan_matches(const TilePlan& plan, const fq_mesh* mesh, fq_csr* const* csrs, int nblocks) {
  if (plan.mesh != mesh || plan.nblocks != nblocks) return false;
  for (int b = 0; b < nblocks; ++b)
    if (plan.csr[b] != csrs[b]) return false;
  return true;
}
bool& tile_plan_compact(TilePlan& plan) { return plan.compact; }
bool& tile_plan_drop(TilePlan& plan) { return plan.drop; }
double tile_plan_build_ms(const TilePlan& plan) { return plan.build_ms; }
size_t tile_plan_cell_visits(const TilePlan& plan) { return plan.tile_cv_cells.n; }

int64_t tile_plan_bytes(const TilePlan& plan) {
  return int64_t(plan.tiles.bytes() + plan.cv_rec.bytes() + plan.gbase.bytes() + plan.stream.bytes());
}
// bytes of the per-step inputs of the fused kernel other than the edge lengths: cell-visit records + record streams
int64_t tile_plan_stream_bytes(const TilePlan& plan) { return int64_t(plan.stream.bytes()); }
int64_t tile_plan_cv_bytes(const TilePlan& plan) { return int64_t(plan.cv_rec.bytes()); }

}  // namespace fq
