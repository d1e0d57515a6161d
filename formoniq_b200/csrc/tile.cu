// tile.cu — tile-fused numeric assembly: K1 (element masses) and K3 (segmented
// reduction into CSR) in ONE persistent kernel, with the element data of a tile
// living only in shared memory.  The element slab of the two-kernel path
// (8*T bytes per cell written and read back through HBM) disappears.
//
// Decomposition = the multi-GPU one, repeated at CTA level: *owner computes*.
//   * vertices are clustered into tiles (closed-form bricks on Kuhn grids,
//     breadth-first clusters on generic meshes);
//   * a tile owns the rows (simplices) whose top vertex it contains, hence
//     whole CSR rows, hence every structural non-zero of those rows;
//   * it evaluates the element masses of ALL cells touching its vertices
//     (owned + halo cells, recomputed by the neighbouring tiles — FP64 work is
//     cheap here, HBM traffic is not) into shared memory, then reduces each of
//     its non-zeros over the contributing (cell, slot) pairs in ascending cell
//     order — the same order as the slab path, so values are bit-identical.
//
// Kernels (same plan, same streams): tile_assemble_alt_kernel (default) runs 8 producer warps (K1) and 16 consumer warps
// (K3) over ONE slab used in two alternating halves, so the FP64 work hides completely behind the gather without
// shrinking the tiles; tile_assemble_kernel (FQ_TILE_KERNEL=s) is the phase-serialised predecessor (K1, barrier, gather,
// barrier on 16 warps); tile_assemble_ws_kernel (=w) the two-slab producer/consumer experiment.
//
// Shared memory holds only the DISTINCT values a cell contributes: for
// HodgeBlocks the masses M_{k-1}, M_k, M_{k+1} (54 doubles for 3-D k = 1 instead of 112
// element entries; FQ_TILE_CORE=h stores dif_both(k+1) instead of M_{k+1}: 74).  dif_test = d*M_k and dif_both
// (operators.rs:201-211) are evaluated by the gather through per-slot "recipes" —
// signed sums of mass entries in the reference's k-ascending gemm order, exact
// because the incidence entries are 0/+-1 (tape.hpp evaluates the same products
// symbolically).
//
// The cell-slot -> nnz map is laid out per tile as ONE contiguous byte stream
// of 2 KB chunks holding warp-sized records {header; dest[64]; entry[L][64]}
// (non-zeros grouped by their number of contributions L, so a warp runs L
// uniform iterations on two independent chains per lane, lanes read
// consecutive 2-byte entries, no per-nnz offsets are stored).  Chunk c of a
// tile belongs to warp c mod NW: every warp streams its own chunks through a
// private double buffer filled by TMA bulk copies (cp.async.bulk + mbarrier),
// prefetching across tile boundaries — no warp ever waits on another warp or
// on a dependent global load during the gather, and HBM sees long sequential
// reads.
//
// Reference path replaced: formoniq/src/galerkin.rs:138-188 (assemble_matrix)
// + hodge.rs:62-72 (the four HodgeBlocks), numeric phase.
#include <cub/cub.cuh>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "elmat_gen.cuh"
#include "internal.hpp"
#include "kuhn.hpp"

namespace fq {

constexpr int kTileMaxBlocks = 4;
// dest codes of the stream: 0 = padding lane, 1 = dropped non-zero (must stay all-zero, galerkin.rs:173),
// d >= 2 = position d - 2 of csr->values
constexpr uint32_t kPadDest = 0u;
constexpr uint32_t kNoDest = 1u;
#ifndef FQ_TILE_CHUNK_BYTES
#define FQ_TILE_CHUNK_BYTES 1536
#endif
constexpr int kChunkBytes = FQ_TILE_CHUNK_BYTES;  // TMA granule of the tile stream; records never straddle a chunk
constexpr int kChunkHdr = 16;       // u32 nrec + padding
constexpr int kRecHdr = 16;         // u32 (L | block << 8 | lanes << 16) + padding
// a record is 16 + lanes * (4 + 2 L) bytes and must fit a chunk after its 16-byte header
constexpr int kMaxLen64 = ((kChunkBytes - 32) / 64 - 4) / 2;   // 2 KB chunks: 64 lanes up to L = 13
constexpr int kMaxLen32 = ((kChunkBytes - 32) / 32 - 4) / 2;   //              32 lanes up to L = 29
constexpr int kMaxLen = ((kChunkBytes - 32) / 16 - 4) / 2;     //              16 lanes up to L = 61
constexpr int kSlotsPerWarp = 2;    // private double buffer of every warp

__host__ __device__ inline uint32_t rec_lanes(uint32_t L) { return L <= uint32_t(kMaxLen64) ? 64u : (L <= uint32_t(kMaxLen32) ? 32u : 16u); }
__host__ __device__ inline uint32_t rec_bytes(uint32_t L) { return uint32_t(kRecHdr) + rec_lanes(L) * (4u + 2u * L); }

struct TileBlockDev {
  double* values;
  int no, ni;      // recipe shape: outer x inner signed terms per slot (1 x 1: entries are pre-translated)
  int recipe_off;  // offset (u16 units) of this block's recipes
  int slot_bits;
};

struct TileParams {
  const uint32_t* tile_cell_ptr;    // [ntiles+1]
  const uint32_t* tile_cell_edges;  // [tile cell slots][NE] edge ids, pre-gathered
  const double* lengths;
  uint32_t edge_lo;
  uint32_t ntiles;
  const uint32_t* tile_chunk_ptr;   // [ntiles+1] chunk index of the tile's stream
  const unsigned char* stream;      // chunks of kChunkBytes
  int cstride;                      // cells capacity of the shared slab
  int nblocks;
  uint32_t ring_off, rec_off, mbar_off;  // byte offsets in dynamic shared memory
  uint32_t slab_bytes;                   // warp-specialised kernel: size of one of its two slabs
  const uint8_t* recipes;           // u16 codes: sign << 15 | distinct * cstride
  int recipe_bytes;
  int debug;                        // development knobs (FQ_TILE_DEBUG): 1 skip K1, 2 skip records, 4 skip stores
  int check_classification;         // 1 when the plan carries the reference's value-dependent pattern
  uint32_t yblock_mask;             // alternating kernel: blocks whose records read second-half slab values only
  uint32_t chunk_rotation;          // alternating kernel: per-tile rotation of the chunk -> warp deal (0: none)
  int* changed;                     // raised when the zero/non-zero classification differs from the plan's
  unsigned int* ticket;             // dynamic tile scheduler
  unsigned long long* stats;        // debug & 8: per-phase warp cycles [k1, bar_k1, chunk_wait, records, bar_top, other]
  TileBlockDev blk[kTileMaxBlocks];
};

struct TileSink {
  double* __restrict__ slab;  // + local cell
  int cstride;
  template <int B, int E>
  __device__ __forceinline__ void put(double v) const {
    slab[E * cstride] = v;
  }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
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

// x with the sign bit flipped when bit 15 of `code` is set (exact negation)
__device__ __forceinline__ double signed_load(const double* __restrict__ p, uint32_t code) {
  const double x = p[code & 0x7FFFu];
  return __hiloint2double(__double2hiint(x) ^ int((code & 0x8000u) << 16), __double2loint(x));
}

// Value of one contribution: a recipe of NO x NI signed stored entries
//   v = (((x00 + x01) + ..) + ((x10 + x11) + ..)) + ..
// (the k-ascending gemm order of operators.rs:201-211 with the +-1 incidence entries folded in).
template <int NO, int NI>
__device__ __forceinline__ double recipe_value(uint32_t e, const double* __restrict__ slab, const uint16_t* __restrict__ brec,
                                               uint32_t sb, uint32_t slot_mask) {
  constexpr int NT4 = (NO * NI + 3) / 4 * 4;  // codes per slot, padded to 8-byte groups
  const double* __restrict__ sc = slab + (e >> sb);
  const uint2* __restrict__ rr = reinterpret_cast<const uint2*>(brec + (e & slot_mask) * NT4);
  uint32_t code[NT4];
#pragma unroll
  for (int w = 0; w < NT4 / 4; ++w) {
    const uint2 c = rr[w];
    code[4 * w + 0] = c.x & 0xFFFFu;
    code[4 * w + 1] = c.x >> 16;
    code[4 * w + 2] = c.y & 0xFFFFu;
    code[4 * w + 3] = c.y >> 16;
  }
  double v = 0.0;
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    double inner = signed_load(sc, code[o * NI]);
#pragma unroll
    for (int q = 1; q < NI; ++q) inner = __dadd_rn(inner, signed_load(sc, code[o * NI + q]));
    v = (o == 0) ? inner : __dadd_rn(v, inner);
  }
  return v;
}
// The two non-zeros of a lane (columns lane and lane + 32 of a wide record; a narrow record runs the
// second chain on the first column and discards it): left-to-right sums over L contributions.
template <int NO, int NI>
__device__ __forceinline__ void gather_record(const uint16_t* __restrict__ ent0, const uint16_t* __restrict__ ent1,
                                              uint32_t stride, uint32_t L, const double* __restrict__ slab,
                                              const uint16_t* __restrict__ brec, uint32_t sb, uint32_t slot_mask, double& acc0,
                                              double& acc1, bool& any0, bool& any1) {
#pragma unroll 1
  for (uint32_t j = 0; j < L; ++j) {
    const uint32_t e0 = ent0[j * stride], e1 = ent1[j * stride];
    const double v0 = recipe_value<NO, NI>(e0, slab, brec, sb, slot_mask);
    const double v1 = recipe_value<NO, NI>(e1, slab, brec, sb, slot_mask);
    any0 = any0 || (v0 != 0.0);
    any1 = any1 || (v1 != 0.0);
    acc0 = __dadd_rn(acc0, v0);
    acc1 = __dadd_rn(acc1, v1);
  }
}
// blocks stored directly: the stream entries are pre-translated to sign | slab offset
__device__ __forceinline__ void gather_record_direct(const uint16_t* __restrict__ ent0, const uint16_t* __restrict__ ent1,
                                                     uint32_t stride, uint32_t L, const double* __restrict__ slab,
                                                     double& acc0, double& acc1, bool& any0, bool& any1) {
#pragma unroll 1
  for (uint32_t j = 0; j < L; ++j) {
    const double x0 = signed_load(slab, ent0[j * stride]);
    const double x1 = signed_load(slab, ent1[j * stride]);
    any0 = any0 || (x0 != 0.0);
    any1 = any1 || (x1 != 0.0);
    acc0 = __dadd_rn(acc0, x0);
    acc1 = __dadd_rn(acc1, x1);
  }
}

// Alternating kernel: where a consumer warp stands in the tile (first-half records, then second-half records)
struct AltState {
  uint64_t* a_empty;
  uint64_t* b_full;
  uint32_t parity;
  bool in_y;
};
__device__ __forceinline__ void mbar_arrive(uint64_t* bar);

// All records of one chunk of the tile stream, processed by one warp.
template <bool ALT = false>
__device__ __forceinline__ void gather_chunk(const unsigned char* __restrict__ chunk, const TileParams& P,
                                             const double* __restrict__ slab, const uint16_t* __restrict__ rec, int lane,
                                             AltState* alt = nullptr) {
  const uint32_t nrec = (P.debug & 2) ? 0u : *reinterpret_cast<const uint32_t*>(chunk);
  const unsigned char* rp = chunk + kChunkHdr;
  for (uint32_t r = 0; r < nrec; ++r) {
    const uint32_t h = *reinterpret_cast<const uint32_t*>(rp);
    const uint32_t L = h & 0xFFu, b = (h >> 8) & 3u, stride = h >> 16;  // stride = lanes of the record: 64, 32 or 16
    const uint32_t* destp = reinterpret_cast<const uint32_t*>(rp + kRecHdr);
    const uint32_t l0 = lane & (stride - 1u);  // lanes beyond a 16-wide record shadow the first ones and never store
    const uint32_t dest0 = uint32_t(lane) < stride ? destp[l0] : kPadDest;
    const uint32_t dest1 = stride == 64u ? destp[lane + 32] : kPadDest;
    const uint16_t* __restrict__ ent0 = reinterpret_cast<const uint16_t*>(rp + kRecHdr + 4 * stride) + l0;
    const uint16_t* __restrict__ ent1 = ent0 + (stride == 64u ? 32 : 0);
    rp += kRecHdr + stride * (4u + 2u * L);
    if (ALT) {
      // first record of this warp that reads the second half of the slab: the warp is done with the first half
      // (released to the producers, who refill it for the next tile) and needs the second half of THIS tile
      if (!alt->in_y && ((P.yblock_mask >> b) & 1u)) {
        __syncwarp();
        if (lane == 0) mbar_arrive(alt->a_empty);
        mbar_wait(alt->b_full, alt->parity);
        alt->in_y = true;
      }
    }
    const TileBlockDev& B = P.blk[b];
    const uint16_t* __restrict__ brec = rec + B.recipe_off;
    const uint32_t sb = B.slot_bits, slot_mask = (1u << sb) - 1u;
    double acc0 = 0.0, acc1 = 0.0;
    bool any0 = false, any1 = false;
    switch (B.no * 8 + B.ni) {
      case 0: break;  // zero space: every contribution is an exact zero
      case 1 * 8 + 1: gather_record_direct(ent0, ent1, stride, L, slab, acc0, acc1, any0, any1); break;
      case 1 * 8 + 2: gather_record<1, 2>(ent0, ent1, stride, L, slab, brec, sb, slot_mask, acc0, acc1, any0, any1); break;
      case 1 * 8 + 3: gather_record<1, 3>(ent0, ent1, stride, L, slab, brec, sb, slot_mask, acc0, acc1, any0, any1); break;
      case 1 * 8 + 4: gather_record<1, 4>(ent0, ent1, stride, L, slab, brec, sb, slot_mask, acc0, acc1, any0, any1); break;
      case 2 * 8 + 2: gather_record<2, 2>(ent0, ent1, stride, L, slab, brec, sb, slot_mask, acc0, acc1, any0, any1); break;
      case 3 * 8 + 3: gather_record<3, 3>(ent0, ent1, stride, L, slab, brec, sb, slot_mask, acc0, acc1, any0, any1); break;
      default: gather_record<4, 4>(ent0, ent1, stride, L, slab, brec, sb, slot_mask, acc0, acc1, any0, any1); break;
    }
    // padding lanes carry zero entries (they read slab[0]) and never store
    if (P.check_classification && ((dest0 != kPadDest && (dest0 != kNoDest) != any0) ||
                                   (dest1 != kPadDest && (dest1 != kNoDest) != any1)))
      *P.changed = 1;
    if (P.debug & 4) continue;
    if (dest0 > kNoDest) B.values[dest0 - 2u] = acc0;
    if (dest1 > kNoDest) B.values[dest1 - 2u] = acc1;
  }
}

template <class Fn, int NE, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) tile_assemble_kernel(Fn fn, const __grid_constant__ TileParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* slab = reinterpret_cast<double*>(smem_raw);
  constexpr int NW = NT / 32;
  constexpr int NEE = NE > 0 ? NE : 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* myring = smem_raw + P.ring_off + size_t(warp) * kSlotsPerWarp * kChunkBytes;  // this warp's double buffer
  uint16_t* rec = reinterpret_cast<uint16_t*>(smem_raw + P.rec_off);
  uint64_t* mybar = reinterpret_cast<uint64_t*>(smem_raw + P.mbar_off) + warp * kSlotsPerWarp;
  __shared__ uint32_t s_hdr[3][8];
  auto fetch_header = [&](uint32_t* h) {  // thread 0: next tile from the dynamic scheduler
    const uint32_t t = atomicAdd(P.ticket, 1u);
    h[0] = t;
    if (t < P.ntiles) {
      h[1] = __ldg(P.tile_cell_ptr + t);
      h[2] = __ldg(P.tile_cell_ptr + t + 1);
      h[3] = __ldg(P.tile_chunk_ptr + t);
      h[4] = __ldg(P.tile_chunk_ptr + t + 1);
    }
  };
  if (lane == 0) {
    for (int s = 0; s < kSlotsPerWarp; ++s) mbar_init(&mybar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid == 0) {
    fetch_header(s_hdr[0]);
    fetch_header(s_hdr[1]);
  }
  for (int i = tid; i < P.recipe_bytes / 2; i += NT) rec[i] = reinterpret_cast<const uint16_t*>(P.recipes)[i];
  // this warp's chunk stream: chunks issued / consumed so far (slot = n & 1, parity = (n >> 1) & 1) and the
  // issue cursor (tile iteration it is on, next chunk, end of that tile's chunks); runs ahead across tiles
  uint32_t n_issued = 0, n_consumed = 0;
  uint32_t cur_it = 0xFFFFFFFFu, cur_chunk = 0, cur_end = 0;
  uint32_t eid_next[NEE];
  bool have_eids = false;
  const bool prof = (P.debug & 8) && lane == 0;
  long long tk[6] = {0, 0, 0, 0, 0, 0};
  long long t_last = clock64();
  auto lap = [&](int which) {
    if (prof) {
      const long long now = clock64();
      tk[which] += now - t_last;
      t_last = now;
    }
  };
  for (uint32_t it = 0;; ++it) {
    lap(5);
    __syncthreads();  // previous tile fully consumed, this tile's header visible
    lap(4);
    const uint32_t* hdr = s_hdr[it % 3];
    const uint32_t t = hdr[0];
    if (t >= P.ntiles) break;
    const uint32_t cbase = hdr[1], nc = hdr[2] - hdr[1];
    const uint32_t c0 = hdr[3], c1 = hdr[4];
    const uint32_t* hnext = s_hdr[(it + 1) % 3];
    auto issue_more = [&]() {  // keep this warp's double buffer full, crossing into the next tile when this one is done
      while (n_issued - n_consumed < uint32_t(kSlotsPerWarp)) {
        if (cur_chunk >= cur_end) {
          if (cur_it == it && hnext[0] < P.ntiles) {
            cur_it = it + 1;
            cur_chunk = hnext[3] + warp;
            cur_end = hnext[4];
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
        cur_chunk += NW;
        ++n_issued;
      }
    };
    if (cur_it != it) {  // the cursor did not run ahead into this tile: start here
      cur_it = it;
      cur_chunk = c0 + warp;
      cur_end = c1;
    }
    issue_more();  // lands while K1 runs
    // ---- K1: element values of the tile's cells -> shared slab [distinct][cell]
    if (c1 > c0 && !(P.debug & 1)) {
      for (uint32_t c = tid; c < nc; c += NT) {
        uint32_t eid[NEE];
        if (have_eids && c == uint32_t(tid)) {  // fetched while this thread waited at the previous tile's barrier
#pragma unroll
          for (int e = 0; e < NE; ++e) eid[e] = eid_next[e];
        } else {
          const uint32_t* ce = P.tile_cell_edges + size_t(cbase + c) * NE;
#pragma unroll
          for (int e = 0; e < NE; ++e) eid[e] = __ldg(ce + e);
        }
        double s[NEE];
#pragma unroll
        for (int e = 0; e < NE; ++e) s[e] = __ldg(P.lengths + (eid[e] - P.edge_lo));
        TileSink sink{slab + c, P.cstride};
        fn(s, sink);
      }
    }
    lap(0);
    __syncthreads();
    lap(1);
    if (tid == 0) fetch_header(s_hdr[(it + 2) % 3]);  // two tiles ahead: its latency hides behind this gather
    // ---- K3: this warp's chunks; one record at a time, two owned structural non-zeros per lane
    for (uint32_t c = c0 + warp; c < c1; c += NW) {
      lap(5);
      mbar_wait(&mybar[n_consumed & 1u], (n_consumed >> 1) & 1u);
      lap(2);
      const unsigned char* chunk = myring + (n_consumed & 1u) * kChunkBytes;
      gather_chunk(chunk, P, slab, rec, lane);
      __syncwarp();  // every lane is done reading the slot before it is refilled
      lap(3);
      ++n_consumed;
      issue_more();
    }
    // edge ids of this thread's cell of the next tile: issued now, they arrive while the warp waits at the tile barrier
    have_eids = !(P.debug & 16) && hnext[0] < P.ntiles;
    if (have_eids && uint32_t(tid) < hnext[2] - hnext[1]) {
      const uint32_t* ce = P.tile_cell_edges + size_t(hnext[1] + tid) * NE;
#pragma unroll
      for (int e = 0; e < NE; ++e) eid_next[e] = __ldg(ce + e);
    }
  }
  if (prof)
    for (int i = 0; i < 6; ++i) atomicAdd(P.stats + i, (unsigned long long)tk[i]);
}


// ---- warp-specialised variant --------------------------------------------------------------------------------
// 8 producer warps evaluate the element values of tile i+1 into one of two shared slabs while 16 consumer warps
// reduce tile i out of the other: the FP64 pipe (K1) and the shared-memory crossbar (gather) work concurrently and
// no warp idles at a CTA barrier.  setmaxnreg gives the producers the 128 registers the straight-line tape needs
// and leaves 56 to each consumer (launch: 768 x 80; the consumers release 16*32*24 = the 8*32*48 the producers acquire).  Hand-offs are mbarriers:
// Measured on B200: 6.7 ms vs 5.8 ms for the phase-serialised kernel (the two slabs halve the tile, K1 of a small tile
// is latency-bound at ~3 us) - kept behind FQ_TILE_KERNEL=w.
//   hdr_ready[b]  (1 arrival)   the tile of buffer b is known (consumers start prefetching its stream)
//   slab_full[b]  (8 arrivals)  every producer warp has stored its cells
//   slab_empty[b] (16 arrivals) every consumer warp is done with the tile
constexpr int kWsProducerWarps = 8, kWsConsumerWarps = 16;
constexpr int kWsThreads = 32 * (kWsProducerWarps + kWsConsumerWarps);

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

template <class Fn, int NE>
__global__ void __launch_bounds__(kWsThreads, 1) tile_assemble_ws_kernel(Fn fn, const __grid_constant__ TileParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NEE = NE > 0 ? NE : 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* slabs[2] = {reinterpret_cast<double*>(smem_raw), reinterpret_cast<double*>(smem_raw + P.slab_bytes)};
  uint16_t* rec = reinterpret_cast<uint16_t*>(smem_raw + P.rec_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + P.mbar_off);
  uint64_t* hdr_ready = bars;         // [2]
  uint64_t* slab_full = bars + 2;     // [2]
  uint64_t* slab_empty = bars + 4;    // [2]
  uint64_t* tma_bar = bars + 6;       // [consumer warp][2]
  __shared__ uint32_t s_hdr[2][8];
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hdr_ready[b], 1);
      mbar_init(&slab_full[b], kWsProducerWarps);
      mbar_init(&slab_empty[b], kWsConsumerWarps);
    }
    for (int i = 0; i < 2 * kWsConsumerWarps; ++i) mbar_init(&tma_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < P.recipe_bytes / 2; i += kWsThreads) rec[i] = reinterpret_cast<const uint16_t*>(P.recipes)[i];
  __syncthreads();
  if (warp < kWsProducerWarps) {
    // ------------------------------------------------------------------ producers: K1
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    uint32_t h[5] = {0xFFFFFFFFu, 0, 0, 0, 0};
    auto fetch = [&]() {  // thread 0: next tile from the dynamic scheduler (latency hidden behind the current K1)
      h[0] = atomicAdd(P.ticket, 1u);
      if (h[0] < P.ntiles) {
        h[1] = __ldg(P.tile_cell_ptr + h[0]);
        h[2] = __ldg(P.tile_cell_ptr + h[0] + 1);
        h[3] = __ldg(P.tile_chunk_ptr + h[0]);
        h[4] = __ldg(P.tile_chunk_ptr + h[0] + 1);
      }
    };
    if (tid == 0) fetch();
    for (uint32_t it = 0;; ++it) {
      const uint32_t b = it & 1u, u = it >> 1;
      if (it >= 2) mbar_wait(&slab_empty[b], (u - 1u) & 1u);  // the consumers left the tile that used this buffer
      if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) s_hdr[b][i] = h[i];
        mbar_arrive(&hdr_ready[b]);
      }
      mbar_wait(&hdr_ready[b], u & 1u);
      const uint32_t t = s_hdr[b][0];
      if (t >= P.ntiles) break;
      const uint32_t cbase = s_hdr[b][1], nc = s_hdr[b][2] - s_hdr[b][1];
      const bool work = s_hdr[b][4] > s_hdr[b][3];
      if (tid == 0) fetch();
      if (work && !(P.debug & 1)) {
        double* slab = slabs[b];
        for (uint32_t c = tid; c < nc; c += 32 * kWsProducerWarps) {
          const uint32_t* ce = P.tile_cell_edges + size_t(cbase + c) * NE;
          uint32_t eid[NEE];
#pragma unroll
          for (int e = 0; e < NE; ++e) eid[e] = __ldg(ce + e);
          double s[NEE];
#pragma unroll
          for (int e = 0; e < NE; ++e) s[e] = __ldg(P.lengths + (eid[e] - P.edge_lo));
          TileSink sink{slab + c, P.cstride};
          fn(s, sink);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&slab_full[b]);
    }
  } else {
    // ------------------------------------------------------------------ consumers: K3
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");  // 16 warps x 24 released = the 12288 registers the producers acquire
    const int cw = warp - kWsProducerWarps;
    unsigned char* myring = smem_raw + P.ring_off + size_t(cw) * kSlotsPerWarp * kChunkBytes;
    uint64_t* mybar = tma_bar + cw * kSlotsPerWarp;
    uint32_t n_issued = 0, n_consumed = 0;
    uint32_t cur_it = 0xFFFFFFFFu, cur_chunk = 0, cur_end = 0;
    for (uint32_t it = 0;; ++it) {
      const uint32_t b = it & 1u, u = it >> 1;
      mbar_wait(&hdr_ready[b], u & 1u);
      const uint32_t t = s_hdr[b][0];
      if (t >= P.ntiles) break;
      const uint32_t c0 = s_hdr[b][3], c1 = s_hdr[b][4];
      auto issue_more = [&]() {  // keep this warp's double buffer full, crossing into the next tile once it is known
        while (n_issued - n_consumed < uint32_t(kSlotsPerWarp)) {
          if (cur_chunk >= cur_end) {
            if (cur_it == it && mbar_test(&hdr_ready[b ^ 1u], ((it + 1u) >> 1) & 1u) && s_hdr[b ^ 1u][0] < P.ntiles) {
              cur_it = it + 1;
              cur_chunk = s_hdr[b ^ 1u][3] + cw;
              cur_end = s_hdr[b ^ 1u][4];
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
          cur_chunk += kWsConsumerWarps;
          ++n_issued;
        }
      };
      if (cur_it != it) {
        cur_it = it;
        cur_chunk = c0 + cw;
        cur_end = c1;
      }
      issue_more();
      mbar_wait(&slab_full[b], u & 1u);
      const double* slab = slabs[b];
      for (uint32_t c = c0 + cw; c < c1; c += kWsConsumerWarps) {
        mbar_wait(&mybar[n_consumed & 1u], (n_consumed >> 1) & 1u);
        gather_chunk(myring + (n_consumed & 1u) * kChunkBytes, P, slab, rec, lane);
        __syncwarp();
        ++n_consumed;
        issue_more();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&slab_empty[b]);
    }
  }
}

// ---- alternating producer/consumer variant (the default) -----------------------------------------------------------
// One full-size slab (the same tiles and record streams as the phase-serialised kernel), logically split in two halves:
//   half 1 = the stored values of M_{k-1} and M_k      read by the records of M_{k-1}, M_k, dif_test(k)
//   half 2 = the stored values of M_{k+1}              read by the records of dif_both(k+1)
// The records of a tile are laid out block by block, so every consumer warp first works through first-half records
// and then through second-half records.  While the consumers are in the second half of tile i the 8 producer warps
// evaluate half 1 of tile i+1 (tape stage B1), and while they are in the first half of tile i+1 the producers
// evaluate its half 2 (stage B2): the FP64 pipe works in the shadow of the shared-memory-bound gather with NO second
// slab, i.e. without shrinking the tiles.  Tiles are dealt statically (tile = blockIdx.x + it * gridDim.x).
//   a_full / b_full   (8 arrivals)   the producers stored half 1 / half 2 of tile it
//   a_empty / b_empty (16 arrivals)  every consumer warp is past its first-half / second-half records of tile it
// Stage A, B1, B2 execute exactly the operations of the unsplit tape: results are bit-identical.
constexpr int kAltCellsPerThread = 2;  // a tile has at most 2 * 256 cells

// NC consumer warps; PR / CR registers per producer / consumer thread after setmaxnreg (the launch allocates
// LR = 65536 / threads rounded down to 8 per thread; what the consumers release must cover what the producers acquire)
template <class Fn, int NE, int NC, int PR, int CR>
__global__ void __launch_bounds__(32 * (kWsProducerWarps + NC), 1) tile_assemble_alt_kernel(Fn fn, const __grid_constant__ TileParams P) {
  constexpr int kThreads = 32 * (kWsProducerWarps + NC);
  constexpr int LR = 65536 / kThreads / 8 * 8;
  static_assert(PR % 8 == 0 && CR % 8 == 0 && CR <= LR && PR >= LR && 32 * NC * (LR - CR) >= 256 * (PR - LR), "register split");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NEE = NE > 0 ? NE : 1;
  constexpr int NM = Fn::kMid;
  constexpr int NP = 32 * kWsProducerWarps;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* slab = reinterpret_cast<double*>(smem_raw);
  uint16_t* rec = reinterpret_cast<uint16_t*>(smem_raw + P.rec_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + P.mbar_off);
  uint64_t* tma_bar = bars;                                       // [consumer warp][2]
  uint64_t* a_full = bars + NC * kSlotsPerWarp;     // the 8 spare barriers of the layout
  uint64_t* a_empty = a_full + 1;
  uint64_t* b_full = a_full + 2;
  uint64_t* b_empty = a_full + 3;
  if (tid == 0) {
    mbar_init(a_full, kWsProducerWarps);
    mbar_init(b_full, kWsProducerWarps);
    mbar_init(a_empty, NC);
    mbar_init(b_empty, NC);
    for (int i = 0; i < kSlotsPerWarp * NC; ++i) mbar_init(&tma_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < P.recipe_bytes / 2; i += kThreads) rec[i] = reinterpret_cast<const uint16_t*>(P.recipes)[i];
  __syncthreads();
  const uint32_t G = gridDim.x, t0 = blockIdx.x;
  if (warp < kWsProducerWarps) {
    // ------------------------------------------------------------------ producers: K1, two cells per thread
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PR));
    auto load_hdr = [&](uint64_t t, uint32_t& cb, uint32_t& nc) {
      cb = 0;
      nc = 0;
      if (t < uint64_t(P.ntiles)) {
        const uint32_t a = __ldg(P.tile_cell_ptr + t), e = __ldg(P.tile_cell_ptr + t + 1);
        const uint32_t c0 = __ldg(P.tile_chunk_ptr + t), c1 = __ldg(P.tile_chunk_ptr + t + 1);
        cb = a;
        nc = (c1 > c0 && !(P.debug & 1)) ? e - a : 0u;
      }
    };
    auto load_ids = [&](uint32_t cb, uint32_t nc, uint32_t (*eid)[NEE]) {
#pragma unroll
      for (int j = 0; j < kAltCellsPerThread; ++j) {
        const uint32_t c = uint32_t(tid) + uint32_t(j) * NP;
        if (c < nc) {
          const uint32_t* ce = P.tile_cell_edges + size_t(cb + c) * NE;
#pragma unroll
          for (int e = 0; e < NE; ++e) eid[j][e] = __ldg(ce + e);
        }
      }
    };
    uint32_t cb0, nc0, cb1, nc1, cb2, nc2;
    load_hdr(t0, cb0, nc0);
    load_hdr(uint64_t(t0) + G, cb1, nc1);
    uint32_t eid[kAltCellsPerThread][NEE];
    load_ids(cb0, nc0, eid);
    for (uint32_t it = 0;; ++it) {
      const uint64_t t = uint64_t(t0) + uint64_t(it) * G;
      if (t >= uint64_t(P.ntiles)) break;
      load_hdr(t + 2 * uint64_t(G), cb2, nc2);  // cell range two tiles ahead (its ids are loaded next iteration)
      // edge lengths and stage A (metric, inverse, volume) of this tile's cells; the consumers are still busy with the
      // previous tile, so this latency is off the critical path
      double mid[kAltCellsPerThread][NM];
#pragma unroll
      for (int j = 0; j < kAltCellsPerThread; ++j) {
        const uint32_t c = uint32_t(tid) + uint32_t(j) * NP;
        if (c < nc0) {
          double sl[NEE];
#pragma unroll
          for (int e = 0; e < NE; ++e) sl[e] = __ldg(P.lengths + (eid[j][e] - P.edge_lo));
          fn.a(sl, mid[j]);
        }
      }
      load_ids(cb1, nc1, eid);  // ids of the next tile: in flight while the two halves are evaluated
      const uint32_t par_prev = (it - 1u) & 1u;
      if (it >= 1) mbar_wait(a_empty, par_prev);  // every consumer warp is past the first-half records of the previous tile
#pragma unroll
      for (int j = 0; j < kAltCellsPerThread; ++j) {
        const uint32_t c = uint32_t(tid) + uint32_t(j) * NP;
        if (c < nc0) {
          TileSink sink{slab + c, P.cstride};
          fn.b1(mid[j], sink);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      if (it >= 1) mbar_wait(b_empty, par_prev);  // ... and past its second-half records
#pragma unroll
      for (int j = 0; j < kAltCellsPerThread; ++j) {
        const uint32_t c = uint32_t(tid) + uint32_t(j) * NP;
        if (c < nc0) {
          TileSink sink{slab + c, P.cstride};
          fn.b2(mid[j], sink);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(b_full);
      cb0 = cb1;
      nc0 = nc1;
      cb1 = cb2;
      nc1 = nc2;
    }
  } else {
    // ------------------------------------------------------------------ consumers: K3
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CR));  // the 16 consumer warps release what the 8 producer warps acquire
    const int cw = warp - kWsProducerWarps;
    unsigned char* myring = smem_raw + P.ring_off + size_t(cw) * kSlotsPerWarp * kChunkBytes;
    uint64_t* mybar = tma_bar + cw * kSlotsPerWarp;
    uint32_t n_issued = 0, n_consumed = 0;
    uint32_t cur_it = 0xFFFFFFFFu, cur_chunk = 0, cur_end = 0;
    auto load_chunks = [&](uint64_t t, uint32_t& a, uint32_t& e) {
      a = 0;
      e = 0;
      if (t < uint64_t(P.ntiles)) {
        a = __ldg(P.tile_chunk_ptr + t);
        e = __ldg(P.tile_chunk_ptr + t + 1);
      }
    };
    uint32_t c0, c1, c0n = 0, c1n = 0;
    load_chunks(t0, c0, c1);
    // Chunk c of a tile goes to the warp (c - rotation) mod NC, the rotation advancing with every tile: a tile has
    // ~3.75 chunks per warp, and without the rotation the same warps would get the extra chunk of every tile.
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
      AltState st{a_empty, b_full, it & 1u, false};
      mbar_wait(a_full, it & 1u);
      for (uint32_t c = c0 + lane_of(it); c < c1; c += NC) {
        mbar_wait(&mybar[n_consumed & 1u], (n_consumed >> 1) & 1u);
        gather_chunk<true>(myring + (n_consumed & 1u) * kChunkBytes, P, slab, rec, lane, &st);
        __syncwarp();
        ++n_consumed;
        issue_more();
      }
      __syncwarp();
      if (!st.in_y) {
        // This warp had no second-half records in the tile.  It must still observe b_full(it) before it arrives on
        // b_empty: otherwise it could run a whole tile ahead of a slow warp and its arrival for tile it+1 would
        // complete phase it of b_empty while that warp still reads the second half of tile it.
        if (lane == 0) mbar_arrive(a_empty);
        mbar_wait(b_full, it & 1u);
      }
      if (lane == 0) mbar_arrive(b_empty);
      c0 = c0n;
      c1 = c1n;
    }
  }
}

// ------------------------------------------------------------------ plan
struct TileBlockPlan {
  int no = 1, ni = 1, recipe_off = 0, slot_bits = 7;
  fq_csr* csr = nullptr;
  size_t nnz_at_build = 0;
  bool dropped_at_build = false;
};

struct TilePlan {
  const fq_mesh* mesh = nullptr;
  int dim = 0, core_k = 0, ndistinct = 0;
  uint32_t ntiles = 0;
  int cstride = 0;
  int nthreads = 512;
  bool ws = false;
  bool alt = false;           // FQ_TILE_KERNEL=a and the block set splits: tile_assemble_alt_kernel
  uint32_t yblock_mask = 0;   // blocks reading only second-half values
  bool pack = false;          // bank-aware lane packing (slab stride = 0 mod 16)
  int stream_warps = 0;       // alternating kernel with 20 / 24 consumer warps (0: 16)
  uint32_t slab_bytes = 0;
  size_t smem_bytes = 0;
  uint32_t ring_off = 0, rec_off = 0, mbar_off = 0;
  int recipe_bytes = 0;
  DevBuf<uint32_t> tile_cell_ptr, tile_cell_edges, tile_chunk_ptr;
  DevBuf<unsigned char> stream;
  DevBuf<uint8_t> recipes;
  DevBuf<int> changed;
  DevBuf<unsigned int> ticket;
  DevBuf<unsigned long long> stats;
  int nblocks = 0;
  TileBlockPlan blk[kTileMaxBlocks];
  int grid = 0;
  void (*launch)(fq_ctx*, const TilePlan&, const TileParams&) = nullptr;
};

#define FQ_DECLARE_CORE(fn, n, k, variant, nin, nd, nout)                                      \
  struct Core_##fn {                                                                           \
    static constexpr int kDistinct = nd;                                                       \
    static constexpr int kMid = fn##_nmid; /* values live across the stage A / stage B cut */  \
    template <class S>                                                                         \
    __device__ __forceinline__ void operator()(const double* __restrict__ s, S& sink) const {  \
      fn(s, sink);                                                                             \
    }                                                                                          \
    __device__ __forceinline__ void a(const double* __restrict__ s, double* __restrict__ mid) const { \
      fn##_a(s, mid);                                                                          \
    }                                                                                          \
    template <class S>                                                                         \
    __device__ __forceinline__ void b1(const double* __restrict__ mid, S& sink) const {        \
      fn##_b1(mid, sink);                                                                      \
    }                                                                                          \
    template <class S>                                                                         \
    __device__ __forceinline__ void b2(const double* __restrict__ mid, S& sink) const {        \
      fn##_b2(mid, sink);                                                                      \
    }                                                                                          \
  };
FQ_GEN_CORE_LIST(FQ_DECLARE_CORE)
#undef FQ_DECLARE_CORE

template <class Fn, int NE, int NT, int MINB>
static void launch_tile_nt(fq_ctx* ctx, const TilePlan& plan, const TileParams& params) {
  static bool attr_set = false;
  if (!attr_set) {
    FQ_CUDA(cudaFuncSetAttribute(tile_assemble_kernel<Fn, NE, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (MINB == 1 ? 227 : 113) * 1024 - 256));
    attr_set = true;
  }
  tile_assemble_kernel<Fn, NE, NT, MINB><<<plan.grid, NT, plan.smem_bytes, ctx->stream>>>(Fn{}, params);
}
template <class Fn, int NE>
static void launch_tile_ws(fq_ctx* ctx, const TilePlan& plan, const TileParams& params) {
  static bool attr_set = false;
  if (!attr_set) {
    FQ_CUDA(cudaFuncSetAttribute(tile_assemble_ws_kernel<Fn, NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    attr_set = true;
  }
  tile_assemble_ws_kernel<Fn, NE><<<plan.grid, kWsThreads, plan.smem_bytes, ctx->stream>>>(Fn{}, params);
}
template <class Fn, int NE, int NC, int PR, int CR>
static void launch_tile_alt_v(fq_ctx* ctx, const TilePlan& plan, const TileParams& params) {
  static bool attr_set = false;
  if (!attr_set) {
    FQ_CUDA(cudaFuncSetAttribute(tile_assemble_alt_kernel<Fn, NE, NC, PR, CR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024 - 256));
    attr_set = true;
  }
  tile_assemble_alt_kernel<Fn, NE, NC, PR, CR><<<plan.grid, 32 * (kWsProducerWarps + NC), plan.smem_bytes, ctx->stream>>>(Fn{}, params);
}
template <class Fn, int NE>
static void launch_tile_alt(fq_ctx* ctx, const TilePlan& plan, const TileParams& params) {
  static int variant = -1;
  if (variant < 0) {
    const char* e = std::getenv("FQ_ALT_REGS");  // tuning (16 consumer warps): 0 = 128/56, 1 = 112/64 (default), 2 = 96/72
    variant = e ? std::atoi(e) : 1;
    if (variant < 0 || variant > 2) variant = 1;
  }
  if (plan.stream_warps == 24)
    launch_tile_alt_v<Fn, NE, 24, 88, 56>(ctx, plan, params);
  else if (plan.stream_warps == 20)
    launch_tile_alt_v<Fn, NE, 20, 88, 64>(ctx, plan, params);
  else if (variant == 2)
    launch_tile_alt_v<Fn, NE, 16, 96, 72>(ctx, plan, params);
  else if (variant == 0)
    launch_tile_alt_v<Fn, NE, 16, 128, 56>(ctx, plan, params);
  else
    launch_tile_alt_v<Fn, NE, 16, 112, 64>(ctx, plan, params);
}
template <class Fn, int NE>
static void launch_tile(fq_ctx* ctx, const TilePlan& plan, const TileParams& params) {
  if (plan.alt)
    launch_tile_alt<Fn, NE>(ctx, plan, params);
  else if (plan.ws)
    launch_tile_ws<Fn, NE>(ctx, plan, params);
  else if (plan.nthreads == 256)
    launch_tile_nt<Fn, NE, 256, 2>(ctx, plan, params);  // two CTAs per SM: one tile's K1 overlaps the other's gather
  else
    launch_tile_nt<Fn, NE, 512, 1>(ctx, plan, params);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

// variant 0: stores M_{k-1}, M_k, M_{k+1};  variant 1: stores M_{k-1}, M_k, dif_both(k+1)
struct CoreEntryRt {
  int n, k, variant, nin, ndistinct, nouts;
  const short* map;
  int split_ok;                      // the stored values split into two halves (gen_elmat.cpp)
  const unsigned long long* half2;   // bit s: distinct slot s belongs to the second half
  void (*launch)(fq_ctx*, const TilePlan&, const TileParams&);
};
#define FQ_CORE_ENTRY(fn, n, k, variant, nin, nd, nout) \
  CoreEntryRt{n, k, variant, nin, nd, nout, fn##_map, fn##_split_ok, fn##_half2, &launch_tile<Core_##fn, nin>},
static const CoreEntryRt g_cores[] = {FQ_GEN_CORE_LIST(FQ_CORE_ENTRY)};
#undef FQ_CORE_ENTRY

// offset of the stored block (kind, g) inside the core's map, or -1 when the core does not store it
static int stored_offset(const CoreEntryRt& core, int kind, int g) {
  const BlockSpec stored[3] = {{KIND_MASS, core.k - 1},
                               {KIND_MASS, core.k},
                               {core.variant == 1 ? int(KIND_DIF_BOTH) : int(KIND_MASS), core.k + 1}};
  int off = 0;
  for (const BlockSpec& b : stored) {
    int tg, rg;
    kind_grades(b.kind, b.grade, tg, rg);
    if (b.kind == kind && b.grade == g) return off;
    off += nlocal(core.n, tg) * nlocal(core.n, rg);
  }
  return -1;
}

// ---- launch configuration (tunable through the environment for sweeps) -------
struct TileConfig {
  int nthreads;
  size_t smem_cta;  // dynamic shared memory budget of the CTA
  bool ws;          // warp-specialised kernel: two slabs, 8 producer + 16 consumer warps
  bool alt = false;           // alternating kernel: the layout of the phase-serialised kernel, 8 + 16 warps
  int stream_warps = 0;       // warps that stream chunks (0: nthreads / 32, or the 16 consumers of the w/p kernels)
};
static TileConfig tile_config() {
  TileConfig c{512, size_t(227) * 1024 - 256, false};
  c.alt = true;  // default: the alternating producer/consumer kernel (FQ_TILE_KERNEL=s: the phase-serialised one)
  if (const char* e = std::getenv("FQ_TILE_THREADS"))
    if (std::atoi(e) == 256) c = TileConfig{256, size_t(113) * 1024 - 256, false};
  if (const char* e = std::getenv("FQ_TILE_KERNEL")) {
    if (e[0] == 'w') c = TileConfig{kWsThreads, size_t(227) * 1024 - 256, true};
    if (e[0] == 's') c.alt = false;  // same plan (16 streaming warps, one slab, tiles of <= 512 cells), 512-thread kernel
  }
  if (c.alt)
    if (const char* e = std::getenv("FQ_ALT_CONSUMERS")) {  // tuning: 20 or 24 consumer warps (larger ring, smaller tiles)
      const int n = std::atoi(e);
      if (n == 20 || n == 24) c.stream_warps = n;
    }
  return c;
}
static size_t tile_fixed_smem(const TileConfig& c) {
  const size_t nwarps = c.stream_warps ? size_t(c.stream_warps)
                                       : (c.ws ? size_t(kWsConsumerWarps) : size_t(c.nthreads) / 32);  // warps that stream chunks
  return nwarps * kSlotsPerWarp * kChunkBytes /*ring*/ + 2048 /*recipes*/ + (nwarps * kSlotsPerWarp + 8) * 8 /*mbarriers*/ +
         512 /*alignment slack*/;
}
int tile_cells_capacity(int ndistinct) {
  const TileConfig c = tile_config();
  const size_t slabs = c.ws ? 2 : 1;
  const int cap = int((c.smem_cta - tile_fixed_smem(c)) / slabs / (size_t(ndistinct) * sizeof(double)));
  return std::min(cap, c.ws ? 32 * kWsProducerWarps : c.nthreads);  // K1 evaluates one cell per thread in one pass
}
static int tile_core_variant() {
  // default: masses only (54 doubles per 3-D cell -> larger tiles); FQ_TILE_CORE=h also stores dif_both(k+1)
  const char* e = std::getenv("FQ_TILE_CORE");
  return (e && e[0] == 'h') ? 1 : 0;
}

// Recipes of one block over the distinct values of core(n, kc).
// Returns codes[nslots][no*ni]; mass entry (g, i, j) -> map code.
static bool build_recipes(int n, int kc, const CoreEntryRt& core, int kind, int g, int& no, int& ni,
                          std::vector<uint8_t>& codes) {
  if (kind == KIND_LUMPED) return false;
  int tg, rg;
  kind_grades(kind, g, tg, rg);
  const int rows = nlocal(n, tg), cols = nlocal(n, rg);
  const int nslots = rows * cols;
  (void)kc;
  if (g < 0 || g > n || nslots == 0) {  // zero space: every entry is an exact zero
    no = 0;
    ni = 0;
    codes.clear();
    return true;
  }
  if (core.ndistinct > 127) return false;
  auto direct_code = [&](int m) -> int { return m < 0 ? 0xFF : ((m & 0xFF) | ((m & 0x100) ? 0x80 : 0)); };
  const int doff = stored_offset(core, kind, g);
  if (doff >= 0) {  // the block itself is stored: every slot is one signed stored value
    no = 1, ni = 1;
    codes.resize(size_t(nslots));
    for (int sl = 0; sl < nslots; ++sl) codes[size_t(sl)] = uint8_t(direct_code(core.map[doff + sl]));
    return true;
  }
  // otherwise a sandwich of the stored mass of grade g
  const int off = stored_offset(core, KIND_MASS, g);
  if (off < 0 || kind == KIND_MASS) return false;
  const int nd = nlocal(n, g);
  auto mcode = [&](int i, int j, int sign) -> int {  // signed mass entry -> code or -1 (zero)
    const int m = core.map[off + i * nd + j];
    if (m < 0) return 0xFF;
    const int slot = m & 0xFF;
    const bool neg = ((m & 0x100) != 0) != (sign < 0);
    return slot | (neg ? 0x80 : 0);
  };
  // boundary operator rows -> list of (coface index, sign), ascending
  struct Inc {
    int idx, sign;
  };
  std::vector<std::vector<Inc>> brow;  // brow[face] = its cofaces of grade g
  if (kind != KIND_MASS) {
    brow.assign(size_t(nlocal(n, g - 1)), {});
    const auto cof = colex_subsets(n + 1, g + 1);
    for (size_t ic = 0; ic < cof.size(); ++ic) {
      const auto el = mask_elems(cof[ic]);
      for (size_t pos = 0; pos < el.size(); ++pos) {
        const uint32_t face = cof[ic] & ~(1u << el[pos]);
        brow[size_t(colex_rank(face))].push_back(Inc{int(ic), (pos & 1) ? -1 : 1});
      }
    }
    for (auto& r : brow) std::sort(r.begin(), r.end(), [](const Inc& a, const Inc& b) { return a.idx < b.idx; });
  }
  const int ncof = (kind == KIND_MASS) ? 1 : (n + 1 - g);
  if (kind == KIND_MASS) {
    no = 1, ni = 1;
  } else if (kind == KIND_DIF_BOTH) {
    no = ncof, ni = ncof;
  } else {
    no = 1, ni = ncof;
  }
  codes.assign(size_t(nslots) * no * ni, 0xFF);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      uint8_t* out = codes.data() + size_t(r * cols + c) * no * ni;
      if (kind == KIND_MASS) {
        out[0] = uint8_t(mcode(r, c, 1));
      } else if (kind == KIND_DIF_TRIAL) {  // (M * B^T)[r][c] = sum_m M[r][m] * B[c][m]
        const auto& bc = brow[size_t(c)];
        for (size_t q = 0; q < bc.size(); ++q) out[q] = uint8_t(mcode(r, bc[q].idx, bc[q].sign));
      } else if (kind == KIND_DIF_TEST) {  // (B * M)[r][c] = sum_k B[r][k] * M[k][c]
        const auto& br = brow[size_t(r)];
        for (size_t q = 0; q < br.size(); ++q) out[q] = uint8_t(mcode(br[q].idx, c, br[q].sign));
      } else {  // B * (M * B^T): outer over k (cofaces of r), inner over m (cofaces of c)
        const auto& br = brow[size_t(r)];
        const auto& bc = brow[size_t(c)];
        for (size_t o = 0; o < br.size(); ++o)
          for (size_t q = 0; q < bc.size(); ++q)
            out[o * size_t(ni) + q] = uint8_t(mcode(br[o].idx, bc[q].idx, br[o].sign * bc[q].sign));
      }
    }
  return true;
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
// type-major slot order inside every tile: key = tile | cell type | cell
__global__ void tile_slot_keys_kernel(const uint32_t* __restrict__ tile_of, const uint32_t* __restrict__ tile_cells, size_t n,
                                      uint32_t period, uint64_t cell_offset, uint64_t* __restrict__ keys,
                                      uint32_t* __restrict__ pos) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t type = period ? (cell_offset + tile_cells[i]) % period : 0;
    keys[i] = (uint64_t(tile_of[i]) << 35) | (type << 32) | uint64_t(tile_cells[i]);
    pos[i] = uint32_t(i);
  }
}
// after the sort: slot i holds the cell at (tile, cell)-sorted position pos[i]
__global__ void tile_slots_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ pos, size_t n,
                                  const uint32_t* __restrict__ tile_cell_ptr, uint32_t* __restrict__ cells_by_slot,
                                  uint32_t* __restrict__ local_of_pos) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t t = uint32_t(keys[i] >> 35);
    cells_by_slot[i] = uint32_t(keys[i]);
    local_of_pos[pos[i]] = uint32_t(i) - tile_cell_ptr[t];
  }
}
__global__ void tile_cell_edges_kernel(const uint32_t* __restrict__ tile_cells, size_t n, const uint32_t* __restrict__ cell_edges,
                                       int ne, uint32_t* __restrict__ out) {
  const size_t total = n * size_t(ne);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += stride)
    out[p] = cell_edges[size_t(tile_cells[p / size_t(ne)]) * ne + p % size_t(ne)];
}

// row -> tile of its top vertex
__global__ void row_tile_kernel(const uint32_t* __restrict__ faces, int nl, const uint32_t* __restrict__ cell_verts, int nv,
                                const uint8_t* __restrict__ top_pos, size_t ncells, const uint32_t* __restrict__ vtile,
                                uint32_t v_lo, uint32_t row_begin, uint32_t row_end, uint32_t* __restrict__ row_tile) {
  const size_t total = ncells * size_t(nl);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += stride) {
    const size_t c = p / size_t(nl);
    const int i = int(p % size_t(nl));
    const uint32_t row = faces[p];
    if (row < row_begin || row >= row_end) continue;
    row_tile[row - row_begin] = vtile[cell_verts[c * nv + top_pos[i]] - v_lo];
  }
}
// sort key of every structural non-zero: tile | L | signature of its slot sequence
__global__ void nnz_key_kernel(const uint32_t* __restrict__ row_ptr, uint32_t nrows, const uint32_t* __restrict__ row_tile,
                               const uint32_t* __restrict__ contrib_ptr, const uint32_t* __restrict__ contrib_src, uint32_t T,
                               int use_sig, const uint32_t* __restrict__ tile_cell_ptr, const uint32_t* __restrict__ tile_cells,
                               const uint32_t* __restrict__ tile_local, uint64_t* __restrict__ key,
                               uint32_t* __restrict__ nnz_id, int* __restrict__ err) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const uint64_t t = row_tile[r];
    for (uint32_t q = row_ptr[r]; q < row_ptr[r + 1]; ++q) {
      const uint32_t p0 = contrib_ptr[q], p1 = contrib_ptr[q + 1];
      uint32_t L = p1 - p0;
      if (L > uint32_t(kMaxLen)) {  // does not fit a 32-lane record of one chunk: the slab path handles this mesh
        atomicExch(err, 4);
        L = kMaxLen;
      }
      uint32_t sig = 0;
      if (use_sig) {  // "type" of the non-zero: its slot sequence and the relative cell offsets of its contributions
        const uint32_t cell0 = contrib_src[p0] / T;
        for (uint32_t p = p0; p < p1; ++p) {
          const uint32_t cell = contrib_src[p] / T;
          sig = (sig * 131u + (contrib_src[p] - cell * T) + 1u) * 31u + (cell - cell0);
        }
      }
      sig = (sig ^ (sig >> 16)) & 0xFFFFu;
      // shared-memory bank (8-byte units, 16 per wavefront) of the slab slot of the first contributing cell
      uint32_t bank = 0;
      if (tile_local) {
        const uint32_t cell = contrib_src[p0] / T;
        uint32_t lo = tile_cell_ptr[t], hi = tile_cell_ptr[t + 1];
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (tile_cells[mid] < cell)
            lo = mid + 1;
          else
            hi = mid;
        }
        bank = tile_local[lo] & 15u;
      }
      // tile | L | signature | occurrence (filled in later) | bank
      key[q] = (t << 36) | (uint64_t(L) << 28) | (uint64_t(sig) << 12) | bank;
      nnz_id[q] = q;
    }
  }
}
// Bank interleaving: within a group of equal (tile, L, signature, bank) the k-th member gets occurrence k; sorting by
// (.., occurrence, bank) then makes consecutive lanes of a record hit different banks of the slab.
__global__ void group_start_kernel(const uint64_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ start) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    start[i] = (i == 0 || key[i] != key[i - 1]) ? i : 0u;
}
__global__ void occurrence_kernel(uint64_t* __restrict__ key, uint32_t n, const uint32_t* __restrict__ start_scan) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t occ = min(i - start_scan[i], 255u);
    key[i] |= uint64_t(occ) << 4;
  }
}
__global__ void run_heads_kernel(const uint64_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ head) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    head[i] = (i == 0 || (key[i] >> 28) != (key[i - 1] >> 28)) ? 1u : 0u;
}
// run r: start, tile, L
__global__ void run_info_kernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ head,
                                const uint32_t* __restrict__ run_scan /*inclusive*/, uint32_t n, uint32_t* __restrict__ run_start,
                                uint32_t* __restrict__ run_tile, uint32_t* __restrict__ run_len_code) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (head[i]) {
      const uint32_t r = run_scan[i] - 1;
      run_start[r] = i;
      run_tile[r] = uint32_t(key[i] >> 36);
      run_len_code[r] = uint32_t(key[i] >> 28) & 0xFFu;
    }
}
// ---- bank-aware lane packing ---------------------------------------------------------------------------------------
// The gather's dominant cost is shared-memory wavefronts of the 8-byte slab loads: 32 lanes read slab[value * cstride +
// cell slot], served as two half-warps of 16 lanes over 16 8-byte bank pairs.  With cstride a multiple of 16 the bank pair
// of EVERY load of a contribution is (cell slot mod 16), whatever value it reads, so a half-warp is conflict-free at
// step j iff its 16 lanes' j-th contributing cells have distinct slots mod 16.  One thread per (tile, L) run deals its
// non-zeros to the half-warp bins of the records the run occupies anyway (first fit: a bin accepts a non-zero when
// none of its L residues is taken yet); what fits nowhere goes to the free lane where it collides least.  Positions are
// a permutation of the lanes the run already had plus its padding lanes: the stream does not grow, the sums do not
// change (each lane still adds its own contributions in ascending cell order).
constexpr int kPackMaxBins = 160, kPackMaxLen = 24;
__global__ void pack_runs_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ run_start,
                                 const uint32_t* __restrict__ run_tile, const uint32_t* __restrict__ run_len_code, uint32_t nruns,
                                 const uint32_t* __restrict__ contrib_ptr, const uint32_t* __restrict__ contrib_src, uint32_t T,
                                 const uint32_t* __restrict__ tile_cell_ptr, const uint32_t* __restrict__ tile_cells,
                                 const uint32_t* __restrict__ tile_local, int enable, uint32_t* __restrict__ ppos,
                                 unsigned long long* __restrict__ stats /*optional: items, contributions, dense collisions, packed collisions, unplaced*/) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nruns; r += stride) {
    const uint32_t r0 = run_start[r], r1 = run_start[r + 1], ntot = r1 - r0;
    const uint32_t L = run_len_code[r];
    const uint32_t lanes = rec_lanes(L);
    const uint32_t nb_tot = (ntot + lanes - 1u) / lanes * (lanes / 16u);  // half-warp bins of the records the run occupies
    if (!enable || L == 0 || L > uint32_t(kPackMaxLen) || ntot <= 1) {
      for (uint32_t i = r0; i < r1; ++i) ppos[i] = i - r0;
      continue;
    }
    const uint32_t t = run_tile[r];
    const uint32_t cb = tile_cell_ptr[t], ce = tile_cell_ptr[t + 1];
    auto residues = [&](uint32_t q, uint8_t* res) {  // slot mod 16 of every contributing cell of non-zero q
      const uint32_t p0 = contrib_ptr[q];
      for (uint32_t j = 0; j < L; ++j) {
        const uint32_t cell = contrib_src[p0 + j] / T;
        uint32_t lo = cb, hi = ce;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (tile_cells[mid] < cell)
            lo = mid + 1;
          else
            hi = mid;
        }
        res[j] = uint8_t(lo < ce ? (tile_local[lo] & 15u) : 0u);
      }
    };
    uint16_t mask[kPackMaxBins][kPackMaxLen];
    uint8_t cnt[kPackMaxBins];
    uint8_t res[kPackMaxLen];
    // long runs are dealt segment by segment (kPackMaxBins bins each; only the last one owns the run's padding lanes)
    constexpr uint32_t kSeg = uint32_t(kPackMaxBins) * 16u;
    for (uint32_t s0 = 0; s0 < ntot; s0 += kSeg) {
      const uint32_t i0 = r0 + s0, n = min(kSeg, ntot - s0), i1 = i0 + n;
      const uint32_t nb = min(uint32_t(kPackMaxBins), nb_tot - s0 / 16u);
      auto clear = [&]() {
        for (uint32_t b = 0; b < nb; ++b) {
          cnt[b] = 0;
          for (uint32_t j = 0; j < L; ++j) mask[b][j] = 0;
        }
      };
      clear();
      if (stats) {  // collisions of the dense order (lane = position in the run), for comparison
        unsigned long long coll = 0;
        for (uint32_t i = i0; i < i1; ++i) {
          residues(perm[i], res);
          const uint32_t b = (i - i0) / 16u;
          for (uint32_t j = 0; j < L; ++j) {
            coll += (mask[b][j] >> res[j]) & 1u;
            mask[b][j] |= uint16_t(1u << res[j]);
          }
        }
        atomicAdd(stats + 0, (unsigned long long)n);
        atomicAdd(stats + 1, (unsigned long long)n * L);
        atomicAdd(stats + 2, coll);
        clear();
      }
      uint32_t first_open = 0, nleft = 0;
      unsigned long long pcoll = 0;
      for (uint32_t i = i0; i < i1; ++i) {
        residues(perm[i], res);
        while (first_open < nb && cnt[first_open] >= 16) ++first_open;
        uint32_t where = 0xFFFFFFFFu;
        for (uint32_t b = first_open; b < nb; ++b) {
          if (cnt[b] >= 16) continue;
          bool ok = true;
          for (uint32_t j = 0; j < L; ++j) ok = ok && !((mask[b][j] >> res[j]) & 1u);
          if (ok) {
            where = b;
            break;
          }
        }
        if (where == 0xFFFFFFFFu) {
          ppos[i] = 0xFFFFFFFFu;
          ++nleft;
          continue;
        }
        for (uint32_t j = 0; j < L; ++j) mask[where][j] |= uint16_t(1u << res[j]);
        ppos[i] = s0 + where * 16u + cnt[where]++;
      }
      for (uint32_t i = i0; i < i1 && nleft; ++i) {  // the rest: the free lane with the fewest collisions
        if (ppos[i] != 0xFFFFFFFFu) continue;
        residues(perm[i], res);
        uint32_t best = 0xFFFFFFFFu, best_cost = 0xFFFFFFFFu;
        for (uint32_t b = 0; b < nb; ++b) {
          if (cnt[b] >= 16) continue;
          uint32_t cost = 0;
          for (uint32_t j = 0; j < L; ++j) cost += (mask[b][j] >> res[j]) & 1u;
          if (cost < best_cost) best_cost = cost, best = b;
        }
        for (uint32_t j = 0; j < L; ++j) mask[best][j] |= uint16_t(1u << res[j]);
        ppos[i] = s0 + best * 16u + cnt[best]++;
        --nleft;
        pcoll += best_cost;
        if (stats) atomicAdd(stats + 4, 1ull);
      }
      if (stats) atomicAdd(stats + 3, pcoll);
    }
  }
}
__global__ void run_nrec_kernel(const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ run_len_code, uint32_t nruns,
                                uint32_t* __restrict__ nrec) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= nruns; r += stride) {
    if (r == nruns) {
      nrec[r] = 0u;
      continue;
    }
    const uint32_t lanes = rec_lanes(run_len_code[r]);
    nrec[r] = (run_start[r + 1] - run_start[r] + lanes - 1u) / lanes;
  }
}
__global__ void rec_len_kernel(const uint32_t* __restrict__ rec_base, const uint32_t* __restrict__ run_len_code, uint32_t nruns,
                               uint8_t* __restrict__ rec_len) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nruns; r += stride)
    for (uint32_t k = rec_base[r]; k < rec_base[r + 1]; ++k) rec_len[k] = uint8_t(run_len_code[r]);
}
__global__ void gather_u32_kernel(const uint32_t* __restrict__ idx, uint32_t n, const uint32_t* __restrict__ src,
                                  uint32_t* __restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = src[idx[i]];
}

struct TileLayoutBlock {
  const uint32_t* rec_tile_ptr;  // [ntiles+1] first record of every tile
  const uint8_t* rec_len;        // [nrec]
  uint32_t* rec_rel;             // [nrec] out: byte offset / 16 of the record within the tile's stream
};
struct TileLayoutArgs {
  TileLayoutBlock blk[kTileMaxBlocks];
  int nblocks;
};
// One thread per tile packs its records into chunks (a record never straddles a chunk).
// pass 0: chunk count of the tile;  pass 1: record offsets, record and chunk headers written into the stream.
__global__ void tile_layout_kernel(TileLayoutArgs A, uint32_t ntiles, int pass, uint32_t* __restrict__ tile_nchunks,
                                   const uint32_t* __restrict__ tile_chunk_ptr, unsigned char* __restrict__ stream) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += stride) {
    uint32_t off = 0;          // byte offset within the tile's stream; 0 = no chunk opened yet
    uint32_t in_chunk = 0;     // records in the open chunk
    unsigned char* base = pass ? stream + size_t(tile_chunk_ptr[t]) * kChunkBytes : nullptr;
    for (int b = 0; b < A.nblocks; ++b) {
      const TileLayoutBlock& B = A.blk[b];
      for (uint32_t k = B.rec_tile_ptr[t]; k < B.rec_tile_ptr[t + 1]; ++k) {
        const uint32_t L = B.rec_len[k];
        const uint32_t size = rec_bytes(L);
        const uint32_t chunk = off / kChunkBytes;
        if (off % kChunkBytes == 0 || off + size > (chunk + 1) * kChunkBytes) {  // open a new chunk
          if (off % kChunkBytes != 0) {
            if (pass) *reinterpret_cast<uint32_t*>(base + size_t(chunk) * kChunkBytes) = in_chunk;
            off = (chunk + 1) * kChunkBytes;
          }
          off += kChunkHdr;
          in_chunk = 0;
        }
        if (pass) {
          B.rec_rel[k] = off / 16u;
          *reinterpret_cast<uint32_t*>(base + off) = L | (uint32_t(b) << 8) | (rec_lanes(L) << 16);
        }
        off += size;
        ++in_chunk;
      }
    }
    if (off % kChunkBytes != 0) {
      if (pass) *reinterpret_cast<uint32_t*>(base + size_t(off / kChunkBytes) * kChunkBytes) = in_chunk;
      off = (off / kChunkBytes + 1) * kChunkBytes;
    }
    if (!pass) tile_nchunks[t] = off / kChunkBytes;
  }
}
// One thread per (sorted) non-zero: its lane of its record.
__global__ void stream_fill_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ ppos, uint32_t n,
                                   const uint32_t* __restrict__ run_scan,
                                   const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ run_tile,
                                   const uint32_t* __restrict__ run_len_code, const uint32_t* __restrict__ rec_base,
                                   const uint32_t* __restrict__ rec_rel, const uint32_t* __restrict__ tile_chunk_ptr,
                                   const uint32_t* __restrict__ contrib_ptr, const uint32_t* __restrict__ contrib_src, uint32_t T,
                                   int slot_bits, const uint32_t* __restrict__ tile_cell_ptr,
                                   const uint32_t* __restrict__ tile_cells, const uint32_t* __restrict__ tile_local,
                                   const uint8_t* __restrict__ keep,
                                   const uint32_t* __restrict__ pos, int drop,
                                   const uint16_t* __restrict__ direct_map /*stored blocks: slot -> sign | slab offset*/,
                                   unsigned char* __restrict__ stream, int* __restrict__ err) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t q = perm[i];
    const uint32_t r = run_scan[i] - 1;
    const uint32_t idx = ppos[i];  // lane of the run dealt by pack_runs_kernel (i - run_start[r] when packing is off)
    const uint32_t lanes = rec_lanes(run_len_code[r]);
    const uint32_t k = rec_base[r] + idx / lanes, lane = idx % lanes;
    const uint32_t t = run_tile[r];
    unsigned char* rp = stream + size_t(tile_chunk_ptr[t]) * kChunkBytes + size_t(rec_rel[k]) * 16u + kRecHdr;
    reinterpret_cast<uint32_t*>(rp)[lane] = drop ? (keep[q] ? pos[q] + 2u : kNoDest) : q + 2u;
    uint16_t* ent = reinterpret_cast<uint16_t*>(rp + 4u * lanes) + lane;
    const uint32_t p0 = contrib_ptr[q], p1 = contrib_ptr[q + 1];
    const uint32_t cb = tile_cell_ptr[t], ce = tile_cell_ptr[t + 1];
    for (uint32_t p = p0; p < p1 && p - p0 < uint32_t(kMaxLen); ++p) {
      const uint32_t src = contrib_src[p];
      const uint32_t cell = src / T, slot = src - cell * T;
      uint32_t lo = cb, hi = ce;  // first index with tile_cells[idx] >= cell
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tile_cells[mid] < cell)
          lo = mid + 1;
        else
          hi = mid;
      }
      if (lo >= ce || tile_cells[lo] != cell) {
        atomicExch(err, 2);
        continue;
      }
      const uint32_t local = tile_local[lo];
      if (direct_map) {
        const uint32_t code = direct_map[slot];
        if ((code & 0x7FFFu) + local > 0x7FFFu) atomicExch(err, 3);
        ent[(p - p0) * lanes] = uint16_t((code & 0x8000u) | ((code & 0x7FFFu) + local));
      } else {
        if ((local << slot_bits) > 0xFFFFu) atomicExch(err, 3);
        ent[(p - p0) * lanes] = uint16_t((local << slot_bits) | slot);
      }
    }
  }
}

static int bits_for32(uint64_t n) {
  int b = 1;
  while ((1ull << b) < n) ++b;
  return b;
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
static void exclusive_scan_u32(fq_ctx* ctx, const uint32_t* in, uint32_t* out, size_t n) {
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, int64_t(n), ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, int64_t(n), ctx->stream));
  fq_count_launch(ctx, 2);
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
}
static void inclusive_scan_u32(fq_ctx* ctx, const uint32_t* in, uint32_t* out, size_t n) {
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, in, out, int64_t(n), ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, in, out, int64_t(n), ctx->stream));
  fq_count_launch(ctx, 2);
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
}

// Closed-form vertex bricks on a Kuhn grid.
// Closed-form vertex bricks on a Kuhn grid.
void tile_cluster_kuhn(fq_ctx* ctx, fq_mesh* mesh, int dim, const size_t* shape, size_t slab_begin, size_t slab_end_held) {
  if (dim > 3 || std::getenv("FQ_NO_TILE")) return;
  int max_distinct = 1;
  for (const CoreEntryRt& e : g_cores)
    if (e.n == dim && (tile_core_variant() == 1 || e.variant == 0)) max_distinct = std::max(max_distinct, e.ndistinct);
  const int cells_capacity = tile_cells_capacity(max_distinct);
  // brick[a] owned vertices per axis; a tile needs every cell with a vertex in the brick
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
        const uint64_t ncell = cells_touching(cand);
        if (ncell > uint64_t(cells_capacity)) continue;
        const double ratio = double(ncell) / double(owned);
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

// Generic meshes: greedy breadth-first growth of vertex clusters on the host (once
// per mesh), each limited to `capacity` incident cells.  cell_verts is the
// grade-0 FaceIncidence table ([ncells][dim+1], global vertex ids).
void tile_cluster_generic(fq_ctx* ctx, fq_mesh* mesh) {
  const int dim = mesh->dim;
  mesh->cluster_tried = true;
  if (dim > 3 || std::getenv("FQ_NO_TILE") || !mesh->cell_faces[0].p || mesh->edge_lo != 0 || mesh->cell_offset != 0) return;
  std::vector<uint32_t> cell_verts(mesh->cell_faces[0].n);
  FQ_CUDA(cudaMemcpyAsync(cell_verts.data(), mesh->cell_faces[0].p, cell_verts.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  int max_distinct = 1;
  for (const CoreEntryRt& e : g_cores)
    if (e.n == dim && (tile_core_variant() == 1 || e.variant == 0)) max_distinct = std::max(max_distinct, e.ndistinct);
  const uint32_t capacity = uint32_t(tile_cells_capacity(max_distinct));
  const size_t nv = size_t(dim) + 1, ncells = mesh->ncells, V = mesh->nsimplices[0];
  if (V == 0 || ncells == 0 || V >= (size_t(1) << 32)) return;
  // vertex -> cells incidence (CSR)
  std::vector<uint32_t> vptr(V + 1, 0);
  for (size_t i = 0; i < ncells * nv; ++i) {
    if (cell_verts[i] >= V) return;  // malformed table: leave the mesh unclustered (slab path)
    vptr[size_t(cell_verts[i]) + 1] += 1;
  }
  for (size_t v = 0; v < V; ++v) vptr[v + 1] += vptr[v];
  std::vector<uint32_t> vcells(ncells * nv), fill(vptr.begin(), vptr.end() - 1);
  for (size_t c = 0; c < ncells; ++c)
    for (size_t j = 0; j < nv; ++j) vcells[fill[size_t(cell_verts[c * nv + j])]++] = uint32_t(c);
  const uint32_t kNone = 0xFFFFFFFFu;
  std::vector<uint32_t> vtile(V, kNone), stamp(ncells, kNone), queue;
  uint32_t ntiles = 0;
  for (size_t seed = 0; seed < V; ++seed) {
    if (vtile[seed] != kNone) continue;
    const uint32_t T = ntiles++;
    uint32_t ncells_T = 0;
    queue.clear();
    queue.push_back(uint32_t(seed));
    for (size_t head = 0; head < queue.size(); ++head) {
      const uint32_t v = queue[head];
      if (vtile[v] != kNone) continue;
      uint32_t fresh = 0;
      for (uint32_t p = vptr[v]; p < vptr[v + 1]; ++p) fresh += stamp[vcells[p]] != T;
      if (ncells_T > 0 && ncells_T + fresh > capacity) continue;  // does not fit: left for a later tile
      vtile[v] = T;
      ncells_T += fresh;
      for (uint32_t p = vptr[v]; p < vptr[v + 1]; ++p) {
        const uint32_t c = vcells[p];
        if (stamp[c] == T) continue;
        stamp[c] = T;
        for (size_t j = 0; j < nv; ++j) {
          const uint32_t w = uint32_t(cell_verts[size_t(c) * nv + j]);
          if (vtile[w] == kNone) queue.push_back(w);
        }
      }
      if (ncells_T >= capacity) break;
    }
  }
  mesh->vertex_tile.alloc(V);
  mesh->vtile_lo = 0;
  mesh->ntiles = ntiles;
  FQ_CUDA(cudaMemcpyAsync(mesh->vertex_tile.p, vtile.data(), V * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
}


static void radix_sort_pairs_u64(fq_ctx* ctx, DevBuf<uint64_t>& keys, DevBuf<uint32_t>& vals, size_t n, int end_bit) {
  DevBuf<uint64_t> keys_alt(n ? n : 1);
  DevBuf<uint32_t> vals_alt(n ? n : 1);
  cub::DoubleBuffer<uint64_t> dk(keys.p, keys_alt.p);
  cub::DoubleBuffer<uint32_t> dv(vals.p, vals_alt.p);
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, int64_t(n), 0, end_bit, ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, int64_t(n), 0, end_bit, ctx->stream));
  fq_count_launch(ctx, (end_bit + 7) / 8 + 1);
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (dk.Current() != keys.p) std::swap(keys, keys_alt);
  if (dv.Current() != vals.p) std::swap(vals, vals_alt);
}

// Per-block intermediate data of the plan build (kept until the stream is filled).
struct BlockBuild {
  DevBuf<uint32_t> perm, ppos, run_scan, run_start, run_tile, run_len, rec_base, rec_tile_ptr, rec_rel;
  DevBuf<uint8_t> rec_len;
  uint32_t nruns = 0, nrec = 0;
};

// Picks the core (grade, variant) that can serve every block with the fewest gather terms.
static const CoreEntryRt* choose_core(int dim, fq_csr* const* csrs, int nblocks) {
  const int pref = tile_core_variant();
  const CoreEntryRt* best = nullptr;
  long best_score = 0;
  for (const CoreEntryRt& core : g_cores) {
    if (core.n != dim) continue;
    if (pref == 0 && core.variant != 0) continue;
    long score = 0;
    bool ok = true;
    for (int b = 0; ok && b < nblocks; ++b) {
      int no, ni;
      std::vector<uint8_t> codes;
      ok = build_recipes(dim, core.k, core, csrs[b]->kind, csrs[b]->grade, no, ni, codes);
      for (uint8_t c : codes) ok = ok && c != 0xFF;
      score += long(no) * ni * 1000;
    }
    if (!ok) continue;
    score += core.ndistinct;
    if (!best || score < best_score) best = &core, best_score = score;
  }
  return best;
}

// Builds the tile plan for the fused blocks `csrs` (their structural phase and,
// when dropping, their cached classification keep/pos must be valid).
// Returns nullptr when the tile path does not apply (falls back to the slab path).
std::shared_ptr<TilePlan> tile_plan_build(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* const* csrs, int nblocks, bool drop) {
  if (std::getenv("FQ_NO_TILE")) return nullptr;
  if (nblocks < 1 || nblocks > kTileMaxBlocks) return nullptr;
  if (!mesh->vertex_tile.p || mesh->ntiles == 0 || !mesh->cell_faces[0].p) return nullptr;
  const int dim = mesh->dim;
  if (dim > 3) return nullptr;
  for (int b = 0; b < nblocks; ++b)
    if (csrs[b]->kind == KIND_LUMPED) return nullptr;
  const CoreEntryRt* core = choose_core(dim, csrs, nblocks);
  if (!core) return nullptr;
  const int kc = core->k;
  const TileConfig cfg = tile_config();
  auto plan = std::make_shared<TilePlan>();
  plan->mesh = mesh;
  plan->dim = dim;
  plan->core_k = kc;
  plan->ndistinct = core->ndistinct;
  plan->nblocks = nblocks;
  plan->launch = core->launch;
  plan->ntiles = uint32_t(mesh->ntiles);
  plan->nthreads = cfg.nthreads;
  plan->ws = cfg.ws;
  plan->stream_warps = cfg.stream_warps;
  if (plan->ntiles >= (1u << 27)) return nullptr;
  // recipes (8-bit codes over the distinct values; widened to slab offsets once the slab stride is known)
  std::vector<std::vector<uint8_t>> codes8(static_cast<size_t>(nblocks));
  size_t recipe_u16 = 0;
  for (int b = 0; b < nblocks; ++b) {
    TileBlockPlan& bp = plan->blk[b];
    std::vector<uint8_t>& codes = codes8[size_t(b)];
    if (!build_recipes(dim, kc, *core, csrs[b]->kind, csrs[b]->grade, bp.no, bp.ni, codes)) return nullptr;
    const bool shape_ok = (bp.no == 0 && bp.ni == 0) || (bp.no == 1 && bp.ni >= 1 && bp.ni <= 4) ||
                          (bp.no == bp.ni && bp.no >= 2 && bp.no <= 4);
    if (!shape_ok) return nullptr;
    if (bp.no * bp.ni > 1) recipe_u16 += size_t(csrs[b]->el_rows * csrs[b]->el_cols) * size_t((bp.no * bp.ni + 3) / 4 * 4);
    bp.csr = csrs[b];
    bp.nnz_at_build = csrs[b]->nnz;
    bp.dropped_at_build = drop;
    const uint32_t T = uint32_t(csrs[b]->el_rows * csrs[b]->el_cols);
    bp.slot_bits = bits_for32(T ? T : 1);
  }
  const size_t recipe_bytes = (recipe_u16 * 2 + 15) / 16 * 16;
  if (recipe_bytes > 2048) return nullptr;
  // alternating kernel: every block must read one half of the slab only, second-half blocks last in the stream
  if (cfg.alt && core->split_ok && cfg.nthreads == 512) {
    bool ok = true;
    uint32_t ymask = 0;
    for (int b = 0; b < nblocks && ok; ++b) {
      bool any1 = false, any2 = false;
      for (uint8_t c : codes8[size_t(b)]) {
        if (c == 0xFF) continue;
        const int slot = c & 0x7F;
        (((core->half2[slot >> 6] >> (slot & 63)) & 1ull) ? any2 : any1) = true;
      }
      if (any1 && any2) ok = false;
      if (any2) ymask |= 1u << b;
    }
    for (int b = 0; b + 1 < nblocks; ++b)
      if (((ymask >> b) & 1u) && !((ymask >> (b + 1)) & 1u)) ok = false;  // a first-half block after a second-half one
    plan->alt = ok;
    plan->yblock_mask = ok ? ymask : 0u;
  }
  const int block = 256;
  const int nv = dim + 1;
  const int ne = int(binom(dim + 1, 2));
  const size_t ncells = mesh->ncells;
  const uint32_t v_lo = uint32_t(mesh->vtile_lo);
  // ---- tile cell lists
  DevBuf<uint32_t> tile_cells, tile_local;  // cells of every tile sorted by id, and their slab slot
  {
    const size_t nkeys = ncells * size_t(nv);
    DevBuf<uint64_t> keys(nkeys), keys_alt(nkeys);
    tile_cell_keys_kernel<<<grid_for(ncells, block, ctx->sm_count), block, 0, ctx->stream>>>(
        mesh->cell_faces[0].p, nv, ncells, mesh->vertex_tile.p, v_lo, keys.p);
    fq_count_launch(ctx);
    cub::DoubleBuffer<uint64_t> dk(keys.p, keys_alt.p);
    size_t tmp_bytes = 0;
    const int end_bit = 64;
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
    tile_cells.alloc(nvalid ? nvalid : 1);
    split_keys_kernel<<<grid_for(nvalid, block, ctx->sm_count), block, 0, ctx->stream>>>(dk.Current(), nvalid, tile_of.p,
                                                                                       tile_cells.p);
    plan->tile_cell_ptr.alloc(size_t(plan->ntiles) + 1);
    seg_ptr_kernel<<<grid_for(size_t(nvalid) + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(
        tile_of.p, nvalid, plan->ntiles, plan->tile_cell_ptr.p);
    // slab slots: type-major inside a tile when the cell numbering is (box, type), else ascending cell id
    DevBuf<uint32_t> cells_by_slot(nvalid ? nvalid : 1);
    tile_local.alloc(nvalid ? nvalid : 1);
    {
      DevBuf<uint64_t> skeys(nvalid ? nvalid : 1);
      DevBuf<uint32_t> spos(nvalid ? nvalid : 1);
      const uint32_t period = std::getenv("FQ_TILE_NO_TYPE_MAJOR") ? 0u : uint32_t(mesh->cell_type_period);
      tile_slot_keys_kernel<<<grid_for(nvalid, block, ctx->sm_count), block, 0, ctx->stream>>>(
          tile_of.p, tile_cells.p, nvalid, period, uint64_t(mesh->cell_offset), skeys.p, spos.p);
      radix_sort_pairs_u64(ctx, skeys, spos, nvalid, 64);
      tile_slots_kernel<<<grid_for(nvalid, block, ctx->sm_count), block, 0, ctx->stream>>>(
          skeys.p, spos.p, nvalid, plan->tile_cell_ptr.p, cells_by_slot.p, tile_local.p);
    }
    plan->tile_cell_edges.alloc(size_t(nvalid ? nvalid : 1) * size_t(ne));
    tile_cell_edges_kernel<<<grid_for(size_t(nvalid) * ne, block, ctx->sm_count), block, 0, ctx->stream>>>(
        cells_by_slot.p, nvalid, mesh->cell_faces[1].p, ne, plan->tile_cell_edges.p);
    fq_count_launch(ctx, 5);
    FQ_CUDA(cudaGetLastError());
    std::vector<uint32_t> h_ptr(size_t(plan->ntiles) + 1);
    FQ_CUDA(cudaMemcpyAsync(h_ptr.data(), plan->tile_cell_ptr.p, h_ptr.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                            ctx->stream));
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
    uint32_t max_cells = 1;
    for (uint32_t t = 0; t < plan->ntiles; ++t) max_cells = std::max(max_cells, h_ptr[t + 1] - h_ptr[t]);
    if (int(max_cells) > tile_cells_capacity(plan->ndistinct)) return nullptr;
    plan->cstride = int(max_cells) | 1;  // odd stride: distinct rows start on different banks
    {
      // bank-aware lane packing wants every load of a contribution on the bank pair of its cell: stride = 0 mod 16
      const char* e = std::getenv("FQ_TILE_PACK");
      const int cs16 = (int(max_cells) + 15) / 16 * 16;
      plan->pack = !(e && e[0] == '0') && !cfg.ws && cs16 <= tile_cells_capacity(plan->ndistinct) + 15 &&
                   size_t(cs16) * size_t(plan->ndistinct) * sizeof(double) + tile_fixed_smem(cfg) <= cfg.smem_cta;
      if (plan->pack) plan->cstride = cs16;
    }
    // dynamic shared memory layout
    const size_t nwarps = plan->stream_warps ? size_t(plan->stream_warps)
                                             : (plan->ws ? size_t(kWsConsumerWarps) : size_t(plan->nthreads) / 32);
    size_t off = size_t(plan->cstride) * size_t(plan->ndistinct) * sizeof(double);
    off = (off + 127) / 128 * 128;
    plan->slab_bytes = uint32_t(off);
    if (plan->ws) off *= 2;
    plan->ring_off = uint32_t(off);
    off += nwarps * kSlotsPerWarp * kChunkBytes;
    plan->rec_off = uint32_t(off);
    off += recipe_bytes;
    off = (off + 7) / 8 * 8;
    plan->mbar_off = uint32_t(off);
    off += (nwarps * kSlotsPerWarp + 8) * 8;
    plan->smem_bytes = off;
    if (plan->smem_bytes > cfg.smem_cta) return nullptr;
  }
  // ---- recipes as slab offsets: code16 = sign << 15 | distinct * cstride
  if (size_t(plan->ndistinct) * size_t(plan->cstride) > 0x8000u) return nullptr;
  std::vector<DevBuf<uint16_t>> direct_map(static_cast<size_t>(nblocks));  // stored blocks: slot -> code16, used by the stream fill
  {
    std::vector<uint16_t> table(recipe_bytes / 2, 0);
    size_t off16 = 0;
    for (int b = 0; b < nblocks; ++b) {
      TileBlockPlan& bp = plan->blk[b];
      const std::vector<uint8_t>& codes = codes8[size_t(b)];
      const int nterms = bp.no * bp.ni;
      const size_t nslots = size_t(csrs[b]->el_rows * csrs[b]->el_cols);
      auto widen = [&](uint8_t c) { return uint16_t(((c & 0x80u) << 8) | uint32_t((c & 0x7Fu) * plan->cstride)); };
      bp.recipe_off = int(off16);
      if (nterms == 1) {
        std::vector<uint16_t> dm(nslots);
        for (size_t sl = 0; sl < nslots; ++sl) dm[sl] = widen(codes[sl]);
        upload_vec(direct_map[size_t(b)], dm);
      } else if (nterms > 1) {
        const int nt4 = (nterms + 3) / 4 * 4;
        for (size_t sl = 0; sl < nslots; ++sl)
          for (int q = 0; q < nterms; ++q) table[off16 + sl * nt4 + q] = widen(codes[sl * nterms + q]);
        off16 += nslots * size_t(nt4);
      }
    }
    std::vector<uint8_t> bytes(recipe_bytes ? recipe_bytes : 16, 0);
    if (recipe_bytes) std::memcpy(bytes.data(), table.data(), recipe_bytes);
    upload_vec(plan->recipes, bytes);
    plan->recipe_bytes = int(recipe_bytes);
  }
  // ---- per block: non-zeros sorted by (tile, L, signature), cut into warp records
  // orderings inside a (tile, L) run — measured neutral on B200 (the gathers stay scattered), off by default
  const int use_sig = std::getenv("FQ_TILE_SIG") ? std::atoi(std::getenv("FQ_TILE_SIG")) : 0;
  const bool use_bank = std::getenv("FQ_TILE_BANK") ? std::atoi(std::getenv("FQ_TILE_BANK")) != 0 : false;
  DevBuf<int> d_err(1);
  FQ_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), ctx->stream));
  std::vector<BlockBuild> bb(static_cast<size_t>(nblocks));
  const bool pack_stats = std::getenv("FQ_TILE_PACK_STATS") != nullptr;  // development aid: packing quality per block
  DevBuf<unsigned long long> d_pack_stats(8);
  FQ_CUDA(cudaMemsetAsync(d_pack_stats.p, 0, d_pack_stats.bytes(), ctx->stream));
  for (int b = 0; b < nblocks; ++b) {
    fq_csr* csr = csrs[b];
    BlockBuild& B = bb[size_t(b)];
    const size_t s_nnz = csr->s_nnz;
    const size_t nrows_local = csr->row_end - csr->row_begin;
    B.rec_tile_ptr.alloc(size_t(plan->ntiles) + 1);
    if (s_nnz == 0) {
      FQ_CUDA(cudaMemsetAsync(B.rec_tile_ptr.p, 0, B.rec_tile_ptr.bytes(), ctx->stream));
      B.rec_len.alloc(1);
      B.rec_rel.alloc(1);
      continue;
    }
    int tg, rg;
    kind_grades(csr->kind, csr->grade, tg, rg);
    const int nt = nlocal(dim, tg);
    const uint32_t T = uint32_t(csr->el_rows * csr->el_cols);
    std::vector<uint8_t> top_pos;  // top position of every local face of the test grade
    for (uint32_t m : colex_subsets(dim + 1, tg + 1)) top_pos.push_back(uint8_t(mask_elems(m).back()));
    DevBuf<uint8_t> d_top;
    upload_vec(d_top, top_pos);
    DevBuf<uint32_t> row_tile(nrows_local ? nrows_local : 1);
    FQ_CUDA(cudaMemsetAsync(row_tile.p, 0, row_tile.bytes(), ctx->stream));
    row_tile_kernel<<<grid_for(ncells * size_t(nt), block, ctx->sm_count), block, 0, ctx->stream>>>(
        mesh->cell_faces[size_t(tg)].p, nt, mesh->cell_faces[0].p, nv, d_top.p, ncells, mesh->vertex_tile.p, v_lo,
        uint32_t(csr->row_begin), uint32_t(csr->row_end), row_tile.p);
    DevBuf<uint64_t> key(s_nnz);
    B.perm.alloc(s_nnz);
    nnz_key_kernel<<<grid_for(nrows_local, block, ctx->sm_count), block, 0, ctx->stream>>>(
        csr->s_row_ptr.p, uint32_t(nrows_local), row_tile.p, csr->contrib_ptr.p, csr->contrib_src.p, T, use_sig,
        use_bank ? plan->tile_cell_ptr.p : nullptr, tile_cells.p, use_bank ? tile_local.p : nullptr, key.p, B.perm.p, d_err.p);
    fq_count_launch(ctx, 2);
    FQ_CUDA(cudaGetLastError());
    radix_sort_pairs_u64(ctx, key, B.perm, s_nnz, 36 + bits_for32(plan->ntiles));
    if (use_bank) {
      DevBuf<uint32_t> gstart(s_nnz), gscan(s_nnz);
      group_start_kernel<<<grid_for(s_nnz, block, ctx->sm_count), block, 0, ctx->stream>>>(key.p, uint32_t(s_nnz), gstart.p);
      size_t tmp_bytes = 0;
      FQ_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tmp_bytes, gstart.p, gscan.p, cub::Max(), int64_t(s_nnz), ctx->stream));
      DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
      FQ_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tmp_bytes, gstart.p, gscan.p, cub::Max(), int64_t(s_nnz), ctx->stream));
      occurrence_kernel<<<grid_for(s_nnz, block, ctx->sm_count), block, 0, ctx->stream>>>(key.p, uint32_t(s_nnz), gscan.p);
      fq_count_launch(ctx, 4);
      FQ_CUDA(cudaStreamSynchronize(ctx->stream));
      radix_sort_pairs_u64(ctx, key, B.perm, s_nnz, 36 + bits_for32(plan->ntiles));
    }
    DevBuf<uint32_t> head(s_nnz);
    B.run_scan.alloc(s_nnz);
    run_heads_kernel<<<grid_for(s_nnz, block, ctx->sm_count), block, 0, ctx->stream>>>(key.p, uint32_t(s_nnz), head.p);
    inclusive_scan_u32(ctx, head.p, B.run_scan.p, s_nnz);
    B.nruns = read_u32(ctx, B.run_scan.p + (s_nnz - 1));
    B.run_start.alloc(size_t(B.nruns) + 1);
    B.run_tile.alloc(B.nruns);
    B.run_len.alloc(B.nruns);
    run_info_kernel<<<grid_for(s_nnz, block, ctx->sm_count), block, 0, ctx->stream>>>(
        key.p, head.p, B.run_scan.p, uint32_t(s_nnz), B.run_start.p, B.run_tile.p, B.run_len.p);
    const uint32_t n32 = uint32_t(s_nnz);
    FQ_CUDA(cudaMemcpyAsync(B.run_start.p + B.nruns, &n32, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    B.ppos.alloc(s_nnz);
    pack_runs_kernel<<<grid_for(B.nruns, 64, ctx->sm_count), 64, 0, ctx->stream>>>(
        B.perm.p, B.run_start.p, B.run_tile.p, B.run_len.p, B.nruns, csr->contrib_ptr.p, csr->contrib_src.p, T,
        plan->tile_cell_ptr.p, tile_cells.p, tile_local.p, plan->pack ? 1 : 0, B.ppos.p, pack_stats ? d_pack_stats.p : nullptr);
    fq_count_launch(ctx);
    if (pack_stats) {
      unsigned long long h[5];
      FQ_CUDA(cudaMemcpyAsync(h, d_pack_stats.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
      FQ_CUDA(cudaStreamSynchronize(ctx->stream));
      std::fprintf(stderr, "[tile pack] block %d: %llu non-zeros, %llu contributions, collisions dense %llu -> packed %llu (%llu unplaced), stride %d\n",
                   b, h[0], h[1], h[2], h[3], h[4], plan->cstride);
      FQ_CUDA(cudaMemsetAsync(d_pack_stats.p, 0, sizeof h, ctx->stream));
    }
    DevBuf<uint32_t> nrec_run(size_t(B.nruns) + 1);
    B.rec_base.alloc(size_t(B.nruns) + 1);
    run_nrec_kernel<<<grid_for(size_t(B.nruns) + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(
        B.run_start.p, B.run_len.p, B.nruns, nrec_run.p);
    exclusive_scan_u32(ctx, nrec_run.p, B.rec_base.p, size_t(B.nruns) + 1);
    B.nrec = read_u32(ctx, B.rec_base.p + B.nruns);
    B.rec_len.alloc(B.nrec ? B.nrec : 1);
    B.rec_rel.alloc(B.nrec ? B.nrec : 1);
    rec_len_kernel<<<grid_for(B.nruns, block, ctx->sm_count), block, 0, ctx->stream>>>(B.rec_base.p, B.run_len.p, B.nruns,
                                                                                      B.rec_len.p);
    // first record of every tile: first run of the tile -> its record base
    DevBuf<uint32_t> first_run(size_t(plan->ntiles) + 1);
    seg_ptr_kernel<<<grid_for(size_t(B.nruns) + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(B.run_tile.p, B.nruns,
                                                                                                 plan->ntiles, first_run.p);
    gather_u32_kernel<<<grid_for(size_t(plan->ntiles) + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(
        first_run.p, plan->ntiles + 1, B.rec_base.p, B.rec_tile_ptr.p);
    fq_count_launch(ctx, 6);
    FQ_CUDA(cudaGetLastError());
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  {
    int h_err = 0;
    FQ_CUDA(cudaMemcpy(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (h_err) return nullptr;  // a non-zero with more contributions than a record holds: keep the slab path
  }
  // ---- layout of the per-tile streams
  TileLayoutArgs la{};
  la.nblocks = nblocks;
  for (int b = 0; b < nblocks; ++b)
    la.blk[b] = TileLayoutBlock{bb[size_t(b)].rec_tile_ptr.p, bb[size_t(b)].rec_len.p, bb[size_t(b)].rec_rel.p};
  DevBuf<uint32_t> tile_nchunks(size_t(plan->ntiles) + 1);
  FQ_CUDA(cudaMemsetAsync(tile_nchunks.p, 0, tile_nchunks.bytes(), ctx->stream));
  tile_layout_kernel<<<grid_for(plan->ntiles, block, ctx->sm_count), block, 0, ctx->stream>>>(la, plan->ntiles, 0,
                                                                                             tile_nchunks.p, nullptr, nullptr);
  fq_count_launch(ctx);
  plan->tile_chunk_ptr.alloc(size_t(plan->ntiles) + 1);
  {
    std::vector<uint32_t> h_n(size_t(plan->ntiles) + 1);
    FQ_CUDA(cudaMemcpyAsync(h_n.data(), tile_nchunks.p, h_n.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t total = 0;
    for (uint32_t t = 0; t < plan->ntiles; ++t) total += h_n[t];
    if (total >= (1ull << 31)) return nullptr;
  }
  exclusive_scan_u32(ctx, tile_nchunks.p, plan->tile_chunk_ptr.p, size_t(plan->ntiles) + 1);
  const uint32_t total_chunks = read_u32(ctx, plan->tile_chunk_ptr.p + plan->ntiles);
  plan->stream.alloc(size_t(total_chunks ? total_chunks : 1) * kChunkBytes);
  FQ_CUDA(cudaMemsetAsync(plan->stream.p, 0, plan->stream.bytes(), ctx->stream));  // padding lanes: dest 0, entries 0
  tile_layout_kernel<<<grid_for(plan->ntiles, block, ctx->sm_count), block, 0, ctx->stream>>>(
      la, plan->ntiles, 1, nullptr, plan->tile_chunk_ptr.p, plan->stream.p);
  fq_count_launch(ctx);
  for (int b = 0; b < nblocks; ++b) {
    fq_csr* csr = csrs[b];
    BlockBuild& B = bb[size_t(b)];
    if (csr->s_nnz == 0) continue;
    const uint32_t T = uint32_t(csr->el_rows * csr->el_cols);
    stream_fill_kernel<<<grid_for(csr->s_nnz, block, ctx->sm_count), block, 0, ctx->stream>>>(
        B.perm.p, B.ppos.p, uint32_t(csr->s_nnz), B.run_scan.p, B.run_start.p, B.run_tile.p, B.run_len.p, B.rec_base.p, B.rec_rel.p,
        plan->tile_chunk_ptr.p, csr->contrib_ptr.p, csr->contrib_src.p, T, plan->blk[b].slot_bits, plan->tile_cell_ptr.p,
        tile_cells.p, tile_local.p, csr->keep.p, drop ? csr->pos.p : nullptr, drop ? 1 : 0,
        plan->blk[b].no * plan->blk[b].ni == 1 ? direct_map[size_t(b)].p : nullptr, plan->stream.p, d_err.p);
    fq_count_launch(ctx);
  }
  FQ_CUDA(cudaGetLastError());
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  int h_err = 0;
  FQ_CUDA(cudaMemcpy(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (h_err) return nullptr;  // a tile exceeds the 15/16-bit local index space: keep the slab path
  plan->changed.alloc(1);
  plan->ticket.alloc(1);
  plan->stats.alloc(8);
  plan->grid = ctx->sm_count * (cfg.nthreads == 256 ? 2 : 1);
  return plan;
}

// Runs the fused kernel.  Returns false when the zero/non-zero classification of
// some entry differs from the plan's (the caller re-runs the slab path and rebuilds).
bool tile_assemble(fq_ctx* ctx, const fq_mesh* mesh, TilePlan& plan) {
  TileParams P{};
  P.tile_cell_ptr = plan.tile_cell_ptr.p;
  P.tile_cell_edges = plan.tile_cell_edges.p;
  P.lengths = mesh->lengths.p;
  P.edge_lo = uint32_t(mesh->edge_lo);
  P.ntiles = plan.ntiles;
  P.tile_chunk_ptr = plan.tile_chunk_ptr.p;
  P.stream = plan.stream.p;
  P.cstride = plan.cstride;
  P.nblocks = plan.nblocks;
  P.ring_off = plan.ring_off;
  P.rec_off = plan.rec_off;
  P.mbar_off = plan.mbar_off;
  P.slab_bytes = plan.slab_bytes;
  P.recipes = plan.recipes.p;
  P.recipe_bytes = plan.recipe_bytes;
  P.debug = std::getenv("FQ_TILE_DEBUG") ? std::atoi(std::getenv("FQ_TILE_DEBUG")) : 0;
  P.check_classification = (plan.blk[0].dropped_at_build && !(P.debug & 7)) ? 1 : 0;
  P.yblock_mask = plan.yblock_mask;
  P.chunk_rotation = std::getenv("FQ_TILE_ROTATE") ? uint32_t(std::atoi(std::getenv("FQ_TILE_ROTATE"))) : 5u;
  P.changed = plan.changed.p;
  P.ticket = plan.ticket.p;
  P.stats = plan.stats.p;
  for (int b = 0; b < plan.nblocks; ++b) {
    const TileBlockPlan& bp = plan.blk[b];
    TileBlockDev& d = P.blk[b];
    d.values = bp.csr->values.p;
    d.no = bp.no;
    d.ni = bp.ni;
    d.recipe_off = bp.recipe_off;
    d.slot_bits = bp.slot_bits;
  }
  {
    ScopedSpan span(ctx, "k13_tile_fused");
    FQ_CUDA(cudaMemsetAsync(plan.changed.p, 0, sizeof(int), ctx->stream));
    FQ_CUDA(cudaMemsetAsync(plan.ticket.p, 0, sizeof(unsigned int), ctx->stream));
    if (P.debug & 8) FQ_CUDA(cudaMemsetAsync(plan.stats.p, 0, 8 * sizeof(unsigned long long), ctx->stream));
    plan.launch(ctx, plan, P);
  }
  int changed = 0;
  FQ_CUDA(cudaMemcpyAsync(&changed, plan.changed.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (P.debug & 8) {  // development aid: where the warps spend their cycles
    unsigned long long h[6];
    FQ_CUDA(cudaMemcpy(h, plan.stats.p, sizeof h, cudaMemcpyDeviceToHost));
    const double tot = double(h[0] + h[1] + h[2] + h[3] + h[4] + h[5]);
    std::fprintf(stderr, "[tile stats] k1 %.1f%%  bar_k1 %.1f%%  chunk_wait %.1f%%  records %.1f%%  bar_top %.1f%%  other %.1f%%  (warp-cycles %.3g)\n",
                 100 * h[0] / tot, 100 * h[1] / tot, 100 * h[2] / tot, 100 * h[3] / tot, 100 * h[4] / tot, 100 * h[5] / tot, tot);
  }
  return changed == 0;
}

bool tile_plan_matches(const TilePlan& plan, const fq_mesh* mesh, fq_csr* const* csrs, int nblocks, bool drop) {
  if (plan.mesh != mesh || plan.nblocks != nblocks) return false;
  for (int b = 0; b < nblocks; ++b) {
    const TileBlockPlan& bp = plan.blk[b];
    if (bp.csr != csrs[b] || bp.dropped_at_build != drop || bp.nnz_at_build != csrs[b]->nnz) return false;
  }
  return true;
}

int64_t tile_plan_bytes(const TilePlan& plan) {
  return int64_t(plan.tile_cell_ptr.bytes() + plan.tile_cell_edges.bytes() + plan.tile_chunk_ptr.bytes() + plan.stream.bytes());
}

}  // namespace fq
