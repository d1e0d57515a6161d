// tile.cu — tile-fused numeric assembly: K1 (element masses) and K3 (segmented
// reduction into CSR) in ONE persistent kernel, with the element data of a tile
// living only in shared memory.  The element slab of the two-kernel path
// (8*T bytes per cell written and read back through HBM) disappears.
//
// Decomposition = the multi-GPU one, repeated at CTA level: *owner computes*.
//   * vertices are clustered into tiles (closed-form bricks on Kuhn grids);
//   * a tile owns the rows (simplices) whose top vertex it contains, hence
//     whole CSR rows, hence every structural non-zero of those rows;
//   * it evaluates the element masses of ALL cells touching its vertices
//     (owned + halo cells, recomputed by the neighbouring tiles — FP64 work is
//     cheap here, HBM traffic is not) into shared memory, then reduces each of
//     its non-zeros over the contributing (cell, slot) pairs in ascending cell
//     order — the same order as the slab path, so values are bit-identical.
//
// Shared memory holds only the DISTINCT values of the masses M_{k-1}, M_k,
// M_{k+1} of a cell (54 doubles for the 3-D k = 1 Hodge blocks instead of 112
// element entries); the sandwiches d*M*D (operators.rs:201-211) are evaluated by
// the gather through per-slot "recipes" — signed sums of mass entries in the
// reference's k-ascending gemm order, exact because the incidence entries are
// 0/+-1 (tape.hpp evaluates the same products symbolically).
//
// Reference path replaced: formoniq/src/galerkin.rs:138-188 (assemble_matrix)
// + hodge.rs:62-72 (the four HodgeBlocks), numeric phase.
#include <cub/cub.cuh>

#include <cstdlib>

#include "elmat_gen.cuh"
#include "internal.hpp"
#include "kuhn.hpp"

namespace fq {

constexpr int kTileThreads = 512;
constexpr int kTileMaxBlocks = 4;
constexpr uint32_t kNoDest = 0xFFFFFFFFu;

struct TileBlockDev {
  const uint32_t* tile_nnz_ptr;  // [ntiles+1] into the tile-ordered nnz arrays
  const uint32_t* tile_con_ptr;  // [ntiles+1] into con_src
  const uint32_t* nnz_dest;      // [s_nnz] position in csr->values, kNoDest = dropped (must stay all-zero)
  const uint16_t* nnz_end;       // [s_nnz] end of the nnz's contributions, relative to tile_con_ptr[tile]
  const uint16_t* con_src;       // [ncontrib] (local cell << slot_bits) | slot
  double* values;
  int no, ni;                    // recipe shape: outer x inner signed terms per slot
  int recipe_off;                // byte offset of this block's recipes (nslots * no * ni codes)
  int slot_bits;
};

struct TileParams {
  const uint32_t* tile_cell_ptr;  // [ntiles+1]
  const uint32_t* tile_cells;     // local cell ids, ascending within a tile
  const uint32_t* cell_edges;
  const double* lengths;
  uint32_t edge_lo;
  uint32_t ntiles;
  int cstride;                    // cells capacity of the shared slab
  int nblocks;
  const uint8_t* recipes;         // code = distinct slot | 0x80 negated; 0xFF = no term
  int recipe_bytes;
  int check_classification;       // 1 when the plan carries the reference's value-dependent pattern
  int* changed;                   // raised when the zero/non-zero classification differs from the plan's
  unsigned int* ticket;           // dynamic tile scheduler
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

template <class Fn, int NE>
__global__ void __launch_bounds__(kTileThreads, 1) tile_assemble_kernel(Fn fn, const __grid_constant__ TileParams P) {
  extern __shared__ double smem[];
  double* slab = smem;
  uint8_t* rec = reinterpret_cast<uint8_t*>(smem + size_t(P.cstride) * Fn::kDistinct);
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_hdr[2 + 3 * kTileMaxBlocks];
  for (int i = threadIdx.x; i < P.recipe_bytes; i += kTileThreads) rec[i] = P.recipes[i];
  for (;;) {
    __syncthreads();  // previous tile fully consumed (and recipes visible)
    if (threadIdx.x == 0) s_tile = atomicAdd(P.ticket, 1u);
    __syncthreads();
    const uint32_t t = s_tile;
    if (t >= P.ntiles) break;
    if (threadIdx.x < 2) s_hdr[threadIdx.x] = P.tile_cell_ptr[t + threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 32 + 3 * P.nblocks) {
      const int j = threadIdx.x - 32, b = j / 3, w = j - 3 * b;
      s_hdr[2 + j] = w == 0 ? P.blk[b].tile_nnz_ptr[t] : (w == 1 ? P.blk[b].tile_nnz_ptr[t + 1] : P.blk[b].tile_con_ptr[t]);
    }
    __syncthreads();
    const uint32_t cbase = s_hdr[0], nc = s_hdr[1] - s_hdr[0];
    uint32_t work = 0;
    for (int b = 0; b < P.nblocks; ++b) work += s_hdr[2 + 3 * b + 1] - s_hdr[2 + 3 * b];
    if (work == 0) continue;
    // ---- K1: element masses of the tile's cells -> shared slab [distinct][cell]
    for (uint32_t c = threadIdx.x; c < nc; c += kTileThreads) {
      const uint32_t cell = __ldg(P.tile_cells + cbase + c);
      const uint32_t* ce = P.cell_edges + size_t(cell) * NE;
      uint32_t eid[NE > 0 ? NE : 1];
#pragma unroll
      for (int e = 0; e < NE; ++e) eid[e] = __ldg(ce + e);
      double s[NE > 0 ? NE : 1];
#pragma unroll
      for (int e = 0; e < NE; ++e) s[e] = __ldg(P.lengths + (eid[e] - P.edge_lo));
      TileSink sink{slab + c, P.cstride};
      fn(s, sink);
    }
    __syncthreads();
    // ---- K3: one thread per owned structural non-zero, contributions in ascending cell order
    for (int b = 0; b < P.nblocks; ++b) {
      const TileBlockDev& B = P.blk[b];
      const uint32_t n0 = s_hdr[2 + 3 * b], n1 = s_hdr[2 + 3 * b + 1], c0 = s_hdr[2 + 3 * b + 2];
      const int no = B.no, ni = B.ni, nterms = no * ni;
      const uint8_t* brec = rec + B.recipe_off;
      const uint32_t slot_mask = (1u << B.slot_bits) - 1u;
      for (uint32_t i = n0 + threadIdx.x; i < n1; i += kTileThreads) {
        const uint32_t e0 = (i == n0) ? 0u : uint32_t(__ldg(B.nnz_end + i - 1));
        const uint32_t e1 = __ldg(B.nnz_end + i);
        const uint32_t dest = __ldg(B.nnz_dest + i);
        double acc = 0.0;
        bool any = false;
        for (uint32_t p = e0; p < e1; ++p) {
          const uint32_t src = __ldg(B.con_src + c0 + p);
          const double* __restrict__ sc = slab + (src >> B.slot_bits);
          const uint8_t* __restrict__ r = brec + (src & slot_mask) * nterms;
          double v = 0.0;
          for (int o = 0; o < no; ++o) {
            double inner = 0.0;
            bool first = true;
            for (int q = 0; q < ni; ++q) {
              const uint32_t code = r[o * ni + q];
              if (code == 0xFFu) continue;
              double x = sc[(code & 0x7Fu) * P.cstride];
              if (code & 0x80u) x = -x;
              inner = first ? x : __dadd_rn(inner, x);
              first = false;
            }
            v = (o == 0) ? inner : __dadd_rn(v, inner);
          }
          any = any || (v != 0.0);
          acc = __dadd_rn(acc, v);
        }
        const bool kept = dest != kNoDest;
        if (P.check_classification && kept != any) *P.changed = 1;
        if (kept) B.values[dest] = acc;
      }
    }
  }
}

// ------------------------------------------------------------------ plan
struct TileBlockPlan {
  DevBuf<uint32_t> tile_nnz_ptr, tile_con_ptr, nnz_dest;
  DevBuf<uint16_t> nnz_end, con_src;
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
  size_t smem_bytes = 0;
  DevBuf<uint32_t> tile_cell_ptr, tile_cells;
  DevBuf<uint8_t> recipes;
  DevBuf<int> changed;
  DevBuf<unsigned int> ticket;
  int nblocks = 0;
  TileBlockPlan blk[kTileMaxBlocks];
  int grid = 0;
  void (*launch)(fq_ctx*, const TilePlan&, const TileParams&) = nullptr;
};

#define FQ_DECLARE_CORE(fn, n, k, nin, nd, nout)                                               \
  struct Core_##fn {                                                                           \
    static constexpr int kDistinct = nd;                                                       \
    template <class S>                                                                         \
    __device__ __forceinline__ void operator()(const double* __restrict__ s, S& sink) const {  \
      fn(s, sink);                                                                             \
    }                                                                                          \
  };
FQ_GEN_CORE_LIST(FQ_DECLARE_CORE)
#undef FQ_DECLARE_CORE

template <class Fn, int NE>
static void launch_tile(fq_ctx* ctx, const TilePlan& plan, const TileParams& params) {
  static bool attr_set = false;
  if (!attr_set) {
    FQ_CUDA(cudaFuncSetAttribute(tile_assemble_kernel<Fn, NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    attr_set = true;
  }
  tile_assemble_kernel<Fn, NE><<<plan.grid, kTileThreads, plan.smem_bytes, ctx->stream>>>(Fn{}, params);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

struct CoreEntryRt {
  int n, k, nin, ndistinct, nouts;
  const short* map;
  void (*launch)(fq_ctx*, const TilePlan&, const TileParams&);
};
#define FQ_CORE_ENTRY(fn, n, k, nin, nd, nout) CoreEntryRt{n, k, nin, nd, nout, fn##_map, &launch_tile<Core_##fn, nin>},
static const CoreEntryRt g_cores[] = {FQ_GEN_CORE_LIST(FQ_CORE_ENTRY)};
#undef FQ_CORE_ENTRY

static const CoreEntryRt* find_core(int n, int k) {
  for (const CoreEntryRt& e : g_cores)
    if (e.n == n && e.k == k) return &e;
  return nullptr;
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
  if (g < kc - 1 || g > kc + 1) return false;
  // offset of mass g inside the core map
  int off = 0;
  for (int gg = kc - 1; gg < g; ++gg) off += nlocal(n, gg) * nlocal(n, gg);
  const int nd = nlocal(n, g);
  auto mcode = [&](int i, int j, int sign) -> int {  // signed mass entry -> code or -1 (zero)
    const int m = core.map[off + i * nd + j];
    if (m < 0) return 0xFF;
    const int slot = m & 0xFF;
    const bool neg = ((m & 0x100) != 0) != (sign < 0);
    return slot | (neg ? 0x80 : 0);
  };
  if (g < 0 || g > n || nslots == 0) {  // zero space: every entry is an exact zero
    no = 0;
    ni = 0;
    codes.clear();
    return true;
  }
  if (core.ndistinct > 127) return false;
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
// ptr[t] = first i with key[i] >= t, for sorted keys; ptr[nkeys_range] = n
__global__ void seg_ptr_kernel(const uint32_t* __restrict__ key, size_t n, uint32_t nseg, uint32_t* __restrict__ ptr) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i <= n; i += stride) {
    const uint32_t hi = (i == n) ? nseg : key[i];
    const uint32_t lo = (i == 0) ? 0u : key[i - 1] + 1;
    for (uint32_t t = lo; t <= hi && t <= nseg; ++t) ptr[t] = uint32_t(i);
  }
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
__global__ void nnz_tile_kernel(const uint32_t* __restrict__ row_ptr, uint32_t nrows, const uint32_t* __restrict__ row_tile,
                                uint32_t* __restrict__ nnz_tile, uint32_t* __restrict__ nnz_id) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const uint32_t t = row_tile[r];
    for (uint32_t q = row_ptr[r]; q < row_ptr[r + 1]; ++q) nnz_tile[q] = t, nnz_id[q] = q;
  }
}
__global__ void nnz_len_kernel(const uint32_t* __restrict__ perm, uint32_t n, const uint32_t* __restrict__ contrib_ptr,
                               uint32_t* __restrict__ len) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t q = perm[i];
    len[i] = contrib_ptr[q + 1] - contrib_ptr[q];
  }
}
__global__ void tile_con_ptr_kernel(const uint32_t* __restrict__ tile_nnz_ptr, uint32_t ntiles, const uint32_t* __restrict__ G,
                                    uint32_t nnz, uint32_t ncontrib, uint32_t* __restrict__ tile_con_ptr) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t <= ntiles; t += stride) {
    const uint32_t i = tile_nnz_ptr[t];
    tile_con_ptr[t] = i < nnz ? G[i] : ncontrib;
  }
}
// per tile-ordered nnz: destination, end offset and the remapped contributions
__global__ void nnz_fill_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ sorted_tile, uint32_t n,
                                const uint32_t* __restrict__ G, const uint32_t* __restrict__ tile_con_ptr,
                                const uint32_t* __restrict__ contrib_ptr, const uint32_t* __restrict__ contrib_src,
                                uint32_t T, int slot_bits, const uint32_t* __restrict__ tile_cell_ptr,
                                const uint32_t* __restrict__ tile_cells, const uint8_t* __restrict__ keep,
                                const uint32_t* __restrict__ pos, int drop, uint32_t* __restrict__ nnz_dest,
                                uint16_t* __restrict__ nnz_end, uint16_t* __restrict__ con_src, int* __restrict__ err) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t q = perm[i], t = sorted_tile[i];
    const uint32_t p0 = contrib_ptr[q], p1 = contrib_ptr[q + 1];
    const uint32_t g0 = G[i];
    const uint32_t end = g0 + (p1 - p0) - tile_con_ptr[t];
    if (end > 0xFFFFu) atomicExch(err, 1);
    nnz_end[i] = uint16_t(end);
    nnz_dest[i] = drop ? (keep[q] ? pos[q] : kNoDest) : q;
    const uint32_t cb = tile_cell_ptr[t], ce = tile_cell_ptr[t + 1];
    for (uint32_t p = p0; p < p1; ++p) {
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
      const uint32_t local = lo - cb;
      if ((local << slot_bits) > 0xFFFFu) atomicExch(err, 3);
      con_src[g0 + (p - p0)] = uint16_t((local << slot_bits) | slot);
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
  if (!h.empty()) FQ_CUDA(cudaMemcpy(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
}

// Closed-form vertex bricks on a Kuhn grid.
int tile_cells_capacity(int ndistinct);

void tile_cluster_kuhn(fq_ctx* ctx, fq_mesh* mesh, int dim, const size_t* shape, size_t slab_begin, size_t slab_end_held) {
  if (dim > 3 || std::getenv("FQ_NO_TILE")) return;
  int max_distinct = 1;
  for (const CoreEntryRt& e : g_cores)
    if (e.n == dim) max_distinct = std::max(max_distinct, e.ndistinct);
  const int cells_capacity = tile_cells_capacity(max_distinct);
  // brick[a] owned vertices per axis; a tile needs the cells of prod(brick[a]+1) boxes
  const int ncelltypes = int(fact(dim));
  std::vector<uint32_t> brick(size_t(dim), 1);
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
        uint64_t boxes = 1, owned = 1;
        for (int b = 0; b < dim; ++b) {
          const uint64_t bb = brick[size_t(b)] + (b == a ? 1 : 0);
          boxes *= bb + 1;
          owned *= bb;
        }
        if (boxes * uint64_t(ncelltypes) > uint64_t(cells_capacity)) continue;
        const double ratio = double(boxes) / double(owned);
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
void tile_cluster_generic(fq_ctx* ctx, fq_mesh* mesh, const uint64_t* cell_verts) {
  const int dim = mesh->dim;
  if (dim > 3 || std::getenv("FQ_NO_TILE") || !cell_verts) return;
  int max_distinct = 1;
  for (const CoreEntryRt& e : g_cores)
    if (e.n == dim) max_distinct = std::max(max_distinct, e.ndistinct);
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

int tile_cells_capacity(int ndistinct) {
  // shared slab: ndistinct doubles per cell, leave room for recipes and headers
  const size_t budget = 227 * 1024 - 256 - 4096;
  int cap = int(budget / (size_t(ndistinct) * sizeof(double)));
  if (cap > 511) cap = 511;
  return cap;
}

static void radix_sort_pairs_u32(fq_ctx* ctx, DevBuf<uint32_t>& keys, DevBuf<uint32_t>& vals, size_t n, int end_bit) {
  DevBuf<uint32_t> keys_alt(n ? n : 1), vals_alt(n ? n : 1);
  cub::DoubleBuffer<uint32_t> dk(keys.p, keys_alt.p), dv(vals.p, vals_alt.p);
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, int64_t(n), 0, end_bit, ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, int64_t(n), 0, end_bit, ctx->stream));
  fq_count_launch(ctx, (end_bit + 7) / 8 + 1);
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (dk.Current() != keys.p) std::swap(keys, keys_alt);
  if (dv.Current() != vals.p) std::swap(vals, vals_alt);
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
  // core grade: the middle grade of the block set (hodge_blocks(k) -> k)
  int gmin = 1 << 30, gmax = -(1 << 30);
  for (int b = 0; b < nblocks; ++b) {
    if (csrs[b]->kind == KIND_LUMPED) return nullptr;
    gmin = std::min(gmin, csrs[b]->grade);
    gmax = std::max(gmax, csrs[b]->grade);
  }
  if (gmax - gmin > 2) return nullptr;
  int kc = (gmax - gmin == 2) ? gmin + 1 : (gmax - gmin == 1 ? gmax : gmin);
  kc = std::max(0, std::min(dim, kc));
  if (gmin < kc - 1 || gmax > kc + 1) return nullptr;
  const CoreEntryRt* core = find_core(dim, kc);
  if (!core) return nullptr;
  auto plan = std::make_shared<TilePlan>();
  plan->mesh = mesh;
  plan->dim = dim;
  plan->core_k = kc;
  plan->ndistinct = core->ndistinct;
  plan->nblocks = nblocks;
  plan->launch = core->launch;
  plan->ntiles = uint32_t(mesh->ntiles);
  // recipes
  std::vector<uint8_t> all_codes;
  for (int b = 0; b < nblocks; ++b) {
    std::vector<uint8_t> codes;
    TileBlockPlan& bp = plan->blk[b];
    if (!build_recipes(dim, kc, *core, csrs[b]->kind, csrs[b]->grade, bp.no, bp.ni, codes)) return nullptr;
    bp.recipe_off = int(all_codes.size());
    all_codes.insert(all_codes.end(), codes.begin(), codes.end());
    bp.csr = csrs[b];
    const uint32_t T = uint32_t(csrs[b]->el_rows * csrs[b]->el_cols);
    bp.slot_bits = bits_for32(T ? T : 1);
  }
  while (all_codes.size() % 8) all_codes.push_back(0xFF);
  upload_vec(plan->recipes, all_codes);
  const int block = 256;
  const int nv = dim + 1;
  const size_t ncells = mesh->ncells;
  const uint32_t v_lo = uint32_t(mesh->vtile_lo);
  // ---- tile cell lists
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
    uint32_t nvalid = 0;
    FQ_CUDA(cudaMemcpyAsync(&nvalid, d_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
    DevBuf<uint32_t> tile_of(nvalid ? nvalid : 1);
    plan->tile_cells.alloc(nvalid ? nvalid : 1);
    split_keys_kernel<<<grid_for(nvalid, block, ctx->sm_count), block, 0, ctx->stream>>>(dk.Current(), nvalid, tile_of.p,
                                                                                       plan->tile_cells.p);
    plan->tile_cell_ptr.alloc(size_t(plan->ntiles) + 1);
    seg_ptr_kernel<<<grid_for(size_t(nvalid) + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(
        tile_of.p, nvalid, plan->ntiles, plan->tile_cell_ptr.p);
    fq_count_launch(ctx, 2);
    FQ_CUDA(cudaGetLastError());
    std::vector<uint32_t> h_ptr(size_t(plan->ntiles) + 1);
    FQ_CUDA(cudaMemcpyAsync(h_ptr.data(), plan->tile_cell_ptr.p, h_ptr.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                            ctx->stream));
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
    uint32_t max_cells = 1;
    for (uint32_t t = 0; t < plan->ntiles; ++t) max_cells = std::max(max_cells, h_ptr[t + 1] - h_ptr[t]);
    if (int(max_cells) > tile_cells_capacity(plan->ndistinct)) return nullptr;
    plan->cstride = int(max_cells) | 1;  // odd stride: distinct rows start on different banks
    plan->smem_bytes = size_t(plan->cstride) * size_t(plan->ndistinct) * sizeof(double) + all_codes.size();
    if (plan->smem_bytes > 227 * 1024 - 256) return nullptr;
  }
  // ---- per block: tile-ordered nnz lists and remapped contributions
  DevBuf<int> d_err(1);
  FQ_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), ctx->stream));
  for (int b = 0; b < nblocks; ++b) {
    fq_csr* csr = csrs[b];
    TileBlockPlan& bp = plan->blk[b];
    const size_t s_nnz = csr->s_nnz;
    const size_t nrows_local = csr->row_end - csr->row_begin;
    bp.tile_nnz_ptr.alloc(size_t(plan->ntiles) + 1);
    bp.tile_con_ptr.alloc(size_t(plan->ntiles) + 1);
    bp.nnz_at_build = csr->nnz;
    bp.dropped_at_build = drop;
    if (s_nnz == 0) {
      FQ_CUDA(cudaMemsetAsync(bp.tile_nnz_ptr.p, 0, bp.tile_nnz_ptr.bytes(), ctx->stream));
      FQ_CUDA(cudaMemsetAsync(bp.tile_con_ptr.p, 0, bp.tile_con_ptr.bytes(), ctx->stream));
      bp.nnz_dest.alloc(1);
      bp.nnz_end.alloc(1);
      bp.con_src.alloc(1);
      continue;
    }
    int tg, rg;
    kind_grades(csr->kind, csr->grade, tg, rg);
    const int nt = nlocal(dim, tg);
    // top position of every local face of the test grade
    std::vector<uint8_t> top_pos;
    for (uint32_t m : colex_subsets(dim + 1, tg + 1)) top_pos.push_back(uint8_t(mask_elems(m).back()));
    DevBuf<uint8_t> d_top;
    upload_vec(d_top, top_pos);
    DevBuf<uint32_t> row_tile(nrows_local ? nrows_local : 1);
    FQ_CUDA(cudaMemsetAsync(row_tile.p, 0, row_tile.bytes(), ctx->stream));
    row_tile_kernel<<<grid_for(ncells * size_t(nt), block, ctx->sm_count), block, 0, ctx->stream>>>(
        mesh->cell_faces[size_t(tg)].p, nt, mesh->cell_faces[0].p, nv, d_top.p, ncells, mesh->vertex_tile.p, v_lo,
        uint32_t(csr->row_begin), uint32_t(csr->row_end), row_tile.p);
    DevBuf<uint32_t> nnz_tile(s_nnz), perm(s_nnz);
    nnz_tile_kernel<<<grid_for(nrows_local, block, ctx->sm_count), block, 0, ctx->stream>>>(
        csr->s_row_ptr.p, uint32_t(nrows_local), row_tile.p, nnz_tile.p, perm.p);
    fq_count_launch(ctx, 2);
    FQ_CUDA(cudaGetLastError());
    radix_sort_pairs_u32(ctx, nnz_tile, perm, s_nnz, bits_for32(plan->ntiles));
    seg_ptr_kernel<<<grid_for(s_nnz + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(nnz_tile.p, s_nnz, plan->ntiles,
                                                                                        bp.tile_nnz_ptr.p);
    DevBuf<uint32_t> len(s_nnz), G(s_nnz);
    nnz_len_kernel<<<grid_for(s_nnz, block, ctx->sm_count), block, 0, ctx->stream>>>(perm.p, uint32_t(s_nnz),
                                                                                    csr->contrib_ptr.p, len.p);
    size_t tmp_bytes = 0;
    FQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, len.p, G.p, int64_t(s_nnz), ctx->stream));
    DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
    FQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, len.p, G.p, int64_t(s_nnz), ctx->stream));
    tile_con_ptr_kernel<<<grid_for(size_t(plan->ntiles) + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(
        bp.tile_nnz_ptr.p, plan->ntiles, G.p, uint32_t(s_nnz), uint32_t(csr->ncontrib), bp.tile_con_ptr.p);
    bp.nnz_dest.alloc(s_nnz);
    bp.nnz_end.alloc(s_nnz);
    bp.con_src.alloc(csr->ncontrib ? csr->ncontrib : 1);
    const uint32_t T = uint32_t(csr->el_rows * csr->el_cols);
    nnz_fill_kernel<<<grid_for(s_nnz, block, ctx->sm_count), block, 0, ctx->stream>>>(
        perm.p, nnz_tile.p, uint32_t(s_nnz), G.p, bp.tile_con_ptr.p, csr->contrib_ptr.p, csr->contrib_src.p, T, bp.slot_bits,
        plan->tile_cell_ptr.p, plan->tile_cells.p, csr->keep.p, drop ? csr->pos.p : nullptr, drop ? 1 : 0, bp.nnz_dest.p,
        bp.nnz_end.p, bp.con_src.p, d_err.p);
    fq_count_launch(ctx, 6);
    FQ_CUDA(cudaGetLastError());
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  int h_err = 0;
  FQ_CUDA(cudaMemcpy(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (h_err) return nullptr;  // a tile exceeds the 16-bit local index space: keep the slab path
  plan->changed.alloc(1);
  plan->ticket.alloc(1);
  plan->grid = ctx->sm_count;
  return plan;
}

// Runs the fused kernel.  Returns false when the zero/non-zero classification of
// some entry differs from the plan's (the caller re-runs the slab path and rebuilds).
bool tile_assemble(fq_ctx* ctx, const fq_mesh* mesh, TilePlan& plan) {
  TileParams P{};
  P.tile_cell_ptr = plan.tile_cell_ptr.p;
  P.tile_cells = plan.tile_cells.p;
  P.cell_edges = mesh->cell_faces[1].p;
  P.lengths = mesh->lengths.p;
  P.edge_lo = uint32_t(mesh->edge_lo);
  P.ntiles = plan.ntiles;
  P.cstride = plan.cstride;
  P.nblocks = plan.nblocks;
  P.recipes = plan.recipes.p;
  P.recipe_bytes = int(plan.recipes.n);
  P.check_classification = plan.blk[0].dropped_at_build ? 1 : 0;
  P.changed = plan.changed.p;
  P.ticket = plan.ticket.p;
  for (int b = 0; b < plan.nblocks; ++b) {
    const TileBlockPlan& bp = plan.blk[b];
    TileBlockDev& d = P.blk[b];
    d.tile_nnz_ptr = bp.tile_nnz_ptr.p;
    d.tile_con_ptr = bp.tile_con_ptr.p;
    d.nnz_dest = bp.nnz_dest.p;
    d.nnz_end = bp.nnz_end.p;
    d.con_src = bp.con_src.p;
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
    plan.launch(ctx, plan, P);
  }
  int changed = 0;
  FQ_CUDA(cudaMemcpyAsync(&changed, plan.changed.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
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
  int64_t total = int64_t(plan.tile_cell_ptr.bytes() + plan.tile_cells.bytes());
  for (int b = 0; b < plan.nblocks; ++b) {
    const TileBlockPlan& bp = plan.blk[b];
    total += int64_t(bp.tile_nnz_ptr.bytes() + bp.tile_con_ptr.bytes() + bp.nnz_dest.bytes() + bp.nnz_end.bytes() +
                     bp.con_src.bytes());
  }
  return total;
}

}  // namespace fq
