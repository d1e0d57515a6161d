// assemble.cu — K2 (symbolic pattern + cell-slot -> nnz map) and K3 (numeric
// scatter as a segmented reduction, no atomics).
//
// Reference semantics (formoniq/src/galerkin.rs:138-188): triplets are emitted
// cell by cell (i outer, j inner), entries equal to 0.0 are dropped, and the
// COO -> CSR conversion sums duplicates.  Here:
//   symbolic: one key (row, col) per (cell, slot); a stable radix sort groups
//     the contributions of every structural non-zero in ascending cell order
//     (so the summation order is deterministic and independent of the
//     partition); run heads give row_ptr / col_idx and the gather lists.
//   numeric: the element slab is produced by K1, then one thread per
//     structural non-zero sums its contributions in order and records whether
//     any of them was non-zero; with drop_exact_zeros the pattern is
//     compacted to exactly the reference's value-dependent pattern.
#include <cub/cub.cuh>

#include <cstdlib>
#include <functional>

#include "internal.hpp"
#include "stream.cuh"

namespace fq {

struct IsOwnedKey {
  __device__ __forceinline__ uint32_t operator()(const uint64_t& k) const { return k != ~0ull ? 1u : 0u; }
};

static int bits_for(uint64_t n) {  // bits needed to represent values < n
  int b = 1;
  while ((1ull << b) < n) ++b;
  return b;
}

__global__ void sym_keys_kernel(const uint32_t* __restrict__ faces_t, const uint32_t* __restrict__ faces_r, int nt,
                                int nr, uint64_t ncontrib, uint32_t row_begin, uint32_t row_end, int col_bits,
                                uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  const uint32_t T = uint32_t(nt * nr);
  for (uint64_t p = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < ncontrib; p += stride) {
    const uint64_t c = p / T;
    const uint32_t slot = uint32_t(p % T);
    const uint32_t i = slot / uint32_t(nr), j = slot % uint32_t(nr);
    const uint32_t row = faces_t[c * nt + i], col = faces_r[c * nr + j];
    uint64_t key;
    if (row < row_begin || row >= row_end)
      key = ~0ull;  // not owned: sorted to the end and cut off
    else
      key = (uint64_t(row - row_begin) << col_bits) | col;
    keys[p] = key;
    vals[p] = uint32_t(p);
  }
}

__global__ void sym_heads_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ head) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t p = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < n; p += stride) {
    const uint64_t k = keys[p];
    head[p] = (k != ~0ull && (p == 0 || keys[p - 1] != k)) ? 1u : 0u;
  }
}

// For every run head: record its start (contrib_ptr), its column, and fill
// row_ptr for the rows that begin at or before it.
__global__ void sym_emit_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ head,
                                const uint32_t* __restrict__ head_scan, uint64_t n, int col_bits,
                                uint32_t* __restrict__ contrib_ptr, uint32_t* __restrict__ col_idx,
                                uint32_t* __restrict__ row_of_nnz) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  const uint64_t col_mask = (1ull << col_bits) - 1;
  for (uint64_t p = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < n; p += stride) {
    if (!head[p]) continue;
    const uint32_t q = head_scan[p];
    const uint64_t k = keys[p];
    contrib_ptr[q] = uint32_t(p);
    col_idx[q] = uint32_t(k & col_mask);
    row_of_nnz[q] = uint32_t(k >> col_bits);
  }
}

// row_ptr[r] = first nnz whose row >= r  (rows are sorted)
__global__ void sym_rowptr_kernel(const uint32_t* __restrict__ row_of_nnz, uint32_t nnz, uint32_t nrows,
                                  uint32_t* __restrict__ row_ptr) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t q = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; q <= nnz; q += stride) {
    const uint32_t r_hi = (q == nnz) ? nrows : row_of_nnz[q];
    const uint32_t r_lo = (q == 0) ? 0u : row_of_nnz[q - 1] + 1;
    for (uint32_t r = r_lo; r <= r_hi && r <= nrows; ++r) row_ptr[r] = uint32_t(q);
  }
}

// Symbolic phase, part 1: shape, element layout and row range of the block; the pattern itself is produced either by
// the tile plan builder (tile.cu) on the first numeric pass or, for the slab path, by K2 below (lazily).
void assemble_symbolic(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, size_t row_begin, size_t row_end,
                       fq_csr* out) {
  const int dim = mesh->dim;
  int tg, rg;
  kind_grades(kind, grade, tg, rg);
  const int nt = nlocal(dim, tg), nr = nlocal(dim, rg);
  const size_t grows = (tg < 0 || tg > dim) ? 0 : mesh->nsimplices[size_t(tg)];
  const size_t gcols = (rg < 0 || rg > dim) ? 0 : mesh->nsimplices[size_t(rg)];
  if (row_end > grows) row_end = grows;
  if (row_begin > row_end) row_begin = row_end;
  out->nrows = grows;
  out->ncols = gcols;
  out->row_begin = row_begin;
  out->row_end = row_end;
  out->kind = kind;
  out->grade = grade;
  out->dim = dim;
  out->el_rows = nt;
  out->el_cols = nr;
  out->ncells = mesh->ncells;
  out->has_plan = true;
  out->compact_valid = false;
  out->pattern_valid = false;
  out->k2_done = false;
  const size_t nrows_local = row_end - row_begin;
  const uint64_t ncontrib_all = uint64_t(mesh->ncells) * uint64_t(nt) * uint64_t(nr);
  if (ncontrib_all == 0 || nrows_local == 0) {
    // pairing() of an empty space: correctly shaped zero matrix (whitney_complex.rs:113-122)
    out->s_row_ptr.alloc(nrows_local + 1);
    FQ_CUDA(cudaMemsetAsync(out->s_row_ptr.p, 0, (nrows_local + 1) * sizeof(uint32_t), ctx->stream));
    out->s_nnz = 0;
    out->ncontrib = 0;
    out->s_col_idx.alloc(1);
    out->contrib_ptr.alloc(1);
    FQ_CUDA(cudaMemsetAsync(out->contrib_ptr.p, 0, sizeof(uint32_t), ctx->stream));
    out->contrib_src.alloc(1);
    out->s_values.alloc(1);
    out->keep.alloc(1);
    out->k2_done = true;
    return;
  }
  FQ_REQUIRE(mesh->cell_faces[size_t(tg)].p && mesh->cell_faces[size_t(rg)].p,
             "mesh was created without the face tables of the required grades");
  FQ_REQUIRE(ncontrib_all < (1ull << 32), "more than 2^32 element entries in one block: not supported");
  FQ_REQUIRE(gcols < (1ull << 32) && grows < (1ull << 32), "more than 2^32 rows/cols: not supported");
}

// Symbolic phase, part 2 (K2): the global-sort pattern + cell-slot -> nnz map of the slab path.
static void ensure_k2(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* out) {
  if (out->k2_done) return;
  const int dim = mesh->dim;
  const int kind = out->kind, grade = out->grade;
  int tg, rg;
  kind_grades(kind, grade, tg, rg);
  const int nt = nlocal(dim, tg), nr = nlocal(dim, rg);
  const size_t gcols = out->ncols;
  const size_t row_begin = out->row_begin, row_end = out->row_end;
  const size_t nrows_local = row_end - row_begin;
  const size_t T = size_t(nt) * size_t(nr);
  const uint64_t ncontrib_all = uint64_t(mesh->ncells) * T;
  out->s_row_ptr.alloc(nrows_local + 1);
  const int col_bits = bits_for(gcols);
  const int row_bits = bits_for(nrows_local + 1);
  FQ_REQUIRE(col_bits + row_bits <= 63, "key overflow");
  const int block = 256;
  ScopedSpan span_sym(ctx, "k2_symbolic");
  DevBuf<uint64_t> keys(ncontrib_all), keys_alt(ncontrib_all);
  DevBuf<uint32_t> vals(ncontrib_all), vals_alt(ncontrib_all);
  sym_keys_kernel<<<grid_for(ncontrib_all, block, ctx->sm_count), block, 0, ctx->stream>>>(
      mesh->cell_faces[size_t(tg)].p, mesh->cell_faces[size_t(rg)].p, nt, nr, ncontrib_all, uint32_t(row_begin),
      uint32_t(row_end), col_bits, keys.p, vals.p);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
  cub::DoubleBuffer<uint64_t> dk(keys.p, keys_alt.p);
  cub::DoubleBuffer<uint32_t> dv(vals.p, vals_alt.p);
  size_t tmp_bytes = 0;
  // non-owned keys are all-ones: sort on 64 bits would be wasteful, so sort on
  // [0, row_bits+col_bits+1) after remapping ~0 -> the bit just above the range.
  // Simpler and exact: all-ones keys have every bit set, so sorting on the low
  // (row_bits + col_bits + 1) bits still places them last.
  const int end_bit = std::min(64, row_bits + col_bits + 1);
  FQ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, int64_t(ncontrib_all), 0, end_bit, ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes);
  FQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, int64_t(ncontrib_all), 0, end_bit, ctx->stream));
  fq_count_launch(ctx, (end_bit + 7) / 8 + 1);
  const uint64_t* skeys = dk.Current();
  const uint32_t* svals = dv.Current();
  // run heads -> nnz ids
  DevBuf<uint32_t> head(ncontrib_all), head_scan(ncontrib_all + 1);
  sym_heads_kernel<<<grid_for(ncontrib_all, block, ctx->sm_count), block, 0, ctx->stream>>>(skeys, ncontrib_all, head.p);
  fq_count_launch(ctx);
  size_t tmp2 = 0;
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, head.p, head_scan.p, int64_t(ncontrib_all), ctx->stream));
  if (tmp.n < tmp2) tmp.alloc(tmp2);
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp2, head.p, head_scan.p, int64_t(ncontrib_all), ctx->stream));
  fq_count_launch(ctx, 2);
  uint32_t last_scan = 0, last_head = 0;
  FQ_CUDA(cudaMemcpyAsync(&last_scan, head_scan.p + (ncontrib_all - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          ctx->stream));
  FQ_CUDA(cudaMemcpyAsync(&last_head, head.p + (ncontrib_all - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  const size_t s_nnz = size_t(last_scan) + last_head;
  out->s_nnz = s_nnz;
  out->s_col_idx.alloc(s_nnz ? s_nnz : 1);
  out->contrib_ptr.alloc(s_nnz + 1);
  DevBuf<uint32_t> row_of_nnz(s_nnz ? s_nnz : 1);
  sym_emit_kernel<<<grid_for(ncontrib_all, block, ctx->sm_count), block, 0, ctx->stream>>>(
      skeys, head.p, head_scan.p, ncontrib_all, col_bits, out->contrib_ptr.p, out->s_col_idx.p, row_of_nnz.p);
  fq_count_launch(ctx);
  // number of owned contributions = first position holding an all-ones key
  // = total - (#non-owned).  Count owned keys with a reduction over a transform.
  {
    // owned keys are sorted first; find the boundary by binary search on the host-free path:
    // count of keys != ~0 via cub::DeviceReduce on a transform iterator.
    cub::TransformInputIterator<uint32_t, IsOwnedKey, const uint64_t*> it(skeys, IsOwnedKey());
    DevBuf<uint32_t> d_count(1);
    size_t tmp3 = 0;
    FQ_CUDA(cub::DeviceReduce::Sum(nullptr, tmp3, it, d_count.p, int64_t(ncontrib_all), ctx->stream));
    if (tmp.n < tmp3) tmp.alloc(tmp3);
    FQ_CUDA(cub::DeviceReduce::Sum(tmp.p, tmp3, it, d_count.p, int64_t(ncontrib_all), ctx->stream));
    fq_count_launch(ctx);
    uint32_t owned = 0;
    FQ_CUDA(cudaMemcpyAsync(&owned, d_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
    out->ncontrib = owned;
  }
  const uint32_t ncontrib32 = uint32_t(out->ncontrib);
  FQ_CUDA(cudaMemcpyAsync(out->contrib_ptr.p + s_nnz, &ncontrib32, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  out->contrib_src.alloc(out->ncontrib ? out->ncontrib : 1);
  FQ_CUDA(cudaMemcpyAsync(out->contrib_src.p, svals, out->ncontrib * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                          ctx->stream));
  sym_rowptr_kernel<<<grid_for(s_nnz + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(
      row_of_nnz.p, uint32_t(s_nnz), uint32_t(nrows_local), out->s_row_ptr.p);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
  out->s_values.alloc(s_nnz ? s_nnz : 1);
  out->keep.alloc(s_nnz ? s_nnz : 1);
  stream_build_blocks(ctx, out->contrib_ptr.p, s_nnz, out->ncontrib, out->gather_blocks, out->ngather_blocks);
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  out->k2_done = true;
}

// ------------------------------------------------------------------ numeric
// K3 is a segmented reduction over the element slab: one segment per structural
// non-zero, items = its contributions in ascending cell order (stream.cuh).
// Every segment also yields the "some contribution != 0.0" flag of galerkin.rs:173.
//
// GatherStructural: values[q] = sum, keep[q] = any   (first pass / no dropping)
// GatherCompacted:  values[pos[q]] = sum for kept entries; if the zero/non-zero
//   classification differs from the cached one (the geometry changed), raise
//   *changed so the host redoes the compaction.
struct GatherStructural {
  static constexpr bool kHasValues = false;
  static constexpr bool kCustomSrc = false;
  static constexpr bool kGated = false;
  __device__ __forceinline__ double load(uint32_t, bool) const { return 0.0; }
  double* __restrict__ values;
  uint8_t* __restrict__ keep;
  __device__ __forceinline__ void store(uint32_t q, double sum, bool any) const {
    values[q] = sum;
    keep[q] = any ? 1 : 0;
  }
};
struct GatherCompacted {
  static constexpr bool kHasValues = false;
  static constexpr bool kCustomSrc = false;
  static constexpr bool kGated = false;
  __device__ __forceinline__ double load(uint32_t, bool) const { return 0.0; }
  double* __restrict__ values;
  const uint8_t* __restrict__ keep;
  const uint32_t* __restrict__ pos;
  int* __restrict__ changed;
  __device__ __forceinline__ void store(uint32_t q, double sum, bool any) const {
    const bool kept = keep[q] != 0;
    if (kept != any) *changed = 1;
    if (kept) values[pos[q]] = sum;
  }
};

__global__ void num_keep_to_u32(const uint8_t* __restrict__ keep, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) out[q] = keep[q];
}

__global__ void num_compact_kernel(const uint8_t* __restrict__ keep, const uint32_t* __restrict__ pos, uint32_t s_nnz,
                                   const uint32_t* __restrict__ s_col, const double* __restrict__ s_val,
                                   uint32_t* __restrict__ col, double* __restrict__ val) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < s_nnz; q += stride)
    if (keep[q]) {
      col[pos[q]] = s_col[q];
      val[pos[q]] = s_val[q];
    }
}
__global__ void num_rowptr_kernel(const uint32_t* __restrict__ s_row_ptr, const uint32_t* __restrict__ pos,
                                  uint32_t nrows, uint32_t* __restrict__ row_ptr) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= nrows; r += stride) row_ptr[r] = pos[s_row_ptr[r]];
}

// keep[] / s_values[] of a structural pass -> the reference's value-dependent pattern (galerkin.rs:173): pos = exclusive
// scan of keep, compacted row_ptr / col_idx / values.
static void compact_pattern(fq_ctx* ctx, fq_csr* csr) {
  const size_t nrows_local = csr->row_end - csr->row_begin;
  const size_t s_nnz = csr->s_nnz;
  const int block = 256;
  ScopedSpan span_compact(ctx, "k3_compact");
  DevBuf<uint32_t> k32(s_nnz + 1);
  if (csr->pos.n != s_nnz + 1) csr->pos.alloc(s_nnz + 1);
  if (csr->d_changed.n != 1) csr->d_changed.alloc(1);
  num_keep_to_u32<<<grid_for(s_nnz, block, ctx->sm_count), block, 0, ctx->stream>>>(csr->keep.p, uint32_t(s_nnz), k32.p);
  FQ_CUDA(cudaMemsetAsync(k32.p + s_nnz, 0, sizeof(uint32_t), ctx->stream));
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, k32.p, csr->pos.p, int64_t(s_nnz + 1), ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes);
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, k32.p, csr->pos.p, int64_t(s_nnz + 1), ctx->stream));
  fq_count_launch(ctx, 3);
  uint32_t nnz = 0;
  FQ_CUDA(cudaMemcpyAsync(&nnz, csr->pos.p + s_nnz, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  csr->dropped = true;
  csr->nnz = nnz;
  csr->row_ptr.alloc(nrows_local + 1);
  csr->col_idx.alloc(nnz ? nnz : 1);
  csr->values.alloc(nnz ? nnz : 1);
  num_compact_kernel<<<grid_for(s_nnz, block, ctx->sm_count), block, 0, ctx->stream>>>(
      csr->keep.p, csr->pos.p, uint32_t(s_nnz), csr->s_col_idx.p, csr->s_values.p, csr->col_idx.p, csr->values.p);
  num_rowptr_kernel<<<grid_for(nrows_local + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(
      csr->s_row_ptr.p, csr->pos.p, uint32_t(nrows_local), csr->row_ptr.p);
  fq_count_launch(ctx, 2);
  FQ_CUDA(cudaGetLastError());
  csr->compact_valid = true;
  csr->pattern_valid = true;
  csr->spmv_ready = false;
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
}

static void gather_structural(fq_ctx* ctx, fq_csr* csr) {
  ScopedSpan span(ctx, "k3_gather");
  stream_reduce(ctx, csr->gather_blocks.p, csr->ngather_blocks, csr->contrib_ptr.p, csr->contrib_src.p, nullptr, csr->slab.p,
                GatherStructural{csr->s_values.p, csr->keep.p});
}

// After K1 filled csr->slab: reduce into the CSR values under the requested pattern semantics.
static void numeric_reduce(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* csr, bool drop_exact_zeros) {
  const size_t nrows_local = csr->row_end - csr->row_begin;
  const size_t s_nnz = csr->s_nnz;
  const int block = 256;
  const int ne = int(binom(mesh->dim + 1, 2));
  csr->assembly_shared_bytes = int64_t(8 * mesh->lengths.n + 4 * size_t(ne) * mesh->ncells);
  csr->assembly_bytes = csr->assembly_shared_bytes + int64_t(4 * csr->ncontrib);
  if (!drop_exact_zeros || s_nnz == 0) {
    // the structural pattern is the result; its index arrays are shared once
    if (csr->dropped || csr->row_ptr.n != nrows_local + 1 || csr->nnz != s_nnz || !csr->pattern_valid) {
      csr->row_ptr.alloc(nrows_local + 1);
      csr->col_idx.alloc(s_nnz ? s_nnz : 1);
      csr->values.alloc(s_nnz ? s_nnz : 1);
      FQ_CUDA(cudaMemcpyAsync(csr->row_ptr.p, csr->s_row_ptr.p, (nrows_local + 1) * sizeof(uint32_t),
                              cudaMemcpyDeviceToDevice, ctx->stream));
      if (s_nnz)
        FQ_CUDA(cudaMemcpyAsync(csr->col_idx.p, csr->s_col_idx.p, s_nnz * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                ctx->stream));
      csr->spmv_ready = false;
    }
    csr->dropped = false;
    csr->pattern_valid = true;
    csr->compact_valid = false;
    csr->nnz = s_nnz;
    if (s_nnz) {
      ScopedSpan span(ctx, "k3_gather");
      stream_reduce(ctx, csr->gather_blocks.p, csr->ngather_blocks, csr->contrib_ptr.p, csr->contrib_src.p, nullptr,
                    csr->slab.p, GatherStructural{csr->values.p, csr->keep.p});
    }
    csr->assembly_bytes += int64_t(8 * s_nnz);
    return;
  }
  if (csr->compact_valid && csr->dropped) {
    // fast path: the cached value-dependent pattern is reused and verified
    {
      ScopedSpan span(ctx, "k3_gather");
      FQ_CUDA(cudaMemsetAsync(csr->d_changed.p, 0, sizeof(int), ctx->stream));
      stream_reduce(ctx, csr->gather_blocks.p, csr->ngather_blocks, csr->contrib_ptr.p, csr->contrib_src.p, nullptr,
                    csr->slab.p, GatherCompacted{csr->values.p, csr->keep.p, csr->pos.p, csr->d_changed.p});
    }
    int changed = 0;
    FQ_CUDA(cudaMemcpyAsync(&changed, csr->d_changed.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
    csr->assembly_bytes += int64_t(8 * csr->nnz);
    if (!changed) return;
  }
  // slow path: structural gather, then compaction to the reference's value-dependent pattern
  gather_structural(ctx, csr);
  compact_pattern(ctx, csr);
  csr->assembly_bytes += int64_t(8 * csr->nnz);
}

// The structural pattern becomes the active one (no dropping): index arrays copied once, values assembled in place.
static void adopt_structural_pattern(fq_ctx* ctx, fq_csr* csr) {
  const size_t nrows_local = csr->row_end - csr->row_begin;
  const size_t s_nnz = csr->s_nnz;
  if (csr->dropped || csr->row_ptr.n != nrows_local + 1 || csr->nnz != s_nnz || !csr->pattern_valid) {
    csr->row_ptr.alloc(nrows_local + 1);
    csr->col_idx.alloc(s_nnz ? s_nnz : 1);
    csr->values.alloc(s_nnz ? s_nnz : 1);
    FQ_CUDA(cudaMemcpyAsync(csr->row_ptr.p, csr->s_row_ptr.p, (nrows_local + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                            ctx->stream));
    if (s_nnz)
      FQ_CUDA(cudaMemcpyAsync(csr->col_idx.p, csr->s_col_idx.p, s_nnz * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    csr->spmv_ready = false;
  }
  csr->dropped = false;
  csr->pattern_valid = true;
  csr->compact_valid = false;
  csr->nnz = s_nnz;
}

// Numeric pass through the tile-fused kernel.  The plan's record streams carry either structural destinations (first
// pass, or no dropping) or the destinations of the value-dependent pattern found by the first pass.
static bool tile_numeric(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* const* csrs, int nblocks, bool drop) {
  for (int attempt = 0; attempt < 2; ++attempt) {
    std::shared_ptr<TilePlan> plan = csrs[0]->tile_plan;
    if (plan && !tile_plan_matches(*plan, mesh, csrs, nblocks)) plan.reset();
    // a plan whose destinations were retargeted to a dropped pattern cannot serve a structural pass: rebuild
    if (plan && tile_plan_compact(*plan) && !drop) plan.reset();
    if (!plan) {
      ScopedSpan span(ctx, "tile_plan_build");
      plan = tile_plan_build(ctx, mesh, csrs, nblocks);
      // generic clustering estimates the cell visits of a tile: when a tile turns out too large, cut smaller ones
      for (int retry = 0; !plan && mesh->cluster_generic && retry < 3; ++retry) {
        fq_mesh* m = const_cast<fq_mesh*>(mesh);
        m->cluster_scale *= 1.35f;
        tile_cluster_generic(ctx, m);
        if (!mesh->vertex_tile.p) break;
        plan = tile_plan_build(ctx, mesh, csrs, nblocks);
      }
      for (int b = 0; b < nblocks; ++b) {
        csrs[b]->tile_plan = plan;
        csrs[b]->pattern_valid = false;
        csrs[b]->compact_valid = false;
        csrs[b]->plan_build_ms = plan ? tile_plan_build_ms(*plan) : 0.0;
        csrs[b]->plan_cell_visits = plan ? tile_plan_cell_visits(*plan) : 0;
      }
      if (!plan) {
        csrs[0]->tile_refused = 1;  // does not apply to this block set / mesh: stay on the slab path
        return false;
      }
      for (int b = 0; b < nblocks; ++b) {  // the slab path's buffers are not needed any more
        csrs[b]->slab.release();
        if (csrs[b]->s_nnz > 0 && csrs[b]->k2_done) {
          csrs[b]->contrib_ptr.release();
          csrs[b]->contrib_src.release();
          csrs[b]->gather_blocks.release();
          csrs[b]->k2_done = false;
        }
      }
    }
    double* vals[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t* keeps[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int b = 0; b < nblocks; ++b) csrs[b]->inv_diag.release();
    if (!drop) {
      for (int b = 0; b < nblocks; ++b) {
        adopt_structural_pattern(ctx, csrs[b]);
        vals[b] = csrs[b]->values.p;
      }
      tile_assemble(ctx, mesh, *plan, vals, nullptr);
      tile_plan_drop(*plan) = false;
      return true;
    }
    if (tile_plan_compact(*plan)) {
      // steady state: the cached value-dependent pattern is reused and verified by the kernel
      for (int b = 0; b < nblocks; ++b) vals[b] = csrs[b]->values.p;
      if (tile_assemble(ctx, mesh, *plan, vals, nullptr)) return true;
      // the zero / non-zero classification changed with the geometry: rebuild the structural plan and redo the pass
      for (int b = 0; b < nblocks; ++b) csrs[b]->tile_plan.reset();
      continue;
    }
    // first pass under dropping semantics: structural values + classification, then compaction
    for (int b = 0; b < nblocks; ++b) {
      fq_csr* csr = csrs[b];
      const size_t n = csr->s_nnz ? csr->s_nnz : 1;
      if (csr->s_values.n != n) csr->s_values.alloc(n);
      if (csr->keep.n != n) csr->keep.alloc(n);
      vals[b] = csr->s_values.p;
      keeps[b] = csr->keep.p;
    }
    tile_assemble(ctx, mesh, *plan, vals, keeps);
    for (int b = 0; b < nblocks; ++b) {
      if (csrs[b]->s_nnz == 0) {
        adopt_structural_pattern(ctx, csrs[b]);
        csrs[b]->dropped = true;
        csrs[b]->compact_valid = true;
        if (csrs[b]->pos.n != 1) {
          csrs[b]->pos.alloc(1);
          FQ_CUDA(cudaMemsetAsync(csrs[b]->pos.p, 0, sizeof(uint32_t), ctx->stream));
        }
        continue;
      }
      compact_pattern(ctx, csrs[b]);
    }
    tile_retarget(ctx, *plan);
    tile_plan_drop(*plan) = true;
    for (int b = 0; b < nblocks; ++b) csrs[b]->s_values.release();  // scratch of the first pass only
    return true;
  }
  return false;
}

void assemble_numeric_multi(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* const* csrs, int nblocks, bool drop_exact_zeros) {
  // ---- fast path: tile-fused K1+K3 (tile.cu); its plan builder is also the symbolic phase
  {
    bool ready = nblocks >= 1 && csrs[0]->tile_refused == 0;
    for (int b = 0; ready && b < nblocks; ++b)
      ready = csrs[b]->has_plan && csrs[b]->ncells == mesh->ncells && csrs[b]->dim == mesh->dim;
    // Uploaded meshes: a one-shot assembly (BilinearForm::assemble: symbolic + one numeric pass) is cheaper through K2 and
    // the slab path than through a tile plan (the plan builder sorts every tile's entries twice: ~60 ms per block at
    // 12.6 M tets against ~30 ms for K2); the fused kernel takes over when the matrix is assembled again.  Meshes from
    // the device Kuhn generator carry closed-form bricks and go through the fused kernel from the first pass.
    if (ready && !mesh->vertex_tile.p && !std::getenv("FQ_TILE_FIRST")) {
      for (int b = 0; ready && b < nblocks; ++b) ready = csrs[b]->slab_passes >= 1;
    }
    if (ready && !mesh->vertex_tile.p && !mesh->cluster_tried)
      tile_cluster_generic(ctx, const_cast<fq_mesh*>(mesh));  // clustered once, on first reuse
    ready = ready && mesh->vertex_tile.p != nullptr;
    if (ready && tile_numeric(ctx, mesh, csrs, nblocks, drop_exact_zeros)) {
      const int ne = int(binom(mesh->dim + 1, 2));
      for (int b = 0; b < nblocks; ++b) {
        csrs[b]->assembly_shared_bytes = int64_t(8 * mesh->lengths.n + 4 * size_t(ne) * mesh->ncells);
        csrs[b]->assembly_bytes = csrs[b]->assembly_shared_bytes + int64_t(4 * csrs[b]->ncontrib + 8 * csrs[b]->nnz);
      }
      return;
    }
  }
  for (int b = 0; b < nblocks; ++b) {
    FQ_REQUIRE(csrs[b]->has_plan, "matrix has no assembly plan (uploaded matrices cannot be re-assembled)");
    FQ_REQUIRE(csrs[b]->ncells == mesh->ncells && csrs[b]->dim == mesh->dim, "mesh does not match the symbolic phase");
    if (!csrs[b]->k2_done) {
      // the tile builder may have left its structural pattern here: K2 recomputes the same one with the gather lists
      csrs[b]->pattern_valid = false;
      csrs[b]->compact_valid = false;
      ensure_k2(ctx, mesh, csrs[b]);
    }
  }
  std::vector<BlockSpec> blocks;
  std::vector<double*> outs;
  bool any_work = false;
  for (int b = 0; b < nblocks; ++b) {
    fq_csr* csr = csrs[b];
    FQ_REQUIRE(csr->has_plan, "matrix has no assembly plan (uploaded matrices cannot be re-assembled)");
    FQ_REQUIRE(csr->ncells == mesh->ncells && csr->dim == mesh->dim, "mesh does not match the symbolic phase");
    csr->inv_diag.release();
    const size_t T = size_t(csr->el_rows) * size_t(csr->el_cols);
    const size_t want = mesh->ncells * T;
    if (csr->slab.n != (want ? want : 1)) csr->slab.alloc(want ? want : 1);  // persistent across numeric calls
    blocks.push_back(BlockSpec{csr->kind, csr->grade});
    outs.push_back(csr->slab.p);
    any_work = any_work || (csr->s_nnz > 0);
  }
  if (any_work) {
    ScopedSpan span(ctx, "k1_elmat");
    if (nblocks == 1 || elmat_has_generated(mesh->dim, blocks)) {
      elmat_to_slabs(ctx, mesh, blocks, 0, mesh->ncells, true, outs.data(), nullptr);
    } else {
      for (int b = 0; b < nblocks; ++b) {
        double* one[1] = {outs[size_t(b)]};
        if (csrs[b]->s_nnz > 0) elmat_to_slabs(ctx, mesh, {blocks[size_t(b)]}, 0, mesh->ncells, true, one, nullptr);
      }
    }
  }
  for (int b = 0; b < nblocks; ++b) {
    numeric_reduce(ctx, mesh, csrs[b], drop_exact_zeros);
    csrs[b]->slab_passes += 1;
  }
}

// Numeric pass of a quadrature form: the element matrices come from `fill` (which writes the cell-major element slab)
// instead of the generated K1 kernels; K3 and the pattern semantics are those of every other form.
void assemble_numeric_custom(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* csr, bool drop_exact_zeros,
                             const std::function<void(double*)>& fill) {
  FQ_REQUIRE(csr->has_plan, "matrix has no assembly plan (uploaded matrices cannot be re-assembled)");
  FQ_REQUIRE(csr->ncells == mesh->ncells && csr->dim == mesh->dim, "mesh does not match the symbolic phase");
  csr->inv_diag.release();
  csr->tile_plan.reset();  // the tile-fused kernel evaluates the closed-form masses: not this form
  csr->tile_refused = 1;
  if (!csr->k2_done) {
    csr->pattern_valid = false;
    csr->compact_valid = false;
    ensure_k2(ctx, mesh, csr);
  }
  const size_t T = size_t(csr->el_rows) * size_t(csr->el_cols);
  const size_t want = mesh->ncells * T;
  if (csr->slab.n != (want ? want : 1)) csr->slab.alloc(want ? want : 1);
  if (csr->s_nnz > 0) fill(csr->slab.p);
  numeric_reduce(ctx, mesh, csr, drop_exact_zeros);
  csr->slab_passes += 1;
}

void assemble_numeric(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* csr, bool drop_exact_zeros) {
  fq_csr* one[1] = {csr};
  assemble_numeric_multi(ctx, mesh, one, 1, drop_exact_zeros);
}

}  // namespace fq
