// blockop.cu — block operators built on the device from assembled blocks.
//
// HodgeBlocks::mixed_hodge_laplacian (formoniq/src/hodge.rs:93-99) stitches
//     [[ M_{k-1},  -dif_test   ],
//      [ dif_test^T,  dif_both ]]
// through CooMatrixExt::block (simplicial/src/linalg.rs:110-167) and one more serial
// COO -> CSR conversion.  Here the transpose is a stable radix sort by column and the
// stitching is a row-wise concatenation, both on the device; entries are copied (or
// negated) exactly, so the result is bit-identical to the reference's stitched matrix.
#include <cub/cub.cuh>

#include "internal.hpp"

namespace fq {

__global__ void expand_rows_kernel(const uint32_t* __restrict__ row_ptr, uint32_t nrows, uint32_t* __restrict__ row_of,
                                   uint32_t* __restrict__ id) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride)
    for (uint32_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) row_of[p] = r, id[p] = p;
}
__global__ void transpose_fill_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ row_of,
                                      const double* __restrict__ val, uint32_t nnz, uint32_t* __restrict__ col_t,
                                      double* __restrict__ val_t) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
    const uint32_t p = perm[i];
    col_t[i] = row_of[p];
    val_t[i] = val[p];
  }
}
// row_ptr[r] = first i with key[i] >= r  (keys sorted)
__global__ void lower_bound_ptr_kernel(const uint32_t* __restrict__ key, uint32_t n, uint32_t nrows, uint32_t* __restrict__ ptr) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i <= n; i += stride) {
    const uint32_t hi = (i == n) ? nrows : key[i];
    const uint32_t lo = (i == 0) ? 0u : key[i - 1] + 1;
    for (uint32_t r = lo; r <= hi && r <= nrows; ++r) ptr[r] = uint32_t(i);
  }
}

void csr_transpose(fq_ctx* ctx, const fq_csr* a, fq_csr* out) {
  FQ_REQUIRE(a->row_begin == 0 && a->row_end == a->nrows, "transpose needs a fully held matrix");
  const uint32_t nnz = uint32_t(a->nnz), nrows = uint32_t(a->nrows), ncols = uint32_t(a->ncols);
  out->nrows = ncols;
  out->ncols = nrows;
  out->row_begin = 0;
  out->row_end = ncols;
  out->nnz = nnz;
  out->row_ptr.alloc(size_t(ncols) + 1);
  out->col_idx.alloc(nnz ? nnz : 1);
  out->values.alloc(nnz ? nnz : 1);
  if (nnz == 0) {
    FQ_CUDA(cudaMemsetAsync(out->row_ptr.p, 0, out->row_ptr.bytes(), ctx->stream));
    return;
  }
  const int block = 256;
  DevBuf<uint32_t> row_of(nnz), id(nnz), keys(nnz), keys_alt(nnz), id_alt(nnz);
  expand_rows_kernel<<<grid_for(nrows, block, ctx->sm_count), block, 0, ctx->stream>>>(a->row_ptr.p, nrows, row_of.p, id.p);
  FQ_CUDA(cudaMemcpyAsync(keys.p, a->col_idx.p, size_t(nnz) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  cub::DoubleBuffer<uint32_t> dk(keys.p, keys_alt.p), dv(id.p, id_alt.p);
  int end_bit = 1;
  while ((1ull << end_bit) < ncols) ++end_bit;
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, int64_t(nnz), 0, end_bit, ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, int64_t(nnz), 0, end_bit, ctx->stream));  // stable
  transpose_fill_kernel<<<grid_for(nnz, block, ctx->sm_count), block, 0, ctx->stream>>>(dv.Current(), row_of.p, a->values.p, nnz,
                                                                                      out->col_idx.p, out->values.p);
  lower_bound_ptr_kernel<<<grid_for(size_t(nnz) + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(dk.Current(), nnz, ncols,
                                                                                                   out->row_ptr.p);
  fq_count_launch(ctx, 8);
  FQ_CUDA(cudaGetLastError());
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
}

// out = [[a00, s01 * a01], [a10, a11]] with a01: n0 x n1, a10: n1 x n0
__global__ void block2x2_ptr_kernel(const uint32_t* __restrict__ p00, const uint32_t* __restrict__ p01,
                                    const uint32_t* __restrict__ p10, const uint32_t* __restrict__ p11, uint32_t n0, uint32_t n1,
                                    uint32_t* __restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t top = p00[n0] + p01[n0];
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= n0 + n1; r += stride)
    out[r] = r <= n0 ? p00[r] + p01[r] : top + p10[r - n0] + p11[r - n0];
}
__global__ void block2x2_fill_kernel(const uint32_t* __restrict__ pa, const uint32_t* __restrict__ ca, const double* __restrict__ va,
                                     double sa, uint32_t shift_a, const uint32_t* __restrict__ pb,
                                     const uint32_t* __restrict__ cb, const double* __restrict__ vb, double sb, uint32_t shift_b,
                                     uint32_t nrows, uint32_t row_shift, const uint32_t* __restrict__ out_ptr,
                                     uint32_t* __restrict__ out_col, double* __restrict__ out_val) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    uint32_t o = out_ptr[r + row_shift];
    for (uint32_t p = pa[r]; p < pa[r + 1]; ++p, ++o) out_col[o] = ca[p] + shift_a, out_val[o] = sa < 0 ? -va[p] : va[p];
    for (uint32_t p = pb[r]; p < pb[r + 1]; ++p, ++o) out_col[o] = cb[p] + shift_b, out_val[o] = sb < 0 ? -vb[p] : vb[p];
  }
}

void csr_block2x2(fq_ctx* ctx, const fq_csr* a00, const fq_csr* a01, double s01, const fq_csr* a10, const fq_csr* a11,
                  fq_csr* out, double s00) {
  const size_t n0 = a00->nrows, n1 = a11->nrows;
  FQ_REQUIRE(a00->ncols == n0 && a11->ncols == n1 && a01->nrows == n0 && a01->ncols == n1 && a10->nrows == n1 && a10->ncols == n0,
             "block shapes do not fit");
  for (const fq_csr* b : {a00, a01, a10, a11}) FQ_REQUIRE(b->row_begin == 0 && b->row_end == b->nrows, "blocks must be fully held");
  const size_t nnz = a00->nnz + a01->nnz + a10->nnz + a11->nnz;
  FQ_REQUIRE(nnz < (size_t(1) << 32) && n0 + n1 < (size_t(1) << 32), "block matrix too large for 32-bit indices");
  out->nrows = out->ncols = n0 + n1;
  out->row_begin = 0;
  out->row_end = n0 + n1;
  out->nnz = nnz;
  out->row_ptr.alloc(n0 + n1 + 1);
  out->col_idx.alloc(nnz ? nnz : 1);
  out->values.alloc(nnz ? nnz : 1);
  const int block = 256;
  block2x2_ptr_kernel<<<grid_for(n0 + n1 + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(
      a00->row_ptr.p, a01->row_ptr.p, a10->row_ptr.p, a11->row_ptr.p, uint32_t(n0), uint32_t(n1), out->row_ptr.p);
  if (n0)
    block2x2_fill_kernel<<<grid_for(n0, block, ctx->sm_count), block, 0, ctx->stream>>>(
        a00->row_ptr.p, a00->col_idx.p, a00->values.p, s00, 0u, a01->row_ptr.p, a01->col_idx.p, a01->values.p, s01, uint32_t(n0),
        uint32_t(n0), 0u, out->row_ptr.p, out->col_idx.p, out->values.p);
  if (n1)
    block2x2_fill_kernel<<<grid_for(n1, block, ctx->sm_count), block, 0, ctx->stream>>>(
        a10->row_ptr.p, a10->col_idx.p, a10->values.p, 1.0, 0u, a11->row_ptr.p, a11->col_idx.p, a11->values.p, 1.0, uint32_t(n0),
        uint32_t(n1), uint32_t(n0), out->row_ptr.p, out->col_idx.p, out->values.p);
  fq_count_launch(ctx, 3);
  FQ_CUDA(cudaGetLastError());
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- restriction to a subset of rows / columns -----------------------------------------------------------------
// RelativeWhitneyComplex::assemble (formoniq/src/whitney_complex.rs:620-624) forms E_test^T A E_trial with the 0/1
// inclusion matrices of the interior DOFs through two sparse products; since every product has a single term
// A_ij * 1 * 1 the result is the sub-matrix A[interior rows, interior cols] — here an index compaction.
__global__ void restrict_colmap_kernel(const uint32_t* __restrict__ keep, uint32_t n, uint32_t* __restrict__ col_map) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) col_map[keep[i]] = i;
}
__global__ void restrict_count_kernel(const uint32_t* __restrict__ rows_keep, uint32_t nr, const uint32_t* __restrict__ row_ptr,
                                      const uint32_t* __restrict__ col_idx, const uint32_t* __restrict__ col_map,
                                      uint32_t* __restrict__ count) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= nr; i += stride) {
    uint32_t c = 0;
    if (i < nr) {
      const uint32_t r = rows_keep[i];
      for (uint32_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) c += col_map[col_idx[p]] != 0xFFFFFFFFu;
    }
    count[i] = c;
  }
}
__global__ void restrict_fill_kernel(const uint32_t* __restrict__ rows_keep, uint32_t nr, const uint32_t* __restrict__ row_ptr,
                                     const uint32_t* __restrict__ col_idx, const double* __restrict__ val,
                                     const uint32_t* __restrict__ col_map, const uint32_t* __restrict__ out_ptr,
                                     uint32_t* __restrict__ out_col, double* __restrict__ out_val) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += stride) {
    const uint32_t r = rows_keep[i];
    uint32_t o = out_ptr[i];
    for (uint32_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
      const uint32_t c = col_map[col_idx[p]];
      if (c != 0xFFFFFFFFu) out_col[o] = c, out_val[o] = val[p], ++o;
    }
  }
}

// rows_keep / cols_keep: ascending device index lists
void csr_restrict(fq_ctx* ctx, const fq_csr* a, const uint32_t* rows_keep, size_t nr, const uint32_t* cols_keep, size_t nc,
                  fq_csr* out) {
  FQ_REQUIRE(a->row_begin == 0 && a->row_end == a->nrows, "restriction needs a fully held matrix");
  const int block = 256;
  DevBuf<uint32_t> col_map(a->ncols ? a->ncols : 1), count(nr + 1);
  FQ_CUDA(cudaMemsetAsync(col_map.p, 0xFF, col_map.bytes(), ctx->stream));
  if (nc) restrict_colmap_kernel<<<grid_for(nc, block, ctx->sm_count), block, 0, ctx->stream>>>(cols_keep, uint32_t(nc), col_map.p);
  restrict_count_kernel<<<grid_for(nr + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(rows_keep, uint32_t(nr), a->row_ptr.p,
                                                                                          a->col_idx.p, col_map.p, count.p);
  out->nrows = nr;
  out->ncols = nc;
  out->row_begin = 0;
  out->row_end = nr;
  out->row_ptr.alloc(nr + 1);
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, count.p, out->row_ptr.p, int64_t(nr + 1), ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, count.p, out->row_ptr.p, int64_t(nr + 1), ctx->stream));
  uint32_t nnz = 0;
  FQ_CUDA(cudaMemcpyAsync(&nnz, out->row_ptr.p + nr, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  out->nnz = nnz;
  out->col_idx.alloc(nnz ? nnz : 1);
  out->values.alloc(nnz ? nnz : 1);
  if (nr)
    restrict_fill_kernel<<<grid_for(nr, block, ctx->sm_count), block, 0, ctx->stream>>>(
        rows_keep, uint32_t(nr), a->row_ptr.p, a->col_idx.p, a->values.p, col_map.p, out->row_ptr.p, out->col_idx.p, out->values.p);
  fq_count_launch(ctx, 5);
  FQ_CUDA(cudaGetLastError());
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- C = A + B on the union pattern (nalgebra-sparse `&a + &b`: HilbertComplex::hdif_gram, whitney_complex.rs:180-183)
// Rows are merged by column (both inputs are sorted); an entry present in both operands is a_ij + b_ij, one present in
// a single operand is copied, explicit zeros stay in the pattern — entry for entry what the reference's spadd produces.
__global__ void add_count_kernel(const uint32_t* __restrict__ arp, const uint32_t* __restrict__ aci,
                                 const uint32_t* __restrict__ brp, const uint32_t* __restrict__ bci, uint32_t nrows,
                                 uint32_t* __restrict__ len) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= nrows; r += stride) {
    if (r == nrows) {
      len[r] = 0;
      continue;
    }
    uint32_t p = arp[r], q = brp[r], n = 0;
    const uint32_t pe = arp[r + 1], qe = brp[r + 1];
    while (p < pe && q < qe) {
      const uint32_t ca = aci[p], cb = bci[q];
      p += ca <= cb;
      q += cb <= ca;
      ++n;
    }
    len[r] = n + (pe - p) + (qe - q);
  }
}
__global__ void add_fill_kernel(const uint32_t* __restrict__ arp, const uint32_t* __restrict__ aci, const double* __restrict__ av,
                                const uint32_t* __restrict__ brp, const uint32_t* __restrict__ bci, const double* __restrict__ bv,
                                uint32_t nrows, const uint32_t* __restrict__ crp, uint32_t* __restrict__ cci,
                                double* __restrict__ cv) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    uint32_t p = arp[r], q = brp[r], o = crp[r];
    const uint32_t pe = arp[r + 1], qe = brp[r + 1];
    while (p < pe || q < qe) {
      const uint32_t ca = p < pe ? aci[p] : 0xFFFFFFFFu, cb = q < qe ? bci[q] : 0xFFFFFFFFu;
      if (ca == cb) {
        cci[o] = ca;
        cv[o] = __dadd_rn(av[p], bv[q]);
        ++p, ++q;
      } else if (ca < cb) {
        cci[o] = ca;
        cv[o] = av[p];
        ++p;
      } else {
        cci[o] = cb;
        cv[o] = bv[q];
        ++q;
      }
      ++o;
    }
  }
}
void csr_add(fq_ctx* ctx, const fq_csr* a, const fq_csr* b, fq_csr* out) {
  FQ_REQUIRE(a->nrows == b->nrows && a->ncols == b->ncols && a->row_begin == b->row_begin && a->row_end == b->row_end,
             "csr_add: the operands must have the same shape and row range");
  const uint32_t nrows = uint32_t(a->row_end - a->row_begin);
  out->nrows = a->nrows;
  out->ncols = a->ncols;
  out->row_begin = a->row_begin;
  out->row_end = a->row_end;
  out->row_ptr.alloc(size_t(nrows) + 1);
  const int block = 256;
  add_count_kernel<<<grid_for(size_t(nrows) + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(
      a->row_ptr.p, a->col_idx.p, b->row_ptr.p, b->col_idx.p, nrows, out->row_ptr.p);
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, out->row_ptr.p, out->row_ptr.p, int64_t(nrows) + 1, ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, out->row_ptr.p, out->row_ptr.p, int64_t(nrows) + 1, ctx->stream));
  uint32_t nnz = 0;
  FQ_CUDA(cudaMemcpyAsync(&nnz, out->row_ptr.p + nrows, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  out->nnz = nnz;
  out->col_idx.alloc(nnz ? nnz : 1);
  out->values.alloc(nnz ? nnz : 1);
  if (nnz)
    add_fill_kernel<<<grid_for(nrows, block, ctx->sm_count), block, 0, ctx->stream>>>(
        a->row_ptr.p, a->col_idx.p, a->values.p, b->row_ptr.p, b->col_idx.p, b->values.p, nrows, out->row_ptr.p, out->col_idx.p,
        out->values.p);
  fq_count_launch(ctx, 4);
  FQ_CUDA(cudaGetLastError());
}

__global__ void row_abs_sum_kernel(const uint32_t* __restrict__ rp, const double* __restrict__ v, uint32_t nrows, double* __restrict__ y) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    double acc = 0.0;
    for (uint32_t p = rp[r]; p < rp[r + 1]; ++p) acc += fabs(v[p]);
    y[r] = acc;
  }
}
void csr_row_abs_sums(fq_ctx* ctx, const fq_csr* a, double* y) {
  const uint32_t nrows = uint32_t(a->row_end - a->row_begin);
  if (!nrows) return;
  row_abs_sum_kernel<<<grid_for(nrows, 256, ctx->sm_count), 256, 0, ctx->stream>>>(a->row_ptr.p, a->values.p, nrows, y);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

}  // namespace fq
