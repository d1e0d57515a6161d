// spmv.cu — K4: CSR SpMV y = A x  (iterative/src/operator.rs:5-14).
//
// "CSR-stream": a CTA owns a block of consecutive rows whose non-zeros fit a
// shared-memory chunk.  Phase 1 streams values / column indices fully
// coalesced and stages the products v*x[col] in shared memory; phase 2 gives
// every row to one thread, which adds the row's products left to right.  That
// is the reference's serial order (one accumulator per row, ascending
// columns, mul then add, no FMA), so y is bit-identical to the CPU path while
// the HBM traffic stays the algorithmic 12 B/nnz + vectors.
#include "internal.hpp"

namespace fq {

constexpr int kSpmvThreads = 256;
constexpr int kSpmvChunk = 2048;  // products staged per CTA (16 KB)

__global__ void __launch_bounds__(kSpmvThreads) spmv_stream_kernel(const uint32_t* __restrict__ rowblocks,
                                                                    const uint32_t* __restrict__ row_ptr,
                                                                    const uint32_t* __restrict__ col_idx,
                                                                    const double* __restrict__ values,
                                                                    const double* __restrict__ x,
                                                                    double* __restrict__ y, uint32_t nblocks) {
  __shared__ double prod[kSpmvChunk];
  __shared__ double red[kSpmvThreads / 32];
  for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const uint32_t r0 = rowblocks[b], r1 = rowblocks[b + 1];
    const uint32_t p0 = row_ptr[r0], p1 = row_ptr[r1];
    if (p1 - p0 <= uint32_t(kSpmvChunk)) {
      for (uint32_t p = p0 + threadIdx.x; p < p1; p += kSpmvThreads)
        prod[p - p0] = __dmul_rn(values[p], __ldg(x + col_idx[p]));
      __syncthreads();
      for (uint32_t r = r0 + threadIdx.x; r < r1; r += kSpmvThreads) {
        const uint32_t b0 = row_ptr[r] - p0, b1 = row_ptr[r + 1] - p0;
        double acc = 0.0;
        for (uint32_t q = b0; q < b1; ++q) acc = __dadd_rn(acc, prod[q]);
        y[r] = acc;
      }
      __syncthreads();
    } else {
      // a single row longer than the chunk: block-wide strided sum (tree order)
      double acc = 0.0;
      for (uint32_t p = p0 + threadIdx.x; p < p1; p += kSpmvThreads)
        acc = __dadd_rn(acc, __dmul_rn(values[p], __ldg(x + col_idx[p])));
      for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, o));
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kSpmvThreads / 32; ++w) s = __dadd_rn(s, red[w]);
        y[r0] = s;
      }
      __syncthreads();
    }
  }
}

void spmv_prepare(fq_ctx* ctx, fq_csr* a) {
  if (a->spmv_ready) return;
  const size_t nrows = a->row_end - a->row_begin;
  std::vector<uint32_t> rp(nrows + 1);
  FQ_CUDA(cudaMemcpyAsync(rp.data(), a->row_ptr.p, (nrows + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<uint32_t> rb;
  rb.push_back(0);
  size_t r = 0;
  while (r < nrows) {
    size_t e = r + 1;  // at least one row per block
    while (e < nrows && e - r < size_t(kSpmvThreads) && rp[e + 1] - rp[r] <= uint32_t(kSpmvChunk)) ++e;
    rb.push_back(uint32_t(e));
    r = e;
  }
  a->nrowblocks = rb.size() - 1;
  a->rowblocks.alloc(rb.size());
  FQ_CUDA(cudaMemcpyAsync(a->rowblocks.p, rb.data(), rb.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  a->spmv_ready = true;
}

void spmv_apply(fq_ctx* ctx, const fq_csr* a, const double* x, double* y) {
  FQ_REQUIRE(a->spmv_ready, "spmv_prepare was not called");
  if (a->nrowblocks == 0) return;
  const int grid = int(std::min<size_t>(a->nrowblocks, size_t(ctx->sm_count) * 8));
  ScopedSpan span(ctx, "k4_spmv");
  spmv_stream_kernel<<<grid, kSpmvThreads, 0, ctx->stream>>>(a->rowblocks.p, a->row_ptr.p, a->col_idx.p, a->values.p, x,
                                                              y, uint32_t(a->nrowblocks));
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

// Inverse diagonal for the Jacobi preconditioner (iterative/src/precond.rs:113-121).
__global__ void inv_diag_kernel(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col_idx,
                                const double* __restrict__ values, uint32_t nrows, uint32_t row_begin,
                                double* __restrict__ inv_diag) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    double d = 1.0;
    for (uint32_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p)
      if (col_idx[p] == r + row_begin) d = __ddiv_rn(1.0, values[p]);
    inv_diag[r] = d;
  }
}

void csr_build_inv_diag(fq_ctx* ctx, fq_csr* a) {
  const size_t nrows = a->row_end - a->row_begin;
  if (a->inv_diag.n == nrows && nrows) return;
  a->inv_diag.alloc(nrows ? nrows : 1);
  if (!nrows) return;
  inv_diag_kernel<<<grid_for(nrows, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
      a->row_ptr.p, a->col_idx.p, a->values.p, uint32_t(nrows), uint32_t(a->row_begin), a->inv_diag.p);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

}  // namespace fq
