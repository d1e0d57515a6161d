// spmv.cu — K4: CSR SpMV y = A x  (iterative/src/operator.rs:5-14).
//
// "CSR-stream" (stream.cuh): a CTA owns a block of consecutive rows whose
// non-zeros fit a shared-memory chunk.  Phase 1 streams values / column indices
// fully coalesced with 8 independent gathers of x in flight per thread and
// stages the products v*x[col] in shared memory; phase 2 gives every row to one
// thread, which adds the row's products left to right.  That is the
// reference's serial order (one accumulator per row, ascending columns, mul
// then add, no FMA), so y is bit-identical to the CPU path while the HBM
// traffic stays the algorithmic 12 B/nnz + vectors.
#include "internal.hpp"
#include "stream.cuh"

namespace fq {

struct SpmvPolicy {
  static constexpr bool kHasValues = true;
  static constexpr bool kCustomSrc = false;
  double* __restrict__ y;
  __device__ __forceinline__ double load(uint32_t) const { return 0.0; }
  __device__ __forceinline__ void store(uint32_t row, double sum, bool) const { y[row] = sum; }
};

// SpMV fused with the halo exchange: x lives distributed over the ranks' windows; columns below / above this rank's
// owned range are loaded straight from the neighbouring GPU's memory (CUDA IPC mapping, NVLink 5 / NVSwitch P2P
// loads issued by the gather itself), so the transfer overlaps the row blocks that do not need it — there is no
// separate exchange step and no staging copy.  Pointers are pre-offset so that ptr[global column] is valid.
struct SpmvPeerPolicy {
  static constexpr bool kHasValues = true;
  static constexpr bool kCustomSrc = true;
  double* __restrict__ y;
  const double* own;    // this rank's window
  const double* lower;  // rank - 1's window (columns < own_lo)
  const double* upper;  // rank + 1's window (columns >= own_hi)
  uint32_t own_lo, own_hi;
  __device__ __forceinline__ double load(uint32_t col) const {
    const double* p = col < own_lo ? lower : (col >= own_hi ? upper : own);
    return p[col];  // plain ld.global: peer memory is not read through the non-coherent path
  }
  __device__ __forceinline__ void store(uint32_t row, double sum, bool) const { y[row] = sum; }
};

void spmv_apply_peer(fq_ctx* ctx, const fq_csr* a, const double* own, const double* lower, const double* upper, size_t own_lo,
                     size_t own_hi, double* y) {
  FQ_REQUIRE(a->spmv_ready, "spmv_prepare was not called");
  if (a->nrowblocks == 0) return;
  ScopedSpan span(ctx, "k4_spmv_peer");
  stream_reduce(ctx, a->rowblocks.p, a->nrowblocks, a->row_ptr.p, a->col_idx.p, a->values.p, nullptr,
                SpmvPeerPolicy{y, own, lower, upper, uint32_t(own_lo), uint32_t(own_hi)});
}

// Stream-ordered flags in (peer-mapped) device memory: the producer of x publishes an epoch after its last write,
// the consumers wait for it before the fused SpMV reads x over NVLink.
__global__ void flag_signal_kernel(volatile double* flag, double value) {
  __threadfence_system();
  *flag = value;
  __threadfence_system();
}
__global__ void flag_wait_kernel(const volatile double* flag, double value, int* timeout) {
  long long spins = 0;
  while (*flag < value) {
    __nanosleep(200);
    if (++spins > (1ll << 24)) {  // ~3 s: never hang the device on a lost peer
      *timeout = 1;
      return;
    }
  }
  __threadfence_system();
}
void flag_signal(fq_ctx* ctx, double* flag, double value) {
  flag_signal_kernel<<<1, 1, 0, ctx->stream>>>(flag, value);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}
void flag_wait(fq_ctx* ctx, const double* flag, double value, int* d_timeout) {
  flag_wait_kernel<<<1, 1, 0, ctx->stream>>>(flag, value, d_timeout);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

void spmv_prepare(fq_ctx* ctx, fq_csr* a) {
  if (a->spmv_ready) return;
  const size_t nrows = a->row_end - a->row_begin;
  stream_build_blocks(ctx, a->row_ptr.p, nrows, a->nnz, a->rowblocks, a->nrowblocks);
  a->spmv_ready = true;
}

void spmv_apply(fq_ctx* ctx, const fq_csr* a, const double* x, double* y) {
  FQ_REQUIRE(a->spmv_ready, "spmv_prepare was not called");
  if (a->nrowblocks == 0) return;
  ScopedSpan span(ctx, "k4_spmv");
  stream_reduce(ctx, a->rowblocks.p, a->nrowblocks, a->row_ptr.p, a->col_idx.p, a->values.p, x, SpmvPolicy{y});
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
