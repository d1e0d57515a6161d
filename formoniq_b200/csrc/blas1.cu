// blas1.cu — K5: the InnerProductSpace operations on device vectors
// (iterative/src/lib.rs:84-141): dot, scale, add_scaled, plus the pointwise
// product a diagonal preconditioner needs.  The dot product is a fixed-shape
// two-stage reduction, so it is deterministic run to run.
#include <algorithm>

#include "internal.hpp"

namespace fq {

constexpr int kRedThreads = 256;
constexpr int kRedBlocksMax = 1184;  // 148 SMs x 8

__global__ void __launch_bounds__(kRedThreads) dot_partial_kernel(const double* __restrict__ x,
                                                                   const double* __restrict__ y, size_t n,
                                                                   double* __restrict__ partial) {
  __shared__ double red[kRedThreads / 32];
  double acc = 0.0;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) acc = fma(x[i], y[i], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kRedThreads / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(kRedThreads) dot_final_kernel(const double* __restrict__ partial, int nparts,
                                                                 double* __restrict__ out) {
  __shared__ double red[kRedThreads];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += kRedThreads) acc += partial[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kRedThreads / 2; o > 0; o >>= 1) {
    if (int(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

double vec_dot(fq_ctx* ctx, const double* x, const double* y, size_t n) {
  if (n == 0) return 0.0;
  if (ctx->reduce_scratch.n < size_t(kRedBlocksMax) + 1) ctx->reduce_scratch.alloc(size_t(kRedBlocksMax) + 1);
  // the partials live in kRedBlocksMax slots and the result in the slot after them, whatever the SM count
  const int grid = std::min(grid_for(n, kRedThreads, ctx->sm_count, 8), kRedBlocksMax);
  dot_partial_kernel<<<grid, kRedThreads, 0, ctx->stream>>>(x, y, n, ctx->reduce_scratch.p);
  dot_final_kernel<<<1, kRedThreads, 0, ctx->stream>>>(ctx->reduce_scratch.p, grid, ctx->reduce_scratch.p + kRedBlocksMax);
  fq_count_launch(ctx, 2);
  FQ_CUDA(cudaMemcpyAsync(ctx->host_scalar, ctx->reduce_scratch.p + kRedBlocksMax, sizeof(double), cudaMemcpyDeviceToHost,
                          ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return *ctx->host_scalar;
}

// The reduction tree of one CTA of dot_partial_kernel, for kernels that fuse other work with a partial inner product and
// must produce the same bits as vec_dot: thread-local accumulator -> warp shuffle tree -> warp sums added in order.
__device__ __forceinline__ double dot_block_reduce(double acc, double* red) {
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kRedThreads / 32; ++w) s += red[w];
  __syncthreads();
  return s;  // valid in thread 0
}
// CG, fused middle of an iteration (krylov.rs:80-88): alpha = rz / pAp; x += alpha p; r -= alpha Ap; z = M^-1 r (Jacobi
// vector d, or z = r when d is null); partial sums of <r, z> and <r, r> in the grid and order of dot_partial_kernel.
// Nothing happens once *done is set.  scal = {rz, pap}.
__global__ void __launch_bounds__(kRedThreads) cg_fused_update_kernel(double* __restrict__ x, double* __restrict__ r,
                                                                       const double* __restrict__ p, const double* __restrict__ ap,
                                                                       double* __restrict__ z, const double* __restrict__ d,
                                                                       const double* __restrict__ rz, const double* __restrict__ pap,
                                                                       const int* __restrict__ done, size_t n,
                                                                       double* __restrict__ part_rz, double* __restrict__ part_rr) {
  __shared__ double red[kRedThreads / 32];
  if (*done) return;
  const double alpha = __ddiv_rn(*rz, *pap);
  const double nalpha = -alpha;
  double acc_rz = 0.0, acc_rr = 0.0;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    x[i] = __dadd_rn(__dmul_rn(alpha, p[i]), x[i]);
    const double ri = __dadd_rn(__dmul_rn(nalpha, ap[i]), r[i]);
    r[i] = ri;
    const double zi = d ? __dmul_rn(d[i], ri) : ri;
    z[i] = zi;
    acc_rz = fma(ri, zi, acc_rz);
    acc_rr = fma(ri, ri, acc_rr);
  }
  const double s0 = dot_block_reduce(acc_rz, red);
  const double s1 = dot_block_reduce(acc_rr, red);
  if (threadIdx.x == 0) {
    part_rz[blockIdx.x] = s0;
    part_rr[blockIdx.x] = s1;
  }
}
// final stage of two inner products at once (the tree of dot_final_kernel), then the scalar tail of the CG iteration:
// beta = rz_next / rz; rz = rz_next; ++iters; the stopping test of the next iteration (krylov.rs:62-66).
// st = {bb, rz, pap, rz_next, rr, residual, beta} (doubles), it = {iters (as two ints' worth: unsigned long long), ...}
__global__ void __launch_bounds__(kRedThreads) cg_fused_final_kernel(const double* __restrict__ part_rz,
                                                                      const double* __restrict__ part_rr, int nparts,
                                                                      double* __restrict__ st, unsigned long long* __restrict__ iters,
                                                                      int* __restrict__ flags, double rtol,
                                                                      unsigned long long max_iters) {
  __shared__ double red0[kRedThreads], red1[kRedThreads];
  if (flags[0]) return;
  double a0 = 0.0, a1 = 0.0;
  for (int i = threadIdx.x; i < nparts; i += kRedThreads) a0 += part_rz[i], a1 += part_rr[i];
  red0[threadIdx.x] = a0, red1[threadIdx.x] = a1;
  __syncthreads();
  for (int o = kRedThreads / 2; o > 0; o >>= 1) {
    if (int(threadIdx.x) < o) red0[threadIdx.x] += red0[threadIdx.x + o], red1[threadIdx.x] += red1[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double rz_next = red0[0], rr = red1[0];
    st[6] = __ddiv_rn(rz_next, st[1]);  // beta
    st[1] = rz_next;
    st[3] = rz_next;
    st[4] = rr;
    const unsigned long long k = ++*iters;
    const double residual = __ddiv_rn(__dsqrt_rn(rr), __dsqrt_rn(st[0]));
    st[5] = residual;
    flags[1] = residual <= rtol ? 1 : 0;
    if (flags[1] || k >= max_iters) flags[0] = 1;
  }
}
void cg_fused_update(fq_ctx* ctx, double* x, double* r, const double* p, const double* ap, double* z, const double* d,
                     const double* rz, const double* pap, const int* done, size_t n, double* part_rz, double* part_rr,
                     double* st, unsigned long long* iters, int* flags, double rtol, size_t max_iters) {
  const int grid = std::min(grid_for(n ? n : 1, kRedThreads, ctx->sm_count, 8), kRedBlocksMax);
  cg_fused_update_kernel<<<grid, kRedThreads, 0, ctx->stream>>>(x, r, p, ap, z, d, rz, pap, done, n, part_rz, part_rr);
  cg_fused_final_kernel<<<1, kRedThreads, 0, ctx->stream>>>(part_rz, part_rr, grid, st, iters, flags, rtol,
                                                            (unsigned long long)max_iters);
  fq_count_launch(ctx, 2);
}

// <x, y> by the same two-stage reduction, the result left in device memory (no host round trip): the device-resident
// Krylov loops of krylov.cu read their scalars from there.  Same grid and order as vec_dot, hence the same bits.
void vec_dot_device(fq_ctx* ctx, const double* x, const double* y, size_t n, double* d_partials, double* d_out) {
  const int grid = std::min(grid_for(n ? n : 1, kRedThreads, ctx->sm_count, 8), kRedBlocksMax);
  dot_partial_kernel<<<grid, kRedThreads, 0, ctx->stream>>>(x, y, n, d_partials);
  dot_final_kernel<<<1, kRedThreads, 0, ctx->stream>>>(d_partials, grid, d_out);
  fq_count_launch(ctx, 2);
}
size_t vec_dot_scratch_doubles() { return size_t(kRedBlocksMax); }

__global__ void scale_kernel(double* __restrict__ x, double alpha, size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = __dmul_rn(x[i], alpha);
}
// y <- alpha*x + y, in the reference's evaluation order (nalgebra axpy: a*x + b*y with b = 1)
__global__ void axpy_kernel(double* __restrict__ y, double alpha, const double* __restrict__ x, size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __dadd_rn(__dmul_rn(alpha, x[i]), y[i]);
}
__global__ void mul_pointwise_kernel(double* __restrict__ z, const double* __restrict__ d, const double* __restrict__ r,
                                     size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) z[i] = __dmul_rn(d[i], r[i]);
}

void vec_scale(fq_ctx* ctx, double* x, double alpha, size_t n) {
  if (!n) return;
  scale_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(x, alpha, n);
  fq_count_launch(ctx);
}
void vec_axpy(fq_ctx* ctx, double* y, double alpha, const double* x, size_t n) {
  if (!n) return;
  axpy_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(y, alpha, x, n);
  fq_count_launch(ctx);
}
void vec_mul_pointwise(fq_ctx* ctx, double* z, const double* d, const double* r, size_t n) {
  if (!n) return;
  mul_pointwise_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(z, d, r, n);
  fq_count_launch(ctx);
}

}  // namespace fq
