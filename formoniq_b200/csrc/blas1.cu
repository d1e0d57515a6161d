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
