// krylov.cu — the reference's Krylov drivers on device vectors: cg (iterative/src/krylov.rs:48-95) and preconditioned
// MINRES (:113-211), written once over an abstract operator / preconditioner / inner product so that the same loops
// serve
//   * one assembled matrix with the Identity or Jacobi preconditioner (iterative/src/precond.rs),
//   * the mixed Hodge-Laplace KKT operator with the AFW block preconditioner diag(hdif_gram(k-1)^-1, hdif_gram(k)^-1)
//     (problems/elliptic.rs:29-47: the reference factorises the two Gram matrices with a sparse Cholesky; here each
//     block is an inner Jacobi-preconditioned CG solve to a tolerance far below the outer one),
//   * the row-partitioned KKT operator of a multi-GPU run: SpMVs on the rank's row blocks after a halo exchange of the
//     two column windows, inner products completed by an all-reduce of one scalar (callbacks supplied by the host
//     side, which owns the communicator: torch.distributed / NCCL).
// Scalars live on the host exactly as in the reference; every vector operation is a kernel from blas1.cu / spmv.cu.
#include <cmath>
#include <cstdlib>
#include <functional>

#include "internal.hpp"

namespace fq {

namespace {
struct Work {
  fq_ctx* ctx;
  size_t n;
  std::vector<DevBuf<double>> bufs;
  double* get() {
    bufs.emplace_back(n ? n : 1);
    FQ_CUDA(cudaMemsetAsync(bufs.back().p, 0, (n ? n : 1) * sizeof(double), ctx->stream));
    return bufs.back().p;
  }
};
void copy(fq_ctx* ctx, double* dst, const double* src, size_t n) {
  if (n) FQ_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
}
}  // namespace

KrylovReport cg_core(fq_ctx* ctx, size_t n, const KrylovOps& ops, const double* b, double rtol, size_t max_iters, double* x) {
  KrylovReport rep;
  if (n) FQ_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), ctx->stream));
  const double b_norm = std::sqrt(ops.dot(b, b));
  if (b_norm == 0.0) {
    rep.converged = true;
    return rep;
  }
  Work w{ctx, n, {}};
  double *r = w.get(), *z = w.get(), *p = w.get(), *ap = w.get();
  copy(ctx, r, b, n);
  ops.precond(r, z);
  copy(ctx, p, z, n);
  double rz = ops.dot(r, z);
  for (;;) {
    rep.residual = std::sqrt(ops.dot(r, r)) / b_norm;
    rep.converged = rep.residual <= rtol;
    if (rep.converged || rep.iters >= max_iters) break;
    ops.apply(p, ap);
    const double alpha = rz / ops.dot(p, ap);
    vec_axpy(ctx, x, alpha, p, n);
    vec_axpy(ctx, r, -alpha, ap, n);
    ops.precond(r, z);
    const double rz_next = ops.dot(r, z);
    const double beta = rz_next / rz;
    vec_scale(ctx, p, beta, n);
    vec_axpy(ctx, p, 1.0, z, n);
    rz = rz_next;
    ++rep.iters;
  }
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return rep;
}

KrylovReport minres_core(fq_ctx* ctx, size_t n, const KrylovOps& ops, const double* b, double rtol, size_t max_iters,
                         double* x) {
  KrylovReport rep;
  const double eps = 2.220446049250313e-16;
  Work w{ctx, n, {}};
  double *r1 = w.get(), *r2 = w.get(), *y = w.get(), *v = w.get(), *yn = w.get();
  double *wv = w.get(), *w2 = w.get(), *wnew = w.get();
  copy(ctx, r1, b, n);
  ops.precond(r1, y);
  const double beta1_sq = ops.dot(r1, y);
  if (n) FQ_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), ctx->stream));
  if (beta1_sq <= 0.0) {
    rep.converged = true;
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
    return rep;
  }
  const double beta1 = std::sqrt(beta1_sq);
  double oldb = 0.0, beta = beta1, dbar = 0.0, epsln = 0.0, phibar = beta1, cs = -1.0, sn = 0.0;
  copy(ctx, r2, r1, n);
  rep.residual = 1.0;
  while (rep.iters < max_iters) {
    ++rep.iters;
    copy(ctx, v, y, n);
    vec_scale(ctx, v, 1.0 / beta, n);
    ops.apply(v, yn);
    if (rep.iters >= 2) vec_axpy(ctx, yn, -beta / oldb, r1, n);
    const double alfa = ops.dot(v, yn);
    vec_axpy(ctx, yn, -alfa / beta, r2, n);
    std::swap(r1, r2);   // r1 = r2
    std::swap(r2, yn);   // r2 = y_next (yn now holds the old r1: scratch)
    ops.precond(r2, y);
    oldb = beta;
    beta = std::sqrt(std::max(ops.dot(r2, y), 0.0));
    const double oldeps = epsln;
    const double delta = cs * dbar + sn * alfa;
    const double gbar = sn * dbar - cs * alfa;
    epsln = sn * beta;
    dbar = -cs * beta;
    const double gamma = std::max(std::sqrt(gbar * gbar + beta * beta), eps);
    cs = gbar / gamma;
    sn = beta / gamma;
    const double phi = cs * phibar;
    phibar *= sn;
    copy(ctx, wnew, v, n);
    vec_axpy(ctx, wnew, -oldeps, w2, n);
    vec_axpy(ctx, wnew, -delta, wv, n);
    vec_scale(ctx, wnew, 1.0 / gamma, n);
    // w2 = w ; w = wnew
    double* t = w2;
    w2 = wv;
    wv = wnew;
    wnew = t;
    vec_axpy(ctx, x, phi, wv, n);
    rep.residual = phibar / beta1;
    if (rep.residual <= rtol) {
      rep.converged = true;
      break;
    }
  }
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return rep;
}

// One (or a few) iterations of a device-resident loop as an executable CUDA graph; nullptr when capture is not possible
// (per-kernel timing on, FQ_KRYLOV_NO_GRAPH) — the caller then launches the kernels directly.  The capture runs on a
// private stream: the context's stream may be the legacy default stream, which cannot be captured, and a graph can be
// launched into any stream afterwards.
namespace {
template <class F>
cudaGraphExec_t capture_iterations(fq_ctx* ctx, F&& body) {
  if (ctx->timing || std::getenv("FQ_KRYLOV_NO_GRAPH")) return nullptr;
  cudaStream_t cap = nullptr;
  if (cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  cudaStream_t run = ctx->stream;
  const int64_t launches_before = ctx->launches;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
    ctx->stream = cap;
    try {
      body();
    } catch (...) {
      ctx->stream = run;
      cudaStreamEndCapture(cap, &graph);
      if (graph) cudaGraphDestroy(graph);
      cudaStreamDestroy(cap);
      (void)cudaGetLastError();
      throw;
    }
    ctx->stream = run;
    if (cudaStreamEndCapture(cap, &graph) == cudaSuccess && graph) {
      if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
      cudaGraphDestroy(graph);
    }
  }
  ctx->launches = launches_before;  // the replays are counted, not the capture
  cudaStreamDestroy(cap);
  if (!exec) (void)cudaGetLastError();
  return exec;
}
}  // namespace

// ---- device-resident CG for one assembled matrix (Identity / Jacobi)
// The loop above synchronises with the host three times per iteration (one per inner product).  Here the scalars of
// the recurrence live in device memory, the updates read them there, and a `done` flag turns every update into a no-op
// once the stopping test of krylov.rs:62-66 fires: one iteration is a fixed sequence of kernels, captured once in a CUDA
// graph and replayed in batches, with one host look at the flag per batch.  The vector updates, the Jacobi product and the
// partial sums of the two inner products that follow them are ONE kernel, their final reduction and the scalar tail of the
// iteration another (blas1.cu: cg_fused_update): six launches per iteration.  Same reductions, same IEEE
// operations on the scalars as cg_core: the iterates, the iteration count and the residual are the same bits.
namespace {
// device scalars of the CG recurrence: st = {bb, rz, pap, rz_next, rr, residual, beta}, iters, flags = {done, converged}
struct CgState {
  double st[7];
  unsigned long long iters;
  int flags[2];
};
__global__ void cg_update_p_kernel(double* __restrict__ p, const double* __restrict__ z, const CgState* __restrict__ s, size_t n) {
  if (s->flags[0]) return;
  const double beta = s->st[6];
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    p[i] = __dadd_rn(__dmul_rn(1.0, z[i]), __dmul_rn(p[i], beta));  // p *= beta, then p += 1.0 * z (krylov.rs:90-91)
}
}  // namespace

static KrylovReport cg_device(fq_ctx* ctx, fq_csr* a, int precond, const double* b, double rtol, size_t max_iters, double* x) {
  const size_t n = a->nrows;
  KrylovReport rep;
  spmv_prepare(ctx, a);
  if (precond == 1) csr_build_inv_diag(ctx, a);
  if (n) FQ_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), ctx->stream));
  Work w{ctx, n, {}};
  double *r = w.get(), *z = w.get(), *p = w.get(), *ap = w.get();
  DevBuf<double> partials(2 * vec_dot_scratch_doubles());
  double* partials2 = partials.p + vec_dot_scratch_doubles();
  DevBuf<CgState> state(1);
  FQ_CUDA(cudaMemsetAsync(state.p, 0, sizeof(CgState), ctx->stream));
  const double* jacobi = precond == 1 ? a->inv_diag.p : nullptr;
  double* st = state.p->st;
  vec_dot_device(ctx, b, b, n, partials.p, &st[0]);
  copy(ctx, r, b, n);
  if (jacobi)
    vec_mul_pointwise(ctx, z, jacobi, r, n);
  else
    copy(ctx, z, r, n);
  copy(ctx, p, z, n);
  vec_dot_device(ctx, r, z, n, partials.p, &st[1]);
  vec_dot_device(ctx, r, r, n, partials.p, &st[4]);
  CgState h{};
  FQ_CUDA(cudaMemcpyAsync(&h, state.p, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  const double b_norm = std::sqrt(h.st[0]);
  if (b_norm == 0.0) {  // krylov.rs:54-57
    rep.converged = true;
    return rep;
  }
  // the stopping test before the first iteration (the fused tail of an iteration does the test of the next one)
  rep.residual = std::sqrt(h.st[4]) / b_norm;
  rep.converged = rep.residual <= rtol;
  if (rep.converged || max_iters == 0) return rep;
  const int grid = grid_for(n, 256, ctx->sm_count);
  auto iteration = [&]() {
    spmv_apply(ctx, a, p, ap);
    vec_dot_device(ctx, p, ap, n, partials.p, &st[2]);
    cg_fused_update(ctx, x, r, p, ap, z, jacobi, &st[1], &st[2], &state.p->flags[0], n, partials.p, partials2, st,
                    &state.p->iters, state.p->flags, rtol, max_iters);
    cg_update_p_kernel<<<grid, 256, 0, ctx->stream>>>(p, z, state.p, n);
    fq_count_launch(ctx);
  };
  // one iteration captured as a graph
  cudaGraphExec_t exec = capture_iterations(ctx, iteration);
  const int launches_per_iteration = 6;
  size_t batch = 4;
  for (;;) {
    for (size_t i = 0; i < batch; ++i) {
      if (exec) {
        FQ_CUDA(cudaGraphLaunch(exec, ctx->stream));
        fq_count_launch(ctx, launches_per_iteration);
      } else {
        iteration();
      }
    }
    FQ_CUDA(cudaMemcpyAsync(&h, state.p, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream));
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h.flags[0]) break;
    if (batch < 64) batch *= 2;
  }
  if (exec) cudaGraphExecDestroy(exec);
  FQ_CUDA(cudaGetLastError());
  rep.iters = size_t(h.iters);
  rep.residual = h.st[5];
  rep.converged = h.flags[1] != 0;
  return rep;
}

// ---- device-resident MINRES for one assembled matrix (Identity / Jacobi): minres_core above with its scalars (Lanczos
// coefficients, Givens rotation, residual estimate) in device memory.  The three-term recurrences rotate their vectors
// with period 3 (r1 <- r2 <- y_next, w2 <- w <- w_new), so THREE iterations are captured as one graph.
namespace {
struct MinresState {
  double beta1, oldb, beta, dbar, epsln, phibar, cs, sn;
  double alfa, t0, oldeps, delta, gamma, phi, residual;
  unsigned long long iters;
  int done, stop_next, converged;
};
__global__ void minres_begin_kernel(MinresState* s, unsigned long long max_iters) {
  if (s->done) return;
  if (s->stop_next || s->iters >= max_iters) {
    s->done = 1;
    return;
  }
  ++s->iters;
}
// v = y * (1 / beta)
__global__ void minres_v_kernel(double* __restrict__ v, const double* __restrict__ y, const MinresState* __restrict__ s, size_t n) {
  if (s->done) return;
  const double inv = __ddiv_rn(1.0, s->beta);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = __dmul_rn(y[i], inv);
}
// yn += (-beta / oldb) * r1   (from the second iteration on)
__global__ void minres_sub_r1_kernel(double* __restrict__ yn, const double* __restrict__ r1, const MinresState* __restrict__ s,
                                     size_t n) {
  if (s->done || s->iters < 2) return;
  const double c = __ddiv_rn(-s->beta, s->oldb);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) yn[i] = __dadd_rn(__dmul_rn(c, r1[i]), yn[i]);
}
// yn += (-alfa / beta) * r2
__global__ void minres_sub_r2_kernel(double* __restrict__ yn, const double* __restrict__ r2, const MinresState* __restrict__ s,
                                     size_t n) {
  if (s->done) return;
  const double c = __ddiv_rn(-s->alfa, s->beta);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) yn[i] = __dadd_rn(__dmul_rn(c, r2[i]), yn[i]);
}
// the scalar recurrences of one step (krylov.rs:160-190), t0 = <r2, y>
__global__ void minres_scalars_kernel(MinresState* s, double rtol) {
  if (s->done) return;
  const double eps = 2.220446049250313e-16;
  s->oldb = s->beta;
  s->beta = __dsqrt_rn(fmax(s->t0, 0.0));
  s->oldeps = s->epsln;
  s->delta = __dadd_rn(__dmul_rn(s->cs, s->dbar), __dmul_rn(s->sn, s->alfa));
  const double gbar = __dadd_rn(__dmul_rn(s->sn, s->dbar), -__dmul_rn(s->cs, s->alfa));
  s->epsln = __dmul_rn(s->sn, s->beta);
  s->dbar = __dmul_rn(-s->cs, s->beta);
  s->gamma = fmax(__dsqrt_rn(__dadd_rn(__dmul_rn(gbar, gbar), __dmul_rn(s->beta, s->beta))), eps);
  s->cs = __ddiv_rn(gbar, s->gamma);
  s->sn = __ddiv_rn(s->beta, s->gamma);
  s->phi = __dmul_rn(s->cs, s->phibar);
  s->phibar = __dmul_rn(s->phibar, s->sn);
  s->residual = __ddiv_rn(s->phibar, s->beta1);
  if (s->residual <= rtol) {
    s->converged = 1;
    s->stop_next = 1;
  }
}
// w_new = (v - oldeps * w2 - delta * w) / gamma ;  x += phi * w_new
__global__ void minres_w_x_kernel(double* __restrict__ wnew, const double* __restrict__ v, const double* __restrict__ w2,
                                  const double* __restrict__ wv, double* __restrict__ x, const MinresState* __restrict__ s,
                                  size_t n) {
  if (s->done) return;
  const double a = -s->oldeps, b = -s->delta, inv = __ddiv_rn(1.0, s->gamma), phi = s->phi;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    double t = v[i];
    t = __dadd_rn(__dmul_rn(a, w2[i]), t);
    t = __dadd_rn(__dmul_rn(b, wv[i]), t);
    t = __dmul_rn(t, inv);
    wnew[i] = t;
    x[i] = __dadd_rn(__dmul_rn(phi, t), x[i]);
  }
}
}  // namespace

static KrylovReport minres_device(fq_ctx* ctx, fq_csr* a, int precond, const double* b, double rtol, size_t max_iters, double* x) {
  const size_t n = a->nrows;
  KrylovReport rep;
  spmv_prepare(ctx, a);
  if (precond == 1) csr_build_inv_diag(ctx, a);
  Work w{ctx, n, {}};
  double *r1 = w.get(), *r2 = w.get(), *y = w.get(), *v = w.get(), *yn = w.get();
  double *wv = w.get(), *w2 = w.get(), *wnew = w.get();
  DevBuf<double> partials(vec_dot_scratch_doubles());
  DevBuf<MinresState> state(1);
  auto apply_precond = [&](const double* r, double* z) {
    if (precond == 0)
      copy(ctx, z, r, n);
    else
      vec_mul_pointwise(ctx, z, a->inv_diag.p, r, n);
  };
  FQ_CUDA(cudaMemsetAsync(state.p, 0, sizeof(MinresState), ctx->stream));
  copy(ctx, r1, b, n);
  apply_precond(r1, y);
  vec_dot_device(ctx, r1, y, n, partials.p, &state.p->t0);
  if (n) FQ_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), ctx->stream));
  MinresState h{};
  FQ_CUDA(cudaMemcpyAsync(&h, state.p, sizeof(MinresState), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h.t0 <= 0.0) {  // krylov.rs:127-131
    rep.converged = true;
    return rep;
  }
  h.beta1 = std::sqrt(h.t0);
  h.oldb = 0.0, h.beta = h.beta1, h.dbar = 0.0, h.epsln = 0.0, h.phibar = h.beta1, h.cs = -1.0, h.sn = 0.0, h.residual = 1.0;
  FQ_CUDA(cudaMemcpyAsync(state.p, &h, sizeof(MinresState), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));  // `h` is reused below
  copy(ctx, r2, r1, n);
  const int grid = grid_for(n, 256, ctx->sm_count);
  auto iteration = [&]() {
    minres_begin_kernel<<<1, 1, 0, ctx->stream>>>(state.p, (unsigned long long)max_iters);
    minres_v_kernel<<<grid, 256, 0, ctx->stream>>>(v, y, state.p, n);
    spmv_apply(ctx, a, v, yn);
    minres_sub_r1_kernel<<<grid, 256, 0, ctx->stream>>>(yn, r1, state.p, n);
    vec_dot_device(ctx, v, yn, n, partials.p, &state.p->alfa);
    minres_sub_r2_kernel<<<grid, 256, 0, ctx->stream>>>(yn, r2, state.p, n);
    std::swap(r1, r2);  // r1 = r2
    std::swap(r2, yn);  // r2 = y_next (yn now holds the old r1: scratch)
    apply_precond(r2, y);
    vec_dot_device(ctx, r2, y, n, partials.p, &state.p->t0);
    minres_scalars_kernel<<<1, 1, 0, ctx->stream>>>(state.p, rtol);
    minres_w_x_kernel<<<grid, 256, 0, ctx->stream>>>(wnew, v, w2, wv, x, state.p, n);
    double* t = w2;  // w2 = w ; w = w_new
    w2 = wv;
    wv = wnew;
    wnew = t;
    fq_count_launch(ctx, 6);
  };
  // the pointer roles come back after three iterations: capture three
  cudaGraphExec_t exec = capture_iterations(ctx, [&]() {
    iteration();
    iteration();
    iteration();
  });
  const int launches_per_iteration = 12;
  size_t batch = 2;  // in units of three iterations
  for (;;) {
    for (size_t i = 0; i < batch; ++i) {
      if (exec) {
        FQ_CUDA(cudaGraphLaunch(exec, ctx->stream));
        fq_count_launch(ctx, 3 * launches_per_iteration);
      } else {
        iteration();
        iteration();
        iteration();
      }
    }
    FQ_CUDA(cudaMemcpyAsync(&h, state.p, sizeof(MinresState), cudaMemcpyDeviceToHost, ctx->stream));
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h.done) break;
    if (batch < 32) batch *= 2;
  }
  if (exec) cudaGraphExecDestroy(exec);
  FQ_CUDA(cudaGetLastError());
  rep.iters = size_t(h.iters);
  rep.residual = h.residual;
  rep.converged = h.converged != 0;
  return rep;
}

// ---- one assembled matrix, Identity / Jacobi
static KrylovOps csr_ops(fq_ctx* ctx, fq_csr* a, int precond, size_t n) {
  spmv_prepare(ctx, a);
  if (precond == 1) csr_build_inv_diag(ctx, a);
  KrylovOps ops;
  ops.apply = [ctx, a](const double* x, double* y) { spmv_apply(ctx, a, x, y); };
  ops.precond = [ctx, a, precond, n](const double* r, double* z) {
    if (precond == 0)
      copy(ctx, z, r, n);
    else
      vec_mul_pointwise(ctx, z, a->inv_diag.p, r, n);
  };
  ops.dot = [ctx, n](const double* u, const double* v) { return vec_dot(ctx, u, v, n); };
  return ops;
}
KrylovReport krylov_cg(fq_ctx* ctx, fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters, fq_vec* x) {
  const size_t n = b->d.n;
  FQ_REQUIRE(a->nrows == a->ncols && a->row_begin == 0 && a->row_end == a->nrows, "cg needs a square, fully held matrix");
  FQ_REQUIRE(n == a->nrows && x->d.n == n, "cg: dimension mismatch");
  if (std::getenv("FQ_KRYLOV_HOST")) return cg_core(ctx, n, csr_ops(ctx, a, precond, n), b->d.p, rtol, max_iters, x->d.p);
  return cg_device(ctx, a, precond, b->d.p, rtol, max_iters, x->d.p);
}
KrylovReport krylov_minres(fq_ctx* ctx, fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters,
                           fq_vec* x) {
  const size_t n = b->d.n;
  FQ_REQUIRE(a->nrows == a->ncols && a->row_begin == 0 && a->row_end == a->nrows,
             "minres needs a square, fully held matrix");
  FQ_REQUIRE(n == a->nrows && x->d.n == n, "minres: dimension mismatch");
  if (std::getenv("FQ_KRYLOV_HOST")) return minres_core(ctx, n, csr_ops(ctx, a, precond, n), b->d.p, rtol, max_iters, x->d.p);
  return minres_device(ctx, a, precond, b->d.p, rtol, max_iters, x->d.p);
}

// ---- MINRES with a block-diagonal preconditioner of inner solves (the AFW preconditioner of elliptic.rs:29-47)
// blocks[i] acts on the segment [offsets[i], offsets[i+1]) of the vectors: z_i = blocks[i]^-1 r_i by Jacobi-CG to
// inner_rtol (nullptr: identity on the segment, e.g. the harmonic border).
KrylovReport krylov_minres_blockdiag(fq_ctx* ctx, fq_csr* a, int nblocks, fq_csr* const* blocks, const size_t* offsets,
                                     double inner_rtol, size_t inner_max_iters, const fq_vec* b, double rtol, size_t max_iters,
                                     fq_vec* x, size_t* inner_iters_total) {
  const size_t n = b->d.n;
  FQ_REQUIRE(a->nrows == a->ncols && a->row_begin == 0 && a->row_end == a->nrows, "minres needs a square, fully held matrix");
  FQ_REQUIRE(n == a->nrows && x->d.n == n, "minres: dimension mismatch");
  FQ_REQUIRE(nblocks >= 1 && offsets[0] == 0 && offsets[nblocks] == n, "block preconditioner: the segments must tile the vector");
  std::vector<KrylovOps> inner(static_cast<size_t>(nblocks));
  for (int i = 0; i < nblocks; ++i) {
    const size_t ni = offsets[i + 1] - offsets[i];
    if (!blocks[i]) continue;
    FQ_REQUIRE(blocks[i]->nrows == ni && blocks[i]->ncols == ni && blocks[i]->row_begin == 0 && blocks[i]->row_end == ni,
               "block preconditioner: block shape does not match its segment");
    inner[size_t(i)] = csr_ops(ctx, blocks[i], 1, ni);
  }
  KrylovOps ops = csr_ops(ctx, a, 0, n);
  size_t total_inner = 0;
  const bool host_scalars = std::getenv("FQ_KRYLOV_HOST") != nullptr;
  ops.precond = [&, ctx](const double* r, double* z) {
    for (int i = 0; i < nblocks; ++i) {
      const size_t off = offsets[i], ni = offsets[i + 1] - off;
      if (!blocks[i]) {
        copy(ctx, z + off, r + off, ni);
        continue;
      }
      const KrylovReport rep = host_scalars ? cg_core(ctx, ni, inner[size_t(i)], r + off, inner_rtol, inner_max_iters, z + off)
                                            : cg_device(ctx, blocks[i], 1, r + off, inner_rtol, inner_max_iters, z + off);
      total_inner += rep.iters;
    }
  };
  const KrylovReport rep = minres_core(ctx, n, ops, b->d.p, rtol, max_iters, x->d.p);
  if (inner_iters_total) *inner_iters_total = total_inner;
  return rep;
}

}  // namespace fq
