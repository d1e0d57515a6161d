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
  return cg_core(ctx, n, csr_ops(ctx, a, precond, n), b->d.p, rtol, max_iters, x->d.p);
}
KrylovReport krylov_minres(fq_ctx* ctx, fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters,
                           fq_vec* x) {
  const size_t n = b->d.n;
  FQ_REQUIRE(a->nrows == a->ncols && a->row_begin == 0 && a->row_end == a->nrows,
             "minres needs a square, fully held matrix");
  FQ_REQUIRE(n == a->nrows && x->d.n == n, "minres: dimension mismatch");
  return minres_core(ctx, n, csr_ops(ctx, a, precond, n), b->d.p, rtol, max_iters, x->d.p);
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
  ops.precond = [&, ctx](const double* r, double* z) {
    for (int i = 0; i < nblocks; ++i) {
      const size_t off = offsets[i], ni = offsets[i + 1] - off;
      if (!blocks[i]) {
        copy(ctx, z + off, r + off, ni);
        continue;
      }
      const KrylovReport rep = cg_core(ctx, ni, inner[size_t(i)], r + off, inner_rtol, inner_max_iters, z + off);
      total_inner += rep.iters;
    }
  };
  const KrylovReport rep = minres_core(ctx, n, ops, b->d.p, rtol, max_iters, x->d.p);
  if (inner_iters_total) *inner_iters_total = total_inner;
  return rep;
}

}  // namespace fq
