// krylov.cu — the reference's Krylov drivers run on device vectors:
// cg (iterative/src/krylov.rs:48-95) and preconditioned MINRES (:113-211).
// Scalars live on the host exactly as in the reference; every vector
// operation is a kernel from blas1.cu / spmv.cu.
#include <cmath>

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
void apply_precond(fq_ctx* ctx, fq_csr* a, int precond, double* z, const double* r, size_t n) {
  if (precond == 0)
    copy(ctx, z, r, n);
  else
    vec_mul_pointwise(ctx, z, a->inv_diag.p, r, n);
}
}  // namespace

KrylovReport krylov_cg(fq_ctx* ctx, fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters, fq_vec* x) {
  const size_t n = b->d.n;
  FQ_REQUIRE(a->nrows == a->ncols && a->row_begin == 0 && a->row_end == a->nrows, "cg needs a square, fully held matrix");
  FQ_REQUIRE(n == a->nrows && x->d.n == n, "cg: dimension mismatch");
  spmv_prepare(ctx, a);
  if (precond == 1) csr_build_inv_diag(ctx, a);
  KrylovReport rep;
  FQ_CUDA(cudaMemsetAsync(x->d.p, 0, n * sizeof(double), ctx->stream));
  const double b_norm = std::sqrt(vec_dot(ctx, b->d.p, b->d.p, n));
  if (b_norm == 0.0) {
    rep.converged = true;
    return rep;
  }
  Work w{ctx, n, {}};
  double *r = w.get(), *z = w.get(), *p = w.get(), *ap = w.get();
  copy(ctx, r, b->d.p, n);
  apply_precond(ctx, a, precond, z, r, n);
  copy(ctx, p, z, n);
  double rz = vec_dot(ctx, r, z, n);
  for (;;) {
    rep.residual = std::sqrt(vec_dot(ctx, r, r, n)) / b_norm;
    rep.converged = rep.residual <= rtol;
    if (rep.converged || rep.iters >= max_iters) break;
    spmv_apply(ctx, a, p, ap);
    const double alpha = rz / vec_dot(ctx, p, ap, n);
    vec_axpy(ctx, x->d.p, alpha, p, n);
    vec_axpy(ctx, r, -alpha, ap, n);
    apply_precond(ctx, a, precond, z, r, n);
    const double rz_next = vec_dot(ctx, r, z, n);
    const double beta = rz_next / rz;
    vec_scale(ctx, p, beta, n);
    vec_axpy(ctx, p, 1.0, z, n);
    rz = rz_next;
    ++rep.iters;
  }
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return rep;
}

KrylovReport krylov_minres(fq_ctx* ctx, fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters,
                           fq_vec* x) {
  const size_t n = b->d.n;
  FQ_REQUIRE(a->nrows == a->ncols && a->row_begin == 0 && a->row_end == a->nrows,
             "minres needs a square, fully held matrix");
  FQ_REQUIRE(n == a->nrows && x->d.n == n, "minres: dimension mismatch");
  spmv_prepare(ctx, a);
  if (precond == 1) csr_build_inv_diag(ctx, a);
  KrylovReport rep;
  const double eps = 2.220446049250313e-16;
  Work w{ctx, n, {}};
  double *r1 = w.get(), *r2 = w.get(), *y = w.get(), *v = w.get(), *yn = w.get();
  double *wv = w.get(), *w2 = w.get(), *wnew = w.get();
  copy(ctx, r1, b->d.p, n);
  apply_precond(ctx, a, precond, y, r1, n);
  const double beta1_sq = vec_dot(ctx, r1, y, n);
  FQ_CUDA(cudaMemsetAsync(x->d.p, 0, n * sizeof(double), ctx->stream));
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
    spmv_apply(ctx, a, v, yn);
    if (rep.iters >= 2) vec_axpy(ctx, yn, -beta / oldb, r1, n);
    const double alfa = vec_dot(ctx, v, yn, n);
    vec_axpy(ctx, yn, -alfa / beta, r2, n);
    std::swap(r1, r2);   // r1 = r2
    std::swap(r2, yn);   // r2 = y_next (yn now holds the old r1: scratch)
    apply_precond(ctx, a, precond, y, r2, n);
    oldb = beta;
    beta = std::sqrt(std::max(vec_dot(ctx, r2, y, n), 0.0));
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
    vec_axpy(ctx, x->d.p, phi, wv, n);
    rep.residual = phibar / beta1;
    if (rep.residual <= rtol) {
      rep.converged = true;
      break;
    }
  }
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return rep;
}

}  // namespace fq
