// quadform.cu — element vectors of the source load on the device: SourceForm::element
// (formoniq/src/operators.rs:607-635) = CellQuadrature::integrate (operators.rs:247-261) of
// inner(f(x), W_sigma(x)) under the induced inner product Lambda^k g^{-1} (metric/src/tensor.rs:127-131,140-157),
// times the cell volume (regge/src/lib.rs:26-28).
//
// What stays on the host, as in the reference: the quadrature rule and the Whitney shape table are reference data built
// once per (dim, grade, rule) (SimplexQuadRule::grundmann_moeller, LsfSamples::whitney: formoniq_b200/quadrature.py), and
// the source field is a user closure (a `Section`) — the caller samples it at the nodes of every cell and passes
// [ncells][nnodes][C(n,k)] components in the cell's reference frame.  The device does the per-cell work
//   g -> g^{-1}, vol;  G = Lambda^k g^{-1} (k x k minors on colex k-subsets);
//   elvec[sigma] = vol * sum_q w_q * f_q^T G W_sigma(q)         (node-outer, dof-inner like the reference)
// and writes the element vectors straight into the staging buffer of the LinearForm plan (matfree.cu), whose per-DOF
// gather then produces the Galerkin vector: one kernel + one segmented reduction, no atomics.
#include "geometry.cuh"
#include "internal.hpp"

namespace fq {

constexpr int kQfMaxComp = 20;   // C(n,k) components of a k-form
constexpr int kQfMaxGrade = 6;

// determinant of the k x k minor of ginv on rows I, columns J (Gaussian elimination, partial pivoting)
__device__ inline double minor_det(const double* ginv, int n, const uint8_t* I, const uint8_t* J, int k) {
  if (k == 0) return 1.0;
  if (k == 1) return ginv[I[0] * n + J[0]];
  double a[kQfMaxGrade * kQfMaxGrade];
  for (int r = 0; r < k; ++r)
    for (int c = 0; c < k; ++c) a[r * k + c] = ginv[I[r] * n + J[c]];
  if (k == 2) return a[0] * a[3] - a[1] * a[2];
  double det = 1.0;
  for (int i = 0; i < k; ++i) {
    int p = i;
    double best = fabs(a[i * k + i]);
    for (int r = i + 1; r < k; ++r)
      if (fabs(a[r * k + i]) > best) best = fabs(a[r * k + i]), p = r;
    if (best == 0.0) return 0.0;
    if (p != i) {
      det = -det;
      for (int c = 0; c < k; ++c) {
        const double t = a[i * k + c];
        a[i * k + c] = a[p * k + c];
        a[p * k + c] = t;
      }
    }
    det *= a[i * k + i];
    for (int r = i + 1; r < k; ++r) {
      const double f = a[r * k + i] / a[i * k + i];
      for (int c = i + 1; c < k; ++c) a[r * k + c] -= f * a[i * k + c];
    }
  }
  return det;
}

__global__ void __launch_bounds__(128) source_elvec_kernel(int n, int k, int ncomp, int ndofs, int nnodes, size_t ncells,
                                                           const uint32_t* __restrict__ cell_edges, const double* __restrict__ lengths,
                                                           uint32_t edge_lo, const uint8_t* __restrict__ subsets /*[ncomp][k]*/,
                                                           const double* __restrict__ weights /*[nnodes]*/,
                                                           const double* __restrict__ shapes /*[nnodes][ndofs][ncomp]*/,
                                                           const double* __restrict__ samples /*[ncells][nnodes][ncomp]*/,
                                                           double* __restrict__ elvecs /*[ncells][ndofs]*/, int* __restrict__ err) {
  const int ne = n * (n + 1) / 2;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t c = size_t(blockIdx.x) * blockDim.x + threadIdx.x; c < ncells; c += stride) {
    double s[kMaxDim * (kMaxDim + 1) / 2];
    for (int e = 0; e < ne; ++e) s[e] = lengths[cell_edges[c * size_t(ne) + e] - edge_lo];
    double ginv[kMaxDim * kMaxDim];
    double vol;
    if (!geometry_generic(n, s, ginv, &vol)) {
      atomicExch(err, 1);
      continue;
    }
    double G[kQfMaxComp * kQfMaxComp];
    for (int i = 0; i < ncomp; ++i)
      for (int j = 0; j < ncomp; ++j) G[i * ncomp + j] = minor_det(ginv, n, subsets + i * k, subsets + j * k, k);
    double* out = elvecs + c * size_t(ndofs);
    for (int d = 0; d < ndofs; ++d) out[d] = 0.0;
    const double* f = samples + c * size_t(nnodes) * ncomp;
    for (int q = 0; q < nnodes; ++q) {
      const double* fq_ = f + size_t(q) * ncomp;
      for (int d = 0; d < ndofs; ++d) {
        const double* w = shapes + (size_t(q) * ndofs + d) * ncomp;
        double val = 0.0;  // left . (G right): source components against the measured shape components (tensor.rs:150-156)
        for (int i = 0; i < ncomp; ++i) {
          double m = 0.0;
          for (int j = 0; j < ncomp; ++j) m += G[i * ncomp + j] * w[j];
          val += fq_[i] * m;
        }
        out[d] += weights[q] * val;
      }
    }
    for (int d = 0; d < ndofs; ++d) out[d] = vol * out[d];
  }
}

// WeightedHodgeMass::element (formoniq/src/operators.rs:477-485): CellQuadrature::integrate_pair (operators.rs:266-290)
// of alpha(x) * inner(W_i, W_j), times the cell volume:  A[i][j] = vol * sum_q w_q * alpha_q * W_i(q)^T G W_j(q).
// One thread per cell writes its cell-major element matrix into the CSR's element slab; the K3 reduction of the slab path
// (assemble.cu) scatters it under the structural pattern of the mass of the same grade.
__global__ void __launch_bounds__(128) weighted_mass_kernel(int n, int k, int ncomp, int ndofs, int nnodes, size_t ncells,
                                                            const uint32_t* __restrict__ cell_edges, const double* __restrict__ lengths,
                                                            uint32_t edge_lo, const uint8_t* __restrict__ subsets,
                                                            const double* __restrict__ weights, const double* __restrict__ shapes,
                                                            const double* __restrict__ coeff /*[ncells][nnodes]*/,
                                                            double* __restrict__ slab /*[ncells][ndofs*ndofs]*/, int* __restrict__ err) {
  const int ne = n * (n + 1) / 2;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t c = size_t(blockIdx.x) * blockDim.x + threadIdx.x; c < ncells; c += stride) {
    double s[kMaxDim * (kMaxDim + 1) / 2];
    for (int e = 0; e < ne; ++e) s[e] = lengths[cell_edges[c * size_t(ne) + e] - edge_lo];
    double ginv[kMaxDim * kMaxDim];
    double vol;
    if (!geometry_generic(n, s, ginv, &vol)) {
      atomicExch(err, 1);
      continue;
    }
    double G[kQfMaxComp * kQfMaxComp];
    for (int i = 0; i < ncomp; ++i)
      for (int j = 0; j < ncomp; ++j) G[i * ncomp + j] = minor_det(ginv, n, subsets + i * k, subsets + j * k, k);
    double* out = slab + c * size_t(ndofs) * ndofs;
    for (int d = 0; d < ndofs * ndofs; ++d) out[d] = 0.0;
    for (int q = 0; q < nnodes; ++q) {
      const double wa = coeff[c * size_t(nnodes) + q];
      for (int j = 0; j < ndofs; ++j) {
        const double* wj = shapes + (size_t(q) * ndofs + j) * ncomp;
        double gw[kQfMaxComp];  // G * W_j(q): the measured column (tensor.rs:150-155)
        for (int a = 0; a < ncomp; ++a) {
          double m = 0.0;
          for (int b = 0; b < ncomp; ++b) m += G[a * ncomp + b] * wj[b];
          gw[a] = m;
        }
        for (int i = 0; i < ndofs; ++i) {
          const double* wi = shapes + (size_t(q) * ndofs + i) * ncomp;
          double val = 0.0;
          for (int a = 0; a < ncomp; ++a) val += wi[a] * gw[a];
          out[i * ndofs + j] += weights[q] * (wa * val);
        }
      }
    }
    for (int d = 0; d < ndofs * ndofs; ++d) out[d] = vol * out[d];
  }
}

void weighted_mass_to_slab(fq_ctx* ctx, const fq_mesh* mesh, int grade, int nnodes, const double* h_weights,
                           const double* h_shapes, const double* h_coeff, double* d_slab) {
  const int n = mesh->dim;
  FQ_REQUIRE(grade >= 0 && grade <= n, "weighted mass: the grade must lie in [0, dim]");
  FQ_REQUIRE(nnodes >= 1, "weighted mass: a quadrature rule has at least one node");
  const int ncomp = int(binom(n, grade)), ndofs = nlocal(n, grade);
  if (ncomp > kQfMaxComp || grade > kQfMaxGrade) throw Error(FQ_ERR_UNSUPPORTED, "weighted mass: more than 20 form components");
  if (mesh->ncells == 0) return;
  std::vector<uint8_t> subsets;
  for (uint32_t m : colex_subsets(n, grade))
    for (int e : mask_elems(m)) subsets.push_back(uint8_t(e));
  if (subsets.empty()) subsets.push_back(0);
  const size_t nn = size_t(nnodes);
  DevBuf<uint8_t> d_subsets(subsets.size());
  DevBuf<double> d_weights(nn), d_shapes(nn * ndofs * ncomp), d_coeff(mesh->ncells * nn);
  DevBuf<int> d_err(1);
  FQ_CUDA(cudaMemcpyAsync(d_subsets.p, subsets.data(), subsets.size(), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaMemcpyAsync(d_weights.p, h_weights, d_weights.bytes(), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaMemcpyAsync(d_shapes.p, h_shapes, d_shapes.bytes(), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaMemcpyAsync(d_coeff.p, h_coeff, d_coeff.bytes(), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), ctx->stream));
  {
    ScopedSpan span(ctx, "qf_weighted_mass");
    const int block = 128;
    weighted_mass_kernel<<<grid_for(mesh->ncells, block, ctx->sm_count), block, 0, ctx->stream>>>(
        n, grade, ncomp, ndofs, nnodes, mesh->ncells, mesh->cell_faces[1].p, mesh->lengths.p, uint32_t(mesh->edge_lo),
        d_subsets.p, d_weights.p, d_shapes.p, d_coeff.p, d_slab, d_err.p);
    fq_count_launch(ctx);
    FQ_CUDA(cudaGetLastError());
  }
  int h = 0;
  FQ_CUDA(cudaMemcpyAsync(&h, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h) throw Error(FQ_ERR_DEGENERATE, "a cell metric is singular");
}

// Fills `elvecs` (device, [ncells][ndofs]) from host tables and host samples.
void source_element_vectors(fq_ctx* ctx, const fq_mesh* mesh, int grade, int nnodes, const double* h_weights,
                            const double* h_shapes, const double* h_samples, double* d_elvecs) {
  const int n = mesh->dim;
  FQ_REQUIRE(grade >= 0 && grade <= n, "source form: the grade must lie in [0, dim]");
  FQ_REQUIRE(nnodes >= 1, "source form: a quadrature rule has at least one node");
  FQ_REQUIRE(mesh->cell_offset == 0, "source form needs a fully held mesh");
  const int ncomp = int(binom(n, grade)), ndofs = nlocal(n, grade);
  if (ncomp > kQfMaxComp || grade > kQfMaxGrade) throw Error(FQ_ERR_UNSUPPORTED, "source form: more than 20 form components");
  if (mesh->ncells == 0) return;
  std::vector<uint8_t> subsets;
  for (uint32_t m : colex_subsets(n, grade))
    for (int e : mask_elems(m)) subsets.push_back(uint8_t(e));
  if (subsets.empty()) subsets.push_back(0);
  DevBuf<uint8_t> d_subsets(subsets.size());
  const size_t nn = size_t(nnodes);
  DevBuf<double> d_weights(nn), d_shapes(nn * ndofs * ncomp), d_samples(mesh->ncells * nn * ncomp);
  DevBuf<int> d_err(1);
  FQ_CUDA(cudaMemcpyAsync(d_subsets.p, subsets.data(), subsets.size(), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaMemcpyAsync(d_weights.p, h_weights, d_weights.bytes(), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaMemcpyAsync(d_shapes.p, h_shapes, d_shapes.bytes(), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaMemcpyAsync(d_samples.p, h_samples, d_samples.bytes(), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), ctx->stream));
  {
    ScopedSpan span(ctx, "qf_source_elvec");
    const int block = 128;
    source_elvec_kernel<<<grid_for(mesh->ncells, block, ctx->sm_count), block, 0, ctx->stream>>>(
        n, grade, ncomp, ndofs, nnodes, mesh->ncells, mesh->cell_faces[1].p, mesh->lengths.p, uint32_t(mesh->edge_lo),
        d_subsets.p, d_weights.p, d_shapes.p, d_samples.p, d_elvecs, d_err.p);
    fq_count_launch(ctx);
    FQ_CUDA(cudaGetLastError());
  }
  int h = 0;
  FQ_CUDA(cudaMemcpyAsync(&h, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h) throw Error(FQ_ERR_DEGENERATE, "a cell metric is singular");
}

}  // namespace fq
