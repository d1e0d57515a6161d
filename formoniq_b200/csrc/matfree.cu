// matfree.cu — the matrix-free peer of assembly: ElementOperator (formoniq/src/matfree.rs:60-179).
//
//   y = sum_K P_K^T A_K P_K x  by gather, in the reference's two stages:
//   stage 1 (per cell)  local_K = A_K * (x gathered at the cell's trial DOFs)       matfree.rs:105-118
//   stage 2 (per DOF)   y_dof   = sum over FaceIncidence::face_cells(dof), in cell order, of local_K[position]
//                                                                                    matfree.rs:120-131
// The element matrices come from the same K1 kernels as assembly (elmat.cu) and are kept in a cell-major slab
// (the reference keeps the cell metrics and re-evaluates element(); the arithmetic per entry is the same).
// The converse incidence (dof -> (cell, position)) is built on the device by one stable radix sort, and stage 2
// is the CSR-stream segmented reduction shared with K3/K4 (stream.cuh): no atomics, fixed summation order.
#include <cub/cub.cuh>

#include "internal.hpp"
#include "stream.cuh"

struct fq_matfree {
  int dim = 0, kind = 0, grade = 0, tg = 0, rg = 0;
  int nt = 0, nr = 0;          // local test / trial DOFs per cell
  size_t nrows = 0, ncols = 0, ncells = 0;
  fq::DevBuf<double> slab;     // [ncells][nt*nr] element matrices
  fq::DevBuf<double> local;    // [ncells][nt]    stage-1 results
  fq::DevBuf<uint32_t> face_ptr, face_src;  // converse incidence: per DOF the (cell*nt + position) places, ascending cells
  fq::DevBuf<uint32_t> blocks;
  size_t nblocks = 0;
  const fq_mesh* mesh = nullptr;
};

namespace fq {

__global__ void mf_keys_kernel(const uint32_t* __restrict__ faces, size_t n, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < n; p += stride) keys[p] = faces[p], vals[p] = uint32_t(p);
}
__global__ void mf_ptr_kernel(const uint32_t* __restrict__ key, uint32_t n, uint32_t nrows, uint32_t* __restrict__ ptr) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i <= n; i += stride) {
    const uint32_t hi = (i == n) ? nrows : key[i];
    const uint32_t lo = (i == 0) ? 0u : key[i - 1] + 1;
    for (uint32_t r = lo; r <= hi && r <= nrows; ++r) ptr[r] = uint32_t(i);
  }
}
// stage 1: one thread per (cell, local row): row i of A_K times the gathered x, j ascending, mul then add (no FMA)
__global__ void mf_local_kernel(const double* __restrict__ slab, const uint32_t* __restrict__ cols, const double* __restrict__ x,
                                size_t ncells, int nt, int nr, double* __restrict__ local) {
  const size_t total = ncells * size_t(nt);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += stride) {
    const size_t c = p / size_t(nt);
    const int i = int(p - c * size_t(nt));
    const double* __restrict__ a = slab + (c * size_t(nt) + i) * size_t(nr);
    const uint32_t* __restrict__ cj = cols + c * size_t(nr);
    double acc = 0.0;
    for (int j = 0; j < nr; ++j) acc = __dadd_rn(acc, __dmul_rn(a[j], __ldg(x + cj[j])));
    local[p] = acc;
  }
}
__global__ void mf_diag_kernel(const double* __restrict__ slab, size_t ncells, int nt, double* __restrict__ local) {
  const size_t total = ncells * size_t(nt);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < total; p += stride) {
    const size_t c = p / size_t(nt);
    const int i = int(p - c * size_t(nt));
    local[p] = slab[(c * size_t(nt) + i) * size_t(nt) + i];
  }
}
struct MfGatherPolicy {
  static constexpr bool kHasValues = false;
  static constexpr bool kCustomSrc = false;
  static constexpr bool kGated = false;
  double* __restrict__ y;
  __device__ __forceinline__ double load(uint32_t, bool) const { return 0.0; }
  __device__ __forceinline__ void store(uint32_t dof, double sum, bool) const { y[dof] = sum; }
};

void matfree_build(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, fq_matfree* op, bool with_slab) {
  const int dim = mesh->dim;
  int tg, rg;
  kind_grades(kind, grade, tg, rg);
  FQ_REQUIRE(tg >= 0 && tg <= dim && rg >= 0 && rg <= dim, "matrix-free operator: both grades must lie in [0, dim]");
  FQ_REQUIRE(mesh->cell_faces[size_t(tg)].p && mesh->cell_faces[size_t(rg)].p, "mesh lacks the face tables of the required grades");
  FQ_REQUIRE(mesh->cell_offset == 0 && mesh->id_lo[size_t(tg)] == 0, "matrix-free operator needs a fully held mesh");
  op->dim = dim;
  op->kind = kind;
  op->grade = grade;
  op->tg = tg;
  op->rg = rg;
  op->nt = nlocal(dim, tg);
  op->nr = nlocal(dim, rg);
  op->nrows = mesh->nsimplices[size_t(tg)];
  op->ncols = mesh->nsimplices[size_t(rg)];
  op->ncells = mesh->ncells;
  op->mesh = mesh;
  const size_t nplaces = op->ncells * size_t(op->nt);
  FQ_REQUIRE(nplaces < (size_t(1) << 32), "more than 2^32 (cell, position) places: not supported");
  if (with_slab) op->slab.alloc(op->ncells * size_t(op->nt) * size_t(op->nr));
  op->local.alloc(nplaces ? nplaces : 1);
  // converse incidence: places sorted by DOF, ascending cell within a DOF (stable sort of cell-major places)
  const int block = 256;
  DevBuf<uint32_t> keys(nplaces), keys_alt(nplaces), vals(nplaces), vals_alt(nplaces);
  mf_keys_kernel<<<grid_for(nplaces, block, ctx->sm_count), block, 0, ctx->stream>>>(mesh->cell_faces[size_t(tg)].p, nplaces, keys.p,
                                                                                   vals.p);
  cub::DoubleBuffer<uint32_t> dk(keys.p, keys_alt.p), dv(vals.p, vals_alt.p);
  int end_bit = 1;
  while ((1ull << end_bit) < op->nrows) ++end_bit;
  size_t tmp_bytes = 0;
  FQ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, int64_t(nplaces), 0, end_bit, ctx->stream));
  DevBuf<uint8_t> tmp(tmp_bytes ? tmp_bytes : 1);
  FQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, int64_t(nplaces), 0, end_bit, ctx->stream));
  op->face_ptr.alloc(op->nrows + 1);
  op->face_src.alloc(nplaces ? nplaces : 1);
  mf_ptr_kernel<<<grid_for(nplaces + 1, block, ctx->sm_count), block, 0, ctx->stream>>>(dk.Current(), uint32_t(nplaces),
                                                                                      uint32_t(op->nrows), op->face_ptr.p);
  FQ_CUDA(cudaMemcpyAsync(op->face_src.p, dv.Current(), nplaces * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  fq_count_launch(ctx, 6);
  stream_build_blocks(ctx, op->face_ptr.p, op->nrows, nplaces, op->blocks, op->nblocks);
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
}

// LinearForm::assemble (formoniq/src/galerkin.rs:279-312): the converse incidence of one grade, no element matrices.
// assemble_vector's reduction "for cells in order: galvec[face] += elvec[position]" is stage 2 of the matrix-free apply.
void vector_plan_build(fq_ctx* ctx, const fq_mesh* mesh, int grade, fq_matfree* op) {
  FQ_REQUIRE(grade >= 0 && grade <= mesh->dim, "linear form: the grade must lie in [0, dim]");
  matfree_build(ctx, mesh, KIND_MASS, grade, op, /*with_slab=*/false);
}
void vector_plan_assemble(fq_ctx* ctx, const fq_matfree* op, const double* h_elvecs, double* y) {
  if (op->nrows == 0) return;
  const size_t nplaces = op->ncells * size_t(op->nt);
  FQ_CUDA(cudaMemcpyAsync(op->local.p, h_elvecs, nplaces * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  ScopedSpan span(ctx, "lf_gather");
  stream_reduce(ctx, op->blocks.p, op->nblocks, op->face_ptr.p, op->face_src.p, nullptr, op->local.p, MfGatherPolicy{y});
}

// SourceForm (operators.rs:607-635) assembled: element vectors by device quadrature (quadform.cu) straight into the
// plan's staging buffer, then the per-DOF gather.
void vector_plan_source(fq_ctx* ctx, const fq_matfree* op, int nnodes, const double* h_weights, const double* h_shapes,
                        const double* h_samples, double* y) {
  if (op->nrows == 0) return;
  source_element_vectors(ctx, op->mesh, op->tg, nnodes, h_weights, h_shapes, h_samples, op->local.p);
  ScopedSpan span(ctx, "lf_gather");
  stream_reduce(ctx, op->blocks.p, op->nblocks, op->face_ptr.p, op->face_src.p, nullptr, op->local.p, MfGatherPolicy{y});
}

void matfree_refresh(fq_ctx* ctx, fq_matfree* op) {  // element matrices from the mesh's current edge lengths
  double* outs[1] = {op->slab.p};
  DevBuf<int> err(1);
  FQ_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), ctx->stream));
  elmat_to_slabs(ctx, op->mesh, {BlockSpec{op->kind, op->grade}}, 0, op->ncells, true, outs, err.p);
  int h = 0;
  FQ_CUDA(cudaMemcpyAsync(&h, err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h) throw Error(FQ_ERR_DEGENERATE, "a cell metric is singular");
}

static void matfree_gather(fq_ctx* ctx, const fq_matfree* op, double* y) {
  ScopedSpan span(ctx, "mf_gather");
  stream_reduce(ctx, op->blocks.p, op->nblocks, op->face_ptr.p, op->face_src.p, nullptr, op->local.p, MfGatherPolicy{y});
}

void matfree_apply(fq_ctx* ctx, const fq_matfree* op, const double* x, double* y) {
  if (op->nrows == 0) return;
  const int block = 256;
  {
    ScopedSpan span(ctx, "mf_local");
    mf_local_kernel<<<grid_for(op->ncells * size_t(op->nt), block, ctx->sm_count), block, 0, ctx->stream>>>(
        op->slab.p, op->mesh->cell_faces[size_t(op->rg)].p, x, op->ncells, op->nt, op->nr, op->local.p);
    fq_count_launch(ctx);
    FQ_CUDA(cudaGetLastError());
  }
  matfree_gather(ctx, op, y);
}

size_t matfree_nrows(const fq_matfree* op) { return op->nrows; }
size_t matfree_ncols(const fq_matfree* op) { return op->ncols; }
void matfree_delete(fq_matfree* op) { delete op; }
fq_matfree* matfree_new() { return new fq_matfree; }

void matfree_diagonal(fq_ctx* ctx, const fq_matfree* op, double* d) {
  FQ_REQUIRE(op->nt == op->nr && op->nrows == op->ncols, "a diagonal needs a square operator");
  if (op->nrows == 0) return;
  const int block = 256;
  mf_diag_kernel<<<grid_for(op->ncells * size_t(op->nt), block, ctx->sm_count), block, 0, ctx->stream>>>(op->slab.p, op->ncells,
                                                                                                       op->nt, op->local.p);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
  matfree_gather(ctx, op, d);
}

}  // namespace fq
