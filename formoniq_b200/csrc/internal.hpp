// internal.hpp — functions shared between the translation units of the library.
#pragma once
#include <functional>
#include "common.cuh"

namespace fq {

// ---- elmat.cu
// Element matrices of cells [c0, c1) (local cell indices) of one block, or of
// the fused blocks of hodge_blocks(k), into per-block cell-major device slabs:
// d_outs[b][(c - c0) * T_b + e].
void elmat_to_slabs(fq_ctx* ctx, const fq_mesh* mesh, const std::vector<BlockSpec>& blocks, size_t c0, size_t c1,
                    bool use_generated, double* const* d_outs, int* d_err);
int elmat_nouts(int dim, const std::vector<BlockSpec>& blocks);
bool elmat_has_generated(int dim, const std::vector<BlockSpec>& blocks);

// ---- kuhn.cu
void kuhn_build_mesh(fq_ctx* ctx, int dim, const size_t* shape, const double* vmin, const double* vmax,
                     const double* ambient_diag, double jitter, size_t slab_begin, size_t slab_end, fq_mesh* mesh);

// ---- assemble.cu
void assemble_symbolic(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, size_t row_begin, size_t row_end,
                       fq_csr* out);
void assemble_numeric(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* csr, bool drop_exact_zeros);
void assemble_numeric_custom(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* csr, bool drop_exact_zeros,
                             const std::function<void(double*)>& fill);
void weighted_mass_to_slab(fq_ctx* ctx, const fq_mesh* mesh, int grade, int nnodes, const double* h_weights,
                           const double* h_shapes, const double* h_coeff, double* d_slab);
// fused numeric phase of several blocks sharing one element kernel launch (HodgeBlocks)
void assemble_numeric_multi(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* const* csrs, int nblocks, bool drop_exact_zeros);

// ---- tile.cu (tile-fused numeric assembly)
void tile_cluster_kuhn(fq_ctx* ctx, fq_mesh* mesh, int dim, const size_t* shape, size_t slab_begin, size_t slab_end_held);
void tile_cluster_generic(fq_ctx* ctx, fq_mesh* mesh);  // lazy: run when a tile plan is first built
// builds the plan AND the structural patterns (s_row_ptr / s_col_idx / s_nnz / ncontrib) of the blocks
std::shared_ptr<TilePlan> tile_plan_build(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* const* csrs, int nblocks);
bool tile_plan_matches(const TilePlan& plan, const fq_mesh* mesh, fq_csr* const* csrs, int nblocks);
bool tile_assemble(fq_ctx* ctx, const fq_mesh* mesh, TilePlan& plan, double* const* values, uint8_t* const* keep);
void tile_retarget(fq_ctx* ctx, TilePlan& plan);
bool& tile_plan_compact(TilePlan& plan);
bool& tile_plan_drop(TilePlan& plan);
double tile_plan_build_ms(const TilePlan& plan);
size_t tile_plan_cell_visits(const TilePlan& plan);
int64_t tile_plan_bytes(const TilePlan& plan);
int64_t tile_plan_stream_bytes(const TilePlan& plan);
int64_t tile_plan_cv_bytes(const TilePlan& plan);

// ---- spmv.cu
void spmv_prepare(fq_ctx* ctx, fq_csr* a);
void spmv_apply(fq_ctx* ctx, const fq_csr* a, const double* x, double* y);
// Epoch flags the fused halo SpMV can handle inside the kernel (all optional): wait until the neighbours' `ready`
// flags reach `epoch` before the first P2P load, write `epoch` to `consumed` once the halo has been pulled.
struct PeerSync {
  const double* ready_lower = nullptr;
  const double* ready_upper = nullptr;
  double epoch = 0.0;
  double* consumed = nullptr;
  int* timeout = nullptr;
};
void spmv_apply_peer(fq_ctx* ctx, fq_csr* a, double* own, const double* lower, const double* upper, size_t held_lo,
                     size_t own_lo, size_t own_hi, size_t held_hi, double* y, const PeerSync& sync = PeerSync());
void flag_signal(fq_ctx* ctx, double* flag, double value);
void flag_wait(fq_ctx* ctx, const double* flag, double value, int* d_timeout);
void csr_build_inv_diag(fq_ctx* ctx, fq_csr* a);

// ---- blockop.cu
void csr_transpose(fq_ctx* ctx, const fq_csr* a, fq_csr* out);
void csr_add(fq_ctx* ctx, const fq_csr* a, const fq_csr* b, fq_csr* out);
void csr_row_abs_sums(fq_ctx* ctx, const fq_csr* a, double* y);
void csr_block2x2(fq_ctx* ctx, const fq_csr* a00, const fq_csr* a01, double s01, const fq_csr* a10, const fq_csr* a11,
                  fq_csr* out, double s00 = 1.0);

void csr_restrict(fq_ctx* ctx, const fq_csr* a, const uint32_t* rows_keep, size_t nr, const uint32_t* cols_keep, size_t nc,
                  fq_csr* out);

// ---- matfree.cu
}  // namespace fq
struct fq_matfree;
namespace fq {
void matfree_build(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, ::fq_matfree* op, bool with_slab = true);
void vector_plan_build(fq_ctx* ctx, const fq_mesh* mesh, int grade, ::fq_matfree* op);
void vector_plan_assemble(fq_ctx* ctx, const ::fq_matfree* op, const double* h_elvecs, double* y);
void vector_plan_source(fq_ctx* ctx, const ::fq_matfree* op, int nnodes, const double* h_weights, const double* h_shapes,
                        const double* h_samples, double* y);
void source_element_vectors(fq_ctx* ctx, const fq_mesh* mesh, int grade, int nnodes, const double* h_weights,
                            const double* h_shapes, const double* h_samples, double* d_elvecs);
void matfree_refresh(fq_ctx* ctx, ::fq_matfree* op);
void matfree_apply(fq_ctx* ctx, const ::fq_matfree* op, const double* x, double* y);
void matfree_diagonal(fq_ctx* ctx, const ::fq_matfree* op, double* d);
size_t matfree_nrows(const ::fq_matfree* op);
size_t matfree_ncols(const ::fq_matfree* op);
void matfree_delete(::fq_matfree* op);
::fq_matfree* matfree_new();

// ---- blas1.cu
double vec_dot(fq_ctx* ctx, const double* x, const double* y, size_t n);
void vec_dot_device(fq_ctx* ctx, const double* x, const double* y, size_t n, double* d_partials, double* d_out);
size_t vec_dot_scratch_doubles();
void cg_fused_update(fq_ctx* ctx, double* x, double* r, const double* p, const double* ap, double* z, const double* d,
                     const double* rz, const double* pap, const int* done, size_t n, double* part_rz, double* part_rr,
                     double* st, unsigned long long* iters, int* flags, double rtol, size_t max_iters);
void vec_scale(fq_ctx* ctx, double* x, double alpha, size_t n);
void vec_axpy(fq_ctx* ctx, double* y, double alpha, const double* x, size_t n);
void vec_mul_pointwise(fq_ctx* ctx, double* z, const double* d, const double* r, size_t n);

// ---- krylov.cu
struct KrylovReport {
  size_t iters = 0;
  double residual = 0.0;
  bool converged = false;
};
// what cg / minres need from their operands (iterative/src/lib.rs:84-157: LinearOperator, ApproxInverse and the inner
// product of the space): y = A x, z = M^-1 r, <u, v> — the last one global when the vectors are distributed
struct KrylovOps {
  std::function<void(const double* x, double* y)> apply;
  std::function<void(const double* r, double* z)> precond;
  std::function<double(const double* u, const double* v)> dot;
};
KrylovReport cg_core(fq_ctx* ctx, size_t n, const KrylovOps& ops, const double* b, double rtol, size_t max_iters, double* x);
KrylovReport minres_core(fq_ctx* ctx, size_t n, const KrylovOps& ops, const double* b, double rtol, size_t max_iters, double* x);
KrylovReport krylov_cg(fq_ctx* ctx, fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters, fq_vec* x);
KrylovReport krylov_minres(fq_ctx* ctx, fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters,
                           fq_vec* x);
KrylovReport krylov_minres_blockdiag(fq_ctx* ctx, fq_csr* a, int nblocks, fq_csr* const* blocks, const size_t* offsets,
                                     double inner_rtol, size_t inner_max_iters, const fq_vec* b, double rtol, size_t max_iters,
                                     fq_vec* x, size_t* inner_iters_total);

inline int grid_for(size_t n, int block, int sm_count, int ctas_per_sm = 8) {
  const size_t want = (n + size_t(block) - 1) / size_t(block);
  const size_t cap = size_t(sm_count) * size_t(ctas_per_sm);
  return int(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace fq
