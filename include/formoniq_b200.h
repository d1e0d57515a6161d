/* formoniq_b200.h — C ABI of the B200-native Galerkin-assembly path.
 *
 * This is the drop-in boundary for luiswirth/formoniq's assembly hot path.
 * The reference has no FFI of its own (pure Rust traits); each entry point
 * below names the reference interface it replaces (paths relative to the
 * reference checkout).  A Rust shim binds these with a plain `extern "C"`
 * block (see INTEGRATION.md and rust/formoniq-b200/src/lib.rs).
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; the message is
 *    available from fq_last_error() (thread-local).  Nothing unwinds across
 *    the ABI.
 *  - handles are opaque and owned by the library; host buffers are owned by
 *    the caller.  Index arrays crossing the ABI are 64-bit (`usize` on the
 *    reference's targets), values are IEEE-754 binary64.
 *  - there is NO CPU fallback: every compute entry point needs a CUDA device
 *    and fails with FQ_ERR_CUDA when none is usable.
 */
#ifndef FORMONIQ_B200_H
#define FORMONIQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FQ_OK 0
#define FQ_ERR_INVALID (-1)  /* bad argument / contract violation (the reference panics) */
#define FQ_ERR_CUDA (-2)     /* CUDA runtime failure or no device */
#define FQ_ERR_DEGENERATE (-3) /* a cell metric was singular (Metric::dual's expect) */
#define FQ_ERR_UNSUPPORTED (-4)

typedef struct fq_ctx fq_ctx;
typedef struct fq_mesh fq_mesh;
typedef struct fq_csr fq_csr;
typedef struct fq_vec fq_vec;
typedef struct fq_hodge fq_hodge;
typedef struct fq_matfree fq_matfree;

/* formoniq/src/operators.rs:169-191 (the four WhitneyPairing constructors) and
 * :27-40 (ScalarLumpedMass).  `grade` is always the grade of the inner
 * product, as in the reference (operators.rs:131-133). */
enum fq_kind { FQ_MASS = 0, FQ_DIF_TRIAL = 1, FQ_DIF_TEST = 2, FQ_DIF_BOTH = 3, FQ_LUMPED = 4 };

const char* fq_last_error(void);
/* number of visible CUDA devices (0 when there is none) */
int fq_device_count(void);

/* ---- context -------------------------------------------------------------
 * One context per device and host thread.  All work is enqueued on the
 * context's stream; fq_ctx_set_stream adopts a caller-owned cudaStream_t
 * (e.g. torch's current stream) so callers can bracket work with their own
 * events.  rank/nranks describe the owner-computes row partition used by the
 * multi-GPU entry points (1 process per GPU). */
int fq_ctx_create(int device, fq_ctx** out);
int fq_ctx_destroy(fq_ctx* ctx);
int fq_ctx_set_stream(fq_ctx* ctx, void* cuda_stream);
int fq_ctx_synchronize(fq_ctx* ctx);
/* kernels launched by this context so far (for bench.py's gpu_launches) */
int64_t fq_ctx_launch_count(const fq_ctx* ctx);
/* per-kernel device timing: when on, the hot kernels are bracketed with CUDA
 * events on the context's stream; the report is a JSON object
 * {"kernel": {"ms": total, "count": launches}, ...} and resets the spans. */
/* Device memory released by the library is kept in a process-wide cache for reuse (multi-GB cudaMalloc/cudaFree
 * calls synchronise the device and cost tens of milliseconds); this hands the cached blocks back to the driver. */
int fq_device_cache_trim(void);
int fq_ctx_set_timing(fq_ctx* ctx, int on);
int fq_ctx_timing_report(fq_ctx* ctx, char* buf, size_t buflen);

/* ---- mesh ----------------------------------------------------------------
 * What `Complex` + `MeshLengthsSq` expose to assembly:
 *   cell_faces[j]  = FaceIncidence::faces_flat of grade j, cell-major with
 *                    stride C(dim+1, j+1)  (simplicial/src/topology/incidence.rs:42-53)
 *                    entries may be NULL for grades the caller will not use;
 *                    grade 1 is required (edge lengths are gathered through it)
 *   edge_lengths_sq = MeshLengthsSq (regge/src/lengths/mesh.rs:34-36), one
 *                    signed squared length per edge of the 1-skeleton. */
int fq_mesh_create(fq_ctx* ctx, int dim, size_t ncells, const size_t* nsimplices /*[dim+1]*/,
                   const uint64_t* const* cell_faces /*[dim+1]*/, const double* edge_lengths_sq, fq_mesh** out);
/* One rank's part of an uploaded mesh under owner-computes (the partition SURVEY 8e describes, for any Complex in the
 * reference's colex skeleton numbering — simplices sorted by their top vertex, simplicial/src/topology/skeleton.rs:50-86 —
 * not only the generated Kuhn grids): the rank owns a contiguous vertex range, hence the contiguous id range
 * [own_lo[j], own_hi[j]) of every grade j (the simplices whose top vertex it owns), and HOLDS the ncells_held cells that
 * touch one of its vertices, in ascending global order.  cell_faces[j] are the rows of the global tables for those cells
 * with GLOBAL ids (cell_faces[dim] = their global cell ids), nsimplices the GLOBAL counts, edge_lengths_sq the global
 * array.  Assembling with the row range fq_mesh_owned_range reports gives the rank's row block, bit-identical to the
 * same rows of the one-GPU matrix (every cell contributing to an owned row is held, in the same order).  The host-side
 * split (vertex ranges balanced by incident cells, held cells, id ranges) is formoniq_b200.dist.partition_mesh. */
int fq_mesh_create_part(fq_ctx* ctx, int dim, size_t ncells_held, const size_t* nsimplices /*[dim+1]*/,
                        const uint64_t* const* cell_faces /*[dim+1]*/, const double* edge_lengths_sq,
                        const size_t* own_lo /*[dim+1]*/, const size_t* own_hi /*[dim+1]*/, fq_mesh** out);
/* Kuhn triangulation of a box grid generated on the device with the
 * reference's colex skeleton numbering (simplicial/src/mesher/grid.rs:77-103,
 * regge/src/mesher/cartesian.rs:158-169,192-208, regge/src/coord/mesh.rs:208-216).
 *   shape[a]  cells along axis a;  vmin/vmax the box;  ambient_diag the diagonal
 *   ambient form (+1 Euclid, -1 for a time axis);  jitter displaces every vertex
 *   by jitter*h_a*pseudo_random(seed=a, index=v) (formoniq/src/linalg/eigen.rs:259-268).
 * slab_begin/slab_end select the box layers [slab_begin, slab_end) along the
 * LAST axis that this context owns (owner-computes partition); pass
 * 0, shape[dim-1] for the whole grid.  Simplex ids stay global.  The mesh also
 * holds the halo layer slab_end (when it exists): its cells touch owned rows.
 * fq_mesh_owned_range gives the rows (simplices of a grade) this slab owns,
 * fq_mesh_held_range the ids its cells reference (owned + halos, contiguous). */
int fq_mesh_create_kuhn(fq_ctx* ctx, int dim, const size_t* shape, const double* vmin, const double* vmax,
                        const double* ambient_diag, double jitter, size_t slab_begin, size_t slab_end,
                        fq_mesh** out);
int fq_mesh_destroy(fq_mesh* mesh);
int fq_mesh_dim(const fq_mesh* mesh);
size_t fq_mesh_ncells(const fq_mesh* mesh);
size_t fq_mesh_nsimplices(const fq_mesh* mesh, int grade);
size_t fq_mesh_nowned_cells(const fq_mesh* mesh);
int fq_mesh_owned_range(const fq_mesh* mesh, int grade, size_t* lo, size_t* hi);
int fq_mesh_held_range(const fq_mesh* mesh, int grade, size_t* lo, size_t* hi);
/* replace the geometry (MeshLengthsSq) of an existing mesh */
int fq_mesh_set_lengths(fq_ctx* ctx, fq_mesh* mesh, const double* edge_lengths_sq);
/* copy device-side tables back (parity hooks for the generator) */
int fq_mesh_download_cell_faces(fq_ctx* ctx, const fq_mesh* mesh, int grade, uint64_t* out);
int fq_mesh_download_lengths(fq_ctx* ctx, const fq_mesh* mesh, double* out);
/* Host-only evaluation of the closed-form Kuhn numbering (no device needed):
 * fills cell_faces of one grade for the whole grid.  Useful to build a
 * reference-side `Complex` without its hash-based `from_cells`. */
int fq_kuhn_cell_faces_host(int dim, const size_t* shape, int grade, uint64_t* out);
int fq_kuhn_counts(int dim, const size_t* shape, size_t* nsimplices /*[dim+1]*/);
/* Host-only: id ranges of one grade for the slab [slab_begin, slab_end):
 * out4 = {held_lo, own_lo, own_hi, held_hi} (what fq_mesh_held_range /
 * fq_mesh_owned_range report for the device mesh). */
int fq_kuhn_slab_ranges(int dim, const size_t* shape, size_t slab_begin, size_t slab_end, int grade, size_t* out4);

/* ---- element matrices ------------------------------------------------------
 * BilinearForm::element for a batch of cells (formoniq/src/galerkin.rs:50,
 * operators.rs:84-94,201-211).  out is host memory, row-major
 * [cell_end-cell_begin][rows][cols].  use_generated=0 forces the generic tape
 * interpreter (any dim/grade); 1 uses the straight-line kernel when one was
 * generated for (dim, kind, grade) and the interpreter otherwise. */
int fq_elmat_shape(int dim, int kind, int grade, int* rows, int* cols);
int fq_elmat_batch(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, size_t cell_begin, size_t cell_end,
                   int use_generated, double* out);

/* ---- assembly --------------------------------------------------------------
 * BilinearForm::assemble / assemble_matrix (formoniq/src/galerkin.rs:52-57,
 * 138-188) split into a symbolic phase (pattern + cell-slot -> nnz map, once
 * per mesh and block) and a numeric phase (values; the hot path).
 * Rows [row_begin,row_end) of the global matrix are assembled (owner-computes;
 * pass 0, SIZE_MAX for all rows).
 * drop_exact_zeros=1 reproduces the reference's `val != 0.0` triplet filter
 * (galerkin.rs:173): an entry exists iff some contribution is non-zero. */
int fq_assemble_symbolic(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, size_t row_begin, size_t row_end,
                         fq_csr** out);
int fq_assemble_numeric(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* csr, int drop_exact_zeros);
/* one-shot convenience: symbolic + numeric */
int fq_assemble(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, int drop_exact_zeros, fq_csr** out);

/* HodgeBlocks::compute (formoniq/src/hodge.rs:62-72): the four blocks of a mixed
 * problem posed at `grade` — mass(grade-1), mass(grade), dif_test(grade),
 * dif_both(grade+1) — with ONE fused element kernel per numeric pass (the cell
 * metric, its inverse and the shared masses are evaluated once per cell instead
 * of once per block).  sigma rows = simplices of grade-1, u rows = grade.
 * fq_hodge_block borrows block `which` (0 mass_sigma, 1 mass_u, 2 dif_test,
 * 3 dif_both); it stays owned by the fq_hodge. */
int fq_hodge_symbolic(fq_ctx* ctx, const fq_mesh* mesh, int grade, size_t sigma_row_begin, size_t sigma_row_end,
                      size_t u_row_begin, size_t u_row_end, fq_hodge** out);
int fq_hodge_numeric(fq_ctx* ctx, const fq_mesh* mesh, fq_hodge* blocks, int drop_exact_zeros);
fq_csr* fq_hodge_block(fq_hodge* blocks, int which);
/* C = A + B on the union pattern, explicit zeros kept (nalgebra-sparse `&a + &b`): HilbertComplex::hdif_gram =
 * mass(k) + dif_both(k + 1), crates/formoniq/src/whitney_complex.rs:180-183, whose sparse factorizations are the blocks
 * of the AFW preconditioner (problems/elliptic.rs:29-47) */
int fq_csr_add(fq_ctx* ctx, const fq_csr* a, const fq_csr* b, fq_csr** out);
/* y_i = sum_j |a_ij| over the held rows; its maximum is the inf-norm the eigen solver scales residuals with
 * (linalg/eigen.rs:359-368) */
int fq_csr_row_abs_sums(fq_ctx* ctx, const fq_csr* a, fq_vec* y);
/* d_i = 1 / a_ii over the held rows of a square matrix: the Jacobi preconditioner of iterative/src/precond.rs:87-121 as a
 * vector, for callers that drive a Krylov method through fq_cg_op / fq_minres_op (a row-partitioned operator applies it
 * with fq_vec_mul).  FQ_ERR_DEGENERATE on a missing or zero diagonal entry, as the reference asserts (:101-104). */
int fq_csr_inv_diagonal(fq_ctx* ctx, fq_csr* a, fq_vec* d);
/* z = d .* r (component-wise product; z may alias r) */
int fq_vec_mul(fq_ctx* ctx, fq_vec* z, const fq_vec* d, const fq_vec* r);

/* HodgeBlocks::mixed_hodge_laplacian (formoniq/src/hodge.rs:93-99): [[M_{k-1}, -dif_test], [dif_test^T, dif_both]]
 * stitched on the device (stable transpose + row-wise concatenation instead of CooMatrixExt::block,
 * simplicial/src/linalg.rs:110-167, and a second COO->CSR); bit-identical entries.  The blocks must be fully held
 * (single-GPU row range).  fq_csr_transpose is the transpose used for the lower-left block. */
int fq_hodge_mixed_laplacian(fq_ctx* ctx, const fq_hodge* blocks, fq_csr** out);
/* the same saddle point with the sigma block-row negated, [[-M_{k-1}, dif_test], [dif_test^T, dif_both]]: the symmetric
 * form assemble_mixed_kkt hands to MINRES (problems/elliptic.rs:101-113) */
int fq_hodge_mixed_kkt_symmetric(fq_ctx* ctx, const fq_hodge* blocks, fq_csr** out);
int fq_csr_transpose(fq_ctx* ctx, const fq_csr* a, fq_csr** out);
/* RelativeWhitneyComplex::assemble (formoniq/src/whitney_complex.rs:620-624): E_test^T A E_trial for the 0/1 inclusions of
 * the interior DOFs = the sub-matrix A[rows_keep, cols_keep] (ascending index lists, e.g. interior_simps of the test and
 * trial grade, whitney_complex.rs:562-575), done as an index compaction on the device instead of two sparse products. */
int fq_csr_restrict(fq_ctx* ctx, const fq_csr* a, const size_t* rows_keep, size_t nrows_keep, const size_t* cols_keep,
                    size_t ncols_keep, fq_csr** out);
int fq_hodge_destroy(fq_hodge* blocks);

/* ---- CSR matrices ----------------------------------------------------------
 * The data contract of nalgebra_sparse::CsrMatrix<f64> handed to faer by
 * formoniq/src/linalg/faer.rs:16-24: row_offsets[nrows+1], col_indices strictly
 * ascending within a row, values. */
int fq_csr_shape(const fq_csr* csr, size_t* nrows, size_t* ncols, size_t* nnz);
int fq_csr_row_range(const fq_csr* csr, size_t* row_begin, size_t* row_end);
int fq_csr_download(fq_ctx* ctx, const fq_csr* csr, size_t* row_offsets, size_t* col_indices, double* values);
int fq_csr_upload(fq_ctx* ctx, size_t nrows, size_t ncols, const size_t* row_offsets, const size_t* col_indices,
                  const double* values, fq_csr** out);
/* The same download enqueued on the context's copy stream after everything submitted so far: it overlaps whatever the
 * caller enqueues next (e.g. the assembly of the next block).  The host buffers (pinned for PCIe-rate copies) and the
 * matrix must stay alive until fq_ctx_wait_downloads returns.  The index arrays cross PCIe as the device's u32 and are
 * widened to usize IN the caller's buffers by a pool of host threads (FQ_HOST_WIDEN_THREADS, default min(16, cores - 1);
 * 0 = widen on the device and move u64): the contents of row_offsets / col_indices are undefined until the wait. */
int fq_csr_download_async(fq_ctx* ctx, const fq_csr* csr, size_t* row_offsets, size_t* col_indices, double* values);
int fq_ctx_wait_downloads(fq_ctx* ctx);
int fq_csr_destroy(fq_csr* csr);
/* algorithmic HBM bytes of the last numeric assembly / of one SpMV */
int64_t fq_csr_assembly_bytes(const fq_csr* csr);
/* the part of fq_csr_assembly_bytes that blocks assembled together in one fused launch (fq_hodge_numeric) read ONCE:
   edge lengths + cell -> edge ids (8 E + 4 C(n+1,2) C); the rest (4 B per element entry of the map, 8 B per non-zero) is
   per block */
int64_t fq_csr_assembly_shared_bytes(const fq_csr* csr);
/* device milliseconds the last tile-plan build of this matrix took (symbolic pattern + record streams; 0: none) */
double fq_csr_plan_build_ms(const fq_csr* csr);
/* cell visits of the tile plan (a cell is evaluated once per tile touching it): element tapes per fused launch */
size_t fq_csr_plan_cell_visits(const fq_csr* csr);
int64_t fq_csr_spmv_bytes(const fq_csr* csr);

/* ---- vectors: iterative::InnerProductSpace (iterative/src/lib.rs:84-141) ---- */
int fq_vec_create(fq_ctx* ctx, size_t n, fq_vec** out); /* zeros_like */
/* non-owning view of caller-owned device memory (e.g. a torch tensor) */
int fq_vec_wrap(fq_ctx* ctx, void* device_ptr, size_t n, fq_vec** out);
int fq_vec_destroy(fq_vec* v);
size_t fq_vec_len(const fq_vec* v);
int fq_vec_upload(fq_ctx* ctx, fq_vec* v, const double* host);
int fq_vec_download(fq_ctx* ctx, const fq_vec* v, double* host);
int fq_vec_copy(fq_ctx* ctx, fq_vec* dst, const fq_vec* src);                  /* clone */
int fq_vec_dot(fq_ctx* ctx, const fq_vec* x, const fq_vec* y, double* out);   /* dot   */
int fq_vec_scale(fq_ctx* ctx, fq_vec* x, double alpha);                        /* scale */
int fq_vec_axpy(fq_ctx* ctx, fq_vec* y, double alpha, const fq_vec* x);       /* add_scaled */
/* raw device pointer (for zero-copy interop with torch / NCCL plumbing) */
void* fq_vec_device_ptr(fq_vec* v);

/* ---- SpMV: iterative::LinearOperator::apply (iterative/src/operator.rs:5-14) ---- */
int fq_spmv(fq_ctx* ctx, const fq_csr* a, const fq_vec* x, fq_vec* y);
/* y = A x where x holds only the window [x_lo, x_lo + len(x)) of the global
 * column space (owned segment + halos of a row-partitioned operator). */
int fq_spmv_window(fq_ctx* ctx, const fq_csr* a, const fq_vec* x, size_t x_lo, fq_vec* y);

/* ---- matrix-free operator: formoniq::matfree::ElementOperator (formoniq/src/matfree.rs:60-179) --------------------
 * y = sum_K P_K^T A_K P_K x by the reference's two-stage gather (per cell A_K * gathered x, then per DOF the sum over
 * FaceIncidence::face_cells in cell order): no global matrix, no atomics.  create walks the mesh once (converse incidence
 * + element matrices); refresh re-evaluates the element matrices after fq_mesh_set_lengths; diagonal is matfree.rs:155-179
 * (the Jacobi preconditioner of a matrix-free operator).  The mesh must outlive the operator. */
int fq_matfree_create(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, fq_matfree** out);
int fq_matfree_refresh(fq_ctx* ctx, fq_matfree* op);
int fq_matfree_destroy(fq_matfree* op);
int fq_matfree_shape(const fq_matfree* op, size_t* nrows, size_t* ncols);
int fq_matfree_apply(fq_ctx* ctx, const fq_matfree* op, const fq_vec* x, fq_vec* y);
int fq_matfree_diagonal(fq_ctx* ctx, const fq_matfree* op, fq_vec* d);

/* ---- LinearForm::assemble: formoniq::galerkin::assemble_vector (formoniq/src/galerkin.rs:279-312) ---------------
 * The Galerkin (load) vector of a linear form of the given grade: ell_sigma = sum over the cells K containing sigma, in
 * cell order, of elvec_K[position of sigma in K] - the reference's "for cell in cells: galvec[face] += elvec[ilocal]"
 * (its `!= 0.0` filter only skips additions of zero), as a per-DOF segmented sum over the converse incidence: no
 * atomics, the reference's summation order, bit-identical to the CPU.  `element_vectors` is a HOST array, cell-major
 * [ncells][C(dim+1, grade+1)] in the local face order of SimplexRef::faces(grade): LinearForm::element evaluates a user
 * closure (a Section sampled at quadrature nodes, operators.rs:607-635) and stays host code, as in the reference.
 * A plan (the converse incidence, one radix sort) serves any number of right-hand sides on the same mesh; the handle is
 * an fq_matfree without element matrices. */
int fq_linear_form_create(fq_ctx* ctx, const fq_mesh* mesh, int grade, fq_matfree** out);
int fq_linear_form_assemble(fq_ctx* ctx, const fq_matfree* plan, const double* element_vectors, fq_vec* out);
int fq_linear_form_destroy(fq_matfree* plan);
/* SourceForm (formoniq/src/operators.rs:607-635) assembled in one call: the element vectors
 *   elvec_K[sigma] = vol_K * sum_q w_q <f(x_q), W_sigma(x_q)>_{Lambda^k g_K^{-1}}      (operators.rs:247-261, tensor.rs:140-157)
 * are evaluated on the device (one thread per cell: metric by polarisation, inverse, volume, k x k minors of g^{-1}) and
 * reduced per DOF by the plan.  Host inputs: `weights[nnodes]` (normalised, SimplexQuadRule::weights), `shapes`
 * [nnodes][C(n+1,k+1)][C(n,k)] = LsfSamples::whitney (derham/src/interpolate/samples.rs:31-40), `samples`
 * [ncells][nnodes][C(n,k)] = the source Section evaluated at the nodes of every cell in the cell's reference frame
 * (the user closure of the reference; the only per-cell host work).  Values agree with the reference to rounding
 * (1e-12 relative), not bitwise: the reference sums through nalgebra's generic tensor contraction. */
int fq_source_form_assemble(fq_ctx* ctx, const fq_matfree* plan, int nnodes, const double* weights, const double* shapes,
                            const double* samples, fq_vec* out);

/* ---- WeightedHodgeMass (formoniq/src/operators.rs:432-486) ----------------------------------------------------------
 * [int_K alpha <W_sigma, W_tau> vol], the varying-coefficient mass, as a numeric pass on a matrix created by
 * fq_assemble_symbolic(FQ_MASS, grade): the element matrices vol_K * sum_q w_q alpha(x_q) W_i(q)^T (Lambda^k g^-1) W_j(q)
 * (CellQuadrature::integrate_pair, operators.rs:266-290) are evaluated on the device, one thread per cell, into the
 * element slab and scattered by the K3 reduction under the same pattern semantics as every other form
 * (`drop_exact_zeros` = galerkin.rs:173).  Host inputs as for fq_source_form_assemble; `coefficient` [ncells][nnodes] is the
 * grade-0 section alpha sampled at the nodes of every cell (the user closure).  A later fq_assemble_numeric on the same
 * handle re-assembles the plain (closed-form) mass. */
int fq_weighted_mass_numeric(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* csr, int nnodes, const double* weights,
                             const double* shapes, const double* coefficient, int drop_exact_zeros);

/* ---- SpMV fused with the halo exchange (one process per GPU, NVLink 5 / NVSwitch peer memory) -----------------
 * The reference is single-process; under the owner-computes row partition the only exchange step of the path is
 * the x halo of `LinearOperator::apply`.  Instead of exchanging halos and then multiplying, the gather of the SpMV
 * loads the columns owned by the neighbouring ranks directly from their windows over NVLink:
 *   fq_vec_ipc_export / fq_vec_ipc_import  map a peer rank's vector (64-byte CUDA IPC handle, shipped by the caller,
 *                                          e.g. with torch.distributed.all_gather_object)
 *   fq_spmv_peer   y = A x; x_window covers [held_lo, ..) of this rank, columns < own_lo come from x_lower (the
 *                  window of rank-1, starting at lower_held_lo), columns >= own_hi from x_upper; NULL = no neighbour.
 *                  The first CTAs of the kernel pull the halo segments into x_window's halo slots (coalesced P2P
 *                  loads) while the others multiply the row blocks that need owned columns only; the boundary row
 *                  blocks wait on a device counter.  (FQ_PEER_DIRECT=1: the gather loads remote columns itself.)
 *   fq_flag_signal / fq_flag_wait  stream-ordered epoch flags in peer-mapped memory: a rank signals after its last
 *                  write of x and waits for its neighbours' epochs before the fused SpMV reads them (and the
 *                  reverse before x is overwritten); fq_flag_check reports a wait that timed out (~3 s). */
int fq_vec_ipc_export(fq_ctx* ctx, const fq_vec* v, unsigned char* handle64);
int fq_vec_ipc_import(fq_ctx* ctx, const unsigned char* handle64, size_t n, fq_vec** out);
int fq_spmv_peer(fq_ctx* ctx, const fq_csr* a, const fq_vec* x_window, size_t held_lo, size_t own_lo, size_t own_hi,
                 const fq_vec* x_lower, size_t lower_held_lo, const fq_vec* x_upper, size_t upper_held_lo, fq_vec* y);
/* The same kernel also doing the synchronisation: its first CTAs wait until ready_lower / ready_upper (the
 * neighbours' published epochs, peer-mapped; NULL = none) reach `epoch`, pull the halo segments into x_window over
 * NVLink while the other CTAs already multiply the interior row blocks, and write `epoch` to `consumed` (this rank's
 * flag, polled by the neighbours before they overwrite x; NULL = none) as soon as the last halo entry has arrived. */
int fq_spmv_peer_epoch(fq_ctx* ctx, const fq_csr* a, const fq_vec* x_window, size_t held_lo, size_t own_lo, size_t own_hi,
                       const fq_vec* x_lower, size_t lower_held_lo, const fq_vec* ready_lower, const fq_vec* x_upper,
                       size_t upper_held_lo, const fq_vec* ready_upper, double epoch, fq_vec* consumed, fq_vec* y);
int fq_flag_signal(fq_ctx* ctx, fq_vec* flag, double value);
int fq_flag_wait(fq_ctx* ctx, const fq_vec* flag, double value);
int fq_flag_check(fq_ctx* ctx);

/* ---- Krylov drivers on device vectors (iterative/src/krylov.rs:48-95, 113-211).
 * precond: 0 identity (iterative/src/precond.rs:16-41), 1 Jacobi (:113-121).
 * report: iters, residual, converged (iterative/src/lib.rs:212-219). */
int fq_cg(fq_ctx* ctx, const fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters, fq_vec* x,
          size_t* iters, double* residual, int* converged);
int fq_minres(fq_ctx* ctx, const fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters, fq_vec* x,
              size_t* iters, double* residual, int* converged);

/* ---- Krylov over user operators: iterative::{cg, minres} are generic over LinearOperator / ApproxInverse /
 * InnerProductSpace (iterative/src/krylov.rs:48-95, 113-211; lib.rs:84-157).  The operator and the preconditioner are
 * callbacks on DEVICE pointers (y = A x, z = M^-1 r; precond NULL = Identity, precond.rs:16-41); `reduce` completes a
 * rank-local inner product (the all-reduce of a distributed space; NULL = single process).  All vector updates of
 * the iteration stay inside the library.  Callbacks return 0 on success. */
typedef int (*fq_apply_fn)(void* user, const double* x_device, double* y_device);
typedef int (*fq_reduce_fn)(void* user, double local, double* global);
int fq_cg_op(fq_ctx* ctx, size_t n, fq_apply_fn apply, fq_apply_fn precond, fq_reduce_fn reduce, void* user, const fq_vec* b,
             double rtol, size_t max_iters, fq_vec* x, size_t* iters, double* residual, int* converged);
int fq_minres_op(fq_ctx* ctx, size_t n, fq_apply_fn apply, fq_apply_fn precond, fq_reduce_fn reduce, void* user,
                 const fq_vec* b, double rtol, size_t max_iters, fq_vec* x, size_t* iters, double* residual, int* converged);
/* MINRES on `a` with a block-diagonal preconditioner of inner solves: segment [offsets[i], offsets[i+1]) of the
 * vector is preconditioned by blocks[i]^-1 (Jacobi-CG to inner_rtol; NULL = identity).  With blocks =
 * {hdif_gram(k-1), hdif_gram(k)} this is the AFW mixed_block_preconditioner of problems/elliptic.rs:29-47, whose blocks
 * the reference factorises with a sparse Cholesky. */
int fq_minres_blockdiag(fq_ctx* ctx, const fq_csr* a, int nblocks, const fq_csr* const* blocks, const size_t* offsets,
                        double inner_rtol, size_t inner_max_iters, const fq_vec* b, double rtol, size_t max_iters, fq_vec* x,
                        size_t* iters, double* residual, int* converged, size_t* inner_iters);
/* non-owning view of v[offset, offset + n) (the sigma / u segments of a mixed vector) */
int fq_vec_view(fq_ctx* ctx, const fq_vec* v, size_t offset, size_t n, fq_vec** out);

#ifdef __cplusplus
}
#endif
#endif /* FORMONIQ_B200_H */
