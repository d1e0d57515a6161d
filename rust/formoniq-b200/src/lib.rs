//! formoniq-b200: the CUDA assembly path behind formoniq's own traits.
//!
//! Three seams, all monomorphised generics in the reference:
//!  1. `BilinearForm::assemble`      (formoniq/src/galerkin.rs:52-57)   -> `GpuPairing`
//!  2. `HilbertComplex::assemble`    (formoniq/src/whitney_complex.rs:66) -> `GpuWhitneyComplex`
//!  3. `LinearOperator` + `InnerProductSpace` (iterative/src/lib.rs:84-157) -> `DeviceCsr`, `DeviceVector`
//!
//! Source form only: this image has no Rust toolchain, so the crate is
//! exercised through the same C ABI from C++/Python tests instead.
#![allow(non_camel_case_types)]
use std::{ffi::{c_char, c_double, c_int, c_void, CStr}, ptr};

use formoniq::{galerkin::{BilinearForm, GalerkinMatrix}, operators::WhitneyPairing};
use iterative::{InnerProductSpace, LinearOperator};
use nalgebra_sparse::CsrMatrix;
use regge::lengths::mesh::MeshLengthsSq;
use simplicial::topology::{complex::Complex, incidence::FaceIncidence};

#[repr(C)] pub struct fq_ctx { _p: [u8; 0] }
#[repr(C)] pub struct fq_mesh { _p: [u8; 0] }
#[repr(C)] pub struct fq_csr { _p: [u8; 0] }
#[repr(C)] pub struct fq_vec { _p: [u8; 0] }
#[repr(C)] pub struct fq_hodge { _p: [u8; 0] }
#[repr(C)] pub struct fq_matfree { _p: [u8; 0] }

// include/formoniq_b200.h
unsafe extern "C" {
  fn fq_last_error() -> *const c_char;
  fn fq_ctx_create(device: c_int, out: *mut *mut fq_ctx) -> c_int;
  fn fq_ctx_destroy(ctx: *mut fq_ctx) -> c_int;
  fn fq_mesh_create(ctx: *mut fq_ctx, dim: c_int, ncells: usize, nsimplices: *const usize,
                    cell_faces: *const *const u64, edge_lengths_sq: *const c_double, out: *mut *mut fq_mesh) -> c_int;
  // one rank's part of a Complex under owner-computes (multi-GPU; host split: formoniq_b200.dist.partition_mesh)
  fn fq_mesh_create_part(ctx: *mut fq_ctx, dim: c_int, ncells_held: usize, nsimplices: *const usize,
                         cell_faces: *const *const u64, edge_lengths_sq: *const c_double, own_lo: *const usize,
                         own_hi: *const usize, out: *mut *mut fq_mesh) -> c_int;
  fn fq_mesh_destroy(mesh: *mut fq_mesh) -> c_int;
  fn fq_assemble(ctx: *mut fq_ctx, mesh: *const fq_mesh, kind: c_int, grade: c_int, drop_exact_zeros: c_int,
                 out: *mut *mut fq_csr) -> c_int;
  fn fq_assemble_symbolic(ctx: *mut fq_ctx, mesh: *const fq_mesh, kind: c_int, grade: c_int, row_begin: usize,
                          row_end: usize, out: *mut *mut fq_csr) -> c_int;
  fn fq_csr_shape(csr: *const fq_csr, nrows: *mut usize, ncols: *mut usize, nnz: *mut usize) -> c_int;
  fn fq_csr_download(ctx: *mut fq_ctx, csr: *const fq_csr, row_offsets: *mut usize, col_indices: *mut usize,
                     values: *mut c_double) -> c_int;
  fn fq_csr_upload(ctx: *mut fq_ctx, nrows: usize, ncols: usize, row_offsets: *const usize, col_indices: *const usize,
                   values: *const c_double, out: *mut *mut fq_csr) -> c_int;
  fn fq_csr_destroy(csr: *mut fq_csr) -> c_int;
  fn fq_vec_create(ctx: *mut fq_ctx, n: usize, out: *mut *mut fq_vec) -> c_int;
  fn fq_vec_destroy(v: *mut fq_vec) -> c_int;
  fn fq_vec_len(v: *const fq_vec) -> usize;
  fn fq_vec_upload(ctx: *mut fq_ctx, v: *mut fq_vec, host: *const c_double) -> c_int;
  fn fq_vec_download(ctx: *mut fq_ctx, v: *const fq_vec, host: *mut c_double) -> c_int;
  fn fq_vec_copy(ctx: *mut fq_ctx, dst: *mut fq_vec, src: *const fq_vec) -> c_int;
  fn fq_vec_dot(ctx: *mut fq_ctx, x: *const fq_vec, y: *const fq_vec, out: *mut c_double) -> c_int;
  fn fq_vec_scale(ctx: *mut fq_ctx, x: *mut fq_vec, alpha: c_double) -> c_int;
  fn fq_vec_axpy(ctx: *mut fq_ctx, y: *mut fq_vec, alpha: c_double, x: *const fq_vec) -> c_int;
  fn fq_spmv(ctx: *mut fq_ctx, a: *const fq_csr, x: *const fq_vec, y: *mut fq_vec) -> c_int;
  fn fq_linear_form_create(ctx: *mut fq_ctx, mesh: *const fq_mesh, grade: c_int, out: *mut *mut fq_matfree) -> c_int;
  fn fq_linear_form_assemble(ctx: *mut fq_ctx, plan: *const fq_matfree, element_vectors: *const c_double,
                             out: *mut fq_vec) -> c_int;
  fn fq_linear_form_destroy(plan: *mut fq_matfree) -> c_int;
  fn fq_weighted_mass_numeric(ctx: *mut fq_ctx, mesh: *const fq_mesh, csr: *mut fq_csr, nnodes: c_int, weights: *const c_double,
                              shapes: *const c_double, coefficient: *const c_double, drop_exact_zeros: c_int) -> c_int;
  fn fq_source_form_assemble(ctx: *mut fq_ctx, plan: *const fq_matfree, nnodes: c_int, weights: *const c_double,
                             shapes: *const c_double, samples: *const c_double, out: *mut fq_vec) -> c_int;
  // HodgeBlocks (hodge.rs:62-99): symbolic once, numeric per geometry (the first pass builds the tile plan, every later pass is one fused kernel)
  fn fq_mesh_set_lengths(ctx: *mut fq_ctx, mesh: *mut fq_mesh, edge_lengths_sq: *const c_double) -> c_int;
  fn fq_hodge_symbolic(ctx: *mut fq_ctx, mesh: *const fq_mesh, grade: c_int, sigma_row_begin: usize, sigma_row_end: usize,
                       u_row_begin: usize, u_row_end: usize, out: *mut *mut fq_hodge) -> c_int;
  fn fq_hodge_numeric(ctx: *mut fq_ctx, mesh: *const fq_mesh, blocks: *mut fq_hodge, drop_exact_zeros: c_int) -> c_int;
  fn fq_hodge_block(blocks: *mut fq_hodge, which: c_int) -> *mut fq_csr;
  fn fq_hodge_mixed_laplacian(ctx: *mut fq_ctx, blocks: *const fq_hodge, out: *mut *mut fq_csr) -> c_int;
  fn fq_hodge_destroy(blocks: *mut fq_hodge) -> c_int;
  // RelativeWhitneyComplex::assemble (whitney_complex.rs:620-624)
  fn fq_csr_restrict(ctx: *mut fq_ctx, a: *const fq_csr, rows_keep: *const usize, nrows_keep: usize,
                     cols_keep: *const usize, ncols_keep: usize, out: *mut *mut fq_csr) -> c_int;
  // hdif_gram (whitney_complex.rs:118-125), symmetrised KKT and the AFW block preconditioner (problems/elliptic.rs:29-47,101-113)
  fn fq_csr_add(ctx: *mut fq_ctx, a: *const fq_csr, b: *const fq_csr, out: *mut *mut fq_csr) -> c_int;
  fn fq_hodge_mixed_kkt_symmetric(ctx: *mut fq_ctx, blocks: *const fq_hodge, out: *mut *mut fq_csr) -> c_int;
  fn fq_minres_blockdiag(ctx: *mut fq_ctx, a: *const fq_csr, nblocks: c_int, blocks: *const *const fq_csr,
                         offsets: *const usize, inner_rtol: c_double, inner_max_iters: usize, b: *const fq_vec,
                         rtol: c_double, max_iters: usize, x: *mut fq_vec, iters: *mut usize, residual: *mut c_double,
                         converged: *mut c_int, inner_iters: *mut usize) -> c_int;
}

/// The reference panics on contract violations (`galerkin.rs:184` unwraps);
/// the shim keeps that behaviour at the Rust boundary.
fn check(rc: c_int) {
  if rc != 0 {
    let msg = unsafe { CStr::from_ptr(fq_last_error()) }.to_string_lossy().into_owned();
    panic!("formoniq_b200 error {rc}: {msg}");
  }
}

/// One CUDA device. `Sync` because every entry point serialises on the
/// context's stream (seam 1 needs `BilinearForm: Sync`).
pub struct Device {
  ctx: *mut fq_ctx,
  /// The last uploaded mesh, keyed by the addresses and sizes of the `Complex` and `MeshLengthsSq` it came from:
  /// `HodgeBlocks::compute` calls `form.assemble(topology, geometry)` four times on the same pair, and materialising
  /// `FaceIncidence` for every grade costs about twice an assembly (galerkin.rs:136-137), so seam 1 must not redo it.
  mesh_cache: std::sync::Mutex<Option<(MeshKey, *mut fq_mesh)>>,
}
#[derive(PartialEq, Eq, Clone, Copy)]
struct MeshKey { topology: usize, geometry: usize, ncells: usize, nedges: usize }
unsafe impl Send for Device {}
unsafe impl Sync for Device {}
impl Device {
  pub fn new(index: i32) -> Self {
    let mut ctx = ptr::null_mut();
    check(unsafe { fq_ctx_create(index, &mut ctx) });
    Self { ctx, mesh_cache: std::sync::Mutex::new(None) }
  }
  /// The device mesh of (topology, geometry): uploaded on first use, reused while the same pair is passed again.
  /// (A caller that mutates the lengths in place keeps the handle and calls `fq_mesh_set_lengths` instead.)
  fn cached_mesh(&self, topology: &Complex, geometry: &MeshLengthsSq) -> *mut fq_mesh {
    let key = MeshKey { topology: topology as *const _ as usize, geometry: geometry as *const _ as usize,
                        ncells: topology.cells().len(), nedges: geometry.vector().len() };
    let mut slot = self.mesh_cache.lock().unwrap();
    if let Some((k, raw)) = *slot { if k == key { return raw; } unsafe { fq_mesh_destroy(raw) }; }
    let raw = upload_mesh(self, topology, geometry);
    *slot = Some((key, raw));
    raw
  }
}
impl Drop for Device {
  fn drop(&mut self) {
    if let Some((_, raw)) = self.mesh_cache.lock().unwrap().take() { unsafe { fq_mesh_destroy(raw) }; }
    unsafe { fq_ctx_destroy(self.ctx) };
  }
}
fn upload_mesh(dev: &Device, topology: &Complex, geometry: &MeshLengthsSq) -> *mut fq_mesh {
  let dim = topology.dim().index();
  let nsimplices: Vec<usize> = (0..=dim).map(|j| topology.nsimplices(j)).collect();
  let tables: Vec<Vec<u64>> = (0..=dim)
    .map(|j| FaceIncidence::new(topology, j).faces_flat().iter().map(|&i| i as u64).collect())
    .collect();
  let ptrs: Vec<*const u64> = tables.iter().map(|t| t.as_ptr()).collect();
  let mut raw = ptr::null_mut();
  check(unsafe {
    fq_mesh_create(dev.ctx, dim as c_int, topology.cells().len(), nsimplices.as_ptr(), ptrs.as_ptr(),
                   geometry.vector().as_ptr(), &mut raw)
  });
  raw
}

/// `Complex` + `MeshLengthsSq` uploaded once: the FaceIncidence tables of every
/// grade (incidence.rs:42-53) and the edge lengths (lengths/mesh.rs:34-36).
pub struct DeviceMesh<'d> { dev: &'d Device, raw: *mut fq_mesh, dim: usize }
impl<'d> DeviceMesh<'d> {
  pub fn new(dev: &'d Device, topology: &Complex, geometry: &MeshLengthsSq) -> Self {
    Self { dev, raw: upload_mesh(dev, topology, geometry), dim: topology.dim().index() }
  }
}
impl Drop for DeviceMesh<'_> { fn drop(&mut self) { unsafe { fq_mesh_destroy(self.raw) }; } }

fn download(dev: &Device, csr: *mut fq_csr) -> GalerkinMatrix {
  let m = download_borrowed(dev, csr);
  unsafe { fq_csr_destroy(csr) };
  m
}
/// Same for a matrix the library keeps (a block of an `fq_hodge` plan).
fn download_borrowed(dev: &Device, csr: *mut fq_csr) -> GalerkinMatrix {
  let (mut nr, mut nc, mut nnz) = (0usize, 0usize, 0usize);
  check(unsafe { fq_csr_shape(csr, &mut nr, &mut nc, &mut nnz) });
  let (mut rp, mut ci, mut va) = (vec![0usize; nr + 1], vec![0usize; nnz], vec![0f64; nnz]);
  check(unsafe { fq_csr_download(dev.ctx, csr, rp.as_mut_ptr(), ci.as_mut_ptr(), va.as_mut_ptr()) });
  // the data contract handed to faer by linalg/faer.rs:16-24
  CsrMatrix::try_from_csr_data(nr, nc, rp, ci, va).unwrap()
}

/// Seam 1: a `BilinearForm` whose *provided* `assemble` (galerkin.rs:52-57) is
/// overridden to run on the GPU; `element` stays the reference's own, so the
/// CPU parity check is one call away.
pub struct GpuPairing<'d> { dev: &'d Device, kind: c_int, grade: c_int, cpu: WhitneyPairing }
impl<'d> GpuPairing<'d> {
  // operators.rs:169-191: the four constructors, grade = grade of the inner product
  pub fn mass(dev: &'d Device, dim: usize, grade: usize) -> Self {
    Self { dev, kind: 0, grade: grade as c_int, cpu: WhitneyPairing::mass(dim, grade) }
  }
  pub fn dif_trial(dev: &'d Device, dim: usize, grade: usize) -> Self {
    Self { dev, kind: 1, grade: grade as c_int, cpu: WhitneyPairing::dif_trial(dim, grade) }
  }
  pub fn dif_test(dev: &'d Device, dim: usize, grade: usize) -> Self {
    Self { dev, kind: 2, grade: grade as c_int, cpu: WhitneyPairing::dif_test(dim, grade) }
  }
  pub fn dif_both(dev: &'d Device, dim: usize, grade: usize) -> Self {
    Self { dev, kind: 3, grade: grade as c_int, cpu: WhitneyPairing::dif_both(dim, grade) }
  }
}
impl BilinearForm for GpuPairing<'_> {
  fn test_grade(&self) -> multialgebra::ExteriorGrade { self.cpu.test_grade() }
  fn trial_grade(&self) -> multialgebra::ExteriorGrade { self.cpu.trial_grade() }
  fn element(&self, metric: &metric::Metric, chart: simplicial::atlas::Chart) -> simplicial::linalg::Matrix {
    self.cpu.element(metric, chart)
  }
  fn assemble(&self, topology: &Complex, geometry: &MeshLengthsSq) -> GalerkinMatrix {
    let mesh = self.dev.cached_mesh(topology, geometry);  // one upload for the four calls of HodgeBlocks::compute
    let mut csr = ptr::null_mut();
    check(unsafe { fq_assemble(self.dev.ctx, mesh, self.kind, self.grade, 1, &mut csr) });
    download(self.dev, csr)
  }
}
// Seam 2 needs no code: `WhitneyComplex::assemble` (whitney_complex.rs:407-409) calls
// `form.assemble(topology, geometry)`, so `complex.pairing(&GpuPairing::mass(..))`,
// `HodgeBlocks::compute` and everything in `problems::*` reach the GPU through seam 1.
// A caller that assembles several blocks on one mesh keeps a `DeviceMesh` and calls
// `fq_hodge_symbolic` / `fq_hodge_numeric` (one fused element kernel for the four blocks).

/// Seam 3a: a vector that does not live in host memory (iterative/src/lib.rs:58-84).
pub struct DeviceVector<'d> { dev: &'d Device, raw: *mut fq_vec }
impl Clone for DeviceVector<'_> {
  fn clone(&self) -> Self {
    let out = self.zeros_like();
    check(unsafe { fq_vec_copy(self.dev.ctx, out.raw, self.raw) });
    out
  }
}
impl Drop for DeviceVector<'_> { fn drop(&mut self) { unsafe { fq_vec_destroy(self.raw) }; } }
impl<'d> DeviceVector<'d> {
  pub fn zeros(dev: &'d Device, n: usize) -> Self {
    let mut raw = ptr::null_mut();
    check(unsafe { fq_vec_create(dev.ctx, n, &mut raw) });
    Self { dev, raw }
  }
  pub fn to_host(&self) -> nalgebra::DVector<f64> {
    let mut host = nalgebra::DVector::zeros(unsafe { fq_vec_len(self.raw) });
    check(unsafe { fq_vec_download(self.dev.ctx, self.raw, host.as_mut_ptr()) });
    host
  }
}
impl InnerProductSpace for DeviceVector<'_> {
  type Scalar = f64;
  fn zeros_like(&self) -> Self {
    let mut raw = ptr::null_mut();
    check(unsafe { fq_vec_create(self.dev.ctx, fq_vec_len(self.raw), &mut raw) });
    Self { dev: self.dev, raw }
  }
  fn dot(&self, other: &Self) -> f64 {
    let mut out = 0.0;
    check(unsafe { fq_vec_dot(self.dev.ctx, self.raw, other.raw, &mut out) });
    out
  }
  fn scale(&mut self, alpha: f64) { check(unsafe { fq_vec_scale(self.dev.ctx, self.raw, alpha) }); }
  fn add_scaled(&mut self, alpha: f64, x: &Self) { check(unsafe { fq_vec_axpy(self.dev.ctx, self.raw, alpha, x.raw) }); }
}

/// Seam 3b: the assembled operator applied on the device (iterative/src/operator.rs:5-14).
pub struct DeviceCsr<'d> { dev: &'d Device, raw: *mut fq_csr, n: usize }
impl<'d> DeviceCsr<'d> {
  pub fn upload(dev: &'d Device, m: &CsrMatrix<f64>) -> Self {
    let mut raw = ptr::null_mut();
    check(unsafe {
      fq_csr_upload(dev.ctx, m.nrows(), m.ncols(), m.row_offsets().as_ptr(), m.col_indices().as_ptr(),
                    m.values().as_ptr(), &mut raw)
    });
    Self { dev, raw, n: m.nrows() }
  }
}
impl Drop for DeviceCsr<'_> { fn drop(&mut self) { unsafe { fq_csr_destroy(self.raw) }; } }
impl<'d> LinearOperator for DeviceCsr<'d> {
  type Space = DeviceVector<'d>;
  fn dim(&self) -> usize { self.n }
  fn apply(&self, x: &Self::Space) -> Self::Space {
    let y = x.zeros_like();
    check(unsafe { fq_spmv(self.dev.ctx, self.raw, x.raw, y.raw) });
    y
  }
}

/// `HodgeBlocks::compute` (hodge.rs:62-72) with the plan kept alive: `refresh` re-runs the numeric phase after
/// `MeshLengthsSq` changed (time stepping, Regge flow); from the second pass on that is ONE fused kernel.
pub struct GpuHodgeBlocks<'d> { dev: &'d Device, mesh: DeviceMesh<'d>, raw: *mut fq_hodge }
impl<'d> GpuHodgeBlocks<'d> {
  pub fn compute(dev: &'d Device, topology: &Complex, geometry: &MeshLengthsSq, grade: usize) -> Self {
    let mesh = DeviceMesh::new(dev, topology, geometry);
    let mut raw = ptr::null_mut();
    check(unsafe { fq_hodge_symbolic(dev.ctx, mesh.raw, grade as c_int, 0, usize::MAX, 0, usize::MAX, &mut raw) });
    check(unsafe { fq_hodge_numeric(dev.ctx, mesh.raw, raw, 1) });
    Self { dev, mesh, raw }
  }
  pub fn refresh(&mut self, geometry: &MeshLengthsSq) {
    check(unsafe { fq_mesh_set_lengths(self.dev.ctx, self.mesh.raw, geometry.vector().as_ptr()) });
    check(unsafe { fq_hodge_numeric(self.dev.ctx, self.mesh.raw, self.raw, 1) });
  }
  /// 0 mass_sigma, 1 mass_u, 2 dif_test, 3 dif_both as `GalerkinMatrix` (host CSR, usize indices)
  pub fn block(&self, which: usize) -> GalerkinMatrix {
    download_borrowed(self.dev, unsafe { fq_hodge_block(self.raw, which as c_int) })
  }
  /// `mixed_hodge_laplacian` (hodge.rs:93-99) stitched on the device, left there for the Krylov solve
  pub fn mixed_hodge_laplacian(&self) -> DeviceCsr<'d> {
    let mut raw = ptr::null_mut();
    check(unsafe { fq_hodge_mixed_laplacian(self.dev.ctx, self.raw, &mut raw) });
    DeviceCsr { dev: self.dev, raw, n: self.block(0).nrows() + self.block(1).nrows() }
  }
}
impl<'d> GpuHodgeBlocks<'d> {
  /// The symmetric saddle point `assemble_mixed_kkt` hands to MINRES (problems/elliptic.rs:101-113): sigma rows negated.
  pub fn mixed_kkt_symmetric(&self) -> DeviceCsr<'d> {
    let mut raw = ptr::null_mut();
    check(unsafe { fq_hodge_mixed_kkt_symmetric(self.dev.ctx, self.raw, &mut raw) });
    DeviceCsr { dev: self.dev, raw, n: self.block(0).nrows() + self.block(1).nrows() }
  }
}
impl Drop for GpuHodgeBlocks<'_> { fn drop(&mut self) { unsafe { fq_hodge_destroy(self.raw) }; } }

impl<'d> DeviceCsr<'d> {
  /// `hdif_gram(k) = mass(k) + dif_both(k + 1)` (whitney_complex.rs:118-125) as a device CSR add on the union pattern.
  pub fn add(&self, other: &DeviceCsr<'d>) -> DeviceCsr<'d> {
    let mut raw = ptr::null_mut();
    check(unsafe { fq_csr_add(self.dev.ctx, self.raw, other.raw, &mut raw) });
    DeviceCsr { dev: self.dev, raw, n: self.n }
  }
}

/// MINRES on the symmetrised KKT operator with the AFW block-diagonal preconditioner (problems/elliptic.rs:29-47): the two
/// `hdif_gram` blocks are solved by inner Jacobi-CG on the device where the reference applies a sparse Cholesky factor.
/// Returns (solution, outer iterations, relative residual, converged).
pub fn minres_afw<'d>(kkt: &DeviceCsr<'d>, hdif_gram_sigma: &DeviceCsr<'d>, hdif_gram_u: &DeviceCsr<'d>, rhs: &DeviceVector<'d>,
                      rtol: f64, max_iters: usize, inner_rtol: f64, inner_max_iters: usize) -> (DeviceVector<'d>, usize, f64, bool) {
  let x = DeviceVector::zeros(kkt.dev, kkt.n);
  let blocks = [hdif_gram_sigma.raw as *const fq_csr, hdif_gram_u.raw as *const fq_csr];
  let offsets = [0usize, hdif_gram_sigma.n, kkt.n];
  let (mut iters, mut residual, mut converged, mut inner) = (0usize, 0f64, 0 as c_int, 0usize);
  check(unsafe { fq_minres_blockdiag(kkt.dev.ctx, kkt.raw, 2, blocks.as_ptr(), offsets.as_ptr(), inner_rtol, inner_max_iters,
                                     rhs.raw, rtol, max_iters, x.raw, &mut iters, &mut residual, &mut converged, &mut inner) });
  (x, iters, residual, converged != 0)
}

// ---- WeightedHodgeMass (formoniq/src/operators.rs:432-486) -------------------------------------------------------------
/// Drop-in for `WeightedHodgeMass::new(coefficient, grade, qr).assemble(topology, geometry)`: the rule and the shape
/// table are built as `WeightedHodgeMass::new` does, the scalar coefficient is sampled at the nodes of every cell with
/// rayon, element quadrature and scatter run on the device.
pub fn assemble_weighted_mass<F: Sync + formoniq::Section>(dev: &Device, mesh: &DeviceMesh, topology: &Complex, coefficient: &F,
                                                           grade: usize, qr: Option<simplicial::atlas::quadrature::SimplexQuadRule>)
  -> GalerkinMatrix {
  use rayon::prelude::*;
  let dim = coefficient.dim();
  let qr = qr.unwrap_or(simplicial::atlas::quadrature::SimplexQuadRule::degree(dim, 1));
  let nodes: Vec<_> = qr.points().map(|b| b.to_coords()).collect();
  let shapes = derham::interpolate::samples::LsfSamples::whitney(dim, grade, &nodes);
  let weights: Vec<f64> = qr.weights().iter().copied().collect();
  let mut table = Vec::new();
  for q in 0..nodes.len() { for w in shapes.at_node(q) { table.extend(w.components().iter().copied()); } }
  let alpha: Vec<f64> = topology.cells().handle_par_iter()
    .flat_map_iter(|cell| nodes.iter().map(|b| coefficient.at(&cell.point(b.clone())).as_scalar()).collect::<Vec<_>>())
    .collect();
  let mut raw = ptr::null_mut();
  check(unsafe { fq_assemble_symbolic(dev.ctx, mesh.raw, 0 /*FQ_MASS*/, grade as c_int, 0, usize::MAX, &mut raw) });
  check(unsafe { fq_weighted_mass_numeric(dev.ctx, mesh.raw, raw, nodes.len() as c_int, weights.as_ptr(), table.as_ptr(),
                                          alpha.as_ptr(), 1) });
  download(dev, raw)
}

// ---- LinearForm::assemble (formoniq/src/galerkin.rs:279-312) -----------------------------------------------------
// `LinearForm::element` evaluates a user `Section` at quadrature nodes and stays host code; the element vectors of all
// cells are evaluated with rayon exactly as `assemble_vector` does and handed over cell-major, the device does the
// per-DOF sums (cells in order, no atomics).  One plan per (mesh, grade) serves every right-hand side.
pub struct GpuLinearFormPlan<'d> { dev: &'d Device, raw: *mut fq_matfree, grade: usize, ndofs: usize }
impl<'d> GpuLinearFormPlan<'d> {
  pub fn new(dev: &'d Device, mesh: &DeviceMesh<'d>, topology: &Complex, grade: usize) -> Self {
    let mut raw = ptr::null_mut();
    check(unsafe { fq_linear_form_create(dev.ctx, mesh.raw, grade as c_int, &mut raw) });
    Self { dev, raw, grade, ndofs: topology.skeleton(grade).len() }
  }
  /// Drop-in for `form.assemble(topology, geometry)` of any `LinearForm` of this plan's grade.
  pub fn assemble(&self, topology: &Complex, geometry: &MeshLengthsSq, form: &impl formoniq::galerkin::LinearForm)
    -> formoniq::galerkin::GalerkinVector {
    use rayon::prelude::*;
    assert_eq!(form.test_grade().index(), self.grade, "the plan was built for another grade");
    let elvecs: Vec<f64> = topology.cells().handle_par_iter()
      .flat_map_iter(|cell| form.element(&geometry.cell_metric(cell), cell).iter().copied().collect::<Vec<_>>())
      .collect();
    let out = DeviceVector::zeros(self.dev, self.ndofs);
    check(unsafe { fq_linear_form_assemble(self.dev.ctx, self.raw, elvecs.as_ptr(), out.raw) });
    formoniq::galerkin::GalerkinVector::new(form.test_grade(), out.to_host())
  }
}
impl<'d> GpuLinearFormPlan<'d> {
  /// Drop-in for `SourceForm::new(source, qr).assemble(topology, geometry)` (operators.rs:607-635): the rule and the
  /// `LsfSamples::whitney` table are built exactly as `SourceForm::new` does, the `Section` is sampled at the nodes of
  /// every cell with rayon, and quadrature + scatter run on the device.
  pub fn assemble_source<F: Sync + formoniq::Section>(&self, topology: &Complex, source: &F,
                                                      qr: Option<simplicial::atlas::quadrature::SimplexQuadRule>)
    -> formoniq::galerkin::GalerkinVector {
    use rayon::prelude::*;
    let dim = source.dim();
    let qr = qr.unwrap_or(simplicial::atlas::quadrature::SimplexQuadRule::degree(dim, 1));
    let nodes: Vec<_> = qr.points().map(|b| b.to_coords()).collect();
    let shapes = derham::interpolate::samples::LsfSamples::whitney(dim, source.grade(), &nodes);
    let weights: Vec<f64> = qr.weights().iter().copied().collect();
    let mut table = Vec::new();
    for q in 0..nodes.len() { for w in shapes.at_node(q) { table.extend(w.components().iter().copied()); } }
    let samples: Vec<f64> = topology.cells().handle_par_iter()
      .flat_map_iter(|cell| nodes.iter().flat_map(|b| source.at(&cell.point(b.clone())).components().iter().copied()
        .collect::<Vec<_>>()).collect::<Vec<_>>())
      .collect();
    let out = DeviceVector::zeros(self.dev, self.ndofs);
    check(unsafe { fq_source_form_assemble(self.dev.ctx, self.raw, nodes.len() as c_int, weights.as_ptr(), table.as_ptr(),
                                           samples.as_ptr(), out.raw) });
    formoniq::galerkin::GalerkinVector::new(source.grade(), out.to_host())
  }
}
impl Drop for GpuLinearFormPlan<'_> { fn drop(&mut self) { unsafe { fq_linear_form_destroy(self.raw) }; } }

// `iterative::krylov::{cg, minres}` now run unmodified on `DeviceCsr` with
// `iterative::precond::Identity<DeviceVector>` (precond.rs:16-41).
#[allow(dead_code)]
fn _uses(_: *mut c_void) {}
