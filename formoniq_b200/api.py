"""Host-side mirror of the reference interfaces for the assembly path, on top
of the C ABI.  Names, argument meaning and error behaviour follow the
reference (paths relative to the reference checkout):

  WhitneyPairing / ScalarLumpedMass   formoniq/src/operators.rs:27-40,150-211
  BilinearForm.assemble               formoniq/src/galerkin.rs:42-58,138-188
  HodgeBlocks.compute                 formoniq/src/hodge.rs:62-72
  DeviceVector (InnerProductSpace)    iterative/src/lib.rs:84-141
  DeviceCsr (LinearOperator)          iterative/src/operator.rs:5-14
  cg / minres / StopCriterion / Report   iterative/src/krylov.rs:48-211, lib.rs:186-219
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from math import comb

import numpy as np

from . import _lib
from ._lib import FQ_DIF_BOTH, FQ_DIF_TEST, FQ_DIF_TRIAL, FQ_LUMPED, FQ_MASS, FormoniqError, check

SIZE_MAX = (1 << 64) - 1


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def nlocal(dim: int, grade: int) -> int:
    return 0 if grade < 0 or grade > dim else comb(dim + 1, grade + 1)


class Context:
    """One CUDA device + stream (fq_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = C.c_void_p()
        check(_lib.lib().fq_ctx_create(device, C.byref(h)))
        self._h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, cuda_stream: int):
        check(_lib.lib().fq_ctx_set_stream(self._h, C.c_void_p(cuda_stream)))

    def wait_downloads(self):
        """Block until every download_async issued on this context has landed in host memory."""
        check(_lib.lib().fq_ctx_wait_downloads(self._h))

    def synchronize(self):
        check(_lib.lib().fq_ctx_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return _lib.lib().fq_ctx_launch_count(self._h)

    def set_timing(self, on: bool):
        check(_lib.lib().fq_ctx_set_timing(self._h, int(on)))

    def timing_report(self) -> dict:
        """Per-kernel device time since the last report: {name: {"ms", "count"}}."""
        import json

        buf = C.create_string_buffer(8192)
        check(_lib.lib().fq_ctx_timing_report(self._h, buf, 8192))
        return json.loads(buf.value.decode())

    def __del__(self):
        try:
            _lib.lib().fq_ctx_destroy(self._h)
        except Exception:
            pass


class Mesh:
    """`Complex` + `MeshLengthsSq` as assembly consumes them: per-grade
    FaceIncidence tables and one signed squared length per edge."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self._h = ctx, handle
        L = _lib.lib()
        self.dim = L.fq_mesh_dim(handle)
        self.ncells = L.fq_mesh_ncells(handle)

    @classmethod
    def from_arrays(cls, ctx: Context, dim: int, nsimplices, cell_faces, edge_lengths_sq) -> "Mesh":
        """cell_faces: sequence indexed by grade of [ncells, C(dim+1, j+1)] integer arrays (None to skip)."""
        ns = np.ascontiguousarray(nsimplices, dtype=np.uint64)
        faces = [None if f is None else np.ascontiguousarray(f, dtype=np.uint64) for f in cell_faces]
        ptrs = (C.c_void_p * (dim + 1))(*[None if f is None else f.ctypes.data for f in faces])
        lengths = np.ascontiguousarray(edge_lengths_sq, dtype=np.float64)
        if lengths.shape[0] != int(ns[1]):
            raise FormoniqError(-1, "one squared length per edge is required")
        h = C.c_void_p()
        check(_lib.lib().fq_mesh_create(ctx._h, dim, int(ns[dim]), _p(ns), C.cast(ptrs, C.c_void_p), _p(lengths), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_part(cls, ctx: Context, dim: int, nsimplices, cell_faces, edge_lengths_sq, part) -> "Mesh":
        """One rank's part of an uploaded mesh (dist.partition_mesh): the held cells' rows of the GLOBAL face tables
        (global ids, global counts, global edge lengths); owned_range(j) then gives the rank's row range of grade j."""
        ns = np.ascontiguousarray(nsimplices, dtype=np.uint64)
        faces = [None if f is None else np.ascontiguousarray(np.asarray(f)[part.cells], dtype=np.uint64) for f in cell_faces]
        if faces[dim] is None:
            faces[dim] = np.ascontiguousarray(part.cells, dtype=np.uint64).reshape(-1, 1)
        ptrs = (C.c_void_p * (dim + 1))(*[None if f is None else f.ctypes.data for f in faces])
        lengths = np.ascontiguousarray(edge_lengths_sq, dtype=np.float64)
        if lengths.shape[0] != int(ns[1]):
            raise FormoniqError(-1, "one squared length per edge is required")
        lo = np.ascontiguousarray([r[0] for r in part.own], dtype=np.uint64)
        hi = np.ascontiguousarray([r[1] for r in part.own], dtype=np.uint64)
        h = C.c_void_p()
        check(_lib.lib().fq_mesh_create_part(ctx._h, dim, int(len(part.cells)), _p(ns), C.cast(ptrs, C.c_void_p), _p(lengths),
                                             _p(lo), _p(hi), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def kuhn(cls, ctx: Context, dim: int, shape, vmin=None, vmax=None, ambient_diag=None, jitter: float = 0.0,
             slab=None) -> "Mesh":
        """CartesianGrid::triangulate + to_edge_lengths_sq generated on the device."""
        if np.isscalar(shape):
            shape = [int(shape)] * dim
        shp = np.ascontiguousarray(shape, dtype=np.uint64)
        vmin = None if vmin is None else np.ascontiguousarray(vmin, dtype=np.float64)
        vmax = None if vmax is None else np.ascontiguousarray(vmax, dtype=np.float64)
        diag = None if ambient_diag is None else np.ascontiguousarray(ambient_diag, dtype=np.float64)
        sb, se = (0, int(shp[dim - 1])) if slab is None else slab
        h = C.c_void_p()
        check(_lib.lib().fq_mesh_create_kuhn(ctx._h, dim, _p(shp), _p(vmin), _p(vmax), _p(diag), float(jitter), sb, se,
                                             C.byref(h)))
        m = cls(ctx, h)
        m.shape = [int(s) for s in shp]
        return m

    def nsimplices(self, grade: int) -> int:
        return _lib.lib().fq_mesh_nsimplices(self._h, grade)

    @property
    def nowned_cells(self) -> int:
        return _lib.lib().fq_mesh_nowned_cells(self._h)

    def owned_range(self, grade: int):
        """Rows (simplices of `grade`) this slab owns under owner-computes."""
        lo, hi = C.c_size_t(), C.c_size_t()
        check(_lib.lib().fq_mesh_owned_range(self._h, grade, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def held_range(self, grade: int):
        """Ids of `grade` referenced by the held cells (owned + halos), contiguous."""
        lo, hi = C.c_size_t(), C.c_size_t()
        check(_lib.lib().fq_mesh_held_range(self._h, grade, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def set_lengths(self, edge_lengths_sq):
        a = np.ascontiguousarray(edge_lengths_sq, dtype=np.float64)
        if a.shape[0] != self.nsimplices(1):
            raise FormoniqError(-1, "one squared length per edge is required")
        check(_lib.lib().fq_mesh_set_lengths(self.ctx._h, self._h, _p(a)))

    def cell_faces(self, grade: int) -> np.ndarray:
        out = np.zeros((self.ncells, nlocal(self.dim, grade)), dtype=np.uint64)
        check(_lib.lib().fq_mesh_download_cell_faces(self.ctx._h, self._h, grade, _p(out)))
        return out

    def lengths(self) -> np.ndarray:
        out = np.full(self.nsimplices(1), np.nan)
        check(_lib.lib().fq_mesh_download_lengths(self.ctx._h, self._h, _p(out)))
        return out

    def __del__(self):
        try:
            _lib.lib().fq_mesh_destroy(self._h)
        except Exception:
            pass


def kuhn_counts(dim: int, shape) -> list[int]:
    if np.isscalar(shape):
        shape = [int(shape)] * dim
    shp = np.ascontiguousarray(shape, dtype=np.uint64)
    out = np.zeros(dim + 1, dtype=np.uint64)
    check(_lib.lib().fq_kuhn_counts(dim, _p(shp), _p(out)))
    return [int(v) for v in out]


def kuhn_slab_ranges(dim: int, shape, slab, grade: int):
    """(held_lo, own_lo, own_hi, held_hi) of one grade for a slab, host-only closed form."""
    if np.isscalar(shape):
        shape = [int(shape)] * dim
    shp = np.ascontiguousarray(shape, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    check(_lib.lib().fq_kuhn_slab_ranges(dim, _p(shp), int(slab[0]), int(slab[1]), grade, _p(out)))
    return tuple(int(v) for v in out)


def kuhn_cell_faces_host(dim: int, shape, grade: int) -> np.ndarray:
    """Closed-form colex numbering of the Kuhn grid, evaluated on the host."""
    if np.isscalar(shape):
        shape = [int(shape)] * dim
    shp = np.ascontiguousarray(shape, dtype=np.uint64)
    ncells = kuhn_counts(dim, shape)[dim]
    out = np.zeros((ncells, nlocal(dim, grade)), dtype=np.uint64)
    check(_lib.lib().fq_kuhn_cell_faces_host(dim, _p(shp), grade, _p(out)))
    return out


@dataclass
class StopCriterion:
    rtol: float
    max_iters: int = 10_000


@dataclass
class Report:
    iters: int
    residual: float
    converged: bool


class DeviceVector:
    """A vector in HBM implementing iterative::InnerProductSpace."""

    def __init__(self, ctx: Context, n: int):
        h = C.c_void_p()
        check(_lib.lib().fq_vec_create(ctx._h, n, C.byref(h)))
        self.ctx, self._h, self.n = ctx, h, n

    @classmethod
    def from_numpy(cls, ctx: Context, a) -> "DeviceVector":
        a = np.ascontiguousarray(a, dtype=np.float64)
        v = cls(ctx, a.shape[0])
        check(_lib.lib().fq_vec_upload(ctx._h, v._h, _p(a)))
        return v

    @classmethod
    def wrap(cls, ctx: Context, device_ptr: int, n: int, keepalive=None) -> "DeviceVector":
        """Non-owning view of caller-owned device memory (e.g. a torch CUDA tensor)."""
        v = cls.__new__(cls)
        h = C.c_void_p()
        check(_lib.lib().fq_vec_wrap(ctx._h, C.c_void_p(device_ptr), n, C.byref(h)))
        v.ctx, v._h, v.n, v._keepalive = ctx, h, n, keepalive
        return v

    @classmethod
    def from_torch(cls, ctx: Context, t) -> "DeviceVector":
        import torch

        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.dim() == 1):
            raise FormoniqError(-1, "need a contiguous 1-D float64 CUDA tensor")
        return cls.wrap(ctx, t.data_ptr(), t.numel(), keepalive=t)

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape[0] != self.n:
            raise FormoniqError(-1, "vector length mismatch")
        check(_lib.lib().fq_vec_upload(self.ctx._h, self._h, _p(a)))

    def to_numpy(self) -> np.ndarray:
        out = np.zeros(self.n)
        check(_lib.lib().fq_vec_download(self.ctx._h, self._h, _p(out)))
        return out

    def __len__(self):
        return self.n

    def zeros_like(self) -> "DeviceVector":
        return DeviceVector(self.ctx, self.n)

    def clone(self) -> "DeviceVector":
        v = DeviceVector(self.ctx, self.n)
        check(_lib.lib().fq_vec_copy(self.ctx._h, v._h, self._h))
        return v

    def view(self, offset: int, n: int) -> "DeviceVector":
        """Non-owning view of self[offset : offset + n] (the sigma / u segments of a mixed vector)."""
        v = DeviceVector.__new__(DeviceVector)
        h = C.c_void_p()
        check(_lib.lib().fq_vec_view(self.ctx._h, self._h, int(offset), int(n), C.byref(h)))
        v.ctx, v._h, v.n, v._keepalive = self.ctx, h, int(n), self
        return v

    def copy_from(self, src: "DeviceVector"):
        check(_lib.lib().fq_vec_copy(self.ctx._h, self._h, src._h))

    def fill_zero(self):
        self.scale(0.0)

    def dot(self, other: "DeviceVector") -> float:
        out = C.c_double()
        check(_lib.lib().fq_vec_dot(self.ctx._h, self._h, other._h, C.byref(out)))
        return out.value

    def scale(self, alpha: float):
        check(_lib.lib().fq_vec_scale(self.ctx._h, self._h, float(alpha)))

    def add_scaled(self, alpha: float, x: "DeviceVector"):
        check(_lib.lib().fq_vec_axpy(self.ctx._h, self._h, float(alpha), x._h))

    def add(self, x: "DeviceVector"):
        self.add_scaled(1.0, x)

    def assign_product(self, d: "DeviceVector", r: "DeviceVector"):
        """self = d .* r (component-wise; the Jacobi preconditioner applied to r)."""
        check(_lib.lib().fq_vec_mul(self.ctx._h, self._h, d._h, r._h))

    def norm(self) -> float:
        return float(np.sqrt(self.dot(self)))

    @property
    def device_ptr(self) -> int:
        return _lib.lib().fq_vec_device_ptr(self._h)

    def __del__(self):
        try:
            _lib.lib().fq_vec_destroy(self._h)
        except Exception:
            pass


class DeviceCsr:
    """GalerkinMatrix in HBM: nalgebra-sparse CSR contract + LinearOperator::apply."""

    def __init__(self, ctx: Context, handle, owner=None):
        self.ctx, self._h, self._owner = ctx, handle, owner  # owner: a HodgeBlocks plan that owns the handle

    @property
    def shape(self):
        nr, nc, nnz = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(_lib.lib().fq_csr_shape(self._h, C.byref(nr), C.byref(nc), C.byref(nnz)))
        return nr.value, nc.value

    @property
    def nnz(self) -> int:
        nnz = C.c_size_t()
        check(_lib.lib().fq_csr_shape(self._h, None, None, C.byref(nnz)))
        return nnz.value

    @property
    def row_range(self):
        b, e = C.c_size_t(), C.c_size_t()
        check(_lib.lib().fq_csr_row_range(self._h, C.byref(b), C.byref(e)))
        return b.value, e.value

    def dim(self) -> int:
        return self.shape[0]

    @classmethod
    def upload(cls, ctx: Context, nrows, ncols, row_offsets, col_indices, values) -> "DeviceCsr":
        rp = np.ascontiguousarray(row_offsets, dtype=np.uint64)
        ci = np.ascontiguousarray(col_indices, dtype=np.uint64)
        va = np.ascontiguousarray(values, dtype=np.float64)
        h = C.c_void_p()
        check(_lib.lib().fq_csr_upload(ctx._h, nrows, ncols, _p(rp), _p(ci), _p(va), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_scipy(cls, ctx: Context, m) -> "DeviceCsr":
        m = m.tocsr()
        m.sort_indices()
        return cls.upload(ctx, m.shape[0], m.shape[1], m.indptr, m.indices, m.data)

    def download(self, out=None):
        """(row_offsets, col_indices, values) as usize/usize/f64 host arrays.

        `out` = (row_offsets, col_indices, values) buffers to fill (uint64, uint64, float64, at least as long as
        needed); pass views of pinned memory to get PCIe-rate copies.  Returns views of the filled parts."""
        b, e = self.row_range
        nnz = self.nnz
        if out is None:
            rp = np.empty(e - b + 1, dtype=np.uint64)
            ci = np.empty(nnz, dtype=np.uint64)
            va = np.empty(nnz)
        else:
            rp, ci, va = out[0][:e - b + 1], out[1][:nnz], out[2][:nnz]
            if rp.dtype != np.uint64 or ci.dtype != np.uint64 or va.dtype != np.float64:
                raise FormoniqError(-1, "download buffers must be uint64 / uint64 / float64")
            if rp.shape[0] != e - b + 1 or ci.shape[0] != nnz or va.shape[0] != nnz:
                raise FormoniqError(-1, "download buffers are too small")
        check(_lib.lib().fq_csr_download(self.ctx._h, self._h, _p(rp), _p(ci), _p(va)))
        return rp, ci, va

    def download_async(self, out):
        """Enqueue the download into `out` = (row_offsets, col_indices, values) buffers (pinned memory) on the copy
        stream; it overlaps the work submitted afterwards.  Call ctx.wait_downloads() before reading the buffers, and
        keep this matrix alive until then.  Returns views of the parts that will be filled."""
        b, e = self.row_range
        nnz = self.nnz
        rp, ci, va = out[0][:e - b + 1], out[1][:nnz], out[2][:nnz]
        if rp.dtype != np.uint64 or ci.dtype != np.uint64 or va.dtype != np.float64:
            raise FormoniqError(-1, "download buffers must be uint64 / uint64 / float64")
        if rp.shape[0] != e - b + 1 or ci.shape[0] != nnz or va.shape[0] != nnz:
            raise FormoniqError(-1, "download buffers are too small")
        check(_lib.lib().fq_csr_download_async(self.ctx._h, self._h, _p(rp), _p(ci), _p(va)))
        return rp, ci, va

    def transpose(self) -> "DeviceCsr":
        h = C.c_void_p()
        check(_lib.lib().fq_csr_transpose(self.ctx._h, self._h, C.byref(h)))
        return DeviceCsr(self.ctx, h)

    def inv_diagonal(self) -> DeviceVector:
        """1 / a_ii of the held rows (Jacobi, iterative/src/precond.rs:87-121); FormoniqError on a zero / missing entry."""
        b, e = self.row_range
        d = DeviceVector(self.ctx, e - b)
        check(_lib.lib().fq_csr_inv_diagonal(self.ctx._h, self._h, d._h))
        return d

    def row_abs_sums(self) -> DeviceVector:
        """y_i = sum_j |a_ij| of the held rows (the inf-norm is its maximum)."""
        b, e = self.row_range
        y = DeviceVector(self.ctx, e - b)
        check(_lib.lib().fq_csr_row_abs_sums(self.ctx._h, self._h, y._h))
        return y

    def __add__(self, other: "DeviceCsr") -> "DeviceCsr":
        """A + B on the union pattern (nalgebra-sparse `&a + &b`), on the device."""
        h = C.c_void_p()
        check(_lib.lib().fq_csr_add(self.ctx._h, self._h, other._h, C.byref(h)))
        return DeviceCsr(self.ctx, h)

    def restrict(self, rows_keep, cols_keep) -> "DeviceCsr":
        """E_test^T A E_trial of RelativeWhitneyComplex::assemble: the sub-matrix on ascending index lists."""
        r = np.ascontiguousarray(rows_keep, dtype=np.uint64)
        c = np.ascontiguousarray(cols_keep, dtype=np.uint64)
        h = C.c_void_p()
        check(_lib.lib().fq_csr_restrict(self.ctx._h, self._h, _p(r), r.shape[0], _p(c), c.shape[0], C.byref(h)))
        return DeviceCsr(self.ctx, h)

    def to_scipy(self):
        import scipy.sparse as sp

        rp, ci, va = self.download()
        b, e = self.row_range
        return sp.csr_matrix((va, ci.astype(np.int64), rp.astype(np.int64)), shape=(e - b, self.shape[1]))

    def numeric(self, mesh: Mesh, drop_exact_zeros: bool = True):
        """Re-run the numeric phase (new geometry, same topology)."""
        check(_lib.lib().fq_assemble_numeric(self.ctx._h, mesh._h, self._h, int(drop_exact_zeros)))

    def apply(self, x: DeviceVector, out: DeviceVector | None = None) -> DeviceVector:
        b, e = self.row_range
        y = out if out is not None else DeviceVector(self.ctx, e - b)
        check(_lib.lib().fq_spmv(self.ctx._h, self._h, x._h, y._h))
        return y

    def apply_window(self, x: DeviceVector, x_lo: int, out: DeviceVector | None = None) -> DeviceVector:
        """y = A x with x holding the column window [x_lo, x_lo+len(x)) (owned + halo segment)."""
        b, e = self.row_range
        y = out if out is not None else DeviceVector(self.ctx, e - b)
        check(_lib.lib().fq_spmv_window(self.ctx._h, self._h, x._h, x_lo, y._h))
        return y

    @property
    def assembly_bytes(self) -> int:
        return _lib.lib().fq_csr_assembly_bytes(self._h)

    @property
    def assembly_shared_bytes(self) -> int:
        """Inputs (edge lengths, cell -> edge ids) that blocks assembled in one fused launch read once."""
        return _lib.lib().fq_csr_assembly_shared_bytes(self._h)

    @property
    def plan_build_ms(self) -> float:
        return _lib.lib().fq_csr_plan_build_ms(self._h)

    @property
    def plan_cell_visits(self) -> int:
        return _lib.lib().fq_csr_plan_cell_visits(self._h)

    @property
    def spmv_bytes(self) -> int:
        return _lib.lib().fq_csr_spmv_bytes(self._h)

    def __del__(self):
        try:
            if self._owner is None:
                _lib.lib().fq_csr_destroy(self._h)
        except Exception:
            pass


class BilinearForm:
    """The reference trait: grades, element matrices, assembly."""

    kind: int
    dim: int
    grade: int

    def test_grade(self) -> int:
        return self.grade - (self.kind in (FQ_DIF_TEST, FQ_DIF_BOTH))

    def trial_grade(self) -> int:
        return self.grade - (self.kind in (FQ_DIF_TRIAL, FQ_DIF_BOTH))

    def element_shape(self):
        return nlocal(self.dim, self.test_grade()), nlocal(self.dim, self.trial_grade())

    def element_batch(self, mesh: Mesh, cell_begin: int = 0, cell_end: int | None = None, use_generated: bool = True):
        """BilinearForm::element for cells [cell_begin, cell_end) -> [ncells, rows, cols]."""
        if mesh.dim != self.dim:
            raise FormoniqError(-1, "form and mesh dimensions differ")  # assert_eq!(self.dim, metric.dim())
        cell_end = mesh.ncells if cell_end is None else cell_end
        r, c = self.element_shape()
        out = np.zeros((cell_end - cell_begin, r, c))
        check(_lib.lib().fq_elmat_batch(mesh.ctx._h, mesh._h, self.kind, self.grade, cell_begin, cell_end,
                                        int(use_generated), _p(out)))
        return out

    def symbolic(self, mesh: Mesh, row_begin: int = 0, row_end: int = SIZE_MAX) -> DeviceCsr:
        h = C.c_void_p()
        check(_lib.lib().fq_assemble_symbolic(mesh.ctx._h, mesh._h, self.kind, self.grade, row_begin, row_end, C.byref(h)))
        return DeviceCsr(mesh.ctx, h)

    def assemble(self, mesh: Mesh, drop_exact_zeros: bool = True) -> DeviceCsr:
        """BilinearForm::assemble -> GalerkinMatrix (reference `!= 0.0` filter by default)."""
        if mesh.dim != self.dim:
            raise FormoniqError(-1, "form and mesh dimensions differ")
        h = C.c_void_p()
        check(_lib.lib().fq_assemble(mesh.ctx._h, mesh._h, self.kind, self.grade, int(drop_exact_zeros), C.byref(h)))
        return DeviceCsr(mesh.ctx, h)


class WhitneyPairing(BilinearForm):
    def __init__(self, dim: int, grade: int, kind: int):
        self.dim, self.grade, self.kind = int(dim), int(grade), kind

    @classmethod
    def mass(cls, dim, grade):
        return cls(dim, grade, FQ_MASS)

    @classmethod
    def dif_trial(cls, dim, grade):
        return cls(dim, grade, FQ_DIF_TRIAL)

    @classmethod
    def dif_test(cls, dim, grade):
        return cls(dim, grade, FQ_DIF_TEST)

    @classmethod
    def dif_both(cls, dim, grade):
        return cls(dim, grade, FQ_DIF_BOTH)


class ScalarLumpedMass(BilinearForm):
    def __init__(self, dim: int):
        self.dim, self.grade, self.kind = int(dim), 0, FQ_LUMPED

    def test_grade(self):
        return 0

    def trial_grade(self):
        return 0


class ElementOperator:
    """formoniq::matfree::ElementOperator: the matrix-free peer of the assembled matrix (matfree.rs:60-179)."""

    def __init__(self, mesh: Mesh, form: "BilinearForm"):
        self.ctx, self.mesh, self.form = mesh.ctx, mesh, form  # the mesh must outlive the operator
        h = C.c_void_p()
        check(_lib.lib().fq_matfree_create(mesh.ctx._h, mesh._h, form.kind, form.grade, C.byref(h)))
        self._h = h
        nr, nc = C.c_size_t(), C.c_size_t()
        check(_lib.lib().fq_matfree_shape(h, C.byref(nr), C.byref(nc)))
        self.nrows, self.ncols = nr.value, nc.value

    def refresh(self):
        """Re-evaluate the element matrices after mesh.set_lengths()."""
        check(_lib.lib().fq_matfree_refresh(self.ctx._h, self._h))

    def apply(self, x: DeviceVector, out: DeviceVector | None = None) -> DeviceVector:
        y = out if out is not None else DeviceVector(self.ctx, self.nrows)
        check(_lib.lib().fq_matfree_apply(self.ctx._h, self._h, x._h, y._h))
        return y

    def diagonal(self) -> DeviceVector:
        d = DeviceVector(self.ctx, self.nrows)
        check(_lib.lib().fq_matfree_diagonal(self.ctx._h, self._h, d._h))
        return d

    def dim(self) -> int:
        return self.nrows

    def __del__(self):
        try:
            _lib.lib().fq_matfree_destroy(self._h)
        except Exception:
            pass


class LinearFormPlan:
    """formoniq::galerkin::assemble_vector (galerkin.rs:279-312) for linear forms of one grade on one mesh: the
    load vector ell_sigma = sum_K elvec_K[position of sigma in K], cells in ascending order (`LinearForm::assemble`).
    The element vectors come from the caller (`LinearForm::element` evaluates a user field: host code)."""

    def __init__(self, mesh: Mesh, grade: int):
        self.ctx, self.mesh, self.grade = mesh.ctx, mesh, grade  # the mesh must outlive the plan
        h = C.c_void_p()
        check(_lib.lib().fq_linear_form_create(mesh.ctx._h, mesh._h, grade, C.byref(h)))
        self._h = h
        nr, nc = C.c_size_t(), C.c_size_t()
        check(_lib.lib().fq_matfree_shape(h, C.byref(nr), C.byref(nc)))
        self.nrows = nr.value

    def assemble(self, element_vectors: np.ndarray, out: DeviceVector | None = None) -> DeviceVector:
        ev = np.ascontiguousarray(element_vectors, dtype=np.float64)
        y = out if out is not None else DeviceVector(self.ctx, self.nrows)
        check(_lib.lib().fq_linear_form_assemble(self.ctx._h, self._h, ev.ctypes.data_as(C.c_void_p), y._h))
        return y

    def __del__(self):
        try:
            _lib.lib().fq_linear_form_destroy(self._h)
        except Exception:
            pass


class WeightedHodgeMass(BilinearForm):
    """formoniq::operators::WeightedHodgeMass (operators.rs:432-486): [int_K alpha <W_sigma, W_tau> vol] by quadrature,
    alpha a scalar (grade-0) coefficient sampled by the caller at the rule's nodes: `coefficient[cell][node]`.
    The default rule is the degree-1 Grundmann-Moeller rule (operators.rs:229-231)."""

    def __init__(self, dim: int, grade: int, degree: int = 1):
        from . import quadrature
        self.dim, self.grade, self.kind = int(dim), int(grade), FQ_MASS
        self.nodes, weights = quadrature.quad_rule(dim, degree)
        self.weights = np.ascontiguousarray(weights)
        self.shapes = np.ascontiguousarray(quadrature.whitney_shapes(dim, grade, self.nodes))

    def numeric(self, mesh: Mesh, a: DeviceCsr, coefficient: np.ndarray, drop_exact_zeros: bool = True) -> DeviceCsr:
        nn = self.shapes.shape[0]
        al = np.ascontiguousarray(coefficient, dtype=np.float64)
        if al.size != mesh.ncells * nn:
            raise FormoniqError(-1, "coefficient must be [ncells][nnodes]")
        check(_lib.lib().fq_weighted_mass_numeric(mesh.ctx._h, mesh._h, a._h, nn, self.weights.ctypes.data_as(C.c_void_p),
                                                  self.shapes.ctypes.data_as(C.c_void_p), al.ctypes.data_as(C.c_void_p),
                                                  int(drop_exact_zeros)))
        return a

    def assemble(self, mesh: Mesh, coefficient: np.ndarray, drop_exact_zeros: bool = True) -> DeviceCsr:  # type: ignore[override]
        """BilinearForm::assemble for this form: symbolic phase of the mass pattern + the quadrature numeric pass."""
        if mesh.dim != self.dim:
            raise FormoniqError(-1, "form and mesh dimensions differ")
        return self.numeric(mesh, self.symbolic(mesh), coefficient, drop_exact_zeros)


class SourceForm:
    """formoniq::operators::SourceForm (operators.rs:607-635): the load [int_K <f, W_sigma> vol]_sigma of a k-form source.
    `SourceForm(dim, grade, degree)` builds the reference data (`CellQuadrature::new` + `LsfSamples::whitney`; the default
    rule is the degree-1 Grundmann-Moeller rule, operators.rs:229-231); `nodes` are the barycentric quadrature nodes at
    which the caller samples its field: `samples[cell][node][component]`, components of f in the cell's reference frame
    on the colex k-subsets of the axes (what `Section::at(point)` returns).  `assemble` = `LinearForm::assemble`."""

    def __init__(self, dim: int, grade: int, degree: int = 1):
        from . import quadrature
        self.dim, self.grade = dim, grade
        self.nodes, self.weights = quadrature.quad_rule(dim, degree)
        self.shapes = np.ascontiguousarray(quadrature.whitney_shapes(dim, grade, self.nodes))
        self.weights = np.ascontiguousarray(self.weights)
        self._plans = {}

    def test_grade(self) -> int:
        return self.grade

    def assemble(self, mesh: Mesh, samples: np.ndarray, plan: "LinearFormPlan | None" = None) -> DeviceVector:
        if plan is None:
            plan = self._plans.get(id(mesh))
            if plan is None or plan.mesh is not mesh:
                plan = self._plans[id(mesh)] = LinearFormPlan(mesh, self.grade)
        nn, nd, nc = self.shapes.shape
        f = np.ascontiguousarray(samples, dtype=np.float64)
        if f.size != mesh.ncells * nn * nc:
            raise FormoniqError(-1, "samples must be [ncells][nnodes][C(dim, grade)]")
        y = DeviceVector(mesh.ctx, plan.nrows)
        check(_lib.lib().fq_source_form_assemble(mesh.ctx._h, plan._h, nn, self.weights.ctypes.data_as(C.c_void_p),
                                                 self.shapes.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p), y._h))
        return y


class _HodgePlan:
    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        try:
            _lib.lib().fq_hodge_destroy(self._h)
        except Exception:
            pass


@dataclass
class HodgeBlocks:
    """The four matrices of a mixed problem posed at `grade` (hodge.rs:62-72),
    assembled with one fused element kernel per numeric pass."""

    n_sigma: int
    n_u: int
    mass_sigma: DeviceCsr
    mass_u: DeviceCsr
    dif_test: DeviceCsr
    dif_both: DeviceCsr
    _plan: _HodgePlan = None

    @classmethod
    def symbolic(cls, mesh: Mesh, grade: int, sigma_rows=(0, SIZE_MAX), u_rows=(0, SIZE_MAX)) -> "HodgeBlocks":
        h = C.c_void_p()
        check(_lib.lib().fq_hodge_symbolic(mesh.ctx._h, mesh._h, grade, sigma_rows[0], sigma_rows[1], u_rows[0], u_rows[1],
                                           C.byref(h)))
        plan = _HodgePlan(h)
        blk = [DeviceCsr(mesh.ctx, C.c_void_p(_lib.lib().fq_hodge_block(h, i)), owner=plan) for i in range(4)]
        return cls(mesh.nsimplices(grade - 1), mesh.nsimplices(grade), blk[0], blk[1], blk[2], blk[3], plan)

    def numeric(self, mesh: Mesh, drop_exact_zeros: bool = True):
        check(_lib.lib().fq_hodge_numeric(mesh.ctx._h, mesh._h, self._plan._h, int(drop_exact_zeros)))

    @property
    def blocks(self):
        return [self.mass_sigma, self.mass_u, self.dif_test, self.dif_both]

    @classmethod
    def compute(cls, mesh: Mesh, grade: int, drop_exact_zeros: bool = True) -> "HodgeBlocks":
        if grade > mesh.dim or grade < 0:
            raise FormoniqError(-1, "grade <= complex.dim() is required")
        hb = cls.symbolic(mesh, grade)
        hb.numeric(mesh, drop_exact_zeros)
        return hb

    def mixed_hodge_laplacian(self, on_device: bool = True, symmetrized: bool = False):
        """[[M_{k-1}, -dif_test], [dif_test^T, dif_both]] (hodge.rs:93-99), stitched on the device
        (on_device=False: on the host with scipy, then uploaded — the cross-check of the tests).
        symmetrized: the sigma block-row negated, [[-M, dif_test], [dif_test^T, dif_both]] — the symmetric saddle point
        assemble_mixed_kkt gives MINRES (problems/elliptic.rs:101-113)."""
        if on_device:
            h = C.c_void_p()
            fn = _lib.lib().fq_hodge_mixed_kkt_symmetric if symmetrized else _lib.lib().fq_hodge_mixed_laplacian
            check(fn(self.mass_u.ctx._h, self._plan._h, C.byref(h)))
            return DeviceCsr(self.mass_u.ctx, h)
        import scipy.sparse as sp

        ms, dt, db = self.mass_sigma.to_scipy(), self.dif_test.to_scipy(), self.dif_both.to_scipy()
        a = sp.bmat([[ms, -dt], [dt.T, db]], format="csr")
        return DeviceCsr.from_scipy(self.mass_u.ctx, a)


def _krylov(fn, op: DeviceCsr, precond, b: DeviceVector, stop: StopCriterion):
    pc = {None: 0, "identity": 0, "jacobi": 1}[precond]
    x = b.zeros_like()
    it, res, conv = C.c_size_t(), C.c_double(), C.c_int()
    check(fn(op.ctx._h, op._h, pc, b._h, float(stop.rtol), int(stop.max_iters), x._h, C.byref(it), C.byref(res),
             C.byref(conv)))
    return x, Report(it.value, res.value, bool(conv.value))


def cg(op: DeviceCsr, precond, b: DeviceVector, stop: StopCriterion):
    """iterative::krylov::cg on device vectors; precond in {None, "jacobi"}."""
    return _krylov(_lib.lib().fq_cg, op, precond, b, stop)


def minres(op: DeviceCsr, precond, b: DeviceVector, stop: StopCriterion):
    """iterative::krylov::minres on device vectors; precond in {None, "jacobi"}."""
    return _krylov(_lib.lib().fq_minres, op, precond, b, stop)


# ---------------------------------------------------------------------------
# Krylov over user operators and the AFW block preconditioner
# ---------------------------------------------------------------------------
_APPLY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
_REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_double, C.POINTER(C.c_double))


def _wrap_apply(ctx: Context, n: int, fn):
    """fn(x: DeviceVector, y: DeviceVector) on views of the device pointers the library hands to the callback."""
    if fn is None:
        return C.cast(None, _APPLY_FN), None

    def cb(_user, xp, yp):
        try:
            fn(DeviceVector.wrap(ctx, xp, n), DeviceVector.wrap(ctx, yp, n))
            return 0
        except Exception as exc:  # never unwind through the C frames
            import traceback

            traceback.print_exc()
            return 1

    c = _APPLY_FN(cb)
    return c, c


def _krylov_op(fn, ctx: Context, n: int, apply, precond, reduce, b: DeviceVector, stop: StopCriterion):
    a_cb, a_keep = _wrap_apply(ctx, n, apply)
    p_cb, p_keep = _wrap_apply(ctx, n, precond)
    if reduce is None:
        r_cb = C.cast(None, _REDUCE_FN)
    else:
        def rcb(_user, local, out):
            try:
                out[0] = float(reduce(float(local)))
                return 0
            except Exception:
                import traceback

                traceback.print_exc()
                return 1

        r_cb = _REDUCE_FN(rcb)
    x = b.zeros_like()
    it, res, conv = C.c_size_t(), C.c_double(), C.c_int()
    check(fn(ctx._h, n, a_cb, p_cb, r_cb, None, b._h, float(stop.rtol), int(stop.max_iters), x._h, C.byref(it), C.byref(res),
             C.byref(conv)))
    del a_keep, p_keep
    return x, Report(it.value, res.value, bool(conv.value))


def cg_op(ctx: Context, n: int, apply, b: DeviceVector, stop: StopCriterion, precond=None, reduce=None):
    """iterative::krylov::cg over a user operator: apply(x, y) / precond(r, z) act on DeviceVector views, reduce(local)
    completes an inner product (all-reduce of a distributed space)."""
    return _krylov_op(_lib.lib().fq_cg_op, ctx, n, apply, precond, reduce, b, stop)


def minres_op(ctx: Context, n: int, apply, b: DeviceVector, stop: StopCriterion, precond=None, reduce=None):
    """iterative::krylov::minres over a user operator (see cg_op)."""
    return _krylov_op(_lib.lib().fq_minres_op, ctx, n, apply, precond, reduce, b, stop)


def minres_blockdiag(op: DeviceCsr, blocks, offsets, b: DeviceVector, stop: StopCriterion, inner: StopCriterion):
    """MINRES on `op` preconditioned by diag(blocks[i]^-1) on the segments [offsets[i], offsets[i+1]) (None = identity),
    each block an inner Jacobi-CG solve: with blocks (hdif_gram(k-1), hdif_gram(k)) the AFW block preconditioner of
    problems/elliptic.rs:29-47.  Returns (x, Report, total inner iterations)."""
    nb = len(blocks)
    arr = (C.c_void_p * nb)(*[None if blk is None else blk._h for blk in blocks])
    offs = (C.c_size_t * (nb + 1))(*[int(o) for o in offsets])
    x = b.zeros_like()
    it, res, conv, inner_it = C.c_size_t(), C.c_double(), C.c_int(), C.c_size_t()
    check(_lib.lib().fq_minres_blockdiag(op.ctx._h, op._h, nb, arr, offs, float(inner.rtol), int(inner.max_iters), b._h,
                                         float(stop.rtol), int(stop.max_iters), x._h, C.byref(it), C.byref(res), C.byref(conv),
                                         C.byref(inner_it)))
    return x, Report(it.value, res.value, bool(conv.value)), inner_it.value


class WhitneyComplex:
    """HilbertComplex over a device mesh (crates/formoniq/src/whitney_complex.rs:55-183): the provided assemblies of the
    trait — pairing / mass / dif_trial / dif_test / dif_both / hdif_gram — each one `assemble(form)` on the device.
    Grades off [0, dim] give correctly shaped zero matrices (whitney_complex.rs:113-122)."""

    def __init__(self, mesh: Mesh, drop_exact_zeros: bool = True):
        self.mesh, self.drop = mesh, drop_exact_zeros

    @property
    def dim(self) -> int:
        return self.mesh.dim

    def assemble(self, form: "BilinearForm") -> DeviceCsr:
        return form.assemble(self.mesh, self.drop)

    pairing = assemble

    def mass(self, grade: int) -> DeviceCsr:
        return self.assemble(WhitneyPairing.mass(self.dim, grade))

    def dif_trial(self, grade: int) -> DeviceCsr:
        return self.assemble(WhitneyPairing.dif_trial(self.dim, grade))

    def dif_test(self, grade: int) -> DeviceCsr:
        return self.assemble(WhitneyPairing.dif_test(self.dim, grade))

    def dif_both(self, grade: int) -> DeviceCsr:
        return self.assemble(WhitneyPairing.dif_both(self.dim, grade))

    def hdif_gram(self, grade: int) -> DeviceCsr:
        """Gram matrix of the graph inner product <u,v> + <du,dv> (whitney_complex.rs:180-183): mass(k) + dif_both(k+1)."""
        return self.mass(grade) + self.dif_both(grade + 1)
