"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(formoniq_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

MASS, DIF_TRIAL, DIF_TEST, DIF_BOTH, LUMPED = 0, 1, 2, 3, 4


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "fq_oracle.hpp")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i64, dbl, vp, i32 = C.c_int64, C.c_double, C.c_void_p, C.c_int
        P = C.POINTER
        L.fqo_last_error.restype = C.c_char_p
        L.fqo_max_threads.restype = i32
        L.fqo_binomial.restype = i64
        L.fqo_combinations.restype = i64
        L.fqo_combinations.argtypes = [i32, i32, vp]
        L.fqo_permutations.restype = i64
        L.fqo_permutations.argtypes = [i32, vp, vp]
        L.fqo_unit_boundary_operator.restype = i64
        L.fqo_unit_boundary_operator.argtypes = [i32, i32, vp, P(i32), P(i32)]
        L.fqo_difbarys_power.restype = i64
        L.fqo_difbarys_power.argtypes = [i32, i32, vp, P(i32), P(i32)]
        L.fqo_pseudo_random.restype = dbl
        L.fqo_pseudo_random.argtypes = [C.c_uint64, C.c_uint64]
        L.fqo_unit_simplex_lengths_sq.argtypes = [i32, vp]
        L.fqo_elmat.restype = i32
        L.fqo_elmat.argtypes = [i32, i32, i32, vp, vp, P(i32), P(i32)]
        L.fqo_cell_geometry.restype = i32
        L.fqo_cell_geometry.argtypes = [i32, vp, vp, vp, P(dbl)]
        L.fqo_complex_from_cells.restype = vp
        L.fqo_complex_from_cells.argtypes = [i32, i64, vp]
        L.fqo_complex_kuhn.restype = vp
        L.fqo_complex_kuhn.argtypes = [i32, vp]
        L.fqo_complex_destroy.argtypes = [vp]
        L.fqo_complex_dim.restype = i32
        L.fqo_complex_dim.argtypes = [vp]
        L.fqo_complex_nsimplices.restype = i64
        L.fqo_complex_nsimplices.argtypes = [vp, i32]
        L.fqo_complex_skeleton.argtypes = [vp, i32, vp]
        L.fqo_complex_cell_faces.argtypes = [vp, i32, vp]
        L.fqo_kuhn_vertex_coords.argtypes = [i32, vp, vp, vp, vp]
        L.fqo_edge_lengths_sq.argtypes = [vp, i32, vp, vp, vp]
        L.fqo_elmat_batch.restype = i32
        L.fqo_elmat_batch.argtypes = [vp, vp, i32, i32, i64, i64, vp]
        L.fqo_assemble.restype = vp
        L.fqo_assemble.argtypes = [vp, vp, i32, i32, i32, i32, vp]
        L.fqo_csr_from_arrays.restype = vp
        L.fqo_csr_from_arrays.argtypes = [i64, i64, vp, vp, vp]
        L.fqo_csr_destroy.argtypes = [vp]
        L.fqo_csr_shape.argtypes = [vp, P(i64), P(i64), P(i64)]
        L.fqo_csr_copy.argtypes = [vp, vp, vp, vp]
        L.fqo_spmv.argtypes = [vp, vp, vp]
        L.fqo_spmv_timed.restype = dbl
        L.fqo_spmv_timed.argtypes = [vp, vp, vp, i32]
        L.fqo_spmv_parallel_timed.restype = dbl
        L.fqo_spmv_parallel_timed.argtypes = [vp, vp, vp, i32, i32]
        for f in (L.fqo_cg, L.fqo_minres):
            f.restype = i32
            f.argtypes = [vp, i32, vp, dbl, i64, vp, P(i64), P(dbl)]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    return lib().fqo_max_threads()


def binomial(n, k):
    return lib().fqo_binomial(n, k)


def nlocal(n, j):
    return 0 if j < 0 or j > n else binomial(n + 1, j + 1)


def kind_grades(kind, k):
    if kind == LUMPED:
        return 0, 0
    return k - (kind in (DIF_TEST, DIF_BOTH)), k - (kind in (DIF_TRIAL, DIF_BOTH))


def combinations(n, card):
    cnt = lib().fqo_combinations(n, card, None)
    out = np.zeros((cnt, card), dtype=np.int32)
    lib().fqo_combinations(n, card, _p(out))
    return out


def permutations(n):
    cnt = lib().fqo_permutations(n, None, None)
    out = np.zeros((cnt, n), dtype=np.int32)
    sg = np.zeros(cnt)
    lib().fqo_permutations(n, _p(out), _p(sg))
    return out, sg


def unit_boundary_operator(n, k):
    r, c = C.c_int(), C.c_int()
    lib().fqo_unit_boundary_operator(n, k, None, C.byref(r), C.byref(c))
    out = np.zeros((r.value, c.value))
    lib().fqo_unit_boundary_operator(n, k, _p(out), C.byref(r), C.byref(c))
    return out


def difbarys_power(n, k):
    r, c = C.c_int(), C.c_int()
    lib().fqo_difbarys_power(n, k, None, C.byref(r), C.byref(c))
    out = np.zeros((r.value, c.value))
    lib().fqo_difbarys_power(n, k, _p(out), C.byref(r), C.byref(c))
    return out


def pseudo_random(seed, index):
    return lib().fqo_pseudo_random(seed, index)


def unit_simplex_lengths_sq(n):
    out = np.zeros(binomial(n + 1, 2))
    lib().fqo_unit_simplex_lengths_sq(n, _p(out))
    return out


def elmat(kind, n, k, lengths_sq):
    lengths_sq = np.ascontiguousarray(lengths_sq, dtype=np.float64)
    r, c = C.c_int(), C.c_int()
    buf = np.zeros(4096)
    rc = lib().fqo_elmat(kind, n, k, _p(lengths_sq), _p(buf), C.byref(r), C.byref(c))
    if rc != 0:
        raise RuntimeError(lib().fqo_last_error().decode())
    return buf[: r.value * c.value].reshape(r.value, c.value).copy()


def cell_geometry(n, lengths_sq):
    lengths_sq = np.ascontiguousarray(lengths_sq, dtype=np.float64)
    g, gi, vol = np.zeros((n, n)), np.zeros((n, n)), C.c_double()
    rc = lib().fqo_cell_geometry(n, _p(lengths_sq), _p(g), _p(gi), C.byref(vol))
    if rc != 0:
        raise RuntimeError("degenerate metric")
    return g, gi, vol.value


class Csr:
    def __init__(self, handle):
        self._h = handle
        nr, nc, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        lib().fqo_csr_shape(handle, C.byref(nr), C.byref(nc), C.byref(nnz))
        self.nrows, self.ncols, self.nnz = nr.value, nc.value, nnz.value
        self._arrays = None

    @classmethod
    def from_arrays(cls, nrows, ncols, row_ptr, col_idx, values):
        rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
        ci = np.ascontiguousarray(col_idx, dtype=np.int64)
        va = np.ascontiguousarray(values, dtype=np.float64)
        return cls(lib().fqo_csr_from_arrays(nrows, ncols, _p(rp), _p(ci), _p(va)))

    def arrays(self):
        if self._arrays is None:
            rp = np.zeros(self.nrows + 1, dtype=np.int64)
            ci = np.zeros(self.nnz, dtype=np.int64)
            va = np.zeros(self.nnz)
            lib().fqo_csr_copy(self._h, _p(rp), _p(ci), _p(va))
            self._arrays = (rp, ci, va)
        return self._arrays

    def to_scipy(self):
        import scipy.sparse as sp

        rp, ci, va = self.arrays()
        return sp.csr_matrix((va, ci, rp), shape=(self.nrows, self.ncols))

    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.nrows)
        lib().fqo_spmv(self._h, _p(x), _p(y))
        return y

    def spmv_timed(self, x, reps=3, nthreads=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.nrows)
        if nthreads:
            return lib().fqo_spmv_parallel_timed(self._h, _p(x), _p(y), reps, nthreads)
        return lib().fqo_spmv_timed(self._h, _p(x), _p(y), reps)

    def _solve(self, fn, b, rtol, max_iters, precond):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.nrows)
        it, res = C.c_int64(), C.c_double()
        conv = fn(self._h, precond, _p(b), rtol, max_iters, _p(x), C.byref(it), C.byref(res))
        return x, dict(iters=it.value, residual=res.value, converged=bool(conv))

    def cg(self, b, rtol=1e-10, max_iters=10000, precond=0):
        return self._solve(lib().fqo_cg, b, rtol, max_iters, precond)

    def minres(self, b, rtol=1e-10, max_iters=10000, precond=0):
        return self._solve(lib().fqo_minres, b, rtol, max_iters, precond)

    def __del__(self):
        try:
            lib().fqo_csr_destroy(self._h)
        except Exception:
            pass


class Complex:
    """Simplicial complex with colex skeleton numbering (reference `Complex`)."""

    def __init__(self, handle):
        self._h = handle
        self.dim = lib().fqo_complex_dim(handle)

    @classmethod
    def kuhn(cls, dim, shape):
        if np.isscalar(shape):
            shape = [int(shape)] * dim
        shape = np.ascontiguousarray(shape, dtype=np.int64)
        cx = cls(lib().fqo_complex_kuhn(dim, _p(shape)))
        cx.shape = shape
        return cx

    @classmethod
    def from_cells(cls, dim, cells):
        cells = np.ascontiguousarray(cells, dtype=np.int64).reshape(-1, dim + 1)
        return cls(lib().fqo_complex_from_cells(dim, cells.shape[0], _p(cells)))

    def nsimplices(self, j):
        return lib().fqo_complex_nsimplices(self._h, j)

    @property
    def ncells(self):
        return self.nsimplices(self.dim)

    def skeleton(self, j):
        out = np.zeros((self.nsimplices(j), j + 1), dtype=np.int64)
        lib().fqo_complex_skeleton(self._h, j, _p(out))
        return out

    def cell_faces(self, j):
        out = np.zeros((self.ncells, nlocal(self.dim, j)), dtype=np.int64)
        lib().fqo_complex_cell_faces(self._h, j, _p(out))
        return out

    def edge_lengths_sq(self, coords, ambient_diag=None):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        ad = coords.shape[1]
        diag = np.ones(ad) if ambient_diag is None else np.ascontiguousarray(ambient_diag, dtype=np.float64)
        out = np.zeros(self.nsimplices(1))
        lib().fqo_edge_lengths_sq(self._h, ad, _p(coords), _p(diag), _p(out))
        return out

    def elmat_batch(self, lengths_sq, kind, k, c0=0, c1=None):
        c1 = self.ncells if c1 is None else c1
        tg, rg = kind_grades(kind, k)
        r, c = nlocal(self.dim, tg), nlocal(self.dim, rg)
        lengths_sq = np.ascontiguousarray(lengths_sq, dtype=np.float64)
        out = np.zeros((c1 - c0, r, c))
        rc = lib().fqo_elmat_batch(self._h, _p(lengths_sq), kind, k, c0, c1, _p(out))
        if rc != 0:
            raise RuntimeError(lib().fqo_last_error().decode())
        return out

    def assemble(self, lengths_sq, kind, k, drop_zeros=True, nthreads=1, times=None):
        lengths_sq = np.ascontiguousarray(lengths_sq, dtype=np.float64)
        t = np.zeros(2)
        h = lib().fqo_assemble(self._h, _p(lengths_sq), kind, k, int(drop_zeros), nthreads, _p(t))
        if not h:
            raise RuntimeError(lib().fqo_last_error().decode())
        if times is not None:
            times[:] = t
        return Csr(h)

    def __del__(self):
        try:
            lib().fqo_complex_destroy(self._h)
        except Exception:
            pass


def kuhn_vertex_coords(dim, shape, vmin=None, vmax=None):
    if np.isscalar(shape):
        shape = [int(shape)] * dim
    shape = np.ascontiguousarray(shape, dtype=np.int64)
    vmin = np.zeros(dim) if vmin is None else np.ascontiguousarray(vmin, dtype=np.float64)
    vmax = np.ones(dim) if vmax is None else np.ascontiguousarray(vmax, dtype=np.float64)
    nv = int(np.prod(shape + 1))
    out = np.zeros((nv, dim))
    lib().fqo_kuhn_vertex_coords(dim, _p(shape), _p(vmin), _p(vmax), _p(out))
    return out


def jitter_coords(coords, shape, amplitude=0.2):
    """Generic-geometry variant: every vertex displaced by
    amplitude*h*pseudo_random(seed=axis, index=vertex)."""
    coords = coords.copy()
    dim = coords.shape[1]
    shape = np.broadcast_to(np.asarray(shape), (dim,))
    for a in range(dim):
        h = (coords[:, a].max() - coords[:, a].min()) / shape[a]
        for v in range(coords.shape[0]):
            coords[v, a] += amplitude * h * pseudo_random(a, v)
    return coords


def assemble_vector(cx: "Complex", grade: int, element_vectors: np.ndarray) -> np.ndarray:
    """Restatement of formoniq::galerkin::assemble_vector (formoniq/src/galerkin.rs:279-312): cells in order, the
    non-zero entries of each element vector as (global face, value) pairs, then `galvec[irow] += val` sequentially."""
    faces = cx.cell_faces(grade)                      # [ncells][C(n+1, grade+1)], SimplexRef::faces order
    ev = np.asarray(element_vectors, dtype=np.float64).reshape(faces.shape)
    out = np.zeros(cx.nsimplices(grade))
    rows = faces.reshape(-1)
    vals = ev.reshape(-1)
    keep = vals != 0.0                                # galerkin.rs:299
    # np.add.at applies the additions one by one in index order == the reference's sequential loop
    np.add.at(out, rows[keep], vals[keep])
    return out


def source_element_vectors(cx: "Complex", lengths_sq, grade: int, weights, shapes, samples) -> np.ndarray:
    """Restatement of SourceForm::element (formoniq/src/operators.rs:624-634) for every cell:
    CellQuadrature::integrate (operators.rs:247-261: elvec[i] += w_q * f(point, W_i(q)) node-outer, then vol * elvec) of
    inner(source, whitney, metric) = source . (Lambda^k g^-1 whitney) (metric/src/tensor.rs:127-131,140-157), with
    Lambda^k g^-1 = the k x k minors of g^-1 on colex k-subsets (multialgebra/src/lib.rs:285-305) and
    vol = cell_volume (regge/src/lib.rs:26-28).  Returns [ncells][C(n+1, grade+1)]."""
    import itertools
    n = cx.dim
    edges = cx.cell_faces(1)
    nn, nd, nc = np.asarray(shapes).shape
    f = np.asarray(samples, dtype=np.float64).reshape(cx.ncells, nn, nc)
    subsets = sorted(itertools.combinations(range(n), grade), key=lambda c: c[::-1])
    out = np.zeros((cx.ncells, nd))
    for c in range(cx.ncells):
        _, gi, vol = cell_geometry(n, np.asarray(lengths_sq)[edges[c]])
        G = np.array([[np.linalg.det(gi[np.ix_(I, J)]) if grade else 1.0 for J in subsets] for I in subsets])
        elvec = np.zeros(nd)
        for q in range(nn):
            for i in range(nd):
                elvec[i] += weights[q] * float(f[c, q] @ (G @ shapes[q][i]))
        out[c] = vol * elvec
    return out


def weighted_mass_elmats(cx: "Complex", lengths_sq, grade: int, weights, shapes, coefficient) -> np.ndarray:
    """Restatement of WeightedHodgeMass::element (formoniq/src/operators.rs:477-485) for every cell:
    CellQuadrature::integrate_pair (operators.rs:266-290: elmat[i][j] += w_q * alpha_q * inner(W_i, W_j) node-outer, then
    vol * elmat), inner as in source_element_vectors.  Returns [ncells][nd][nd]."""
    import itertools
    n = cx.dim
    edges = cx.cell_faces(1)
    nn, nd, nc = np.asarray(shapes).shape
    al = np.asarray(coefficient, dtype=np.float64).reshape(cx.ncells, nn)
    subsets = sorted(itertools.combinations(range(n), grade), key=lambda c: c[::-1])
    out = np.zeros((cx.ncells, nd, nd))
    for c in range(cx.ncells):
        _, gi, vol = cell_geometry(n, np.asarray(lengths_sq)[edges[c]])
        G = np.array([[np.linalg.det(gi[np.ix_(I, J)]) if grade else 1.0 for J in subsets] for I in subsets])
        elmat = np.zeros((nd, nd))
        for q in range(nn):
            W = np.asarray(shapes[q])                     # [dof][component]
            elmat += weights[q] * (al[c, q] * (W @ (G @ W.T)))
        out[c] = vol * elmat
    return out


def assemble_from_elmats(cx: "Complex", test_grade: int, trial_grade: int, elmats: np.ndarray, drop_zeros: bool = True):
    """assemble_matrix (formoniq/src/galerkin.rs:138-188) on given element matrices: triplets in cell order, `!= 0.0`
    filter, duplicates summed in order, CSR with ascending columns.  Returns scipy CSR (sorted, no explicit pattern loss)."""
    import scipy.sparse as sp
    rows_f, cols_f = cx.cell_faces(test_grade), cx.cell_faces(trial_grade)
    nr, nc = rows_f.shape[1], cols_f.shape[1]
    r = np.repeat(rows_f, nc, axis=1).reshape(-1)
    c = np.tile(cols_f, (1, nr)).reshape(-1)
    v = np.asarray(elmats, dtype=np.float64).reshape(-1)
    if drop_zeros:
        keep = v != 0.0
        r, c, v = r[keep], c[keep], v[keep]
    # stable sort by (row, col): duplicates stay in cell order and are summed sequentially
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    key = r.astype(np.int64) * cx.nsimplices(trial_grade) + c
    heads = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
    vals = np.zeros(len(heads))
    seg = np.repeat(np.arange(len(heads)), np.diff(np.r_[heads, len(key)]))
    np.add.at(vals, seg, v)
    indptr = np.zeros(cx.nsimplices(test_grade) + 1, dtype=np.int64)
    np.add.at(indptr, r[heads] + 1, 1)
    return sp.csr_matrix((vals, c[heads], np.cumsum(indptr)), shape=(cx.nsimplices(test_grade), cx.nsimplices(trial_grade)))
