"""Block shift-invert Lanczos on device vectors: `sparse_shift_invert_eigen`
(formoniq/src/linalg/eigen.rs:60-176) — the solver behind `elliptic::solve_evp`
(BASELINE config 5).

Written over an abstract pencil (operator applications, global inner products,
seeds, inner solve), so the same driver runs on one GPU (CsrPencil) and on a
row-partitioned KKT pencil across GPUs (dist.DistKktPencil).  The inner solve
M^-1 v, M = A - shift*B, is either the reference's division of labour — sparse
factorisation and triangular solves as third-party host code (faer's sparse LU
there, SuperLU through scipy here) — or MINRES on the shifted
operator, SpMV-only and on the device; everything else
— the B-products, the two-pass B-orthogonalisation against the whole basis
(O(dim) dots/axpys per step), the Ritz combinations and the backward-error
residuals — runs on the device through the library's SpMV and BLAS-1 kernels.
Step for step this follows eigen.rs: same seeds (splitmix64 `pseudo_random`),
same Krylov cap `max(4k, 2k+20)`, same restart and refinement rules, same
tolerances."""
from __future__ import annotations

import numpy as np

from .api import DeviceCsr, DeviceVector

RESIDUAL_TOL = 1e-9          # eigen.rs:96
MAX_RESTART_CYCLES = 100     # eigen.rs:97
BREAKDOWN_TOL_SQ = 1e-20     # eigen.rs:201
SEED_TOL = 1e-24             # eigen.rs:226
MASK = (1 << 64) - 1


class EigenError(RuntimeError):
    """eigen.rs:12-21: SingularPencil, NoFiniteEigenvalue, NotConverged."""

    def __init__(self, kind: str, **info):
        super().__init__(f"{kind}: {info}")
        self.kind, self.info = kind, info


def pseudo_random(seed: int, n, start: int = 0) -> np.ndarray:
    """eigen.rs:259-268, vectorised over index = start..start+n-1 (a rank of a distributed run takes its slice of the
    one global seed vector)."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (np.uint64((seed * 0x9E3779B97F4A7C15) & MASK) + idx * np.uint64(0xD1B54A32D192ED03) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2.0 - 1.0


def _inf_norm(m) -> float:
    return float(abs(m).sum(axis=1).max()) if m.nnz else 0.0


def _perturbed_shift(shift: float, eps0: float, attempt: int) -> float:  # eigen.rs:314-324
    if attempt == 0:
        return shift
    step = eps0 * 2.0 ** ((attempt - 1) // 2)
    return shift + step if attempt % 2 == 1 else shift - step


def _factor_with_retry(a, b, shift: float, a_norm: float):
    """eigen.rs:327-357: sparse LU of A - shift*B, a perturbed shift on a singular or ill-conditioned factorisation."""
    import scipy.sparse.linalg as spla

    eps0 = np.sqrt(np.finfo(np.float64).eps) * max(a_norm, 1.0)
    n = a.shape[0]
    for attempt in range(16):
        cur = _perturbed_shift(shift, eps0, attempt)
        m = (a - cur * b).tocsc()
        try:
            lu = spla.splu(m)
        except RuntimeError:
            continue
        probe = pseudo_random(MASK, n)
        resolved = lu.solve(m @ probe)
        if np.all(np.isfinite(resolved)) and np.linalg.norm(resolved - probe) <= 1e-6 * max(np.linalg.norm(probe), 1.0):
            return lu, cur
    raise EigenError("SingularPencil", shift=shift)


class CsrPencil:
    """(A, B) as two assembled device matrices on one GPU.  inner = "lu": the reference's division of labour (sparse LU of
    A - shift*B on the host, SuperLU standing in for faer); inner = "minres": the inner solves are MINRES on the
    shifted operator x -> A x - shift * B x, i.e. SpMV-only and on the device."""

    def __init__(self, a: DeviceCsr, b: DeviceCsr, inner: str = "lu", inner_rtol: float = 1e-13, inner_max_iters: int = 200000,
                 negate_rows: int = 0):
        """negate_rows: for inner = "minres", the leading rows of A (and of the right-hand side) are negated inside the
        inner solve — the mixed Hodge-Laplacian [[M, -D], [D^T, K]] becomes the symmetric [[-M, D], [D^T, K]] MINRES needs
        (problems/elliptic.rs:101-113); B must vanish on those rows (it does: B = diag(0, M_k))."""
        self.negate_rows = int(negate_rows)
        n = a.shape[0]
        if a.shape != (n, n) or b.shape != (n, n):
            raise ValueError("A and B must be square and of one size")
        self.a, self.b, self.n, self.n_global, self.ctx = a, b, n, n, a.ctx
        self.inner, self.inner_rtol, self.inner_max_iters = inner, inner_rtol, inner_max_iters
        self.inner_iterations = 0
        if inner == "lu":
            self._ah, self._bh = a.to_scipy(), b.to_scipy()      # host copies for the third-party factorisation
            self.a_norm, self.b_norm = _inf_norm(self._ah), _inf_norm(self._bh)
        else:
            self.a_norm = float(a.row_abs_sums().to_numpy().max()) if n else 0.0
            self.b_norm = float(b.row_abs_sums().to_numpy().max()) if n else 0.0

    def a_apply(self, x):
        return self.a.apply(x)

    def b_apply(self, x):
        return self.b.apply(x)

    def dot(self, x, y) -> float:
        return x.dot(y)

    def seed(self, s: int):
        return DeviceVector.from_numpy(self.ctx, pseudo_random(s, self.n))

    def prepare(self, shift: float) -> float:
        if self.inner == "lu":
            self._lu, used = _factor_with_retry(self._ah, self._bh, shift, self.a_norm)
            return used
        self._shift = shift
        return shift

    def solve(self, v):
        if self.inner == "lu":   # M^-1 v: host triangular solves, vectors cross PCIe once each way
            return DeviceVector.from_numpy(self.ctx, self._lu.solve(v.to_numpy()))
        from .api import StopCriterion, minres_op

        tmp = v.zeros_like()
        m = self.negate_rows

        def shifted(x, y):
            self.a.apply(x, y)
            if m:
                y.view(0, m).scale(-1.0)
            self.b.apply(x, tmp)
            y.add_scaled(-self._shift, tmp)

        rhs = v
        if m:
            rhs = v.clone()
            rhs.view(0, m).scale(-1.0)
        x, rep = minres_op(self.ctx, self.n, shifted, rhs, StopCriterion(self.inner_rtol, self.inner_max_iters))
        self.inner_iterations += rep.iters
        if not rep.converged:
            raise EigenError("SingularPencil", shift=self._shift, inner_residual=rep.residual)
        return x


def _b_orthogonalize(pencil, v: DeviceVector, basis, bbasis):
    """eigen.rs:245-255: two passes of modified Gram-Schmidt against a B-orthonormal basis; returns the coefficients."""
    coeffs = [0.0] * len(basis)
    for _ in range(2):
        for j, (vj, bvj) in enumerate(zip(basis, bbasis)):
            c = pencil.dot(v, bvj)
            coeffs[j] += c
            v.add_scaled(-c, vj)
    return coeffs


def _combine(vecs, coeff, dim: int) -> DeviceVector:  # eigen.rs:270-276
    out = vecs[0].zeros_like()
    for l in range(dim):
        out.add_scaled(float(coeff[l]), vecs[l])
    return out


def _residual(pencil, lam: float, x: DeviceVector) -> float:  # eigen.rs:280-296
    r = pencil.a_apply(x)
    r.add_scaled(-lam, pencil.b_apply(x))
    xnorm = np.sqrt(max(pencil.dot(x, x), 0.0))
    rnorm = np.sqrt(max(pencil.dot(r, r), 0.0))
    scale = pencil.a_norm * xnorm + abs(lam) * pencil.b_norm * xnorm
    return rnorm / scale if scale > 0.0 else rnorm


def shift_invert_lanczos(pencil, shift: float, k: int):
    """Block shift-invert Lanczos over an abstract pencil (CsrPencil here, dist.DistKktPencil across GPUs): the `k`
    eigenpairs of A x = lambda B x closest to `shift`.  Step for step eigen.rs:60-176."""
    n = pencil.n_global
    k = min(k, n)
    if k == 0:
        return np.zeros(0), []
    used_shift = pencil.prepare(shift)
    target_dim = min(max(4 * k, 2 * k + 20), n)
    # seed_block (eigen.rs:213-240)
    basis, bbasis = [], []
    seed = 0
    while len(basis) < k and seed < k * 32 + 32:
        v = pencil.seed(seed)
        seed += 1
        _b_orthogonalize(pencil, v, basis, bbasis)
        bv = pencil.b_apply(v)
        norm_sq = pencil.dot(v, bv)
        if norm_sq > SEED_TOL:
            norm = np.sqrt(norm_sq)
            v.scale(1.0 / norm)
            bv.scale(1.0 / norm)
            basis.append(v)
            bbasis.append(bv)
    if not basis:
        raise EigenError("NoFiniteEigenvalue")
    proj_cap = min(target_dim + 2 * k, n)
    proj = np.zeros((proj_cap, proj_cap))
    dim = 0
    worst = np.inf
    for _cycle in range(MAX_RESTART_CYCLES + 1):
        while dim < target_dim and dim < len(basis):
            # expand (eigen.rs:183-208)
            w = pencil.solve(bbasis[dim])
            h = _b_orthogonalize(pencil, w, basis, bbasis)
            for j, hj in enumerate(h):
                proj[j, dim] = hj
                proj[dim, j] = hj
            bw = pencil.b_apply(w)
            beta_sq = pencil.dot(w, bw)
            if beta_sq > BREAKDOWN_TOL_SQ:
                beta = np.sqrt(beta_sq)
                w.scale(1.0 / beta)
                bw.scale(1.0 / beta)
                basis.append(w)
                bbasis.append(bw)
            dim += 1
        exhausted = dim < target_dim
        theta, s = np.linalg.eigh(proj[:dim, :dim])
        order = sorted(range(dim), key=lambda i: -abs(theta[i]))
        take = min(k, dim)
        pairs = []
        for idx in order[:take]:
            lam = used_shift + 1.0 / theta[idx]
            y = _combine(basis, s[:, idx], dim)
            res = _residual(pencil, lam, y)
            if res > RESIDUAL_TOL:   # refine only when the raw pair misses (eigen.rs:121-136)
                x = pencil.solve(pencil.b_apply(y))
                bnorm = np.sqrt(max(pencil.dot(x, pencil.b_apply(x)), 0.0))
                if bnorm > 0.0:
                    x.scale(1.0 / bnorm)
                y, res = x, _residual(pencil, lam, x)
            pairs.append((lam, y, res))
        worst = max(p[2] for p in pairs)
        if worst <= RESIDUAL_TOL or exhausted:
            pairs.sort(key=lambda p: p[0])
            return np.array([p[0] for p in pairs]), [p[1] for p in pairs]
        keep = max(min(2 * k, max(target_dim - 1, 0)), 1)
        new_basis = [_combine(basis, s[:, idx], dim) for idx in order[:keep]]
        new_bbasis = [_combine(bbasis, s[:, idx], dim) for idx in order[:keep]]
        basis, bbasis = new_basis, new_bbasis
        proj = np.zeros((proj_cap, proj_cap))
        dim = 0
    raise EigenError("NotConverged", shift=shift, residual=worst)


def sparse_shift_invert_eigen(a: DeviceCsr, b: DeviceCsr, shift: float, k: int, inner: str = "lu", negate_rows: int = 0):
    """The `k` eigenpairs of A x = lambda B x closest to `shift` (A symmetric, B symmetric positive semi-definite).

    Returns (eigenvalues ascending, list of B-normalised eigenvectors as DeviceVector).  Raises EigenError like the
    reference's Result (SingularPencil / NoFiniteEigenvalue / NotConverged).  inner = "lu" factorises A - shift*B on the
    host like the reference; inner = "minres" keeps the inner solves on the device (SpMV-based MINRES)."""
    if a.shape[0] == 0 or k == 0:
        return np.zeros(0), []
    return shift_invert_lanczos(CsrPencil(a, b, inner, negate_rows=negate_rows), shift, k)
