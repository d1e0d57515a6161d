"""Block shift-invert Lanczos on device vectors: `sparse_shift_invert_eigen`
(formoniq/src/linalg/eigen.rs:60-176) — the solver behind `elliptic::solve_evp`
(BASELINE config 5).

Division of labour, as in the reference: the sparse factorisation of
M = A - shift*B and its triangular solves are third-party host code (faer's
sparse LU there, SuperLU through scipy here — SURVEY §7 H6); everything else
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


def pseudo_random(seed: int, n: int) -> np.ndarray:
    """eigen.rs:259-268, vectorised over index = 0..n-1."""
    idx = np.arange(n, dtype=np.uint64)
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


def _b_orthogonalize(v: DeviceVector, basis, bbasis):
    """eigen.rs:245-255: two passes of modified Gram-Schmidt against a B-orthonormal basis; returns the coefficients."""
    coeffs = [0.0] * len(basis)
    for _ in range(2):
        for j, (vj, bvj) in enumerate(zip(basis, bbasis)):
            c = v.dot(bvj)
            coeffs[j] += c
            v.add_scaled(-c, vj)
    return coeffs


def _combine(vecs, coeff, dim: int) -> DeviceVector:  # eigen.rs:270-276
    out = vecs[0].zeros_like()
    for l in range(dim):
        out.add_scaled(float(coeff[l]), vecs[l])
    return out


def _residual(a: DeviceCsr, b: DeviceCsr, lam: float, x: DeviceVector, a_norm: float, b_norm: float) -> float:  # eigen.rs:280-296
    r = a.apply(x)
    r.add_scaled(-lam, b.apply(x))
    xnorm = x.norm()
    scale = a_norm * xnorm + abs(lam) * b_norm * xnorm
    return r.norm() / scale if scale > 0.0 else r.norm()


def sparse_shift_invert_eigen(a: DeviceCsr, b: DeviceCsr, shift: float, k: int):
    """The `k` eigenpairs of A x = lambda B x closest to `shift` (A symmetric, B symmetric positive semi-definite).

    Returns (eigenvalues ascending, list of B-normalised eigenvectors as DeviceVector).  Raises EigenError like the
    reference's Result (SingularPencil / NoFiniteEigenvalue / NotConverged)."""
    n = a.shape[0]
    if a.shape != (n, n) or b.shape != (n, n):
        raise ValueError("A and B must be square and of one size")
    ctx = a.ctx
    k = min(k, n)
    if k == 0:
        return np.zeros(0), []
    ah, bh = a.to_scipy(), b.to_scipy()          # host copies for the third-party factorisation
    a_norm, b_norm = _inf_norm(ah), _inf_norm(bh)
    lu, used_shift = _factor_with_retry(ah, bh, shift, a_norm)
    target_dim = min(max(4 * k, 2 * k + 20), n)

    def solve(v: DeviceVector) -> DeviceVector:   # M^-1 v: host triangular solves, vectors cross PCIe once each way
        return DeviceVector.from_numpy(ctx, lu.solve(v.to_numpy()))

    # seed_block (eigen.rs:213-240)
    basis, bbasis = [], []
    seed = 0
    while len(basis) < k and seed < k * 32 + 32:
        v = DeviceVector.from_numpy(ctx, pseudo_random(seed, n))
        seed += 1
        _b_orthogonalize(v, basis, bbasis)
        bv = b.apply(v)
        norm_sq = v.dot(bv)
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
            w = solve(bbasis[dim])
            h = _b_orthogonalize(w, basis, bbasis)
            for j, hj in enumerate(h):
                proj[j, dim] = hj
                proj[dim, j] = hj
            bw = b.apply(w)
            beta_sq = w.dot(bw)
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
            res = _residual(a, b, lam, y, a_norm, b_norm)
            if res > RESIDUAL_TOL:   # refine only when the raw pair misses (eigen.rs:121-136)
                x = solve(b.apply(y))
                bnorm = np.sqrt(max(x.dot(b.apply(x)), 0.0))
                if bnorm > 0.0:
                    x.scale(1.0 / bnorm)
                y, res = x, _residual(a, b, lam, x, a_norm, b_norm)
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
