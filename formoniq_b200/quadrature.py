"""Reference data of the quadrature forms, built once per (dim, grade, rule) on the host — as in the reference.

* `grundmann_moeller(dim, s)`: SimplexQuadRule::grundmann_moeller (simplicial/src/atlas/quadrature.rs:39-76), the
  symmetric rule exact for polynomials of degree 2s+1 on the n-simplex, barycentric nodes, weights normalised to 1.
* `whitney_shapes(dim, grade, nodes)`: LsfSamples::whitney (derham/src/interpolate/samples.rs:31-40): the Whitney shape
  functions of one grade at the nodes, in the reference frame of the cell (vertex 0 at the origin, vertex i at e_i), as
  components on the colex k-subsets of the n axes; DOF order = colex (k+1)-subsets of the vertices
  (WhitneyLsf::at_bary, derham/src/interpolate/form.rs:108-112, by the deletion formula of form.rs:131-135:
  W_sigma = k! sum_i (-1)^i lambda_{sigma_i} dlambda_{sigma_0} ^ .. omit i .. ^ dlambda_{sigma_k}).
The per-cell work (metric, inverse, minors, quadrature sum, scatter) is the device's: fq_source_form_assemble.
"""
from __future__ import annotations

import itertools
import math

import numpy as np


def _colex_subsets(n: int, card: int):
    return sorted(itertools.combinations(range(n), card), key=lambda c: c[::-1])


def _compositions(nparts: int, degree: int):
    """Composition::all (multiindex/src/composition.rs:115-119): multisets of size `degree` over `nparts` symbols as
    non-decreasing words in colex order, read as parts[p] = multiplicity of p."""
    words = sorted(itertools.combinations_with_replacement(range(nparts), degree), key=lambda w: w[::-1])
    for w in words:
        parts = [0] * nparts
        for p in w:
            parts[p] += 1
        yield parts


def grundmann_moeller(dim: int, s: int):
    """(nodes [npoints][dim+1] barycentric, weights [npoints] summing to 1)."""
    n, d = dim, 2 * s + 1
    points, weights = [], []
    for i in range(s + 1):
        denominator = float(d + n - 2 * i)
        weight = (-1.0) ** i * 2.0 ** (-2 * s) * denominator ** d / (math.factorial(i) * math.factorial(d + n - i))
        for beta in _compositions(n + 1, s - i):
            points.append([(2 * b + 1) / denominator for b in beta])
            weights.append(weight)
    w = np.array(weights)
    return np.array(points, dtype=np.float64).reshape(len(weights), n + 1), w / w.sum()


def quad_rule(dim: int, degree: int = 1):
    """SimplexQuadRule::degree (quadrature.rs:80-82): the minimal-index rule exact for the given degree."""
    return grundmann_moeller(dim, degree // 2)


def whitney_shapes(dim: int, grade: int, nodes: np.ndarray) -> np.ndarray:
    """[nnodes][C(dim+1, grade+1)][C(dim, grade)]"""
    n, k = dim, grade
    dlam = np.zeros((n + 1, n))          # differentials of the barycentric coordinates in the reference frame
    dlam[0, :] = -1.0
    for i in range(1, n + 1):
        dlam[i, i - 1] = 1.0
    dofs = _colex_subsets(n + 1, k + 1)
    comps = _colex_subsets(n, k)
    out = np.zeros((len(nodes), len(dofs), max(len(comps), 1) if k <= n else 0))
    kf = float(math.factorial(k))
    for q, lam in enumerate(np.asarray(nodes, dtype=np.float64)):
        for a, sigma in enumerate(dofs):
            for i in range(k + 1):
                rest = [sigma[j] for j in range(k + 1) if j != i]
                for c, axes in enumerate(comps):
                    m = dlam[np.ix_(rest, list(axes))] if k > 0 else np.zeros((0, 0))
                    wedge = float(np.linalg.det(m)) if k > 0 else 1.0
                    out[q, a, c] += kf * (-1.0) ** i * lam[sigma[i]] * wedge
    return out
