"""Shared helpers for the tests: reference-style problems built with the oracle."""
import numpy as np

from oracle import oracle as O


def kuhn_problem(dim, shape, scale=1.0, jitter=False, minkowski=False):
    """Oracle-side Kuhn mesh: (complex, lengths, coords, ambient_diag, vmax)."""
    cx = O.Complex.kuhn(dim, shape)
    vmax = np.full(dim, float(scale))
    if minkowski:
        vmax[0] = 0.7 * scale  # CAUSAL_TIME_SCALE, regge/src/mesher/cartesian.rs:21
    coords = O.kuhn_vertex_coords(dim, shape, None, vmax)
    if jitter:
        coords = O.jitter_coords(coords, shape)
    diag = np.ones(dim)
    if minkowski:
        diag[0] = -1.0
    return cx, cx.edge_lengths_sq(coords, diag), coords, diag, vmax


def all_blocks(dim):
    """Every (kind, grade) the reference can ask of a dim-complex."""
    out = []
    for k in range(0, dim + 2):
        for kind in (O.MASS, O.DIF_TRIAL, O.DIF_TEST, O.DIF_BOTH):
            if k == 0 and kind != O.MASS:
                continue
            if k == dim + 1 and kind != O.DIF_BOTH:
                continue
            out.append((kind, k))
    return out


def mesh_from_oracle(fq, ctx, cx, lengths):
    ns = [cx.nsimplices(j) for j in range(cx.dim + 1)]
    faces = [cx.cell_faces(j) for j in range(cx.dim + 1)]
    return fq.Mesh.from_arrays(ctx, cx.dim, ns, faces, lengths)


def same_bits_mod_zero_sign(a, b):
    """Bitwise equality up to the sign of zero (and both finite)."""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all(np.isfinite(a))) and bool(np.all(a == b))
