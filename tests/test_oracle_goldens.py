"""Pins the CPU oracle (oracle/) against the golden vectors and known-answer
tests the reference holds for the assembly path.  Every test names the
reference file:line it restates (paths relative to /root/reference/)."""
import itertools
import math

import numpy as np
import pytest

from oracle import oracle as O

EPS = np.finfo(float).eps


def rel_eq(a, b, eps=EPS, max_relative=EPS):
    """approx::assert_relative_eq! semantics (default epsilon = max_relative = f64::EPSILON)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = np.abs(a - b)
    big = np.maximum(np.abs(a), np.abs(b))
    return bool(np.all((d <= eps) | (d <= big * max_relative)))


# ---------------------------------------------------------------- combinatorics
def test_permutation_colex_order_is_frozen():
    # crates/multiindex/src/permutation.rs:220-234
    p, s = O.permutations(3)
    assert p.tolist() == [[2, 1, 0], [1, 2, 0], [2, 0, 1], [0, 2, 1], [1, 0, 2], [0, 1, 2]]
    assert s.tolist() == [-1.0, 1.0, 1.0, -1.0, -1.0, 1.0]


def test_permutations_are_lex_on_reversed_word():
    # crates/multiindex/src/permutation.rs:236-250
    for n in range(0, 6):
        p, _ = O.permutations(n)
        rev = [tuple(reversed(w)) for w in p.tolist()]
        assert rev == sorted(rev)
        assert len(rev) == math.factorial(n)


def test_combination_rank_matches_enumeration():
    # crates/multiindex/src/combination.rs:206-215
    for n in range(0, 8):
        for c in range(0, n + 1):
            combs = O.combinations(n, c)
            assert len(combs) == math.comb(n, c)
            for r, s in enumerate(combs.tolist()):
                assert sum(math.comb(v, i + 1) for i, v in enumerate(s)) == r
            keys = [tuple(reversed(s)) for s in combs.tolist()]
            assert keys == sorted(keys)


def test_unit_boundary_squares_to_zero():
    # crates/simplicial/src/topology/simplex.rs:223-238 (d∘d = 0)
    for n in range(1, 6):
        for k in range(2, n + 1):
            assert not np.any(O.unit_boundary_operator(n, k - 1) @ O.unit_boundary_operator(n, k))


# ---------------------------------------------------------------- element goldens
def unit_elmat(kind, n, k):
    return O.elmat(kind, n, k, O.unit_simplex_lengths_sq(n))


def test_hodge_mass_dim2_grade1():
    # crates/formoniq/src/operators.rs:931-943
    exp = np.array([[1 / 3, 1 / 6, 0], [1 / 6, 1 / 3, 0], [0, 0, 1 / 6]])
    got = unit_elmat(O.MASS, 2, 1)
    assert rel_eq(got, exp)
    assert np.array_equal(got == 0.0, exp == 0.0)  # exact zeros stay exact


def test_dif_trial_n2_k1():
    # crates/formoniq/src/operators.rs:945-959
    exp = np.array([[-1 / 2, 1 / 3, 1 / 6], [-1 / 2, 1 / 6, 1 / 3], [0, -1 / 6, 1 / 6]])
    assert rel_eq(unit_elmat(O.DIF_TRIAL, 2, 1), exp)


def test_dif_test_n2_k1():
    # crates/formoniq/src/operators.rs:961-975
    exp = np.array([[-1 / 2, -1 / 2, 0], [1 / 3, 1 / 6, -1 / 6], [1 / 6, 1 / 3, 1 / 6]])
    assert rel_eq(unit_elmat(O.DIF_TEST, 2, 1), exp)


def test_hodge_mass0_is_scalar_mass():
    # crates/formoniq/src/operators.rs:920-929
    for n in range(0, 4):
        nv = n + 1
        q = np.full((nv, nv), 1.0 / (nv * (nv + 1)))
        np.fill_diagonal(q, 2.0 / (nv * (nv + 1)))
        assert rel_eq(unit_elmat(O.MASS, n, 0), q / math.factorial(n))


def test_laplacian_refcell_dims_1_to_10():
    # crates/formoniq/tests/unit_elmat.rs:31-50
    for n in range(1, 11):
        exp = np.zeros((n + 1, n + 1))
        exp[0, 0] = n
        for i in range(1, n + 1):
            exp[i, 0] = exp[0, i] = -1
            exp[i, i] = 1
        exp = exp * (1.0 / math.factorial(n))
        assert rel_eq(unit_elmat(O.DIF_BOTH, n, 1), exp), n


def test_mass_refcell():
    # crates/formoniq/tests/unit_elmat.rs:52-80
    mats = [
        [[1.0]],
        [[1 / 3, 1 / 6], [1 / 6, 1 / 3]],
        [[1 / 12, 1 / 24, 1 / 24], [1 / 24, 1 / 12, 1 / 24], [1 / 24, 1 / 24, 1 / 12]],
        [[1 / 60, 1 / 120, 1 / 120, 1 / 120], [1 / 120, 1 / 60, 1 / 120, 1 / 120],
         [1 / 120, 1 / 120, 1 / 60, 1 / 120], [1 / 120, 1 / 120, 1 / 120, 1 / 60]],
    ]
    for n, m in enumerate(mats):
        assert rel_eq(unit_elmat(O.MASS, n, 0), np.array(m)), n


def test_lumped_mass_refcell():
    # crates/formoniq/tests/unit_elmat.rs:82-90
    for n in range(1, 11):
        exp = np.eye(n + 1) * (1.0 / math.factorial(n) / (n + 1))
        assert rel_eq(unit_elmat(O.LUMPED, n, 0), exp)


def independent_mass(n, k, s):
    """Independent numpy formula for the Whitney mass on one cell:
    M = vol * sum_{a,b} (k!)^2 (-1)^{a+b} det(G[s\\a, t\\b]) Q[s_a,t_b],
    G = D g^-1 D^T the Gramian of the barycentric differentials (Cauchy-Binet
    form of operators.rs:84-94; the check of operators.rs:977-997 by hand)."""
    g, _, _ = O.cell_geometry(n, s) if n else (np.zeros((0, 0)), None, None)
    D = np.zeros((n + 1, n))
    D[0, :] = -1
    D[1:, :] = np.eye(n)
    G = D @ np.linalg.inv(g) @ D.T if n else np.zeros((1, 1))
    nv = n + 1
    Q = np.full((nv, nv), 1.0 / (nv * (nv + 1))) + np.eye(nv) / (nv * (nv + 1))
    vol = math.sqrt(abs(np.linalg.det(g))) / math.factorial(n) if n else 1.0
    dofs = list(itertools.combinations(range(nv), k + 1))
    dofs.sort(key=lambda t: tuple(reversed(t)))
    M = np.zeros((len(dofs), len(dofs)))
    kf = math.factorial(k)
    for i, si in enumerate(dofs):
        for j, sj in enumerate(dofs):
            acc = 0.0
            for a in range(k + 1):
                for b in range(k + 1):
                    ra = [v for t, v in enumerate(si) if t != a]
                    rb = [v for t, v in enumerate(sj) if t != b]
                    minor = np.linalg.det(G[np.ix_(ra, rb)]) if k else 1.0
                    acc += kf * kf * (-1) ** (a + b) * minor * Q[si[a], sj[b]]
            M[i, j] = vol * acc
    return M


def random_cell_lengths(n, rng, lorentz=False):
    """Squared edge lengths of a random non-degenerate simplex in R^n."""
    while True:
        x = rng.normal(size=(n + 1, n))
        eta = np.ones(n)
        if lorentz:
            eta[0] = -1.0
        s = np.zeros(math.comb(n + 1, 2))
        for j in range(1, n + 1):
            for i in range(j):
                d = x[j] - x[i]
                s[i + math.comb(j, 2)] = np.sum(eta * d * d)
        g, _, vol = O.cell_geometry(n, s)
        if vol > 1e-2 and (not lorentz or np.min(np.abs(s)) > 1e-2):
            return s


@pytest.mark.parametrize("lorentz", [False, True])
def test_mass_matches_independent_formula_on_random_cells(lorentz):
    rng = np.random.default_rng(7)
    for n in range(1, 5):
        for k in range(0, n + 1):
            for _ in range(3):
                s = random_cell_lengths(n, rng, lorentz)
                got = O.elmat(O.MASS, n, k, s)
                exp = independent_mass(n, k, s)
                scale = np.abs(exp).max()
                assert np.abs(got - exp).max() <= 1e-11 * scale, (n, k)


def test_sandwiches_are_boundary_conjugations():
    # crates/formoniq/src/operators.rs:201-211 (A = ∂ M D)
    rng = np.random.default_rng(3)
    for n in range(1, 5):
        for k in range(1, n + 1):
            s = random_cell_lengths(n, rng)
            M = O.elmat(O.MASS, n, k, s)
            B = O.unit_boundary_operator(n, k)
            sc = np.abs(M).max()
            assert np.abs(O.elmat(O.DIF_TRIAL, n, k, s) - M @ B.T).max() <= 1e-13 * sc
            assert np.abs(O.elmat(O.DIF_TEST, n, k, s) - B @ M).max() <= 1e-13 * sc
            assert np.abs(O.elmat(O.DIF_BOTH, n, k, s) - B @ M @ B.T).max() <= 1e-13 * sc * (n + 1)


# ---------------------------------------------------------------- mesh goldens
def test_unit_cube_mesh():
    # crates/regge/src/mesher/cartesian.rs:265-296, simplicial/src/mesher/grid.rs:118-141
    cx = O.Complex.kuhn(3, 1)
    assert cx.skeleton(3).tolist() == [[0, 1, 3, 7], [0, 2, 3, 7], [0, 1, 5, 7], [0, 4, 5, 7], [0, 2, 6, 7],
                                      [0, 4, 6, 7]]
    coords = O.kuhn_vertex_coords(3, 1)
    exp = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]]
    assert coords.tolist() == exp


def test_unit_square_mesh():
    # crates/regge/src/mesher/cartesian.rs:298-330
    cx = O.Complex.kuhn(2, 2)
    assert cx.skeleton(2).tolist() == [[0, 1, 4], [0, 3, 4], [1, 2, 5], [1, 4, 5], [3, 4, 7], [3, 6, 7],
                                      [4, 5, 8], [4, 7, 8]]
    coords = O.kuhn_vertex_coords(2, 2)
    exp = [[0, 0], [.5, 0], [1, 0], [0, .5], [.5, .5], [1, .5], [0, 1], [.5, 1], [1, 1]]
    assert coords.tolist() == exp


def test_kuhn_counts_and_euler():
    # crates/simplicial/src/mesher/grid.rs:143-160 + SURVEY Appendix C
    for dim in range(1, 5):
        for N in range(1, 4 if dim < 4 else 3):
            cx = O.Complex.kuhn(dim, N)
            assert cx.ncells == math.factorial(dim) * N ** dim
            assert cx.nsimplices(0) == (N + 1) ** dim
            chi = sum((-1) ** j * cx.nsimplices(j) for j in range(dim + 1))
            assert chi == 1
    N = 3
    cx = O.Complex.kuhn(3, N)
    assert cx.nsimplices(1) == 7 * N ** 3 + 9 * N ** 2 + 3 * N
    assert cx.nsimplices(2) == 12 * N ** 3 + 6 * N ** 2


def test_skeletons_are_colex_sorted_and_faces_consistent():
    # crates/simplicial/src/topology/complex.rs:371-392, skeleton.rs:61-63
    cx = O.Complex.kuhn(3, 2)
    for j in range(4):
        sk = cx.skeleton(j)
        keys = [tuple(reversed(s)) for s in sk.tolist()]
        assert keys == sorted(set(keys))
        cf = cx.cell_faces(j)
        cells = cx.skeleton(3)
        subs = O.combinations(4, j + 1)
        for c in (0, 5, len(cells) - 1):
            for l, sub in enumerate(subs):
                assert sk[cf[c, l]].tolist() == cells[c][sub].tolist()


# ---------------------------------------------------------------- assembly
def kuhn_problem(dim, N, scale=1.0, jitter=False, minkowski=False):
    cx = O.Complex.kuhn(dim, N)
    vmax = np.full(dim, scale)
    if minkowski:
        vmax[0] = 0.7 * scale
    coords = O.kuhn_vertex_coords(dim, N, None, vmax)
    if jitter:
        coords = O.jitter_coords(coords, N)
    diag = np.ones(dim)
    if minkowski:
        diag[0] = -1.0
    return cx, cx.edge_lengths_sq(coords, diag), coords


def fem3d_galmat(N):
    # crates/formoniq/tests/fem3d.rs:18-103, restated in numpy
    nvd = N + 1
    h = 1.0 / N
    nv = nvd ** 3
    xyz = np.zeros((nv, 3))
    for z in range(nvd):
        for y in range(nvd):
            for x in range(nvd):
                xyz[x + nvd * (y + nvd * z)] = (h * x, h * y, h * z)
    tets = [[0, 1, 3, 7], [0, 1, 5, 7], [0, 2, 3, 7], [0, 2, 6, 7], [0, 4, 5, 7], [0, 4, 6, 7]]
    tet_vol = h ** 3 / 6
    A = np.zeros((nv, nv))
    for zb in range(N):
        for yb in range(N):
            for xb in range(N):
                box = [(xb + i) + nvd * ((yb + j) + nvd * (zb + k)) for k in range(2) for j in range(2)
                       for i in range(2)]
                for t in tets:
                    iv = [box[i] for i in t]
                    P = [xyz[i] for i in iv]
                    ns = []
                    for i in range(4):
                        face = [P[j] for j in range(4) if j != i]
                        sign = 1.0 if i % 2 == 0 else -1.0
                        ns.append(sign * np.cross(face[1] - face[0], face[2] - face[0]))
                    el = np.array([[ns[i] @ ns[j] for j in range(4)] for i in range(4)]) / (36.0 * tet_vol)
                    for a, ga in enumerate(iv):
                        for b, gb in enumerate(iv):
                            A[ga, gb] += el[a, b]
    return A


def test_feec_vs_fem3d():
    # crates/formoniq/tests/fem3d.rs:9-16 (epsilon 1e-12); N limited for CPU time
    for N in range(1, 5):
        cx, s, _ = kuhn_problem(3, N)
        feec = cx.assemble(s, O.DIF_BOTH, 1).to_scipy().toarray()
        fem = fem3d_galmat(N)
        assert rel_eq(feec, fem, eps=1e-12), N


def kron_sum_laplacian(dim, nv):
    L1 = 2 * np.eye(nv, dtype=int) - np.eye(nv, k=1, dtype=int) - np.eye(nv, k=-1, dtype=int)
    out = np.zeros((nv ** dim, nv ** dim), dtype=int)
    for a in range(dim):
        mats = [np.eye(nv, dtype=int)] * dim
        mats = [L1 if b == a else np.eye(nv, dtype=int) for b in range(dim)]
        m = mats[-1]
        for b in range(dim - 2, -1, -1):
            m = np.kron(m, mats[b])
        out += m
    return out


def test_feec_vs_fdm_interior():
    # crates/formoniq/tests/fdm.rs:172-222 (integer Laplacian stencil, dims 1..4)
    for N in range(1, 4):
        for dim in range(1, 5):
            if dim == 4 and N > 2:
                continue
            Nb = N + 2
            cx, s, _ = kuhn_problem(dim, Nb, scale=float(Nb))
            A = cx.assemble(s, O.DIF_BOTH, 1).to_scipy().toarray()
            M = cx.assemble(s, O.MASS, 0).to_scipy().toarray()
            b = M @ np.ones(cx.nsimplices(0))
            A = A / b[:, None]
            nvd = Nb + 1
            idx = np.arange(nvd ** dim)
            interior = np.ones(nvd ** dim, bool)
            for a in range(dim):
                c = (idx // nvd ** a) % nvd
                interior &= (c != 0) & (c != Nb)
            A = A[np.ix_(interior, interior)]
            assert np.abs(A - np.round(A)).max() <= 10e-12
            assert np.array_equal(np.round(A).astype(int), kron_sum_laplacian(dim, N + 1)), (dim, N)


def global_boundary(cx, k):
    """∂_k : k-simplices -> (k-1)-simplices (complex.rs:173-179)."""
    sk, skm = cx.skeleton(k), cx.skeleton(k - 1)
    lookup = {tuple(s): i for i, s in enumerate(skm.tolist())}
    B = np.zeros((len(skm), len(sk)))
    for j, s in enumerate(sk.tolist()):
        for i in range(len(s)):
            B[lookup[tuple(s[:i] + s[i + 1:])], j] = (-1) ** i
    return B


@pytest.mark.parametrize("jitter", [False, True])
def test_local_route_equals_global_route(jitter):
    # crates/formoniq/src/whitney_complex.rs:733-772, :780-815 (tol 1e-12*scale)
    for dim, N in ((2, 3), (3, 2)):
        cx, s, _ = kuhn_problem(dim, N, jitter=jitter)
        for k in range(1, dim + 1):
            M = cx.assemble(s, O.MASS, k).to_scipy().toarray()
            D = global_boundary(cx, k).T  # d_{k-1} = ∂_k^T
            both = cx.assemble(s, O.DIF_BOTH, k).to_scipy().toarray()
            test = cx.assemble(s, O.DIF_TEST, k).to_scipy().toarray()
            trial = cx.assemble(s, O.DIF_TRIAL, k).to_scipy().toarray()
            sc = np.abs(M).max() * 10
            assert np.abs(both - D.T @ M @ D).max() <= 1e-12 * sc
            assert np.abs(test - D.T @ M).max() <= 1e-12 * sc
            assert np.abs(trial - M @ D).max() <= 1e-12 * sc


def test_mass_matrices_are_symmetric_and_sum_to_volume():
    for dim, N in ((2, 4), (3, 2)):
        cx, s, _ = kuhn_problem(dim, N, jitter=True)
        M0 = cx.assemble(s, O.MASS, 0).to_scipy().toarray()
        assert abs(M0.sum() - sum(
            O.cell_geometry(dim, s[cx.cell_faces(1)[c]])[2] for c in range(cx.ncells))) < 1e-12
        for k in range(dim + 1):
            M = cx.assemble(s, O.MASS, k).to_scipy().toarray()
            assert np.abs(M - M.T).max() <= 1e-14


def test_lorentzian_lengths_and_mass():
    # crates/formoniq/src/galerkin.rs:401-422: signed Regge data; det g < 0
    for dim in (2, 3):
        cx, s, coords = kuhn_problem(dim, 2, minkowski=True)
        assert (s < 0).any() and (s > 0).any() and not (s == 0).any()
        for k in range(dim + 1):
            M = cx.assemble(s, O.MASS, k).to_scipy().toarray()
            assert np.isfinite(M).all() and np.abs(M - M.T).max() < 1e-13
            exp0 = independent_mass(dim, k, s[cx.cell_faces(1)[0]])
            got0 = cx.elmat_batch(s, O.MASS, k, 0, 1)[0]
            assert np.abs(got0 - exp0).max() <= 1e-12 * np.abs(exp0).max()


def test_drop_filter_and_structural_counts():
    # crates/formoniq/src/galerkin.rs:173 + SURVEY §7 H1 / Appendix C
    N = 4
    cx, s, _ = kuhn_problem(3, N)
    structural = {(O.MASS, 0): 15 * N ** 3 + 21 * N ** 2 + 9 * N + 1, (O.MASS, 1): 115 * N ** 3 + 45 * N ** 2 + 3 * N,
                  (O.DIF_TEST, 1): 50 * N ** 3 + 36 * N ** 2 + 6 * N, (O.DIF_BOTH, 2): 115 * N ** 3 + 45 * N ** 2 + 3 * N}
    for (kind, k), nnz in structural.items():
        full = cx.assemble(s, kind, k, drop_zeros=False)
        ref = cx.assemble(s, kind, k, drop_zeros=True)
        assert full.nnz == nnz
        assert ref.nnz <= nnz
        assert np.abs(full.to_scipy().toarray() - ref.to_scipy().toarray()).max() == 0.0
    assert cx.assemble(s, O.MASS, 1).nnz < structural[(O.MASS, 1)]  # dyadic mesh: exact zeros dropped
    cxj, sj, _ = kuhn_problem(3, 3, jitter=True)
    for (kind, k) in structural:
        assert cxj.assemble(sj, kind, k).nnz == cxj.assemble(sj, kind, k, drop_zeros=False).nnz


def test_threaded_assembly_is_identical():
    # galerkin.rs:150-159,181: ordered concat makes the result thread-count independent
    cx, s, _ = kuhn_problem(3, 3, jitter=True)
    a = cx.assemble(s, O.MASS, 1, nthreads=1).arrays()
    b = cx.assemble(s, O.MASS, 1, nthreads=5).arrays()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


# ---------------------------------------------------------------- SpMV / Krylov
def test_spmv_vs_dense():
    # crates/iterative/src/operator.rs:21-29
    cx, s, _ = kuhn_problem(3, 2, jitter=True)
    A = cx.assemble(s, O.MASS, 1)
    x = np.array([((7 * i) % 13) - 6 for i in range(A.ncols)], float)  # matfree.rs:205-207
    assert np.abs(A.spmv(x) - A.to_scipy().toarray() @ x).max() <= 1e-13


def test_cg_and_minres_solve():
    # crates/iterative/src/krylov.rs:224-320 (laws: residual below rtol, solution matches direct)
    cx, s, _ = kuhn_problem(2, 6, jitter=True)
    M = cx.assemble(s, O.MASS, 1)
    b = np.array([(i % 7) - 3.0 for i in range(M.nrows)])
    dense = M.to_scipy().toarray()
    exact = np.linalg.solve(dense, b)
    for solver in (M.cg, M.minres):
        for pc in (0, 1):
            x, rep = solver(b, rtol=1e-12, precond=pc)
            assert rep["converged"]
            assert np.abs(x - exact).max() <= 1e-8 * np.abs(exact).max()
    x, rep = M.cg(np.zeros(M.nrows))
    assert rep["iters"] == 0 and rep["converged"] and not x.any()


def test_pseudo_random_is_splitmix():
    # crates/formoniq/src/linalg/eigen.rs:259-268
    def ref(seed, index):
        m = (1 << 64) - 1
        z = (seed * 0x9E3779B97F4A7C15 + index * 0xD1B54A32D192ED03 + 0x9E3779B97F4A7C15) & m
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        z ^= z >> 31
        return (z >> 11) / float(1 << 53) * 2.0 - 1.0
    for seed in range(3):
        for idx in (0, 1, 17, 12345):
            assert O.pseudo_random(seed, idx) == ref(seed, idx)


def test_assemble_vector_law_of_the_reference():
    # crates/formoniq/src/galerkin.rs:330-372: with the source a Whitney form lambda_tau the assembled load is the
    # column tau of the mass matrix; and the restatement equals the plain sequential loop of galerkin.rs:305-309.
    for dim, n, grade in ((2, 4, 1), (3, 3, 1), (3, 2, 2)):
        cx, s, _ = kuhn_problem(dim, n, jitter=True)
        nl = O.nlocal(dim, grade)
        faces = cx.cell_faces(grade)
        elm = cx.elmat_batch(s, O.MASS, grade)
        mass = cx.assemble(s, O.MASS, grade).to_scipy().toarray()
        for tau in (0, cx.nsimplices(grade) // 2, cx.nsimplices(grade) - 1):
            ev = np.zeros((cx.ncells, nl))
            for c in range(cx.ncells):
                hit = np.flatnonzero(faces[c] == tau)
                if hit.size:
                    ev[c] = elm[c][:, hit[0]]
            got = O.assemble_vector(cx, grade, ev)
            assert np.abs(got - mass[:, tau]).max() <= 1e-13 * np.abs(mass[:, tau]).max()
        rng = np.random.default_rng(dim)
        ev = rng.standard_normal((cx.ncells, nl))
        ev[rng.random(ev.shape) < 0.3] = 0.0
        naive = np.zeros(cx.nsimplices(grade))
        for c in range(cx.ncells):
            for i in range(nl):
                if ev[c, i] != 0.0:
                    naive[faces[c, i]] += ev[c, i]
        assert np.array_equal(O.assemble_vector(cx, grade, ev), naive)


def test_grundmann_moeller_and_whitney_tables():
    # simplicial/src/atlas/quadrature.rs:176-196 (the rule integrates every barycentric monomial of degree <= 2s+1
    # exactly, in every dimension) and the de Rham duality of the lowest-order Whitney forms (int_tau W_sigma = delta)
    from formoniq_b200 import quadrature as Q
    for dim in range(0, 5):
        for s in range(0, 4):
            pts, w = Q.grundmann_moeller(dim, s)
            assert abs(w.sum() - 1.0) < 1e-14 and pts.shape == (len(w), dim + 1)
            assert np.abs(pts.sum(axis=1) - 1.0).max() < 1e-15
            for deg in range(0, 2 * s + 2):
                for alpha in Q._compositions(dim + 1, deg):
                    val = sum(wi * np.prod(p ** np.array(alpha)) for p, wi in zip(pts, w))
                    exact = math.factorial(dim) * np.prod([math.factorial(a) for a in alpha]) / math.factorial(dim + deg)
                    assert abs(val - exact) < 1e-12
    for dim in (1, 2, 3, 4):
        verts = np.vstack([np.zeros(dim), np.eye(dim)])
        dofs = Q._colex_subsets(dim + 1, 2)
        for b, (p, q) in enumerate(dofs):
            lam = np.zeros(dim + 1)
            lam[p] = lam[q] = 0.5
            w_mid = Q.whitney_shapes(dim, 1, [lam])[0]          # [dof][axis] at the midpoint of edge b
            for a in range(len(dofs)):
                assert abs(w_mid[a] @ (verts[q] - verts[p]) - (1.0 if a == b else 0.0)) < 1e-14
        centroid = np.full(dim + 1, 1.0 / (dim + 1))
        assert np.allclose(Q.whitney_shapes(dim, 0, [centroid])[0][:, 0], centroid)        # W_v = lambda_v
        assert np.allclose(Q.whitney_shapes(dim, dim, [centroid])[0], math.factorial(dim))  # the volume form n! dx


def test_source_form_law_of_the_reference():
    # crates/formoniq/src/galerkin.rs:330-372: the source a Whitney form lambda_tau, the load is the column tau of the
    # mass matrix (the integrand is quadratic: the degree-3 rule is exact) - pins rule + shape table + Lambda^k g^-1 +
    # volume of the restatement against the golden-pinned mass matrices, Riemannian and Lorentzian.
    from formoniq_b200 import quadrature as Q
    for dim, n, grade, mink in ((2, 3, 1, False), (3, 2, 1, False), (3, 2, 2, False), (2, 3, 0, False), (3, 2, 3, False),
                                (3, 2, 1, True)):
        cx, s, _ = kuhn_problem(dim, n, jitter=not mink, minkowski=mink)
        nodes, weights = Q.quad_rule(dim, 3)
        shapes = Q.whitney_shapes(dim, grade, nodes)
        faces = cx.cell_faces(grade)
        mass = cx.assemble(s, O.MASS, grade).to_scipy().toarray()
        for tau in (0, cx.nsimplices(grade) // 2):
            samples = np.zeros((cx.ncells, len(weights), shapes.shape[2]))
            for c in range(cx.ncells):
                hit = np.flatnonzero(faces[c] == tau)
                if hit.size:
                    samples[c] = shapes[:, hit[0], :]
            ev = O.source_element_vectors(cx, s, grade, weights, shapes, samples)
            load = O.assemble_vector(cx, grade, ev)
            assert np.abs(load - mass[:, tau]).max() <= 1e-12 * np.abs(mass[:, tau]).max(), (dim, grade, mink)


def test_weighted_hodge_mass_on_a_constant_is_the_closed_form():
    # crates/formoniq/src/operators.rs:897-918: on alpha = c the degree-2 quadrature returns c times the exact HodgeMass,
    # at every dimension and grade (here on jittered and Lorentzian cells too); and assemble_from_elmats reproduces the
    # oracle's assembled mass from the oracle's own element matrices (pattern and values).
    from formoniq_b200 import quadrature as Q
    for dim, n, mink in ((1, 4, False), (2, 3, False), (3, 2, False), (3, 2, True)):
        cx, s, _ = kuhn_problem(dim, n, jitter=not mink, minkowski=mink)
        for grade in range(dim + 1):
            nodes, weights = Q.quad_rule(dim, 2)
            shapes = Q.whitney_shapes(dim, grade, nodes)
            exact = cx.elmat_batch(s, O.MASS, grade)
            for cval in (1.0, 2.5):
                got = O.weighted_mass_elmats(cx, s, grade, weights, shapes, np.full((cx.ncells, len(weights)), cval))
                assert np.abs(got - cval * exact).max() <= 1e-12 * np.abs(exact).max(), (dim, grade, mink)
            ref = cx.assemble(s, O.MASS, grade)
            erp, eci, eva = ref.arrays()
            mine = O.assemble_from_elmats(cx, grade, grade, exact)
            assert np.array_equal(mine.indptr, erp) and np.array_equal(mine.indices, eci)
            assert np.array_equal(mine.data, eva)
