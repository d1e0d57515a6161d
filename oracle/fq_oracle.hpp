// fq_oracle.hpp — CPU ORACLE for the formoniq Galerkin-assembly hot path.
//
// TEST INFRASTRUCTURE ONLY.  This header is a plain C++17 restatement of the
// reference's (luiswirth/formoniq, Rust) CPU algorithm, written from reading
// the reference sources.  It exists so the CUDA path can be checked against
// the reference's arithmetic.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may link or call it; the
// product library (formoniq_b200/csrc) never includes this file.
//
// Pinning status: the reference is a Rust workspace and no Rust toolchain is
// available (nor are its crates vendored), so the reference itself cannot be
// run here.  The oracle is pinned against every golden literal the reference's
// own tests hold for this path (tests/test_oracle_goldens.py lists them with
// file:line).  The *values* are pinned by those goldens; the exact
// floating-point operation order of the third-party pieces (nalgebra
// try_inverse / determinant / gemm, nalgebra-sparse convert_coo_csr) is
// restated from their published algorithms at the versions pinned in
// Cargo.lock (nalgebra 0.35.0, nalgebra-sparse 0.12.0) and is NOT bit-pinned;
// the sparsity pattern is asserted nowhere in the reference's tests, so the
// pattern is "parity unpinned" by the reference and pinned only by this
// restatement of galerkin.rs:173 (the `val != 0.0` filter).
//
// All citations are relative to /root/reference/.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

namespace fqo {

using idx_t = int64_t;

// ---------------------------------------------------------------------------
// Dense row-major matrix (nalgebra DMatrix stand-in; layout is irrelevant to
// the arithmetic, only the loop orders below are).
// ---------------------------------------------------------------------------
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() = default;
  Mat(int r_, int c_, double v = 0.0) : r(r_), c(c_), a(size_t(r_) * c_, v) {}
  double& operator()(int i, int j) { return a[size_t(i) * c + j]; }
  double operator()(int i, int j) const { return a[size_t(i) * c + j]; }
  Mat transpose() const {
    Mat t(c, r);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < c; ++j) t(j, i) = (*this)(i, j);
    return t;
  }
};

// ---------------------------------------------------------------------------
// A0. Combinatorics (crates/multiindex)
// ---------------------------------------------------------------------------

// crates/multiindex/src/count.rs:27-36
inline idx_t binomial(int n, int k) {
  if (k < 0 || n < 0 || k > n) return 0;
  idx_t res = 1;
  for (int i = 1; i <= k; ++i) res = res * (n - k + i) / i;
  return res;
}
inline idx_t factorial(int n) {
  idx_t f = 1;
  for (int i = 2; i <= n; ++i) f *= i;
  return f;
}

using Comb = std::vector<int>;  // strictly ascending

// crates/multiindex/src/combination.rs:154-179 — all `card`-subsets of
// {0..n-1} in colexicographic order (bitsets in increasing integer order).
inline std::vector<Comb> combinations(int n, int card) {
  std::vector<Comb> out;
  if (card < 0 || card > n) return out;
  if (card == 0) {
    out.push_back({});
    return out;
  }
  Comb cur(card);
  for (int i = 0; i < card; ++i) cur[i] = i;
  for (;;) {
    out.push_back(cur);
    // colex successor: bump the lowest element that can move up without
    // colliding, reset everything below it.
    int i = 0;
    while (i + 1 < card && cur[i] + 1 == cur[i + 1]) ++i;
    if (i == card - 1 && cur[i] + 1 >= n) break;
    ++cur[i];
    for (int j = 0; j < i; ++j) cur[j] = j;
  }
  return out;
}

// crates/multiindex/src/monotone.rs:355-361 — rank = sum_i C(s_i, i+1).
inline idx_t rank_of(const Comb& s) {
  idx_t r = 0;
  for (size_t i = 0; i < s.size(); ++i) r += binomial(s[i], int(i) + 1);
  return r;
}

struct Deletion {
  double sign;
  int vertex;
  Comb rest;
};
// crates/multiindex/src/monotone.rs:617-638 — position i: sign (-1)^i,
// deleted symbol s_i, remainder.
inline std::vector<Deletion> deletions(const Comb& s) {
  std::vector<Deletion> out;
  for (size_t i = 0; i < s.size(); ++i) {
    Deletion d;
    d.sign = (i % 2 == 0) ? 1.0 : -1.0;
    d.vertex = s[i];
    for (size_t j = 0; j < s.size(); ++j)
      if (j != i) d.rest.push_back(s[j]);
    out.push_back(std::move(d));
  }
  return out;
}

struct Perm {
  std::vector<int> p;
  double sign;
};
// crates/multiindex/src/permutation.rs:160-181 — colex order = reversal of
// each word of the lexicographic enumeration; sign by inversions (:100-110).
inline std::vector<Perm> permutations_all(int n) {
  std::vector<Perm> out;
  std::vector<int> word(n);
  std::iota(word.begin(), word.end(), 0);
  do {
    Perm q;
    q.p.assign(word.rbegin(), word.rend());
    int inv = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < j; ++i)
        if (q.p[i] > q.p[j]) ++inv;
    q.sign = (inv % 2 == 0) ? 1.0 : -1.0;
    out.push_back(std::move(q));
  } while (std::next_permutation(word.begin(), word.end()));
  return out;
}

// crates/multialgebra/src/lib.rs:447-456 — Leibniz determinant:
// fold(0.0,+) over sigma of sign * fold(1.0,*)_i A[i][sigma(i)].
inline double det_leibniz(const Mat& m) {
  assert(m.r == m.c);
  const int n = m.r;
  double acc = 0.0;
  for (const Perm& s : permutations_all(n)) {
    double prod = 1.0;
    for (int i = 0; i < n; ++i) prod = prod * m(i, s.p[i]);
    acc = acc + s.sign * prod;
  }
  return acc;
}

// ---------------------------------------------------------------------------
// nalgebra 0.35.0 small dense kernels (third-party; restated, see header).
// ---------------------------------------------------------------------------

// gemm for "small" shapes (any dim <= 5): column by column gemv, each gemv an
// axcpy chain: y = (1*A[:,0])*x0 ; y = (1*A[:,j])*xj + 1*y  (SURVEY A4).
// The same loop is used for every shape; for all dims > 5 nalgebra hands over
// to matrixmultiply whose order is unspecified (value parity only there).
inline Mat gemm(const Mat& A, const Mat& B) {
  assert(A.c == B.r);
  Mat C(A.r, B.c, 0.0);
  for (int j = 0; j < B.c; ++j) {
    for (int kk = 0; kk < A.c; ++kk) {
      const double x = B(kk, j);
      for (int i = 0; i < A.r; ++i) {
        const double term = (1.0 * A(i, kk)) * x;
        if (kk == 0)
          C(i, j) = term;
        else
          C(i, j) = term + 1.0 * C(i, j);
      }
    }
  }
  return C;
}

// LU with partial pivoting (nalgebra linalg/lu.rs), used for n >= 4 det and
// n >= 5 inverse.  Returns false if singular.
struct LU {
  int n;
  Mat lu;
  std::vector<int> piv;  // row swaps
  int nswaps = 0;
  bool ok = true;
  explicit LU(const Mat& m) : n(m.r), lu(m), piv(m.r) {
    for (int i = 0; i < n; ++i) {
      int p = i;
      double best = std::fabs(lu(i, i));
      for (int r = i + 1; r < n; ++r)
        if (std::fabs(lu(r, i)) > best) best = std::fabs(lu(r, i)), p = r;
      piv[i] = p;
      if (best == 0.0) {
        ok = false;
        continue;
      }
      if (p != i) {
        ++nswaps;
        for (int c = 0; c < n; ++c) std::swap(lu(i, c), lu(p, c));
      }
      const double d = lu(i, i);
      for (int r = i + 1; r < n; ++r) lu(r, i) = lu(r, i) / d;
      for (int c = i + 1; c < n; ++c) {
        const double pc = lu(i, c);
        for (int r = i + 1; r < n; ++r) lu(r, c) = lu(r, c) - lu(r, i) * pc;
      }
    }
  }
  double determinant() const {
    double d = 1.0;
    for (int i = 0; i < n; ++i) d = d * lu(i, i);
    return (nswaps % 2) ? -d : d;
  }
  Mat inverse() const {
    Mat inv(n, n, 0.0);
    for (int col = 0; col < n; ++col) {
      std::vector<double> b(n, 0.0);
      b[col] = 1.0;
      for (int i = 0; i < n; ++i) std::swap(b[i], b[piv[i]]);
      for (int i = 0; i < n; ++i)
        for (int r = i + 1; r < n; ++r) b[r] = b[r] - lu(r, i) * b[i];
      for (int i = n - 1; i >= 0; --i) {
        b[i] = b[i] / lu(i, i);
        for (int r = 0; r < i; ++r) b[r] = b[r] - lu(r, i) * b[i];
      }
      for (int r = 0; r < n; ++r) inv(r, col) = b[r];
    }
    return inv;
  }
};

// nalgebra linalg/determinant.rs — closed forms n <= 3, LU beyond.
inline double determinant(const Mat& m) {
  const int n = m.r;
  if (n == 0) return 1.0;
  if (n == 1) return m(0, 0);
  if (n == 2) return m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1);
  if (n == 3) {
    const double minor_1 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
    const double minor_2 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
    const double minor_3 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
    return m(0, 0) * minor_1 - m(0, 1) * minor_2 + m(0, 2) * minor_3;
  }
  return LU(m).determinant();
}

// nalgebra linalg/inverse.rs — closed forms n <= 4, LU beyond.
inline bool try_inverse(const Mat& m, Mat& out) {
  const int n = m.r;
  out = Mat(n, n);
  if (n == 0) return true;
  if (n == 1) {
    if (m(0, 0) == 0.0) return false;
    out(0, 0) = 1.0 / m(0, 0);
    return true;
  }
  if (n == 2) {
    const double m11 = m(0, 0), m12 = m(0, 1), m21 = m(1, 0), m22 = m(1, 1);
    const double d = m11 * m22 - m21 * m12;
    if (d == 0.0) return false;
    out(0, 0) = m22 / d;
    out(0, 1) = -m12 / d;
    out(1, 0) = -m21 / d;
    out(1, 1) = m11 / d;
    return true;
  }
  if (n == 3) {
    const double m11 = m(0, 0), m12 = m(0, 1), m13 = m(0, 2);
    const double m21 = m(1, 0), m22 = m(1, 1), m23 = m(1, 2);
    const double m31 = m(2, 0), m32 = m(2, 1), m33 = m(2, 2);
    const double minor_m12_m23 = m22 * m33 - m32 * m23;
    const double minor_m11_m23 = m21 * m33 - m31 * m23;
    const double minor_m11_m22 = m21 * m32 - m31 * m22;
    const double d = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    if (d == 0.0) return false;
    out(0, 0) = minor_m12_m23 / d;
    out(0, 1) = (m13 * m32 - m33 * m12) / d;
    out(0, 2) = (m12 * m23 - m22 * m13) / d;
    out(1, 0) = -minor_m11_m23 / d;
    out(1, 1) = (m11 * m33 - m31 * m13) / d;
    out(1, 2) = (m13 * m21 - m23 * m11) / d;
    out(2, 0) = minor_m11_m22 / d;
    out(2, 1) = (m12 * m31 - m32 * m11) / d;
    out(2, 2) = (m11 * m22 - m21 * m12) / d;
    return true;
  }
  if (n == 4) {
    // do_inverse4: the unrolled cofactor expansion (column-major m[0..16]).
    double s[16], o[16];
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) s[j * 4 + i] = m(i, j);
    o[0] = s[5] * s[10] * s[15] - s[5] * s[11] * s[14] - s[9] * s[6] * s[15] + s[9] * s[7] * s[14] +
           s[13] * s[6] * s[11] - s[13] * s[7] * s[10];
    o[1] = -s[1] * s[10] * s[15] + s[1] * s[11] * s[14] + s[9] * s[2] * s[15] - s[9] * s[3] * s[14] -
           s[13] * s[2] * s[11] + s[13] * s[3] * s[10];
    o[2] = s[1] * s[6] * s[15] - s[1] * s[7] * s[14] - s[5] * s[2] * s[15] + s[5] * s[3] * s[14] +
           s[13] * s[2] * s[7] - s[13] * s[3] * s[6];
    o[3] = -s[1] * s[6] * s[11] + s[1] * s[7] * s[10] + s[5] * s[2] * s[11] - s[5] * s[3] * s[10] -
           s[9] * s[2] * s[7] + s[9] * s[3] * s[6];
    o[4] = -s[4] * s[10] * s[15] + s[4] * s[11] * s[14] + s[8] * s[6] * s[15] - s[8] * s[7] * s[14] -
           s[12] * s[6] * s[11] + s[12] * s[7] * s[10];
    o[5] = s[0] * s[10] * s[15] - s[0] * s[11] * s[14] - s[8] * s[2] * s[15] + s[8] * s[3] * s[14] +
           s[12] * s[2] * s[11] - s[12] * s[3] * s[10];
    o[6] = -s[0] * s[6] * s[15] + s[0] * s[7] * s[14] + s[4] * s[2] * s[15] - s[4] * s[3] * s[14] -
           s[12] * s[2] * s[7] + s[12] * s[3] * s[6];
    o[7] = s[0] * s[6] * s[11] - s[0] * s[7] * s[10] - s[4] * s[2] * s[11] + s[4] * s[3] * s[10] +
           s[8] * s[2] * s[7] - s[8] * s[3] * s[6];
    o[8] = s[4] * s[9] * s[15] - s[4] * s[11] * s[13] - s[8] * s[5] * s[15] + s[8] * s[7] * s[13] +
           s[12] * s[5] * s[11] - s[12] * s[7] * s[9];
    o[9] = -s[0] * s[9] * s[15] + s[0] * s[11] * s[13] + s[8] * s[1] * s[15] - s[8] * s[3] * s[13] -
           s[12] * s[1] * s[11] + s[12] * s[3] * s[9];
    o[10] = s[0] * s[5] * s[15] - s[0] * s[7] * s[13] - s[4] * s[1] * s[15] + s[4] * s[3] * s[13] +
            s[12] * s[1] * s[7] - s[12] * s[3] * s[5];
    o[11] = -s[0] * s[5] * s[11] + s[0] * s[7] * s[9] + s[4] * s[1] * s[11] - s[4] * s[3] * s[9] -
            s[8] * s[1] * s[7] + s[8] * s[3] * s[5];
    o[12] = -s[4] * s[9] * s[14] + s[4] * s[10] * s[13] + s[8] * s[5] * s[14] - s[8] * s[6] * s[13] -
            s[12] * s[5] * s[10] + s[12] * s[6] * s[9];
    o[13] = s[0] * s[9] * s[14] - s[0] * s[10] * s[13] - s[8] * s[1] * s[14] + s[8] * s[2] * s[13] +
            s[12] * s[1] * s[10] - s[12] * s[2] * s[9];
    o[14] = -s[0] * s[5] * s[14] + s[0] * s[6] * s[13] + s[4] * s[1] * s[14] - s[4] * s[2] * s[13] -
            s[12] * s[1] * s[6] + s[12] * s[2] * s[5];
    o[15] = s[0] * s[5] * s[10] - s[0] * s[6] * s[9] - s[4] * s[1] * s[10] + s[4] * s[2] * s[9] +
            s[8] * s[1] * s[6] - s[8] * s[2] * s[5];
    const double det = s[0] * o[0] + s[1] * o[4] + s[2] * o[8] + s[3] * o[12];
    if (det == 0.0) return false;
    const double inv_det = 1.0 / det;
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) out(i, j) = o[j * 4 + i] * inv_det;
    return true;
  }
  LU lu(m);
  if (!lu.ok) return false;
  out = lu.inverse();
  return true;
}

// ---------------------------------------------------------------------------
// Reference-cell tables
// ---------------------------------------------------------------------------

// crates/simplicial/src/atlas.rs:174-182
inline Mat unit_difbarys(int n) {
  Mat d(n + 1, n, 0.0);
  for (int j = 0; j < n; ++j) d(0, j) = -1.0;
  for (int i = 0; i < n; ++i) d(i + 1, i) = 1.0;
  return d;
}
// crates/simplicial/src/atlas.rs:198-204
inline Mat unit_bary_gramian(int n) {
  const int nv = n + 1;
  const double scale = 1.0 / double(nv * (nv + 1));
  Mat q(nv, nv, scale);
  for (int i = 0; i < nv; ++i) q(i, i) = 2.0 * scale;
  return q;
}
// crates/simplicial/src/atlas.rs:102-105
inline double unit_simplex_volume(int n) { return 1.0 / double(factorial(n)); }

// crates/multialgebra/src/lib.rs:493-514 — k-th compound matrix.
inline Mat exterior_power(const Mat& map, int k) {
  const idx_t nr = binomial(map.r, k), nc = binomial(map.c, k);
  Mat out(int(nr), int(nc), 0.0);
  if (k < 0) return out;
  const auto rows = combinations(map.r, k), cols = combinations(map.c, k);
  Mat minor(k, k);
  for (size_t i = 0; i < rows.size(); ++i)
    for (size_t j = 0; j < cols.size(); ++j) {
      for (int ii = 0; ii < k; ++ii)
        for (int jj = 0; jj < k; ++jj) minor(ii, jj) = map(rows[i][ii], cols[j][jj]);
      out(int(i), int(j)) = det_leibniz(minor);
    }
  return out;
}

// crates/multialgebra/src/lib.rs:285-305 (alternating branch): all k x k
// minors of a square form on colex k-subsets.
inline Mat induced_form_alternating(const Mat& single, int k) {
  return exterior_power(single, k);
}

// crates/simplicial/src/topology/simplex.rs:223-238
inline Mat unit_boundary_operator(int n, int k) {
  const int nrows = (k - 1 >= 0 && k - 1 <= n) ? int(binomial(n + 1, k)) : 0;
  const int ncols = (k >= 0 && k <= n) ? int(binomial(n + 1, k + 1)) : 0;
  Mat b(nrows, ncols, 0.0);
  if (nrows == 0 || ncols == 0) return b;
  const auto cofaces = combinations(n + 1, k + 1);
  for (size_t ic = 0; ic < cofaces.size(); ++ic)
    for (const Deletion& d : deletions(cofaces[ic])) b(int(rank_of(d.rest)), int(ic)) = d.sign;
  return b;
}

// ---------------------------------------------------------------------------
// A1. Regge metric (crates/regge/src/lengths/simplex.rs:308-326)
// ---------------------------------------------------------------------------
inline int edge_index(int vi, int vj) {  // simplex.rs:209
  const int lo = std::min(vi, vj), hi = std::max(vi, vj);
  return int(binomial(lo, 1) + binomial(hi, 2));
}
inline Mat metric_from_lengths(int n, const double* s) {
  Mat g(n, n, 0.0);
  for (int i = 0; i < n; ++i) g(i, i) = s[edge_index(0, i + 1)];
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      const double s0i = s[edge_index(0, i + 1)];
      const double s0j = s[edge_index(0, j + 1)];
      const double sij = s[edge_index(i + 1, j + 1)];
      const double val = 0.5 * (s0i + s0j - sij);
      g(i, j) = val;
      g(j, i) = val;
    }
  return g;
}
// SimplexLengthsSq::unit(dim): the unit simplex (edges from vertex 0 have
// squared length 1, the others 2) — used by the golden element matrices.
inline std::vector<double> unit_simplex_lengths_sq(int n) {
  std::vector<double> s(size_t(binomial(n + 1, 2)));
  for (int j = 1; j <= n; ++j)
    for (int i = 0; i < j; ++i) s[edge_index(i, j)] = (i == 0) ? 1.0 : 2.0;
  return s;
}

// crates/regge/src/lib.rs:26-28 + metric/src/lib.rs:188-190
inline double cell_volume(const Mat& g) {
  return unit_simplex_volume(g.r) * std::sqrt(std::fabs(determinant(g)));
}

// ---------------------------------------------------------------------------
// Element matrices
// ---------------------------------------------------------------------------
enum Kind { MASS = 0, DIF_TRIAL = 1, DIF_TEST = 2, DIF_BOTH = 3, LUMPED = 4 };

// Number of local dofs of grade j on an n-cell (0 off the range [0,n]).
inline int nlocal(int n, int j) { return (j < 0 || j > n) ? 0 : int(binomial(n + 1, j + 1)); }

// Per-(n,k) tables of HodgeMass::new (crates/formoniq/src/operators.rs:70-79).
struct HodgeMassTables {
  int n, k;
  std::vector<Comb> dofs;  // combinations(n+1, k+1)
  Mat difbarys_power;      // Lambda^k(unit_difbarys) : C(n+1,k) x C(n,k)
  Mat bary_gramian;        // Q
  double scale;            // k! as f64
  // WhitneyExpansion::column (form.rs:195-200) tabulated per dof:
  // (coefficient sign*k!, rank of the blade, vertex).
  struct Term {
    double value;
    int blade_rank;
    int vertex;
  };
  std::vector<std::vector<Term>> columns;
  HodgeMassTables(int n_, int k_)
      : n(n_), k(k_), dofs(combinations(n_ + 1, k_ + 1)),
        difbarys_power(exterior_power(unit_difbarys(n_), k_)), bary_gramian(unit_bary_gramian(n_)),
        scale(double(factorial(k_))) {
    for (const Comb& d : dofs) {
      std::vector<Term> col;
      for (const Deletion& del : deletions(d))
        col.push_back(Term{del.sign * scale, int(rank_of(del.rest)), del.vertex});
      columns.push_back(std::move(col));
    }
  }
};

// crates/derham/src/interpolate/form.rs:222-232 (pullback) with :195-200.
inline Mat whitney_pullback(const HodgeMassTables& t, const Mat& blade, const Mat& bary) {
  const int nd = int(t.dofs.size());
  Mat out(nd, nd, 0.0);
  for (int i = 0; i < nd; ++i)
    for (int j = 0; j < nd; ++j) {
      double acc = 0.0;
      for (const auto& a : t.columns[size_t(i)])
        for (const auto& b : t.columns[size_t(j)]) {
          const double term = a.value * b.value * blade(a.blade_rank, b.blade_rank) * bary(a.vertex, b.vertex);
          acc = acc + term;
        }
      out(i, j) = acc;
    }
  return out;
}

// crates/formoniq/src/operators.rs:84-94
inline Mat hodge_mass_element(const HodgeMassTables& t, const Mat& g) {
  Mat ginv;
  if (!try_inverse(g, ginv)) throw std::runtime_error("degenerate metric");
  const Mat form_gramian = induced_form_alternating(ginv, t.k);  // metric/src/tensor.rs:127-131
  const Mat blade_gramian = gemm(gemm(t.difbarys_power, form_gramian), t.difbarys_power.transpose());
  Mat p = whitney_pullback(t, blade_gramian, t.bary_gramian);
  const double vol = cell_volume(g);
  for (double& v : p.a) v = vol * v;
  return p;
}

struct PairingTables {
  int n, k, kind;
  HodgeMassTables mass;
  Mat row;  // boundary (codif), if test side differentiated
  Mat col;  // boundary^T (dif), if trial side differentiated
  bool has_row, has_col;
  PairingTables(int n_, int k_, int kind_)
      : n(n_), k(k_), kind(kind_), mass(n_, k_), has_row(kind_ == DIF_TEST || kind_ == DIF_BOTH),
        has_col(kind_ == DIF_TRIAL || kind_ == DIF_BOTH) {
    if (has_row) row = unit_boundary_operator(n_, k_);
    if (has_col) col = unit_boundary_operator(n_, k_).transpose();
  }
  int test_grade() const { return k - (has_row ? 1 : 0); }
  int trial_grade() const { return k - (has_col ? 1 : 0); }
};

// crates/formoniq/src/operators.rs:201-211 / :27-40 (lumped)
inline Mat pairing_element(const PairingTables& t, const Mat& g) {
  Mat mass = hodge_mass_element(t.mass, g);
  if (t.has_col) mass = gemm(mass, t.col);
  if (t.has_row) mass = gemm(t.row, mass);
  return mass;
}
inline Mat lumped_element(const Mat& g) {
  const int nv = g.r + 1;
  const double v = cell_volume(g) / double(nv);
  Mat m(nv, nv, 0.0);
  for (int i = 0; i < nv; ++i) m(i, i) = v;
  return m;
}

inline void kind_grades(int kind, int k, int& test, int& trial) {
  if (kind == LUMPED) {
    test = trial = 0;
    return;
  }
  test = k - ((kind == DIF_TEST || kind == DIF_BOTH) ? 1 : 0);
  trial = k - ((kind == DIF_TRIAL || kind == DIF_BOTH) ? 1 : 0);
}

// ---------------------------------------------------------------------------
// Topology: simplicial complex from cells (crates/simplicial/src/topology)
// ---------------------------------------------------------------------------

// Colex comparison of vertex words (simplex.rs:126-129): from the largest
// vertex downward.
inline bool colex_less(const int64_t* a, const int64_t* b, int len) {
  for (int i = len - 1; i >= 0; --i) {
    if (a[i] != b[i]) return a[i] < b[i];
  }
  return false;
}

struct Complex {
  int dim = 0;
  // skeleton[j]: flat sorted vertex words of the j-simplices, stride j+1.
  std::vector<std::vector<int64_t>> skeleton;
  // cell_faces[j]: for every cell (colex order) the global ids of its
  // C(n+1,j+1) faces in local colex order (FaceIncidence::faces_flat,
  // crates/simplicial/src/topology/incidence.rs:42-53,60-113).
  std::vector<std::vector<int64_t>> cell_faces;
  idx_t nsimplices(int j) const {
    return (j < 0 || j > dim) ? 0 : idx_t(skeleton[j].size() / size_t(j + 1));
  }
  idx_t ncells() const { return nsimplices(dim); }
};

// Complex::from_cells_unchecked (complex.rs:298-336) + Skeleton::new
// (skeleton.rs:50-86): each skeleton = sort + dedup of all sub-simplices of
// all cells in colex order; cells themselves re-sorted.
inline Complex complex_from_cells(int dim, const std::vector<int64_t>& cells_in) {
  Complex cx;
  cx.dim = dim;
  const int nv = dim + 1;
  const size_t nc_in = cells_in.size() / size_t(nv);
  cx.skeleton.resize(dim + 1);
  cx.cell_faces.resize(dim + 1);
  for (int j = 0; j <= dim; ++j) {
    const int len = j + 1;
    const auto subs = combinations(nv, len);
    std::vector<int64_t> words;
    words.reserve(nc_in * subs.size() * len);
    for (size_t c = 0; c < nc_in; ++c) {
      int64_t sorted[16];
      for (int i = 0; i < nv; ++i) sorted[i] = cells_in[c * nv + i];
      std::sort(sorted, sorted + nv);
      for (const Comb& s : subs)
        for (int i = 0; i < len; ++i) words.push_back(sorted[s[i]]);
    }
    const size_t nw = words.size() / len;
    std::vector<size_t> order(nw);
    std::iota(order.begin(), order.end(), size_t(0));
    std::sort(order.begin(), order.end(), [&](size_t x, size_t y) {
      return colex_less(&words[x * len], &words[y * len], len);
    });
    std::vector<int64_t>& sk = cx.skeleton[j];
    for (size_t o = 0; o < nw; ++o) {
      const int64_t* w = &words[order[o] * len];
      if (!sk.empty() && std::memcmp(&sk[sk.size() - len], w, sizeof(int64_t) * len) == 0) continue;
      sk.insert(sk.end(), w, w + len);
    }
  }
  // local -> global maps by binary search in the sorted skeleton
  const idx_t ncells = cx.ncells();
  const std::vector<int64_t>& cells = cx.skeleton[dim];
  for (int j = 0; j <= dim; ++j) {
    const int len = j + 1;
    const auto subs = combinations(nv, len);
    const std::vector<int64_t>& sk = cx.skeleton[j];
    const idx_t ns = cx.nsimplices(j);
    std::vector<int64_t>& cf = cx.cell_faces[j];
    cf.resize(size_t(ncells) * subs.size());
    for (idx_t c = 0; c < ncells; ++c)
      for (size_t l = 0; l < subs.size(); ++l) {
        int64_t w[16];
        for (int i = 0; i < len; ++i) w[i] = cells[size_t(c) * nv + subs[l][i]];
        idx_t lo = 0, hi = ns;
        while (lo < hi) {
          const idx_t mid = (lo + hi) / 2;
          if (colex_less(&sk[size_t(mid) * len], w, len))
            lo = mid + 1;
          else
            hi = mid;
        }
        assert(lo < ns && std::memcmp(&sk[size_t(lo) * len], w, sizeof(int64_t) * len) == 0);
        cf[size_t(c) * subs.size() + l] = lo;
      }
  }
  return cx;
}

// Kuhn (Freudenthal) triangulation of a box grid with `shape[a]` cells along
// axis a (crates/simplicial/src/mesher/grid.rs:77-103; vertex linearisation
// axis 0 fastest, multiindex/src/cartesian.rs:94-105).
inline std::vector<int64_t> kuhn_cells(int dim, const int64_t* shape) {
  std::vector<int64_t> vstride(dim);
  int64_t acc = 1, nboxes = 1;
  for (int a = 0; a < dim; ++a) {
    vstride[a] = acc;
    acc *= shape[a] + 1;
    nboxes *= shape[a];
  }
  std::vector<int64_t> cells;
  const auto perms = permutations_all(dim);
  std::vector<int64_t> bc(dim);
  for (int64_t ibox = 0; ibox < nboxes; ++ibox) {
    int64_t rem = ibox, origin = 0;
    for (int a = 0; a < dim; ++a) {
      bc[a] = rem % shape[a];
      rem /= shape[a];
      origin += bc[a] * vstride[a];
    }
    for (const Perm& p : perms) {
      int64_t v = origin;
      cells.push_back(v);
      for (int i = 0; i < dim; ++i) {
        v += vstride[p.p[i]];
        cells.push_back(v);
      }
    }
  }
  return cells;
}

// Vertex coordinates of the grid (crates/regge/src/mesher/cartesian.rs:158-169):
// x_a = (c_a / N_a) * side_a + min_a.  Column-major [dim] per vertex.
inline std::vector<double> kuhn_vertex_coords(int dim, const int64_t* shape, const double* min,
                                              const double* max) {
  int64_t nvert = 1;
  for (int a = 0; a < dim; ++a) nvert *= shape[a] + 1;
  std::vector<double> x(size_t(nvert) * dim);
  for (int64_t v = 0; v < nvert; ++v) {
    int64_t rem = v;
    for (int a = 0; a < dim; ++a) {
      const int64_t c = rem % (shape[a] + 1);
      rem /= shape[a] + 1;
      const double frac = double(c) / double(shape[a]);
      x[size_t(v) * dim + a] = frac * (max[a] - min[a]) + min[a];
    }
  }
  return x;
}

// splitmix64-style probe of crates/formoniq/src/linalg/eigen.rs:259-268.
inline double pseudo_random(uint64_t seed, uint64_t index) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + index * 0xD1B54A32D192ED03ull + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return double(z >> 11) / double(1ull << 53) * 2.0 - 1.0;
}

// Edge lengths from coordinates (crates/regge/src/coord/mesh.rs:208-216,
// crates/metric/src/lib.rs:350-361): s = (d^T G) d with d = x[vj]-x[vi],
// vi<vj; `ambient_diag` holds the diagonal ambient form (+1 Euclid, -1 time).
inline std::vector<double> edge_lengths_sq(const Complex& cx, int ambient_dim, const double* coords,
                                           const double* ambient_diag) {
  const idx_t ne = cx.nsimplices(1);
  std::vector<double> s(size_t(ne), 0.0);
  for (idx_t e = 0; e < ne; ++e) {
    const int64_t vi = cx.skeleton[1][size_t(e) * 2], vj = cx.skeleton[1][size_t(e) * 2 + 1];
    // v^T * G: row vector t_j = sum_i (1*d_i)*G_ij  (gemm, only the diagonal
    // term is non-zero; zero terms add +-0 exactly)
    double acc = 0.0;
    for (int a = 0; a < ambient_dim; ++a) {
      const double d = coords[size_t(vj) * ambient_dim + a] - coords[size_t(vi) * ambient_dim + a];
      const double t = d * ambient_diag[a];
      const double term = t * d;
      acc = (a == 0) ? term : term + acc;
    }
    s[size_t(e)] = acc;
  }
  return s;
}

// ---------------------------------------------------------------------------
// A8. Assembly (crates/formoniq/src/galerkin.rs:138-188) and COO->CSR
// (nalgebra-sparse 0.12 convert_coo_csr; duplicates summed in cell order).
// ---------------------------------------------------------------------------
struct Csr {
  idx_t nrows = 0, ncols = 0;
  std::vector<int64_t> row_ptr, col_idx;
  std::vector<double> values;
};

struct Triplets {
  std::vector<int64_t> rows, cols;
  std::vector<double> vals;
};

inline void cell_metric(const Complex& cx, const double* lengths_sq, idx_t cell, Mat& g) {
  const int n = cx.dim;
  const int ne = int(binomial(n + 1, 2));
  double s[128];
  for (int e = 0; e < ne; ++e) s[e] = lengths_sq[cx.cell_faces[1][size_t(cell) * ne + e]];
  g = metric_from_lengths(n, s);
}

// Element matrix of one cell, row-major into out[rows*cols].
inline Mat cell_element(const Complex& cx, const double* lengths_sq, int kind, int k, idx_t cell,
                        const PairingTables* tables) {
  Mat g;
  if (cx.dim >= 1)
    cell_metric(cx, lengths_sq, cell, g);
  else
    g = Mat(0, 0);
  if (kind == LUMPED) return lumped_element(g);
  return pairing_element(*tables, g);
}

// Triplets of cells [c0,c1) in cell order with the `!= 0.0` filter
// (galerkin.rs:160-181).  drop_zeros=false keeps every structural entry.
inline void assemble_triplets(const Complex& cx, const double* lengths_sq, int kind, int k, idx_t c0,
                              idx_t c1, bool drop_zeros, Triplets& out) {
  int tg, rg;
  kind_grades(kind, k, tg, rg);
  const int nt = nlocal(cx.dim, tg), nr = nlocal(cx.dim, rg);
  if (nt == 0 || nr == 0) return;
  PairingTables tables(cx.dim, kind == LUMPED ? 0 : k, kind == LUMPED ? MASS : kind);
  for (idx_t c = c0; c < c1; ++c) {
    const Mat el = cell_element(cx, lengths_sq, kind, k, c, &tables);
    const int64_t* trow = &cx.cell_faces[tg][size_t(c) * nt];
    const int64_t* tcol = &cx.cell_faces[rg][size_t(c) * nr];
    for (int i = 0; i < nt; ++i)
      for (int j = 0; j < nr; ++j) {
        const double v = el(i, j);
        if (!drop_zeros || v != 0.0) {
          out.rows.push_back(trow[i]);
          out.cols.push_back(tcol[j]);
          out.vals.push_back(v);
        }
      }
  }
}

// COO -> CSR: counting sort by row (stable), per-row stable sort by column,
// duplicates combined left to right (value = value + next).
inline Csr coo_to_csr(idx_t nrows, idx_t ncols, const Triplets& t) {
  Csr m;
  m.nrows = nrows;
  m.ncols = ncols;
  const size_t nt = t.rows.size();
  std::vector<int64_t> offs(size_t(nrows) + 1, 0);
  for (size_t i = 0; i < nt; ++i) ++offs[size_t(t.rows[i]) + 1];
  for (idx_t r = 0; r < nrows; ++r) offs[size_t(r) + 1] += offs[size_t(r)];
  std::vector<int64_t> ucol(nt);
  std::vector<double> uval(nt);
  {
    std::vector<int64_t> cur(offs.begin(), offs.end() - 1);
    for (size_t i = 0; i < nt; ++i) {
      const size_t p = size_t(cur[size_t(t.rows[i])]++);
      ucol[p] = t.cols[i];
      uval[p] = t.vals[i];
    }
  }
  m.row_ptr.assign(1, 0);
  std::vector<size_t> perm;
  for (idx_t r = 0; r < nrows; ++r) {
    const size_t b = size_t(offs[size_t(r)]), e = size_t(offs[size_t(r) + 1]);
    perm.resize(e - b);
    std::iota(perm.begin(), perm.end(), b);
    std::stable_sort(perm.begin(), perm.end(), [&](size_t x, size_t y) { return ucol[x] < ucol[y]; });
    size_t i = 0;
    while (i < perm.size()) {
      const int64_t col = ucol[perm[i]];
      double v = uval[perm[i]];
      ++i;
      while (i < perm.size() && ucol[perm[i]] == col) {
        v = v + uval[perm[i]];
        ++i;
      }
      m.col_idx.push_back(col);
      m.values.push_back(v);
    }
    m.row_ptr.push_back(int64_t(m.col_idx.size()));
  }
  return m;
}

inline Csr assemble_matrix(const Complex& cx, const double* lengths_sq, int kind, int k,
                           bool drop_zeros = true) {
  int tg, rg;
  kind_grades(kind, k, tg, rg);
  Triplets t;
  assemble_triplets(cx, lengths_sq, kind, k, 0, cx.ncells(), drop_zeros, t);
  return coo_to_csr(cx.nsimplices(tg), cx.nsimplices(rg), t);
}

// ---------------------------------------------------------------------------
// A10. SpMV and the Krylov drivers of crates/iterative
// ---------------------------------------------------------------------------

// crates/iterative/src/operator.rs:12-14 -> nalgebra-sparse serial CSR*dense.
inline void spmv(const Csr& a, const double* x, double* y) {
  for (idx_t i = 0; i < a.nrows; ++i) {
    double acc = 0.0;
    for (int64_t p = a.row_ptr[size_t(i)]; p < a.row_ptr[size_t(i) + 1]; ++p)
      acc = acc + a.values[size_t(p)] * x[a.col_idx[size_t(p)]];
    y[i] = acc;
  }
}

inline double dot(const std::vector<double>& a, const std::vector<double>& b) {
  double s = 0.0;
  for (size_t i = 0; i < a.size(); ++i) s = s + a[i] * b[i];
  return s;
}
inline void axpy(std::vector<double>& y, double alpha, const std::vector<double>& x) {
  for (size_t i = 0; i < y.size(); ++i) y[i] = alpha * x[i] + y[i];
}
inline void scale(std::vector<double>& y, double alpha) {
  for (double& v : y) v = v * alpha;
}

struct Report {
  int64_t iters = 0;
  double residual = 0.0;
  bool converged = false;
};
using ApplyFn = std::function<void(const std::vector<double>&, std::vector<double>&)>;

// crates/iterative/src/krylov.rs:48-95
inline Report cg(const ApplyFn& op, const ApplyFn& precond, const std::vector<double>& b, double rtol,
                 int64_t max_iters, std::vector<double>& x) {
  const size_t n = b.size();
  x.assign(n, 0.0);
  Report rep;
  const double b_norm = std::sqrt(dot(b, b));
  if (b_norm == 0.0) {
    rep.converged = true;
    return rep;
  }
  std::vector<double> r = b, z(n), p, ap(n);
  precond(r, z);
  p = z;
  double rz = dot(r, z);
  for (;;) {
    rep.residual = std::sqrt(dot(r, r)) / b_norm;
    rep.converged = rep.residual <= rtol;
    if (rep.converged || rep.iters >= max_iters) break;
    op(p, ap);
    const double alpha = rz / dot(p, ap);
    axpy(x, alpha, p);
    axpy(r, -alpha, ap);
    precond(r, z);
    const double rz_next = dot(r, z);
    const double beta = rz_next / rz;
    scale(p, beta);
    axpy(p, 1.0, z);
    rz = rz_next;
    ++rep.iters;
  }
  return rep;
}

// crates/iterative/src/krylov.rs:113-211
inline Report minres(const ApplyFn& op, const ApplyFn& precond, const std::vector<double>& b,
                     double rtol, int64_t max_iters, std::vector<double>& x) {
  const size_t n = b.size();
  Report rep;
  const double eps = 2.220446049250313e-16;
  std::vector<double> r1 = b, y(n);
  precond(r1, y);
  const double beta1_sq = dot(r1, y);
  x.assign(n, 0.0);
  if (beta1_sq <= 0.0) {
    rep.converged = true;
    return rep;
  }
  const double beta1 = std::sqrt(beta1_sq);
  double oldb = 0.0, beta = beta1, dbar = 0.0, epsln = 0.0, phibar = beta1, cs = -1.0, sn = 0.0;
  std::vector<double> w(n, 0.0), w2(n, 0.0), r2 = r1, v, y_next(n);
  rep.residual = 1.0;
  while (rep.iters < max_iters) {
    ++rep.iters;
    v = y;
    scale(v, 1.0 / beta);
    op(v, y_next);
    if (rep.iters >= 2) axpy(y_next, -beta / oldb, r1);
    const double alfa = dot(v, y_next);
    axpy(y_next, -alfa / beta, r2);
    r1 = r2;
    r2 = y_next;
    precond(r2, y);
    oldb = beta;
    beta = std::sqrt(std::max(dot(r2, y), 0.0));
    const double oldeps = epsln;
    const double delta = cs * dbar + sn * alfa;
    const double gbar = sn * dbar - cs * alfa;
    epsln = sn * beta;
    dbar = -cs * beta;
    const double gamma = std::max(std::sqrt(gbar * gbar + beta * beta), eps);
    cs = gbar / gamma;
    sn = beta / gamma;
    const double phi = cs * phibar;
    phibar *= sn;
    std::vector<double> wnew = v;
    axpy(wnew, -oldeps, w2);
    axpy(wnew, -delta, w);
    scale(wnew, 1.0 / gamma);
    w2 = w;
    w = wnew;
    axpy(x, phi, w);
    rep.residual = phibar / beta1;
    if (rep.residual <= rtol) {
      rep.converged = true;
      break;
    }
  }
  return rep;
}

}  // namespace fqo
