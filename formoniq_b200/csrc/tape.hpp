// tape.hpp — host-side "element tape" compiler (product code, no CUDA needed).
//
// For a runtime (dim n, grade k, operator kind) this builds the straight-line
// FP64 program that maps one cell's signed squared edge lengths to its
// Whitney element matrix, in exactly the operation order of the reference:
//   regge/src/lengths/simplex.rs:308-326   metric by polarisation
//   metric/src/lib.rs:250-257              g^-1 (nalgebra closed forms n<=3)
//   multialgebra/src/lib.rs:285-305,447-456  Lambda^k g^-1 by Leibniz minors
//   formoniq/src/operators.rs:84-94        H = DP*FG*DP^T, M = vol * pullback
//   derham/src/interpolate/form.rs:222-232 pullback C^T (H (x) Q) C
//   formoniq/src/operators.rs:201-211      sandwiches  d * M * D
// Operations that are exact in IEEE-754 (multiplication by 0, +-1 and powers of
// two, addition of an exact zero, negation) are evaluated symbolically, so the
// emitted program contains only the roundings the reference performs and is
// bit-identical to it up to the sign of exact zeros.
//
// The same tape feeds (a) the generic interpreter kernel (any n, k at run
// time) and (b) the build-time generator that prints it as straight-line CUDA
// for the hot (n,k) combinations (gen_elmat.cpp -> elmat_gen.cuh).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace fq {

enum Kind : int { KIND_MASS = 0, KIND_DIF_TRIAL = 1, KIND_DIF_TEST = 2, KIND_DIF_BOTH = 3, KIND_LUMPED = 4 };

// ---------------------------------------------------------------- combinatorics
inline int64_t binom(int n, int k) {
  if (k < 0 || n < 0 || k > n) return 0;
  int64_t r = 1;
  for (int i = 0; i < k; ++i) r = r * (n - i) / (i + 1);
  return r;
}
inline int64_t fact(int n) {
  int64_t f = 1;
  for (int i = 2; i <= n; ++i) f *= i;
  return f;
}
// Colex-ordered `card`-subsets of {0..n-1} as bitmasks (increasing integer
// order of equal-popcount masks is colex order; Gosper's hack).
inline std::vector<uint32_t> colex_subsets(int n, int card) {
  std::vector<uint32_t> out;
  if (card < 0 || card > n) return out;
  if (card == 0) {
    out.push_back(0);
    return out;
  }
  uint64_t m = (1ull << card) - 1, lim = 1ull << n;
  while (m < lim) {
    out.push_back(uint32_t(m));
    const uint64_t c = m & (~m + 1), r = m + c;
    m = (((r ^ m) >> 2) / c) | r;
  }
  return out;
}
inline std::vector<int> mask_elems(uint32_t m) {
  std::vector<int> e;
  for (int i = 0; i < 32; ++i)
    if (m >> i & 1) e.push_back(i);
  return e;
}
// colex rank of a subset = sum_i C(s_i, i+1)
inline int64_t colex_rank(uint32_t m) {
  int64_t r = 0;
  int i = 0;
  for (int b = 0; b < 32; ++b)
    if (m >> b & 1) r += binom(b, ++i);
  return r;
}
inline int nlocal(int n, int j) { return (j < 0 || j > n) ? 0 : int(binom(n + 1, j + 1)); }
inline int edge_slot(int i, int j) { return int(binom(i < j ? i : j, 1) + binom(i < j ? j : i, 2)); }

struct SignedPerm {
  std::vector<int> p;
  int sign;
};
// S_m in the reference's colex order: the reversed words of the lexicographic
// enumeration (multiindex/src/permutation.rs:160-181).
inline std::vector<SignedPerm> perms_colex(int m) {
  std::vector<SignedPerm> out;
  std::vector<int> w(m);
  for (int i = 0; i < m; ++i) w[i] = i;
  for (;;) {
    SignedPerm sp;
    sp.p.assign(w.rbegin(), w.rend());
    int inv = 0;
    for (int a = 0; a < m; ++a)
      for (int b = a + 1; b < m; ++b) inv += sp.p[a] > sp.p[b];
    sp.sign = (inv & 1) ? -1 : 1;
    out.push_back(sp);
    if (!std::next_permutation(w.begin(), w.end())) break;
  }
  return out;
}

// ---------------------------------------------------------------- tape
enum TapeOpcode : uint8_t {
  OP_ADD = 0,   // r[d] = r[a] + r[b]
  OP_SUB = 1,   // r[d] = r[a] - r[b]
  OP_MUL = 2,   // r[d] = r[a] * r[b]
  OP_MULC = 3,  // r[d] = r[a] * consts[b]
  OP_DIV = 4,   // r[d] = r[a] / r[b]
  OP_SQRTABS = 5,  // r[d] = sqrt(|r[a]|)
  OP_LOADC = 6,    // r[d] = consts[b]
  OP_STORE = 7,    // out[d] = r[a]
  OP_STOREN = 8,   // out[d] = -r[a]
  OP_STOREC = 9,   // out[d] = consts[b]
};
struct TapeOp {
  uint8_t op;
  uint32_t d, a, b;
};

// A finished program.  Inputs occupy registers [0, ninputs).
struct Tape {
  int n = 0, k = 0;
  int ninputs = 0;   // C(n+1,2) lengths (n<=3) or n*n+1 (ginv row-major, vol) for n>=4
  bool inputs_are_lengths = true;
  int nregs = 0;     // after register allocation
  int nouts = 0;
  std::vector<TapeOp> ops;
  std::vector<double> consts;
  // statistics (FP64 instruction classes actually emitted)
  int n_addsub = 0, n_mul = 0, n_div = 0, n_sqrt = 0;
};

// Symbolic value: sign * 2^exp * (register | constant).
struct Val {
  bool is_const = true;
  double c = 0.0;
  int reg = -1;
  int sign = 1;
  int exp = 0;
};

class TapeBuilder {
 public:
  std::vector<TapeOp> ops;  // SSA form: d is a fresh id for value ops
  std::vector<double> consts;
  int next_reg = 0;
  int nouts = 0;

  static Val constant(double c) {
    Val v;
    v.is_const = true;
    v.c = c;
    return v;
  }
  Val input() {
    Val v;
    v.is_const = false;
    v.reg = next_reg++;
    return v;
  }
  static bool is_zero(const Val& v) { return v.is_const && v.c == 0.0; }
  static bool pow2(double c, int& e) {
    if (c == 0.0 || !std::isfinite(c)) return false;
    int ex;
    const double m = std::frexp(std::fabs(c), &ex);
    if (m != 0.5) return false;
    e = ex - 1;
    return true;
  }
  static Val neg(Val v) {
    if (v.is_const)
      v.c = -v.c;
    else
      v.sign = -v.sign;
    return v;
  }

  int const_index(double c) {
    for (size_t i = 0; i < consts.size(); ++i)
      if (std::memcmp(&consts[i], &c, sizeof(double)) == 0) return int(i);
    consts.push_back(c);
    return int(consts.size() - 1);
  }
  // hash-consed emission
  int emit(uint8_t op, int a, int b) {
    if ((op == OP_ADD || op == OP_MUL) && a > b) std::swap(a, b);
    const auto key = std::make_tuple(op, a, b);
    auto it = cse_.find(key);
    if (it != cse_.end()) return it->second;
    const int d = next_reg++;
    ops.push_back(TapeOp{op, uint32_t(d), uint32_t(a), uint32_t(b)});
    cse_[key] = d;
    return d;
  }
  // Materialise the 2^exp factor (exact) so that exp == 0.
  Val flat(Val v) {
    if (v.is_const || v.exp == 0) return v;
    const int r = emit(OP_MULC, v.reg, const_index(std::ldexp(1.0, v.exp)));
    v.reg = r;
    v.exp = 0;
    return v;
  }
  Val to_reg(Val v) {
    if (!v.is_const) return v;
    Val r;
    r.is_const = false;
    r.sign = v.c < 0 || (v.c == 0 && std::signbit(v.c)) ? -1 : 1;
    r.reg = emit(OP_LOADC, 0, const_index(std::fabs(v.c)));
    return r;
  }

  Val add(Val a, Val b) {
    if (a.is_const && b.is_const) return constant(a.c + b.c);
    if (is_zero(a)) return b;
    if (is_zero(b)) return a;
    if (a.is_const) a = to_reg(a);
    if (b.is_const) b = to_reg(b);
    if (a.exp != b.exp) {
      a = flat(a);
      b = flat(b);
    }
    Val r;
    r.is_const = false;
    r.exp = a.exp;
    if (a.sign == b.sign) {
      r.reg = emit(OP_ADD, a.reg, b.reg);
      r.sign = a.sign;
    } else {
      if (a.reg == b.reg) return constant(0.0);  // x - x
      // canonical orientation so that x-y and -(y-x) share one register
      const int lo = std::min(a.reg, b.reg), hi = std::max(a.reg, b.reg);
      r.reg = emit(OP_SUB, lo, hi);
      const int sign_lo = (a.reg == lo) ? a.sign : b.sign;
      r.sign = sign_lo;  // (+lo) + (-hi) = lo - hi ; (-lo) + (+hi) = -(lo - hi)
    }
    return r;
  }
  Val sub(Val a, Val b) { return add(a, neg(b)); }
  Val mul(Val a, Val b) {
    if (a.is_const && b.is_const) return constant(a.c * b.c);
    if (!a.is_const && b.is_const) std::swap(a, b);
    if (a.is_const) {
      if (a.c == 0.0) return constant(0.0);
      int e;
      Val r = b;
      if (a.c < 0) r.sign = -r.sign;
      if (pow2(a.c, e)) {
        r.exp += e;
        return r;
      }
      r.reg = emit(OP_MULC, b.reg, const_index(std::fabs(a.c)));
      return r;
    }
    Val r;
    r.is_const = false;
    r.reg = emit(OP_MUL, a.reg, b.reg);
    r.sign = a.sign * b.sign;
    r.exp = a.exp + b.exp;
    return r;
  }
  Val div(Val a, Val b) {
    if (a.is_const && b.is_const) return constant(a.c / b.c);
    if (is_zero(a)) return constant(0.0);
    if (b.is_const) {
      int e;
      if (pow2(b.c, e)) {  // exact
        Val r = a;
        if (b.c < 0) r.sign = -r.sign;
        r.exp -= e;
        return r;
      }
      b = to_reg(b);
    }
    if (a.is_const) {
      int e = 0;
      if (pow2(a.c, e)) {  // (+-2^e)/x = +-2^e * (1/x)
        Val one = to_reg(constant(1.0));
        Val r;
        r.is_const = false;
        r.reg = emit(OP_DIV, one.reg, b.reg);
        r.sign = (a.c < 0 ? -1 : 1) * b.sign;
        r.exp = e - b.exp;
        return r;
      }
      a = to_reg(a);
    }
    Val r;
    r.is_const = false;
    r.reg = emit(OP_DIV, a.reg, b.reg);
    r.sign = a.sign * b.sign;
    r.exp = a.exp - b.exp;
    return r;
  }
  Val sqrt_abs(Val a) {
    if (a.is_const) return constant(std::sqrt(std::fabs(a.c)));
    if (a.exp & 1) a = flat(a);
    Val r;
    r.is_const = false;
    r.reg = emit(OP_SQRTABS, a.reg, 0);
    r.sign = 1;
    r.exp = a.exp / 2;
    return r;
  }
  void store(int out, Val v) {
    if (v.is_const) {
      ops.push_back(TapeOp{OP_STOREC, uint32_t(out), 0, uint32_t(const_index(v.c))});
    } else {
      v = flat(v);
      ops.push_back(TapeOp{uint8_t(v.sign < 0 ? OP_STOREN : OP_STORE), uint32_t(out), uint32_t(v.reg), 0});
    }
    nouts = std::max(nouts, out + 1);
  }

  // Dead-code elimination + linear-scan register allocation.
  Tape finish(int n, int k, int ninputs, bool inputs_are_lengths) {
    const int nssa = next_reg;
    std::vector<char> live(size_t(nssa), 0);
    auto is_store = [](uint8_t op) { return op == OP_STORE || op == OP_STOREN || op == OP_STOREC; };
    auto uses_b = [](uint8_t op) { return op == OP_ADD || op == OP_SUB || op == OP_MUL || op == OP_DIV; };
    auto uses_a = [](uint8_t op) { return op != OP_LOADC && op != OP_STOREC; };
    for (size_t i = ops.size(); i-- > 0;) {
      const TapeOp& o = ops[i];
      const bool keep = is_store(o.op) || live[o.d];
      if (!keep) continue;
      if (uses_a(o.op)) live[o.a] = 1;
      if (uses_b(o.op)) live[o.b] = 1;
    }
    std::vector<TapeOp> kept;
    for (const TapeOp& o : ops)
      if (is_store(o.op) || live[o.d]) kept.push_back(o);
    // last use
    std::vector<int> last(size_t(nssa), -1);
    for (size_t i = 0; i < kept.size(); ++i) {
      const TapeOp& o = kept[i];
      if (uses_a(o.op)) last[o.a] = int(i);
      if (uses_b(o.op)) last[o.b] = int(i);
    }
    Tape t;
    t.n = n;
    t.k = k;
    t.ninputs = ninputs;
    t.inputs_are_lengths = inputs_are_lengths;
    t.consts = consts;
    t.nouts = nouts;
    std::vector<int> phys(size_t(nssa), -1);
    std::vector<int> free_regs;
    int nphys = ninputs;
    for (int i = 0; i < ninputs; ++i) phys[i] = i;
    for (size_t i = 0; i < kept.size(); ++i) {
      TapeOp o = kept[i];
      const uint32_t sa = o.a, sb = o.b;
      if (uses_a(o.op)) o.a = uint32_t(phys[sa]);
      if (uses_b(o.op)) o.b = uint32_t(phys[sb]);
      // release operands whose last use is here (inputs included)
      if (uses_a(o.op) && last[sa] == int(i)) free_regs.push_back(phys[sa]);
      if (uses_b(o.op) && sb != sa && last[sb] == int(i)) free_regs.push_back(phys[sb]);
      if (!is_store(o.op)) {
        int r;
        if (!free_regs.empty()) {
          r = free_regs.back();
          free_regs.pop_back();
        } else {
          r = nphys++;
        }
        phys[o.d] = r;
        o.d = uint32_t(r);
        if (last[kept[i].d] < 0) free_regs.push_back(r);  // never read (cannot happen after DCE)
      }
      switch (o.op) {
        case OP_ADD: case OP_SUB: ++t.n_addsub; break;
        case OP_MUL: case OP_MULC: ++t.n_mul; break;
        case OP_DIV: ++t.n_div; break;
        case OP_SQRTABS: ++t.n_sqrt; break;
        default: break;
      }
      t.ops.push_back(o);
    }
    t.nregs = nphys;
    ssa_ops_ = kept;
    return t;
  }
  // SSA ops after DCE (for the code generator).
  const std::vector<TapeOp>& ssa_ops() const { return ssa_ops_; }

 private:
  std::map<std::tuple<uint8_t, int, int>, int> cse_;
  std::vector<TapeOp> ssa_ops_;
};

// ---------------------------------------------------------------- symbolic matrices
struct SMat {
  int r = 0, c = 0;
  std::vector<Val> a;
  SMat() = default;
  SMat(int r_, int c_) : r(r_), c(c_), a(size_t(r_) * c_, TapeBuilder::constant(0.0)) {}
  Val& operator()(int i, int j) { return a[size_t(i) * c + j]; }
  const Val& operator()(int i, int j) const { return a[size_t(i) * c + j]; }
  SMat transpose() const {
    SMat t(c, r);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < c; ++j) t(j, i) = (*this)(i, j);
    return t;
  }
};

// nalgebra small-gemm order: C_ij = (1*A_i0)*B_0j, then ((1*A_ik)*B_kj) + 1*C_ij.
inline SMat sgemm(TapeBuilder& tb, const SMat& A, const SMat& B) {
  SMat C(A.r, B.c);
  const Val one = TapeBuilder::constant(1.0);
  for (int j = 0; j < B.c; ++j)
    for (int kk = 0; kk < A.c; ++kk)
      for (int i = 0; i < A.r; ++i) {
        const Val term = tb.mul(tb.mul(one, A(i, kk)), B(kk, j));
        C(i, j) = (kk == 0) ? term : tb.add(term, tb.mul(one, C(i, j)));
      }
  return C;
}
// Leibniz determinant in the reference's permutation order.
inline Val sdet(TapeBuilder& tb, const SMat& m) {
  Val acc = TapeBuilder::constant(0.0);
  for (const SignedPerm& s : perms_colex(m.r)) {
    Val prod = TapeBuilder::constant(1.0);
    for (int i = 0; i < m.r; ++i) prod = tb.mul(prod, m(i, s.p[i]));
    acc = tb.add(acc, tb.mul(TapeBuilder::constant(double(s.sign)), prod));
  }
  return acc;
}
// k-th compound matrix on colex subsets.
inline SMat scompound(TapeBuilder& tb, const SMat& m, int k) {
  const auto rows = colex_subsets(m.r, k), cols = colex_subsets(m.c, k);
  SMat out(int(rows.size()), int(cols.size()));
  for (size_t i = 0; i < rows.size(); ++i)
    for (size_t j = 0; j < cols.size(); ++j) {
      const auto ri = mask_elems(rows[i]), cj = mask_elems(cols[j]);
      SMat minor(k, k);
      for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b) minor(a, b) = m(ri[a], cj[b]);
      out(int(i), int(j)) = sdet(tb, minor);
    }
  return out;
}
inline SMat const_difbarys(int n) {
  SMat d(n + 1, n);
  for (int j = 0; j < n; ++j) d(0, j) = TapeBuilder::constant(-1.0);
  for (int i = 0; i < n; ++i) d(i + 1, i) = TapeBuilder::constant(1.0);
  return d;
}
// boundary operator of the reference cell: rows = (k-1)-faces, cols = k-faces.
inline SMat const_boundary(int n, int k) {
  SMat b(nlocal(n, k - 1), nlocal(n, k));
  if (b.r == 0 || b.c == 0) return b;
  const auto cof = colex_subsets(n + 1, k + 1);
  for (size_t ic = 0; ic < cof.size(); ++ic) {
    const auto el = mask_elems(cof[ic]);
    for (size_t pos = 0; pos < el.size(); ++pos) {
      const uint32_t face = cof[ic] & ~(1u << el[pos]);
      b(int(colex_rank(face)), int(ic)) = TapeBuilder::constant((pos & 1) ? -1.0 : 1.0);
    }
  }
  return b;
}

struct Geometry {
  SMat ginv;
  Val vol;
};

// Geometry stage for n <= 3 from the cell's edge lengths (tape inputs).
inline Geometry build_geometry_from_lengths(TapeBuilder& tb, int n, const std::vector<Val>& s) {
  Geometry geo;
  SMat g(n, n);
  for (int i = 0; i < n; ++i) g(i, i) = s[edge_slot(0, i + 1)];
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      const Val v = tb.mul(TapeBuilder::constant(0.5),
                           tb.sub(tb.add(s[edge_slot(0, i + 1)], s[edge_slot(0, j + 1)]), s[edge_slot(i + 1, j + 1)]));
      g(i, j) = v;
      g(j, i) = v;
    }
  Val det;
  geo.ginv = SMat(n, n);
  if (n == 0) {
    det = TapeBuilder::constant(1.0);
  } else if (n == 1) {
    det = g(0, 0);
    geo.ginv(0, 0) = tb.div(TapeBuilder::constant(1.0), g(0, 0));
  } else if (n == 2) {
    det = tb.sub(tb.mul(g(0, 0), g(1, 1)), tb.mul(g(1, 0), g(0, 1)));
    geo.ginv(0, 0) = tb.div(g(1, 1), det);
    geo.ginv(0, 1) = tb.div(TapeBuilder::neg(g(0, 1)), det);
    geo.ginv(1, 0) = tb.div(TapeBuilder::neg(g(1, 0)), det);
    geo.ginv(1, 1) = tb.div(g(0, 0), det);
  } else if (n == 3) {
    const Val m11 = g(0, 0), m12 = g(0, 1), m13 = g(0, 2), m21 = g(1, 0), m22 = g(1, 1), m23 = g(1, 2),
              m31 = g(2, 0), m32 = g(2, 1), m33 = g(2, 2);
    const Val mi1 = tb.sub(tb.mul(m22, m33), tb.mul(m32, m23));
    const Val mi2 = tb.sub(tb.mul(m21, m33), tb.mul(m31, m23));
    const Val mi3 = tb.sub(tb.mul(m21, m32), tb.mul(m31, m22));
    det = tb.add(tb.sub(tb.mul(m11, mi1), tb.mul(m12, mi2)), tb.mul(m13, mi3));
    geo.ginv(0, 0) = tb.div(mi1, det);
    geo.ginv(0, 1) = tb.div(tb.sub(tb.mul(m13, m32), tb.mul(m33, m12)), det);
    geo.ginv(0, 2) = tb.div(tb.sub(tb.mul(m12, m23), tb.mul(m22, m13)), det);
    geo.ginv(1, 0) = tb.div(TapeBuilder::neg(mi2), det);
    geo.ginv(1, 1) = tb.div(tb.sub(tb.mul(m11, m33), tb.mul(m31, m13)), det);
    geo.ginv(1, 2) = tb.div(tb.sub(tb.mul(m13, m21), tb.mul(m23, m11)), det);
    geo.ginv(2, 0) = tb.div(mi3, det);
    geo.ginv(2, 1) = tb.div(tb.sub(tb.mul(m12, m31), tb.mul(m32, m11)), det);
    geo.ginv(2, 2) = tb.div(tb.sub(tb.mul(m11, m22), tb.mul(m21, m12)), det);
  } else {
    throw std::runtime_error("build_geometry_from_lengths: n <= 3 only");
  }
  geo.vol = tb.mul(TapeBuilder::constant(1.0 / double(fact(n))), tb.sqrt_abs(det));
  return geo;
}

// Whitney mass of grade k (operators.rs:84-94).
inline SMat build_mass(TapeBuilder& tb, int n, int k, const Geometry& geo) {
  const int nv = n + 1;
  // FG = Lambda^k g^-1
  const SMat FG = scompound(tb, geo.ginv, k);
  // DP = Lambda^k(difbarys): exact constants
  const SMat DP = scompound(tb, const_difbarys(n), k);
  const SMat H = sgemm(tb, sgemm(tb, DP, FG), DP.transpose());
  // Q
  const double qs = 1.0 / double(nv * (nv + 1));
  const auto dofs = colex_subsets(nv, k + 1);
  const double kf = double(fact(k));
  struct Term {
    double coef;
    int blade, vertex;
  };
  std::vector<std::vector<Term>> cols;
  for (uint32_t d : dofs) {
    std::vector<Term> col;
    const auto el = mask_elems(d);
    for (size_t pos = 0; pos < el.size(); ++pos)
      col.push_back(Term{((pos & 1) ? -1.0 : 1.0) * kf, int(colex_rank(d & ~(1u << el[pos]))), el[pos]});
    cols.push_back(col);
  }
  const int nd = int(dofs.size());
  SMat M(nd, nd);
  for (int i = 0; i < nd; ++i)
    for (int j = 0; j < nd; ++j) {
      Val acc = TapeBuilder::constant(0.0);
      for (const Term& a : cols[size_t(i)])
        for (const Term& b : cols[size_t(j)]) {
          const double q = (a.vertex == b.vertex) ? 2.0 * qs : qs;
          const Val term = tb.mul(tb.mul(TapeBuilder::constant(a.coef * b.coef), H(a.blade, b.blade)),
                                  TapeBuilder::constant(q));
          acc = tb.add(acc, term);
        }
      // vol * acc; the exact 2^e factor of acc is folded into vol once so the
      // product needs no per-entry rescale (bitwise the same value).
      if (acc.is_const) {
        M(i, j) = tb.mul(geo.vol, acc);
      } else {
        Val v = geo.vol;
        v.exp += acc.exp;
        acc.exp = 0;
        M(i, j) = tb.mul(tb.flat(v), acc);
      }
    }
  return M;
}

inline void kind_grades(int kind, int k, int& test, int& trial) {
  if (kind == KIND_LUMPED) {
    test = trial = 0;
    return;
  }
  test = k - ((kind == KIND_DIF_TEST || kind == KIND_DIF_BOTH) ? 1 : 0);
  trial = k - ((kind == KIND_DIF_TRIAL || kind == KIND_DIF_BOTH) ? 1 : 0);
}

// One requested block of a (possibly fused) tape.
struct BlockSpec {
  int kind, grade;
};
struct BlockLayout {
  int kind, grade, rows, cols, out_offset;
};

// Build the tape for a list of blocks on an n-cell; outputs are the blocks'
// row-major element matrices concatenated.  Blocks share the geometry and any
// common mass (hash-consing), which is what fuses HodgeBlocks (hodge.rs:62-72).
inline Tape build_tape(int n, const std::vector<BlockSpec>& blocks, std::vector<BlockLayout>* layout,
                       TapeBuilder* keep_builder = nullptr) {
  TapeBuilder local;
  TapeBuilder& tb = keep_builder ? *keep_builder : local;
  Geometry geo;
  int ninputs;
  bool from_lengths = n <= 3;
  if (from_lengths) {
    std::vector<Val> s;
    ninputs = int(binom(n + 1, 2));
    for (int e = 0; e < ninputs; ++e) s.push_back(tb.input());
    geo = build_geometry_from_lengths(tb, n, s);
  } else {
    ninputs = n * n + 1;
    geo.ginv = SMat(n, n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) geo.ginv(i, j) = tb.input();
    geo.vol = tb.input();
  }
  std::map<int, SMat> masses;
  auto mass_of = [&](int k) -> const SMat& {
    auto it = masses.find(k);
    if (it == masses.end()) it = masses.emplace(k, build_mass(tb, n, k, geo)).first;
    return it->second;
  };
  int off = 0;
  for (const BlockSpec& b : blocks) {
    int tg, rg;
    kind_grades(b.kind, b.grade, tg, rg);
    SMat el;
    if (b.kind == KIND_LUMPED) {
      const int nv = n + 1;
      el = SMat(nv, nv);
      const Val v = tb.div(geo.vol, TapeBuilder::constant(double(nv)));
      for (int i = 0; i < nv; ++i) el(i, i) = v;
    } else {
      const int rows = nlocal(n, tg), cols = nlocal(n, rg);
      if (rows == 0 || cols == 0 || b.grade < 0 || b.grade > n) {
        el = SMat(rows, cols);
      } else {
        el = mass_of(b.grade);
        if (b.kind == KIND_DIF_TRIAL || b.kind == KIND_DIF_BOTH)
          el = sgemm(tb, el, const_boundary(n, b.grade).transpose());
        if (b.kind == KIND_DIF_TEST || b.kind == KIND_DIF_BOTH) el = sgemm(tb, const_boundary(n, b.grade), el);
      }
    }
    if (layout) layout->push_back(BlockLayout{b.kind, b.grade, el.r, el.c, off});
    for (int i = 0; i < el.r; ++i)
      for (int j = 0; j < el.c; ++j) tb.store(off + i * el.c + j, el(i, j));
    off += el.r * el.c;
  }
  return tb.finish(n, blocks.empty() ? 0 : blocks[0].grade, ninputs, from_lengths);
}

// The four blocks of a mixed problem posed at grade k (hodge.rs:62-72).
inline std::vector<BlockSpec> hodge_blocks(int k) {
  return {{KIND_MASS, k - 1}, {KIND_MASS, k}, {KIND_DIF_TEST, k}, {KIND_DIF_BOTH, k + 1}};
}

// ---------------------------------------------------------------- block sets of the tile-fused kernel (tile.cu)
// The tile kernel stores, for every (cell, owned local row) of a block, the DISTINCT values of that row of the
// block's element matrix ("column slots"): entries of one row that the tape proves to be the same value (same
// SSA register, same sign) share a slot, exact zeros have none.  Blocks are evaluated in stage groups: one group
// per mass grade the blocks derive from (mass(g), dif_*(g) share M_g), in block order.
struct SetBlock {
  int kind = 0, grade = 0, tg = 0, rg = 0, rows = 0, cols = 0, out_offset = 0;
  int group = -1;        // stage group, -1 for an empty block (a grade off [0, n]: the zero space)
  int tclass = -1;       // row class: index of the test grade among the set's distinct test grades
  int d = 0;             // column slots per row (max over the rows)
  std::vector<int> cs;   // [rows*cols] column slot of every entry, -1 = exact zero
};
struct SetPut {          // one generated store: value of `op` (index into the SSA ops) -> (block, row, column slot)
  int op, block, row, slot;
};
struct SetLayout {
  int n = 0, ninputs = 0, ngroups = 0, nclasses = 0;
  int class_grade[4] = {-1, -1, -1, -1};
  std::vector<SetBlock> blocks;
  std::vector<SetPut> puts;   // in tape order
};
inline SetLayout set_layout(int n, const std::vector<BlockSpec>& specs, TapeBuilder* keep_builder = nullptr,
                            Tape* keep_tape = nullptr) {
  TapeBuilder local;
  TapeBuilder& tb = keep_builder ? *keep_builder : local;
  std::vector<BlockLayout> layout;
  const Tape t = build_tape(n, specs, &layout, &tb);
  if (keep_tape) *keep_tape = t;
  SetLayout L;
  L.n = n;
  L.ninputs = t.ninputs;
  std::vector<int> group_grade;
  for (const BlockLayout& bl : layout) {
    SetBlock b;
    b.kind = bl.kind, b.grade = bl.grade, b.rows = bl.rows, b.cols = bl.cols, b.out_offset = bl.out_offset;
    kind_grades(bl.kind, bl.grade, b.tg, b.rg);
    b.cs.assign(size_t(b.rows) * size_t(b.cols), -1);
    if (b.rows > 0 && b.cols > 0) {
      int g = -1;
      for (size_t i = 0; i < group_grade.size(); ++i)
        if (group_grade[i] == bl.grade && bl.kind != KIND_LUMPED) g = int(i);
      if (g < 0) {
        group_grade.push_back(bl.kind == KIND_LUMPED ? -1000 : bl.grade);
        g = int(group_grade.size()) - 1;
      }
      b.group = g;
      for (int c = 0; c < L.nclasses; ++c)
        if (L.class_grade[c] == b.tg) b.tclass = c;
      if (b.tclass < 0) {
        if (L.nclasses == 4) throw std::runtime_error("set_layout: too many test grades");
        L.class_grade[L.nclasses] = b.tg;
        b.tclass = L.nclasses++;
      }
    }
    L.blocks.push_back(b);
  }
  L.ngroups = int(group_grade.size());
  // value identity of a store: (opcode class, register / constant index)
  std::vector<std::map<std::tuple<int, uint32_t>, int>> slots;  // per (block, row)
  std::vector<int> row_base(L.blocks.size() + 1, 0);
  for (size_t b = 0; b < L.blocks.size(); ++b) row_base[b + 1] = row_base[b] + L.blocks[b].rows;
  slots.resize(size_t(row_base.back()));
  const std::vector<TapeOp>& ops = tb.ssa_ops();
  for (size_t i = 0; i < ops.size(); ++i) {
    const TapeOp& o = ops[i];
    if (o.op != OP_STORE && o.op != OP_STOREN && o.op != OP_STOREC) continue;
    int blk = -1;
    for (size_t b = 0; b < L.blocks.size(); ++b)
      if (int(o.d) >= L.blocks[b].out_offset && int(o.d) < L.blocks[b].out_offset + L.blocks[b].rows * L.blocks[b].cols)
        blk = int(b);
    if (blk < 0) throw std::runtime_error("set_layout: store outside every block");
    SetBlock& B = L.blocks[size_t(blk)];
    const int e = int(o.d) - B.out_offset, r = e / B.cols;
    if (o.op == OP_STOREC && tb.consts[o.b] == 0.0) continue;  // exact zero: no slot
    const auto key = std::make_tuple(int(o.op), o.op == OP_STOREC ? o.b : o.a);
    auto& m = slots[size_t(row_base[size_t(blk)] + r)];
    auto it = m.find(key);
    if (it == m.end()) {
      it = m.emplace(key, int(m.size())).first;
      L.puts.push_back(SetPut{int(i), blk, r, it->second});
    }
    B.cs[size_t(e)] = it->second;
    B.d = std::max(B.d, it->second + 1);
  }
  return L;
}

}  // namespace fq
