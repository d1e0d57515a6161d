// tape_host_check.cpp — CPU-side check of the product's tape compiler
// (formoniq_b200/csrc/tape.hpp) against the oracle (oracle/fq_oracle.hpp).
// TEST ONLY: interprets the tape on the host, which the product never does.
// Prints "OK <ncases>" or the first mismatch; exit code 0/1.
#include "../../formoniq_b200/csrc/tape.hpp"
#include "../../oracle/fq_oracle.hpp"

#include <cinttypes>
#include <random>

static std::vector<double> run_tape(const fq::Tape& t, const std::vector<double>& in) {
  std::vector<double> r(size_t(t.nregs), 0.0), out(size_t(t.nouts), 0.0);
  for (int i = 0; i < t.ninputs; ++i) r[i] = in[i];
  for (const fq::TapeOp& o : t.ops) {
    switch (o.op) {
      case fq::OP_ADD: r[o.d] = r[o.a] + r[o.b]; break;
      case fq::OP_SUB: r[o.d] = r[o.a] - r[o.b]; break;
      case fq::OP_MUL: r[o.d] = r[o.a] * r[o.b]; break;
      case fq::OP_MULC: r[o.d] = r[o.a] * t.consts[o.b]; break;
      case fq::OP_DIV: r[o.d] = r[o.a] / r[o.b]; break;
      case fq::OP_SQRTABS: r[o.d] = std::sqrt(std::fabs(r[o.a])); break;
      case fq::OP_LOADC: r[o.d] = t.consts[o.b]; break;
      case fq::OP_STORE: out[o.d] = r[o.a]; break;
      case fq::OP_STOREN: out[o.d] = -r[o.a]; break;
      case fq::OP_STOREC: out[o.d] = t.consts[o.b]; break;
    }
  }
  return out;
}

static std::vector<double> random_lengths(int n, std::mt19937_64& rng, bool lorentz) {
  std::normal_distribution<double> nd;
  for (;;) {
    std::vector<double> x(size_t(n + 1) * n);
    for (double& v : x) v = nd(rng);
    std::vector<double> s(size_t(fqo::binomial(n + 1, 2)));
    bool bad = false;
    for (int j = 1; j <= n; ++j)
      for (int i = 0; i < j; ++i) {
        double acc = 0;
        for (int a = 0; a < n; ++a) {
          const double d = x[size_t(j) * n + a] - x[size_t(i) * n + a];
          acc += ((lorentz && a == 0) ? -1.0 : 1.0) * d * d;
        }
        if (std::fabs(acc) < 1e-2) bad = true;
        s[size_t(fqo::edge_index(i, j))] = acc;
      }
    if (bad) continue;
    if (n >= 1 && fqo::cell_volume(fqo::metric_from_lengths(n, s.data())) < 1e-2) continue;
    return s;
  }
}
// Kuhn-cell lengths with dyadic h (exact zeros in the mass)
static std::vector<double> kuhn_lengths(int n, double h) {
  std::vector<double> s(size_t(fqo::binomial(n + 1, 2)));
  for (int j = 1; j <= n; ++j)
    for (int i = 0; i < j; ++i) s[size_t(fqo::edge_index(i, j))] = double(j - i) * h * h;
  return s;
}

int main() {
  std::mt19937_64 rng(12345);
  int64_t ncases = 0;
  for (int n = 0; n <= 6; ++n) {
    std::vector<std::vector<fq::BlockSpec>> specs;
    for (int k = 0; k <= n + 1; ++k)
      for (int kind = 0; kind < 4; ++kind) {
        if (k == 0 && kind != fq::KIND_MASS) continue;
        if (k == n + 1 && kind != fq::KIND_DIF_BOTH) continue;
        if (n >= 5 && k > 2 && k < n - 1) continue;  // keep the CPU run short
        specs.push_back({{kind, k}});
      }
    specs.push_back({{fq::KIND_LUMPED, 0}});
    for (int k = 0; k <= n && n <= 4; ++k) specs.push_back(fq::hodge_blocks(k));
    for (const auto& spec : specs) {
      std::vector<fq::BlockLayout> layout;
      const fq::Tape t = fq::build_tape(n, spec, &layout);
      for (int trial = 0; trial < 6; ++trial) {
        std::vector<double> s;
        if (trial == 0) s = fqo::unit_simplex_lengths_sq(n);
        else if (trial == 1) s = kuhn_lengths(n, 0.125);
        else if (trial == 2) s = kuhn_lengths(n, 1.0 / 3.0);
        else s = random_lengths(n, rng, trial == 5 && n >= 2);
        const fqo::Mat g = n >= 1 ? fqo::metric_from_lengths(n, s.data()) : fqo::Mat(0, 0);
        std::vector<double> in;
        if (t.inputs_are_lengths) {
          in = s;
        } else {
          fqo::Mat gi;
          if (!fqo::try_inverse(g, gi)) return 2;
          in = gi.a;
          in.push_back(fqo::cell_volume(g));
        }
        const std::vector<double> out = run_tape(t, in);
        for (const fq::BlockLayout& bl : layout) {
          fqo::Mat ref;
          // a grade off [0,n] is the zero space: the reference never evaluates
          // an element there (whitney_complex.rs:113-122)
          if (bl.rows == 0 || bl.cols == 0) continue;
          if (bl.kind == fq::KIND_LUMPED) {
            ref = fqo::lumped_element(g);
          } else {
            fqo::PairingTables pt(n, bl.grade, bl.kind);
            ref = fqo::pairing_element(pt, g);
          }
          if (ref.r != bl.rows || ref.c != bl.cols) {
            std::printf("SHAPE n=%d kind=%d k=%d: tape %dx%d oracle %dx%d\n", n, bl.kind, bl.grade, bl.rows, bl.cols,
                        ref.r, ref.c);
            return 1;
          }
          for (int e = 0; e < ref.r * ref.c; ++e) {
            const double a = out[size_t(bl.out_offset + e)], b = ref.a[size_t(e)];
            if (!(a == b)) {  // bitwise up to the sign of zero
              std::printf("MISMATCH n=%d kind=%d k=%d trial=%d entry=%d tape=%.17g oracle=%.17g\n", n, bl.kind,
                          bl.grade, trial, e, a, b);
              return 1;
            }
            ++ncases;
          }
        }
      }
    }
  }
  std::printf("OK %" PRId64 "\n", ncases);
  return 0;
}
