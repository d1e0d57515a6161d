// staged_core_check.cpp — CPU-side check of the GENERATED core tapes the tile kernels execute
// (formoniq_b200/csrc/elmat_gen.cuh, built by gen_elmat.cpp): the staged form  stage A -> (stage B1, stage B2)  must
// store exactly the bits of the unsplit function into exactly the same slots, B1 must write the first-half slots
// and B2 the second-half slots of the core's `_half2` mask and nothing else.  The device intrinsics are mapped to
// plain IEEE operations (compiled with -ffp-contract=off, like nvcc -fmad=false); TEST ONLY.
// Prints "OK <cores> <cells>" or the first mismatch; exit code 0/1.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#define __device__
#define __forceinline__ inline
#define __restrict__
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }

#include "../../formoniq_b200/csrc/elmat_gen.cuh"

struct SlotSink {
  std::vector<double> v;
  std::vector<int> written;
  explicit SlotSink(int n) : v(size_t(n), 0.0), written(size_t(n), 0) {}
  template <int B, int E>
  void put(double x) {
    v[size_t(E)] = x;
    written[size_t(E)] += 1;
  }
  template <int B, int N>
  void flush() {}
};

static bool same_bits(double a, double b) { return std::memcmp(&a, &b, sizeof a) == 0; }

// squared edge lengths of a random non-degenerate n-simplex (Riemannian or with one time-like axis)
static std::vector<double> random_lengths(int n, std::mt19937_64& rng, bool lorentz) {
  std::normal_distribution<double> nd;
  std::vector<double> x(size_t(n + 1) * n);
  for (double& c : x) c = nd(rng);
  std::vector<double> s;
  for (int j = 1; j <= n; ++j)  // colex pair order: (0,1), (0,2), (1,2), (0,3), ...
    for (int i = 0; i < j; ++i) {
      double acc = 0;
      for (int a = 0; a < n; ++a) {
        const double d = x[size_t(j) * n + a] - x[size_t(i) * n + a];
        acc += ((lorentz && a == 0) ? -0.49 : 1.0) * d * d;
      }
      s.push_back(acc);
    }
  return s;
}

static int g_cells = 0;

#define CHECK_CORE(fn, n, k, variant, nin, nd, nout)                                                        \
  {                                                                                                         \
    std::mt19937_64 rng(1000 * n + 10 * k + variant);                                                       \
    for (int rep = 0; rep < 200; ++rep) {                                                                   \
      const std::vector<double> s = random_lengths(n, rng, rep % 3 == 2);                                   \
      SlotSink full(nd), staged(nd), h1(nd), h2(nd);                                                        \
      fn(s.data(), full);                                                                                   \
      double mid[fn##_nmid];                                                                                \
      fn##_a(s.data(), mid);                                                                                \
      fn##_b1(mid, h1);                                                                                     \
      fn##_b2(mid, h2);                                                                                     \
      for (int slot = 0; slot < nd; ++slot) {                                                               \
        const bool second = (fn##_half2[slot >> 6] >> (slot & 63)) & 1ull;                                  \
        const SlotSink& h = second ? h2 : h1;                                                               \
        const SlotSink& other = second ? h1 : h2;                                                           \
        if (full.written[size_t(slot)] != 1 || h.written[size_t(slot)] != 1 || other.written[size_t(slot)] != 0 || \
            !same_bits(full.v[size_t(slot)], h.v[size_t(slot)])) {                                          \
          std::printf("MISMATCH %s slot %d rep %d: full %a (%d) staged %a (%d/%d) second=%d\n", #fn, slot, rep, \
                      full.v[size_t(slot)], full.written[size_t(slot)], h.v[size_t(slot)], h.written[size_t(slot)], \
                      other.written[size_t(slot)], int(second));                                            \
          return 1;                                                                                         \
        }                                                                                                   \
      }                                                                                                     \
      if (!fn##_split_ok) {                                                                                 \
        std::printf("core %s reports split_ok = 0\n", #fn);                                                 \
        return 1;                                                                                           \
      }                                                                                                     \
      ++g_cells;                                                                                            \
    }                                                                                                       \
    ++ncores;                                                                                               \
  }

int main() {
  int ncores = 0;
  FQ_GEN_CORE_LIST(CHECK_CORE)
  std::printf("OK %d %d\n", ncores, g_cells);
  return 0;
}
