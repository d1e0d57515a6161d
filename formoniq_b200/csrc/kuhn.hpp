// kuhn.hpp — closed-form colex numbering of the Kuhn (Freudenthal)
// triangulation of a box grid (host tables; product code, no CUDA needed).
//
// The reference builds a `Complex` by generating every sub-simplex of every
// cell, sorting + deduplicating them (simplicial/src/topology/skeleton.rs:50-86,
// complex.rs:298-336) and then finds face ids by hashing vertex words
// (skeleton.rs:121).  On a Kuhn grid (simplicial/src/mesher/grid.rs:77-103)
// the same numbering has a closed form, which is what lets the mesh tables be
// generated on the device at >= 10 M cells:
//
//  * vertices are linearised axis 0 fastest; every j-simplex is a chain
//        v_0 < v_1 < ... < v_j,  v_i = v_0 + off(T_i),  0 != T_1 c T_2 c ... c T_j
//    of axis subsets.  Seen from its top vertex w = v_j it is the descending
//    chain U_1 c ... c U_j with v_{j-i} = w - off(U_i).
//  * colex order compares v_j first, then v_{j-1}, ...  (simplex.rs:126-129).
//    off() is monotone in the subset's bitmask, so among the simplices with a
//    common top vertex the order is the descending lexicographic order of
//    (U_1, ..., U_j) — independent of the grid size.
//  * a chain type is present at w iff U_j only uses axes along which w is not
//    on the lower boundary.  Hence
//        id(simplex) = vbase_j[w] + rank_j[B(w)][type],
//    with B(w) the mask of axes with w_a >= 1 and vbase_j the exclusive prefix
//    sum over vertices of the number of valid types.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <vector>

#include "tape.hpp"  // binom, colex_subsets, fact

namespace fq {

struct KuhnGrade {
  int j = 0;
  int ntypes = 0;
  std::vector<std::vector<uint8_t>> chains;  // [ntypes][j]  U_1..U_j
  std::vector<uint16_t> rank_in;             // [2^n][ntypes] rank among the types valid for mask B
  std::vector<uint32_t> cnt;                 // [2^n] number of valid types
};

struct KuhnTables {
  int n = 0;
  int ncelltypes = 0;                         // n!
  std::vector<KuhnGrade> grades;              // [n+1]
  std::vector<std::vector<uint8_t>> cell_P;   // [n!][n+1] ascending chain P_0=0 c ... c P_n=full
  // per grade j: [n!][nlocal(n,j)]
  std::vector<std::vector<uint16_t>> ftype;
  std::vector<std::vector<uint8_t>> ftop;

  explicit KuhnTables(int n_) : n(n_) {
    if (n < 1 || n > 6) throw std::runtime_error("Kuhn generator supports 1 <= dim <= 6");
    const uint32_t full = (1u << n) - 1;
    grades.resize(size_t(n) + 1);
    std::vector<std::map<std::vector<uint8_t>, int>> type_index(size_t(n) + 1);
    for (int j = 0; j <= n; ++j) {
      KuhnGrade& g = grades[size_t(j)];
      g.j = j;
      if (j == 0) {
        g.ntypes = 1;
        g.chains.assign(1, {});
      } else {
        std::vector<std::vector<uint8_t>> acc;
        std::vector<uint8_t> cur;
        // enumerate strict chains of non-empty subsets of length j
        struct Rec {
          static void go(int j, uint32_t full, std::vector<uint8_t>& cur, std::vector<std::vector<uint8_t>>& acc) {
            if (int(cur.size()) == j) {
              acc.push_back(cur);
              return;
            }
            const uint32_t prev = cur.empty() ? 0u : cur.back();
            for (uint32_t m = 1; m <= full; ++m)
              if ((m & prev) == prev && m != prev) {
                cur.push_back(uint8_t(m));
                go(j, full, cur, acc);
                cur.pop_back();
              }
          }
        };
        Rec::go(j, full, cur, acc);
        std::sort(acc.begin(), acc.end(), [](const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {
          return a > b;  // descending lexicographic
        });
        g.chains = acc;
        g.ntypes = int(acc.size());
      }
      for (int t = 0; t < g.ntypes; ++t) type_index[size_t(j)][g.chains[size_t(t)]] = t;
      g.rank_in.assign(size_t(full + 1) * g.ntypes, 0);
      g.cnt.assign(size_t(full) + 1, 0);
      for (uint32_t B = 0; B <= full; ++B) {
        uint32_t r = 0;
        for (int t = 0; t < g.ntypes; ++t) {
          const uint32_t top = j == 0 ? 0u : g.chains[size_t(t)].back();
          g.rank_in[size_t(B) * g.ntypes + t] = uint16_t(r);
          if ((top & ~B) == 0) ++r;
        }
        g.cnt[B] = r;
      }
    }
    // cell types: the n-chains in sorted order; P_m = full \ U_{n-m}
    const KuhnGrade& gc = grades[size_t(n)];
    ncelltypes = gc.ntypes;
    cell_P.resize(size_t(ncelltypes));
    for (int t = 0; t < ncelltypes; ++t) {
      std::vector<uint8_t>& P = cell_P[size_t(t)];
      P.assign(size_t(n) + 1, 0);
      for (int m = 0; m <= n; ++m) P[size_t(m)] = m == n ? uint8_t(full) : uint8_t(full & ~gc.chains[size_t(t)][size_t(n - m - 1)]);
    }
    ftype.resize(size_t(n) + 1);
    ftop.resize(size_t(n) + 1);
    for (int j = 0; j <= n; ++j) {
      const auto subs = colex_subsets(n + 1, j + 1);
      ftype[size_t(j)].assign(size_t(ncelltypes) * subs.size(), 0);
      ftop[size_t(j)].assign(size_t(ncelltypes) * subs.size(), 0);
      for (int t = 0; t < ncelltypes; ++t)
        for (size_t l = 0; l < subs.size(); ++l) {
          const auto pos = mask_elems(subs[l]);  // s_0 < ... < s_j
          const uint8_t top = cell_P[size_t(t)][size_t(pos[size_t(j)])];
          std::vector<uint8_t> U;
          for (int i = 1; i <= j; ++i) U.push_back(uint8_t(top & ~cell_P[size_t(t)][size_t(pos[size_t(j - i)])]));
          ftype[size_t(j)][size_t(t) * subs.size() + l] = uint16_t(type_index[size_t(j)].at(U));
          ftop[size_t(j)][size_t(t) * subs.size() + l] = top;
        }
    }
  }
};

// Grid bookkeeping shared by the host evaluator and the device generator.
struct KuhnGrid {
  int n = 0;
  uint64_t shape[8] = {0}, vstride[8] = {0};
  uint64_t nverts = 0, nboxes = 0;
  KuhnGrid(int n_, const size_t* shape_) : n(n_) {
    uint64_t vs = 1, nb = 1;
    for (int a = 0; a < n; ++a) {
      if (shape_[a] < 1) throw std::runtime_error("Kuhn grid needs >= 1 cell per axis");
      shape[a] = shape_[a];
      vstride[a] = vs;
      vs *= shape_[a] + 1;
      nb *= shape_[a];
    }
    nverts = vs;
    nboxes = nb;
  }
  uint32_t lower_mask(uint64_t w) const {  // B(w): axes with coordinate >= 1
    uint32_t B = 0;
    for (int a = 0; a < n; ++a) {
      if (w % (shape[a] + 1) != 0) B |= 1u << a;
      w /= shape[a] + 1;
    }
    return B;
  }
  uint64_t mask_offset(uint32_t m) const {
    uint64_t o = 0;
    for (int a = 0; a < n; ++a)
      if (m >> a & 1) o += vstride[a];
    return o;
  }
};

// vbase_j[w] for all vertices (+ total at the end), host version.
inline std::vector<uint64_t> kuhn_vbase_host(const KuhnTables& kt, const KuhnGrid& g, int j) {
  std::vector<uint64_t> vb(size_t(g.nverts) + 1);
  uint64_t acc = 0;
  for (uint64_t w = 0; w < g.nverts; ++w) {
    vb[size_t(w)] = acc;
    acc += kt.grades[size_t(j)].cnt[g.lower_mask(w)];
  }
  vb[size_t(g.nverts)] = acc;
  return vb;
}

}  // namespace fq
