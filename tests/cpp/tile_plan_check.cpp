// tile_plan_check.cpp — CPU-side check of the tile plan format and of the generated staged block-set functions the
// tile-fused kernel executes (formoniq_b200/csrc/tile_plan.hpp, elmat_gen.cuh): TEST ONLY.
//   1. builds Kuhn meshes with the oracle (oracle/fq_oracle.hpp), clusters their vertices into bricks,
//   2. builds the plan with the host reference builder (the device builder must produce the same bytes),
//   3. interprets the plan the way the kernel does: per tile, every cell visit runs the generated stage functions
//      (device intrinsics mapped to plain IEEE operations, -ffp-contract=off) into a slab, every record lane sums its
//      contributions left to right and stores to `dest`,
//   4. compares pattern and values with the oracle's assembly (structural pattern bit for bit, values bitwise up to
//      the sign of zero; the "any contribution != 0" flags must reproduce the reference's value-dependent pattern).
// Prints "OK <cases> <nnz>" or the first mismatch; exit code 0/1.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
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
#include "../../formoniq_b200/csrc/tile_plan.hpp"
#include "../../oracle/fq_oracle.hpp"

using namespace fq;

struct HostSink {
  double* slab;
  const tp::SetDesc* S;
  const tp::HostPlan* P;
  const tp::TileHdr* H;
  uint32_t cv;           // cell visit within the tile
  bool bad = false;
  // row slot of (class C, local row R) of this cell visit the way the producers find it: the group record's first
  // slot + the number of lower cell visits of the group that own the row
  uint32_t row_slot(int C, int R) const {
    const uint32_t G = cv >> 5;
    uint32_t rs = P->gbase[(size_t(H->gb_slot) + G) * tp::kGroupWords + size_t(C) * tp::kMaxLocal + R];
    for (uint32_t i = 32u * G; i < cv; ++i) {
      const uint32_t word = P->cv_rec[size_t(H->cv_begin + i) * S->cv_words + S->ne];
      rs += (word >> (8 * C) >> R) & 1u;
    }
    return rs;
  }
  template <int B, int R, int CS, int DD, int C>
  void put(double v) {
    const tp::BlockDesc& D = S->blk[B];
    if (DD != D.d || C != D.rclass) bad = true;
    const uint32_t word = P->cv_rec[size_t(H->cv_begin + cv) * S->cv_words + S->ne];
    if (!((word >> (8 * C) >> R) & 1u)) return;
    slab[P->slab_base[B] + uint32_t(CS) * P->plane[C] + row_slot(C, R)] = v;
  }
};

struct SetFn {
  int n, fused_k, kind, grade, nmid;
  void (*a)(const double*, double*);
  void (*g[3])(const double*, HostSink&);
};
#define FQ_SET_ROW(fn, n, fk, kind, grade, nin) \
  SetFn{n, fk, kind, grade, fn##_nmid, &fn##_a, {&fn##_g0<HostSink>, &fn##_g1<HostSink>, &fn##_g2<HostSink>}},
static const SetFn g_sets[] = {FQ_GEN_SET_LIST(FQ_SET_ROW)};

static const SetFn* find_set(int n, int fused_k, int kind, int grade) {
  for (const SetFn& s : g_sets)
    if (s.n == n && s.fused_k == fused_k && (fused_k >= 0 ? s.grade == grade : (s.kind == kind && s.grade == grade))) return &s;
  return nullptr;
}

static bool same_bits_mod_zero(double a, double b) {
  if (a == 0.0 && b == 0.0) return true;
  return std::memcmp(&a, &b, sizeof a) == 0;
}

static long g_nnz = 0;

// shape: boxes per axis; brick: owned vertices per axis of a tile
static bool run_case(int n, const std::vector<int64_t>& shape, const std::vector<int>& brick, int variant, int fused_k, int kind,
                     int grade) {
  const std::vector<int64_t> cells = fqo::kuhn_cells(n, shape.data());
  const fqo::Complex cx = fqo::complex_from_cells(n, cells);
  std::vector<double> lo(size_t(n), 0.0), hi(size_t(n), 1.0), diag(size_t(n), 1.0);
  if (variant == 2) diag[0] = -1.0, hi[0] = 0.7;
  std::vector<double> coords = fqo::kuhn_vertex_coords(n, shape.data(), lo.data(), hi.data());
  const size_t V = size_t(cx.nsimplices(0));
  if (variant == 1)
    for (size_t v = 0; v < V; ++v)
      for (int a = 0; a < n; ++a) coords[v * n + a] += 0.2 / double(shape[size_t(a)]) * fqo::pseudo_random(uint64_t(a), v);
  const std::vector<double> len = fqo::edge_lengths_sq(cx, n, coords.data(), diag.data());
  // u32 face tables
  std::vector<std::vector<uint32_t>> faces(size_t(n) + 1);
  for (int g = 0; g <= n; ++g) faces[size_t(g)].assign(cx.cell_faces[size_t(g)].begin(), cx.cell_faces[size_t(g)].end());
  // vertex bricks
  std::vector<uint32_t> vtile(V);
  std::vector<int64_t> nb(size_t(n), 1);
  uint32_t ntiles = 1;
  for (int a = 0; a < n; ++a) {
    nb[size_t(a)] = (shape[size_t(a)] + 1 + brick[size_t(a)] - 1) / brick[size_t(a)];
    ntiles *= uint32_t(nb[size_t(a)]);
  }
  for (size_t v = 0; v < V; ++v) {
    size_t rem = v;
    uint32_t t = 0, mul = 1;
    for (int a = 0; a < n; ++a) {
      const int64_t c = int64_t(rem % size_t(shape[size_t(a)] + 1));
      rem /= size_t(shape[size_t(a)] + 1);
      t += uint32_t(c / brick[size_t(a)]) * mul;
      mul *= uint32_t(nb[size_t(a)]);
    }
    vtile[v] = t;
  }
  const size_t ncells = size_t(cx.ncells());
  std::vector<std::vector<uint32_t>> tcells(ntiles);
  for (size_t c = 0; c < ncells; ++c)
    for (int j = 0; j <= n; ++j) {
      const uint32_t t = vtile[faces[0][c * (n + 1) + j]];
      if (tcells[t].empty() || tcells[t].back() != uint32_t(c)) tcells[t].push_back(uint32_t(c));
    }
  std::vector<uint32_t> cv_ptr(1, 0), cv_cells;
  for (uint32_t t = 0; t < ntiles; ++t) {
    std::sort(tcells[t].begin(), tcells[t].end());
    tcells[t].erase(std::unique(tcells[t].begin(), tcells[t].end()), tcells[t].end());
    cv_cells.insert(cv_cells.end(), tcells[t].begin(), tcells[t].end());
    cv_ptr.push_back(uint32_t(cv_cells.size()));
  }
  const std::vector<BlockSpec> specs = fused_k >= 0 ? hodge_blocks(fused_k) : std::vector<BlockSpec>{{kind, grade}};
  tp::SetDesc S = tp::make_set(n, specs);
  for (int b = 0; b < S.nblocks; ++b) {
    S.blk[b].row_begin = 0;
    S.blk[b].row_end = S.blk[b].empty ? 0u : uint32_t(cx.nsimplices(S.blk[b].tg));
  }
  if (!tp::finish_classes(S)) {
    std::printf("unsupported set\n");
    return false;
  }
  tp::HostMesh M;
  M.ncells = ncells;
  for (int g = 0; g <= n; ++g) M.faces[g] = faces[size_t(g)].data();
  M.vertex_tile = vtile.data();
  M.ntiles = ntiles;
  M.tile_cv_ptr = cv_ptr.data();
  M.tile_cv_cells = cv_cells.data();
  tp::HostBuilder builder(S, M);
  const tp::HostPlan P = builder.build(tp::kSlabCapacity);
  const SetFn* fn = find_set(n, fused_k, kind, grade);
  if (!fn) {
    std::printf("no generated set for n=%d fused_k=%d kind=%d grade=%d\n", n, fused_k, kind, grade);
    return false;
  }
  // interpret
  std::vector<std::vector<double>> values(size_t(S.nblocks));
  std::vector<std::vector<uint8_t>> keep(size_t(S.nblocks));
  std::vector<std::vector<int>> written(size_t(S.nblocks));
  for (int b = 0; b < S.nblocks; ++b) {
    values[size_t(b)].assign(P.col_idx[b].size(), 0.0);
    keep[size_t(b)].assign(P.col_idx[b].size(), 0);
    written[size_t(b)].assign(P.col_idx[b].size(), 0);
  }
  std::vector<double> slab(size_t(P.max_slab) + 8, 0.0);
  for (uint32_t t = 0; t < ntiles; ++t) {
    const tp::TileHdr& H = P.tiles[t];
    std::fill(slab.begin(), slab.end(), std::nan(""));  // reading a slot nobody stored must show
    slab[0] = slab[1] = 0.0;
    for (uint32_t i = 0; i < H.ncv; ++i) {
      const uint32_t* rec = P.cv_rec.data() + size_t(H.cv_begin + i) * S.cv_words;
      double s[6], mid[32];
      for (int e = 0; e < S.ne; ++e) s[e] = len[rec[e]];
      fn->a(s, mid);
      HostSink sink{slab.data(), &S, &P, &H, i, false};
      for (int g = 0; g < 3; ++g) fn->g[g](mid, sink);
      if (sink.bad) {
        std::printf("generated put<> carries a wrong slot count or row class\n");
        return false;
      }
    }
    int last_block = -1;
    for (uint32_t c = 0; c < H.nchunks; ++c) {
      const unsigned char* chunk = P.stream.data() + size_t(H.chunk_begin + c) * tp::kChunkBytes;
      uint32_t nrec;
      std::memcpy(&nrec, chunk, 4);
      const unsigned char* rp = chunk + tp::kChunkHdr;
      for (uint32_t r = 0; r < nrec; ++r) {
        uint32_t h;
        std::memcpy(&h, rp, 4);
        const uint32_t L = h & 0xFFu, b = (h >> 8) & 3u, lanes = h >> 16;
        if (lanes != tp::rec_lanes(L) || int(b) < last_block) {
          std::printf("bad record header\n");
          return false;
        }
        last_block = int(b);
        for (uint32_t lane = 0; lane < lanes; ++lane) {
          uint32_t dest;
          std::memcpy(&dest, rp + tp::kRecHdr + 4 * lane, 4);
          double acc = 0.0;
          bool any = false;
          for (uint32_t j = 0; j < L; ++j) {
            uint16_t code;
            std::memcpy(&code, rp + tp::kRecHdr + 4 * lanes + 2 * (j * lanes + lane), 2);
            if (dest == tp::kPadDest && code != 0) {
              std::printf("padding lane with a code\n");
              return false;
            }
            const double x = slab[code];
            any = any || (x != 0.0);
            acc = acc + x;
          }
          if (dest == tp::kPadDest) continue;
          if (dest - 2 >= values[b].size()) {
            std::printf("dest out of range\n");
            return false;
          }
          values[b][dest - 2] = acc;
          keep[b][dest - 2] = any ? 1 : 0;
          written[b][dest - 2] += 1;
        }
        rp += tp::rec_bytes(L);
        if (rp > chunk + tp::kChunkBytes) {
          std::printf("record past its chunk\n");
          return false;
        }
      }
    }
  }
  // compare with the oracle
  for (int b = 0; b < S.nblocks; ++b) {
    const tp::BlockDesc& B = S.blk[b];
    if (B.empty) continue;
    const fqo::Csr ref = fqo::assemble_matrix(cx, len.data(), B.kind, B.grade, false);
    const fqo::Csr refd = fqo::assemble_matrix(cx, len.data(), B.kind, B.grade, true);
    if (ref.row_ptr.size() != P.row_ptr[b].size() || ref.col_idx.size() != P.col_idx[b].size()) {
      std::printf("n=%d block %d: pattern size %zu/%zu vs oracle %zu/%zu\n", n, b, P.row_ptr[b].size(), P.col_idx[b].size(),
                  ref.row_ptr.size(), ref.col_idx.size());
      return false;
    }
    for (size_t i = 0; i < ref.row_ptr.size(); ++i)
      if (int64_t(P.row_ptr[b][i]) != ref.row_ptr[i]) {
        std::printf("n=%d block %d: row_ptr[%zu]\n", n, b, i);
        return false;
      }
    size_t kept = 0;
    for (size_t q = 0; q < ref.col_idx.size(); ++q) {
      if (int64_t(P.col_idx[b][q]) != ref.col_idx[q] || written[size_t(b)][q] != 1) {
        std::printf("n=%d block %d: col_idx / coverage at %zu (written %d)\n", n, b, q, written[size_t(b)][q]);
        return false;
      }
      if (!same_bits_mod_zero(values[size_t(b)][q], ref.values[q])) {
        std::printf("n=%d block %d (kind %d grade %d): value[%zu] %a vs %a\n", n, b, B.kind, B.grade, q, values[size_t(b)][q],
                    ref.values[q]);
        return false;
      }
      if (keep[size_t(b)][q]) {
        if (kept >= refd.col_idx.size() || refd.col_idx[kept] != ref.col_idx[q] ||
            !same_bits_mod_zero(refd.values[kept], ref.values[q])) {
          std::printf("n=%d block %d: kept pattern differs at %zu\n", n, b, q);
          return false;
        }
        ++kept;
      }
    }
    if (kept != refd.col_idx.size()) {
      std::printf("n=%d block %d: kept %zu vs reference nnz %zu\n", n, b, kept, refd.col_idx.size());
      return false;
    }
    g_nnz += long(ref.col_idx.size());
  }
  return true;
}

int main() {
  int cases = 0;
  struct Case {
    int n;
    std::vector<int64_t> shape;
    std::vector<int> brick;
  };
  const std::vector<Case> meshes = {
      {1, {9}, {4}}, {2, {5, 4}, {3, 2}}, {2, {7, 6}, {4, 4}}, {3, {3, 2, 3}, {2, 2, 2}}, {3, {4, 5, 3}, {4, 3, 3}}, {3, {2, 2, 2}, {3, 3, 3}}};
  for (const Case& m : meshes)
    for (int variant = 0; variant < 3; ++variant) {
      for (int k = 0; k <= m.n; ++k) {
        if (!run_case(m.n, m.shape, m.brick, variant, k, -1, k)) {
          std::printf("FAILED: n=%d variant=%d hodge k=%d\n", m.n, variant, k);
          return 1;
        }
        ++cases;
      }
      if (variant == 1)
        for (int k = 0; k <= m.n + 1; ++k)
          for (int kind = 0; kind < 4; ++kind) {
            if (kind != 0 && k == 0) continue;
            if (k == m.n + 1 && kind != 3) continue;
            if (!find_set(m.n, -1, kind, k)) continue;  // an all-zero block has no generated set
            if (!run_case(m.n, m.shape, m.brick, variant, -1, kind, k)) {
              std::printf("FAILED: n=%d variant=%d kind=%d k=%d\n", m.n, variant, kind, k);
              return 1;
            }
            ++cases;
          }
    }
  std::printf("OK %d %ld\n", cases, g_nnz);
  return 0;
}
