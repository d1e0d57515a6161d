// tile_plan.hpp — plan format of the tile-fused numeric assembly (tile.cu) and its host reference builder.
// Pure C++ (no CUDA): shared by the device code (format constants, block-set description) and by the CPU tests,
// which build a plan with the host builder, interpret the record streams against the oracle's element matrices
// and compare the result with the oracle's assembly (tests/cpp/tile_plan_check.cpp).
//
// Decomposition (owner computes, as between GPUs): vertices are clustered into tiles; a tile owns the rows
// (simplices) whose top vertex it contains, i.e. whole CSR rows, and visits every cell touching one of its vertices
// ("cell visit", cv).  Reference path replaced: formoniq/src/galerkin.rs:138-188, hodge.rs:62-72.
//
//   slab (shared memory, doubles)   [0, 2) = 0.0;  block b: slab_base[b] + column slot * plane[class] + row slot, where a
//                                   row slot is one (cell visit, owned local row) pair of the block's row class (= test
//                                   grade), d_b the number of distinct values per row (tape.hpp: set_layout) and the plane
//                                   stride is 1 mod 16 doubles.  Row slots are numbered (group of 32 cell visits, local
//                                   row, cell visit): the 32 producer lanes of a group store one row's value to
//                                   consecutive doubles (conflict-free), and the column slots of one row slot fall on
//                                   consecutive banks for the consumers.
//   cv record (u32 words)           eid[NE] edge ids of the cell | owned-row masks, 8 bits per row class
//   group record (u16)              first row slot of every (class, local row) of a group of 32 cell visits
//   stream (per tile)               chunks of kChunkBytes: {u32 nrec; pad to 16} then records
//                                   {u32 L | block << 8 | lanes << 16; pad to 16; u32 dest[lanes]; u16 code[L][lanes]}
//                                   blocks in order, inside a block runs of equal L (ascending), inside a run CSR order;
//                                   dest = 0 padding lane, 1 dropped non-zero, d >= 2 position d - 2 of the values;
//                                   code = slab index of a contribution, contributions in ascending cell-visit order
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "tape.hpp"

namespace fq {
namespace tp {

constexpr int kMaxBlocks = 4;
constexpr int kMaxClasses = 2;             // row classes = distinct test grades of a block set
constexpr int kMaxLocal = 6;            // local faces of one grade, n <= 3
constexpr int kGroupWords = kMaxClasses * kMaxLocal;  // u16 entries of a group record
constexpr int kChunkBytes = 1536;       // TMA granule of the stream; a record never straddles a chunk
constexpr int kChunkHdr = 16;
constexpr int kRecHdr = 16;
constexpr int kMaxLen64 = ((kChunkBytes - 32) / 64 - 4) / 2;  // 64-lane records up to L = 9
constexpr int kMaxLen32 = ((kChunkBytes - 32) / 32 - 4) / 2;  // 32-lane records up to L = 21
constexpr int kMaxLen = ((kChunkBytes - 32) / 16 - 4) / 2;    // 16-lane records up to L = 45
constexpr uint32_t kPadDest = 0u, kNoDest = 1u;
constexpr int kZeroSlots = 2;           // slab doubles holding 0.0 (padding lanes and exact-zero entries read them)
constexpr int kMaxCv = 512;             // cell visits of a tile
constexpr int kMaxEntries = 9216;       // owned entries of one block in one tile (device sort: 512 threads x 18)
// shared memory of the fused kernel: slab | TMA ring (two chunks per consumer warp, at most 16 warps) | mbarriers
constexpr int kMaxConsumerWarps = 16;
constexpr int kSlotsPerWarp = 2;
constexpr int kMaxGroups = 3;
constexpr size_t kSmemCta = size_t(227) * 1024 - 256;
constexpr size_t kRingBytesMax = size_t(kMaxConsumerWarps) * kSlotsPerWarp * kChunkBytes;
constexpr size_t kBarBytes = (size_t(kMaxConsumerWarps) * kSlotsPerWarp + 2 * kMaxGroups + 2) * 8;
constexpr uint32_t kSlabCapacity = uint32_t((kSmemCta - kRingBytesMax - kBarBytes - 256) / 8);  // doubles

#if defined(__CUDACC__)
#define FQ_TP_HD __host__ __device__
#else
#define FQ_TP_HD
#endif
FQ_TP_HD inline uint32_t rec_lanes(uint32_t L) { return L <= uint32_t(kMaxLen64) ? 64u : (L <= uint32_t(kMaxLen32) ? 32u : 16u); }
FQ_TP_HD inline uint32_t rec_bytes(uint32_t L) { return uint32_t(kRecHdr) + rec_lanes(L) * (4u + 2u * L); }

struct TileHdr {  // 32 bytes
  uint32_t cv_begin, ncv;
  uint32_t chunk_begin, nchunks;
  uint32_t gb_slot;   // first group record of the tile: (cv_begin >> 5) + tile index
  uint32_t pad[3];
};
inline uint32_t round_plane(uint32_t rs_max) {  // smallest stride >= rs_max that is 1 mod 16
  return (rs_max + 14u) / 16u * 16u + 1u;
}

struct BlockDesc {
  int kind, grade, tg, rg;
  int nt, nr;           // local rows / columns
  int group;            // stage group (-1: empty block, not in the stream)
  int d;                // column slots per row
  int rclass;           // row class (grade tg, row range)
  int empty;            // zero space: no rows or no columns
  uint32_t row_begin, row_end;
  uint8_t cs[kMaxLocal * kMaxLocal];  // [r * nr + j] column slot, 0xFF = exact zero
};
struct SetDesc {
  int n, nv, ne;        // cell dimension, vertices and edges per cell
  int nblocks, ngroups, nclasses;
  int nl[4];            // local faces per grade
  uint8_t top[4][kMaxLocal];  // top vertex position of every local face, per grade
  BlockDesc blk[kMaxBlocks];
  int class_grade[kMaxClasses];
  uint32_t class_lo[kMaxClasses], class_hi[kMaxClasses];
  int cv_words;         // ne + 1
};

// Description of a block set on n-cells; row ranges default to everything (set them, then call finish_classes).
inline SetDesc make_set(int n, const std::vector<BlockSpec>& specs) {
  if (n < 1 || n > 3 || specs.empty() || int(specs.size()) > kMaxBlocks) throw std::runtime_error("make_set: unsupported set");
  const SetLayout L = set_layout(n, specs);
  SetDesc S;
  std::memset(&S, 0, sizeof S);
  S.n = n, S.nv = n + 1, S.ne = int(binom(n + 1, 2));
  S.nblocks = int(specs.size());
  S.ngroups = L.ngroups;
  for (int g = 0; g <= n; ++g) {
    S.nl[g] = nlocal(n, g);
    const auto sub = colex_subsets(n + 1, g + 1);
    for (size_t r = 0; r < sub.size(); ++r) S.top[g][r] = uint8_t(mask_elems(sub[r]).back());
  }
  for (int b = 0; b < S.nblocks; ++b) {
    const SetBlock& sb = L.blocks[size_t(b)];
    BlockDesc& B = S.blk[b];
    B.kind = sb.kind, B.grade = sb.grade, B.tg = sb.tg, B.rg = sb.rg;
    B.nt = sb.rows, B.nr = sb.cols, B.group = sb.group, B.d = sb.d;
    B.empty = (sb.rows == 0 || sb.cols == 0) ? 1 : 0;
    B.row_begin = 0, B.row_end = 0xFFFFFFFFu;
    std::memset(B.cs, 0xFF, sizeof B.cs);
    for (size_t e = 0; e < sb.cs.size(); ++e) B.cs[e] = sb.cs[e] < 0 ? uint8_t(0xFF) : uint8_t(sb.cs[e]);
  }
  return S;
}
// Row classes = test grades; blocks of one class must cover the same row range.  false: unsupported set.
inline bool finish_classes(SetDesc& S) {
  S.nclasses = 0;
  for (int b = 0; b < S.nblocks; ++b) {
    BlockDesc& B = S.blk[b];
    B.rclass = -1;
    if (B.empty) continue;
    for (int c = 0; c < S.nclasses; ++c)
      if (S.class_grade[c] == B.tg) {
        if (S.class_lo[c] != B.row_begin || S.class_hi[c] != B.row_end) return false;
        B.rclass = c;
      }
    if (B.rclass < 0) {
      if (S.nclasses == kMaxClasses) return false;
      S.class_grade[S.nclasses] = B.tg, S.class_lo[S.nclasses] = B.row_begin, S.class_hi[S.nclasses] = B.row_end;
      B.rclass = S.nclasses++;
    }
  }
  S.cv_words = S.ne + 1;
  return true;
}

// Per-grade row-slot budgets of a tile and the plane strides that follow from them.  A row slot is one (cell visit,
// owned local row) pair; an interior Kuhn vertex brings S_g = dim! * C(dim+1, g+1) of them for grade g.  The slab holds
// sum_b d_b * plane[tg(b)] doubles whatever the tile, so each tile is limited per grade: R_g = V* * S_g with V* the
// number of interior vertices the most demanding Hodge set of this dimension can hold.  Then any Hodge set (and any
// single block) fits.  The planes are compile-time constants of the generated producers (gen_elmat.cpp prints them).
struct DimBudget {
  uint32_t vstar = 0;
  uint32_t budget[4] = {0, 0, 0, 0};  // [grade] row slots of a tile
  uint32_t plane[4] = {1, 1, 1, 1};   // [grade] plane stride (doubles, 1 mod 16)
};
inline DimBudget dim_budget(int dim) {
  DimBudget db;
  uint32_t max_nr[4] = {0, 0, 0, 0};
  double sg[4] = {0, 0, 0, 0};
  for (int g = 0; g <= dim; ++g) sg[g] = double(fact(dim)) * double(binom(dim + 1, g + 1));
  // planes are a little larger than the budgets (rounded up to 1 mod 16): find the largest V* whose planes still fit
  for (uint32_t v = 1; v < 4096; ++v) {
    bool ok = true;
    for (int k = 0; k <= dim && ok; ++k) {
      const SetDesc S = make_set(dim, hodge_blocks(k));
      uint64_t total = kZeroSlots;
      for (int b = 0; b < S.nblocks; ++b) {
        const BlockDesc& B = S.blk[b];
        if (B.empty) continue;
        total += uint64_t(round_plane(uint32_t(v * sg[B.tg]))) * uint64_t(B.d);
        max_nr[B.tg] = std::max(max_nr[B.tg], uint32_t(B.nr));
        if (uint64_t(v * sg[B.tg]) * uint64_t(B.nr) > uint64_t(kMaxEntries)) ok = false;  // the builder's sort
        if (v * sg[B.tg] > 65000.0) ok = false;
      }
      if (total > kSlabCapacity) ok = false;
    }
    if (!ok) break;
    db.vstar = v;
  }
  for (int g = 0; g <= dim; ++g) {
    db.budget[g] = uint32_t(db.vstar * sg[g]);
    db.plane[g] = round_plane(db.budget[g]);
  }
  return db;
}

// Plane strides and block regions of the slab: the planes of the dimension's budgets (compile-time constants of the
// generated producers), regions in block order.
inline void set_slab_layout(const SetDesc& S, const DimBudget& db, uint32_t* plane, uint32_t* slab_base, uint32_t& total) {
  for (int c = 0; c < kMaxClasses; ++c) plane[c] = c < S.nclasses ? db.plane[S.class_grade[c]] : 1u;
  uint32_t slab = kZeroSlots;
  for (int b = 0; b < kMaxBlocks; ++b) {
    slab_base[b] = slab;
    if (b < S.nblocks && !S.blk[b].empty) slab += plane[S.blk[b].rclass] * uint32_t(S.blk[b].d);
  }
  total = slab;
}

// ---------------------------------------------------------------- host reference builder
struct HostMesh {
  size_t ncells = 0;
  const uint32_t* faces[4] = {nullptr, nullptr, nullptr, nullptr};  // [ncells][nl[g]] global face ids
  const uint32_t* vertex_tile = nullptr;                            // tile of vertex v: vertex_tile[v - v_lo]
  uint32_t v_lo = 0;
  uint32_t ntiles = 0;
  const uint32_t* tile_cv_ptr = nullptr;    // [ntiles+1]
  const uint32_t* tile_cv_cells = nullptr;  // cells of every tile, ascending
};
struct HostPlan {
  std::vector<TileHdr> tiles;
  std::vector<uint32_t> cv_rec;               // [ncv_total][cv_words]
  std::vector<unsigned char> stream;
  std::vector<uint32_t> row_ptr[kMaxBlocks];  // structural pattern of every block (local rows)
  std::vector<uint32_t> col_idx[kMaxBlocks];
  std::vector<uint16_t> gbase;                // group records [gb_slot + group][kGroupWords]
  uint32_t max_slab = 0;                      // slab doubles: kZeroSlots + the regions of all blocks
  uint32_t rs_max[kMaxClasses] = {0, 0};      // largest row-slot count of every row class over the tiles
  uint32_t plane[kMaxClasses] = {1, 1};       // plane strides (doubles)
  uint32_t slab_base[kMaxBlocks] = {0, 0, 0, 0};  // region of every block: the SAME for all tiles (a tile's producers
                                              // refill one stage group's region while the consumers still read the others)
  uint64_t nentries[kMaxBlocks] = {0, 0, 0, 0};  // owned element entries (contributions) of every block
};

// Sequential placement of records into the chunks of one tile's stream (the device builder computes the same
// positions in closed form per run of equal-sized records).
struct Cursor {
  uint32_t off = 0;          // next free byte of the open chunk
  uint32_t chunk_start = 0;  // byte offset of the open chunk
  uint32_t in_chunk = 0;     // records in the open chunk
  bool open = false;
  uint32_t next_start = 0;   // where the next chunk begins
  // returns the byte offset of a record of `size` bytes; writes the header of a chunk it closes when `base` is given
  uint32_t place(uint32_t size, unsigned char* base) {
    if (!open || off + size > chunk_start + uint32_t(kChunkBytes)) {
      close(base);
      chunk_start = next_start;
      next_start = chunk_start + uint32_t(kChunkBytes);
      off = chunk_start + uint32_t(kChunkHdr);
      in_chunk = 0;
      open = true;
    }
    const uint32_t at = off;
    off += size;
    ++in_chunk;
    return at;
  }
  void close(unsigned char* base) {
    if (open && base) std::memcpy(base + chunk_start, &in_chunk, 4);
    open = false;
  }
  uint32_t nchunks() const { return next_start / uint32_t(kChunkBytes); }
};

class HostBuilder {
 public:
  HostBuilder(const SetDesc& set, const HostMesh& mesh) : S(set), M(mesh) {}
  // Runs both passes; throws std::runtime_error when a tile exceeds a format limit.
  HostPlan build(uint32_t slab_capacity_doubles) {
    HostPlan P;
    P.tiles.assign(M.ntiles, TileHdr{});
    for (int b = 0; b < S.nblocks; ++b) {
      const BlockDesc& B = S.blk[b];
      P.row_ptr[b].assign(size_t(B.empty ? 0 : B.row_end - B.row_begin) + 1, 0u);
    }
    set_slab_layout(S, dim_budget(S.n), P.plane, P.slab_base, P.max_slab);
    if (P.max_slab > slab_capacity_doubles || P.max_slab > 0x10000u) throw std::runtime_error("tile plan: the set exceeds the shared slab");
    std::vector<uint32_t> nchunks(M.ntiles, 0);
    for (uint32_t t = 0; t < M.ntiles; ++t) tile(t, false, slab_capacity_doubles, P, nchunks[t]);
    for (int c = 0; c < S.nclasses; ++c)
      if (P.rs_max[c] > P.plane[c]) throw std::runtime_error("tile plan: a tile exceeds its row-slot budget");
    for (int b = 0; b < S.nblocks; ++b) {  // row lengths -> exclusive scan
      uint32_t run = 0;
      for (uint32_t& v : P.row_ptr[b]) {
        const uint32_t len = v;
        v = run;
        run += len;
      }
      P.col_idx[b].assign(run, 0u);
    }
    uint32_t chunk = 0;
    for (uint32_t t = 0; t < M.ntiles; ++t) {
      P.tiles[t].chunk_begin = chunk;
      P.tiles[t].nchunks = nchunks[t];
      chunk += nchunks[t];
    }
    P.stream.assign(size_t(chunk) * kChunkBytes, 0);
    P.cv_rec.assign(size_t(M.tile_cv_ptr[M.ntiles]) * size_t(S.cv_words), 0u);
    P.gbase.assign((size_t(M.tile_cv_ptr[M.ntiles] >> 5) + M.ntiles + 1) * kGroupWords, uint16_t(0));
    for (uint32_t t = 0; t < M.ntiles; ++t) {
      uint32_t n2 = 0;
      tile(t, true, slab_capacity_doubles, P, n2);
      if (n2 != nchunks[t]) throw std::runtime_error("tile plan: passes disagree");
    }
    return P;
  }

 private:
  const SetDesc& S;
  const HostMesh& M;
  struct Ent {
    uint64_t key;
    uint16_t code;
  };
  void tile(uint32_t t, bool emit, uint32_t cap, HostPlan& P, uint32_t& nchunks_out) {
    const uint32_t cv0 = M.tile_cv_ptr[t], ncv = M.tile_cv_ptr[t + 1] - cv0;
    if (ncv > uint32_t(kMaxCv)) throw std::runtime_error("tile plan: too many cell visits in a tile");
    TileHdr& H = P.tiles[t];
    H.cv_begin = cv0, H.ncv = ncv;
    H.gb_slot = (cv0 >> 5) + t;
    // row classes: owned-row masks of every cell visit; row slots numbered (group of 32 cell visits, local row, cell visit)
    std::vector<uint32_t> mask[kMaxClasses];
    std::vector<uint16_t> rs[kMaxClasses];  // [cv * kMaxLocal + row]
    const uint32_t ngroups = (ncv + 31u) / 32u;
    for (int c = 0; c < S.nclasses; ++c) {
      const int g = S.class_grade[c], nl = S.nl[g];
      mask[c].assign(ncv, 0);
      rs[c].assign(size_t(ncv) * kMaxLocal, 0);
      for (uint32_t i = 0; i < ncv; ++i) {
        const size_t cell = M.tile_cv_cells[cv0 + i];
        uint32_t m = 0;
        for (int r = 0; r < nl; ++r) {
          const uint32_t row = M.faces[g][cell * nl + r];
          const uint32_t topv = M.faces[0][cell * S.nv + S.top[g][r]];
          if (M.vertex_tile[topv - M.v_lo] == t && row >= S.class_lo[c] && row < S.class_hi[c]) m |= 1u << r;
        }
        mask[c][i] = m;
      }
      uint32_t counter = 0;
      for (uint32_t G = 0; G < ngroups; ++G)
        for (int r = 0; r < nl; ++r) {
          if (emit) P.gbase[(size_t(H.gb_slot) + G) * kGroupWords + size_t(c) * kMaxLocal + r] = uint16_t(counter);
          for (uint32_t i = 32u * G; i < std::min(ncv, 32u * G + 32u); ++i)
            if (mask[c][i] >> r & 1u) rs[c][size_t(i) * kMaxLocal + r] = uint16_t(counter++);
        }
      if (counter > 0xFFFFu) throw std::runtime_error("tile plan: too many row slots in a tile");
      P.rs_max[c] = std::max(P.rs_max[c], counter);
    }
    (void)cap;
    if (emit)
      for (uint32_t i = 0; i < ncv; ++i) {
        const size_t cell = M.tile_cv_cells[cv0 + i];
        uint32_t* rec = P.cv_rec.data() + size_t(cv0 + i) * S.cv_words;
        for (int e = 0; e < S.ne; ++e) rec[e] = M.faces[1][cell * S.ne + e];
        uint32_t word = 0;
        for (int c = 0; c < S.nclasses; ++c) word |= mask[c][i] << (8 * c);
        rec[S.ne] = word;
      }
    // blocks
    unsigned char* sbase = emit ? P.stream.data() + size_t(H.chunk_begin) * kChunkBytes : nullptr;
    Cursor cur;
    for (int b = 0; b < S.nblocks; ++b) {
      const BlockDesc& B = S.blk[b];
      if (B.empty) continue;
      const int c = B.rclass;
      std::vector<Ent> ents;
      for (uint32_t i = 0; i < ncv; ++i) {
        const size_t cell = M.tile_cv_cells[cv0 + i];
        const uint32_t m = mask[c][i];
        for (int r = 0; r < B.nt; ++r) {
          if (!(m >> r & 1u)) continue;
          const uint32_t slot = rs[c][size_t(i) * kMaxLocal + r];
          const uint32_t row = M.faces[B.tg][cell * B.nt + r];
          for (int j = 0; j < B.nr; ++j) {
            const uint32_t col = M.faces[B.rg][cell * B.nr + j];
            const uint8_t cs = B.cs[r * B.nr + j];
            const uint32_t code = cs == 0xFF ? 0u : P.slab_base[b] + uint32_t(cs) * P.plane[c] + slot;
            ents.push_back(Ent{(uint64_t(row) << 32) | col, uint16_t(code)});
          }
        }
      }
      if (!emit) P.nentries[b] += ents.size();
      if (ents.size() > size_t(kMaxEntries)) throw std::runtime_error("tile plan: too many entries of one block in a tile");
      std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b2) { return a.key < b2.key; });
      struct Nz {
        uint32_t first, L, dest;
      };
      std::vector<Nz> nz;
      for (size_t i = 0; i < ents.size();) {
        size_t j = i;
        while (j < ents.size() && ents[j].key == ents[i].key) ++j;
        const uint32_t row = uint32_t(ents[i].key >> 32), col = uint32_t(ents[i].key);
        uint32_t dest = 0;
        const size_t rl = row - B.row_begin;
        if (!emit) {
          P.row_ptr[b][rl] += 1;
        } else {
          uint32_t rank = 0;  // rank of this column inside its row: non-zeros of the row seen so far
          for (size_t k = nz.size(); k-- > 0;) {
            if (uint32_t(ents[nz[k].first].key >> 32) != row) break;
            ++rank;
          }
          const uint32_t q = P.row_ptr[b][rl] + rank;
          P.col_idx[b][q] = col;
          dest = q + 2u;
        }
        nz.push_back(Nz{uint32_t(i), uint32_t(j - i), dest});
        i = j;
      }
      std::vector<uint32_t> order(nz.size());
      for (size_t i = 0; i < nz.size(); ++i) order[i] = uint32_t(i);
      std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b2) { return nz[a].L < nz[b2].L; });
      for (size_t p0 = 0; p0 < order.size();) {
        const uint32_t L = nz[order[p0]].L;
        if (L > uint32_t(kMaxLen)) throw std::runtime_error("tile plan: a non-zero has too many contributions");
        size_t p1 = p0;
        while (p1 < order.size() && nz[order[p1]].L == L) ++p1;
        const uint32_t lanes = rec_lanes(L), size = rec_bytes(L);
        for (size_t r0 = p0; r0 < p1; r0 += lanes) {
          const uint32_t at = cur.place(size, sbase);
          if (emit) {
            unsigned char* rp = sbase + at;
            const uint32_t h = L | (uint32_t(b) << 8) | (lanes << 16);
            std::memcpy(rp, &h, 4);
            for (size_t p = r0; p < std::min(p1, r0 + lanes); ++p) {
              const Nz& z = nz[order[p]];
              const uint32_t lane = uint32_t(p - r0);
              std::memcpy(rp + kRecHdr + 4 * lane, &z.dest, 4);
              for (uint32_t j = 0; j < L; ++j)
                std::memcpy(rp + kRecHdr + 4 * lanes + 2 * (j * lanes + lane), &ents[z.first + j].code, 2);
            }
          }
        }
        p0 = p1;
      }
    }
    cur.close(sbase);
    nchunks_out = cur.nchunks();
  }
};

}  // namespace tp
}  // namespace fq
