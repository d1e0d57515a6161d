// gen_elmat.cpp — build-time tool: prints the element tapes of the hot (n,k)
// combinations as straight-line CUDA device functions (elmat_gen.cuh).
// Every arithmetic statement is an explicit round-to-nearest intrinsic so the
// compiler can neither contract (FMA) nor re-associate it.
//
//   g++ -O1 -std=c++17 gen_elmat.cpp -o gen_elmat && ./gen_elmat > elmat_gen.cuh
#include "tile_plan.hpp"

#include <cinttypes>
#include <cstdlib>
#include <functional>
#include <sstream>

using namespace fq;

struct Entry {
  std::string name;
  int n;
  int fused_k;  // >=0: hodge_blocks(k); -1: single
  int kind, grade;
  int ninputs, nouts;
  int n_addsub, n_mul, n_div, n_sqrt;
};

static std::string reg(uint32_t r) { return "r" + std::to_string(r); }
static std::string cst(double c) {
  char buf[64];
  std::snprintf(buf, sizeof buf, "%a", c);  // exact hex float
  return buf;
}

static void emit_function(const std::string& name, int n, const std::vector<BlockSpec>& blocks, Entry& e) {
  TapeBuilder tb;
  std::vector<BlockLayout> layout;
  const Tape t = build_tape(n, blocks, &layout, &tb);
  // out index -> (block, entry); a block is flushed after its last store
  auto block_of = [&](uint32_t out) {
    for (size_t b = 0; b < layout.size(); ++b)
      if (int(out) >= layout[b].out_offset && int(out) < layout[b].out_offset + layout[b].rows * layout[b].cols) return int(b);
    return -1;
  };
  std::vector<int> remaining(layout.size(), 0);
  for (size_t b = 0; b < layout.size(); ++b) remaining[b] = layout[b].rows * layout[b].cols;
  auto after_store = [&](uint32_t out) {
    const int b = block_of(out);
    if (--remaining[size_t(b)] == 0)
      std::printf("  sink.template flush<%d, %d>();\n", b, layout[size_t(b)].rows * layout[size_t(b)].cols);
  };
  auto put = [&](uint32_t out, const std::string& v) {
    const int b = block_of(out);
    std::printf("  sink.template put<%d, %d>(%s);\n", b, int(out) - layout[size_t(b)].out_offset, v.c_str());
    after_store(out);
  };
  e.ninputs = t.ninputs;
  e.nouts = t.nouts;
  e.n_addsub = t.n_addsub;
  e.n_mul = t.n_mul;
  e.n_div = t.n_div;
  e.n_sqrt = t.n_sqrt;
  std::printf("// n=%d  inputs=%d outputs=%d  ops: add/sub=%d mul=%d div=%d sqrt=%d\n", n, t.ninputs, t.nouts,
              t.n_addsub, t.n_mul, t.n_div, t.n_sqrt);
  std::printf("template <class Sink>\n__device__ __forceinline__ void %s(const double* __restrict__ s, Sink& sink) {\n",
              name.c_str());
  for (int i = 0; i < t.ninputs; ++i) std::printf("  const double r%d = s[%d];\n", i, i);
  for (const TapeOp& o : tb.ssa_ops()) {
    switch (o.op) {
      case OP_ADD:
        std::printf("  const double %s = __dadd_rn(%s, %s);\n", reg(o.d).c_str(), reg(o.a).c_str(), reg(o.b).c_str());
        break;
      case OP_SUB:
        std::printf("  const double %s = __dsub_rn(%s, %s);\n", reg(o.d).c_str(), reg(o.a).c_str(), reg(o.b).c_str());
        break;
      case OP_MUL:
        std::printf("  const double %s = __dmul_rn(%s, %s);\n", reg(o.d).c_str(), reg(o.a).c_str(), reg(o.b).c_str());
        break;
      case OP_MULC:
        std::printf("  const double %s = __dmul_rn(%s, %s);\n", reg(o.d).c_str(), reg(o.a).c_str(),
                    cst(tb.consts[o.b]).c_str());
        break;
      case OP_DIV:
        std::printf("  const double %s = __ddiv_rn(%s, %s);\n", reg(o.d).c_str(), reg(o.a).c_str(), reg(o.b).c_str());
        break;
      case OP_SQRTABS:
        std::printf("  const double %s = __dsqrt_rn(fabs(%s));\n", reg(o.d).c_str(), reg(o.a).c_str());
        break;
      case OP_LOADC:
        std::printf("  const double %s = %s;\n", reg(o.d).c_str(), cst(tb.consts[o.b]).c_str());
        break;
      case OP_STORE:
        put(o.d, reg(o.a));
        break;
      case OP_STOREN:
        put(o.d, "-" + reg(o.a));
        break;
      case OP_STOREC:
        put(o.d, cst(tb.consts[o.b]));
        break;
    }
  }
  std::printf("}\n\n");
}


// Staged "set" functions for the tile-fused kernel (tile.cu): the element matrices of a block set, every
// distinct value of every row stored once (tape.hpp: set_layout).  Stage A = the geometry head of the tape (up to
// the last division / square root; `mid` carries the values that are live across the cut), then one function per
// stage group (the blocks deriving from one mass grade): the backward slice of that group's stores.  An operation
// needed by several groups is evaluated by each of them on the same operands: same bits as the unsplit tape.
struct SetEntry {
  std::string name;
  int n, fused_k, kind, grade, ninputs, ngroups;
};
static bool emit_set(const std::string& name, int n, const std::vector<BlockSpec>& blocks, SetEntry& e) {
  TapeBuilder tb;
  Tape t;
  const SetLayout L = set_layout(n, blocks, &tb, &t);
  if (L.puts.empty() || L.ngroups > 3 || L.nclasses > 2) return false;
  const std::vector<TapeOp> ops = tb.ssa_ops();
  int cut = -1;
  for (size_t i = 0; i < ops.size(); ++i)
    if (ops[i].op == OP_DIV || ops[i].op == OP_SQRTABS) cut = int(i);
  auto is_store = [](const TapeOp& o) { return o.op == OP_STORE || o.op == OP_STOREN || o.op == OP_STOREC; };
  auto uses = [&](const TapeOp& o, uint32_t* u) -> int {
    switch (o.op) {
      case OP_ADD: case OP_SUB: case OP_MUL: case OP_DIV: u[0] = o.a; u[1] = o.b; return 2;
      case OP_MULC: case OP_SQRTABS: case OP_STORE: case OP_STOREN: u[0] = o.a; return 1;
      default: return 0;
    }
  };
  std::map<int, const SetPut*> put_at;
  for (const SetPut& p : L.puts) put_at[p.op] = &p;
  std::map<uint32_t, int> def_pos;
  for (int i = 0; i < t.ninputs; ++i) def_pos[uint32_t(i)] = -1;
  for (size_t i = 0; i < ops.size(); ++i)
    if (!is_store(ops[i])) def_pos[ops[i].d] = int(i);
  // group masks by backward slicing from the registered puts
  std::map<uint32_t, int> need;
  for (size_t i = ops.size(); i-- > 0;) {
    const TapeOp& o = ops[i];
    uint32_t u[2];
    const int nu = uses(o, u);
    int mask = 0;
    if (is_store(o)) {
      auto it = put_at.find(int(i));
      if (it == put_at.end()) continue;  // a duplicate of an earlier put of the same row, or an exact zero
      mask = 1 << L.blocks[size_t(it->second->block)].group;
    } else {
      mask = need.count(o.d) ? need[o.d] : 0;
    }
    for (int q = 0; q < nu; ++q) need[u[q]] |= mask;
  }
  std::map<uint32_t, int> mid_of;
  for (size_t i = 0; i < ops.size(); ++i) {
    const TapeOp& o = ops[i];
    if (int(i) <= cut && !is_store(o)) continue;
    if (is_store(o) && !put_at.count(int(i))) continue;
    if (!is_store(o) && !(need.count(o.d) && need[o.d])) continue;
    uint32_t u[2];
    const int nu = uses(o, u);
    for (int q = 0; q < nu; ++q)
      if (def_pos.at(u[q]) <= cut && !mid_of.count(u[q])) {
        const int idx = int(mid_of.size());
        mid_of[u[q]] = idx;
      }
  }
  auto line = [&](const TapeOp& o) -> std::string {
    char buf[256];
    switch (o.op) {
      case OP_ADD: std::snprintf(buf, sizeof buf, "  const double %s = __dadd_rn(%s, %s);\n", reg(o.d).c_str(), reg(o.a).c_str(), reg(o.b).c_str()); break;
      case OP_SUB: std::snprintf(buf, sizeof buf, "  const double %s = __dsub_rn(%s, %s);\n", reg(o.d).c_str(), reg(o.a).c_str(), reg(o.b).c_str()); break;
      case OP_MUL: std::snprintf(buf, sizeof buf, "  const double %s = __dmul_rn(%s, %s);\n", reg(o.d).c_str(), reg(o.a).c_str(), reg(o.b).c_str()); break;
      case OP_MULC: std::snprintf(buf, sizeof buf, "  const double %s = __dmul_rn(%s, %s);\n", reg(o.d).c_str(), reg(o.a).c_str(), cst(tb.consts[o.b]).c_str()); break;
      case OP_DIV: std::snprintf(buf, sizeof buf, "  const double %s = __ddiv_rn(%s, %s);\n", reg(o.d).c_str(), reg(o.a).c_str(), reg(o.b).c_str()); break;
      case OP_SQRTABS: std::snprintf(buf, sizeof buf, "  const double %s = __dsqrt_rn(fabs(%s));\n", reg(o.d).c_str(), reg(o.a).c_str()); break;
      case OP_LOADC: std::snprintf(buf, sizeof buf, "  const double %s = %s;\n", reg(o.d).c_str(), cst(tb.consts[o.b]).c_str()); break;
      default: buf[0] = 0; break;
    }
    return buf;
  };
  std::ostringstream sa;
  std::vector<std::ostringstream> sg(3);
  for (int i = 0; i < t.ninputs; ++i) sa << "  const double r" << i << " = s[" << i << "];\n";
  // stage A keeps only what some group (or mid) needs
  for (size_t i = 0; i < ops.size() && int(i) <= cut; ++i) {
    const TapeOp& o = ops[i];
    if (is_store(o)) continue;
    if (need.count(o.d) && need[o.d]) sa << line(o);
  }
  // Stage groups: the stores are emitted column by column (all blocks of the group for column 0, then column 1, ...),
  // each preceded by the not yet emitted part of its dependency cone (depth first, operands in tape order).  Same
  // operations on the same operands as the tape, scheduled so that few values are live at a time: a sandwich column
  // (dif_test = d * M_k) only needs the same column of the mass, so the 36 entries of M_k never have to be live at once.
  std::map<uint32_t, size_t> op_of;  // SSA value -> op index
  for (size_t i = 0; i < ops.size(); ++i)
    if (!is_store(ops[i])) op_of[ops[i].d] = i;
  for (int g = 0; g < L.ngroups && g < 3; ++g) {
    struct PutRef {
      int col, block, row;
      const SetPut* p;
    };
    std::vector<PutRef> order;
    for (const SetPut& p : L.puts) {
      const SetBlock& B = L.blocks[size_t(p.block)];
      if (B.group != g) continue;
      int col = 0;
      for (int j = 0; j < B.cols; ++j)
        if (B.cs[size_t(p.row * B.cols + j)] == p.slot) {
          col = j;
          break;
        }
      order.push_back(PutRef{col, p.block, p.row, &p});
    }
    std::stable_sort(order.begin(), order.end(), [](const PutRef& a, const PutRef& b) {
      if (a.col != b.col) return a.col < b.col;
      if (a.block != b.block) return a.block < b.block;
      return a.row < b.row;
    });
    std::map<uint32_t, bool> done;
    for (const auto& kv : mid_of) done[kv.first] = true;
    // not yet emitted part of the dependency cone of v, as op indices
    std::function<void(uint32_t, std::vector<size_t>&)> collect = [&](uint32_t v, std::vector<size_t>& out) {
      if (done.count(v)) return;
      const size_t at = op_of.at(v);
      if (int(at) <= cut) throw std::runtime_error("set: a stage-A value is missing from mid");
      done[v] = true;
      out.push_back(at);
      uint32_t u[2];
      const int nu = uses(ops[at], u);
      for (int q = 0; q < nu; ++q) collect(u[q], out);
    };
    // Default: the whole group as one batch in tape order, stores in tape order (row-major per block).  Measured on
    // B200: column batches (FQ_GEN_ORDER=col at generation time: fewer live values) make the producers 2x slower.
    const char* gen_order = std::getenv("FQ_GEN_ORDER");
    const bool tape_order = !(gen_order && gen_order[0] == 'c');
    if (tape_order)
      std::stable_sort(order.begin(), order.end(), [](const PutRef& a, const PutRef& b) { return a.p->op < b.p->op; });
    for (size_t p0 = 0; p0 < order.size();) {
      size_t p1 = p0;
      while (p1 < order.size() && (tape_order || order[p1].col == order[p0].col)) ++p1;
      // one column of every block of the group: its operations in TAPE order (independent chains next to each other,
      // which is what gives the FP64 pipe its instruction-level parallelism), then its stores
      std::vector<size_t> batch;
      for (size_t p = p0; p < p1; ++p) {
        const TapeOp& o = ops[size_t(order[p].p->op)];
        if (o.op == OP_STORE || o.op == OP_STOREN) collect(o.a, batch);
      }
      std::sort(batch.begin(), batch.end());
      for (size_t at : batch) sg[size_t(g)] << line(ops[at]);
      for (size_t p = p0; p < p1; ++p) {
        const PutRef& pr = order[p];
        const TapeOp& o = ops[size_t(pr.p->op)];
        const std::string v = (o.op == OP_STORE || o.op == OP_STOREN) ? ((o.op == OP_STOREN ? "-" : "") + reg(o.a)) : cst(tb.consts[o.b]);
      sg[size_t(g)] << "  sink.template put<" << pr.p->block << ", " << pr.p->row << ", " << pr.p->slot << ", "
                    << L.blocks[size_t(pr.p->block)].d << ", " << L.blocks[size_t(pr.p->block)].tclass << ">(" << v << ");\n";
      }
      p0 = p1;
    }
  }
  std::printf("// set %s: n=%d inputs=%d groups=%d  ops: add/sub=%d mul=%d div=%d sqrt=%d\n", name.c_str(), n, t.ninputs, L.ngroups,
              t.n_addsub, t.n_mul, t.n_div, t.n_sqrt);
  const int nmid = int(mid_of.size());
  std::printf("constexpr int %s_nmid = %d;\nconstexpr int %s_ngroups = %d;\n", name.c_str(), nmid > 0 ? nmid : 1, name.c_str(), L.ngroups);
  {
    // slab layout of the tile-fused kernel for this set: plane strides of the row classes and block regions
    // (tile_plan.hpp: dim_budget / set_slab_layout) as compile-time constants, so every store is one instruction
    tp::SetDesc S = tp::make_set(n, blocks);
    for (int b = 0; b < S.nblocks; ++b) S.blk[b].row_begin = 0, S.blk[b].row_end = 1;
    if (!tp::finish_classes(S)) return false;
    uint32_t plane[tp::kMaxClasses], sb[tp::kMaxBlocks], total;
    tp::set_slab_layout(S, tp::dim_budget(n), plane, sb, total);
    std::printf("constexpr unsigned %s_plane0 = %u, %s_plane1 = %u;\n", name.c_str(), plane[0], name.c_str(), plane[1]);
    std::printf("constexpr unsigned %s_sb0 = %u, %s_sb1 = %u, %s_sb2 = %u, %s_sb3 = %u;\n", name.c_str(), sb[0], name.c_str(), sb[1],
                name.c_str(), sb[2], name.c_str(), sb[3]);
  }
  // local rows of the (at most two) row classes = test grades of the set
  std::printf("constexpr int %s_nclasses = %d;\nconstexpr int %s_crows0 = %d;\nconstexpr int %s_crows1 = %d;\n", name.c_str(), L.nclasses,
              name.c_str(), L.nclasses > 0 ? nlocal(n, L.class_grade[0]) : 0, name.c_str(), L.nclasses > 1 ? nlocal(n, L.class_grade[1]) : 0);
  std::printf("__device__ __forceinline__ void %s_a(const double* __restrict__ s, double* __restrict__ mid) {\n%s", name.c_str(),
              sa.str().c_str());
  for (const auto& kv : mid_of) std::printf("  mid[%d] = %s;\n", kv.second, reg(kv.first).c_str());
  std::printf("}\n");
  for (int g = 0; g < 3; ++g) {
    std::printf("template <class Sink>\n__device__ __forceinline__ void %s_g%d(const double* __restrict__ mid, Sink& sink) {\n",
                name.c_str(), g);
    if (g < L.ngroups)
      for (const auto& kv : mid_of) std::printf("  const double %s = mid[%d]; (void)%s;\n", reg(kv.first).c_str(), kv.second, reg(kv.first).c_str());
    std::printf("%s}\n", sg[size_t(g)].str().c_str());
  }
  std::printf("\n");
  e.name = name;
  e.n = n;
  e.ninputs = t.ninputs;
  e.ngroups = L.ngroups;
  return true;
}

int main() {
  std::printf("// GENERATED by gen_elmat.cpp from tape.hpp — do not edit.\n#pragma once\n\n");
  std::vector<Entry> entries;
  for (int n = 1; n <= 3; ++n) {
    for (int k = 0; k <= n + 1; ++k)
      for (int kind = 0; kind < 4; ++kind) {
        if (kind != KIND_MASS && k == 0) continue;
        if (k == n + 1 && kind != KIND_DIF_BOTH) continue;
        Entry e{};
        e.name = "fq_el_n" + std::to_string(n) + "_kind" + std::to_string(kind) + "_k" + std::to_string(k);
        e.n = n;
        e.fused_k = -1;
        e.kind = kind;
        e.grade = k;
        emit_function(e.name, n, {{kind, k}}, e);
        entries.push_back(e);
      }
    {
      Entry e{};
      e.name = "fq_el_n" + std::to_string(n) + "_lumped";
      e.n = n;
      e.fused_k = -1;
      e.kind = KIND_LUMPED;
      e.grade = 0;
      emit_function(e.name, n, {{KIND_LUMPED, 0}}, e);
      entries.push_back(e);
    }
    for (int k = 0; k <= n; ++k) {
      Entry e{};
      e.name = "fq_el_n" + std::to_string(n) + "_hodge" + std::to_string(k);
      e.n = n;
      e.fused_k = k;
      e.kind = -1;
      e.grade = k;
      emit_function(e.name, n, hodge_blocks(k), e);
      entries.push_back(e);
    }
  }
  // dim 4 (BASELINE config 3: 4-D k = 2): the fused Hodge tapes as straight-line code as well; their inputs are g^-1
  // (row-major) and the volume from the generic geometry stage (geometry.cuh: nalgebra's 4x4 cofactor inverse), so the
  // 3 459 operations of hodge_blocks(2) run out of registers instead of the interpreter's global scratch slab
  for (int k = 0; k <= 4; ++k) {
    Entry e{};
    e.name = "fq_el_n4_hodge" + std::to_string(k);
    e.n = 4;
    e.fused_k = k;
    e.kind = -1;
    e.grade = k;
    emit_function(e.name, 4, hodge_blocks(k), e);
    entries.push_back(e);
  }
  // X-macro list: (function, n, fused_k, kind, grade, ninputs, nouts)
  std::printf("#define FQ_GEN_ELMAT_LIST(X) \\\n");
  for (const Entry& e : entries)
    std::printf("  X(%s, %d, %d, %d, %d, %d, %d) \\\n", e.name.c_str(), e.n, e.fused_k, e.kind, e.grade, e.ninputs,
                e.nouts);
  std::printf("\n");
  // staged block sets of the tile-fused kernel: hodge_blocks(k) and every single block
  std::vector<SetEntry> sets;
  for (int n = 1; n <= 3; ++n) {
    for (int k = 0; k <= n; ++k) {
      SetEntry e{};
      e.fused_k = k, e.kind = -1, e.grade = k;
      if (emit_set("fq_set_n" + std::to_string(n) + "_hodge" + std::to_string(k), n, hodge_blocks(k), e)) sets.push_back(e);
    }
    for (int k = 0; k <= n + 1; ++k)
      for (int kind = 0; kind < 4; ++kind) {
        if (kind != KIND_MASS && k == 0) continue;
        if (k == n + 1 && kind != KIND_DIF_BOTH) continue;
        SetEntry e{};
        e.fused_k = -1, e.kind = kind, e.grade = k;
        if (emit_set("fq_set_n" + std::to_string(n) + "_kind" + std::to_string(kind) + "_k" + std::to_string(k), n, {{kind, k}}, e))
          sets.push_back(e);
      }
  }
  // X-macro list: (function, n, fused_k, kind, grade, ninputs)
  std::printf("#define FQ_GEN_SET_LIST(X) \\\n");
  for (const SetEntry& e : sets)
    std::printf("  X(%s, %d, %d, %d, %d, %d) \\\n", e.name.c_str(), e.n, e.fused_k, e.kind, e.grade, e.ninputs);
  std::printf("\n");
  return 0;
}
