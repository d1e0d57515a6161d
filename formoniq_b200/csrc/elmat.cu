// elmat.cu — K1: batched per-cell element matrices.
//
// Two paths share one arithmetic definition (the tape of tape.hpp):
//   * generated straight-line kernels (elmat_gen.cuh) for dim <= 3, thread per
//     cell, every intermediate in registers;
//   * a generic tape interpreter for any runtime (dim, grade): thread per cell,
//     register file in a coalesced global scratch slab, for dim >= 4 preceded
//     by a generic geometry stage (closed-form 4x4 inverse, LU beyond).
// All FP64 arithmetic is explicit round-to-nearest (no FMA contraction), so the
// zero / non-zero classification of every entry matches the reference's
// (galerkin.rs:173 makes the CSR pattern depend on it).
#include <map>
#include <mutex>

#include "elmat_gen.cuh"
#include "geometry.cuh"
#include "internal.hpp"

namespace fq {

// ------------------------------------------------------------------ sinks
// A warp evaluates 32 consecutive cells, one per lane.  Each block's entries
// are staged in shared memory ([entry][lane], padded to 33 so both the
// lane-major writes and the cell-major reads are bank-conflict free) and then
// flushed as one contiguous, fully coalesced run of 32*T doubles of that
// block's cell-major slab  slab_b[(cell) * T_b + entry].
constexpr int kStagePad = 33;
constexpr int kMaxBlocks = 4;
struct SlabPtrs {
  double* p[kMaxBlocks];
};
struct WarpStageSink {
  double* __restrict__ stage;  // this warp's staging area [maxT][33]
  SlabPtrs slabs;              // per block: slab base + first_cell_of_this_warp * T_b is added in flush
  size_t cell0;                // first cell (slab-relative) of this warp's group
  int lane, nvalid;
  template <int B, int E>
  __device__ __forceinline__ void put(double v) const {
    stage[E * kStagePad + lane] = v;
  }
  template <int B, int TB>
  __device__ __forceinline__ void flush() const {
    __syncwarp();
    double* __restrict__ g = slabs.p[B] + cell0 * size_t(TB);
    const int total = nvalid * TB;
#pragma unroll 4
    for (int k = lane; k < total; k += 32) {
      const int c = k / TB, e = k - c * TB;
      g[k] = stage[e * kStagePad + c];
    }
    __syncwarp();
  }
};

template <class Fn, int NE>
__global__ void __launch_bounds__(128) elmat_gen_kernel(Fn fn, const uint32_t* __restrict__ cell_edges,
                                                         const double* __restrict__ lengths, uint32_t edge_lo,
                                                         size_t c0, size_t ncells, int max_t, SlabPtrs slabs) {
  extern __shared__ double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* stage = stage_all + size_t(warp) * size_t(max_t) * kStagePad;
  const size_t ngroups = (ncells + 31) / 32;
  const size_t wstride = size_t(gridDim.x) * (blockDim.x >> 5);
  for (size_t grp = size_t(blockIdx.x) * (blockDim.x >> 5) + warp; grp < ngroups; grp += wstride) {
    const size_t first = grp * 32;
    const int nvalid = int(ncells - first < 32 ? ncells - first : 32);
    // lanes past the end recompute the last valid cell; their results are never flushed
    const size_t i = first + size_t(lane < nvalid ? lane : nvalid - 1);
    const uint32_t* ce = cell_edges + (c0 + i) * NE;
    double s[NE > 0 ? NE : 1];
#pragma unroll
    for (int e = 0; e < NE; ++e) s[e] = __ldg(lengths + (ce[e] - edge_lo));
    WarpStageSink sink{stage, slabs, first, lane, nvalid};
    fn(s, sink);
  }
}

// dim >= 4: the tape's inputs are g^-1 (row-major) and the volume; the generic geometry stage runs in the same thread
template <class Fn, int N>
__global__ void __launch_bounds__(128) elmat_gen_geo_kernel(Fn fn, const uint32_t* __restrict__ cell_edges,
                                                             const double* __restrict__ lengths, uint32_t edge_lo,
                                                             size_t c0, size_t ncells, int max_t, SlabPtrs slabs,
                                                             int* __restrict__ err) {
  extern __shared__ double stage_all[];
  constexpr int NE = N * (N + 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* stage = stage_all + size_t(warp) * size_t(max_t) * kStagePad;
  const size_t ngroups = (ncells + 31) / 32;
  const size_t wstride = size_t(gridDim.x) * (blockDim.x >> 5);
  for (size_t grp = size_t(blockIdx.x) * (blockDim.x >> 5) + warp; grp < ngroups; grp += wstride) {
    const size_t first = grp * 32;
    const int nvalid = int(ncells - first < 32 ? ncells - first : 32);
    const size_t i = first + size_t(lane < nvalid ? lane : nvalid - 1);
    const uint32_t* ce = cell_edges + (c0 + i) * NE;
    double len[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) len[e] = __ldg(lengths + (ce[e] - edge_lo));
    double s[N * N + 1];
    if (!geometry_generic(N, len, s, &s[N * N])) {
      if (err) atomicExch(err, 1);
#pragma unroll
      for (int e = 0; e <= N * N; ++e) s[e] = 0.0;
    }
    WarpStageSink sink{stage, slabs, first, lane, nvalid};
    fn(s, sink);
  }
}

#define FQ_DECLARE_FN(fn, n, fk, kind, grade, nin, nout)                                        \
  struct Fn_##fn {                                                                              \
    template <class S>                                                                          \
    __device__ __forceinline__ void operator()(const double* __restrict__ s, S& sink) const {   \
      fn(s, sink);                                                                              \
    }                                                                                           \
  };
FQ_GEN_ELMAT_LIST(FQ_DECLARE_FN)
#undef FQ_DECLARE_FN

struct GenEntry {
  int n, fused_k, kind, grade, nin, nout;
  void (*launch)(fq_ctx*, const fq_mesh*, size_t, size_t, int, const SlabPtrs&, int*);
};

template <class Fn, int N>
static void launch_gen_geo(fq_ctx* ctx, const fq_mesh* mesh, size_t c0, size_t c1, int max_t, const SlabPtrs& slabs, int* d_err) {
  const size_t nc = c1 - c0;
  if (nc == 0) return;
  const int block = 128;
  const size_t smem = size_t(block / 32) * size_t(max_t) * kStagePad * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    FQ_CUDA(cudaFuncSetAttribute(elmat_gen_geo_kernel<Fn, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  const int grid = grid_for((nc + 31) / 32 * 32, block, ctx->sm_count, 4);
  elmat_gen_geo_kernel<Fn, N><<<grid, block, smem, ctx->stream>>>(Fn{}, mesh->cell_faces[1].p, mesh->lengths.p,
                                                                  uint32_t(mesh->edge_lo), c0, nc, max_t, slabs, d_err);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}
template <class Fn, int NE>
static void launch_gen(fq_ctx* ctx, const fq_mesh* mesh, size_t c0, size_t c1, int max_t, const SlabPtrs& slabs) {
  const size_t nc = c1 - c0;
  if (nc == 0) return;
  const int block = 128;
  const size_t smem = size_t(block / 32) * size_t(max_t) * kStagePad * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    FQ_CUDA(cudaFuncSetAttribute(elmat_gen_kernel<Fn, NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  const int grid = grid_for((nc + 31) / 32 * 32, block, ctx->sm_count, 4);
  elmat_gen_kernel<Fn, NE><<<grid, block, smem, ctx->stream>>>(Fn{}, mesh->cell_faces[1].p, mesh->lengths.p,
                                                               uint32_t(mesh->edge_lo), c0, nc, max_t, slabs);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

#define FQ_ENTRY(fn, n, fk, kind, grade, nin, nout)                                                        \
  GenEntry{n, fk, kind, grade, nin, nout,                                                                  \
           [](fq_ctx* ctx, const fq_mesh* mesh, size_t c0, size_t c1, int max_t, const SlabPtrs& slabs, int* d_err) { \
             if (n >= 4)                                                                                   \
               launch_gen_geo<Fn_##fn, (n >= 4 ? n : 4)>(ctx, mesh, c0, c1, max_t, slabs, d_err);          \
             else                                                                                          \
               launch_gen<Fn_##fn, (n >= 4 ? 0 : nin)>(ctx, mesh, c0, c1, max_t, slabs);                   \
           }},
static const GenEntry g_entries[] = {FQ_GEN_ELMAT_LIST(FQ_ENTRY)};
#undef FQ_ENTRY

static const GenEntry* find_generated(int dim, const std::vector<BlockSpec>& blocks) {
  int fused_k = -2;
  if (blocks.size() == 4) {
    const auto hb = hodge_blocks(blocks[1].grade);
    bool same = true;
    for (int i = 0; i < 4; ++i) same = same && hb[i].kind == blocks[i].kind && hb[i].grade == blocks[i].grade;
    if (same) fused_k = blocks[1].grade;
  }
  for (const GenEntry& e : g_entries) {
    if (e.n != dim) continue;
    if (fused_k >= 0) {
      if (e.fused_k == fused_k) return &e;
    } else if (blocks.size() == 1 && e.fused_k < 0 && e.kind == blocks[0].kind &&
               (e.kind == KIND_LUMPED || e.grade == blocks[0].grade)) {
      return &e;
    }
  }
  return nullptr;
}

static int block_nouts(int dim, const BlockSpec& b) {
  int tg, rg;
  kind_grades(b.kind, b.grade, tg, rg);
  return nlocal(dim, tg) * nlocal(dim, rg);
}

bool elmat_has_generated(int dim, const std::vector<BlockSpec>& blocks) { return find_generated(dim, blocks) != nullptr; }

int elmat_nouts(int dim, const std::vector<BlockSpec>& blocks) {
  int total = 0;
  for (const BlockSpec& b : blocks) total += block_nouts(dim, b);
  return total;
}

// ------------------------------------------------------------------ generic geometry (dim >= 4): geometry.cuh

// ------------------------------------------------------------------ interpreter
__global__ void __launch_bounds__(128) elmat_interp_kernel(const uint4* __restrict__ ops, int nops,
                                                            const double* __restrict__ consts, int ninputs,
                                                            int inputs_are_lengths, int n, int ne,
                                                            const uint32_t* __restrict__ cell_edges,
                                                            const double* __restrict__ lengths, uint32_t edge_lo,
                                                            size_t c0, size_t ncells, double* __restrict__ scratch,
                                                            int nouts, double* __restrict__ out, int* __restrict__ err) {
  const size_t nthreads = size_t(gridDim.x) * blockDim.x;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  double* reg = scratch + tid;  // reg[r * nthreads]
  for (size_t i = tid; i < ncells; i += nthreads) {
    const uint32_t* ce = cell_edges + (c0 + i) * size_t(ne);
    if (inputs_are_lengths) {
      for (int e = 0; e < ne; ++e) reg[size_t(e) * nthreads] = lengths[ce[e] - edge_lo];
    } else {
      double s[kMaxDim * (kMaxDim + 1) / 2];
      double gi[kMaxDim * kMaxDim];
      double vol;
      for (int e = 0; e < ne; ++e) s[e] = lengths[ce[e] - edge_lo];
      if (!geometry_generic(n, s, gi, &vol)) {
        if (err) atomicExch(err, 1);
        vol = 0.0;
        for (int e = 0; e < n * n; ++e) gi[e] = 0.0;
      }
      for (int e = 0; e < n * n; ++e) reg[size_t(e) * nthreads] = gi[e];
      reg[size_t(n * n) * nthreads] = vol;
    }
    double* o = out + i * size_t(nouts);
    for (int p = 0; p < nops; ++p) {
      const uint4 op = __ldg(ops + p);
      switch (op.x) {
        case OP_ADD: reg[size_t(op.y) * nthreads] = __dadd_rn(reg[size_t(op.z) * nthreads], reg[size_t(op.w) * nthreads]); break;
        case OP_SUB: reg[size_t(op.y) * nthreads] = __dsub_rn(reg[size_t(op.z) * nthreads], reg[size_t(op.w) * nthreads]); break;
        case OP_MUL: reg[size_t(op.y) * nthreads] = __dmul_rn(reg[size_t(op.z) * nthreads], reg[size_t(op.w) * nthreads]); break;
        case OP_MULC: reg[size_t(op.y) * nthreads] = __dmul_rn(reg[size_t(op.z) * nthreads], __ldg(consts + op.w)); break;
        case OP_DIV: reg[size_t(op.y) * nthreads] = __ddiv_rn(reg[size_t(op.z) * nthreads], reg[size_t(op.w) * nthreads]); break;
        case OP_SQRTABS: reg[size_t(op.y) * nthreads] = __dsqrt_rn(fabs(reg[size_t(op.z) * nthreads])); break;
        case OP_LOADC: reg[size_t(op.y) * nthreads] = __ldg(consts + op.w); break;
        case OP_STORE: o[op.y] = reg[size_t(op.z) * nthreads]; break;
        case OP_STOREN: o[op.y] = -reg[size_t(op.z) * nthreads]; break;
        case OP_STOREC: o[op.y] = __ldg(consts + op.w); break;
      }
    }
  }
}

struct TapeDev {
  Tape tape;
  DevBuf<uint4> ops;
  DevBuf<double> consts;
};
static std::mutex g_tape_mu;
static std::map<std::vector<int>, TapeDev*> g_tapes;

static TapeDev* get_tape(fq_ctx* ctx, int dim, const std::vector<BlockSpec>& blocks) {
  std::vector<int> key{ctx->device, dim};
  for (const BlockSpec& b : blocks) key.push_back(b.kind), key.push_back(b.grade);
  std::lock_guard<std::mutex> lk(g_tape_mu);
  auto it = g_tapes.find(key);
  if (it != g_tapes.end()) return it->second;
  TapeDev* td = new TapeDev;
  td->tape = build_tape(dim, blocks, nullptr);
  std::vector<uint4> hops;
  for (const TapeOp& o : td->tape.ops) hops.push_back(make_uint4(o.op, o.d, o.a, o.b));
  td->ops.alloc(hops.size() ? hops.size() : 1);
  td->consts.alloc(td->tape.consts.size() ? td->tape.consts.size() : 1);
  FQ_CUDA(cudaDeviceSynchronize());  // recycled blocks: order the blocking copies after everything in flight
  if (!hops.empty()) FQ_CUDA(cudaMemcpy(td->ops.p, hops.data(), hops.size() * sizeof(uint4), cudaMemcpyHostToDevice));
  if (!td->tape.consts.empty())
    FQ_CUDA(cudaMemcpy(td->consts.p, td->tape.consts.data(), td->tape.consts.size() * sizeof(double),
                       cudaMemcpyHostToDevice));
  g_tapes[key] = td;
  return td;
}

// Runs the interpreter for ONE block into its cell-major slab.
static void interp_block(fq_ctx* ctx, const fq_mesh* mesh, const BlockSpec& blk, size_t c0, size_t c1, double* d_out,
                         int* d_err) {
  const int dim = mesh->dim;
  const int nouts = block_nouts(dim, blk);
  if (nouts == 0) return;
  TapeDev* td = get_tape(ctx, dim, {blk});
  const Tape& t = td->tape;
  FQ_REQUIRE(t.nouts <= nouts, "tape output count mismatch");
  const int block = 128;
  // bound the scratch slab: at most ~256 MB of interpreter registers
  size_t max_threads = (size_t(256) << 20) / (sizeof(double) * size_t(t.nregs > 0 ? t.nregs : 1));
  max_threads = max_threads / block * block;
  if (max_threads < size_t(block)) max_threads = block;
  size_t want = ((c1 - c0) + block - 1) / block * block;
  const size_t cap = size_t(ctx->sm_count) * 8 * block;
  if (want > cap) want = cap;
  if (want > max_threads) want = max_threads;
  DevBuf<double> scratch(want * size_t(t.nregs > 0 ? t.nregs : 1));
  const int ne = int(binom(dim + 1, 2));
  elmat_interp_kernel<<<int(want / block), block, 0, ctx->stream>>>(
      td->ops.p, int(t.ops.size()), td->consts.p, t.ninputs, t.inputs_are_lengths ? 1 : 0, dim, ne,
      mesh->cell_faces[1].p, mesh->lengths.p, uint32_t(mesh->edge_lo), c0, c1 - c0, scratch.p, nouts, d_out, d_err);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));  // scratch is freed on return
}

void elmat_to_slabs(fq_ctx* ctx, const fq_mesh* mesh, const std::vector<BlockSpec>& blocks, size_t c0, size_t c1,
                    bool use_generated, double* const* d_outs, int* d_err) {
  const int dim = mesh->dim;
  FQ_REQUIRE(dim >= 1 && dim <= kMaxDim, "element kernels support 1 <= dim <= 10");
  FQ_REQUIRE(c0 <= c1 && c1 <= mesh->ncells, "cell range out of bounds");
  FQ_REQUIRE(int(blocks.size()) <= kMaxBlocks, "too many fused blocks");
  if (c1 == c0 || elmat_nouts(dim, blocks) == 0) return;
  if (use_generated) {
    if (const GenEntry* e = find_generated(dim, blocks)) {
      SlabPtrs sp{};
      int max_t = 1;
      for (size_t b = 0; b < blocks.size(); ++b) {
        sp.p[b] = d_outs[b];
        max_t = std::max(max_t, block_nouts(dim, blocks[b]));
      }
      e->launch(ctx, mesh, c0, c1, max_t, sp, d_err);
      return;
    }
  }
  for (size_t b = 0; b < blocks.size(); ++b) interp_block(ctx, mesh, blocks[b], c0, c1, d_outs[b], d_err);
}

}  // namespace fq
