// common.cuh — shared plumbing of libformoniq_b200: error handling, context,
// device buffers, handle structs.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/formoniq_b200.h"
#include "host_widen.hpp"
#include "tape.hpp"

namespace fq {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& m);

#define FQ_CUDA(expr)                                                                                       \
  do {                                                                                                      \
    cudaError_t err__ = (expr);                                                                             \
    if (err__ != cudaSuccess)                                                                               \
      throw ::fq::Error(FQ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__) + " (" + __FILE__ + \
                                         ":" + std::to_string(__LINE__) + ")");                             \
  } while (0)

#define FQ_REQUIRE(cond, msg)                                  \
  do {                                                         \
    if (!(cond)) throw ::fq::Error(FQ_ERR_INVALID, (msg));     \
  } while (0)

// Wrap an extern "C" body: exceptions never cross the ABI.
#define FQ_API_BEGIN try {
#define FQ_API_END                                  \
  return FQ_OK;                                     \
  }                                                 \
  catch (const ::fq::Error& e) {                    \
    ::fq::set_last_error(e.what());                 \
    return e.code;                                  \
  }                                                 \
  catch (const std::exception& e) {                 \
    ::fq::set_last_error(e.what());                 \
    return FQ_ERR_INVALID;                          \
  }

// Device memory comes from a process-wide caching allocator (capi.cu): multi-GB cudaMalloc/cudaFree calls cost tens of
// milliseconds each and synchronise the device, which dominated the symbolic phase and the one-shot assembly.
void* dev_alloc(size_t bytes);
void dev_free(void* p);
void dev_cache_trim();

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  bool owned = true;  // false: wraps caller-owned device memory
  DevBuf() = default;
  explicit DevBuf(size_t n_) { alloc(n_); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), owned(o.owned) { o.p = nullptr, o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p, n = o.n, owned = o.owned;
      o.p = nullptr, o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n_) {
    release();
    owned = true;
    n = n_;
    if (n) p = static_cast<T*>(dev_alloc(n * sizeof(T)));
  }
  void release() {
    if (p && owned) dev_free(p);
    p = nullptr, n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

}  // namespace fq

struct fq_span {
  int name_id;
  cudaEvent_t a, b;
};
struct fq_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  int64_t launches = 0;
  fq::DevBuf<double> reduce_scratch;  // two-stage reductions
  double* host_scalar = nullptr;      // pinned
  fq::DevBuf<int> d_flag_timeout;     // raised by a peer-flag wait that gave up
  // asynchronous result downloads (fq_csr_download_async): a second stream so that the D2H copies of one matrix
  // overlap the assembly of the next; the widened staging pieces live until fq_ctx_wait_downloads
  cudaStream_t copy_stream = nullptr;
  std::vector<fq::DevBuf<uint64_t>> pending_staging;
  // index arrays cross PCIe as u32 and are widened to usize in place on host threads (host_widen.hpp); created on the
  // first asynchronous download, absent when host widening is off (then the device widens and u64 crosses the bus)
  fq::HostWidener* widener = nullptr;
  bool widener_probed = false;
  // optional per-kernel timing (CUDA events on the launching stream)
  bool timing = false;
  std::vector<std::string> span_names;
  std::vector<fq_span> spans;
};

inline void fq_count_launch(fq_ctx* ctx, int n = 1) { ctx->launches += n; }

namespace fq {
struct TilePlan;  // tile.cu: plan of the tile-fused numeric assembly
// Brackets the kernels launched in its scope with two events when timing is on.
struct ScopedSpan {
  fq_ctx* ctx;
  fq_span sp{};
  bool on;
  ScopedSpan(fq_ctx* c, const char* name) : ctx(c), on(c->timing) {
    if (!on) return;
    int id = -1;
    for (size_t i = 0; i < c->span_names.size(); ++i)
      if (c->span_names[i] == name) id = int(i);
    if (id < 0) {
      c->span_names.push_back(name);
      id = int(c->span_names.size()) - 1;
    }
    sp.name_id = id;
    cudaEventCreate(&sp.a);
    cudaEventCreate(&sp.b);
    cudaEventRecord(sp.a, c->stream);
  }
  ~ScopedSpan() {
    if (!on) return;
    cudaEventRecord(sp.b, ctx->stream);
    ctx->spans.push_back(sp);
  }
};
}  // namespace fq

struct fq_mesh {
  int dim = 0;
  size_t ncells = 0;                 // cells held (owned + halo layer)
  size_t nowned_cells = 0;           // cells of the owned box layers (== ncells for full meshes)
  size_t cell_offset = 0;            // global index of local cell 0 (slab meshes)
  std::vector<size_t> nsimplices;    // global counts per grade
  // cell_faces[j]: [ncells][nlocal(dim,j)] global face ids (u32)
  std::vector<fq::DevBuf<uint32_t>> cell_faces;
  // signed squared edge lengths for edge ids [edge_lo, edge_lo + lengths.n)
  fq::DevBuf<double> lengths;
  size_t edge_lo = 0;
  // id ranges referenced by the local cells, per grade (global for full meshes)
  std::vector<size_t> id_lo, id_hi;
  // rows owned under the owner-computes slab partition (== id range of the
  // simplices whose top vertex lies in the slab's own vertex layers)
  std::vector<size_t> own_lo, own_hi;
  // vertex clustering for the tile-fused assembly (tile.cu): tile id of the held
  // vertices [vtile_lo, vtile_lo + vertex_tile.n); empty when not clustered
  fq::DevBuf<uint32_t> vertex_tile;
  size_t vtile_lo = 0;
  size_t ntiles = 0;
  // cells are numbered (box, type) with `cell_type_period` types per box (Kuhn grids: dim!); 0 = unknown.
  // The tile kernel numbers a tile's cells type-major so that same-type gathers hit distinct banks.
  int cell_type_period = 0;
  bool cluster_tried = false;  // generic meshes are clustered lazily, when a tile plan is first built
  bool cluster_generic = false;  // tiles = runs of consecutive vertices sized by weight (tile.cu: tile_cluster_generic)
  float cluster_scale = 0.0f;    // weight scale of the generic clustering; raised when a tile exceeds a budget
};

struct fq_vec {
  fq::DevBuf<double> d;
  void* ipc_base = nullptr;  // non-null: d.p is a CUDA IPC mapping of a peer's vector (closed on destroy)
};

// One block's symbolic data: structural pattern + gather lists.
struct fq_csr {
  size_t nrows = 0, ncols = 0;     // global shape
  size_t row_begin = 0, row_end = 0;  // rows held (local row r <-> global row_begin + r)
  // active pattern (what the consumer sees)
  size_t nnz = 0;
  fq::DevBuf<uint32_t> row_ptr;  // [nrows_local+1]
  fq::DevBuf<uint32_t> col_idx;  // [nnz] global column ids
  fq::DevBuf<double> values;     // [nnz]
  // ---- assembly plan (empty for uploaded matrices)
  bool has_plan = false;
  int kind = 0, grade = 0, dim = 0;
  int el_rows = 0, el_cols = 0;
  size_t ncells = 0;
  size_t s_nnz = 0;                 // structural nnz
  fq::DevBuf<uint32_t> s_row_ptr;   // [nrows_local+1]
  fq::DevBuf<uint32_t> s_col_idx;   // [s_nnz]
  fq::DevBuf<uint32_t> contrib_ptr; // [s_nnz+1] segments of contrib_src
  fq::DevBuf<uint32_t> contrib_src; // [ncontrib] cell*T+slot, sorted by (nnz, cell)
  size_t ncontrib = 0;
  fq::DevBuf<double> s_values;      // [s_nnz] structural values (scratch when dropping)
  fq::DevBuf<uint8_t> keep;         // [s_nnz] "some contribution != 0.0" (galerkin.rs:173)
  fq::DevBuf<double> slab;          // [ncells][el_rows*el_cols] element slab, persistent
  fq::DevBuf<uint32_t> pos;         // [s_nnz+1] structural -> compacted index (exclusive scan of keep)
  fq::DevBuf<int> d_changed;        // device flag: classification differs from the cached pattern
  fq::DevBuf<uint32_t> gather_blocks;  // stream blocks over contrib_ptr (K3)
  size_t ngather_blocks = 0;
  bool dropped = false;
  bool compact_valid = false;       // pos / keep / compacted pattern describe the last geometry
  bool pattern_valid = false;
  int64_t assembly_bytes = 0;
  int64_t assembly_shared_bytes = 0;  // inputs a fused launch reads once for all its blocks
  // SpMV row blocks (CSR-stream)
  fq::DevBuf<uint32_t> rowblocks;
  size_t nrowblocks = 0;
  bool spmv_ready = false;
  // fused halo-exchange SpMV: {free_lo, free_hi, ext0, ext1} for the owned range below, copy-completion counter
  fq::DevBuf<uint32_t> peer_split;
  fq::DevBuf<unsigned long long> peer_counter;
  size_t peer_own_lo = 0, peer_own_hi = 0, peer_launches = 0;
  // Jacobi (inverse diagonal), built on demand
  fq::DevBuf<double> inv_diag;
  // tile-fused numeric path (shared by the blocks assembled together)
  std::shared_ptr<fq::TilePlan> tile_plan;
  int tile_refused = 0;
  int slab_passes = 0;  // numeric passes done through the slab path since the symbolic phase
  // K2 (the global-sort symbolic phase feeding the slab path) runs lazily: the tile plan builder produces the
  // structural pattern itself, so a matrix that only ever goes through the fused kernel never pays for K2
  bool k2_done = false;
  double plan_build_ms = 0.0;  // device time of the last tile plan build (symbolic + plan), 0 when none
  size_t plan_cell_visits = 0; // cell visits (cells x tiles touching them) of the tile plan: element tapes per fused launch
};
