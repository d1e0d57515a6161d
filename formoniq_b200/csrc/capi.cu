// capi.cu — the extern "C" boundary declared in include/formoniq_b200.h.
#include <cstring>
#include <limits>
#include <map>
#include <mutex>

#include "internal.hpp"
#include "kuhn.hpp"

namespace fq {
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }

// ---- caching device allocator ------------------------------------------------
// Blocks are rounded up (512 B below 1 MB, 2 MB above) and kept per (device, size) when released; reuse is
// stream-ordered because every context enqueues its work on one stream and the API calls that release large
// temporaries synchronise first.  cudaMalloc failures trim the cache and retry once.
namespace {
struct DevCache {
  std::mutex mu;
  std::map<std::pair<int, size_t>, std::vector<void*>> free_blocks;
  std::map<void*, std::pair<int, size_t>> live;
  size_t cached_bytes = 0;
};
DevCache& dev_cache() {
  static DevCache* c = new DevCache;  // leaked on purpose: outlives every static destructor that may release buffers
  return *c;
}
size_t round_size(size_t b) {
  const size_t g = b < (size_t(1) << 20) ? 512 : (size_t(2) << 20);
  return (b + g - 1) / g * g;
}
}  // namespace
void dev_cache_trim() {
  DevCache& c = dev_cache();
  std::lock_guard<std::mutex> lk(c.mu);
  for (auto& kv : c.free_blocks)
    for (void* p : kv.second) cudaFree(p);
  c.free_blocks.clear();
  c.cached_bytes = 0;
}
void* dev_alloc(size_t bytes) {
  DevCache& c = dev_cache();
  int dev = 0;
  FQ_CUDA(cudaGetDevice(&dev));
  const size_t sz = round_size(bytes);
  {
    std::lock_guard<std::mutex> lk(c.mu);
    // best fit: the smallest cached block that is large enough and wastes at most a quarter
    for (auto it = c.free_blocks.lower_bound({dev, sz}); it != c.free_blocks.end() && it->first.first == dev; ++it) {
      if (it->first.second > sz + sz / 4 + (size_t(2) << 20)) break;
      if (it->second.empty()) continue;
      void* p = it->second.back();
      it->second.pop_back();
      c.cached_bytes -= it->first.second;
      c.live[p] = it->first;
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t err = cudaMalloc(&p, sz);
  if (err != cudaSuccess) {
    cudaGetLastError();
    dev_cache_trim();
    err = cudaMalloc(&p, sz);
  }
  if (err != cudaSuccess)
    throw Error(FQ_ERR_CUDA, std::string("cudaMalloc of ") + std::to_string(sz) + " bytes: " + cudaGetErrorString(err));
  std::lock_guard<std::mutex> lk(c.mu);
  c.live[p] = {dev, sz};
  return p;
}
void dev_free(void* p) {
  if (!p) return;
  DevCache& c = dev_cache();
  std::lock_guard<std::mutex> lk(c.mu);
  auto it = c.live.find(p);
  if (it == c.live.end()) {
    cudaFree(p);
    return;
  }
  const auto key = it->second;
  c.live.erase(it);
  // keep at most 96 GB cached per process; beyond that hand the block back to the driver
  if (c.cached_bytes + key.second > (size_t(96) << 30)) {
    cudaFree(p);
    return;
  }
  c.free_blocks[key].push_back(p);
  c.cached_bytes += key.second;
}

__global__ void widen_u32_kernel(const uint32_t* __restrict__ in, size_t n, uint64_t* __restrict__ out) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}
// `limit`: every value must be below it (face ids against the simplex count of their grade): a malformed table
// would otherwise become out-of-bounds device reads in every later kernel
__global__ void narrow_u64_kernel(const uint64_t* __restrict__ in, size_t n, uint32_t* __restrict__ out, uint64_t limit,
                                  int* __restrict__ overflow) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t v = in[i];
    if ((v >> 32) || v >= limit) *overflow = 1;
    out[i] = uint32_t(v);
  }
}

// host u64 array -> device u32 array (chunked through a staging buffer)
static void upload_narrow(fq_ctx* ctx, const uint64_t* host, size_t n, DevBuf<uint32_t>& dst,
                          uint64_t limit = uint64_t(1) << 32) {
  dst.alloc(n ? n : 1);
  if (!n) return;
  const size_t chunk = size_t(1) << 24;
  DevBuf<uint64_t> stage(std::min(n, chunk));
  DevBuf<int> ovf(1);
  FQ_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int), ctx->stream));
  for (size_t off = 0; off < n; off += chunk) {
    const size_t m = std::min(chunk, n - off);
    FQ_CUDA(cudaMemcpyAsync(stage.p, host + off, m * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    narrow_u64_kernel<<<grid_for(m, 256, ctx->sm_count), 256, 0, ctx->stream>>>(stage.p, m, dst.p + off, limit, ovf.p);
    fq_count_launch(ctx);
  }
  int h = 0;
  FQ_CUDA(cudaMemcpyAsync(&h, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  FQ_REQUIRE(h == 0, "index out of range (face id >= the simplex count of its grade, or beyond 32 bits)");
}
// device u32 array -> host u64 array: widened on the device in pieces of up to 2 GB, each copied back with one
// cudaMemcpyAsync (PCIe-bound when the destination is pinned or registered host memory)
static void download_widen(fq_ctx* ctx, const uint32_t* dev, size_t n, uint64_t* host) {
  if (!n) return;
  const size_t chunk = size_t(1) << 28;
  DevBuf<uint64_t> stage[2];
  stage[0].alloc(std::min(n, chunk));
  if (n > chunk) stage[1].alloc(std::min(n - chunk, chunk));
  cudaEvent_t done[2];
  FQ_CUDA(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
  FQ_CUDA(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
  int k = 0;
  for (size_t off = 0; off < n; off += chunk, k ^= 1) {
    const size_t m = std::min(chunk, n - off);
    if (off >= 2 * chunk) FQ_CUDA(cudaEventSynchronize(done[k]));  // the staging piece is free again
    widen_u32_kernel<<<grid_for(m, 256, ctx->sm_count), 256, 0, ctx->stream>>>(dev + off, m, stage[k].p);
    fq_count_launch(ctx);
    FQ_CUDA(cudaMemcpyAsync(host + off, stage[k].p, m * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    FQ_CUDA(cudaEventRecord(done[k], ctx->stream));
  }
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaEventDestroy(done[0]);
  cudaEventDestroy(done[1]);
}
}  // namespace fq

using namespace fq;

extern "C" {

const char* fq_last_error(void) { return g_last_error.c_str(); }

int fq_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int fq_ctx_create(int device, fq_ctx** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(out, "null output");
  int n = 0;
  FQ_CUDA(cudaGetDeviceCount(&n));
  if (n <= 0) throw Error(FQ_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
  FQ_REQUIRE(device >= 0 && device < n, "device index out of range");
  FQ_CUDA(cudaSetDevice(device));
  fq_ctx* ctx = new fq_ctx;
  ctx->device = device;
  FQ_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->own_stream = true;
  cudaDeviceProp prop;
  FQ_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  FQ_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->host_scalar), 64));
  *out = ctx;
  FQ_API_END
}
int fq_ctx_destroy(fq_ctx* ctx) {
  FQ_API_BEGIN
  if (ctx) {
    if (ctx->copy_stream) {
      cudaStreamSynchronize(ctx->copy_stream);
      cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->widener) {
      ctx->widener->wait_idle();
      delete ctx->widener;
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->host_scalar) cudaFreeHost(ctx->host_scalar);
    delete ctx;
  }
  FQ_API_END
}
int fq_ctx_set_stream(fq_ctx* ctx, void* cuda_stream) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx, "null context");
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->own_stream = false;
  FQ_API_END
}
int fq_ctx_synchronize(fq_ctx* ctx) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx, "null context");
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  FQ_API_END
}
int64_t fq_ctx_launch_count(const fq_ctx* ctx) { return ctx ? ctx->launches : 0; }
int fq_device_cache_trim(void) {
  FQ_API_BEGIN
  cudaDeviceSynchronize();
  dev_cache_trim();
  FQ_API_END
}
int fq_ctx_set_timing(fq_ctx* ctx, int on) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx, "null context");
  ctx->timing = on != 0;
  FQ_API_END
}
int fq_ctx_timing_report(fq_ctx* ctx, char* buf, size_t buflen) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && buf && buflen > 2, "bad argument");
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<double> ms(ctx->span_names.size(), 0.0);
  std::vector<int64_t> cnt(ctx->span_names.size(), 0);
  for (fq_span& sp : ctx->spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) {
      ms[size_t(sp.name_id)] += double(t);
      cnt[size_t(sp.name_id)] += 1;
    }
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  ctx->spans.clear();
  std::string out = "{";
  for (size_t i = 0; i < ms.size(); ++i) {
    char tmp[256];
    std::snprintf(tmp, sizeof tmp, "%s\"%s\": {\"ms\": %.6f, \"count\": %lld}", i ? ", " : "", ctx->span_names[i].c_str(),
                  ms[i], static_cast<long long>(cnt[i]));
    out += tmp;
  }
  out += "}";
  FQ_REQUIRE(out.size() + 1 <= buflen, "timing report buffer too small");
  std::memcpy(buf, out.c_str(), out.size() + 1);
  FQ_API_END
}

// ---------------------------------------------------------------- mesh
static int mesh_create_impl(fq_ctx* ctx, int dim, size_t ncells, const size_t* nsimplices, const uint64_t* const* cell_faces,
                            const double* edge_lengths_sq, const size_t* own_lo, const size_t* own_hi, fq_mesh** out);
int fq_mesh_create(fq_ctx* ctx, int dim, size_t ncells, const size_t* nsimplices, const uint64_t* const* cell_faces,
                   const double* edge_lengths_sq, fq_mesh** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(nsimplices && dim >= 1 && dim <= 10 && nsimplices[dim] == ncells, "nsimplices[dim] must equal ncells");
  return mesh_create_impl(ctx, dim, ncells, nsimplices, cell_faces, edge_lengths_sq, nullptr, nullptr, out);
  FQ_API_END
}
int fq_mesh_create_part(fq_ctx* ctx, int dim, size_t ncells_held, const size_t* nsimplices, const uint64_t* const* cell_faces,
                        const double* edge_lengths_sq, const size_t* own_lo, const size_t* own_hi, fq_mesh** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(nsimplices && own_lo && own_hi && dim >= 1 && dim <= 10, "null argument");
  FQ_REQUIRE(ncells_held <= nsimplices[dim], "more held cells than cells");
  for (int j = 0; j <= dim; ++j)
    FQ_REQUIRE(own_lo[j] <= own_hi[j] && own_hi[j] <= nsimplices[j], "owned id range outside the skeleton");
  return mesh_create_impl(ctx, dim, ncells_held, nsimplices, cell_faces, edge_lengths_sq, own_lo, own_hi, out);
  FQ_API_END
}
static int mesh_create_impl(fq_ctx* ctx, int dim, size_t ncells, const size_t* nsimplices, const uint64_t* const* cell_faces,
                            const double* edge_lengths_sq, const size_t* own_lo, const size_t* own_hi, fq_mesh** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && out && nsimplices && cell_faces, "null argument");
  FQ_REQUIRE(dim >= 1 && dim <= 10, "1 <= dim <= 10");
  FQ_REQUIRE(cell_faces[1] && edge_lengths_sq, "grade-1 faces and edge lengths are required");
  FQ_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<fq_mesh> m(new fq_mesh);
  m->dim = dim;
  m->ncells = ncells;
  m->nowned_cells = ncells;
  m->nsimplices.assign(nsimplices, nsimplices + dim + 1);
  m->cell_faces.resize(size_t(dim) + 1);
  m->id_lo.assign(size_t(dim) + 1, 0);
  m->id_hi.assign(nsimplices, nsimplices + dim + 1);
  m->own_lo = m->id_lo;
  m->own_hi = m->id_hi;
  if (own_lo && own_hi) {  // a rank's part: ids stay global, the held cells are a subset
    m->own_lo.assign(own_lo, own_lo + dim + 1);
    m->own_hi.assign(own_hi, own_hi + dim + 1);
    m->nowned_cells = own_hi[dim] - own_lo[dim];
  }
  for (int j = 0; j <= dim; ++j) {
    FQ_REQUIRE(nsimplices[j] < (size_t(1) << 32), "more than 2^32 simplices of one grade: not supported");
    if (cell_faces[j]) upload_narrow(ctx, cell_faces[j], ncells * size_t(nlocal(dim, j)), m->cell_faces[size_t(j)], nsimplices[j]);
  }
  m->lengths.alloc(nsimplices[1] ? nsimplices[1] : 1);
  m->edge_lo = 0;
  FQ_CUDA(cudaMemcpyAsync(m->lengths.p, edge_lengths_sq, nsimplices[1] * sizeof(double), cudaMemcpyHostToDevice,
                          ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = m.release();
  FQ_API_END
}

int fq_mesh_create_kuhn(fq_ctx* ctx, int dim, const size_t* shape, const double* vmin, const double* vmax,
                        const double* ambient_diag, double jitter, size_t slab_begin, size_t slab_end, fq_mesh** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && out && shape, "null argument");
  FQ_REQUIRE(dim >= 1 && dim <= 6, "Kuhn generator supports 1 <= dim <= 6");
  FQ_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<fq_mesh> m(new fq_mesh);
  kuhn_build_mesh(ctx, dim, shape, vmin, vmax, ambient_diag, jitter, slab_begin, slab_end, m.get());
  *out = m.release();
  FQ_API_END
}
int fq_mesh_destroy(fq_mesh* mesh) {
  delete mesh;
  return FQ_OK;
}
int fq_mesh_dim(const fq_mesh* mesh) { return mesh ? mesh->dim : -1; }
size_t fq_mesh_ncells(const fq_mesh* mesh) { return mesh ? mesh->ncells : 0; }
size_t fq_mesh_nsimplices(const fq_mesh* mesh, int grade) {
  if (!mesh || grade < 0 || grade > mesh->dim) return 0;
  return mesh->nsimplices[size_t(grade)];
}
size_t fq_mesh_nowned_cells(const fq_mesh* mesh) { return mesh ? mesh->nowned_cells : 0; }
int fq_mesh_owned_range(const fq_mesh* mesh, int grade, size_t* lo, size_t* hi) {
  FQ_API_BEGIN
  FQ_REQUIRE(mesh && grade >= 0 && grade <= mesh->dim, "bad argument");
  if (lo) *lo = mesh->own_lo[size_t(grade)];
  if (hi) *hi = mesh->own_hi[size_t(grade)];
  FQ_API_END
}
int fq_mesh_held_range(const fq_mesh* mesh, int grade, size_t* lo, size_t* hi) {
  FQ_API_BEGIN
  FQ_REQUIRE(mesh && grade >= 0 && grade <= mesh->dim, "bad argument");
  if (lo) *lo = mesh->id_lo[size_t(grade)];
  if (hi) *hi = mesh->id_hi[size_t(grade)];
  FQ_API_END
}
int fq_mesh_set_lengths(fq_ctx* ctx, fq_mesh* mesh, const double* edge_lengths_sq) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && edge_lengths_sq, "null argument");
  // host array covers all edges; the mesh keeps [edge_lo, edge_lo + n)
  FQ_CUDA(cudaMemcpyAsync(mesh->lengths.p, edge_lengths_sq + mesh->edge_lo,
                          (mesh->id_hi[1] - mesh->id_lo[1]) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  FQ_API_END
}
int fq_mesh_download_cell_faces(fq_ctx* ctx, const fq_mesh* mesh, int grade, uint64_t* out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && out && grade >= 0 && grade <= mesh->dim, "bad argument");
  FQ_REQUIRE(mesh->cell_faces[size_t(grade)].p, "grade not present");
  download_widen(ctx, mesh->cell_faces[size_t(grade)].p, mesh->ncells * size_t(nlocal(mesh->dim, grade)), out);
  FQ_API_END
}
int fq_mesh_download_lengths(fq_ctx* ctx, const fq_mesh* mesh, double* out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && out, "null argument");
  // writes the held range at its global position
  FQ_CUDA(cudaMemcpyAsync(out + mesh->edge_lo, mesh->lengths.p, (mesh->id_hi[1] - mesh->id_lo[1]) * sizeof(double),
                          cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  FQ_API_END
}

int fq_kuhn_counts(int dim, const size_t* shape, size_t* nsimplices) {
  FQ_API_BEGIN
  FQ_REQUIRE(shape && nsimplices, "null argument");
  const KuhnTables kt(dim);
  const uint32_t full = (1u << dim) - 1;
  for (int j = 0; j <= dim; ++j) {
    uint64_t total = 0;
    for (uint32_t B = 0; B <= full; ++B) {
      uint64_t nv = 1;
      for (int a = 0; a < dim; ++a) nv *= (B >> a & 1u) ? shape[a] : 1;
      total += nv * kt.grades[size_t(j)].cnt[B];
    }
    nsimplices[j] = size_t(total);
  }
  FQ_API_END
}

int fq_kuhn_slab_ranges(int dim, const size_t* shape, size_t slab_begin, size_t slab_end, int grade, size_t* out4) {
  FQ_API_BEGIN
  FQ_REQUIRE(shape && out4 && grade >= 0 && grade <= dim, "bad argument");
  FQ_REQUIRE(slab_begin < slab_end && slab_end <= shape[dim - 1], "invalid slab range");
  const KuhnTables kt(dim);
  const KuhnGrid g(dim, shape);
  // number of simplices of `grade` whose top vertex lies in vertex layers [0, z)
  auto below = [&](uint64_t z) {
    const uint32_t full = (1u << dim) - 1;
    uint64_t total = 0;
    for (uint32_t B = 0; B <= full; ++B) {
      uint64_t nv = 1;
      for (int a = 0; a + 1 < dim; ++a) nv *= (B >> a & 1u) ? shape[a] : 1;
      // last axis: layers [0, z) contain 1 layer with coordinate 0 (if z > 0) and z-1 with coordinate >= 1
      const uint64_t last = (B >> (dim - 1) & 1u) ? (z > 0 ? z - 1 : 0) : (z > 0 ? 1 : 0);
      total += nv * last * kt.grades[size_t(grade)].cnt[B];
    }
    return total;
  };
  const size_t halo_end = slab_end < shape[dim - 1] ? slab_end + 1 : slab_end;
  out4[0] = size_t(below(slab_begin));                                  // held_lo
  out4[1] = size_t(slab_begin == 0 ? 0 : below(slab_begin + 1));        // own_lo
  out4[2] = size_t(below(slab_end + 1));                                // own_hi
  out4[3] = size_t(below(halo_end + 1));                                // held_hi
  FQ_API_END
}

int fq_kuhn_cell_faces_host(int dim, const size_t* shape, int grade, uint64_t* out) {
  FQ_API_BEGIN
  FQ_REQUIRE(shape && out, "null argument");
  FQ_REQUIRE(grade >= 0 && grade <= dim, "grade out of range");
  const KuhnTables kt(dim);
  const KuhnGrid g(dim, shape);
  const KuhnGrade& kg = kt.grades[size_t(grade)];
  const std::vector<uint64_t> vbase = kuhn_vbase_host(kt, g, grade);
  const int nl = nlocal(dim, grade);
  const uint64_t ncells = g.nboxes * uint64_t(kt.ncelltypes);
  for (uint64_t c = 0; c < ncells; ++c) {
    uint64_t box = c / uint64_t(kt.ncelltypes);
    const int t = int(c % uint64_t(kt.ncelltypes));
    uint64_t oc[8];
    for (int a = 0; a < dim; ++a) {
      oc[a] = box % g.shape[a];
      box /= g.shape[a];
    }
    for (int l = 0; l < nl; ++l) {
      const uint32_t top = kt.ftop[size_t(grade)][size_t(t) * nl + l];
      uint64_t w = 0;
      uint32_t B = 0;
      for (int a = 0; a < dim; ++a) {
        const uint64_t ca = oc[a] + ((top >> a) & 1u);
        w += ca * g.vstride[a];
        if (ca) B |= 1u << a;
      }
      out[c * nl + l] = vbase[size_t(w)] + kg.rank_in[size_t(B) * kg.ntypes + kt.ftype[size_t(grade)][size_t(t) * nl + l]];
    }
  }
  FQ_API_END
}

// ---------------------------------------------------------------- element matrices
int fq_elmat_shape(int dim, int kind, int grade, int* rows, int* cols) {
  FQ_API_BEGIN
  FQ_REQUIRE(rows && cols, "null argument");
  int tg, rg;
  kind_grades(kind, grade, tg, rg);
  *rows = nlocal(dim, tg);
  *cols = nlocal(dim, rg);
  FQ_API_END
}

int fq_elmat_batch(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, size_t cell_begin, size_t cell_end,
                   int use_generated, double* out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && out, "null argument");
  FQ_REQUIRE(kind >= 0 && kind <= 4, "unknown kind");
  FQ_REQUIRE(cell_begin <= cell_end && cell_end <= mesh->ncells, "cell range out of bounds");
  FQ_CUDA(cudaSetDevice(ctx->device));
  const std::vector<BlockSpec> blocks{{kind, grade}};
  const int nouts = elmat_nouts(mesh->dim, blocks);
  const size_t nc = cell_end - cell_begin;
  if (nc == 0 || nouts == 0) return FQ_OK;
  DevBuf<double> slab(nc * size_t(nouts));
  DevBuf<int> err(1);
  FQ_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), ctx->stream));
  double* outs[1] = {slab.p};
  elmat_to_slabs(ctx, mesh, blocks, cell_begin, cell_end, use_generated != 0, outs, err.p);
  int h = 0;
  FQ_CUDA(cudaMemcpyAsync(&h, err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaMemcpyAsync(out, slab.p, slab.bytes(), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h) throw Error(FQ_ERR_DEGENERATE, "a cell metric is singular");
  FQ_API_END
}

// ---------------------------------------------------------------- assembly
int fq_assemble_symbolic(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, size_t row_begin, size_t row_end,
                         fq_csr** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && out, "null argument");
  FQ_REQUIRE(kind >= 0 && kind <= 4, "unknown kind");
  FQ_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<fq_csr> m(new fq_csr);
  assemble_symbolic(ctx, mesh, kind, grade, row_begin, row_end, m.get());
  *out = m.release();
  FQ_API_END
}
int fq_assemble_numeric(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* csr, int drop_exact_zeros) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && csr, "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  assemble_numeric(ctx, mesh, csr, drop_exact_zeros != 0);
  FQ_API_END
}
int fq_assemble(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, int drop_exact_zeros, fq_csr** out) {
  const int rc = fq_assemble_symbolic(ctx, mesh, kind, grade, 0, std::numeric_limits<size_t>::max(), out);
  if (rc != FQ_OK) return rc;
  const int rc2 = fq_assemble_numeric(ctx, mesh, *out, drop_exact_zeros);
  if (rc2 != FQ_OK) {
    delete *out;
    *out = nullptr;
  }
  return rc2;
}

struct fq_hodge {
  int grade = 0;
  std::unique_ptr<fq_csr> blocks[4];
};

int fq_hodge_symbolic(fq_ctx* ctx, const fq_mesh* mesh, int grade, size_t sigma_row_begin, size_t sigma_row_end,
                      size_t u_row_begin, size_t u_row_end, fq_hodge** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && out, "null argument");
  FQ_REQUIRE(grade >= 0 && grade <= mesh->dim, "grade <= complex.dim() is required");  // hodge.rs:63
  FQ_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<fq_hodge> h(new fq_hodge);
  h->grade = grade;
  const auto specs = hodge_blocks(grade);
  for (int b = 0; b < 4; ++b) {
    h->blocks[b].reset(new fq_csr);
    const bool sigma_rows = (b == 0 || b == 2);
    assemble_symbolic(ctx, mesh, specs[size_t(b)].kind, specs[size_t(b)].grade, sigma_rows ? sigma_row_begin : u_row_begin,
                      sigma_rows ? sigma_row_end : u_row_end, h->blocks[b].get());
  }
  *out = h.release();
  FQ_API_END
}
int fq_hodge_numeric(fq_ctx* ctx, const fq_mesh* mesh, fq_hodge* blocks, int drop_exact_zeros) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && blocks, "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  fq_csr* ptrs[4] = {blocks->blocks[0].get(), blocks->blocks[1].get(), blocks->blocks[2].get(), blocks->blocks[3].get()};
  assemble_numeric_multi(ctx, mesh, ptrs, 4, drop_exact_zeros != 0);
  FQ_API_END
}
int fq_csr_transpose(fq_ctx* ctx, const fq_csr* a, fq_csr** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && out, "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<fq_csr> t(new fq_csr);
  csr_transpose(ctx, a, t.get());
  *out = t.release();
  FQ_API_END
}
int fq_csr_row_abs_sums(fq_ctx* ctx, const fq_csr* a, fq_vec* y) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && y, "null argument");
  FQ_REQUIRE(y->d.n == a->row_end - a->row_begin, "row_abs_sums: dimension mismatch");
  FQ_CUDA(cudaSetDevice(ctx->device));
  csr_row_abs_sums(ctx, a, y->d.p);
  FQ_API_END
}
int fq_csr_inv_diagonal(fq_ctx* ctx, fq_csr* a, fq_vec* d) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && d, "null argument");
  const size_t nrows = a->row_end - a->row_begin;
  FQ_REQUIRE(d->d.n == nrows, "inv_diagonal: dimension mismatch");
  FQ_REQUIRE(a->nrows == a->ncols, "inv_diagonal needs a square matrix");
  FQ_CUDA(cudaSetDevice(ctx->device));
  csr_build_inv_diag(ctx, a);
  if (nrows)
    FQ_CUDA(cudaMemcpyAsync(d->d.p, a->inv_diag.p, nrows * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  FQ_API_END
}
int fq_vec_mul(fq_ctx* ctx, fq_vec* z, const fq_vec* d, const fq_vec* r) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && z && d && r, "null argument");
  FQ_REQUIRE(z->d.n == d->d.n && z->d.n == r->d.n, "vec_mul: dimension mismatch");
  FQ_CUDA(cudaSetDevice(ctx->device));
  vec_mul_pointwise(ctx, z->d.p, d->d.p, r->d.p, z->d.n);
  FQ_API_END
}
int fq_csr_add(fq_ctx* ctx, const fq_csr* a, const fq_csr* b, fq_csr** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && b && out, "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<fq_csr> c(new fq_csr);
  csr_add(ctx, a, b, c.get());
  *out = c.release();
  FQ_API_END
}
int fq_csr_restrict(fq_ctx* ctx, const fq_csr* a, const size_t* rows_keep, size_t nrows_keep, const size_t* cols_keep,
                    size_t ncols_keep, fq_csr** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && out && (rows_keep || nrows_keep == 0) && (cols_keep || ncols_keep == 0), "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  for (size_t i = 0; i < nrows_keep; ++i)
    FQ_REQUIRE(rows_keep[i] < a->nrows && (i == 0 || rows_keep[i] > rows_keep[i - 1]), "rows_keep must be ascending and in range");
  for (size_t i = 0; i < ncols_keep; ++i)
    FQ_REQUIRE(cols_keep[i] < a->ncols && (i == 0 || cols_keep[i] > cols_keep[i - 1]), "cols_keep must be ascending and in range");
  DevBuf<uint32_t> dr, dc;
  upload_narrow(ctx, reinterpret_cast<const uint64_t*>(rows_keep), nrows_keep, dr);
  upload_narrow(ctx, reinterpret_cast<const uint64_t*>(cols_keep), ncols_keep, dc);
  std::unique_ptr<fq_csr> r(new fq_csr);
  csr_restrict(ctx, a, dr.p, nrows_keep, dc.p, ncols_keep, r.get());
  *out = r.release();
  FQ_API_END
}
int fq_hodge_mixed_laplacian(fq_ctx* ctx, const fq_hodge* blocks, fq_csr** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && blocks && out, "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  const fq_csr* m_sigma = blocks->blocks[0].get();
  const fq_csr* dif_test = blocks->blocks[2].get();
  const fq_csr* dif_both = blocks->blocks[3].get();
  fq_csr dt_t;
  csr_transpose(ctx, dif_test, &dt_t);
  std::unique_ptr<fq_csr> a(new fq_csr);
  csr_block2x2(ctx, m_sigma, dif_test, -1.0, &dt_t, dif_both, a.get());
  *out = a.release();
  FQ_API_END
}
int fq_hodge_mixed_kkt_symmetric(fq_ctx* ctx, const fq_hodge* blocks, fq_csr** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && blocks && out, "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  fq_csr dt_t;
  csr_transpose(ctx, blocks->blocks[2].get(), &dt_t);
  std::unique_ptr<fq_csr> a(new fq_csr);
  csr_block2x2(ctx, blocks->blocks[0].get(), blocks->blocks[2].get(), 1.0, &dt_t, blocks->blocks[3].get(), a.get(), -1.0);
  *out = a.release();
  FQ_API_END
}
fq_csr* fq_hodge_block(fq_hodge* blocks, int which) {
  return (blocks && which >= 0 && which < 4) ? blocks->blocks[which].get() : nullptr;
}
int fq_hodge_destroy(fq_hodge* blocks) {
  delete blocks;
  return FQ_OK;
}

// ---------------------------------------------------------------- CSR
int fq_csr_shape(const fq_csr* csr, size_t* nrows, size_t* ncols, size_t* nnz) {
  FQ_API_BEGIN
  FQ_REQUIRE(csr, "null argument");
  if (nrows) *nrows = csr->nrows;
  if (ncols) *ncols = csr->ncols;
  if (nnz) *nnz = csr->nnz;
  FQ_API_END
}
int fq_csr_row_range(const fq_csr* csr, size_t* row_begin, size_t* row_end) {
  FQ_API_BEGIN
  FQ_REQUIRE(csr, "null argument");
  if (row_begin) *row_begin = csr->row_begin;
  if (row_end) *row_end = csr->row_end;
  FQ_API_END
}
int fq_csr_download(fq_ctx* ctx, const fq_csr* csr, size_t* row_offsets, size_t* col_indices, double* values) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && csr, "null argument");
  static_assert(sizeof(size_t) == sizeof(uint64_t), "usize must be 64-bit");
  const size_t nrows_local = csr->row_end - csr->row_begin;
  if (row_offsets) download_widen(ctx, csr->row_ptr.p, nrows_local + 1, reinterpret_cast<uint64_t*>(row_offsets));
  if (col_indices) download_widen(ctx, csr->col_idx.p, csr->nnz, reinterpret_cast<uint64_t*>(col_indices));
  if (values && csr->nnz) {
    FQ_CUDA(cudaMemcpyAsync(values, csr->values.p, csr->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  FQ_API_END
}
// device u32 -> host u64 on the copy stream.  With host widening the u32 array is copied into the upper half of the
// caller's buffer and a host callback hands it to the widening pool (half the PCIe bytes, no device staging);
// otherwise the whole array is widened on the device into one staging buffer kept until the wait.
struct WidenTicket {
  fq::HostWidener* pool;
  uint64_t* buf;
  size_t n;
};
static void CUDART_CB widen_callback(void* p) {
  WidenTicket* t = static_cast<WidenTicket*>(p);
  t->pool->enqueue(t->buf, t->n);
  delete t;
}
static void download_widen_async(fq_ctx* ctx, const uint32_t* dev, size_t n, uint64_t* host) {
  if (!n) return;
  if (!ctx->widener_probed) {
    ctx->widener_probed = true;
    const int threads = fq::host_widen_threads();
    if (threads > 0) ctx->widener = new fq::HostWidener(threads);
  }
  if (ctx->widener) {
    FQ_CUDA(cudaMemcpyAsync(reinterpret_cast<uint32_t*>(host) + n, dev, n * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                            ctx->copy_stream));
    WidenTicket* ticket = new WidenTicket{ctx->widener, host, n};
    const cudaError_t err = cudaLaunchHostFunc(ctx->copy_stream, widen_callback, ticket);
    if (err != cudaSuccess) {
      delete ticket;
      FQ_CUDA(err);
    }
    return;
  }
  ctx->pending_staging.emplace_back(n);
  uint64_t* stage = ctx->pending_staging.back().p;
  widen_u32_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->copy_stream>>>(dev, n, stage);
  fq_count_launch(ctx);
  FQ_CUDA(cudaMemcpyAsync(host, stage, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
}
int fq_csr_download_async(fq_ctx* ctx, const fq_csr* csr, size_t* row_offsets, size_t* col_indices, double* values) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && csr, "null argument");
  if (!ctx->copy_stream) FQ_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  // everything enqueued so far on the compute stream (the assembly of this matrix) precedes the copies
  cudaEvent_t ready;
  FQ_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  FQ_CUDA(cudaEventRecord(ready, ctx->stream));
  FQ_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ready, 0));
  FQ_CUDA(cudaEventDestroy(ready));
  const size_t nrows_local = csr->row_end - csr->row_begin;
  if (row_offsets) download_widen_async(ctx, csr->row_ptr.p, nrows_local + 1, reinterpret_cast<uint64_t*>(row_offsets));
  if (col_indices) download_widen_async(ctx, csr->col_idx.p, csr->nnz, reinterpret_cast<uint64_t*>(col_indices));
  if (values && csr->nnz)
    FQ_CUDA(cudaMemcpyAsync(values, csr->values.p, csr->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
  FQ_API_END
}
int fq_ctx_wait_downloads(fq_ctx* ctx) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx, "null argument");
  if (ctx->copy_stream) FQ_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  if (ctx->widener) ctx->widener->wait_idle();
  ctx->pending_staging.clear();
  FQ_API_END
}
int fq_csr_upload(fq_ctx* ctx, size_t nrows, size_t ncols, const size_t* row_offsets, const size_t* col_indices,
                  const double* values, fq_csr** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && row_offsets && out, "null argument");
  FQ_REQUIRE(nrows < (size_t(1) << 32) && ncols < (size_t(1) << 32), "more than 2^32 rows/cols: not supported");
  const size_t nnz = row_offsets[nrows];
  FQ_REQUIRE(nnz < (size_t(1) << 32), "more than 2^32 non-zeros: not supported");
  FQ_REQUIRE(nnz == 0 || (col_indices && values), "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<fq_csr> m(new fq_csr);
  m->nrows = nrows;
  m->ncols = ncols;
  m->row_begin = 0;
  m->row_end = nrows;
  m->nnz = nnz;
  upload_narrow(ctx, reinterpret_cast<const uint64_t*>(row_offsets), nrows + 1, m->row_ptr);
  upload_narrow(ctx, reinterpret_cast<const uint64_t*>(col_indices), nnz, m->col_idx);
  m->values.alloc(nnz ? nnz : 1);
  if (nnz) FQ_CUDA(cudaMemcpyAsync(m->values.p, values, nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = m.release();
  FQ_API_END
}
int fq_csr_destroy(fq_csr* csr) {
  delete csr;
  return FQ_OK;
}
int64_t fq_csr_assembly_bytes(const fq_csr* csr) { return csr ? csr->assembly_bytes : 0; }
int64_t fq_csr_assembly_shared_bytes(const fq_csr* csr) { return csr ? csr->assembly_shared_bytes : 0; }
double fq_csr_plan_build_ms(const fq_csr* csr) { return csr ? csr->plan_build_ms : 0.0; }
size_t fq_csr_plan_cell_visits(const fq_csr* csr) { return csr ? csr->plan_cell_visits : 0; }
int64_t fq_csr_spmv_bytes(const fq_csr* csr) {
  if (!csr) return 0;
  const size_t nrows_local = csr->row_end - csr->row_begin;
  return int64_t(12 * csr->nnz + 4 * (nrows_local + 1) + 8 * nrows_local + 8 * csr->ncols);
}

// ---------------------------------------------------------------- vectors
int fq_vec_create(fq_ctx* ctx, size_t n, fq_vec** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && out, "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<fq_vec> v(new fq_vec);
  v->d.alloc(n ? n : 1);
  v->d.n = n;
  FQ_CUDA(cudaMemsetAsync(v->d.p, 0, (n ? n : 1) * sizeof(double), ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = v.release();
  FQ_API_END
}
int fq_vec_wrap(fq_ctx* ctx, void* device_ptr, size_t n, fq_vec** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && out && (device_ptr || n == 0), "null argument");
  std::unique_ptr<fq_vec> v(new fq_vec);
  v->d.p = static_cast<double*>(device_ptr);
  v->d.n = n;
  v->d.owned = false;
  *out = v.release();
  FQ_API_END
}
int fq_vec_destroy(fq_vec* v) {
  if (v && v->ipc_base) cudaIpcCloseMemHandle(v->ipc_base);
  delete v;
  return FQ_OK;
}
int fq_vec_ipc_export(fq_ctx* ctx, const fq_vec* v, unsigned char* handle64) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && v && handle64 && v->d.p, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  FQ_CUDA(cudaIpcGetMemHandle(&h, v->d.p));
  std::memcpy(handle64, &h, 64);
  FQ_API_END
}
int fq_vec_ipc_import(fq_ctx* ctx, const unsigned char* handle64, size_t n, fq_vec** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && handle64 && out, "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* p = nullptr;
  FQ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  std::unique_ptr<fq_vec> v(new fq_vec);
  v->d.p = static_cast<double*>(p);
  v->d.n = n;
  v->d.owned = false;
  v->ipc_base = p;
  *out = v.release();
  FQ_API_END
}
int fq_flag_signal(fq_ctx* ctx, fq_vec* flag, double value) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && flag && flag->d.n >= 1, "null argument");
  flag_signal(ctx, flag->d.p, value);
  FQ_API_END
}
int fq_flag_wait(fq_ctx* ctx, const fq_vec* flag, double value) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && flag && flag->d.n >= 1, "null argument");
  if (ctx->d_flag_timeout.n != 1) {
    ctx->d_flag_timeout.alloc(1);
    FQ_CUDA(cudaMemsetAsync(ctx->d_flag_timeout.p, 0, sizeof(int), ctx->stream));
  }
  flag_wait(ctx, flag->d.p, value, ctx->d_flag_timeout.p);
  FQ_API_END
}
int fq_flag_check(fq_ctx* ctx) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx, "null argument");
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->d_flag_timeout.n == 1) {
    int h = 0;
    FQ_CUDA(cudaMemcpy(&h, ctx->d_flag_timeout.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (h) throw Error(FQ_ERR_CUDA, "a peer flag wait timed out");
  }
  FQ_API_END
}
size_t fq_vec_len(const fq_vec* v) { return v ? v->d.n : 0; }
int fq_vec_upload(fq_ctx* ctx, fq_vec* v, const double* host) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && v && (host || v->d.n == 0), "null argument");
  if (v->d.n) FQ_CUDA(cudaMemcpyAsync(v->d.p, host, v->d.n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  FQ_API_END
}
int fq_vec_download(fq_ctx* ctx, const fq_vec* v, double* host) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && v && (host || v->d.n == 0), "null argument");
  if (v->d.n) FQ_CUDA(cudaMemcpyAsync(host, v->d.p, v->d.n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  FQ_API_END
}
int fq_vec_copy(fq_ctx* ctx, fq_vec* dst, const fq_vec* src) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && dst && src && dst->d.n == src->d.n, "vector length mismatch");
  if (src->d.n)
    FQ_CUDA(cudaMemcpyAsync(dst->d.p, src->d.p, src->d.n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  FQ_API_END
}
int fq_vec_dot(fq_ctx* ctx, const fq_vec* x, const fq_vec* y, double* out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && x && y && out && x->d.n == y->d.n, "vector length mismatch");
  *out = vec_dot(ctx, x->d.p, y->d.p, x->d.n);
  FQ_API_END
}
int fq_vec_scale(fq_ctx* ctx, fq_vec* x, double alpha) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && x, "null argument");
  vec_scale(ctx, x->d.p, alpha, x->d.n);
  FQ_API_END
}
int fq_vec_axpy(fq_ctx* ctx, fq_vec* y, double alpha, const fq_vec* x) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && x && y && x->d.n == y->d.n && x != y, "axpy needs two distinct vectors of one length");
  vec_axpy(ctx, y->d.p, alpha, x->d.p, x->d.n);
  FQ_API_END
}
void* fq_vec_device_ptr(fq_vec* v) { return v ? v->d.p : nullptr; }

// ---------------------------------------------------------------- SpMV / Krylov
int fq_spmv(fq_ctx* ctx, const fq_csr* a, const fq_vec* x, fq_vec* y) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && x && y, "null argument");
  FQ_REQUIRE(x->d.n == a->ncols && y->d.n == a->row_end - a->row_begin, "spmv: dimension mismatch");
  FQ_REQUIRE(x != y, "spmv: x and y must be distinct");
  spmv_prepare(ctx, const_cast<fq_csr*>(a));
  spmv_apply(ctx, a, x->d.p, y->d.p);
  FQ_API_END
}
int fq_spmv_window(fq_ctx* ctx, const fq_csr* a, const fq_vec* x, size_t x_lo, fq_vec* y) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && x && y, "null argument");
  FQ_REQUIRE(x_lo + x->d.n <= a->ncols && y->d.n == a->row_end - a->row_begin, "spmv_window: dimension mismatch");
  FQ_REQUIRE(x != y, "spmv: x and y must be distinct");
  spmv_prepare(ctx, const_cast<fq_csr*>(a));
  // column ids are global: shift the base so that x[col] addresses the window
  spmv_apply(ctx, a, x->d.p - x_lo, y->d.p);
  FQ_API_END
}

int fq_spmv_peer(fq_ctx* ctx, const fq_csr* a, const fq_vec* x_window, size_t held_lo, size_t own_lo, size_t own_hi,
                 const fq_vec* x_lower, size_t lower_held_lo, const fq_vec* x_upper, size_t upper_held_lo, fq_vec* y) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && x_window && y, "null argument");
  FQ_REQUIRE(held_lo <= own_lo && own_lo <= own_hi && own_hi <= a->ncols, "spmv_peer: bad ranges");
  FQ_REQUIRE(y->d.n == a->row_end - a->row_begin, "spmv_peer: dimension mismatch");
  FQ_REQUIRE(x_window != y, "spmv: x and y must be distinct");
  spmv_prepare(ctx, const_cast<fq_csr*>(a));
  // pointers pre-offset so that ptr[global column] addresses the right window
  FQ_REQUIRE(held_lo + x_window->d.n >= own_hi, "spmv_peer: window shorter than the owned range");
  double* own = const_cast<double*>(x_window->d.p) - held_lo;  // the halo slots of the window are filled by the kernel
  const double* lower = x_lower ? x_lower->d.p - lower_held_lo : nullptr;
  const double* upper = x_upper ? x_upper->d.p - upper_held_lo : nullptr;
  spmv_apply_peer(ctx, const_cast<fq_csr*>(a), own, lower, upper, held_lo, own_lo, own_hi, held_lo + x_window->d.n, y->d.p);
  FQ_API_END
}

int fq_spmv_peer_epoch(fq_ctx* ctx, const fq_csr* a, const fq_vec* x_window, size_t held_lo, size_t own_lo, size_t own_hi,
                       const fq_vec* x_lower, size_t lower_held_lo, const fq_vec* ready_lower, const fq_vec* x_upper,
                       size_t upper_held_lo, const fq_vec* ready_upper, double epoch, fq_vec* consumed, fq_vec* y) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && x_window && y, "null argument");
  FQ_REQUIRE(held_lo <= own_lo && own_lo <= own_hi && own_hi <= a->ncols, "spmv_peer: bad ranges");
  FQ_REQUIRE(y->d.n == a->row_end - a->row_begin, "spmv_peer: dimension mismatch");
  FQ_REQUIRE(x_window != y, "spmv: x and y must be distinct");
  FQ_REQUIRE(held_lo + x_window->d.n >= own_hi, "spmv_peer: window shorter than the owned range");
  FQ_REQUIRE((!ready_lower || x_lower) && (!ready_upper || x_upper), "spmv_peer: a ready flag without its window");
  spmv_prepare(ctx, const_cast<fq_csr*>(a));
  double* own = const_cast<double*>(x_window->d.p) - held_lo;
  const double* lower = x_lower ? x_lower->d.p - lower_held_lo : nullptr;
  const double* upper = x_upper ? x_upper->d.p - upper_held_lo : nullptr;
  PeerSync sync;
  sync.ready_lower = ready_lower ? ready_lower->d.p : nullptr;
  sync.ready_upper = ready_upper ? ready_upper->d.p : nullptr;
  sync.epoch = epoch;
  sync.consumed = consumed ? consumed->d.p : nullptr;
  if (ctx->d_flag_timeout.n != 1) {
    ctx->d_flag_timeout.alloc(1);
    FQ_CUDA(cudaMemsetAsync(ctx->d_flag_timeout.p, 0, sizeof(int), ctx->stream));
  }
  sync.timeout = ctx->d_flag_timeout.p;
  spmv_apply_peer(ctx, const_cast<fq_csr*>(a), own, lower, upper, held_lo, own_lo, own_hi, held_lo + x_window->d.n, y->d.p,
                  sync);
  FQ_API_END
}

// ---------------------------------------------------------------- matrix-free ElementOperator (matfree.rs)
int fq_matfree_create(fq_ctx* ctx, const fq_mesh* mesh, int kind, int grade, fq_matfree** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && out, "null argument");
  FQ_REQUIRE(kind >= 0 && kind <= 4, "unknown kind");
  FQ_CUDA(cudaSetDevice(ctx->device));
  fq_matfree* op = matfree_new();
  try {
    matfree_build(ctx, mesh, kind, grade, op);
    matfree_refresh(ctx, op);
  } catch (...) {
    matfree_delete(op);
    throw;
  }
  *out = op;
  FQ_API_END
}
int fq_linear_form_create(fq_ctx* ctx, const fq_mesh* mesh, int grade, fq_matfree** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && out, "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  fq_matfree* op = matfree_new();
  try {
    vector_plan_build(ctx, mesh, grade, op);
  } catch (...) {
    matfree_delete(op);
    throw;
  }
  *out = op;
  FQ_API_END
}
int fq_linear_form_assemble(fq_ctx* ctx, const fq_matfree* plan, const double* element_vectors, fq_vec* out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && plan && out, "null argument");
  FQ_REQUIRE(out->d.n == matfree_nrows(plan), "the load vector has one entry per simplex of the form's grade");  // galerkin.rs:285-286
  FQ_REQUIRE(element_vectors || matfree_nrows(plan) == 0, "null element vectors");
  FQ_CUDA(cudaSetDevice(ctx->device));
  vector_plan_assemble(ctx, plan, element_vectors, out->d.p);
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));  // the host buffer may be reused on return
  FQ_API_END
}
int fq_source_form_assemble(fq_ctx* ctx, const fq_matfree* plan, int nnodes, const double* weights, const double* shapes,
                            const double* samples, fq_vec* out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && plan && out && weights && shapes, "null argument");
  FQ_REQUIRE(out->d.n == matfree_nrows(plan), "the load vector has one entry per simplex of the form's grade");
  FQ_REQUIRE(samples || matfree_nrows(plan) == 0, "null source samples");
  FQ_CUDA(cudaSetDevice(ctx->device));
  vector_plan_source(ctx, plan, nnodes, weights, shapes, samples, out->d.p);
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));  // the host buffers may be reused on return
  FQ_API_END
}
int fq_weighted_mass_numeric(fq_ctx* ctx, const fq_mesh* mesh, fq_csr* csr, int nnodes, const double* weights,
                             const double* shapes, const double* coefficient, int drop_exact_zeros) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && mesh && csr && weights && shapes, "null argument");
  FQ_REQUIRE(csr->kind == KIND_MASS, "the weighted mass is assembled on the pattern of the mass of its grade");
  FQ_REQUIRE(coefficient || mesh->ncells == 0, "null coefficient samples");
  FQ_CUDA(cudaSetDevice(ctx->device));
  const int grade = csr->grade;
  assemble_numeric_custom(ctx, mesh, csr, drop_exact_zeros != 0, [&](double* slab) {
    weighted_mass_to_slab(ctx, mesh, grade, nnodes, weights, shapes, coefficient, slab);
  });
  FQ_API_END
}
int fq_linear_form_destroy(fq_matfree* plan) { return fq_matfree_destroy(plan); }

int fq_matfree_refresh(fq_ctx* ctx, fq_matfree* op) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && op, "null argument");
  matfree_refresh(ctx, op);
  FQ_API_END
}
int fq_matfree_destroy(fq_matfree* op) {
  if (op) matfree_delete(op);
  return FQ_OK;
}
int fq_matfree_shape(const fq_matfree* op, size_t* nrows, size_t* ncols) {
  FQ_API_BEGIN
  FQ_REQUIRE(op, "null argument");
  if (nrows) *nrows = matfree_nrows(op);
  if (ncols) *ncols = matfree_ncols(op);
  FQ_API_END
}
int fq_matfree_apply(fq_ctx* ctx, const fq_matfree* op, const fq_vec* x, fq_vec* y) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && op && x && y, "null argument");
  FQ_REQUIRE(x->d.n == matfree_ncols(op) && y->d.n == matfree_nrows(op), "operator and vector disagree");  // matfree.rs:101
  FQ_REQUIRE(x != y, "apply: x and y must be distinct");
  matfree_apply(ctx, op, x->d.p, y->d.p);
  FQ_API_END
}
int fq_matfree_diagonal(fq_ctx* ctx, const fq_matfree* op, fq_vec* d) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && op && d, "null argument");
  FQ_REQUIRE(d->d.n == matfree_nrows(op), "diagonal: dimension mismatch");
  matfree_diagonal(ctx, op, d->d.p);
  FQ_API_END
}

static int krylov_common(bool is_cg, fq_ctx* ctx, const fq_csr* a, int precond, const fq_vec* b, double rtol,
                         size_t max_iters, fq_vec* x, size_t* iters, double* residual, int* converged) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && b && x, "null argument");
  FQ_REQUIRE(precond == 0 || precond == 1, "precond: 0 identity, 1 Jacobi");
  fq_csr* am = const_cast<fq_csr*>(a);
  const KrylovReport r = is_cg ? krylov_cg(ctx, am, precond, b, rtol, max_iters, x)
                               : krylov_minres(ctx, am, precond, b, rtol, max_iters, x);
  if (iters) *iters = r.iters;
  if (residual) *residual = r.residual;
  if (converged) *converged = r.converged ? 1 : 0;
  FQ_API_END
}
int fq_cg(fq_ctx* ctx, const fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters, fq_vec* x,
          size_t* iters, double* residual, int* converged) {
  return krylov_common(true, ctx, a, precond, b, rtol, max_iters, x, iters, residual, converged);
}
int fq_minres(fq_ctx* ctx, const fq_csr* a, int precond, const fq_vec* b, double rtol, size_t max_iters, fq_vec* x,
              size_t* iters, double* residual, int* converged) {
  return krylov_common(false, ctx, a, precond, b, rtol, max_iters, x, iters, residual, converged);
}

}  // extern "C"

// ---- Krylov over user operators (iterative::{cg, minres} are generic over LinearOperator / ApproxInverse /
// InnerProductSpace, iterative/src/krylov.rs:48,113): the operator, the preconditioner and the completion of an inner
// product (the all-reduce of a distributed space) are callbacks on device pointers
static int krylov_op_common(bool is_cg, fq_ctx* ctx, size_t n, fq_apply_fn apply, fq_apply_fn precond, fq_reduce_fn reduce,
                            void* user, const fq_vec* b, double rtol, size_t max_iters, fq_vec* x, size_t* iters,
                            double* residual, int* converged) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && apply && b && x, "null argument");
  FQ_REQUIRE(b->d.n == n && x->d.n == n, "dimension mismatch");
  FQ_CUDA(cudaSetDevice(ctx->device));
  KrylovOps ops;
  ops.apply = [=](const double* xin, double* y) {
    if (apply(user, xin, y) != 0) throw Error(FQ_ERR_INVALID, "operator callback failed");
  };
  ops.precond = [=](const double* r, double* z) {
    if (!precond) {
      if (n) FQ_CUDA(cudaMemcpyAsync(z, r, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    } else if (precond(user, r, z) != 0) {
      throw Error(FQ_ERR_INVALID, "preconditioner callback failed");
    }
  };
  ops.dot = [=](const double* u, const double* v) {
    double local = vec_dot(ctx, u, v, n), global = local;
    if (reduce && reduce(user, local, &global) != 0) throw Error(FQ_ERR_INVALID, "reduction callback failed");
    return global;
  };
  const KrylovReport rep = is_cg ? cg_core(ctx, n, ops, b->d.p, rtol, max_iters, x->d.p)
                                 : minres_core(ctx, n, ops, b->d.p, rtol, max_iters, x->d.p);
  if (iters) *iters = rep.iters;
  if (residual) *residual = rep.residual;
  if (converged) *converged = rep.converged ? 1 : 0;
  FQ_API_END
}
int fq_cg_op(fq_ctx* ctx, size_t n, fq_apply_fn apply, fq_apply_fn precond, fq_reduce_fn reduce, void* user, const fq_vec* b,
             double rtol, size_t max_iters, fq_vec* x, size_t* iters, double* residual, int* converged) {
  return krylov_op_common(true, ctx, n, apply, precond, reduce, user, b, rtol, max_iters, x, iters, residual, converged);
}
int fq_minres_op(fq_ctx* ctx, size_t n, fq_apply_fn apply, fq_apply_fn precond, fq_reduce_fn reduce, void* user,
                 const fq_vec* b, double rtol, size_t max_iters, fq_vec* x, size_t* iters, double* residual, int* converged) {
  return krylov_op_common(false, ctx, n, apply, precond, reduce, user, b, rtol, max_iters, x, iters, residual, converged);
}
int fq_minres_blockdiag(fq_ctx* ctx, const fq_csr* a, int nblocks, const fq_csr* const* blocks, const size_t* offsets,
                        double inner_rtol, size_t inner_max_iters, const fq_vec* b, double rtol, size_t max_iters, fq_vec* x,
                        size_t* iters, double* residual, int* converged, size_t* inner_iters) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && a && blocks && offsets && b && x, "null argument");
  FQ_CUDA(cudaSetDevice(ctx->device));
  std::vector<fq_csr*> blk(static_cast<size_t>(nblocks > 0 ? nblocks : 0));
  for (int i = 0; i < nblocks; ++i) blk[size_t(i)] = const_cast<fq_csr*>(blocks[i]);
  const KrylovReport rep = krylov_minres_blockdiag(ctx, const_cast<fq_csr*>(a), nblocks, blk.data(), offsets, inner_rtol,
                                                   inner_max_iters, b, rtol, max_iters, x, inner_iters);
  if (iters) *iters = rep.iters;
  if (residual) *residual = rep.residual;
  if (converged) *converged = rep.converged ? 1 : 0;
  FQ_API_END
}
// non-owning view of a segment of a vector
int fq_vec_view(fq_ctx* ctx, const fq_vec* v, size_t offset, size_t n, fq_vec** out) {
  FQ_API_BEGIN
  FQ_REQUIRE(ctx && v && out, "null argument");
  FQ_REQUIRE(offset + n <= v->d.n, "view out of range");
  std::unique_ptr<fq_vec> w(new fq_vec);
  w->d.p = v->d.p + offset;
  w->d.n = n;
  w->d.owned = false;
  *out = w.release();
  FQ_API_END
}
