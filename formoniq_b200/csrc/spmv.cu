// spmv.cu — K4: CSR SpMV y = A x  (iterative/src/operator.rs:5-14).
//
// "CSR-stream" (stream.cuh): a CTA owns a block of consecutive rows whose
// non-zeros fit a shared-memory chunk.  Phase 1 streams values / column indices
// fully coalesced with 8 independent gathers of x in flight per thread and
// stages the products v*x[col] in shared memory; phase 2 gives every row to one
// thread, which adds the row's products left to right.  That is the
// reference's serial order (one accumulator per row, ascending columns, mul
// then add, no FMA), so y is bit-identical to the CPU path while the HBM
// traffic stays the algorithmic 12 B/nnz + vectors.
#include "internal.hpp"
#include "stream.cuh"

namespace fq {

struct SpmvPolicy {
  static constexpr bool kHasValues = true;
  static constexpr bool kCustomSrc = false;
  static constexpr bool kGated = false;
  double* __restrict__ y;
  __device__ __forceinline__ double load(uint32_t, bool) const { return 0.0; }
  __device__ __forceinline__ void store(uint32_t row, double sum, bool) const { y[row] = sum; }
};

// SpMV fused with the halo exchange: x lives distributed over the ranks' windows; columns below / above this rank's
// owned range are loaded straight from the neighbouring GPU's memory (CUDA IPC mapping, NVLink 5 / NVSwitch P2P
// loads issued by the gather itself), so the transfer overlaps the row blocks that do not need it — there is no
// separate exchange step and no staging copy.  Pointers are pre-offset so that ptr[global column] is valid.
struct SpmvPeerPolicy {
  static constexpr bool kHasValues = true;
  static constexpr bool kCustomSrc = true;
  static constexpr bool kGated = false;
  double* __restrict__ y;
  const double* own;    // this rank's window
  const double* lower;  // rank - 1's window (columns < own_lo)
  const double* upper;  // rank + 1's window (columns >= own_hi)
  uint32_t own_lo, own_hi;
  __device__ __forceinline__ double load(uint32_t col, bool late) const {
    const double* p = col < own_lo ? lower : (col >= own_hi ? upper : own);
    return p[col];  // plain ld.global: peer memory is not read through the non-coherent path
  }
  __device__ __forceinline__ void store(uint32_t row, double sum, bool) const { y[row] = sum; }
};

// Second form of the fused kernel (default): the first two CTAs per SM first pull their share (a few entries per thread, all
// in flight together: a single NVLink round trip) of the two halo segments out of the neighbours' windows with coalesced P2P loads into this
// rank's own window and then works on its row blocks, most of which reference owned columns only; the few blocks at either end of the row range that touch halo
// columns wait on a device-scope counter for the copy (the first ones belong to the copying CTAs themselves, the last
// ones come up when the copy is long done).  The gather then always reads local
// HBM, so remote latency (~2-3 us per dependent load) is paid once, in bulk, and hidden behind the interior rows.
struct SpmvHaloPolicy {
  static constexpr bool kHasValues = true;
  static constexpr bool kCustomSrc = true;
  static constexpr bool kGated = true;
  double* __restrict__ y;
  double* own;          // this rank's window, pre-offset: own[global column]; the halo slots are filled here
  const double* lower;  // rank - 1's window
  const double* upper;  // rank + 1's window
  uint32_t held_lo, own_lo, own_hi, held_hi;
  const uint32_t* split;        // {first, one-past-last} row block free of halo columns
  unsigned long long* counter;  // copy CTAs that have finished, accumulated over launches
  unsigned long long target;    // value of *counter once this launch's copy is complete
  PeerSync sync;                // optional epoch flags handled by the kernel itself (null pointers: none)
  uint32_t ncopy;               // CTAs taking part in the copy: the first ones dispatched, and they never wait before
                                // their share is done, so the gate cannot deadlock whatever the residency
  __device__ __forceinline__ void prologue() const {
    if (blockIdx.x >= ncopy) return;
    if (sync.ready_lower || sync.ready_upper) {
      // the neighbours publish an epoch after their last write of x: wait for it before the first P2P load
      if (threadIdx.x == 0) {
        long long spins = 0;
        const volatile double* f0 = sync.ready_lower;
        const volatile double* f1 = sync.ready_upper;
        while ((f0 && *f0 < sync.epoch) || (f1 && *f1 < sync.epoch)) {
          __nanosleep(100);
          if (++spins > (1ll << 23)) {  // ~3 s: never hang the device on a lost peer
            *sync.timeout = 1;
            break;
          }
        }
        __threadfence_system();
      }
      __syncthreads();
    }
    const uint32_t nl = own_lo - held_lo, total = nl + (held_hi - own_hi);
    const uint32_t stride = ncopy * kStreamThreads;
    for (uint32_t base = blockIdx.x * kStreamThreads + threadIdx.x; base < total; base += 4 * stride) {
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t k = base + u * stride;
        if (k < total) v[u] = __ldcg((k < nl ? lower + held_lo : upper + own_hi - nl) + k);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t k = base + u * stride;
        if (k < total) own[k < nl ? held_lo + k : own_hi + (k - nl)] = v[u];
      }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long done = atomicAdd(counter, 1ull) + 1;
      if (sync.consumed && done == target) {
        // last share copied: the neighbours may overwrite their x (they wait for this epoch before doing so)
        __threadfence_system();
        *reinterpret_cast<volatile double*>(sync.consumed) = sync.epoch;
      }
    }
  }
  __device__ __forceinline__ bool is_late(uint32_t b) const { return b < __ldg(split) || b >= __ldg(split + 1); }
  __device__ __forceinline__ void gate_wait() const {
    if (threadIdx.x == 0) {
      unsigned long long seen;
      do {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(counter) : "memory");
        if (seen < target) __nanosleep(64);
      } while (seen < target);
    }
    __syncthreads();
  }
  __device__ __forceinline__ double load(uint32_t col, bool) const {
    // coherent (weak) loads: the halo slots are written during this launch, and the acquire in gate_wait followed
    // by the CTA barrier orders every thread's later loads after the copy (PTX memory model; ld.global.nc would not be)
    return own[col];
  }
  __device__ __forceinline__ void store(uint32_t row, double sum, bool) const { y[row] = sum; }
};

// ext[0] = 1 + last row with a column below own_lo (0: none), ext[1] = first row with a column >= own_hi (nrows: none);
// columns are ascending inside a row, so the first / last entry decide.
__global__ void peer_rows_kernel(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col_idx, uint32_t nrows,
                                 uint32_t own_lo, uint32_t own_hi, uint32_t* __restrict__ ext) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const uint32_t b = row_ptr[r], e = row_ptr[r + 1];
    if (b == e) continue;
    if (col_idx[b] < own_lo) atomicMax(ext, r + 1);
    if (col_idx[e - 1] >= own_hi) atomicMin(ext + 1, r);
  }
}
__global__ void peer_split_kernel(const uint32_t* __restrict__ blocks, uint32_t nblocks, uint32_t nrows,
                                  const uint32_t* __restrict__ ext, uint32_t* __restrict__ split) {
  // first block starting at or after row ext[0]
  uint32_t lo = 0, hi = nblocks;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (blocks[mid] < ext[0]) lo = mid + 1; else hi = mid;
  }
  split[0] = lo;
  if (ext[1] >= nrows) {
    split[1] = nblocks;
    return;
  }
  // block containing row ext[1]: last b with blocks[b] <= ext[1]
  lo = 0, hi = nblocks;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (blocks[mid] <= ext[1]) lo = mid + 1; else hi = mid;
  }
  split[1] = lo ? lo - 1 : 0;
}

static void peer_prepare(fq_ctx* ctx, fq_csr* a, size_t own_lo, size_t own_hi) {
  if (a->peer_split.n && a->peer_own_lo == own_lo && a->peer_own_hi == own_hi) return;
  const size_t nrows = a->row_end - a->row_begin;
  a->peer_split.alloc(4);
  const uint32_t init[4] = {0u, 0u, 0u, uint32_t(nrows)};
  FQ_CUDA(cudaMemcpyAsync(a->peer_split.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));  // `init` is on the stack
  peer_rows_kernel<<<grid_for(nrows, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
      a->row_ptr.p, a->col_idx.p, uint32_t(nrows), uint32_t(own_lo), uint32_t(own_hi), a->peer_split.p + 2);
  peer_split_kernel<<<1, 1, 0, ctx->stream>>>(a->rowblocks.p, uint32_t(a->nrowblocks), uint32_t(nrows), a->peer_split.p + 2,
                                              a->peer_split.p);
  fq_count_launch(ctx, 2);
  FQ_CUDA(cudaGetLastError());
  if (!a->peer_counter.n) {
    a->peer_counter.alloc(1);
    FQ_CUDA(cudaMemsetAsync(a->peer_counter.p, 0, sizeof(unsigned long long), ctx->stream));
    a->peer_launches = 0;
  }
  a->peer_own_lo = own_lo, a->peer_own_hi = own_hi;
}

void spmv_apply_peer(fq_ctx* ctx, fq_csr* a, double* own, const double* lower, const double* upper, size_t held_lo,
                     size_t own_lo, size_t own_hi, size_t held_hi, double* y, const PeerSync& sync) {
  FQ_REQUIRE(a->spmv_ready, "spmv_prepare was not called");
  if (a->nrowblocks == 0) return;
  static const bool direct = [] {
    const char* e = getenv("FQ_PEER_DIRECT");
    return e && *e && *e != '0';
  }();
  if (direct) {
    FQ_REQUIRE(!sync.ready_lower && !sync.ready_upper && !sync.consumed, "FQ_PEER_DIRECT: in-kernel epoch flags are not supported");
    ScopedSpan span(ctx, "k4_spmv_peer_direct");
    stream_reduce(ctx, a->rowblocks.p, a->nrowblocks, a->row_ptr.p, a->col_idx.p, a->values.p, nullptr,
                  SpmvPeerPolicy{y, own, lower ? lower : own, upper ? upper : own, uint32_t(lower ? own_lo : 0), uint32_t(upper ? own_hi : a->ncols)});
    return;
  }
  // no neighbour on a side: that halo segment is empty
  if (!lower) held_lo = own_lo;
  if (!upper) held_hi = own_hi;
  peer_prepare(ctx, a, own_lo, own_hi);
  ScopedSpan span(ctx, "k4_spmv_peer");
  const size_t cap = size_t(ctx->sm_count) * 8;
  const size_t grid = a->nrowblocks < cap ? a->nrowblocks : cap;
  const size_t two_per_sm = size_t(ctx->sm_count) * 2;
  const size_t ncopy = grid < two_per_sm ? grid : two_per_sm;
  a->peer_launches += 1;
  stream_reduce(ctx, a->rowblocks.p, a->nrowblocks, a->row_ptr.p, a->col_idx.p, a->values.p, nullptr,
                SpmvHaloPolicy{y, own, lower ? lower : own, upper ? upper : own, uint32_t(held_lo), uint32_t(own_lo),
                               uint32_t(own_hi), uint32_t(held_hi), a->peer_split.p, a->peer_counter.p,
                               (unsigned long long)(a->peer_launches * ncopy), sync, uint32_t(ncopy)});
}

// Stream-ordered flags in (peer-mapped) device memory: the producer of x publishes an epoch after its last write,
// the consumers wait for it before the fused SpMV reads x over NVLink.
__global__ void flag_signal_kernel(volatile double* flag, double value) {
  __threadfence_system();
  *flag = value;
  __threadfence_system();
}
__global__ void flag_wait_kernel(const volatile double* flag, double value, int* timeout) {
  long long spins = 0;
  while (*flag < value) {
    __nanosleep(200);
    if (++spins > (1ll << 24)) {  // ~3 s: never hang the device on a lost peer
      *timeout = 1;
      return;
    }
  }
  __threadfence_system();
}
void flag_signal(fq_ctx* ctx, double* flag, double value) {
  flag_signal_kernel<<<1, 1, 0, ctx->stream>>>(flag, value);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}
void flag_wait(fq_ctx* ctx, const double* flag, double value, int* d_timeout) {
  flag_wait_kernel<<<1, 1, 0, ctx->stream>>>(flag, value, d_timeout);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
}

void spmv_prepare(fq_ctx* ctx, fq_csr* a) {
  if (a->spmv_ready) return;
  const size_t nrows = a->row_end - a->row_begin;
  stream_build_blocks(ctx, a->row_ptr.p, nrows, a->nnz, a->rowblocks, a->nrowblocks);
  a->spmv_ready = true;
}

void spmv_apply(fq_ctx* ctx, const fq_csr* a, const double* x, double* y) {
  FQ_REQUIRE(a->spmv_ready, "spmv_prepare was not called");
  if (a->nrowblocks == 0) return;
  ScopedSpan span(ctx, "k4_spmv");
  stream_reduce(ctx, a->rowblocks.p, a->nrowblocks, a->row_ptr.p, a->col_idx.p, a->values.p, x, SpmvPolicy{y});
}

// Inverse diagonal for the Jacobi preconditioner (iterative/src/precond.rs:113-121).
__global__ void inv_diag_kernel(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col_idx,
                                const double* __restrict__ values, uint32_t nrows, uint32_t row_begin,
                                double* __restrict__ inv_diag, int* __restrict__ bad) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    double d = 0.0;
    for (uint32_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p)
      if (col_idx[p] == r + row_begin) d = values[p];
    if (d == 0.0) *bad = 1;  // no stored diagonal entry, or an explicit zero
    inv_diag[r] = __ddiv_rn(1.0, d);
  }
}

void csr_build_inv_diag(fq_ctx* ctx, fq_csr* a) {
  const size_t nrows = a->row_end - a->row_begin;
  if (a->inv_diag.n == nrows && nrows) return;
  a->inv_diag.alloc(nrows ? nrows : 1);
  if (!nrows) return;
  DevBuf<int> d_bad(1);
  FQ_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), ctx->stream));
  inv_diag_kernel<<<grid_for(nrows, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
      a->row_ptr.p, a->col_idx.p, a->values.p, uint32_t(nrows), uint32_t(a->row_begin), a->inv_diag.p, d_bad.p);
  fq_count_launch(ctx);
  FQ_CUDA(cudaGetLastError());
  int bad = 0;
  FQ_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (bad) {
    a->inv_diag.release();
    // the reference asserts a non-zero diagonal (iterative/src/precond.rs:101-104); dropped exact zeros count as missing
    throw Error(FQ_ERR_DEGENERATE, "Jacobi preconditioner: the matrix has a missing or zero diagonal entry");
  }
}

}  // namespace fq
