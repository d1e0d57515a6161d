// kuhn.cu — device generator for Kuhn grids: FaceIncidence tables of every
// grade and the signed squared edge lengths, in the reference's colex
// numbering, without ever materialising or sorting the simplices (kuhn.hpp).
#include <cub/cub.cuh>

#include "internal.hpp"
#include "kuhn.hpp"

namespace fq {

struct GridDev {
  int n;
  uint32_t shape[6];
  uint32_t vstride[6];
  uint64_t nverts;
};

__device__ __forceinline__ uint32_t dev_lower_mask(const GridDev& g, uint64_t w) {
  uint32_t B = 0;
  for (int a = 0; a < g.n; ++a) {
    const uint32_t nv = g.shape[a] + 1;
    if (w % nv != 0) B |= 1u << a;
    w /= nv;
  }
  return B;
}

__global__ void kuhn_cnt_kernel(GridDev g, const uint32_t* __restrict__ cnt_by_mask, uint32_t* __restrict__ cnt) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t w = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; w < g.nverts; w += stride)
    cnt[w] = cnt_by_mask[dev_lower_mask(g, w)];
}

// cell_faces of one grade for local cells [0, ncells); global cell = cell_offset + c.
__global__ void kuhn_faces_kernel(GridDev g, int ncelltypes, int nl, int ntypes, const uint16_t* __restrict__ ftype,
                                  const uint8_t* __restrict__ ftop, const uint16_t* __restrict__ rank_in,
                                  const uint32_t* __restrict__ vbase, uint64_t cell_offset, uint64_t ncells,
                                  uint32_t* __restrict__ out) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t c = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; c < ncells; c += stride) {
    const uint64_t gc = cell_offset + c;
    uint64_t box = gc / uint64_t(ncelltypes);
    const int t = int(gc % uint64_t(ncelltypes));
    uint32_t oc[6];
    for (int a = 0; a < g.n; ++a) {
      oc[a] = uint32_t(box % g.shape[a]);
      box /= g.shape[a];
    }
    for (int l = 0; l < nl; ++l) {
      const uint32_t top = ftop[t * nl + l];
      uint64_t w = 0;
      uint32_t B = 0;
      for (int a = 0; a < g.n; ++a) {
        const uint32_t ca = oc[a] + ((top >> a) & 1u);
        w += uint64_t(ca) * g.vstride[a];
        if (ca != 0) B |= 1u << a;
      }
      out[c * nl + l] = vbase[w] + rank_in[B * ntypes + ftype[t * nl + l]];
    }
  }
}

struct CoordDev {
  double vmin[6], side[6], diag[6], jit_h[6];  // jit_h = jitter * h_a (0 when jitter is off)
  int use_jitter;
};

__device__ __forceinline__ double dev_pseudo_random(uint64_t seed, uint64_t index) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + index * 0xD1B54A32D192ED03ull + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return __dsub_rn(__dmul_rn(__ddiv_rn(double(z >> 11), 9007199254740992.0), 2.0), 1.0);
}
// x_a(v) = (c_a / N_a) * side_a + min_a   (cartesian.rs:158-169), optionally jittered
__device__ __forceinline__ double dev_coord(const GridDev& g, const CoordDev& cd, int a, uint32_t ca, uint64_t v) {
  double x = __dadd_rn(__dmul_rn(__ddiv_rn(double(ca), double(g.shape[a])), cd.side[a]), cd.vmin[a]);
  if (cd.use_jitter) x = __dadd_rn(x, __dmul_rn(cd.jit_h[a], dev_pseudo_random(uint64_t(a), v)));
  return x;
}

// Edge lengths for the edges whose top vertex lies in [v_begin, v_end).
__global__ void kuhn_lengths_kernel(GridDev g, CoordDev cd, int ntypes, const uint8_t* __restrict__ chain_top,
                                    const uint16_t* __restrict__ rank_in, const uint32_t* __restrict__ vbase,
                                    uint64_t v_begin, uint64_t v_end, uint32_t edge_lo, double* __restrict__ lengths) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t w = v_begin + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; w < v_end; w += stride) {
    uint32_t wc[6];
    uint64_t rem = w;
    uint32_t B = 0;
    for (int a = 0; a < g.n; ++a) {
      wc[a] = uint32_t(rem % (g.shape[a] + 1));
      rem /= g.shape[a] + 1;
      if (wc[a] != 0) B |= 1u << a;
    }
    double xw[6];
    for (int a = 0; a < g.n; ++a) xw[a] = dev_coord(g, cd, a, wc[a], w);
    for (int t = 0; t < ntypes; ++t) {
      const uint32_t U = chain_top[t];
      if (U & ~B) continue;
      uint64_t vi = w;
      for (int a = 0; a < g.n; ++a)
        if ((U >> a) & 1u) vi -= g.vstride[a];
      // s = sum_a ((d_a * G_a) * d_a), accumulated as term + acc (metric/src/lib.rs:350-361)
      double acc = 0.0;
      for (int a = 0; a < g.n; ++a) {
        const uint32_t ca = wc[a] - ((U >> a) & 1u);
        const double xi = dev_coord(g, cd, a, ca, vi);
        const double d = __dsub_rn(xw[a], xi);
        const double term = __dmul_rn(__dmul_rn(d, cd.diag[a]), d);
        acc = (a == 0) ? term : __dadd_rn(term, acc);
      }
      lengths[vbase[w] + rank_in[B * ntypes + t] - edge_lo] = acc;
    }
  }
}

template <class T>
static void upload(DevBuf<T>& d, const std::vector<T>& h) {
  d.alloc(h.size() ? h.size() : 1);
  // blocks may be recycled by the caching allocator: order this blocking copy after everything in flight
  FQ_CUDA(cudaDeviceSynchronize());
  if (!h.empty()) FQ_CUDA(cudaMemcpy(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
}

void kuhn_build_mesh(fq_ctx* ctx, int dim, const size_t* shape, const double* vmin, const double* vmax,
                     const double* ambient_diag, double jitter, size_t slab_begin, size_t slab_end, fq_mesh* mesh) {
  const KuhnTables kt(dim);
  const KuhnGrid grid(dim, shape);
  FQ_REQUIRE(slab_begin < slab_end && slab_end <= shape[dim - 1], "invalid slab range");
  GridDev g{};
  g.n = dim;
  for (int a = 0; a < dim; ++a) {
    FQ_REQUIRE(grid.vstride[a] < (1ull << 32) && shape[a] < (1ull << 31), "grid too large for 32-bit vertex ids");
    g.shape[a] = uint32_t(shape[a]);
    g.vstride[a] = uint32_t(grid.vstride[a]);
  }
  FQ_REQUIRE(grid.nverts < (1ull << 32), "grid too large for 32-bit vertex ids");
  g.nverts = grid.nverts;
  const uint64_t boxes_per_layer = grid.nboxes / shape[dim - 1];
  const uint64_t cells_per_layer = boxes_per_layer * uint64_t(kt.ncelltypes);
  mesh->dim = dim;
  // owner-computes: the slab owns box layers [slab_begin, slab_end) and the
  // simplices whose top vertex lies in vertex layers (slab_begin, slab_end]
  // (plus layer 0 for the first slab).  The cells incident to those simplices
  // also include the box layer slab_end (its bottom faces), held as a halo.
  const size_t own_end = slab_end;
  if (slab_end < shape[dim - 1]) slab_end += 1;
  mesh->cell_offset = size_t(cells_per_layer * slab_begin);
  mesh->ncells = size_t(cells_per_layer * (slab_end - slab_begin));
  mesh->nowned_cells = size_t(cells_per_layer * (own_end - slab_begin));
  mesh->nsimplices.assign(size_t(dim) + 1, 0);
  mesh->cell_faces.clear();
  mesh->cell_faces.resize(size_t(dim) + 1);
  mesh->id_lo.assign(size_t(dim) + 1, 0);
  mesh->id_hi.assign(size_t(dim) + 1, 0);
  mesh->own_lo.assign(size_t(dim) + 1, 0);
  mesh->own_hi.assign(size_t(dim) + 1, 0);
  // vertex layers touched by the slab: z in [slab_begin, slab_end]; owned rows
  // are the simplices whose top vertex has z in (slab_begin, slab_end], plus
  // the bottom layer z = 0 for the first slab.
  const uint64_t layer = grid.vstride[dim - 1];
  const uint64_t v_lo = layer * slab_begin, v_hi = layer * (slab_end + 1);
  const uint64_t own_v_lo = slab_begin == 0 ? 0 : layer * (slab_begin + 1);
  const uint64_t own_v_hi = layer * (own_end + 1);
  const int block = 256;
  DevBuf<uint32_t> cnt(size_t(grid.nverts) + 1);
  std::vector<DevBuf<uint32_t>> vbase(size_t(dim) + 1);
  DevBuf<uint8_t> cub_tmp;
  for (int j = 0; j <= dim; ++j) {
    const KuhnGrade& kg = kt.grades[size_t(j)];
    DevBuf<uint32_t> cnt_by_mask;
    upload(cnt_by_mask, kg.cnt);
    kuhn_cnt_kernel<<<grid_for(grid.nverts, block, ctx->sm_count), block, 0, ctx->stream>>>(g, cnt_by_mask.p, cnt.p);
    fq_count_launch(ctx);
    FQ_CUDA(cudaMemsetAsync(cnt.p + grid.nverts, 0, sizeof(uint32_t), ctx->stream));
    // totals must fit 32 bits: check on host with 64-bit arithmetic
    uint64_t total = 0;
    {
      // sum over masks of cnt[B] * (#vertices with lower mask B)
      const uint32_t full = (1u << dim) - 1;
      for (uint32_t B = 0; B <= full; ++B) {
        uint64_t nv = 1;
        for (int a = 0; a < dim; ++a) nv *= (B >> a & 1u) ? shape[a] : 1;
        total += nv * kg.cnt[B];
      }
    }
    FQ_REQUIRE(total < (1ull << 32), "more than 2^32 simplices of one grade: not supported");
    mesh->nsimplices[size_t(j)] = size_t(total);
    vbase[size_t(j)].alloc(size_t(grid.nverts) + 1);
    size_t tmp_bytes = 0;
    FQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.p, vbase[size_t(j)].p, int(grid.nverts + 1),
                                          ctx->stream));
    if (cub_tmp.n < tmp_bytes) cub_tmp.alloc(tmp_bytes);
    FQ_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp_bytes, cnt.p, vbase[size_t(j)].p, int(grid.nverts + 1),
                                          ctx->stream));
    fq_count_launch(ctx, 2);
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
    // id ranges: vbase at the slab's vertex bounds
    uint32_t h[4];
    FQ_CUDA(cudaMemcpy(&h[0], vbase[size_t(j)].p + v_lo, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    FQ_CUDA(cudaMemcpy(&h[1], vbase[size_t(j)].p + v_hi, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    FQ_CUDA(cudaMemcpy(&h[2], vbase[size_t(j)].p + own_v_lo, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    FQ_CUDA(cudaMemcpy(&h[3], vbase[size_t(j)].p + own_v_hi, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    mesh->id_lo[size_t(j)] = h[0];
    mesh->id_hi[size_t(j)] = h[1];
    mesh->own_lo[size_t(j)] = h[2];
    mesh->own_hi[size_t(j)] = h[3];
  }
  // cells are themselves the grade-dim simplices: owned = all local cells
  for (int j = 0; j <= dim; ++j) {
    const KuhnGrade& kg = kt.grades[size_t(j)];
    const int nl = nlocal(dim, j);
    DevBuf<uint16_t> ftype, rank_in;
    DevBuf<uint8_t> ftop;
    upload(ftype, kt.ftype[size_t(j)]);
    upload(ftop, kt.ftop[size_t(j)]);
    upload(rank_in, kg.rank_in);
    mesh->cell_faces[size_t(j)].alloc(mesh->ncells * size_t(nl));
    kuhn_faces_kernel<<<grid_for(mesh->ncells, block, ctx->sm_count), block, 0, ctx->stream>>>(
        g, kt.ncelltypes, nl, kg.ntypes, ftype.p, ftop.p, rank_in.p, vbase[size_t(j)].p, mesh->cell_offset,
        mesh->ncells, mesh->cell_faces[size_t(j)].p);
    fq_count_launch(ctx);
    FQ_CUDA(cudaGetLastError());
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  // edge lengths
  {
    const KuhnGrade& kg = kt.grades[1];
    std::vector<uint8_t> tops;
    for (const auto& ch : kg.chains) tops.push_back(ch.back());
    DevBuf<uint8_t> d_tops;
    DevBuf<uint16_t> rank_in;
    upload(d_tops, tops);
    upload(rank_in, kg.rank_in);
    CoordDev cd{};
    for (int a = 0; a < dim; ++a) {
      cd.vmin[a] = vmin ? vmin[a] : 0.0;
      const double vmx = vmax ? vmax[a] : 1.0;
      cd.side[a] = vmx - cd.vmin[a];
      cd.diag[a] = ambient_diag ? ambient_diag[a] : 1.0;
      // oracle.jitter_coords: h = (max - min) / shape over the generated coordinates
      const double h = ((double(shape[a]) / double(shape[a])) * cd.side[a] + cd.vmin[a] - cd.vmin[a]) / double(shape[a]);
      cd.jit_h[a] = jitter * h;
    }
    cd.use_jitter = jitter != 0.0;
    mesh->edge_lo = mesh->id_lo[1];
    const size_t ne = mesh->id_hi[1] - mesh->id_lo[1];
    mesh->lengths.alloc(ne ? ne : 1);
    kuhn_lengths_kernel<<<grid_for(v_hi - v_lo, block, ctx->sm_count), block, 0, ctx->stream>>>(
        g, cd, kg.ntypes, d_tops.p, rank_in.p, vbase[1].p, v_lo, v_hi, uint32_t(mesh->edge_lo), mesh->lengths.p);
    fq_count_launch(ctx);
    FQ_CUDA(cudaGetLastError());
    FQ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  // vertex bricks for the tile-fused numeric assembly (tile.cu)
  mesh->cell_type_period = kt.ncelltypes;
  tile_cluster_kuhn(ctx, mesh, dim, shape, slab_begin, slab_end);
}

}  // namespace fq
