// oracle_capi.cpp — C ABI over fq_oracle.hpp so tests and bench.py can drive
// the CPU oracle through ctypes.  TEST INFRASTRUCTURE ONLY (see the header of
// fq_oracle.hpp).  Also holds the timed, rayon-like multi-threaded assembly
// used as `cpu_baseline` (std::thread static chunks over cells, ordered concat,
// then the *serial* COO->CSR, mirroring galerkin.rs:160-187).
#include "fq_oracle.hpp"

#include <chrono>
#include <thread>

using namespace fqo;

namespace {
double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
thread_local std::string g_err;
}  // namespace

extern "C" {

const char* fqo_last_error() { return g_err.c_str(); }

int fqo_max_threads() {
  const unsigned n = std::thread::hardware_concurrency();
  return n ? int(n) : 1;
}

// ---- tables -------------------------------------------------------------
int64_t fqo_binomial(int n, int k) { return binomial(n, k); }

// combinations(n, card) flattened, returns count
int64_t fqo_combinations(int n, int card, int32_t* out) {
  const auto cs = combinations(n, card);
  if (out) {
    size_t p = 0;
    for (const Comb& c : cs)
      for (int v : c) out[p++] = v;
  }
  return int64_t(cs.size());
}
int64_t fqo_permutations(int n, int32_t* out, double* signs) {
  const auto ps = permutations_all(n);
  size_t p = 0;
  for (size_t i = 0; i < ps.size(); ++i) {
    if (out)
      for (int v : ps[i].p) out[p++] = v;
    if (signs) signs[i] = ps[i].sign;
  }
  return int64_t(ps.size());
}
// boundary operator of the reference cell, row-major; returns rows*cols
int64_t fqo_unit_boundary_operator(int n, int k, double* out, int* rows, int* cols) {
  const Mat b = unit_boundary_operator(n, k);
  *rows = b.r;
  *cols = b.c;
  if (out) std::copy(b.a.begin(), b.a.end(), out);
  return int64_t(b.a.size());
}
int64_t fqo_difbarys_power(int n, int k, double* out, int* rows, int* cols) {
  const Mat b = exterior_power(unit_difbarys(n), k);
  *rows = b.r;
  *cols = b.c;
  if (out) std::copy(b.a.begin(), b.a.end(), out);
  return int64_t(b.a.size());
}
double fqo_pseudo_random(uint64_t seed, uint64_t index) { return pseudo_random(seed, index); }
void fqo_unit_simplex_lengths_sq(int n, double* out) {
  const auto s = unit_simplex_lengths_sq(n);
  std::copy(s.begin(), s.end(), out);
}

// ---- element matrices ---------------------------------------------------
// One element matrix from the cell's C(n+1,2) signed squared edge lengths
// (local colex pair order).  out is row-major rows x cols.
int fqo_elmat(int kind, int n, int k, const double* lengths_sq, double* out, int* rows, int* cols) {
  try {
    const Mat g = (n >= 1) ? metric_from_lengths(n, lengths_sq) : Mat(0, 0);
    Mat el;
    if (kind == LUMPED) {
      el = lumped_element(g);
    } else {
      PairingTables t(n, k, kind);
      el = pairing_element(t, g);
    }
    *rows = el.r;
    *cols = el.c;
    if (out) std::copy(el.a.begin(), el.a.end(), out);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// metric / inverse / volume of one cell (debug + parity hooks)
int fqo_cell_geometry(int n, const double* lengths_sq, double* g_out, double* ginv_out, double* vol) {
  const Mat g = metric_from_lengths(n, lengths_sq);
  Mat gi;
  if (!try_inverse(g, gi)) return -1;
  std::copy(g.a.begin(), g.a.end(), g_out);
  std::copy(gi.a.begin(), gi.a.end(), ginv_out);
  *vol = cell_volume(g);
  return 0;
}

// ---- complex ------------------------------------------------------------
void* fqo_complex_from_cells(int dim, int64_t ncells, const int64_t* cells) {
  std::vector<int64_t> c(cells, cells + size_t(ncells) * (dim + 1));
  return new Complex(complex_from_cells(dim, c));
}
void* fqo_complex_kuhn(int dim, const int64_t* shape) {
  return new Complex(complex_from_cells(dim, kuhn_cells(dim, shape)));
}
void fqo_complex_destroy(void* h) { delete static_cast<Complex*>(h); }
int fqo_complex_dim(void* h) { return static_cast<Complex*>(h)->dim; }
int64_t fqo_complex_nsimplices(void* h, int j) { return static_cast<Complex*>(h)->nsimplices(j); }
void fqo_complex_skeleton(void* h, int j, int64_t* out) {
  const auto& s = static_cast<Complex*>(h)->skeleton[j];
  std::copy(s.begin(), s.end(), out);
}
void fqo_complex_cell_faces(void* h, int j, int64_t* out) {
  const auto& s = static_cast<Complex*>(h)->cell_faces[j];
  std::copy(s.begin(), s.end(), out);
}
void fqo_kuhn_vertex_coords(int dim, const int64_t* shape, const double* min, const double* max,
                            double* out) {
  const auto x = kuhn_vertex_coords(dim, shape, min, max);
  std::copy(x.begin(), x.end(), out);
}
void fqo_edge_lengths_sq(void* h, int ambient_dim, const double* coords, const double* ambient_diag,
                         double* out) {
  const auto s = edge_lengths_sq(*static_cast<Complex*>(h), ambient_dim, coords, ambient_diag);
  std::copy(s.begin(), s.end(), out);
}

// batch of element matrices for cells [c0,c1), row-major [cell][rows][cols]
int fqo_elmat_batch(void* h, const double* lengths_sq, int kind, int k, int64_t c0, int64_t c1,
                    double* out) {
  try {
    const Complex& cx = *static_cast<Complex*>(h);
    PairingTables t(cx.dim, kind == LUMPED ? 0 : k, kind == LUMPED ? MASS : kind);
    size_t p = 0;
    for (int64_t c = c0; c < c1; ++c) {
      const Mat el = cell_element(cx, lengths_sq, kind, k, c, &t);
      std::copy(el.a.begin(), el.a.end(), out + p);
      p += el.a.size();
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// ---- assembly -----------------------------------------------------------
// Multi-threaded like the reference: per-thread triplet buffers over static
// contiguous cell chunks, concatenated in chunk (= cell) order, then the
// serial COO->CSR.  times[0] = element+triplet phase, times[1] = COO->CSR.
void* fqo_assemble(void* h, const double* lengths_sq, int kind, int k, int drop_zeros, int nthreads,
                   double* times) {
  try {
    const Complex& cx = *static_cast<Complex*>(h);
    int tg, rg;
    kind_grades(kind, k, tg, rg);
    const idx_t nc = cx.ncells();
    if (nthreads < 1) nthreads = 1;
    std::vector<Triplets> parts(static_cast<size_t>(nthreads));
    const double t0 = now_s();
    {
      std::vector<std::thread> pool;
      for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&, t] {
          const idx_t c0 = nc * t / nthreads, c1 = nc * (t + 1) / nthreads;
          assemble_triplets(cx, lengths_sq, kind, k, c0, c1, drop_zeros != 0, parts[size_t(t)]);
        });
      for (auto& th : pool) th.join();
    }
    Triplets all;
    size_t total = 0;
    for (const auto& p : parts) total += p.rows.size();
    all.rows.reserve(total);
    all.cols.reserve(total);
    all.vals.reserve(total);
    for (auto& p : parts) {
      all.rows.insert(all.rows.end(), p.rows.begin(), p.rows.end());
      all.cols.insert(all.cols.end(), p.cols.begin(), p.cols.end());
      all.vals.insert(all.vals.end(), p.vals.begin(), p.vals.end());
      Triplets().rows.swap(p.rows);
      Triplets().cols.swap(p.cols);
      Triplets().vals.swap(p.vals);
    }
    const double t1 = now_s();
    Csr* m = new Csr(coo_to_csr(cx.nsimplices(tg), cx.nsimplices(rg), all));
    const double t2 = now_s();
    if (times) {
      times[0] = t1 - t0;
      times[1] = t2 - t1;
    }
    return m;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void* fqo_csr_from_arrays(int64_t nrows, int64_t ncols, const int64_t* row_ptr, const int64_t* col_idx,
                          const double* values) {
  Csr* m = new Csr;
  m->nrows = nrows;
  m->ncols = ncols;
  m->row_ptr.assign(row_ptr, row_ptr + nrows + 1);
  const int64_t nnz = row_ptr[nrows];
  m->col_idx.assign(col_idx, col_idx + nnz);
  m->values.assign(values, values + nnz);
  return m;
}
void fqo_csr_destroy(void* m) { delete static_cast<Csr*>(m); }
void fqo_csr_shape(void* m, int64_t* nrows, int64_t* ncols, int64_t* nnz) {
  const Csr& a = *static_cast<Csr*>(m);
  *nrows = a.nrows;
  *ncols = a.ncols;
  *nnz = int64_t(a.values.size());
}
void fqo_csr_copy(void* m, int64_t* row_ptr, int64_t* col_idx, double* values) {
  const Csr& a = *static_cast<Csr*>(m);
  std::copy(a.row_ptr.begin(), a.row_ptr.end(), row_ptr);
  std::copy(a.col_idx.begin(), a.col_idx.end(), col_idx);
  std::copy(a.values.begin(), a.values.end(), values);
}

// ---- SpMV / Krylov --------------------------------------------------------
void fqo_spmv(void* m, const double* x, double* y) { spmv(*static_cast<Csr*>(m), x, y); }
// timed repetitions of the serial reference SpMV; returns seconds per apply
double fqo_spmv_timed(void* m, const double* x, double* y, int reps) {
  const Csr& a = *static_cast<Csr*>(m);
  const double t0 = now_s();
  for (int r = 0; r < reps; ++r) spmv(a, x, y);
  return (now_s() - t0) / reps;
}
// row-parallel variant ("better than reference", labelled as such by callers)
double fqo_spmv_parallel_timed(void* m, const double* x, double* y, int reps, int nthreads) {
  const Csr& a = *static_cast<Csr*>(m);
  const double t0 = now_s();
  if (nthreads < 1) nthreads = 1;
  for (int r = 0; r < reps; ++r) {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back([&, t] {
        const idx_t r0 = a.nrows * t / nthreads, r1 = a.nrows * (t + 1) / nthreads;
        for (idx_t i = r0; i < r1; ++i) {
          double acc = 0.0;
          for (int64_t p = a.row_ptr[size_t(i)]; p < a.row_ptr[size_t(i) + 1]; ++p)
            acc = acc + a.values[size_t(p)] * x[a.col_idx[size_t(p)]];
          y[i] = acc;
        }
      });
    for (auto& th : pool) th.join();
  }
  return (now_s() - t0) / reps;
}

// precond: 0 = identity, 1 = Jacobi (inverse diagonal of the matrix)
static ApplyFn make_precond(const Csr& a, int precond) {
  if (precond == 0) return [](const std::vector<double>& r, std::vector<double>& z) { z = r; };
  std::vector<double> dinv(size_t(a.nrows), 1.0);
  for (idx_t i = 0; i < a.nrows; ++i)
    for (int64_t p = a.row_ptr[size_t(i)]; p < a.row_ptr[size_t(i) + 1]; ++p)
      if (a.col_idx[size_t(p)] == i) dinv[size_t(i)] = 1.0 / a.values[size_t(p)];
  return [dinv](const std::vector<double>& r, std::vector<double>& z) {
    z.resize(r.size());
    for (size_t i = 0; i < r.size(); ++i) z[i] = dinv[i] * r[i];
  };
}
int fqo_cg(void* m, int precond, const double* b, double rtol, int64_t max_iters, double* x,
           int64_t* iters, double* residual) {
  const Csr& a = *static_cast<Csr*>(m);
  std::vector<double> bb(b, b + a.nrows), xx;
  ApplyFn op = [&a](const std::vector<double>& v, std::vector<double>& y) {
    y.resize(size_t(a.nrows));
    spmv(a, v.data(), y.data());
  };
  const Report r = cg(op, make_precond(a, precond), bb, rtol, max_iters, xx);
  std::copy(xx.begin(), xx.end(), x);
  *iters = r.iters;
  *residual = r.residual;
  return r.converged ? 1 : 0;
}
int fqo_minres(void* m, int precond, const double* b, double rtol, int64_t max_iters, double* x,
               int64_t* iters, double* residual) {
  const Csr& a = *static_cast<Csr*>(m);
  std::vector<double> bb(b, b + a.nrows), xx;
  ApplyFn op = [&a](const std::vector<double>& v, std::vector<double>& y) {
    y.resize(size_t(a.nrows));
    spmv(a, v.data(), y.data());
  };
  const Report r = minres(op, make_precond(a, precond), bb, rtol, max_iters, xx);
  std::copy(xx.begin(), xx.end(), x);
  *iters = r.iters;
  *residual = r.residual;
  return r.converged ? 1 : 0;
}

}  // extern "C"
