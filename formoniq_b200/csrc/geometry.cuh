// geometry.cuh — metric, inverse and volume of one cell from its squared edge lengths, any dimension <= 10.
// Shared by the generic element path (elmat.cu, dim >= 4) and the quadrature forms (quadform.cu).
#pragma once

namespace fq {

constexpr int kMaxDim = 10;

// Metric by polarisation (regge/src/lengths/simplex.rs:308-326), inverse and volume ((1/n!) sqrt|det g|,
// regge/src/lib.rs:26-28) for one cell; mirrors the closed-form / LU split of nalgebra (4x4 cofactors, LU with partial
// pivoting beyond).  Returns false on a singular metric.
__device__ inline bool geometry_generic(int n, const double* s, double* ginv /*n*n row-major*/, double* vol) {
  double g[kMaxDim * kMaxDim];
  auto eidx = [](int i, int j) { return i + j * (j - 1) / 2; };  // i<j : C(i,1)+C(j,2)
  for (int i = 0; i < n; ++i) g[i * n + i] = s[eidx(0, i + 1)];
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      const double v = 0.5 * ((s[eidx(0, i + 1)] + s[eidx(0, j + 1)]) - s[eidx(i + 1, j + 1)]);
      g[i * n + j] = v;
      g[j * n + i] = v;
    }
  // LU (partial pivoting) — determinant always, inverse for n >= 5
  double lu[kMaxDim * kMaxDim];
  int piv[kMaxDim];
  for (int i = 0; i < n * n; ++i) lu[i] = g[i];
  int nswaps = 0;
  bool ok = true;
  for (int i = 0; i < n; ++i) {
    int p = i;
    double best = fabs(lu[i * n + i]);
    for (int r = i + 1; r < n; ++r)
      if (fabs(lu[r * n + i]) > best) best = fabs(lu[r * n + i]), p = r;
    piv[i] = p;
    if (best == 0.0) {
      ok = false;
      continue;
    }
    if (p != i) {
      ++nswaps;
      for (int c = 0; c < n; ++c) {
        const double t = lu[i * n + c];
        lu[i * n + c] = lu[p * n + c];
        lu[p * n + c] = t;
      }
    }
    const double d = lu[i * n + i];
    for (int r = i + 1; r < n; ++r) lu[r * n + i] = lu[r * n + i] / d;
    for (int c = i + 1; c < n; ++c) {
      const double pc = lu[i * n + c];
      for (int r = i + 1; r < n; ++r) lu[r * n + c] = lu[r * n + c] - lu[r * n + i] * pc;
    }
  }
  if (!ok) return false;
  double det = 1.0;
  for (int i = 0; i < n; ++i) det = det * lu[i * n + i];
  if (nswaps & 1) det = -det;
  double nf = 1.0;
  for (int i = 2; i <= n; ++i) nf *= double(i);
  *vol = (1.0 / nf) * sqrt(fabs(det));
  if (n == 4) {
    double m[16], o[16];
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) m[j * 4 + i] = g[i * 4 + j];
    o[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] +
           m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    o[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] -
           m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    o[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] +
           m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    o[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] -
           m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    o[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] -
           m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    o[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] +
           m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    o[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] -
           m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    o[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] +
           m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    o[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] +
           m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    o[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] -
           m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    o[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] +
            m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    o[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] -
            m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    o[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] -
            m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    o[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] +
            m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    o[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] -
            m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    o[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] +
            m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double d4 = m[0] * o[0] + m[1] * o[4] + m[2] * o[8] + m[3] * o[12];
    if (d4 == 0.0) return false;
    const double inv = 1.0 / d4;
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) ginv[i * 4 + j] = o[j * 4 + i] * inv;
    return true;
  }
  // inverse by LU solves, column by column
  for (int col = 0; col < n; ++col) {
    double b[kMaxDim];
    for (int i = 0; i < n; ++i) b[i] = (i == col) ? 1.0 : 0.0;
    for (int i = 0; i < n; ++i) {
      const double t = b[i];
      b[i] = b[piv[i]];
      b[piv[i]] = t;
    }
    for (int i = 0; i < n; ++i)
      for (int r = i + 1; r < n; ++r) b[r] = b[r] - lu[r * n + i] * b[i];
    for (int i = n - 1; i >= 0; --i) {
      b[i] = b[i] / lu[i * n + i];
      for (int r = 0; r < i; ++r) b[r] = b[r] - lu[r * n + i] * b[i];
    }
    for (int r = 0; r < n; ++r) ginv[r * n + col] = b[r];
  }
  return true;
}

}  // namespace fq
