// Stage B -- small dense linear algebra for one Team (a CTA on the GPU, one lane in the host model): register-tiled
// matrix products out of shared memory and a Gauss-Jordan solver with partial pivoting that needs ONE team barrier per pivot
// column.  Used by the layer-operator kernel (vl_layerops.cuh) and the interaction-principle solver (vl_add.cuh).
#pragma once
#include "vl_common.cuh"
#include <cstring>

namespace vlk {

#ifdef AB_HOST_MODEL
inline void vl_atomic_max(unsigned long long *p, unsigned long long v) { if (v > *p) *p = v; }
inline unsigned long long vl_dbits(double x) { unsigned long long u; memcpy(&u, &x, 8); return u; }
#else
DEVFN void vl_atomic_max(unsigned long long *p, unsigned long long v) { atomicMax(p, v); }
DEVFN unsigned long long vl_dbits(double x) { return (unsigned long long)__double_as_longlong(x); }
#endif
// pivot-candidate key: magnitude bits with the low 7 bits replaced by the row number
DEVFN unsigned long long vl_pivot_key(double v, int r) { return (vl_dbits(fabs(v)) & ~127ull) | (unsigned long long)r; }

// acc(r, c) = sum_k A[r*lda + k] * B[k*ldb + c] for an M x N result; every team item owns a TM x TN register tile and hands the
// finished sums to epi(r, c, acc).  Items run over the column tiles first, so the lanes of a warp read neighbouring B columns
// (distinct banks) and broadcast the few distinct A rows they share.
template <int TM, int TN, class Team, class Epi>
DEVFN void team_gemm(const Team &T, int M, int N, int K, const double *AB_RESTRICT A, int lda, const double *AB_RESTRICT B, int ldb, Epi epi) {
  const int tn = (N + TN - 1) / TN, tm = (M + TM - 1) / TM;
  for (int item = T.rank; item < tm * tn; item += T.size) {
    const int r0 = (item / tn) * TM, c0 = (item % tn) * TN;
    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) acc[i][j] = 0.0;
    const bool full = (r0 + TM <= M) && (c0 + TN <= N);
    if (full) {
      const double *a = A + (size_t)r0 * lda, *b = B + c0;
#pragma unroll 4
      for (int k = 0; k < K; k++) {
        double av[TM], bv[TN];
#pragma unroll
        for (int i = 0; i < TM; i++) av[i] = a[(size_t)i * lda + k];
#pragma unroll
        for (int j = 0; j < TN; j++) bv[j] = b[(size_t)k * ldb + j];
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
          for (int j = 0; j < TN; j++) acc[i][j] += av[i] * bv[j];
      }
    } else {
      for (int k = 0; k < K; k++) {
        double av[TM], bv[TN];
#pragma unroll
        for (int i = 0; i < TM; i++) av[i] = (r0 + i < M) ? A[(size_t)(r0 + i) * lda + k] : 0.0;
#pragma unroll
        for (int j = 0; j < TN; j++) bv[j] = (c0 + j < N) ? B[(size_t)k * ldb + c0 + j] : 0.0;
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
          for (int j = 0; j < TN; j++) acc[i][j] += av[i] * bv[j];
      }
    }
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++)
        if (r0 + i < M && c0 + j < N) epi(r0 + i, c0 + j, acc[i][j]);
  }
}

// tile shape by problem size: 4 x 3 tiles cover 36 x 36 with 108 items (one pass of a 128-thread team)
template <class Team, class Epi>
DEVFN void team_gemm_auto(const Team &T, int M, int N, int K, const double *A, int lda, const double *B, int ldb, Epi epi) {
  if (M <= 12 || T.size > 192) team_gemm<2, 3>(T, M, N, K, A, lda, B, ldb, epi);
  else team_gemm<4, 3>(T, M, N, K, A, lda, B, ldb, epi);
}

// Gauss-Jordan elimination of the n x W array G (row stride ld; the first n columns are the matrix, the rest right-hand sides) with
// partial pivoting over the rows that have not been pivots yet.  Rows are never swapped or scaled: piv[c] is the row that ends up
// holding unknown c, and the solution is X(c, j) = G[piv[c]*ld + n + j] / G[piv[c]*ld + c].  Columns <= c are dead after step c
// (nobody reads or writes them again except the pivot element), so a step needs a single barrier: the threads that update column
// c + 1 also publish the next pivot candidate through an atomic max on (|value| bits, row) -- the low 7 bits of the magnitude are
// replaced by the row number, which perturbs the choice among candidates that agree to 1e-14 only.
// best: 3 words of team scratch; returns false when a pivot column has no non-zero candidate.
template <class Team>
DEVFN bool team_gauss_jordan(const Team &T, int n, int W, double *AB_RESTRICT G, int ld, int *piv, unsigned long long *best) {
  // items: (row r, column group cg); a team of 128 lanes and n = 36 gives 3 column groups
  int CG = T.size / n;
  if (CG < 1) CG = 1;
  const int nitem = n * CG;
  for (int i = T.rank; i < 3; i += T.size) best[i] = 0ull;
  T.sync();
  for (int r = T.rank; r < n; r += T.size) {
    vl_atomic_max(&best[0], vl_pivot_key(G[(size_t)r * ld], r));
  }
  unsigned long long used0 = 0ull, used1 = 0ull;   // rows that have been pivots (n <= 128)
  bool ok = true;
  for (int c = 0; c < n; c++) {
    T.sync();
    const unsigned long long key = best[c % 3];
    if ((key >> 7) == 0ull) { ok = false; break; }   // every candidate is zero (or denormal): singular
    const int p = (int)(key & 127ull);
    if (p < 64) used0 |= 1ull << p; else used1 |= 1ull << (p - 64);
    if (T.rank == 0) { piv[c] = p; best[(c + 2) % 3] = 0ull; }
    const double *prow = G + (size_t)p * ld;
    const double ipiv = 1.0 / prow[c];
    for (int item = T.rank; item < nitem; item += T.size) {
      const int r = item % n, cg = item / n;
      if (r == p) continue;
      double *row = G + (size_t)r * ld;
      const double f = row[c] * ipiv;
      if (f != 0.0) {
        for (int j = c + 1 + cg; j < W; j += CG) row[j] -= f * prow[j];
      }
      if (cg == 0 && c + 1 < n) {
        const bool u = r < 64 ? ((used0 >> r) & 1ull) : ((used1 >> (r - 64)) & 1ull);
        if (!u) {
          vl_atomic_max(&best[(c + 1) % 3], vl_pivot_key(row[c + 1], r));
        }
      }
    }
  }
  T.sync();
  return ok;
}

}  // namespace vlk
