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
      double *AB_RESTRICT row = G + (size_t)r * ld;          // r != p: the two rows never overlap
      const double *AB_RESTRICT pr = prow;
      const double f = row[c] * ipiv;
      if (f != 0.0) {
        int j = c + 1 + cg;
        for (; j + 3 * CG < W; j += 4 * CG) {                 // loads first: four independent updates in flight
          const double p0 = pr[j], p1 = pr[j + CG], p2 = pr[j + 2 * CG], p3 = pr[j + 3 * CG];
          const double a0 = row[j], a1 = row[j + CG], a2 = row[j + 2 * CG], a3 = row[j + 3 * CG];
          row[j] = a0 - f * p0; row[j + CG] = a1 - f * p1; row[j + 2 * CG] = a2 - f * p2; row[j + 3 * CG] = a3 - f * p3;
        }
        for (; j < W; j += CG) row[j] -= f * pr[j];
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

// ---------------------------------------------------------------------------------------------------------------------------------
// E = A^-1 for an n x n matrix in team scratch (row stride ld).  Generic form (host model, warp teams): Gauss-Jordan on [A | 1] in
// the work array wk (n x ldw doubles, ldw = (2n)|1) + 3 pivot words; A is left untouched.
HDFN int inv_ldw(int n) { return (2 * n) | 1; }
HDFN size_t inv_work_doubles_generic(int n) { return (size_t)n * inv_ldw(n) + n + 4; }
HDFN size_t inv_work_doubles_cta(int n) { return (size_t)n * (n | 1) + 3 * n + 2; }       // scratch of the CTA form below
template <class Team>
DEVFN bool team_inverse_generic(const Team &T, int n, const double *A, int lda, double *E, int lde, double *wk, int *piv) {
  const int ldw = inv_ldw(n);
  double *dinv = wk + (size_t)n * ldw;
  unsigned long long *best = reinterpret_cast<unsigned long long *>(dinv + n);
  for (int idx = T.rank; idx < n * n; idx += T.size) {
    const int r = idx / n, c = idx % n;
    wk[(size_t)r * ldw + c] = A[(size_t)r * lda + c]; wk[(size_t)r * ldw + n + c] = (r == c) ? 1.0 : 0.0;
  }
  if (!team_gauss_jordan(T, n, 2 * n, wk, ldw, piv, best)) return false;
  for (int e = T.rank; e < n; e += T.size) dinv[e] = 1.0 / wk[(size_t)piv[e] * ldw + e];
  T.sync();
  for (int idx = T.rank; idx < n * n; idx += T.size) { const int e = idx / n, r = idx % n; E[(size_t)e * lde + r] = wk[(size_t)piv[e] * ldw + n + r] * dinv[e]; }
  T.sync();
  return true;
}

#ifndef AB_HOST_MODEL
// 1/x to full double precision from the hardware approximation + two Newton steps (the quotient only scales elimination multipliers)
DEVFN double vl_fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = y * (2.0 - x * y);
  y = y * (2.0 - x * y);
  return y;
}

// CTA form: in-place Gauss-Jordan inversion with the matrix REGISTER-RESIDENT and mirrored in shared memory.  Thread (r, cg) of the
// first n*CG threads (CG = blockDim / n column groups) keeps the elements a(r, cg + CG*jj), jj < NJ, of row r in registers for the
// whole elimination and stores every updated element to the mirror Mw (n x ldm, ldm odd: conflict-free for lanes over rows), from
// which the other threads take the pivot row (broadcast loads); the next pivot column (multipliers, pivot candidates) travels through
// a double-buffered column array, so that column c of the mirror can be overwritten in step c -- ONE barrier per pivot column.  Partial pivoting over the rows that have not been pivots (every warp finds the pivot by itself: float magnitudes
// with the row number in the low 6 bits, one warp-wide integer max), no row swaps, no row scaling: piv[c] = pivot row of step c,
// stepof[r] = its inverse; the physical column c holds, after step c, the column of the inverse that belongs to row piv[c] (the
// unit column it replaces is implicit before).  E(stepof[r], piv[j]) = a(r, j) / pivot(stepof[r]).
// The pivot row's threads leave their mirror entries alone in their own step (the row is being read by all; racecheck-clean).
// N_, CG_ > 0: compile-time n and column-group count (n = 36 with 128 threads: 3 groups of 12 columns).
// Scratch: Mw (n x ldm doubles), dinv_s (n), column buffer (2n); ints piv (n), stepof (n).
HDFN int inv_ldm(int n) { return n | 1; }
template <int NJ, int N_, int CG_>
DEVFN bool cta_inverse(int n_, const double *A, int lda, double *E, int lde, double *wk, int *piv) {
  const int n = N_ > 0 ? N_ : n_, CG = CG_ > 0 ? CG_ : (int)blockDim.x / n, ldm = inv_ldm(n);
  const int rank = threadIdx.x, lane = threadIdx.x & 31;
  const bool active = rank < n * CG;
  const int r = active ? rank % n : 0, cg = active ? rank / n : 0;
  double *Mw = wk, *dinv_s = wk + (size_t)n * ldm, *colb = dinv_s + n;
  int *stepof = piv + n;
  double a[NJ];
  double *mrow = Mw + (size_t)r * ldm + cg;
#pragma unroll
  for (int jj = 0; jj < NJ; jj++) {
    const int j = cg + CG * jj;
    a[jj] = (active && j < n) ? A[(size_t)r * lda + j] : 0.0;
    if (active && j < n) mrow[CG * jj] = a[jj];
  }
  if (active && cg == 0) colb[r] = a[0];                       // column 0
  unsigned used0 = 0u, used1 = 0u;                             // rows that have been pivots (n <= 64), the same in every thread
  bool ok = true;
  for (int c = 0; c < n; c++) {
    __syncthreads();                                           // every update of step c - 1 is in the mirror / the column buffer
    const double *cs = colb + (c & 1) * n;                     // column c as of the end of step c - 1
    double *cn = colb + ((c + 1) & 1) * n;
    // pivot search, redundantly in every warp: rows lane and lane + 32 of column c
    unsigned key = 0u;
    if (lane < n && !((used0 >> lane) & 1u)) key = (__float_as_uint(fabsf((float)cs[lane])) & ~63u) | (unsigned)lane;
    if (lane + 32 < n && !((used1 >> lane) & 1u)) {
      const unsigned k2 = (__float_as_uint(fabsf((float)cs[lane + 32])) & ~63u) | (unsigned)(lane + 32);
      key = k2 > key ? k2 : key;
    }
    key = __reduce_max_sync(0xffffffffu, key);
    if ((key >> 6) == 0u) { ok = false; break; }
    const int p = (int)(key & 63u);
    if (p < 32) used0 |= 1u << p; else used1 |= 1u << (p - 32);
    if (rank == 0) { piv[c] = p; stepof[p] = c; }
    const double ipiv = vl_fast_rcp(cs[p]);
    if (active) {
      const bool isp = (r == p);
      const double f = isp ? 0.0 : cs[r] * ipiv;
      const double vc = isp ? 1.0 : -f;                        // what column c becomes (the implicit unit column of row p)
      const double *pr = Mw + (size_t)p * ldm + cg;
      const int dc = c - cg;                                   // this thread holds column c iff dc == CG * jj for some jj
#pragma unroll
      for (int jj = 0; jj < NJ; jj++) {
        if (N_ > 0 || cg + CG * jj < n) {
          const double v = a[jj] - f * pr[CG * jj];            // (for column c the loaded pivot-row value is not used)
          a[jj] = (CG * jj == dc) ? vc : v;
          if (!isp) mrow[CG * jj] = a[jj];                     // (the pivot row is being read by everybody and is never a pivot row again)
          if (CG * jj == dc + 1) cn[r] = a[jj];                // column c + 1 for the next step
        }
      }
      if (isp && dc >= 0 && dc % CG == 0) dinv_s[c] = ipiv;
    }
  }
  __syncthreads();
  if (ok && active) {
    const int cr = stepof[r];
    const double di = dinv_s[cr];
#pragma unroll
    for (int jj = 0; jj < NJ; jj++) { const int j = cg + CG * jj; if (j < n) E[(size_t)cr * lde + piv[j]] = a[jj] * di; }
  }
  __syncthreads();
  return ok;
}

// Two independent n x n inverses in lock step (in place: A0 <- A0^-1, A1 <- A1^-1): the same algorithm as cta_inverse with both
// matrices advancing through ONE barrier per pivot column -- the step is a latency chain (barrier, pivot search, reciprocal, pivot-row
// loads), and two chains in flight per thread halve what K2b pays for its S+ and S- (measured in DESIGN.md section 4).
// wk0 / wk1: one inverse work area each (mirror n x ldm, dinv n, column buffer 2n; the mirrors may live anywhere), piv0 / piv1: 2n ints each.
template <int NJ, int N_, int CG_>
DEVFN bool cta_inverse2(int n_, double *A0, double *A1, int lda, double *Mw0, double *sm0, double *Mw1, double *sm1, int *piv0, int *piv1) {
  const int n = N_ > 0 ? N_ : n_, CG = CG_ > 0 ? CG_ : (int)blockDim.x / n, ldm = inv_ldm(n);
  const int rank = threadIdx.x, lane = threadIdx.x & 31;
  const bool active = rank < n * CG;
  const int r = active ? rank % n : 0, cg = active ? rank / n : 0;
  double *dinv0 = sm0, *colb0 = sm0 + n, *dinv1 = sm1, *colb1 = sm1 + n;
  int *stepof0 = piv0 + n, *stepof1 = piv1 + n;
  double a0[NJ], a1[NJ];
  double *mrow0 = Mw0 + (size_t)r * ldm + cg, *mrow1 = Mw1 + (size_t)r * ldm + cg;
#pragma unroll
  for (int jj = 0; jj < NJ; jj++) {
    const int j = cg + CG * jj;
    const bool in = active && j < n;
    a0[jj] = in ? A0[(size_t)r * lda + j] : 0.0; a1[jj] = in ? A1[(size_t)r * lda + j] : 0.0;
    if (in) { mrow0[CG * jj] = a0[jj]; mrow1[CG * jj] = a1[jj]; }
  }
  if (active && cg == 0) { colb0[r] = a0[0]; colb1[r] = a1[0]; }
  unsigned u00 = 0u, u01 = 0u, u10 = 0u, u11 = 0u;             // rows that have been pivots, per matrix (n <= 64)
  bool ok = true;
  for (int c = 0; c < n; c++) {
    __syncthreads();
    const double *cs0 = colb0 + (c & 1) * n, *cs1 = colb1 + (c & 1) * n;
    double *cn0 = colb0 + ((c + 1) & 1) * n, *cn1 = colb1 + ((c + 1) & 1) * n;
    unsigned k0 = 0u, k1 = 0u;
    if (lane < n) {
      if (!((u00 >> lane) & 1u)) k0 = (__float_as_uint(fabsf((float)cs0[lane])) & ~63u) | (unsigned)lane;
      if (!((u10 >> lane) & 1u)) k1 = (__float_as_uint(fabsf((float)cs1[lane])) & ~63u) | (unsigned)lane;
    }
    if (lane + 32 < n) {
      if (!((u01 >> lane) & 1u)) { const unsigned t = (__float_as_uint(fabsf((float)cs0[lane + 32])) & ~63u) | (unsigned)(lane + 32); k0 = t > k0 ? t : k0; }
      if (!((u11 >> lane) & 1u)) { const unsigned t = (__float_as_uint(fabsf((float)cs1[lane + 32])) & ~63u) | (unsigned)(lane + 32); k1 = t > k1 ? t : k1; }
    }
    k0 = __reduce_max_sync(0xffffffffu, k0); k1 = __reduce_max_sync(0xffffffffu, k1);
    if ((k0 >> 6) == 0u || (k1 >> 6) == 0u) { ok = false; break; }
    const int p0 = (int)(k0 & 63u), p1 = (int)(k1 & 63u);
    if (p0 < 32) u00 |= 1u << p0; else u01 |= 1u << (p0 - 32);
    if (p1 < 32) u10 |= 1u << p1; else u11 |= 1u << (p1 - 32);
    if (rank == 0) { piv0[c] = p0; stepof0[p0] = c; piv1[c] = p1; stepof1[p1] = c; }
    const double ip0 = vl_fast_rcp(cs0[p0]), ip1 = vl_fast_rcp(cs1[p1]);
    if (active) {
      const bool is0 = (r == p0), is1 = (r == p1);
      const double f0 = is0 ? 0.0 : cs0[r] * ip0, f1 = is1 ? 0.0 : cs1[r] * ip1;
      const double vc0 = is0 ? 1.0 : -f0, vc1 = is1 ? 1.0 : -f1;
      const double *pr0 = Mw0 + (size_t)p0 * ldm + cg, *pr1 = Mw1 + (size_t)p1 * ldm + cg;
      const int dc = c - cg;
#pragma unroll
      for (int jj = 0; jj < NJ; jj++) {
        if (N_ > 0 || cg + CG * jj < n) {
          const double v0 = a0[jj] - f0 * pr0[CG * jj], v1 = a1[jj] - f1 * pr1[CG * jj];
          a0[jj] = (CG * jj == dc) ? vc0 : v0; a1[jj] = (CG * jj == dc) ? vc1 : v1;
          if (!is0) mrow0[CG * jj] = a0[jj];
          if (!is1) mrow1[CG * jj] = a1[jj];
          if (CG * jj == dc + 1) { cn0[r] = a0[jj]; cn1[r] = a1[jj]; }
        }
      }
      if (dc >= 0 && dc % CG == 0) { if (is0) dinv0[c] = ip0; if (is1) dinv1[c] = ip1; }
    }
  }
  __syncthreads();
  if (ok && active) {
    const int cr0 = stepof0[r], cr1 = stepof1[r];
    const double d0 = dinv0[cr0], d1 = dinv1[cr1];
#pragma unroll
    for (int jj = 0; jj < NJ; jj++) {
      const int j = cg + CG * jj;
      if (j < n) { A0[(size_t)cr0 * lda + piv0[j]] = a0[jj] * d0; A1[(size_t)cr1 * lda + piv1[j]] = a1[jj] * d1; }
    }
  }
  __syncthreads();
  return ok;
}

// C = A B for a CTA out of shared memory with 16-byte loads: A (M x K, lda even), B (K x N, ldb even), both 16-byte aligned; every
// item owns a TM x 2 register tile, two k per step.  epi2(r, c, v0, v1) receives the finished pair (r, c), (r, c + 1) (c even; the
// second value belongs to column c + 1, which the caller ignores when c + 1 == N).
template <int TM, class Epi2>
DEVFN void cta_gemm2(int M, int N, int K, const double *AB_RESTRICT A, int lda, const double *AB_RESTRICT B, int ldb, Epi2 epi2) {
  const int tn = (N + 1) / 2, tm = (M + TM - 1) / TM;
  for (int item = threadIdx.x; item < tm * tn; item += blockDim.x) {
    const int r0 = (item / tn) * TM, c0 = (item % tn) * 2;
    double acc0[TM], acc1[TM];
#pragma unroll
    for (int i = 0; i < TM; i++) { acc0[i] = 0.0; acc1[i] = 0.0; }
    const double *ar[TM];
#pragma unroll
    for (int i = 0; i < TM; i++) ar[i] = A + (size_t)(r0 + i < M ? r0 + i : M - 1) * lda;
    const double *b = B + c0;
    int k = 0;
#pragma unroll 2
    for (; k + 1 < K; k += 2) {
      const double2 b0 = *reinterpret_cast<const double2 *>(b + (size_t)k * ldb), b1 = *reinterpret_cast<const double2 *>(b + (size_t)(k + 1) * ldb);
#pragma unroll
      for (int i = 0; i < TM; i++) {
        const double2 a2 = *reinterpret_cast<const double2 *>(ar[i] + k);
        acc0[i] += a2.x * b0.x; acc1[i] += a2.x * b0.y;
        acc0[i] += a2.y * b1.x; acc1[i] += a2.y * b1.y;
      }
    }
    if (k < K) {
      const double2 b0 = *reinterpret_cast<const double2 *>(b + (size_t)k * ldb);
#pragma unroll
      for (int i = 0; i < TM; i++) { const double av = ar[i][k]; acc0[i] += av * b0.x; acc1[i] += av * b0.y; }
    }
#pragma unroll
    for (int i = 0; i < TM; i++) if (r0 + i < M) epi2(r0 + i, c0, acc0[i], acc1[i]);
  }
}
#endif

// ---- what the solvers call: one leading dimension for every n x n team matrix (even, >= n + 2: 16-byte rows, room for the pair tiles)
HDFN int mat_ld(int n) { return ((n + 1) / 2) * 2 + 2; }
HDFN size_t inv_work_doubles(int n, bool cta) { return cta ? inv_work_doubles_cta(n) : inv_work_doubles_generic(n); }

// E = A^-1 (E may be A itself for CTA teams); piv: 2n ints
template <class Team>
DEVFN bool team_inverse(const Team &T, int n, double *A, double *E, int ld, double *wk, int *piv) {
#ifndef AB_HOST_MODEL
  if constexpr (Team::is_cta) {
    const int CG = T.size / n, nj = (n + CG - 1) / CG;
    if (n == 36 && CG == 3) return cta_inverse<12, 36, 3>(n, A, ld, E, ld, wk, piv);
    if (nj <= 6) return cta_inverse<6, 0, 0>(n, A, ld, E, ld, wk, piv);
    if (nj <= 10) return cta_inverse<10, 0, 0>(n, A, ld, E, ld, wk, piv);
    return cta_inverse<13, 0, 0>(n, A, ld, E, ld, wk, piv);
  }
#endif
  return team_inverse_generic(T, n, A, ld, E, ld, wk, piv);
}

// CTA teams: A0 <- A0^-1 and A1 <- A1^-1 together (cta_inverse2).  Mw1: n x inv_ldm(n) doubles for the second mirror, sm1: 3n doubles,
// piv: 4n ints; wk as for team_inverse.
#ifndef AB_HOST_MODEL
template <class Team>
DEVFN bool team_inverse_pair(const Team &T, int n, double *A0, double *A1, int ld, double *wk, double *Mw1, double *sm1, int *piv) {
  const int CG = T.size / n, nj = (n + CG - 1) / CG;
  double *Mw0 = wk, *sm0 = wk + (size_t)n * inv_ldm(n);
  if (n == 36 && CG == 3) return cta_inverse2<12, 36, 3>(n, A0, A1, ld, Mw0, sm0, Mw1, sm1, piv, piv + 2 * n);
  if (nj <= 6) return cta_inverse2<6, 0, 0>(n, A0, A1, ld, Mw0, sm0, Mw1, sm1, piv, piv + 2 * n);
  if (nj <= 10) return cta_inverse2<10, 0, 0>(n, A0, A1, ld, Mw0, sm0, Mw1, sm1, piv, piv + 2 * n);
  return cta_inverse2<13, 0, 0>(n, A0, A1, ld, Mw0, sm0, Mw1, sm1, piv, piv + 2 * n);
}
#endif

// C(r, c) = sum_k A(r, k) B(k, c) handed to epi(r, c, value); A, B in team scratch with the leading dimension mat_ld(n)
template <class Team, class Epi>
DEVFN void team_mm(const Team &T, int M, int N, int K, const double *A, int lda, const double *B, int ldb, Epi epi) {
#ifndef AB_HOST_MODEL
  if constexpr (Team::is_cta) {
    auto epi2 = [&](int r, int c, double v0, double v1) { epi(r, c, v0); if (c + 1 < N) epi(r, c + 1, v1); };
    if (M * ((N + 1) / 2) <= 3 * (int)blockDim.x) cta_gemm2<3>(M, N, K, A, lda, B, ldb, epi2);
    else cta_gemm2<6>(M, N, K, A, lda, B, ldb, epi2);
    return;
  }
#endif
  team_gemm_auto(T, M, N, K, A, lda, B, ldb, epi);
}

}  // namespace vlk
