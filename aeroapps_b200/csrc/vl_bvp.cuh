// Stage B, kernel K3: one Team (a CTA on the GPU) per pixel and Fourier component: boundary-value problem,
// post-processing to the TOA upwelling Stokes vector and the Fourier convergence test.
//   BVP_MATRIXSETUP_MASTER / BVP_SOLUTION_MASTER   vlidort_bvproblem.f90:147-2314
//   GET_BOA_SOURCE / VLIDORT_UPUSER_INTENSITY      vlidort_intensity.f90:145-646,1096-1499
//   WHOLELAYER_STERM_UP                            vlidort_PostProcessing.f90:139-454
//   VLIDORT_DIRECTRADIANCE                         vlidort_miscsetups.f90:1296-1617
//   VLIDORT_CONVERGE_OBSGEO                        vlidort_converge.f90:825-1231
// The reference stores the whole matrix in LAPACK band form (9*n-2 x NTOTAL) and calls DGBTRF/DGBTRS.  Here the
// block structure is used instead: the rows of layer interface k only touch the unknowns of layers k and k+1,
// so a window of 3n rows x (4n+1) columns in shared memory is eliminated layer by layer with partial pivoting
// over the window (the same pivot candidates DGBTRF sees), the right-hand side is carried as the last column
// (one beam -> the forward solve is fused into the factorisation) and only the 2n pivot rows per layer are
// spilled to HBM for the back substitution, which is fused with the upward source-function recursion.
#pragma once
#include "vl_eigen.cuh"

namespace vlk {

struct VLSurf {                  // BRDF Fourier components per pixel (null pointers for Lambertian)
  const double *brdf_f;          // [p][m][q][i][j]   q = 4*o1+o2 < nq
  const double *brdf_f_0;        // [p][m][q][i]
  const double *user_brdf_f;     // [p][m][q][j]
  const double *user_brdf_f_0;   // [p][m][q]         (not needed: the direct bounce comes from the FO code)
  int nmfour;                    // number of Fourier components stored (2N)
  int nq;                        // 1: only the (1,1) element is non-zero (scalar kernels), 16: Mueller layout
};

struct VLConv {                  // per pixel, persistent over the Fourier loop
  double *total;                 // [p][4]
  int *state;                    // [p][4]: 0 active(1)/done(0), 1 testconv, 2 fourier_saved, 3 failure code
};

HDFN int bvp_cols(int n) { return 4 * n + 1; }
HDFN int bvp_stride(int n) { return (4 * n + 1) | 1; }                         // odd row stride
HDFN size_t bvp_smem_doubles(int n) { return (size_t)3 * n * bvp_stride(n) + 8 * n + 16; }
HDFN size_t bvp_work_doubles(int n, int L) { return (size_t)L * 2 * n * bvp_cols(n); }

template <int S, class Team>
DEVFN void bvp_pixel(const Team &T, const VLPix &X, const VLSol &Z, const VLQuad &QD, const VLSurf &SF, const VLConv &CV,
                     int lambertian, double accuracy, int m, int p, double *scr, int *iscr, double *Ug) {
  const int N = QD.N, n = S * N, n2 = 2 * n, L = X.L, NM = X.NM;
  const int R = 3 * n, Cc = bvp_cols(n), Cs = bvp_stride(n), RH = 4 * n;      // RH = rhs column
  double *W = scr;
  double *xcur = W + (size_t)R * Cs;      // 2n
  double *xnext = xcur + n2;              // 2n
  double *red = xnext + n2;               // 4n scratch
  int *state = iscr;                      // R : -1 unused, 0 active, 1+t pivot number t
  int *pivrow = state + R;                // 2n
  int *freeslot = pivrow + n2;            // 2n
  const double *sc = X.scal + (size_t)p * 16;
  const int *ib = X.ibuf + (size_t)p * 4;
  int *cst = CV.state + (size_t)p * 4;
  const double cos_sza = sc[0], albedo = sc[4], FLUXF = sc[5];
  const int cutoff = ib[0];
  const bool refl_direct = ib[2] != 0;
  bool include_surface = true;
  if (lambertian) { if (m != 0) include_surface = false; else if (albedo == 0.0) include_surface = false; }
  const double SFAC = (m == 0) ? 2.0 : 1.0, DELTA_FACTOR = (m == 0) ? 1.0 : 2.0;
  const bool localbeam = include_surface && refl_direct;
  const double *BF = nullptr, *BF0 = nullptr, *UBF = nullptr;
  const int nq = SF.nq;
  auto BFq = [&](int o1, int o2, int i, int jj) -> double {
    return nq == 16 ? BF[((size_t)(4 * o1 + o2) * N + i) * N + jj] : ((o1 | o2) == 0 ? BF[(size_t)i * N + jj] : 0.0);
  };
  auto UBFq = [&](int o1, int o2, int jj) -> double {
    return nq == 16 ? UBF[(size_t)(4 * o1 + o2) * N + jj] : ((o1 | o2) == 0 ? UBF[jj] : 0.0);
  };
  if (!lambertian) {
    BF = SF.brdf_f + ((size_t)p * SF.nmfour + m) * SF.nq * N * N;
    BF0 = SF.brdf_f_0 + ((size_t)p * SF.nmfour + m) * SF.nq * N;
    UBF = SF.user_brdf_f + ((size_t)p * SF.nmfour + m) * SF.nq * N;
  }
  double attn = 0.0;
  if (localbeam) attn = FLUXF * (4.0 * cos_sza / DELTA_FACTOR) / PI4 * sc[3];
  const size_t q0 = (size_t)p * L;

  // full-stream accessors: XPOS(i2,o,e) / XNEG(i2,o,e) of layer q from the stored downwelling halves
  auto xpf = [&](size_t q, int e, int r2) -> double {   // r2 = S*i2 + o, i2 in [0,2N)
    if (r2 < n) return Z.xp[(q * n + e) * n + r2];
    const int r = r2 - n; const double v = Z.xn[(q * n + e) * n + r];
    return ((r % S) > 1) ? -v : v;
  };
  auto xnf = [&](size_t q, int e, int r2) -> double {
    if (r2 < n) return Z.xn[(q * n + e) * n + r2];
    const int r = r2 - n; const double v = Z.xp[(q * n + e) * n + r];
    return ((r % S) > 1) ? -v : v;
  };

  for (int r = T.rank; r < R; r += T.size) state[r] = -1;
  for (int idx = T.rank; idx < R * Cs; idx += T.size) W[idx] = 0.0;
  T.sync();
  // TOA rows -> slots 0..n-1 ; first free slots n..3n-1
  for (int idx = T.rank; idx < n * n; idx += T.size) {
    const int e = idx / n, r = idx % n;
    W[(size_t)r * Cs + e] = Z.xp[(q0 * n + e) * n + r];
    W[(size_t)r * Cs + n + e] = Z.xn[(q0 * n + e) * n + r] * Z.tde[q0 * n + e];
  }
  for (int r = T.rank; r < n; r += T.size) { W[(size_t)r * Cs + RH] = -Z.wupper[q0 * n2 + r]; state[r] = 0; }
  for (int t = T.rank; t < n2; t += T.size) freeslot[t] = n + t;
  T.sync();

  bool failed = false;
  for (int k = 0; k < L; k++) {
    const size_t qa = q0 + k, qb = q0 + k + 1;
    const bool last = (k == L - 1);
    // ---- load the new rows into the free slots
    if (!last) {
      for (int idx = T.rank; idx < n * n2; idx += T.size) {
        const int e = idx / n2, r2 = idx % n2;
        double *row = W + (size_t)freeslot[r2] * Cs;
        row[e] = Z.tde[qa * n + e] * xpf(qa, e, r2);
        row[n + e] = xnf(qa, e, r2);
        row[n2 + e] = -xpf(qb, e, r2);
        row[n2 + n + e] = -Z.tde[qb * n + e] * xnf(qb, e, r2);
      }
      for (int r2 = T.rank; r2 < n2; r2 += T.size) {
        W[(size_t)freeslot[r2] * Cs + RH] = Z.wupper[qb * n2 + r2] - Z.wlower[qa * n2 + r2];
        state[freeslot[r2]] = 0;
      }
    } else {
      // surface rows (upwelling streams i+N): X(i+N) - R2_HOM ; rhs = -WLOWER(i+N) + R2_BEAM + direct beam
      for (int idx = T.rank; idx < n * n; idx += T.size) {
        const int e = idx / n, r = idx % n, i = r / S, o1 = r % S;
        double rp = 0.0, rm = 0.0;
        if (include_surface) {
          if (lambertian) {
            if (o1 == 0) {
              for (int jj = 0; jj < N; jj++) { const double a = QD.x[jj] * QD.w[jj]; rp += a * Z.xp[(qa * n + e) * n + S * jj]; rm += a * Z.xn[(qa * n + e) * n + S * jj]; }
              rp *= SFAC * albedo; rm *= SFAC * albedo;
            }
          } else {
            for (int jj = 0; jj < N; jj++)
              for (int o2 = 0; o2 < S; o2++) {
                const double a = QD.x[jj] * QD.w[jj] * BFq(o1, o2, i, jj);
                rp += a * Z.xp[(qa * n + e) * n + S * jj + o2]; rm += a * Z.xn[(qa * n + e) * n + S * jj + o2];
              }
            rp *= SFAC; rm *= SFAC;
          }
        }
        double *row = W + (size_t)freeslot[r] * Cs;
        row[e] = Z.tde[qa * n + e] * (xpf(qa, e, n + r) - rp);
        row[n + e] = xnf(qa, e, n + r) - rm;
      }
      for (int r = T.rank; r < n; r += T.size) {
        const int i = r / S, o1 = r % S;
        double v = -Z.wlower[qa * n2 + n + r];
        if (include_surface) {
          double rb = 0.0;
          if (lambertian) {
            if (o1 == 0) { for (int jj = 0; jj < N; jj++) rb += QD.x[jj] * QD.w[jj] * Z.wlower[qa * n2 + S * jj]; rb *= SFAC * albedo; }
          } else {
            for (int jj = 0; jj < N; jj++) for (int o2 = 0; o2 < S; o2++) rb += QD.x[jj] * QD.w[jj] * BFq(o1, o2, i, jj) * Z.wlower[qa * n2 + S * jj + o2];
            rb *= SFAC;
          }
          v += rb;
          if (localbeam) {
            if (lambertian) { if (m == 0 && o1 == 0) v += attn * albedo; }
            else v += attn * (nq == 16 ? BF0[(size_t)(4 * o1) * N + i] : (o1 == 0 ? BF0[i] : 0.0));
          }
        }
        W[(size_t)freeslot[r] * Cs + RH] = v;
        state[freeslot[r]] = 0;
      }
    }
    T.sync();
    // ---- eliminate the 2n columns of layer k (partial pivoting over the active rows of the window)
    const int nwarp = (T.size + 31) / 32, warp = T.rank / 32, lane = T.rank % 32;
    for (int jc = 0; jc < n2; jc++) {
      double v = -1.0; int idx = R;
      for (int r = T.rank; r < R; r += T.size)
        if (state[r] == 0) { const double a = fabs(W[(size_t)r * Cs + jc]); if (a > v) { v = a; idx = r; } }
      v = T.argmax(v, idx);
      if (!(v > 0.0)) { failed = true; break; }
      if (T.rank == 0) { state[idx] = 1 + jc; pivrow[jc] = idx; }
      T.sync();
      const double *prow = W + (size_t)idx * Cs;
      const double ipiv = 1.0 / prow[jc];
      if (T.size == 1) {
        for (int r = 0; r < R; r++) if (state[r] == 0) {
          double *row = W + (size_t)r * Cs; const double f = row[jc] * ipiv;
          if (f != 0.0) for (int c = jc + 1; c < Cc; c++) row[c] -= f * prow[c];
        }
      } else {
        double pv[8];
        const int c0 = jc + 1 + lane;
#pragma unroll
        for (int u = 0; u < 8; u++) { const int c = c0 + 32 * u; pv[u] = c < Cc ? prow[c] : 0.0; }
        for (int r = warp; r < R; r += nwarp) if (state[r] == 0) {
          double *row = W + (size_t)r * Cs; const double f = row[jc] * ipiv;
#pragma unroll
          for (int u = 0; u < 8; u++) { const int c = c0 + 32 * u; if (c < Cc) row[c] -= f * pv[u]; }
        }
      }
      T.sync();
    }
    if (failed) break;
    // ---- spill the pivot rows (U factor + transformed rhs), recycle their slots, shift the carry rows
    double *Uk = Ug + (size_t)k * n2 * Cc;
    for (int idx = T.rank; idx < n2 * Cc; idx += T.size) {
      const int t = idx / Cc, c = idx % Cc;
      Uk[idx] = W[(size_t)pivrow[t] * Cs + c];
    }
    T.sync();
    for (int t = T.rank; t < n2; t += T.size) { freeslot[t] = pivrow[t]; state[pivrow[t]] = -1; }
    for (int idx = T.rank; idx < R * n2; idx += T.size) {
      const int r = idx / n2, c = idx % n2;
      if (state[r] == 0) { double *row = W + (size_t)r * Cs; row[c] = row[n2 + c]; }
    }
    T.sync();
    for (int idx = T.rank; idx < R * n2; idx += T.size) {
      const int r = idx / n2, c = idx % n2;
      W[(size_t)r * Cs + n2 + c] = 0.0;
      if (state[r] != 0) W[(size_t)r * Cs + c] = 0.0;
    }
    T.sync();
  }
  if (failed) { if (T.rank == 0) { cst[0] = 0; cst[3] = 4; } return; }

  // ---- back substitution fused with the upwelling source-function recursion
  const double usec = sc[2];
  (void)usec;
  double cum[S];
#pragma unroll
  for (int o = 0; o < S; o++) cum[o] = 0.0;
  for (int t = T.rank; t < n2; t += T.size) xnext[t] = 0.0;
  T.sync();
  double *Ub = W;                                         // 2n x (2n+1) block copy: U_kk | b
  const int Ubs = n2 + 1;
  for (int k = L - 1; k >= 0; k--) {
    const size_t qa = q0 + k;
    const double *Uk = Ug + (size_t)k * n2 * Cc;
    for (int idx = T.rank; idx < n2 * n2; idx += T.size) { const int t = idx / n2, c = idx % n2; Ub[(size_t)t * Ubs + c] = Uk[(size_t)t * Cc + c]; }
    for (int t = T.rank; t < n2; t += T.size) {
      double b = Uk[(size_t)t * Cc + RH];
      const double *un = Uk + (size_t)t * Cc + n2;
      for (int c = 0; c < n2; c++) b -= un[c] * xnext[c];
      Ub[(size_t)t * Ubs + n2] = b;
    }
    T.sync();
    // upper-triangular solve (row t has its pivot in column t), column oriented
    for (int t = n2 - 1; t >= 0; t--) {
      const double xt = Ub[(size_t)t * Ubs + n2] / Ub[(size_t)t * Ubs + t];
      T.sync();
      if (T.rank == 0) xcur[t] = xt;
      for (int r = T.rank; r < t; r += T.size) Ub[(size_t)r * Ubs + n2] -= Ub[(size_t)r * Ubs + t] * xt;
      T.sync();
    }
    // xcur = [LCON(1..n), MCON(1..n)] of layer k
    if (k == L - 1) {
      // BOA source: surface-reflected diffuse field (GET_BOA_SOURCE)
      if (include_surface) {
        for (int r = T.rank; r < n; r += T.size) {
          double sh = 0.0;
          for (int e = 0; e < n; e++) sh += xcur[e] * Z.xp[(qa * n + e) * n + r] * Z.tde[qa * n + e] + xcur[n + e] * Z.xn[(qa * n + e) * n + r];
          const int i = r / S;
          red[r] = QD.x[i] * QD.w[i] * (Z.wlower[qa * n2 + r] + sh);
        }
        T.sync();
        if (lambertian) {
          if (m == 0) { double rr = 0.0; for (int i = 0; i < N; i++) rr += red[S * i]; cum[0] = SFAC * albedo * rr; }
        } else {
#pragma unroll
          for (int o1 = 0; o1 < S; o1++) {
            double rr = 0.0;
            for (int jj = 0; jj < N; jj++) for (int o2 = 0; o2 < S; o2++) rr += red[S * jj + o2] * UBFq(o1, o2, jj);
            cum[o1] = SFAC * rr;
          }
        }
        T.sync();
      }
    }
    // layer source term
    double src[S];
#pragma unroll
    for (int o = 0; o < S; o++) src[o] = 0.0;
    if (X.layscat[((size_t)p * (NM + 1) + m) * L + k]) {
      const bool beam = (k + 1 <= cutoff);
      const double wdel = X.tdmubar[pixL(X, p) + k], itr = X.itrans[pixL(X, p) + k], itw = -itr * wdel;
      const double emu = X.emult_up[pixL(X, p) + k], sgp = X.sigma_p[pixL(X, p) + k], itu = X.itrans_userm[pixL(X, p) + k];
      const double udel = X.tduserm[pixL(X, p) + k], dl = X.deltau[pixL(X, p) + k];
      double part[S];
#pragma unroll
      for (int o = 0; o < S; o++) part[o] = 0.0;
      for (int e = T.rank; e < n; e += T.size) {
        const double h1 = Z.hm1[qa * n + e], h2 = Z.hm2[qa * n + e];
        double cd = xcur[e] * h2, cu = xcur[n + e] * h1;
        if (beam) {
          const double gm = Z.gam_m[qa * n + e], gp = Z.gam_p[qa * n + e];
          double SD;
          if (fabs(gm) < TAYLOR_SMALL) SD = itu * taylor_series_2(gm, sgp, dl, 1.0, wdel * udel, 1.0);
          else SD = (itr * h2 - emu) / gm;
          const double SU = (itw * h1 + emu) / gp;
          cd += SD * Z.aterm[qa * n + e]; cu += SU * Z.bterm[qa * n + e];
        }
#pragma unroll
        for (int o = 0; o < S; o++) part[o] += cd * Z.uh_dn[(qa * n + e) * S + o] + cu * Z.uh_up[(qa * n + e) * S + o];
      }
#pragma unroll
      for (int o = 0; o < S; o++) src[o] = T.sum(part[o]);
    }
    {
      const double udel = X.tduserm[pixL(X, p) + k];
#pragma unroll
      for (int o = 0; o < S; o++) cum[o] = src[o] + udel * cum[o];
    }
    T.sync();
    for (int t = T.rank; t < n2; t += T.size) xnext[t] = xcur[t];
    T.sync();
  }
  // ---- Fourier summation and convergence (VLIDORT_CONVERGE_OBSGEO); FLUX_MULTIPLIER = DELTA_FACTOR
  if (T.rank == 0) {
    double *tot = CV.total + (size_t)p * 4;
    double sf[S];
#pragma unroll
    for (int o = 0; o < S; o++) sf[o] = DELTA_FACTOR * cum[o];
    if (m == 0) {
#pragma unroll
      for (int o = 0; o < S; o++) tot[o] = sf[o] + sc[6 + o] + sc[9 + o];
      cst[1] = 0; cst[2] = 0;
      if (NM == 0) cst[0] = 0;
    } else {
      const double arg = sc[12] * (double)m * DEG_TO_RAD, cf = cos(arg), sf_ = sin(arg);
      const double t0 = cf * sf[0], tn0 = tot[0] + t0;
      int count = 0;
      if (t0 != 0.0) { if (fabs(t0 / tn0) < accuracy) count = 1; } else count = 1;
      tot[0] = tn0;
      if (S > 1) { tot[1] += cf * sf[1]; tot[2] += sf_ * sf[2]; }
      if (count == 1) {
        cst[1] += 1;
        if (cst[1] == 2) { cst[0] = 0; cst[2] = m; }
      } else { cst[1] = 0; cst[2] = 2 * N - 1; }
      if (m >= NM) cst[0] = 0;
    }
  }
}

}  // namespace vlk
