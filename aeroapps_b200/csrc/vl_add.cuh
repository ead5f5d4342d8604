// Stage B, kernel K3 (default for nstokes*nstreams <= 64): the boundary-value problem of one pixel and Fourier component solved by
// the INTERACTION PRINCIPLE instead of the reference's band LU (BVP_MATRIXSETUP_MASTER / BVP_SOLUTION_MASTER,
// VM/vlidort_bvproblem.f90:147-2314), followed by the post-processing to the TOA upwelling Stokes vector
// (VM/vlidort_intensity.f90:145-646,1096-1499; VM/vlidort_PostProcessing.f90:139-454) and the Fourier convergence test
// (VM/vlidort_converge.f90:825-1231).
//
// The reference's boundary matrix couples [LCON, MCON] of neighbouring layers through 2n continuity rows per interface
// (2n x 72 unknowns, band half-width 3n-1, 5184 chained pivot columns at n = 36).  K2b (vl_layerops.cuh) has already reduced every
// layer to its n x n reflection / transmission operators R_k, Tm_k and source vectors; here the SAME linear system is solved by
// block elimination over interfaces:
//   up-sweep, k = L .. 1:   Rs_k = R_k + Tm_k Rs_{k+1} E_k Tm_k,   E_k = (1 - R_k Rs_{k+1})^-1     (reflection of everything below)
//                           Ss_k = sg_up + Tm_k (Ss_{k+1} + Rs_{k+1} y_k),   [Y_k | y_k] = E_k [Tm_k | R_k Ss_{k+1} + sg_dn]
//                           Rs_{L+1}, Ss_{L+1} = the surface (Lambertian / BRDF Fourier component, BVP_SURFACE_SETUP_HOM/_BEAM)
//   down-sweep, k = 1 .. L: I_dn(bot_k) = Y_k I_dn(top_k) + y_k,   u(bot_k) = Rs_{k+1} I_dn(bot_k) + Ss_{k+1},  I_dn(top_1) = 0
//                           a + b = S+ (in_dn + in_u),  a - b = S- (in_dn - in_u)                 (LCON, MCON of the layer)
// Work: one n x n inversion and four n x n products per layer (10 n^3 flops against the 37 n^3 of the block-bidiagonal LU), 36
// chained pivot columns per layer instead of 72, and 60 KB of shared memory per pixel instead of 125 KB, so three pixels share an SM.  The layer source terms are accumulated top-down with the running user-stream
// transmittance (the reference recurses bottom-up: same sum, different association).
#pragma once
#include "vl_bvp.cuh"
#include "vl_layerops.cuh"

namespace vlk {

// team scratch (doubles): Rs, Rk, Tk, QE (1 - R Rs, then its inverse, then Rs Y), Yb: n x ld each (ld = mat_ld(n)); 14 vectors;
// inverse work; ints: 2n
HDFN size_t add_smem_doubles(int n, bool cta) { return 5 * (size_t)n * mat_ld(n) + 14 * n + inv_work_doubles(n, cta); }
// per-team global scratch (doubles): per layer Y^T, (Rs Y)^T (n x n each), y, Rs y + Ss (n each)
HDFN size_t add_work_doubles(int n, int L) { return (size_t)L * (2 * (size_t)n * n + 2 * n); }

template <int S, class Team>
DEVFN void add_pixel(const Team &T, const VLPix &X, const VLSol &Z, const VLOps &O, const VLQuad &QD, const VLSurf &SF, const VLConv &CV,
                     int lambertian, double accuracy, int m, int p, double *scr, int *piv, double *Wg) {
  const int N = QD.N, n = S * N, L = X.L, NM = X.NM, ld = mat_ld(n);
  double *Rs = scr, *Rk = Rs + (size_t)n * ld, *Tk = Rk + (size_t)n * ld, *QE = Tk + (size_t)n * ld, *Yb = QE + (size_t)n * ld, *Zb = QE;
  double *Ss = Yb + (size_t)n * ld, *Sn = Ss + n, *su = Sn + n, *sd = su + n, *vv = sd + n, *yv = vv + n, *zz = yv + n, *xv = zz + n,
         *xb = xv + n, *ub = xb + n, *inp = ub + n, *inm = inp + n, *av = inm + n, *bv = av + n;
  double *wk = bv + n;
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
  if (!lambertian) {
    BF = SF.brdf_f + ((size_t)p * SF.nmfour + m) * SF.nq * N * N;
    BF0 = SF.brdf_f_0 + ((size_t)p * SF.nmfour + m) * SF.nq * N;
    UBF = SF.user_brdf_f + ((size_t)p * SF.nmfour + m) * SF.nq * N;
  }
  auto UBFq = [&](int o1, int o2, int jj) -> double {
    return nq == 16 ? UBF[(size_t)(4 * o1 + o2) * N + jj] : ((o1 | o2) == 0 ? UBF[jj] : 0.0);
  };
  double attn = 0.0;
  if (localbeam) attn = FLUXF * (4.0 * cos_sza / DELTA_FACTOR) / PI4 * sc[3];
  const size_t q0 = (size_t)p * L;
  const size_t wl = 2 * (size_t)n * n + 2 * n;              // scratch doubles per layer

  // ---- surface: Rs = D * (reflection matrix), Ss = D * (reflected direct beam)      (BVP_SURFACE_SETUP_HOM / _BEAM :439-626,1625-1742)
  for (int idx = T.rank; idx < n * n; idx += T.size) {
    const int r = idx / n, c = idx % n, i = r / S, o1 = r % S, j = c / S, o2 = c % S;
    double v = 0.0;
    if (include_surface) {
      const double a = QD.x[j] * QD.w[j];
      if (lambertian) v = ((o1 | o2) == 0) ? SFAC * albedo * a : 0.0;
      else v = SFAC * a * (nq == 16 ? BF[((size_t)(4 * o1 + o2) * N + i) * N + j] : ((o1 | o2) == 0 ? BF[(size_t)i * N + j] : 0.0));
      if (o1 > 1) v = -v;
    }
    Rs[(size_t)r * ld + c] = v;
  }
  for (int r = T.rank; r < n; r += T.size) {
    const int i = r / S, o1 = r % S;
    double v = 0.0;
    if (localbeam) {
      if (lambertian) v = (o1 == 0) ? attn * albedo : 0.0;
      else v = attn * (nq == 16 ? BF0[(size_t)(4 * o1) * N + i] : (o1 == 0 ? BF0[i] : 0.0));
      if (o1 > 1) v = -v;
    }
    Ss[r] = v;
  }
  T.sync();

  // ---- up-sweep
  for (int k = L - 1; k >= 0; k--) {
    const size_t q = q0 + k;
    const double *Rg = O.rk + q * n * n, *Tg = O.tk + q * n * n;
    for (int idx = T.rank; idx < n * n; idx += T.size) {
      const int r = idx / n, c = idx % n;
      Rk[(size_t)r * ld + c] = Rg[idx]; Tk[(size_t)r * ld + c] = Tg[idx];
    }
    for (int r = T.rank; r < n; r += T.size) { su[r] = O.sgu[q * n + r]; sd[r] = O.sgd[q * n + r]; }
    T.sync();
    // QE = 1 - R_k Rs,   vv = R_k Ss + sg_dn
    team_mm(T, n, n, n, Rk, ld, Rs, ld, [&](int r, int c, double acc) { QE[(size_t)r * ld + c] = (r == c ? 1.0 : 0.0) - acc; });
    for (int r = T.rank; r < n; r += T.size) {
      const double *Rr = Rk + (size_t)r * ld;
      double v = sd[r];
      for (int c = 0; c < n; c++) v += Rr[c] * Ss[c];
      vv[r] = v;
    }
    T.sync();
    if (!team_inverse(T, n, QE, QE, ld, wk, piv)) { if (T.rank == 0) { cst[0] = 0; cst[3] = 4; } return; }     // QE <- E_k
    // Yb = E_k Tm_k,  yv = E_k vv
    team_mm(T, n, n, n, QE, ld, Tk, ld, [&](int r, int c, double acc) { Yb[(size_t)r * ld + c] = acc; });
    for (int r = T.rank; r < n; r += T.size) {
      const double *Er = QE + (size_t)r * ld;
      double v = 0.0;
      for (int c = 0; c < n; c++) v += Er[c] * vv[c];
      yv[r] = v;
    }
    T.sync();
    // Y^T and y to the team's global scratch for the down-sweep;  Zb = Rs Y (overwrites E),  zz = Rs y
    double *Wk = Wg + (size_t)k * wl;
    for (int idx = T.rank; idx < n * n; idx += T.size) { const int j = idx / n, c = idx % n; Wk[idx] = Yb[(size_t)c * ld + j]; }
    team_mm(T, n, n, n, Rs, ld, Yb, ld, [&](int r, int c, double acc) { Zb[(size_t)r * ld + c] = acc; });
    for (int r = T.rank; r < n; r += T.size) {
      const double *Rr = Rs + (size_t)r * ld;
      double v = 0.0;
      for (int c = 0; c < n; c++) v += Rr[c] * yv[c];
      zz[r] = v;
      Wk[2 * (size_t)n * n + r] = yv[r];
      Wk[2 * (size_t)n * n + n + r] = v + Ss[r];               // u(bot_k) = (Rs Y) I_dn(top_k) + this
    }
    T.sync();
    for (int idx = T.rank; idx < n * n; idx += T.size) { const int j = idx / n, r = idx % n; Wk[(size_t)n * n + idx] = Zb[(size_t)r * ld + j]; }
    for (int r = T.rank; r < n; r += T.size) {
      const double *Tr = Tk + (size_t)r * ld;
      double v = su[r];
      for (int c = 0; c < n; c++) v += Tr[c] * (Ss[c] + zz[c]);
      Sn[r] = v;
    }
    // Rs <- R_k + Tm_k (Rs Y)       (nobody reads the old Rs any more)
    team_mm(T, n, n, n, Tk, ld, Zb, ld, [&](int r, int c, double acc) { Rs[(size_t)r * ld + c] = Rk[(size_t)r * ld + c] + acc; });
    T.sync();
    for (int r = T.rank; r < n; r += T.size) Ss[r] = Sn[r];
    T.sync();
  }

  // ---- down-sweep fused with the layer source terms (WHOLELAYER_STERM_UP) accumulated top-down
  double part[S];
#pragma unroll
  for (int o = 0; o < S; o++) part[o] = 0.0;
  for (int r = T.rank; r < n; r += T.size) xv[r] = 0.0;       // no diffuse downwelling at TOA
  double trans = 1.0;
  T.sync();
  for (int k = 0; k < L; k++) {
    const size_t q = q0 + k;
    const double *Wk = Wg + (size_t)k * wl;
    const double *YT = Wk, *ZT = Wk + (size_t)n * n, *yv = Wk + 2 * (size_t)n * n, *zv = yv + n;
    const bool scat = X.layscat[((size_t)p * (NM + 1) + m) * L + k] != 0;
    for (int item = T.rank; item < 2 * n; item += T.size) {
      const int which = item / n, r = item % n;
      const double *MT = which ? ZT : YT;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int c = 0;
      for (; c + 3 < n; c += 4) {
        s0 += MT[(size_t)c * n + r] * xv[c]; s1 += MT[(size_t)(c + 1) * n + r] * xv[c + 1];
        s2 += MT[(size_t)(c + 2) * n + r] * xv[c + 2]; s3 += MT[(size_t)(c + 3) * n + r] * xv[c + 3];
      }
      for (; c < n; c++) s0 += MT[(size_t)c * n + r] * xv[c];
      const double v = (s0 + s1) + (s2 + s3) + (which ? zv[r] : yv[r]);
      if (which) ub[r] = v; else xb[r] = v;
    }
    T.sync();
    if (scat) {
      const double *WU = Z.wupper + q * 2 * n, *WL = Z.wlower + q * 2 * n;
      for (int r = T.rank; r < n; r += T.size) {
        const double sg = ((r % S) > 1) ? -1.0 : 1.0;
        const double hd = xv[r] - WU[r], hu = ub[r] - sg * WL[n + r];
        inp[r] = hd + hu; inm[r] = hd - hu;
      }
      T.sync();
      const double *SpT = O.spT + q * n * n, *SmT = O.smT + q * n * n;
      for (int item = T.rank; item < 2 * n; item += T.size) {
        const int which = item / n, e = item % n;
        const double *MT = which ? SmT : SpT, *in = which ? inm : inp;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int r = 0;
        for (; r + 3 < n; r += 4) {
          s0 += MT[(size_t)r * n + e] * in[r]; s1 += MT[(size_t)(r + 1) * n + e] * in[r + 1];
          s2 += MT[(size_t)(r + 2) * n + e] * in[r + 2]; s3 += MT[(size_t)(r + 3) * n + e] * in[r + 3];
        }
        for (; r < n; r++) s0 += MT[(size_t)r * n + e] * in[r];
        const double v = (s0 + s1) + (s2 + s3);
        if (which) bv[e] = v; else av[e] = v;                  // a + b, a - b
      }
      T.sync();
      const bool beam = (k + 1 <= cutoff);
      const double wdel = X.tdmubar[pixL(X, p) + k], itr = X.itrans[pixL(X, p) + k], itw = -itr * wdel;
      const double emu = X.emult_up[pixL(X, p) + k], sgp = X.sigma_p[pixL(X, p) + k], itu = X.itrans_userm[pixL(X, p) + k];
      const double udel = X.tduserm[pixL(X, p) + k], dl = X.deltau[pixL(X, p) + k];
      for (int e = T.rank; e < n; e += T.size) {
        const double lc = 0.5 * (av[e] + bv[e]), mc = 0.5 * (av[e] - bv[e]);
        const double h1 = Z.hm1[q * n + e], h2 = Z.hm2[q * n + e];
        double cd = lc * h2, cu = mc * h1;
        if (beam) {
          const double gm = Z.gam_m[q * n + e], gp = Z.gam_p[q * n + e];
          double SD;
          if (fabs(gm) < TAYLOR_SMALL) SD = itu * taylor_series_2(gm, sgp, dl, 1.0, wdel * udel, 1.0);
          else SD = (itr * h2 - emu) / gm;
          const double SU = (itw * h1 + emu) / gp;
          cd += SD * Z.aterm[q * n + e]; cu += SU * Z.bterm[q * n + e];
        }
#pragma unroll
        for (int o = 0; o < S; o++) part[o] += trans * (cd * Z.uh_dn[(q * n + e) * S + o] + cu * Z.uh_up[(q * n + e) * S + o]);
      }
    }
    trans *= X.tduserm[pixL(X, p) + k];
    T.sync();
    for (int r = T.rank; r < n; r += T.size) xv[r] = xb[r];
    T.sync();
  }
  // ---- BOA source: surface-reflected downwelling diffuse field at the bottom (GET_BOA_SOURCE); xv = I_dn(bot_L)
  double boa[S];
#pragma unroll
  for (int o = 0; o < S; o++) boa[o] = 0.0;
  if (include_surface) {
    for (int r = T.rank; r < n; r += T.size) { const int i = r / S; inp[r] = QD.x[i] * QD.w[i] * xv[r]; }
    T.sync();
    if (lambertian) {
      double rr = 0.0;
      for (int i = 0; i < N; i++) rr += inp[S * i];
      boa[0] = SFAC * albedo * rr;
    } else {
#pragma unroll
      for (int o1 = 0; o1 < S; o1++) {
        double rr = 0.0;
        for (int jj = 0; jj < N; jj++) for (int o2 = 0; o2 < S; o2++) rr += inp[S * jj + o2] * UBFq(o1, o2, jj);
        boa[o1] = SFAC * rr;
      }
    }
  }
  double cum[S];
#pragma unroll
  for (int o = 0; o < S; o++) cum[o] = T.sum(part[o]) + trans * boa[o];
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
