// Stage B, kernel K2: one Team (a warp on the GPU) per (pixel, layer) problem of one Fourier component m.
// Computes what VLIDORT_QHOM_SOLUTION, _QHOM_NORMS, _UHOM_SOLUTION, HMULT_MASTER and _GBEAM_SOLUTION compute
// (vlidort_solutions.f90:150-1793,2649-2935), but NOT the way the reference does it:
//
//   SAB = M^-1 (I - Zs W/2),  DAB = M^-1 (I - Zd W/2),   Zs,Zd = sum_l Pi_i B_l (I +- (-1)^(l-m) D) Pi_j
// with Zs, Zd SYMMETRIC for nstokes <= 3 (B_l is block diagonal {I,Q}|{U} and commutes with D).  With the diagonal
// similarity  E = T^-1 M^-1/2 (P Q) M^1/2 T  (T = W^1/2),  P,Q symmetric and Q positive definite, so
//   Q = C C^T (Cholesky),   G = C^T P C symmetric,   G = V diag(k^2) V^T  (parallel cyclic Jacobi),
//   eigenvector of EIGENMAT:  sigma = T^-1 M^-1/2 C^-T v,      SAB sigma = T^-1 M^-1/2 C v
// i.e. the non-symmetric Hessenberg/QR eigensolver (ASYMTX) of the reference is replaced by a symmetric Jacobi
// solve that is regular, branch-free and whose eigenvectors are exactly orthogonal in the metric the Green's
// function normalisation (NORM_SAVED) assumes -- also inside degenerate clusters.
#pragma once
#include "vl_setup.cuh"

namespace vlk {

struct VLSol {                 // per problem q = p*L + k (structure of arrays)
  int n, S, N;
  double *keig, *tde;          // [q][n]
  double *xp, *xn;             // [q][e][r]   r = S*i + o, streams i < N (the i+N half follows by symmetry)
  double *uh_dn, *uh_up;       // [q][e][S]   UHOM_UPDN, UHOM_UPUP
  double *hm1, *hm2, *aterm, *bterm, *gam_m, *gam_p;   // [q][n]
  double *wupper, *wlower;     // [q][2N*S]
  int *fail;                   // [p] set !=0 when an eigenproblem fails
};

struct VLQuad {                // quadrature (pixel independent)
  int N;
  double x[32], w[32];
};

// shared/scratch size in doubles for one problem
HDFN int eigen_scratch_doubles(int n, int NM) { return 3 * n * n + 8 * n + 6 * (NM + 1) + 8; }

template <int S, class Team>
DEVFN void eigen_problem(const Team &T, const VLPix &X, const VLSol &Z, const VLQuad &QD, const double *AB_RESTRICT pq, int m,
                         int p, int k, double *scr) {
  const int N = QD.N, n = S * N, NM = X.NM, L = X.L;
  const size_t q = (size_t)p * L + k;
  double *Qm = scr, *Pm = Qm + n * n, *V = Pm + n * n;
  double *lam = V + n * n, *sr = lam + n, *dpi = sr + n, *dmi = dpi + n, *gup = dmi + n, *gdn = gup + n, *cs = gdn + n;   // cs: 2*(n/2+1) <= 2n
  double *piu = cs + 2 * n;           // 3*(NM+1)  user-stream P,R,T
  double *pi0 = piu + 3 * (NM + 1);   // 3*(NM+1)  solar P,R,T
  double *misc = pi0 + 3 * (NM + 1);
  const double *og = og_ptr(X, p);                          // [(l*L + k)*4]
  const double *sc = X.scal + (size_t)p * 16;
  const int *ib = X.ibuf + (size_t)p * 4;
  const double usec = sc[2], FLUXF = sc[5];
  const bool layscat = X.layscat[((size_t)p * (NM + 1) + m) * L + k] != 0;
  const double deltau = X.deltau[pixL(X, p) + k];

  // ---- generalized spherical functions at the user and solar cosines
  if (T.rank == 0) pi_prt(m, NM, sc[1], piu);
  if (T.rank == (T.size > 1 ? 1 : 0)) pi_prt(m, NM, sc[0], pi0);
  for (int r = T.rank; r < n; r += T.size) { const int i = r / S; sr[r] = sqrt(QD.w[i] * QD.x[i]); }
  T.sync();

  // ---- symmetric matrices P (from Zd) and Q (from Zs)
  for (int idx = T.rank; idx < N * N; idx += T.size) {
    const int i = idx / N, j = idx % N;
    double zs[S * S], zd[S * S];
#pragma unroll
    for (int t = 0; t < S * S; t++) { zs[t] = 0.0; zd[t] = 0.0; }
    for (int l = m; l <= NM; l++) {
      const double *o4 = og + ((size_t)l * L + k) * 4;
      const double *pi_ = pq + ((size_t)l * N + i) * 3, *pj_ = pq + ((size_t)l * N + j) * 3;
      const double es = ((l - m) & 1) ? 0.0 : 2.0, ed = 2.0 - es;        // I +- ps*D on the {I,Q} block; swapped on U
      if (S == 1) {
        const double t = o4[0] * pi_[0] * pj_[0];
        zs[0] += es * t; zd[0] += ed * t;
      } else {
        const double Pi = pi_[0], Ri = pi_[1], Ti = pi_[2], Pj = pj_[0], Rj = pj_[1], Tj = pj_[2];
        const double b11 = o4[0], b12 = o4[1], b22 = o4[2], b33 = o4[3];
        // Y = B * Pi_j  (3x3);  B = [[b11,b12,0],[b12,b22,0],[0,0,b33]], Pi = [[P,0,0],[0,R,T],[0,T,R]]
        const double y00 = b11 * Pj, y01 = b12 * Rj, y02 = b12 * Tj;
        const double y10 = b12 * Pj, y11 = b22 * Rj, y12 = b22 * Tj;
        const double y21 = b33 * Tj, y22 = b33 * Rj;                      // y20 = 0
        // Pi_i * E * Y with E = diag(e1,e1,e2): rows of Y scaled
        const double e1s = es, e2s = ed, e1d = ed, e2d = es;
        // row 0 of Pi_i: (Pi,0,0); row 1: (0,Ri,Ti); row 2: (0,Ti,Ri)
        const double a00 = Pi * y00, a01 = Pi * y01, a02 = Pi * y02;
        const double u10 = Ri * y10, u11 = Ri * y11, u12 = Ri * y12, v11 = Ti * y21, v12 = Ti * y22;
        const double u20 = Ti * y10, u21 = Ti * y11, u22 = Ti * y12, w21 = Ri * y21, w22 = Ri * y22;
        zs[0] += e1s * a00; zs[1] += e1s * a01; zs[2] += e1s * a02;
        zs[3] += e1s * u10; zs[4] += e1s * u11 + e2s * v11; zs[5] += e1s * u12 + e2s * v12;
        zs[6] += e1s * u20; zs[7] += e1s * u21 + e2s * w21; zs[8] += e1s * u22 + e2s * w22;
        zd[0] += e1d * a00; zd[1] += e1d * a01; zd[2] += e1d * a02;
        zd[3] += e1d * u10; zd[4] += e1d * u11 + e2d * v11; zd[5] += e1d * u12 + e2d * v12;
        zd[6] += e1d * u20; zd[7] += e1d * u21 + e2d * w21; zd[8] += e1d * u22 + e2d * w22;
      }
    }
    const double hw = 0.5 * sqrt(QD.w[i] * QD.w[j]), im = 1.0 / sqrt(QD.x[i] * QD.x[j]);
#pragma unroll
    for (int o1 = 0; o1 < S; o1++)
#pragma unroll
      for (int o2 = 0; o2 < S; o2++) {
        const int r = S * i + o1, c = S * j + o2;
        const double dlt = (r == c) ? 1.0 : 0.0;
        Qm[r * n + c] = (dlt - hw * zs[o1 * S + o2]) * im;
        Pm[r * n + c] = (dlt - hw * zd[o1 * S + o2]) * im;
      }
  }
  T.sync();
  // ---- Cholesky Q = C C^T (lower triangle of Qm)
  bool ok = true;
  for (int c = 0; c < n; c++) {
    const double d = Qm[c * n + c];
    if (!(d > 0.0)) { ok = false; break; }
    const double id = 1.0 / sqrt(d);
    T.sync();
    for (int r = c + T.rank; r < n; r += T.size) Qm[r * n + c] *= id;      // includes the diagonal: sqrt(d)
    T.sync();
    const int rem = n - c - 1;
    for (int idx = T.rank; idx < rem * rem; idx += T.size) {
      const int r = c + 1 + idx / rem, cc = c + 1 + idx % rem;
      if (cc <= r) Qm[r * n + cc] -= Qm[r * n + c] * Qm[cc * n + c];
    }
    T.sync();
  }
  if (!ok) { if (T.rank == 0) Z.fail[p] = 1; return; }
  // ---- G = C^T P C : W = P C into V, then G = C^T W into Pm
  for (int idx = T.rank; idx < n * n; idx += T.size) {
    const int r = idx / n, c = idx % n;
    double s = 0.0;
    for (int t = c; t < n; t++) s += Pm[r * n + t] * Qm[t * n + c];
    V[idx] = s;
  }
  T.sync();
  for (int idx = T.rank; idx < n * n; idx += T.size) {
    const int r = idx / n, c = idx % n;
    double s = 0.0;
    for (int t = r; t < n; t++) s += Qm[t * n + r] * V[t * n + c];
    Pm[idx] = s;
  }
  T.sync();
  for (int idx = T.rank; idx < n * n; idx += T.size) V[idx] = (idx / n == idx % n) ? 1.0 : 0.0;
  T.sync();
  // ---- parallel cyclic Jacobi on G (Pm), eigenvectors accumulated in V
  {
    const int np = n + (n & 1), half = np / 2;
    for (int sweep = 0; sweep < 20; sweep++) {
      double offmax = 0.0;
      for (int st = 0; st < np - 1; st++) {
        for (int t = T.rank; t < half; t += T.size) {
          int a, b;
          if (t == 0) { a = np - 1; b = st; } else { a = (st + t) % (np - 1); b = (st - t + np - 1) % (np - 1); }
          const int pp = a < b ? a : b, qq2 = a < b ? b : a;
          double c_ = 1.0, s_ = 0.0;
          if (qq2 < n) {
            const double gpq = Pm[pp * n + qq2], gpp = Pm[pp * n + pp], gqq = Pm[qq2 * n + qq2];
            const double rel = fabs(gpq) / sqrt(fabs(gpp * gqq) + 1.0e-300);
            if (rel > offmax) offmax = rel;
            if (rel > 1.0e-17) {
              const double th = (gqq - gpp) / (2.0 * gpq);
              const double tt = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
              c_ = 1.0 / sqrt(tt * tt + 1.0); s_ = tt * c_;
            }
          }
          cs[2 * t] = c_; cs[2 * t + 1] = s_;
        }
        T.sync();
        for (int idx = T.rank; idx < half * n; idx += T.size) {           // rows: G <- J^T G
          const int t = idx / n, c = idx % n;
          int a, b;
          if (t == 0) { a = np - 1; b = st; } else { a = (st + t) % (np - 1); b = (st - t + np - 1) % (np - 1); }
          const int pp = a < b ? a : b, qq2 = a < b ? b : a;
          if (qq2 < n) {
            const double c_ = cs[2 * t], s_ = cs[2 * t + 1];
            const double gp = Pm[pp * n + c], gq = Pm[qq2 * n + c];
            Pm[pp * n + c] = c_ * gp - s_ * gq; Pm[qq2 * n + c] = s_ * gp + c_ * gq;
          }
        }
        T.sync();
        for (int idx = T.rank; idx < half * n; idx += T.size) {           // columns: G <- G J, V <- V J
          const int t = idx / n, r = idx % n;
          int a, b;
          if (t == 0) { a = np - 1; b = st; } else { a = (st + t) % (np - 1); b = (st - t + np - 1) % (np - 1); }
          const int pp = a < b ? a : b, qq2 = a < b ? b : a;
          if (qq2 < n) {
            const double c_ = cs[2 * t], s_ = cs[2 * t + 1];
            const double gp = Pm[r * n + pp], gq = Pm[r * n + qq2];
            Pm[r * n + pp] = c_ * gp - s_ * gq; Pm[r * n + qq2] = s_ * gp + c_ * gq;
            const double vp = V[r * n + pp], vq = V[r * n + qq2];
            V[r * n + pp] = c_ * vp - s_ * vq; V[r * n + qq2] = s_ * vp + c_ * vq;
          }
        }
        T.sync();
      }
      int dummy = 0;
      offmax = T.argmax(offmax, dummy);
      if (offmax < 1.0e-15) break;
    }
  }
  for (int e = T.rank; e < n; e += T.size) lam[e] = Pm[e * n + e];
  T.sync();
  // ---- back to the eigenvectors of EIGENMAT:  y = C v (into Pm), x = C^-T v (in place in V); column e per lane
  for (int e = T.rank; e < n; e += T.size) {
    for (int r = n - 1; r >= 0; r--) {                                   // y first (needs all of v)
      double s = 0.0;
      for (int t = 0; t <= r; t++) s += Qm[r * n + t] * V[t * n + e];
      Pm[r * n + e] = s;
    }
    for (int r = n - 1; r >= 0; r--) {
      double s = V[r * n + e];
      for (int t = r + 1; t < n; t++) s -= Qm[t * n + r] * V[t * n + e];
      V[r * n + e] = s / Qm[r * n + r];
    }
  }
  T.sync();
  // ---- X+ / X- (streams 1..N), eigenvalue k, whole-layer transmittance
  bool bad = false;
  for (int e = T.rank; e < n; e += T.size) {
    const double l2 = lam[e];
    if (!(l2 > 0.0)) { bad = true; continue; }
    const double ke = sqrt(l2);
    double nrm = 0.0;
    for (int r = 0; r < n; r++) { const double s = V[r * n + e] / sr[r]; nrm += s * s; }
    const double inr = 1.0 / sqrt(nrm);
    for (int r = 0; r < n; r++) {
      const double sg = V[r * n + e] / sr[r] * inr, df = Pm[r * n + e] / (sr[r] * ke) * inr;
      Pm[r * n + e] = 0.5 * (sg + df);       // XPOS(i,o,e)
      V[r * n + e] = 0.5 * (sg - df);        // XNEG(i,o,e)
    }
    Z.keig[q * n + e] = ke;
    const double hlp = ke * deltau, tde = hlp > MAX_TAU_PATH ? 0.0 : exp(-hlp);
    Z.tde[q * n + e] = tde;
    lam[e] = ke;
    // HMULT (whole layer, upwelling user stream): vlidort_solutions.f90:1715-1745, zeta :1290-1305
    const double rho_p = usec + ke, rho_m = usec - ke, udel = X.tduserm[pixL(X, p) + k];
    const bool hs = fabs(rho_m) < TAYLOR_SMALL;
    Z.hm1[q * n + e] = hs ? taylor_series_1(rho_m, deltau, udel, usec) : usec * (tde - udel) / rho_m;
    Z.hm2[q * n + e] = usec * (1.0 - tde * udel) / rho_p;
  }
  if (bad) Z.fail[p] = 1;
  T.sync();
  double *XPg = Z.xp + q * n * n, *XNg = Z.xn + q * n * n;
  for (int idx = T.rank; idx < n * n; idx += T.size) {
    const int e = idx / n, r = idx % n;
    XPg[idx] = Pm[r * n + e]; XNg[idx] = V[r * n + e];
  }
  // ---- user-angle homogeneous solutions UHOM_UPDN / UHOM_UPUP
  for (int e = T.rank; e < n; e += T.size) {
    double dn[S], up[S];
#pragma unroll
    for (int o = 0; o < S; o++) { dn[o] = 0.0; up[o] = 0.0; }
    if (layscat) {
      for (int l = m; l <= NM; l++) {
        const double ps = ((l - m) & 1) ? -1.0 : 1.0;
        const double *o4 = og + ((size_t)l * L + k) * 4;
        double hs[S];
#pragma unroll
        for (int o = 0; o < S; o++) hs[o] = 0.0;
        for (int j = 0; j < N; j++) {
          const double *pj_ = pq + ((size_t)l * N + j) * 3;
          const double a5 = 0.5 * QD.w[j];
          if (S == 1) hs[0] += a5 * pj_[0] * (Pm[j * n + e] + ps * V[j * n + e]);
          else {
            const double Pj = pj_[0], Rj = pj_[1], Tj = pj_[2];
            const double p0 = Pm[(3 * j) * n + e], p1 = Pm[(3 * j + 1) * n + e], p2 = Pm[(3 * j + 2) * n + e];
            const double m0 = V[(3 * j) * n + e], m1 = V[(3 * j + 1) * n + e], m2 = V[(3 * j + 2) * n + e];
            // XQP*X+  +  (D Pi ps)*X-
            hs[0] += a5 * (Pj * p0 + ps * Pj * m0);
            hs[1] += a5 * (Rj * p1 + Tj * p2 + ps * (Rj * m1 + Tj * m2));
            hs[2] += a5 * (Tj * p1 + Rj * p2 - ps * (Tj * m1 + Rj * m2));
          }
        }
        const double *pu = piu + 3 * l;
        if (S == 1) {
          const double g0 = o4[0] * hs[0];
          dn[0] += ps * pu[0] * g0; up[0] += pu[0] * g0;
        } else {
          const double g0 = o4[0] * hs[0] + o4[1] * hs[1], g1 = o4[1] * hs[0] + o4[2] * hs[1], g2 = o4[3] * hs[2];
          const double Pu = pu[0], Ru = pu[1], Tu = pu[2];
          // XUM = D Pi_u D ps ;  XUP_PRE = D Pi_u
          dn[0] += ps * Pu * g0; dn[1] += ps * (Ru * g1 - Tu * g2); dn[2] += ps * (-Tu * g1 + Ru * g2);
          up[0] += Pu * g0; up[1] += Ru * g1 + Tu * g2; up[2] += -(Tu * g1 + Ru * g2);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < S; o++) { Z.uh_dn[(q * n + e) * S + o] = dn[o]; Z.uh_up[(q * n + e) * S + o] = up[o]; }
  }
  // ---- Green's function particular integral (solar beam)
  const bool active = layscat && (k + 1 <= ib[0]);
  double *WU = Z.wupper + q * (2 * n), *WL = Z.wlower + q * (2 * n);
  if (!active) {
    for (int r = T.rank; r < 2 * n; r += T.size) { WU[r] = 0.0; WL[r] = 0.0; }
    for (int e = T.rank; e < n; e += T.size) { Z.aterm[q * n + e] = 0.0; Z.bterm[q * n + e] = 0.0; Z.gam_m[q * n + e] = 1.0; Z.gam_p[q * n + e] = 1.0; }
    (void)misc;
    return;
  }
  const double F1 = FLUXF / PI4;
  for (int r = T.rank; r < n; r += T.size) {
    const int i = r / S, o1 = r % S;
    double a = 0.0, b = 0.0;
    for (int l = m; l <= NM; l++) {
      const double ps = ((l - m) & 1) ? -1.0 : 1.0;
      const double *o4 = og + ((size_t)l * L + k) * 4;
      const double *pi_ = pq + ((size_t)l * N + i) * 3;
      const double P0 = pi0[3 * l];
      double v;
      if (S == 1) v = pi_[0] * o4[0] * P0;
      else {
        const double h0 = o4[0] * P0, h1 = o4[1] * P0;                    // HELP_QFUNC(l,:) = (og11, og12, 0) * P0
        v = (o1 == 0) ? pi_[0] * h0 : (o1 == 1 ? pi_[1] * h1 : pi_[2] * h1);
      }
      a += v; b += ps * v;
    }
    dpi[r] = F1 * a; dmi[r] = F1 * b;
  }
  T.sync();
  {
    const double secbar = X.avsec[pixL(X, p) + k], cst = X.itrans[pixL(X, p) + k], wdel = X.tdmubar[pixL(X, p) + k];
    for (int e = T.rank; e < n; e += T.size) {
      double nrm = 0.0, s1 = 0.0, s2 = 0.0;
      for (int i = 0; i < N; i++) {
        double t1 = 0.0, t2 = 0.0, h1 = 0.0, h2 = 0.0;
#pragma unroll
        for (int o = 0; o < S; o++) {
          const double xp = Pm[(S * i + o) * n + e], xn = V[(S * i + o) * n + e];
          t1 += xp * xp; t2 += xn * xn;
          h1 += dpi[S * i + o] * xp + dmi[S * i + o] * xn;
          h2 += dmi[S * i + o] * xp + dpi[S * i + o] * xn;
        }
        nrm += QD.x[i] * QD.w[i] * (t1 - t2);
        s1 += QD.w[i] * h1; s2 += QD.w[i] * h2;
      }
      const double at = s1 / nrm, bt = s2 / nrm, ke = lam[e], zdel = Z.tde[q * n + e];
      const double gp = secbar + ke, gm = secbar - ke;
      const double cf = fabs(gm) < TAYLOR_SMALL ? taylor_series_1(gm, deltau, wdel, cst) : cst * (zdel - wdel) / gm;
      const double df = cst * (1.0 - zdel * wdel) / gp;
      Z.aterm[q * n + e] = at; Z.bterm[q * n + e] = bt; Z.gam_m[q * n + e] = gm; Z.gam_p[q * n + e] = gp;
      gdn[e] = cf * at; gup[e] = df * bt;
    }
  }
  T.sync();
  for (int r = T.rank; r < n; r += T.size) {
    const int o = r % S;
    double a = 0, b = 0, c = 0, d = 0;
    for (int e = 0; e < n; e++) {
      const double xp = Pm[r * n + e], xn = V[r * n + e];
      a += gup[e] * xn; b += gup[e] * xp; c += gdn[e] * xn; d += gdn[e] * xp;
    }
    const double sg = o > 1 ? -1.0 : 1.0;
    WU[r] = a; WU[n + r] = sg * b; WL[n + r] = sg * c; WL[r] = d;
  }
}

}  // namespace vlk
