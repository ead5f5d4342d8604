/* TEST INFRASTRUCTURE ONLY -- numerical helpers of the CPU oracle for Stage B (VLIDORT 2.8.3).
 * Restatements of (paths relative to /root/reference/src/Components/rtms/RTSI/VLIDORT2p8p3/):
 *   vlidort_main/regular/vlidort_aux.f90:960-1044   GETQUAD2
 *   vlidort_main/regular/vlidort_aux.f90:386-958    ASYMTX (DISORT real non-symmetric eigensolver)
 *   vlidort_main/regular/vlidort_Taylor.f90:119-270 TAYLOR_SERIES_1 / _2
 *   vlidort_main/regular/lapack_tools.f90           DGBTRF/DGBTRS semantics (band LU, partial pivoting)
 */
#pragma once
#include <cmath>
#include <vector>
#include <algorithm>
#include <cstring>

namespace vl {

static const double DEG_TO_RAD = 1.7453292519943e-02;   /* vlidort_pars.f90:397 (truncated on purpose) */
static const double PIE = 180.0 * DEG_TO_RAD;
static const double PI2 = 2.0 * PIE, PI4 = 4.0 * PIE;
static const double TAYLOR_SMALL = 1.0e-4, OMEGA_SMALLNUM = 1.0e-15;
static const double MAX_TAU_SPATH = 32.0, MAX_TAU_UPATH = 32.0, MAX_TAU_QPATH = 32.0;

/* GETQUAD2: Gauss-Legendre on (A,B), roots ascending */
inline void getquad2(double A, double B, int N, double *roots, double *wgts) {
  const double QEPS = 1.0e-13;
  int N2 = (N + 1) / 2;
  double NR = (double)N, MIDPT = 0.5 * (B + A), SFAC = 0.5 * (B - A);
  for (int M = 1; M <= N2; M++) {
    double MR = (double)M;
    double XX = PIE * (MR - 0.25) / (NR + 0.5);
    double s = std::sin(XX);
    double X = (1.0 - (NR - 1.0) / (8.0 * NR * NR * NR) - 1.0 / (384.0 * NR * NR * NR * NR) * (39.0 - 28.0 / (s * s))) * std::cos(XX);
    double LP = 1.0, LPM1 = 0.0, DLP_DX = 0.0;
    for (;;) {
      LPM1 = 0.0; LP = 1.0;
      for (int I = 1; I <= N; I++) {
        double IR = (double)I, LPM2 = LPM1;
        LPM1 = LP;
        LP = ((2.0 * IR - 1.0) * X * LPM1 - (IR - 1.0) * LPM2) / IR;
      }
      DLP_DX = NR * (X * LP - LPM1) / (X * X - 1.0);
      double XOLD = X;
      X = XOLD - LP / DLP_DX;
      if (std::fabs(X - XOLD) <= QEPS) break;
    }
    roots[M - 1] = MIDPT - SFAC * X;
    wgts[M - 1] = (2.0 * SFAC) / ((1.0 - X * X) * DLP_DX * DLP_DX);
    int NM1 = N - M + 1;
    roots[NM1 - 1] = MIDPT + SFAC * X;
    wgts[NM1 - 1] = wgts[M - 1];
  }
}

inline double taylor_series_1(int order, double eps, double delta, double udel, double sm) {
  int mterms = order + 1;
  double d[16];
  d[0] = 1.0;
  for (int m = 1; m <= mterms; m++) d[m] = delta * d[m - 1] / (double)m;
  double mult = d[1], power = 1.0;
  for (int m = 2; m <= mterms; m++) { power *= eps; mult += power * d[m]; }
  return mult * udel * sm;
}

inline double taylor_series_2(int order, double eps, double y, double delta, double fac1, double fac2, double sm) {
  int mterms = order + 1;
  if (std::fabs(y) < TAYLOR_SMALL) mterms++;
  double d[16], ac[16], cc[16], term1[16];
  d[0] = 1.0;
  for (int m = 1; m <= mterms; m++) d[m] = delta * d[m - 1] / (double)m;
  if (std::fabs(y) < TAYLOR_SMALL) {
    double power = 1.0, power2 = 1.0, mult = d[2];
    for (int m = 3; m <= mterms; m++) {
      power = power * (eps - y);
      power2 = power - y * power2;
      mult += d[m] * power2;
    }
    return mult * fac1 * sm;
  }
  double y1 = 1.0 / y;
  ac[0] = 1.0;
  for (int m = 1; m <= mterms; m++) ac[m] = y1 * ac[m - 1];
  cc[0] = 1.0;
  for (int m = 1; m <= mterms; m++) { cc[m] = 0.0; for (int j = 0; j <= m; j++) cc[m] += ac[j] * d[m - j]; }
  for (int m = 0; m <= mterms; m++) term1[m] = fac1 * ac[m] - fac2 * cc[m];
  double power = 1.0, mult = term1[1];
  for (int m = 2; m <= mterms; m++) { power = eps * power; mult += power * term1[m]; }
  return mult * sm * y1;
}

/* ASYMTX.  a: M x M column-major (lda = M), destroyed; evec (M x M col-major), eval(M).
 * returns IER (0 ok, >0 eigenvalue index not converged); *bad set on complex 2x2 / bad input. */
inline int asymtx(double TOL, double *a, int M, double *evec, double *eval, bool *bad) {
  const double C1 = 0.4375, C2 = 0.5, C3 = 0.75, C4 = 0.95, C5 = 16.0, C6 = 256.0;
  *bad = false;
  int IER = 0;
  std::vector<double> wkv(4 * M + 4, 0.0);
  double *wk = wkv.data();
#define A_(i, j) a[((i) - 1) + (size_t)((j) - 1) * M]
#define E_(i, j) evec[((i) - 1) + (size_t)((j) - 1) * M]
#define W_(i) wk[(i)]
  if (M == 1) { eval[0] = A_(1, 1); E_(1, 1) = 1.0; return 0; }
  if (M == 2) {
    double d = (A_(1, 1) - A_(2, 2));
    double DISCRI = d * d + 4.0 * A_(1, 2) * A_(2, 1);
    if (DISCRI < 0.0) { *bad = true; return 0; }
    double SGN = 1.0;
    if (A_(1, 1) < A_(2, 2)) SGN = -1.0;
    eval[0] = 0.5 * (A_(1, 1) + A_(2, 2) + SGN * std::sqrt(DISCRI));
    eval[1] = 0.5 * (A_(1, 1) + A_(2, 2) - SGN * std::sqrt(DISCRI));
    E_(1, 1) = 1.0; E_(2, 2) = 1.0;
    if (A_(1, 1) == A_(2, 2) && (A_(2, 1) == 0.0 || A_(1, 2) == 0.0)) {
      double RNORM = std::fabs(A_(1, 1)) + std::fabs(A_(1, 2)) + std::fabs(A_(2, 1)) + std::fabs(A_(2, 2));
      double W = TOL * RNORM;
      E_(2, 1) = A_(2, 1) / W; E_(1, 2) = -A_(1, 2) / W;
    } else {
      E_(2, 1) = A_(2, 1) / (eval[0] - A_(2, 2));
      E_(1, 2) = A_(1, 2) / (eval[1] - A_(1, 1));
    }
    return 0;
  }
  for (int I = 1; I <= M; I++) { eval[I - 1] = 0.0; for (int J = 1; J <= M; J++) E_(I, J) = 0.0; E_(I, I) = 1.0; }
  double RNORM = 0.0;
  int L = 1, K = M;
  /* rows isolating an eigenvalue, pushed down */
  for (bool again = true; again;) {
    again = false;
    for (int J = K; J >= 1; J--) {
      double ROW = 0.0;
      for (int I = 1; I <= K; I++) if (I != J) ROW += std::fabs(A_(J, I));
      if (ROW == 0.0) {
        W_(K) = J;
        if (J != K) {
          for (int I = 1; I <= K; I++) std::swap(A_(I, J), A_(I, K));
          for (int I = L; I <= M; I++) std::swap(A_(J, I), A_(K, I));
        }
        K--; again = true; break;
      }
    }
  }
  /* columns isolating an eigenvalue, pushed left */
  for (bool again = true; again;) {
    again = false;
    for (int J = L; J <= K; J++) {
      double COL = 0.0;
      for (int I = L; I <= K; I++) if (I != J) COL += std::fabs(A_(I, J));
      if (COL == 0.0) {
        W_(L) = J;
        if (J != L) {
          for (int I = 1; I <= K; I++) std::swap(A_(I, J), A_(I, L));
          for (int I = L; I <= M; I++) std::swap(A_(J, I), A_(L, I));
        }
        L++; again = true; break;
      }
    }
  }
  for (int I = L; I <= K; I++) W_(I) = 1.0;
  /* balance */
  for (bool NOCONV = true; NOCONV;) {
    NOCONV = false;
    for (int I = L; I <= K; I++) {
      double COL = 0.0, ROW = 0.0;
      for (int J = L; J <= K; J++) if (J != I) { COL += std::fabs(A_(J, I)); ROW += std::fabs(A_(I, J)); }
      double F = 1.0, G = ROW / C5, H = COL + ROW;
      while (COL < G) { F *= C5; COL *= C6; }
      G = ROW * C5;
      while (COL >= G) { F /= C5; COL /= C6; }
      if ((COL + ROW) / F < C4 * H) {
        W_(I) = W_(I) * F; NOCONV = true;
        for (int J = L; J <= M; J++) A_(I, J) /= F;
        for (int J = 1; J <= K; J++) A_(J, I) *= F;
      }
    }
  }
  if (!(K - 1 < L + 1)) {
    /* Householder reduction to upper Hessenberg */
    for (int N = L + 1; N <= K - 1; N++) {
      double H = 0.0, SCALE = 0.0;
      W_(N + M) = 0.0;
      for (int I = N; I <= K; I++) SCALE += std::fabs(A_(I, N - 1));
      if (SCALE != 0.0) {
        for (int I = K; I >= N; I--) { W_(I + M) = A_(I, N - 1) / SCALE; H += W_(I + M) * W_(I + M); }
        double G = -std::copysign(std::sqrt(H), W_(N + M));
        H = H - W_(N + M) * G;
        W_(N + M) = W_(N + M) - G;
        for (int J = N; J <= M; J++) {
          double F = 0.0;
          for (int I = K; I >= N; I--) F += W_(I + M) * A_(I, J);
          for (int I = N; I <= K; I++) A_(I, J) -= W_(I + M) * F / H;
        }
        for (int I = 1; I <= K; I++) {
          double F = 0.0;
          for (int J = K; J >= N; J--) F += W_(J + M) * A_(I, J);
          for (int J = N; J <= K; J++) A_(I, J) -= W_(J + M) * F / H;
        }
        W_(N + M) = SCALE * W_(N + M);
        A_(N, N - 1) = SCALE * G;
      }
    }
    for (int N = K - 2; N >= L; N--) {
      int N1 = N + 1, N2 = N + 2;
      double F = A_(N1, N);
      if (F != 0.0) {
        F = F * W_(N1 + M);
        for (int I = N2; I <= K; I++) W_(I + M) = A_(I, N);
        if (N1 <= K) {
          for (int J = 1; J <= M; J++) {
            double G = 0.0;
            for (int I = N1; I <= K; I++) G += W_(I + M) * E_(I, J);
            G = G / F;
            for (int I = N1; I <= K; I++) E_(I, J) += G * W_(I + M);
          }
        }
      }
    }
  }
  /* 350 */
  {
    int N = 1;
    for (int I = 1; I <= M; I++) {
      for (int J = N; J <= M; J++) RNORM += std::fabs(A_(I, J));
      N = I;
      if (I < L || I > K) eval[I - 1] = A_(I, I);
    }
  }
  {
    int N = K;
    double T = 0.0;
    double P = 0, Q = 0, R = 0, S = 0, X = 0, Y = 0, Z = 0, W = 0;
    while (N >= L) {                       /* 380 */
      int IN = 0, N1 = N - 1, N2 = N - 2;
      bool next_root = false;
      while (!next_root) {                 /* 390 */
        int LB = L;
        for (int I = L; I <= N; I++) {
          LB = N + L - I;
          if (LB == L) break;
          S = std::fabs(A_(LB - 1, LB - 1)) + std::fabs(A_(LB, LB));
          if (S == 0.0) S = RNORM;
          if (std::fabs(A_(LB, LB - 1)) <= TOL * S) break;
        }
        X = A_(N, N);
        if (LB == N) {
          A_(N, N) = X + T; eval[N - 1] = A_(N, N);
          N = N1; next_root = true; continue;
        }
        Y = A_(N1, N1);
        W = A_(N, N1) * A_(N1, N);
        if (LB == N1) {
          P = (Y - X) * C2;
          Q = P * P + W;
          Z = std::sqrt(std::fabs(Q));
          A_(N, N) = X + T; X = A_(N, N);
          A_(N1, N1) = Y + T;
          Z = P + std::copysign(Z, P);
          eval[N1 - 1] = X + Z;
          eval[N - 1] = eval[N1 - 1];
          if (Z != 0.0) eval[N - 1] = X - W / Z;
          X = A_(N, N1);
          R = std::sqrt(X * X + Z * Z);
          P = X / R; Q = Z / R;
          for (int J = N1; J <= M; J++) { Z = A_(N1, J); A_(N1, J) = Q * Z + P * A_(N, J); A_(N, J) = Q * A_(N, J) - P * Z; }
          for (int I = 1; I <= N; I++) { Z = A_(I, N1); A_(I, N1) = Q * Z + P * A_(I, N); A_(I, N) = Q * A_(I, N) - P * Z; }
          for (int I = L; I <= K; I++) { Z = E_(I, N1); E_(I, N1) = Q * Z + P * E_(I, N); E_(I, N) = Q * E_(I, N) - P * Z; }
          N = N2; next_root = true; continue;
        }
        if (IN == 30) { IER = N; goto done; }
        if (IN == 10 || IN == 20) {
          T = T + X;
          for (int I = L; I <= N; I++) A_(I, I) -= X;
          S = std::fabs(A_(N, N1)) + std::fabs(A_(N1, N2));
          X = C3 * S; Y = X; W = -C1 * S * S;
        }
        IN++;
        int I = LB;
        for (int J = LB; J <= N2; J++) {
          I = N2 + LB - J;
          Z = A_(I, I);
          R = X - Z; S = Y - Z;
          P = (R * S - W) / A_(I + 1, I) + A_(I, I + 1);
          Q = A_(I + 1, I + 1) - Z - R - S;
          R = A_(I + 2, I + 1);
          S = std::fabs(P) + std::fabs(Q) + std::fabs(R);
          P /= S; Q /= S; R /= S;
          if (I == LB) break;
          double UU = std::fabs(A_(I, I - 1)) * (std::fabs(Q) + std::fabs(R));
          double VV = std::fabs(P) * (std::fabs(A_(I - 1, I - 1)) + std::fabs(Z) + std::fabs(A_(I + 1, I + 1)));
          if (UU <= TOL * VV) break;
        }
        A_(I + 2, I) = 0.0;
        for (int J = I + 3; J <= N; J++) { A_(J, J - 2) = 0.0; A_(J, J - 3) = 0.0; }
        for (int KA = I; KA <= N1; KA++) {
          bool NOTLAS = (KA != N1);
          if (KA == I) {
            S = std::copysign(std::sqrt(P * P + Q * Q + R * R), P);
            if (LB != I) A_(KA, KA - 1) = -A_(KA, KA - 1);
          } else {
            P = A_(KA, KA - 1);
            Q = A_(KA + 1, KA - 1);
            R = 0.0;
            if (NOTLAS) R = A_(KA + 2, KA - 1);
            X = std::fabs(P) + std::fabs(Q) + std::fabs(R);
            if (X == 0.0) continue;
            P /= X; Q /= X; R /= X;
            S = std::copysign(std::sqrt(P * P + Q * Q + R * R), P);
            A_(KA, KA - 1) = -S * X;
          }
          P = P + S; X = P / S; Y = Q / S; Z = R / S; Q = Q / P; R = R / P;
          for (int J = KA; J <= M; J++) {
            P = A_(KA, J) + Q * A_(KA + 1, J);
            if (NOTLAS) { P += R * A_(KA + 2, J); A_(KA + 2, J) -= P * Z; }
            A_(KA + 1, J) -= P * Y;
            A_(KA, J) -= P * X;
          }
          for (int II = 1; II <= std::min(N, KA + 3); II++) {
            P = X * A_(II, KA) + Y * A_(II, KA + 1);
            if (NOTLAS) { P += Z * A_(II, KA + 2); A_(II, KA + 2) -= P * R; }
            A_(II, KA + 1) -= P * Q;
            A_(II, KA) -= P;
          }
          for (int II = L; II <= K; II++) {
            P = X * E_(II, KA) + Y * E_(II, KA + 1);
            if (NOTLAS) { P += Z * E_(II, KA + 2); E_(II, KA + 2) -= P * R; }
            E_(II, KA + 1) -= P * Q;
            E_(II, KA) -= P;
          }
        }
      }
    }
  }
  /* 530: back-substitution for the eigenvectors of the triangular form */
  if (RNORM != 0.0) {
    for (int N = M; N >= 1; N--) {
      int N2 = N;
      A_(N, N) = 1.0;
      for (int I = N - 1; I >= 1; I--) {
        double W = A_(I, I) - eval[N - 1];
        if (W == 0.0) W = TOL * RNORM;
        double R = A_(I, N);
        for (int J = N2; J <= N - 1; J++) R += A_(I, J) * A_(J, N);
        A_(I, N) = -R / W;
        N2 = I;
      }
    }
    for (int I = 1; I <= M; I++)
      if (I < L || I > K) for (int J = I; J <= M; J++) E_(I, J) = A_(I, J);
    if (K != 0) {
      for (int J = M; J >= L; J--)
        for (int I = L; I <= K; I++) {
          double Z = 0.0;
          for (int N = L; N <= std::min(J, K); N++) Z += E_(I, N) * A_(N, J);
          E_(I, J) = Z;
        }
    }
  }
  for (int I = L; I <= K; I++) for (int J = 1; J <= M; J++) E_(I, J) *= W_(I);
  for (int I = L - 1; I >= 1; I--) {
    int J = (int)W_(I);
    if (I != J) for (int N = 1; N <= M; N++) std::swap(E_(I, N), E_(J, N));
  }
  for (int I = K + 1; I <= M; I++) {
    int J = (int)W_(I);
    if (I != J) for (int N = 1; N <= M; N++) std::swap(E_(I, N), E_(J, N));
  }
done:
#undef A_
#undef E_
#undef W_
  return IER;
}

/* General band matrix, LAPACK GB storage with room for fill-in: A(i,j) at ab[(kv + i - j) + ldab*j],
 * kv = kl+ku, ldab = 2*kl+ku+1 (0-based i,j).  DGBTF2 / DGBTRS('N') semantics. */
struct BandLU {
  int n = 0, kl = 0, ku = 0, kv = 0, ldab = 0;
  std::vector<double> ab;
  std::vector<int> ipiv;
  void init(int n_, int kl_, int ku_) {
    n = n_; kl = kl_; ku = ku_; kv = kl + ku; ldab = 2 * kl + ku + 1;
    ab.assign((size_t)ldab * n, 0.0); ipiv.assign(n, 0);
  }
  inline double &at(int i, int j) { return ab[(size_t)(kv + i - j) + (size_t)ldab * j]; }
  int factor() {
    int info = 0, ju = 0;
    for (int j = 0; j < n; j++) {
      int km = std::min(kl, n - 1 - j);
      int jp = 0; double best = -1.0;
      for (int p = 0; p <= km; p++) { double v = std::fabs(ab[(size_t)(kv + p) + (size_t)ldab * j]); if (v > best) { best = v; jp = p; } }
      ipiv[j] = j + jp;
      double piv = ab[(size_t)(kv + jp) + (size_t)ldab * j];
      if (piv != 0.0) {
        ju = std::max(ju, std::min(j + ku + jp, n - 1));
        if (jp != 0) for (int c = j; c <= ju; c++) std::swap(at(j + jp, c), at(j, c));
        if (km > 0) {
          double r = 1.0 / at(j, j);
          for (int p = 1; p <= km; p++) at(j + p, j) *= r;
          for (int c = j + 1; c <= ju; c++) {
            double ajc = at(j, c);
            if (ajc != 0.0) for (int p = 1; p <= km; p++) at(j + p, c) -= at(j + p, j) * ajc;
          }
        }
      } else if (info == 0) info = j + 1;
    }
    return info;
  }
  void solve(double *b) {
    for (int j = 0; j < n; j++) {
      int lm = std::min(kl, n - 1 - j), l = ipiv[j];
      if (l != j) std::swap(b[l], b[j]);
      double bj = b[j];
      for (int p = 1; p <= lm; p++) b[j + p] -= bj * at(j + p, j);
    }
    for (int j = n - 1; j >= 0; j--) {
      b[j] /= at(j, j);
      double bj = b[j];
      for (int i = std::max(0, j - kv); i < j; i++) b[i] -= bj * at(i, j);
    }
  }
};

}  // namespace vl
