// Stage B, kernel K2b: one Team per (pixel, layer) problem of one Fourier component: the layer's discrete-ordinate solution
// (K2: X+, X-, exp(-k delta), Green's-function particular integral) is turned into the layer's INTERACTION OPERATORS
//     reflection R, transmission Tm (n x n), solution-recovery operators S+, S- (n x n), source vectors sg_up, sg_dn (n)
// which is everything the boundary-value problem needs from the layer (vl_add.cuh).
//
// Notation (n = nstokes*nstreams; P(r,e) = XPOS of the downwelling stream r, M(r,e) = XNEG, t_e = exp(-k_e delta); the upwelling
// halves follow by the symmetry XPOS(i+N) = D XNEG(i), XNEG(i+N) = D XPOS(i), D = diag(1,1,-1) per stream, VM/vlidort_solutions.f90:
// 726-756).  With u = D * (upwelling field) the layer equations of BVP_MATRIX_SETUP / BVP_COLUMN_SETUP (VM/vlidort_bvproblem.f90:
// 627-876,1743-2121) read, a = LCON, b = MCON:
//     I_dn(top) - w_dn(top) = P a + M t b        u(bot) - w_u(bot) = M t a + P b          (incoming)
//     u(top)    - w_u(top)  = M a + P t b        I_dn(bot) - w_dn(bot) = P t a + M b      (outgoing)
// Sum and difference decouple them:  in+ = (P + M t)(a + b),  in- = (P - M t)(a - b),  out+ = (M + P t)(a + b),  out- = (M - P t)(a - b)
//     S+ = (P + M t)^-1,  S- = (P - M t)^-1,  G+ = (M + P t) S+,  G- = (M - P t) S-,   R = (G+ + G-)/2,  Tm = (G+ - G-)/2
//     u(top) = R I_dn(top) + Tm u(bot) + sg_up,    I_dn(bot) = Tm I_dn(top) + R u(bot) + sg_dn
//     sg_up = w_u(top) - R w_dn(top) - Tm w_u(bot),   sg_dn = w_dn(bot) - Tm w_dn(top) - R w_u(bot)
// The same linear system as the reference's band matrix, block by block; only the elimination order differs.
#pragma once
#include "vl_eigen.cuh"
#include "vl_linalg.cuh"

namespace vlk {

struct VLOps {                   // per problem q = p*L + k
  double *rk, *tk;               // [q][r][c]  row-major n x n
  double *spT, *smT;             // [q][r][e]  = S+(e,r), S-(e,r): transposed so that lanes over e read consecutive words
  double *sgu, *sgd;             // [q][n]
};

HDFN int ops_ldn(int n) { return n | 1; }
HDFN int ops_ldw(int n) { return (2 * n + 1) | 1; }
// team scratch in doubles: Aug (n x ldw), Bm, Sb, Gp (n x ldn each), 5 vectors, 3 pivot words; ints: n pivots
HDFN size_t ops_smem_doubles(int n) { return (size_t)n * ops_ldw(n) + 3 * (size_t)n * ops_ldn(n) + 5 * n + 4; }

template <int S, class Team>
DEVFN void layer_ops_problem(const Team &T, const VLPix &X, const VLSol &Z, const VLOps &O, int p, int k, double *scr, int *piv) {
  const int n = Z.n, L = X.L, ldn = ops_ldn(n), ldw = ops_ldw(n);
  const size_t q = (size_t)p * L + k;
  double *Aug = scr, *Bm = Aug + (size_t)n * ldw, *Sb = Bm + (size_t)n * ldn, *Gp = Sb + (size_t)n * ldn;
  double *wdt = Gp + (size_t)n * ldn, *wut = wdt + n, *wdb = wut + n, *wub = wdb + n;   // w_dn(top), w_u(top), w_dn(bot), w_u(bot)
  double *dinv = wub + n;
  unsigned long long *best = reinterpret_cast<unsigned long long *>(dinv + n);
  const double *XP = Z.xp + q * n * n, *XN = Z.xn + q * n * n, *td = Z.tde + q * n;
  const double *WU = Z.wupper + q * 2 * n, *WL = Z.wlower + q * 2 * n;
  for (int r = T.rank; r < n; r += T.size) {
    const double sg = ((r % S) > 1) ? -1.0 : 1.0;              // u = D * upwelling
    wdt[r] = WU[r]; wut[r] = sg * WU[n + r]; wdb[r] = WL[r]; wub[r] = sg * WL[n + r];
  }
  for (int sgn = 0; sgn < 2; sgn++) {
    const double s = sgn ? -1.0 : 1.0;
    for (int idx = T.rank; idx < n * n; idx += T.size) {
      const int e = idx / n, r = idx % n;                      // global [e][r]: consecutive lanes read consecutive words
      const double P = XP[idx], M = XN[idx], t = td[e];
      Aug[(size_t)r * ldw + e] = P + s * M * t;
      Bm[(size_t)r * ldn + e] = M + s * P * t;
    }
    for (int idx = T.rank; idx < n * n; idx += T.size) { const int r = idx / n, c = idx % n; Aug[(size_t)r * ldw + n + c] = (r == c) ? 1.0 : 0.0; }
    if (!team_gauss_jordan(T, n, 2 * n, Aug, ldw, piv, best)) { if (T.rank == 0) Z.fail[p] = 1; return; }
    // S(e, r) = Aug[piv[e]][n + r] / Aug[piv[e]][e]
    double *ST = sgn ? O.smT + q * n * n : O.spT + q * n * n;
    for (int e = T.rank; e < n; e += T.size) dinv[e] = 1.0 / Aug[(size_t)piv[e] * ldw + e];
    T.sync();
    for (int idx = T.rank; idx < n * n; idx += T.size) {
      const int r = idx / n, e = idx % n;
      const double v = Aug[(size_t)piv[e] * ldw + n + r] * dinv[e];
      Sb[(size_t)e * ldn + r] = v;
      ST[idx] = v;
    }
    T.sync();
    if (!sgn) {
      team_gemm_auto(T, n, n, n, Bm, ldn, Sb, ldn, [&](int r, int c, double acc) { Gp[(size_t)r * ldn + c] = acc; });
    } else {
      // R -> Gp (in place, element by element), Tm -> the dead augmented array
      team_gemm_auto(T, n, n, n, Bm, ldn, Sb, ldn, [&](int r, int c, double acc) {
        const double gp = Gp[(size_t)r * ldn + c];
        Gp[(size_t)r * ldn + c] = 0.5 * (gp + acc);
        Aug[(size_t)r * ldw + c] = 0.5 * (gp - acc);
      });
    }
    T.sync();
  }
  double *Rg = O.rk + q * n * n, *Tg = O.tk + q * n * n;
  for (int idx = T.rank; idx < n * n; idx += T.size) {
    const int r = idx / n, c = idx % n;
    Rg[idx] = Gp[(size_t)r * ldn + c]; Tg[idx] = Aug[(size_t)r * ldw + c];
  }
  for (int r = T.rank; r < n; r += T.size) {
    const double *Rr = Gp + (size_t)r * ldn, *Tr = Aug + (size_t)r * ldw;
    double su = wut[r], sd = wdb[r];
    for (int c = 0; c < n; c++) { su -= Rr[c] * wdt[c] + Tr[c] * wub[c]; sd -= Tr[c] * wdt[c] + Rr[c] * wub[c]; }
    O.sgu[q * n + r] = su; O.sgd[q * n + r] = sd;
  }
  T.sync();
}

}  // namespace vlk
