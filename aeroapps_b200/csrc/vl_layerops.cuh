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
  double *rt;                    // [q][2][n][ld]  R then Tm, row-major with the team-scratch leading dimension ld = mat_ld(n), followed by
                                 //                sg_up, sg_dn [2][n]: one contiguous block of ops_block_doubles(n) per problem (one bulk copy)
  double *spT, *smT;             // [q][r][e]  = S+(e,r), S-(e,r): transposed so that lanes over e read consecutive words
};
HDFN size_t ops_block_doubles(int n) { return 2 * (size_t)n * mat_ld(n) + 2 * n; }

// team scratch in doubles: Am (later E), Bm, Gp (later R), Tb: n x ld each (ld = mat_ld(n)); 4 boundary vectors; inverse work; ints: 2n
// (CTA teams: + 3n + 2 for the small arrays of the second inverse of the pair; ints: 4n)
HDFN size_t ops_smem_doubles(int n, bool cta) { return 4 * (size_t)n * mat_ld(n) + 4 * n + inv_work_doubles(n, cta) + (cta ? 3 * (size_t)n + 2 : 0); }

template <int S, class Team>
DEVFN void layer_ops_problem(const Team &T, const VLPix &X, const VLSol &Z, const VLOps &O, int p, int k, double *scr, int *piv) {
  const int n = Z.n, L = X.L, ld = mat_ld(n);
  const size_t q = (size_t)p * L + k;
  double *Am = scr, *Bm = Am + (size_t)n * ld, *Gp = Bm + (size_t)n * ld, *Tb = Gp + (size_t)n * ld;
  double *wdt = Tb + (size_t)n * ld, *wut = wdt + n, *wdb = wut + n, *wub = wdb + n;   // w_dn(top), w_u(top), w_dn(bot), w_u(bot)
  double *wk = wub + n;
  const double *XP = Z.xp + q * n * n, *XN = Z.xn + q * n * n, *td = Z.tde + q * n;
  const double *WU = Z.wupper + q * 2 * n, *WL = Z.wlower + q * 2 * n;
  for (int r = T.rank; r < n; r += T.size) {
    const double sg = ((r % S) > 1) ? -1.0 : 1.0;              // u = D * upwelling
    wdt[r] = WU[r]; wut[r] = sg * WU[n + r]; wdb[r] = WL[r]; wub[r] = sg * WL[n + r];
  }
#ifndef AB_HOST_MODEL
  if constexpr (Team::is_cta) {
    // CTA teams: S+ and S- are inverted together (one barrier per pivot column for both).  Slots: Am <- P + M t, Bm <- P - M t, the second
    // mirror in Gp..Tb; then Gp <- (M + P t), Tb <- Gp S+ (= G+); Gp <- (M - P t), R -> Tb, Tm -> Am.
    if (2 * (size_t)n * ld >= (size_t)n * inv_ldm(n)) {
      double *sm1 = wk + inv_work_doubles(n, true);
      for (int idx = T.rank; idx < n * n; idx += T.size) {
        const int e = idx / n, r = idx % n;
        const double P = XP[idx], Mt = XN[idx] * td[e];
        Am[(size_t)r * ld + e] = P + Mt; Bm[(size_t)r * ld + e] = P - Mt;
      }
      T.sync();
      if (!team_inverse_pair(T, n, Am, Bm, ld, wk, Gp, sm1, piv)) { if (T.rank == 0) Z.fail[p] = 1; return; }
      double *SP = O.spT + q * n * n, *SM = O.smT + q * n * n;
      for (int idx = T.rank; idx < n * n; idx += T.size) {
        const int r = idx / n, e = idx % n;
        SP[idx] = Am[(size_t)e * ld + r]; SM[idx] = Bm[(size_t)e * ld + r];
        Gp[(size_t)e * ld + r] = XN[idx] + XP[idx] * td[r];            // global [r][e] read as (eigen r, stream e): Gp(stream, eigen) = M + P t
      }
      T.sync();
      team_mm(T, n, n, n, Gp, ld, Am, ld, [&](int r, int c, double acc) { Tb[(size_t)r * ld + c] = acc; });          // G+
      T.sync();
      for (int idx = T.rank; idx < n * n; idx += T.size) { const int e = idx / n, r = idx % n; Gp[(size_t)r * ld + e] = XN[idx] - XP[idx] * td[e]; }
      T.sync();
      team_mm(T, n, n, n, Gp, ld, Bm, ld, [&](int r, int c, double acc) {                                           // G-
        const double gp = Tb[(size_t)r * ld + c];
        Tb[(size_t)r * ld + c] = 0.5 * (gp + acc);                     // R
        Am[(size_t)r * ld + c] = 0.5 * (gp - acc);                     // Tm (S+ is dead: its product is done and it is in HBM)
      });
      T.sync();
      double *Rg = O.rt + q * ops_block_doubles(n), *Tg = Rg + (size_t)n * ld, *sg = Tg + (size_t)n * ld;
      for (int idx = T.rank; idx < n * ld; idx += T.size) {
        const bool pad = (idx % ld) >= n;
        Rg[idx] = pad ? 0.0 : Tb[idx]; Tg[idx] = pad ? 0.0 : Am[idx];
      }
      for (int r = T.rank; r < n; r += T.size) {
        const double *Rr = Tb + (size_t)r * ld, *Tr = Am + (size_t)r * ld;
        double su = wut[r], sd = wdb[r];
        for (int c = 0; c < n; c++) { su -= Rr[c] * wdt[c] + Tr[c] * wub[c]; sd -= Tr[c] * wdt[c] + Rr[c] * wub[c]; }
        sg[r] = su; sg[n + r] = sd;
      }
      T.sync();
      return;
    }
  }
#endif
  for (int sgn = 0; sgn < 2; sgn++) {
    const double s = sgn ? -1.0 : 1.0;
    for (int idx = T.rank; idx < n * n; idx += T.size) {
      const int e = idx / n, r = idx % n;                      // global [e][r]: consecutive lanes read consecutive words
      const double P = XP[idx], M = XN[idx], t = td[e];
      Am[(size_t)r * ld + e] = P + s * M * t;
      Bm[(size_t)r * ld + e] = M + s * P * t;
    }
    T.sync();
    if (!team_inverse(T, n, Am, Am, ld, wk, piv)) { if (T.rank == 0) Z.fail[p] = 1; return; }      // Am <- S = (P +- M t)^-1, S(e, r)
    double *ST = sgn ? O.smT + q * n * n : O.spT + q * n * n;
    for (int idx = T.rank; idx < n * n; idx += T.size) { const int r = idx / n, e = idx % n; ST[idx] = Am[(size_t)e * ld + r]; }
    if (!sgn) {
      team_mm(T, n, n, n, Bm, ld, Am, ld, [&](int r, int c, double acc) { Gp[(size_t)r * ld + c] = acc; });
    } else {
      // R -> Gp (in place, element by element), Tm -> Tb
      team_mm(T, n, n, n, Bm, ld, Am, ld, [&](int r, int c, double acc) {
        const double gp = Gp[(size_t)r * ld + c];
        Gp[(size_t)r * ld + c] = 0.5 * (gp + acc);
        Tb[(size_t)r * ld + c] = 0.5 * (gp - acc);
      });
    }
    T.sync();
  }
  double *Rg = O.rt + q * ops_block_doubles(n), *Tg = Rg + (size_t)n * ld, *sg = Tg + (size_t)n * ld;
  for (int idx = T.rank; idx < n * ld; idx += T.size) {
    const bool pad = (idx % ld) >= n;                          // zero padding: the block is copied as a whole by the solver
    Rg[idx] = pad ? 0.0 : Gp[idx]; Tg[idx] = pad ? 0.0 : Tb[idx];
  }
  for (int r = T.rank; r < n; r += T.size) {
    const double *Rr = Gp + (size_t)r * ld, *Tr = Tb + (size_t)r * ld;
    double su = wut[r], sd = wdb[r];
    for (int c = 0; c < n; c++) { su -= Rr[c] * wdt[c] + Tr[c] * wub[c]; sd -= Tr[c] * wdt[c] + Rr[c] * wub[c]; }
    sg[r] = su; sg[n + r] = sd;
  }
  T.sync();
}

}  // namespace vlk
