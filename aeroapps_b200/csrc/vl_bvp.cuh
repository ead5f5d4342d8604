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

// optional phase timers of K3 (thread 0 of every CTA, clock64 deltas summed per phase); tools/bvp_timing.py
#if defined(AB200_BVP_TIMING) && !defined(AB_HOST_MODEL)
__device__ unsigned long long g_bvp_tim[16];
#define VLT_DECL long long vlt_t = clock64(), vlt_n = 0; (void)vlt_n;
#define VLT(i) { if (threadIdx.x == 0) { vlt_n = clock64(); atomicAdd(&g_bvp_tim[i], (unsigned long long)(vlt_n - vlt_t)); vlt_t = vlt_n; } }
#else
#define VLT_DECL
#define VLT(i)
#endif

struct VLSurf {                  // BRDF Fourier components per pixel (null pointers for Lambertian)
  const double *brdf_f;          // [p][m][q][i][j]   q = 4*o1+o2 < nq
  const double *brdf_f_0;        // [p][m][q][i]
  const double *user_brdf_f;     // [p][m][q][j]
  const double *user_brdf_f_0;   // [p][m][q]         (not needed: the direct bounce comes from the FO code)
  int nmfour;                    // number of Fourier components stored (2N)
  int nq;                        // 1: only the (1,1) element is non-zero (scalar kernels), 16: Mueller layout
};

struct VLConv {                  // per pixel-view v, persistent over the Fourier loop
  double *total;                 // [v][4]
  int *state;                    // [v][4]: 0 active(1)/done(0), 1 testconv, 2 fourier_saved, 3 failure code
  int *col_active;               // [c]: 1 while any view of pixel c is still in the Fourier loop
};

HDFN int bvp_cols(int n) { return 4 * n + 1; }
HDFN int bvp_stride(int n) { return (4 * n + 1) | 1; }                         // odd row stride
constexpr int BVP_NB = 8;                                                        // panel width of the blocked elimination
constexpr int BVP_LS = BVP_NB + 1;   // row stride of the multiplier panels: odd, so that the per-row stores of a panel column are conflict-free
// big region: forward phase = window + double-buffered panel arrays; backward phase = two buffers of pivot rows
HDFN size_t bvp_big_doubles(int n) {
  const size_t fwd = (size_t)3 * n * bvp_stride(n) + 2 * ((size_t)3 * n * BVP_LS + (size_t)BVP_NB * bvp_stride(n) + BVP_NB * BVP_NB);
  const size_t bwd = (size_t)2 * (2 * n) * bvp_cols(n);
  return ((fwd > bwd ? fwd : bwd) + 1) & ~(size_t)1;
}
HDFN size_t bvp_smem_doubles(int n) { return bvp_big_doubles(n) + 8 * n + 16; }
// "large n" layout (3*nstreams*nstokes rows no longer fit the 227 KB of shared memory, e.g. nstokes 3 x 16 streams): the
// window lives in a per-CTA HBM scratch (L1/L2 resident), shared memory keeps the panel arrays and ONE buffer of pivot rows
HDFN size_t bvp_big_doubles_g(int n) {
  const size_t fwd = 2 * ((size_t)3 * n * BVP_LS + (size_t)BVP_NB * bvp_stride(n) + BVP_NB * BVP_NB);
  const size_t bwd = (size_t)(2 * n) * bvp_cols(n);
  return ((fwd > bwd ? fwd : bwd) + 1) & ~(size_t)1;
}
HDFN size_t bvp_smem_doubles_g(int n) { return bvp_big_doubles_g(n) + 8 * n + 16; }
HDFN size_t bvp_window_doubles(int n) { return (size_t)3 * n * bvp_stride(n); }
HDFN size_t bvp_work_doubles(int n, int L) { return (size_t)L * 2 * n * bvp_cols(n); }

// ---- panel factorisation of the blocked elimination, executed by the panel sub-team (ranks < Team::PW).
// Lane `rank` owns the window rows rank + PW*u.  upd: first apply the previous panel (LpP, UbP; nbp columns) to the
// nb columns jc0.. of this panel.  Multipliers go to LpN (R x NB, zero for rows that are not active), the
// multipliers of the pivot rows inside the panel to LUpN, the pivot rows' own panel entries back to W.
template <int RU, class Team>
DEVFN void bvp_panel(const Team &T, double *AB_RESTRICT W, int Cs, int R, int *state, int *pivrow, int jc0, int nb, bool upd, int nbp,
                     const double *AB_RESTRICT LpP, const double *AB_RESTRICT UbP, double *AB_RESTRICT LpN, double *AB_RESTRICT LUpN, int *iflag,
                     int *rflagN) {
  constexpr int PW = Team::PW;
  double pv[RU][BVP_NB];
  bool act[RU], nz[RU];
#pragma unroll
  for (int u = 0; u < RU; u++) {
    const int r = T.rank + PW * u;
    act[u] = (r < R) && state[r] == 0;
    nz[u] = false;
#pragma unroll
    for (int j = 0; j < BVP_NB; j++) pv[u][j] = (act[u] && j < nb) ? W[(size_t)r * Cs + jc0 + j] : 0.0;
    if (upd && act[u]) {
#pragma unroll
      for (int jj = 0; jj < BVP_NB; jj++) {
        const double l = jj < nbp ? LpP[r * BVP_LS + jj] : 0.0;
#pragma unroll
        for (int j = 0; j < BVP_NB; j++) if (j < nb) pv[u][j] -= l * UbP[jj * Cs + jc0 + j];
      }
    }
  }
  bool pfail = false;
#pragma unroll
  for (int j = 0; j < BVP_NB; j++) {
    if (j < nb && !pfail) {
      double best = -1.0; int bidx = R;
#pragma unroll
      for (int u = 0; u < RU; u++) { const double a = fabs(pv[u][j]); if (act[u] && a > best) { best = a; bidx = T.rank + PW * u; } }
      best = T.pargmax(best, bidx);
      if (!(best > 0.0)) pfail = true;
      else {
        const int owner = bidx % PW, ustar = bidx / PW;          // ustar is uniform over the sub-team
        const bool mine = (T.rank == owner);
        double prow[BVP_NB];
#pragma unroll
        for (int jj = 0; jj < BVP_NB; jj++) prow[jj] = 0.0;
#pragma unroll
        for (int u = 0; u < RU; u++)
          if (u == ustar) {                                      // uniform branch: plain register copies
#pragma unroll
            for (int jj = 0; jj < BVP_NB; jj++) prow[jj] = pv[u][jj];
            if (mine) act[u] = false;
          }
        T.pbcast_n(prow, j, owner);
        const double ipiv = 1.0 / prow[j];
        if (mine) {                                              // this row becomes pivot row j of the panel
          state[bidx] = 1 + jc0 + j; pivrow[jc0 + j] = bidx;
#pragma unroll
          for (int jj = 0; jj < BVP_NB; jj++) if (jj >= j && jj < nb) W[(size_t)bidx * Cs + jc0 + jj] = prow[jj];
        }
#pragma unroll
        for (int u = 0; u < RU; u++) {
          const int r = T.rank + PW * u;
          double l = 0.0;
          if (act[u]) {
            l = pv[u][j] * ipiv;
            nz[u] = nz[u] || (l != 0.0);
#pragma unroll
            for (int jj = 0; jj < BVP_NB; jj++) if (jj > j) pv[u][jj] -= l * prow[jj];
          }
          if (r < R) LpN[r * BVP_LS + j] = l;                    // multipliers straight to shared memory
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < RU; u++) { const int r = T.rank + PW * u; if (r < R) LpN[r * BVP_LS + j] = 0.0; }
    }
  }
#pragma unroll
  for (int u = 0; u < RU; u++) { const int r = T.rank + PW * u; if (r < R) rflagN[r] = nz[u] ? 1 : 0; }
  T.psync();
  // multipliers of the pivot rows inside the panel -> LUpN; then clear those rows so the trailing update skips them
  if (!pfail) {
    for (int idx = T.rank; idx < BVP_NB * BVP_NB; idx += PW) {
      const int j = idx / BVP_NB, jj = idx % BVP_NB;
      if (j < nb) LUpN[idx] = (jj < j) ? LpN[pivrow[jc0 + j] * BVP_LS + jj] : 0.0;
    }
    T.psync();
    for (int idx = T.rank; idx < BVP_NB * BVP_NB; idx += PW) {
      const int j = idx / BVP_NB, jj = idx % BVP_NB;
      if (j < nb) LpN[pivrow[jc0 + j] * BVP_LS + jj] = 0.0;
    }
  }
  if (pfail && T.rank == 0) iflag[0] = 1;
}

template <class Team>
DEVFN void bvp_panel_dispatch(const Team &T, double *W, int Cs, int R, int *state, int *pivrow, int jc0, int nb, bool upd, int nbp,
                              const double *LpP, const double *UbP, double *LpN, double *LUpN, int *iflag, int *rflagN) {
  if constexpr (Team::PW == 1) bvp_panel<Team::PRU>(T, W, Cs, R, state, pivrow, jc0, nb, upd, nbp, LpP, UbP, LpN, LUpN, iflag, rflagN);
  else {
    if (R <= Team::PW) bvp_panel<1>(T, W, Cs, R, state, pivrow, jc0, nb, upd, nbp, LpP, UbP, LpN, LUpN, iflag, rflagN);
    else bvp_panel<2>(T, W, Cs, R, state, pivrow, jc0, nb, upd, nbp, LpP, UbP, LpN, LUpN, iflag, rflagN);
  }
}

// pivot rows of a factorised panel, column-parallel: U(j, c) = W(pr_j, c) - sum_{j' < j} L(pr_j, j') U(j', c), c >= jc0 + nb
template <class Team>
DEVFN void bvp_urows(const Team &T, double *AB_RESTRICT W, int Cs, int Cc, const int *pivrow, int jc0, int nb, const double *AB_RESTRICT LUpN,
                     double *AB_RESTRICT UbN) {
  for (int c = jc0 + nb + T.rank; c < Cc; c += T.size) {
    double uv[BVP_NB];
#pragma unroll
    for (int j = 0; j < BVP_NB; j++) {
      if (j < nb) {
        const int pr = pivrow[jc0 + j];
        double v = W[(size_t)pr * Cs + c];
#pragma unroll
        for (int jj = 0; jj < BVP_NB; jj++) if (jj < j) v -= LUpN[j * BVP_NB + jj] * uv[jj];
        uv[j] = v;
        W[(size_t)pr * Cs + c] = v;
        UbN[j * Cs + c] = v;
      } else { uv[j] = 0.0; UbN[j * Cs + c] = 0.0; }
    }
  }
}

template <int S, class Team>
DEVFN void bvp_backward(const Team &T, const VLPix &X, const VLSol &Z, const VLQuad &QD, const VLSurf &SF, const VLConv &CV,
                        int lambertian, double accuracy, int m, int p, double *W, double *xcur, double *xnext, double *red, const double *Ug,
                        int nbuf = 2);

template <int S, class Team, bool WGLOB = false>
DEVFN void bvp_pixel(const Team &T, const VLPix &X, const VLSol &Z, const VLQuad &QD, const VLSurf &SF, const VLConv &CV,
                     int lambertian, double accuracy, int m, int p, double *scr, int *iscr, double *Ug, double *wglob_ = nullptr) {
  // WGLOB is a compile-time switch so that, in the default layout, the compiler still sees W as a shared-memory pointer
  double *wglob = WGLOB ? wglob_ : nullptr;
  const int N = QD.N, n = S * N, n2 = 2 * n, L = X.L, NM = X.NM;
  const int R = 3 * n, Cc = bvp_cols(n), Cs = bvp_stride(n), RH = 4 * n;      // RH = rhs column
  // wglob != nullptr: "large n" layout, the window in HBM scratch and only the panel arrays in the team scratch
  double *W = wglob ? wglob : scr;
  double *xcur = scr + (wglob ? bvp_big_doubles_g(n) : bvp_big_doubles(n));  // 2n
  double *xnext = xcur + n2;              // 2n
  double *red = xnext + n2;               // 4n scratch
  double *Lp = wglob ? scr : W + (size_t)R * Cs;   // 2 x (R x NB) multipliers of the current / next panel
  double *Ub = Lp + (size_t)2 * R * BVP_LS;   // 2 x (NB x Cs) finished pivot rows of the current / next panel
  double *LUp = Ub + (size_t)2 * BVP_NB * Cs; // 2 x (NB x NB) multipliers of the pivot rows inside the panel
  int *state = iscr;                      // R : -1 unused, 0 active, 1+t pivot number t
  int *pivrow = state + R;                // 2n
  int *freeslot = pivrow + n2;            // 2n
  int *iflag = freeslot + n2;             // failure flag of the panel factorisation
  int *gctr = iflag + 2;                  // work counter of the trailing update
  int *rflag = iflag + 4;                 // 2 x R: row has non-zero multipliers in the current / next panel
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

  VLT_DECL
  if (T.rank == 0) iflag[0] = 0;
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
  VLT(0)

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
    VLT(1)
    // ---- eliminate the 2n columns of layer k: blocked (panels of BVP_NB columns), partial pivoting over the active
    //      rows of the window, one panel of look-ahead: while the other warps apply panel P to the trailing columns,
    //      the panel sub-team (warp 0) applies it to the columns of panel P+1 only and factorises that panel out of
    //      its registers (pivot search by warp-wide integer reductions).
    {
      const int npanel = (n2 + BVP_NB - 1) / BVP_NB;
      const bool is_panel = T.rank < Team::PW;
      const int LW = T.size < 32 ? T.size : 32;
      const int lane = T.rank % LW;
      if (is_panel) bvp_panel_dispatch(T, W, Cs, R, state, pivrow, 0, (n2 < BVP_NB ? n2 : BVP_NB), false, 0, Lp, Ub, Lp, LUp, iflag, rflag);
      T.sync();
      if (iflag[0]) { failed = true; break; }
      bvp_urows(T, W, Cs, Cc, pivrow, 0, (n2 < BVP_NB ? n2 : BVP_NB), LUp, Ub);
      T.sync();
      VLT(2)
      for (int P = 0; P < npanel; P++) {
        const int jc0 = P * BVP_NB, nb = (n2 - jc0 < BVP_NB) ? n2 - jc0 : BVP_NB;
        const bool has_next = P + 1 < npanel;
        const int jc1 = jc0 + nb, nb1 = has_next ? ((n2 - jc1 < BVP_NB) ? n2 - jc1 : BVP_NB) : 0;
        double *LpC = Lp + (size_t)(P & 1) * R * BVP_LS, *UbC = Ub + (size_t)(P & 1) * BVP_NB * Cs;
        double *LpN = Lp + (size_t)((P + 1) & 1) * R * BVP_LS, *UbN = Ub + (size_t)((P + 1) & 1) * BVP_NB * Cs;
        double *LUpN = LUp + ((P + 1) & 1) * BVP_NB * BVP_NB;
        const int *rflagC = rflag + (P & 1) * R;
        if (has_next) {
          // everyone: apply panel P to the columns of panel P+1 first, so that warp 0 can start on it right away
          for (int idx = T.rank; idx < R * BVP_NB; idx += T.size) {
            const int r = idx / BVP_NB, j = idx % BVP_NB;
            if (j < nb1) {
              double a = W[(size_t)r * Cs + jc1 + j];
#pragma unroll
              for (int jj = 0; jj < BVP_NB; jj++) a -= LpC[r * BVP_LS + jj] * UbC[jj * Cs + jc1 + j];
              W[(size_t)r * Cs + jc1 + j] = a;
            }
          }
          if (T.rank == 0) gctr[0] = 0;
          T.sync();
        } else {
          if (T.rank == 0) gctr[0] = 0;
          T.sync();
        }
        VLT(3)
        if (is_panel && has_next) bvp_panel_dispatch(T, W, Cs, R, state, pivrow, jc1, nb1, false, nb, LpC, UbC, LpN, LUpN, iflag, rflag + ((P + 1) & 1) * R);
        VLT(4)
        {
          // trailing update: W(r, c) -= sum_j L(r, j) U(j, c) for c >= cbeg; groups of RG rows are handed out to the
          // warps through a shared counter (the panel warps join in when their panel is done)
          constexpr int RG = 4;
          const int cbeg = jc1 + nb1;
          for (;;) {
            const int g = T.wfetch(gctr);
            if (g * RG >= R) break;
            const int r0 = g * RG;
            bool any = false;                          // rows that are not active carry zero multipliers
#pragma unroll
            for (int rr = 0; rr < RG; rr++) any = any || (r0 + rr < R && rflagC[r0 + rr] != 0);
            if (!any) continue;
            double l[RG][BVP_NB];
#pragma unroll
            for (int rr = 0; rr < RG; rr++)
#pragma unroll
              for (int j = 0; j < BVP_NB; j++) l[rr][j] = (r0 + rr < R) ? LpC[(r0 + rr) * BVP_LS + j] : 0.0;
            for (int c = cbeg + lane; c < Cc; c += 2 * LW) {
              const int c2 = c + LW;
              const bool two = c2 < Cc;
              double u0[BVP_NB], u1[BVP_NB];
#pragma unroll
              for (int j = 0; j < BVP_NB; j++) { u0[j] = UbC[j * Cs + c]; u1[j] = two ? UbC[j * Cs + c2] : 0.0; }
#pragma unroll
              for (int rr = 0; rr < RG; rr++) {
                if (r0 + rr < R) {
                  double *row = W + (size_t)(r0 + rr) * Cs;
                  double a0 = row[c], a1 = two ? row[c2] : 0.0;
#pragma unroll
                  for (int j = 0; j < BVP_NB; j++) { a0 -= l[rr][j] * u0[j]; a1 -= l[rr][j] * u1[j]; }
                  row[c] = a0;
                  if (two) row[c2] = a1;
                }
              }
            }
          }
        }
        T.sync();
        VLT(5)
        if (iflag[0]) { failed = true; break; }
        if (has_next) {
          bvp_urows(T, W, Cs, Cc, pivrow, jc1, nb1, LUpN, UbN);
          T.sync();
        }
        VLT(6)
      }
    }
    if (failed) break;
    // ---- spill the pivot rows (U factor + transformed rhs), recycle their slots, shift the carry rows
    {
      double *Uk = Ug + (size_t)k * n2 * Cc;
      const int LW = T.size < 32 ? T.size : 32, lane = T.rank % LW, wid = T.rank / LW, nw = T.size / LW;
      for (int t = wid; t < n2; t += nw) {
        const double *src = W + (size_t)pivrow[t] * Cs; double *dst = Uk + (size_t)t * Cc;
        for (int c = lane; c < Cc; c += LW) dst[c] = src[c];
      }
      T.sync();
      for (int t = T.rank; t < n2; t += T.size) { freeslot[t] = pivrow[t]; state[pivrow[t]] = -1; }
      T.sync();
      // carry rows: the next layer's block moves to the front; everything else is cleared for the new rows
      for (int r = wid; r < R; r += nw) {
        double *row = W + (size_t)r * Cs;
        if (state[r] == 0) { for (int c = lane; c < n2; c += LW) { row[c] = row[n2 + c]; row[n2 + c] = 0.0; } }
        else { for (int c = lane; c < 2 * n2; c += LW) row[c] = 0.0; }
      }
      T.sync();
    }
    VLT(7)
  }
  if (failed) { if (T.rank == 0) { cst[0] = 0; cst[3] = 4; } return; }

  bvp_backward<S>(T, X, Z, QD, SF, CV, lambertian, accuracy, m, p, scr, xcur, xnext, red, Ug, wglob ? 1 : 2);
}

// Back substitution (pivot rows of every layer in Ug, [k][t][0..Cc) in logical column order) fused with the upwelling
// source-function recursion, the Fourier summation and the convergence test.  big: team scratch of bvp_big_doubles(n)
// (two buffers of pivot rows), xcur/xnext/red: 2n, 2n, 4n doubles.
template <int S, class Team>
DEVFN void bvp_backward(const Team &T, const VLPix &X, const VLSol &Z, const VLQuad &QD, const VLSurf &SF, const VLConv &CV,
                        int lambertian, double accuracy, int m, int p, double *W, double *xcur, double *xnext, double *red, const double *Ug,
                        int nbuf) {
  const int N = QD.N, n = S * N, n2 = 2 * n, L = X.L, NM = X.NM;
  const int Cc = bvp_cols(n), RH = 4 * n;
  const double *sc = X.scal + (size_t)p * 16;
  const int *ib = X.ibuf + (size_t)p * 4;
  int *cst = CV.state + (size_t)p * 4;
  const double albedo = sc[4];
  const int cutoff = ib[0];
  bool include_surface = true;
  if (lambertian) { if (m != 0) include_surface = false; else if (albedo == 0.0) include_surface = false; }
  const double SFAC = (m == 0) ? 2.0 : 1.0, DELTA_FACTOR = (m == 0) ? 1.0 : 2.0;
  const double *UBF = nullptr;
  const int nq = SF.nq;
  auto UBFq = [&](int o1, int o2, int jj) -> double {
    return nq == 16 ? UBF[(size_t)(4 * o1 + o2) * N + jj] : ((o1 | o2) == 0 ? UBF[jj] : 0.0);
  };
  if (!lambertian) UBF = SF.user_brdf_f + ((size_t)p * SF.nmfour + m) * SF.nq * N;
  const size_t q0 = (size_t)p * L;
  VLT_DECL
  double cum[S];
#pragma unroll
  for (int o = 0; o < S; o++) cum[o] = 0.0;
  for (int t = T.rank; t < n2; t += T.size) xnext[t] = 0.0;
  T.sync();
  // pivot rows of layer k, [t][0..Cc) (Cs == Cc), rhs moved to column n2; two buffers: layer k-1 streams in (TMA bulk
  // copy) while layer k is solved
  const int Ubs = Cc;
  const size_t usz = (size_t)n2 * Cc;
  T.load_begin(W, Ug + (size_t)(L - 1) * usz, usz);
  for (int k = L - 1; k >= 0; k--) {
    const size_t qa = q0 + k;
    double *Ut = W + (size_t)(nbuf == 2 ? ((L - 1 - k) & 1) : 0) * usz;
    T.load_wait();
    if (k > 0 && nbuf == 2) T.load_begin(W + (size_t)((L - k) & 1) * usz, Ug + (size_t)(k - 1) * usz, usz);
    for (int t = T.rank; t < n2; t += T.size) {
      const double *un = Ut + (size_t)t * Ubs + n2;
      double b = Ut[(size_t)t * Ubs + RH];
      for (int c = 0; c < n2; c++) b -= un[c] * xnext[c];
      Ut[(size_t)t * Ubs + n2] = b;
    }
    T.sync();
    VLT(8)
    // upper-triangular solve (row t has its pivot in column t): blocks of 8 unknowns, the small diagonal block
    // solved by one thread out of registers, the rest of the right-hand side updated by everyone
    {
      constexpr int TB = 8;
      for (int t1 = n2; t1 > 0; t1 -= TB) {
        const int t0 = t1 - TB > 0 ? t1 - TB : 0, nt = t1 - t0;
        if (T.rank == 0) {
          double uu[TB][TB], bb[TB], xx[TB];
#pragma unroll
          for (int a = 0; a < TB; a++) {
            bb[a] = a < nt ? Ut[(size_t)(t0 + a) * Ubs + n2] : 0.0;
#pragma unroll
            for (int b = 0; b < TB; b++) uu[a][b] = (a < nt && b < nt && b >= a) ? Ut[(size_t)(t0 + a) * Ubs + t0 + b] : (a == b ? 1.0 : 0.0);
          }
#pragma unroll
          for (int a = 0; a < TB; a++) uu[a][a] = 1.0 / uu[a][a];
#pragma unroll
          for (int a = TB - 1; a >= 0; a--) {
            xx[a] = bb[a] * uu[a][a];
#pragma unroll
            for (int b = 0; b < TB; b++) if (b < a) bb[b] -= uu[b][a] * xx[a];
          }
#pragma unroll
          for (int a = 0; a < TB; a++) if (a < nt) xcur[t0 + a] = xx[a];
        }
        T.sync();
        for (int r = T.rank; r < t0; r += T.size) {
          double b = Ut[(size_t)r * Ubs + n2];
          for (int a = 0; a < nt; a++) b -= Ut[(size_t)r * Ubs + t0 + a] * xcur[t0 + a];
          Ut[(size_t)r * Ubs + n2] = b;
        }
        T.sync();
      }
    }
    VLT(9)
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
    if (k > 0 && nbuf == 1) T.load_begin(W, Ug + (size_t)(k - 1) * usz, usz);      // single buffer: no overlap
    VLT(10)
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
