// Stage B, kernel K2: one Team (a warp on the GPU) per (pixel, layer) problem of one Fourier component m.
// Computes what VLIDORT_QHOM_SOLUTION, _QHOM_NORMS, _UHOM_SOLUTION, HMULT_MASTER and _GBEAM_SOLUTION compute
// (vlidort_solutions.f90:150-1793,2649-2935), but NOT the way the reference does it:
//
//   SAB = M^-1 (I - Zs W/2),  DAB = M^-1 (I - Zd W/2),   Zs,Zd = sum_l Pi_i B_l (I +- (-1)^(l-m) D) Pi_j
// with Zs, Zd SYMMETRIC for nstokes <= 3 (B_l is block diagonal {I,Q}|{U} and commutes with D).  With the diagonal
// similarity  E = T^-1 M^-1/2 (P Q) M^1/2 T  (T = W^1/2),  P,Q symmetric and Q positive definite, so
//   Q = C C^T,  P = Lp Lp^T (Cholesky),   G = C^T P C = A A^T with A = C^T Lp,
//   one-sided (Hestenes) Jacobi on the columns of A:  A J = U diag(k)   ->  G = U diag(k^2) U^T
//   eigenvector of EIGENMAT:  sigma = T^-1 M^-1/2 C^-T u,      SAB sigma = T^-1 M^-1/2 C u
// i.e. the non-symmetric Hessenberg/QR eigensolver (ASYMTX) of the reference is replaced by a symmetric Jacobi
// solve that is regular, branch-free and whose eigenvectors are exactly orthogonal in the metric the Green's
// function normalisation (NORM_SAVED) assumes -- also inside degenerate clusters.
#pragma once
#include <vector>
#include "vl_setup.cuh"

namespace vlk {

struct VLSol {                 // per problem q = c*L + k (pixel c, layer k; structure of arrays); G view geometries per pixel
  int n, S, N;
  double *keig, *tde;          // [q][n]
  double *xp, *xn;             // [q][e][r]   r = S*i + o, streams i < N (the i+N half follows by symmetry)
  double *uh_dn, *uh_up;       // [q][g][e][S]   UHOM_UPDN, UHOM_UPUP per view
  double *hm1, *hm2;           // [q][g][n]      whole-layer multipliers per view
  double *aterm, *bterm, *gam_m, *gam_p;   // [q][n]
  double *wupper, *wlower;     // [q][2N*S]
  int *fail;                   // [p] set !=0 when an eigenproblem fails
  // generalized spherical functions P, R, T (l = 0..NM) of the current Fourier component at the user cosines [v][3(NM+1)] and at the solar
  // cosine [c][3(NM+1)], filled once per Fourier term by k_vl_pi_angles; null: every problem computes its own (host model)
  const double *piu_g, *pi0_g;
  // multi-geometry calls: per-problem workspace [q][(NM+1)][n][3] for the view-independent moment sums of the user-stream solutions
  // (z_l(e), see eigen_problem); null: every view builds its operator matrices (single-geometry calls, host model without it)
  double *zws;
};

#ifdef AB_HM_SWEEP_HIST
static long hm_sweep_hist[16][32]; static int hm_sweep_last; static double hm_off_hist[32];
#endif
struct VLQuad {                // quadrature (pixel independent)
  int N;
  double x[32], w[32];
  const unsigned short *jsched;  // device: Jacobi pair schedule of the two-warp teams (jacobi_schedule), [jrounds][half]; nullptr = none
  int jrounds;
};

// Host: the recursive fixed / moving tournament for n columns (see jacobi_duo_resident).  out[round * half + slot] = a | b << 6 |
// load_a << 12 | store_a << 13 or 0xFFFF, half = ceil(n / 2) slots per round, followed by one entry per round: 1 = the round must end
// with a team barrier (a column crosses from one warp's slots to the other's), 0 = a warp-level barrier is enough.  Slots
// [0, ppw) belong to warp 0, the rest to warp 1.  The top-level phase (F against M) is split in two halves that keep the moving
// columns inside one warp (F0 x M0 with F1 x M1, then F0 x M1 with F1 x M0), and every later phase is warp-local by construction, so
// a sweep needs a handful of team barriers instead of one per round.  Returns the number of rounds.
inline int jacobi_schedule(int n, std::vector<unsigned short> &out) {
  const int half = (n + (n & 1)) / 2, ppw = (half + 1) / 2;
  std::vector<std::vector<int>> ra, rb;                    // [round][slot] -> column or -1
  auto ensure = [&](size_t r) { while (ra.size() <= r) { ra.emplace_back(half, -1); rb.emplace_back(half, -1); } };
  struct Job { std::vector<int> cols; int slot0, round0; };
  std::vector<Job> todo;
  { Job j; for (int c = 0; c < n; c++) j.cols.push_back(c); j.slot0 = 0; j.round0 = 0; todo.push_back(j); }
  while (!todo.empty()) {
    Job j = todo.back(); todo.pop_back();
    const int sz = (int)j.cols.size();
    if (sz < 2) continue;
    const int hf = sz / 2, nm = sz - hf;
    int used = nm;                                         // rounds of this job's cross phase
    const int f0 = ppw - j.slot0, f1 = hf - f0;            // fixed columns on warp 0 / warp 1 (when the job spans both)
    const int m0 = nm / 2, m1 = nm - m0;
    if (j.slot0 < ppw && f1 > 0 && m0 >= f0 && m0 >= f1) { // (m1 >= m0) warp-local halves
      for (int ph = 0; ph < 2; ph++) {
        const int base = ph == 0 ? 0 : m1;                 // phase a takes max(m0, m1) = m1 rounds, phase b the same
        for (int s = 0; s < m1; s++) {
          ensure((size_t)j.round0 + base + s);
          for (int t = 0; t < hf; t++) {
            const bool w1 = t >= f0;                       // slot on warp 1
            const bool first = (ph == 0) ? !w1 : w1;       // which half of M this slot meets in this phase: M0 (first) or M1
            const int off = first ? 0 : m0, len = first ? m0 : m1, tl = w1 ? t - f0 : t;
            if (s >= len) continue;                        // the shorter half idles in the last round of the phase
            ra[j.round0 + base + s][j.slot0 + t] = j.cols[t];
            rb[j.round0 + base + s][j.slot0 + t] = j.cols[hf + off + (tl + s) % len];
          }
        }
      }
      used = 2 * m1;
    } else {
      for (int s = 0; s < nm; s++) {
        ensure((size_t)j.round0 + s);
        for (int t = 0; t < hf; t++) { ra[j.round0 + s][j.slot0 + t] = j.cols[t]; rb[j.round0 + s][j.slot0 + t] = j.cols[hf + (t + s) % nm]; }
      }
    }
    Job f, m;
    f.cols.assign(j.cols.begin(), j.cols.begin() + hf); f.slot0 = j.slot0; f.round0 = j.round0 + used;
    m.cols.assign(j.cols.begin() + hf, j.cols.end()); m.slot0 = j.slot0 + hf / 2; m.round0 = j.round0 + used;
    // keep a whole sub-tournament on one warp when it fits: the moving half starts on warp 1's first slot
    if (sz == n && hf / 2 <= ppw && nm / 2 <= half - ppw) m.slot0 = ppw;
    todo.push_back(f); todo.push_back(m);
  }
  const int R = (int)ra.size();
  out.assign((size_t)R * half + R, 0xFFFF);
  for (int r = 0; r < R; r++)
    for (int t = 0; t < half; t++) {
      if (ra[r][t] < 0) continue;
      const bool load = r == 0 || ra[r - 1][t] != ra[r][t];
      const bool store = r == R - 1 || ra[r + 1][t] != ra[r][t];
      out[(size_t)r * half + t] = (unsigned short)(ra[r][t] | (rb[r][t] << 6) | (load ? 0x1000 : 0) | (store ? 0x2000 : 0));
    }
  // team barrier after round r iff round r + 1 reads (from shared memory) a column that the OTHER warp has written since the last one
  std::vector<int> writer(n, -1);                          // warp that stored the column since the last team barrier
  for (int r = 0; r < R; r++) {
    for (int t = 0; t < half; t++) {
      if (ra[r][t] < 0) continue;
      const int w = t >= ppw;
      writer[rb[r][t]] = w;
      if (out[(size_t)r * half + t] & 0x2000) writer[ra[r][t]] = w;
    }
    bool need = (r == R - 1);                              // the sweep ends with one (norms are recomputed by the whole team)
    if (r + 1 < R)
      for (int t = 0; t < half && !need; t++) {
        if (ra[r + 1][t] < 0) continue;
        const int w = t >= ppw;
        const int cb = rb[r + 1][t], ca = ra[r + 1][t];
        if (writer[cb] >= 0 && writer[cb] != w) need = true;
        if ((out[(size_t)(r + 1) * half + t] & 0x1000) && writer[ca] >= 0 && writer[ca] != w) need = true;
      }
    out[(size_t)R * half + r] = need ? 1 : 0;
    if (need) std::fill(writer.begin(), writer.end(), -1);
  }
  return R;
}

// shared/scratch size in doubles for one problem
HDFN int eig_lda(int n) { return n | 1; }                                  // odd leading dimension: conflict-free column access
HDFN int eigen_scratch_doubles(int n, int NM) { return n * (n + 1) + n * eig_lda(n) + 18 * n + 6 * (NM + 1) + 8; }

// one Jacobi rotation of the column pair (a,b) of A (length n).  NR > 0: both columns live in registers (n <= NR).
// The squared column norms travel with the columns (na2, nb2: recomputed exactly at the start of every sweep, updated here
// with  |a'|^2 = |a|^2 - t (a.b),  |b'|^2 = |b|^2 + t (a.b)), so a rotation needs one dot product instead of three.
template <int NR>
DEVFN double jacobi_pair(double *AB_RESTRICT a, double *AB_RESTRICT b, int n, double &na2, double &nb2) {
  double al = na2, be = nb2, ga = 0.0;
  if (NR > 0) {
    double ra[NR > 0 ? NR : 1], rb[NR > 0 ? NR : 1];
    const bool full = (n == NR);                   // warp-uniform: no row masking needed
    if (full) {
#pragma unroll
      for (int r = 0; r < NR; r++) { ra[r] = a[r]; rb[r] = b[r]; }
    } else {
#pragma unroll
      for (int r = 0; r < NR; r++) { ra[r] = r < n ? a[r] : 0.0; rb[r] = r < n ? b[r] : 0.0; }
    }
    // eight partial sums: independent FMA chains
    double gap[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < NR; r++) gap[r & 7] += ra[r] * rb[r];
    ga = ((gap[0] + gap[1]) + (gap[2] + gap[3])) + ((gap[4] + gap[5]) + (gap[6] + gap[7]));
    // rel = |cos(a,b)|; the rotation uses  t = 2 ga / (d + sign(d) sqrt(d^2 + 4 ga^2)),  d = be - al  (one sqrt, one
    // divide and one rsqrt on the dependent chain; rel itself is off the chain)
    const double rel = fabs(ga) * rsqrt(al * be + 1.0e-300);
    if (rel > 1.0e-15) {
      const double d = be - al, g2 = 2.0 * ga;
      const double rt = sqrt(d * d + g2 * g2);
      const double t = g2 / (d >= 0.0 ? d + rt : d - rt);
      const double c = rsqrt(1.0 + t * t), s = c * t;
      na2 = al - t * ga; nb2 = be + t * ga;
      if (full) {
#pragma unroll
        for (int r = 0; r < NR; r++) { a[r] = c * ra[r] - s * rb[r]; b[r] = s * ra[r] + c * rb[r]; }
      } else {
#pragma unroll
        for (int r = 0; r < NR; r++)
          if (r < n) { a[r] = c * ra[r] - s * rb[r]; b[r] = s * ra[r] + c * rb[r]; }
      }
    }
    return rel;
  } else {
    for (int r = 0; r < n; r++) ga += a[r] * b[r];
    const double rel = fabs(ga) / sqrt(al * be + 1.0e-300);
    if (rel > 1.0e-15) {
      const double ze = (be - al) / (2.0 * ga);
      const double t = (ze >= 0.0 ? 1.0 : -1.0) / (fabs(ze) + sqrt(1.0 + ze * ze));
      const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
      na2 = al - t * ga; nb2 = be + t * ga;
      for (int r = 0; r < n; r++) { const double x = a[r], y = b[r]; a[r] = c * x - s * y; b[r] = s * x + c * y; }
    }
    return rel;
  }
}

#ifndef AB_HOST_MODEL
// the same rotation with the rows of the pair split over two lanes 16 apart (h = 0: rows [0, nh), h = 1: rows [nh, n)):
// half the registers and half the dependent work per lane; the dot product is completed with one shuffle.  NRH >= rows per lane.
template <int NRH>
DEVFN double jacobi_halfpair(double *AB_RESTRICT a, double *AB_RESTRICT b, int r0, int nr, double &na2, double &nb2, bool writer, unsigned mask) {
  double ra[NRH], rb[NRH];
#pragma unroll
  for (int r = 0; r < NRH; r++) { ra[r] = r < nr ? a[r0 + r] : 0.0; rb[r] = r < nr ? b[r0 + r] : 0.0; }
  double gap[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int r = 0; r < NRH; r++) gap[r & 3] += ra[r] * rb[r];
  double ga = (gap[0] + gap[1]) + (gap[2] + gap[3]);
  ga += __shfl_xor_sync(mask, ga, 16);
  const double al = na2, be = nb2;
  const double rel = fabs(ga) * rsqrt(al * be + 1.0e-300);
  if (rel > 1.0e-15) {
    const double d = be - al, g2 = 2.0 * ga;
    const double rt = sqrt(d * d + g2 * g2);
    const double t = g2 / (d >= 0.0 ? d + rt : d - rt);
    const double c = rsqrt(1.0 + t * t), s = c * t;
    if (writer) { na2 = al - t * ga; nb2 = be + t * ga; }
#pragma unroll
    for (int r = 0; r < NRH; r++)
      if (r < nr) { a[r0 + r] = c * ra[r] - s * rb[r]; b[r0 + r] = s * ra[r] + c * rb[r]; }
  }
  return rel;
}
// Jacobi rounds of a two-warp team with a RESIDENT column.  The pair schedule (jacobi_schedule below, built on the host) is a
// recursive "fixed / moving" tournament: the columns are split in halves F and M, slot t keeps F[t] while the columns of M pass by
// (|M| rounds: every F-M pair), then F and M are scheduled the same way side by side.  Every pair of columns meets exactly once per
// sweep (n (n-1)/2 rotations in 38 rounds at n = 36 against the 35 of the circle method; the same sweep counts), but the fixed column
// of a slot stays in the registers of its two lanes for a whole phase: per rotation ONE column is loaded from and stored to shared
// memory instead of two, which is what binds this kernel (column traffic of the rotations = 3/4 of its shared-memory wavefronts).
// Entry: a | b << 6 | load_a << 12 | store_a << 13, 0xFFFF = slot idle in this round.
// LP lanes per pair: 2 (row halves on lanes 16 apart) or 3 (row thirds on neighbouring lanes: 27 busy lanes per warp at n = 36, a
// third fewer instructions per rotation); NRH >= rows per lane.
template <int NRH, int LP, class Team>
DEVFN void jacobi_duo_resident(const Team &T, double *AB_RESTRICT A, int LDA, int n, double *AB_RESTRICT nrm2,
                               const unsigned short *AB_RESTRICT sch, int nrounds) {
  const int np = n + (n & 1), half = np / 2;
  const int ppw = (half + 1) / 2, w = T.rank >> 5, lane = T.rank & 31;
  const int h = LP == 2 ? lane >> 4 : lane % 3, tl = LP == 2 ? (lane & 15) : lane / 3, lbase = lane - h;
  const int t = w * ppw + tl;
  const bool act = tl < ppw && t < half;
  const int nseg = (n + LP - 1) / LP, r0 = h * nseg, nr = (n - r0 < nseg) ? (n - r0 > 0 ? n - r0 : 0) : nseg;
  (void)lbase;
  double ra[NRH], na2 = 0.0;
#pragma unroll
  for (int r = 0; r < NRH; r++) ra[r] = 0.0;
  for (int sweep = 0; sweep < 30; sweep++) {
    for (int c = T.rank; c < n; c += T.size) {               // exact squared norms at the start of every sweep
      const double *col = A + c * LDA;
      double s0 = 0.0, s1 = 0.0;
      for (int r = 0; r + 1 < n; r += 2) { s0 += col[r] * col[r]; s1 += col[r + 1] * col[r + 1]; }
      if (n & 1) s0 += col[n - 1] * col[n - 1];
      nrm2[c] = s0 + s1;
    }
    T.sync();
    double offmax = 0.0;
    unsigned e = act ? (unsigned)sch[t] : 0xFFFFu;
    for (int rd = 0; rd < nrounds; rd++) {
      const unsigned en = (act && rd + 1 < nrounds) ? (unsigned)sch[(rd + 1) * half + t] : 0xFFFFu;     // next round's entry, off the chain
      const bool run = e != 0xFFFFu;
      const unsigned rmask = __ballot_sync(0xffffffffu, run);
      if (run) {
        const int a = e & 63, b = (e >> 6) & 63;
        double *pa = A + a * LDA + r0, *pb = A + b * LDA + r0;
        if (e & 0x1000u) {
#pragma unroll
          for (int r = 0; r < NRH; r++) ra[r] = r < nr ? pa[r] : 0.0;
          na2 = nrm2[a];
        }
        double rb[NRH];
#pragma unroll
        for (int r = 0; r < NRH; r++) rb[r] = r < nr ? pb[r] : 0.0;
        double nb2 = nrm2[b];
        double gap[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int r = 0; r < NRH; r++) gap[r & 3] += ra[r] * rb[r];
        double ga = (gap[0] + gap[1]) + (gap[2] + gap[3]);
        if (LP == 2) ga += __shfl_xor_sync(rmask, ga, 16);
        else ga = (__shfl_sync(rmask, ga, lbase) + __shfl_sync(rmask, ga, lbase + 1)) + __shfl_sync(rmask, ga, lbase + 2);   // same order on all three lanes
        const double al = na2, be = nb2;
        // rel = |cos(a,b)| only feeds the convergence test; the rotation is gated by the equivalent product form and its parameters
        // come from two reciprocal square roots on the dependent chain (the textbook t = g2 / (d + sign(d) sqrt(d^2 + g2^2)),
        // c = 1 / sqrt(1 + t^2) has a square root, a division and a reciprocal square root in sequence):
        //   rho^-1 = rsqrt(d^2 + g2^2),  cos 2theta = |d| / rho,  c = cos theta = sqrt((1 + cos 2theta) / 2),  s = sign(d) g2 / (2 rho c)
        const double ab = al * be;
        const double rel = fabs(ga) * rsqrt(ab + 1.0e-300);
        if (rel > offmax) offmax = rel;
        const bool rot = ga * ga > 1.0e-30 * ab;
        if (rot) {
          const double d = be - al, g2 = 2.0 * ga;
          const double inv = rsqrt(d * d + g2 * g2);
          const double x = 0.5 + 0.5 * (fabs(d) * inv);
          const double ic = rsqrt(x);
          const double c = x * ic, sn = (d >= 0.0 ? 0.5 : -0.5) * (g2 * inv) * ic;
          const double tt = sn * ic;
          na2 = al - tt * ga; nb2 = be + tt * ga;
#pragma unroll
          for (int r = 0; r < NRH; r++) {
            const double ua = ra[r], ub = rb[r];
            ra[r] = c * ua - sn * ub;
            if (r < nr) pb[r] = sn * ua + c * ub;
          }
        }
        __syncwarp(rmask);                                   // (all running lanes) the partner lane has read the old norm of b
        if (rot && h == 0) nrm2[b] = nb2;
        if (e & 0x2000u) {
#pragma unroll
          for (int r = 0; r < NRH; r++) if (r < nr) pa[r] = ra[r];
          if (h == 0) nrm2[a] = na2;
        }
      }
      if (sch[nrounds * half + rd]) T.sync(); else __syncwarp();   // team barrier only where a column changes warps (host flag)
      e = en;
    }
    int dummy = 0;
    offmax = T.argmax(offmax, dummy);
    // quadratic convergence: a sweep whose largest pre-rotation |cos(a_p,a_q)| was below 1e-8 leaves ~1e-16 behind
    if (offmax < 1.0e-8) break;
  }
}
#endif

// back-transformation of eigenvector column e:  y = C u (in place in the column),  x = C^-T u (returned through xg)
template <int NR>
DEVFN void eig_backtransform(const double *AB_RESTRICT B1, int LDB, const double *AB_RESTRICT rdiag, double *AB_RESTRICT col, double isig, int n,
                             double *AB_RESTRICT xg) {
  if (NR > 0) {
    double u[NR > 0 ? NR : 1];
#pragma unroll
    for (int r = 0; r < NR; r++) u[r] = r < n ? col[r] * isig : 0.0;
    // y = C u: rows are independent; two partial sums per row halve the dependent chain
#pragma unroll
    for (int r = 0; r < NR; r++)
      if (r < n) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int t = 0; t <= r; t += 2) { s0 += B1[r * LDB + t] * u[t]; if (t + 1 <= r) s1 += B1[r * LDB + t + 1] * u[t + 1]; }
        col[r] = s0 + s1;
      }
    // x = C^-T u: back substitution, the reciprocal diagonal comes from rdiag (no division on the chain)
#pragma unroll
    for (int r = NR - 1; r >= 0; r--)
      if (r < n) {
        double s0 = u[r], s1 = 0.0;
#pragma unroll
        for (int t = r + 1; t < NR; t += 2) {
          if (t < n) s0 -= B1[t * LDB + r] * u[t];
          if (t + 1 < NR && t + 1 < n) s1 -= B1[(t + 1) * LDB + r] * u[t + 1];
        }
        u[r] = (s0 + s1) * rdiag[r];
      }
#pragma unroll
    for (int r = 0; r < NR; r++) if (r < n) xg[r] = u[r];
  } else {
    for (int r = 0; r < n; r++) xg[r] = col[r] * isig;
    for (int r = n - 1; r >= 0; r--) {
      double s = 0.0;
      for (int t = 0; t <= r; t++) s += B1[r * LDB + t] * xg[t];
      col[r] = s;
    }
    for (int r = n - 1; r >= 0; r--) {
      double s = xg[r];
      for (int t = r + 1; t < n; t++) s -= B1[t * LDB + r] * xg[t];
      xg[r] = s * rdiag[r];
    }
  }
}

template <int S, class Team>
DEVFN void eigen_problem(const Team &T, const VLPix &X, const VLSol &Z, const VLQuad &QD, const double *AB_RESTRICT pq, int m,
                         int p, int k, double *scr) {
  const int N = QD.N, n = S * N, NM = X.NM, L = X.L, LDB = n + 1, LDA = eig_lda(n);
  const size_t q = (size_t)p * L + k;
  // B1: Q / C in the lower triangle (r,c<=r) at [r*LDB+c];  P / Lp (r,c<=r) transposed above it at [c*LDB+r+1]
  double *B1 = scr, *A = B1 + n * LDB;
  double *sig = A + n * LDA, *sr = sig + n, *dpi = sr + n, *dmi = dpi + n, *gup = dmi + n, *gdn = gup + n;
  double *KK = gdn + n;               // 12n: user-stream operators [(2*ud+pn)*S+o][r]
  double *piu_s = KK + 12 * n;        // 3*(NM+1)  user-stream P,R,T (when not precomputed)
  double *pi0_s = piu_s + 3 * (NM + 1);   // 3*(NM+1)  solar P,R,T (when not precomputed)
  const double *pi0 = Z.pi0_g ? Z.pi0_g + (size_t)p * 3 * (NM + 1) : pi0_s;
  const double *og = og_ptr(X, p);                          // [(l*L + k)*4]
  const int G = X.G;
  const double *sc = X.scal + (size_t)view0(X, p) * 16;     // sun / flux: the same for every view of the pixel
  const int *ib = X.ibuf + (size_t)view0(X, p) * 4;
  const double FLUXF = sc[5];
  const bool layscat = X.layscat[((size_t)p * (NM + 1) + m) * L + k] != 0;
  const double deltau = X.deltau[pixL(X, p) + k];
  double *XPg = Z.xp + q * n * n, *XNg = Z.xn + q * n * n;

  // ---- generalized spherical functions at the solar cosine (the user cosines follow per view further down)
  if (!Z.pi0_g && T.rank == (T.size > 1 ? 1 : 0)) pi_prt(m, NM, sc[0], pi0_s);
  for (int r = T.rank; r < n; r += T.size) { const int i = r / S; sr[r] = 1.0 / sqrt(QD.w[i] * QD.x[i]); }      // sr holds 1 / sqrt(w mu)

  // ---- lower triangles of the symmetric matrices Q (from Zs) and P (from Zd); item = stream pair i >= j
  for (int idx = T.rank; idx < N * (N + 1) / 2; idx += T.size) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= idx) i++;
    const int j = idx - i * (i + 1) / 2;
    double zs[S * S], zd[S * S];
#pragma unroll
    for (int t = 0; t < S * S; t++) { zs[t] = 0.0; zd[t] = 0.0; }
    for (int l = m; l <= NM; l++) {
      const double *o4 = og + ((size_t)l * L + k) * 4;
      const double *pi_ = pq + ((size_t)l * N + i) * 3, *pj_ = pq + ((size_t)l * N + j) * 3;
      const bool even = ((l - m) & 1) == 0;                                // I +- ps*D on the {I,Q} block; swapped on U
      if (S == 1) {
        const double t = 2.0 * o4[0] * pi_[0] * pj_[0];
        if (even) zs[0] += t; else zd[0] += t;
      } else {
        const double Pi = pi_[0], Ri = pi_[1], Ti = pi_[2], Pj = pj_[0], Rj = pj_[1], Tj = pj_[2];
        const double b11 = 2.0 * o4[0], b12 = 2.0 * o4[1], b22 = 2.0 * o4[2], b33 = 2.0 * o4[3];
        // Pi_i E B Pi_j with B = [[b11,b12,0],[b12,b22,0],[0,0,b33]], Pi = [[P,0,0],[0,R,T],[0,T,R]], E = diag(e1,e1,e2)
        const double a00 = Pi * b11 * Pj, a01 = Pi * b12 * Rj, a02 = Pi * b12 * Tj;
        const double u10 = Ri * b12 * Pj, u11 = Ri * b22 * Rj, u12 = Ri * b22 * Tj, v11 = Ti * b33 * Tj, v12 = Ti * b33 * Rj;
        const double u20 = Ti * b12 * Pj, u21 = Ti * b22 * Rj, u22 = Ti * b22 * Tj, w21 = Ri * b33 * Tj, w22 = Ri * b33 * Rj;
        double *z1 = even ? zs : zd, *z2 = even ? zd : zs;                 // e1 part / e2 part
        z1[0] += a00; z1[1] += a01; z1[2] += a02; z1[3] += u10; z1[4] += u11; z1[5] += u12; z1[6] += u20; z1[7] += u21; z1[8] += u22;
        z2[4] += v11; z2[5] += v12; z2[7] += w21; z2[8] += w22;
      }
    }
    const double hw = 0.5 * sqrt(QD.w[i] * QD.w[j]), im = 1.0 / sqrt(QD.x[i] * QD.x[j]);
#pragma unroll
    for (int o1 = 0; o1 < S; o1++)
#pragma unroll
      for (int o2 = 0; o2 < S; o2++) {
        const int r = S * i + o1, c = S * j + o2;
        if (c <= r) {
          const double dlt = (r == c) ? 1.0 : 0.0;
          B1[r * LDB + c] = (dlt - hw * zs[o1 * S + o2]) * im;
          B1[c * LDB + r + 1] = (dlt - hw * zd[o1 * S + o2]) * im;
        }
      }
  }
  T.sync();
  // ---- Cholesky Q = C C^T and P = Lp Lp^T together, left-looking, in place; item = (matrix, row).  The columns stay UNSCALED during
  //      the factorisation (w(r,t) = L(r,t) sqrt(d_t), d_t = the pivot): s = a(r,c) - sum_t w(r,t) w(c,t) / d_t, so a column needs one
  //      team barrier instead of three and no reciprocal square root on the chain; one parallel pass scales everything at the end
  bool ok = true;
  double *idq = gup, *idp = gdn;                           // reciprocal pivots of Q and P (both vectors are only written further down)
  for (int c = 0; c < n; c++) {
    const int rem = n - c;
    for (int idx = T.rank; idx < 2 * rem; idx += T.size) {
      const int mat = idx >= rem, r = c + idx - mat * rem;
      double s0, s1 = 0.0;
      if (!mat) {
        const double *lr = B1 + r * LDB, *lc = B1 + c * LDB;
        s0 = lr[c];
        int t = 0;
        for (; t + 1 < c; t += 2) { s0 -= lr[t] * lc[t] * idq[t]; s1 -= lr[t + 1] * lc[t + 1] * idq[t + 1]; }
        if (t < c) s0 -= lr[t] * lc[t] * idq[t];
        s0 += s1;
        B1[r * LDB + c] = s0;
        if (r == c) idq[c] = 1.0 / s0;
      } else {
        s0 = B1[c * LDB + r + 1];
        int t = 0;
        for (; t + 1 < c; t += 2) {
          s0 -= B1[t * LDB + r + 1] * B1[t * LDB + c + 1] * idp[t]; s1 -= B1[(t + 1) * LDB + r + 1] * B1[(t + 1) * LDB + c + 1] * idp[t + 1];
        }
        if (t < c) s0 -= B1[t * LDB + r + 1] * B1[t * LDB + c + 1] * idp[t];
        s0 += s1;
        B1[c * LDB + r + 1] = s0;
        if (r == c) idp[c] = 1.0 / s0;
      }
    }
    T.sync();
    const double dq = B1[c * LDB + c], dp = B1[c * LDB + c + 1];
    if (!(dq > 0.0) || !(dp > 0.0)) { ok = false; break; }
  }
  if (!ok) { if (T.rank == 0) Z.fail[p] = 1; return; }
  // L(r,t) = w(r,t) / sqrt(d_t) (the diagonal becomes sqrt(d_t)): B1(i,j) with j <= i belongs to C (column j), with j > i to Lp (column i)
  for (int t = T.rank; t < n; t += T.size) { idq[t] = sqrt(idq[t]); idp[t] = sqrt(idp[t]); }
  T.sync();
  for (int idx = T.rank; idx < n * LDB; idx += T.size) {
    const int i = idx / LDB, j = idx - i * LDB;
    B1[idx] *= (j <= i) ? idq[j] : idp[i];
  }
  T.sync();
  // ---- A = C^T Lp  (column-major, A[c*LDA + r])
  {
    int r = T.rank, c = 0;
    while (r >= n) { r -= n; c++; }
    while (c < n) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int t = (r > c ? r : c);
      for (; t + 3 < n; t += 4) {
        s0 += B1[t * LDB + r] * B1[c * LDB + t + 1]; s1 += B1[(t + 1) * LDB + r] * B1[c * LDB + t + 2];
        s2 += B1[(t + 2) * LDB + r] * B1[c * LDB + t + 3]; s3 += B1[(t + 3) * LDB + r] * B1[c * LDB + t + 4];
      }
      for (; t < n; t++) s0 += B1[t * LDB + r] * B1[c * LDB + t + 1];
      A[c * LDA + r] = (s0 + s1) + (s2 + s3);
      r += T.size;
      while (r >= n) { r -= n; c++; }
    }
  }
  T.sync();
#ifdef AB_HM_SWEEP_HIST
  if (getenv("AB_HM_DUMP") && (k % 12 == 0)) { FILE *f = fopen(getenv("AB_HM_DUMP"), "ab"); int hd[4] = {n, m, k, LDA}; fwrite(hd, 4, 4, f); fwrite(A, 8, (size_t)n * LDA, f); fclose(f); }
#endif
  // ---- one-sided Jacobi
  bool jacobi_done = false;
#if !defined(AB_HOST_MODEL) && !defined(AB200_JACOBI_PAIRS)
  if constexpr (Team::size == 64) {
    // two-warp teams: row halves of a pair on two lanes, the slot's fixed column resident in registers (schedule from the host)
    if (QD.jsched != nullptr) {
      if (n <= 24) jacobi_duo_resident<8, 3>(T, A, LDA, n, sig, QD.jsched, QD.jrounds);
      else if (Team::maxhalf <= 18 || n <= 36) jacobi_duo_resident<12, 3>(T, A, LDA, n, sig, QD.jsched, QD.jrounds);
      else jacobi_duo_resident<(Team::maxhalf > 18 ? Team::maxhalf : 18), 2>(T, A, LDA, n, sig, QD.jsched, QD.jrounds);
      jacobi_done = true;
    }
  }
#endif
  // circle-method ordering, one column pair per lane (sub-warp and one-warp teams, the host model, AB200_JACOBI_PAIRS builds)
  if (!jacobi_done) {
    const int np = n + (n & 1), half = np / 2;
    double *nrm2 = sig;                                     // squared column norms (sig is only written after the sweeps)
    for (int sweep = 0; sweep < 30; sweep++) {
      double offmax = 0.0;
      for (int c = T.rank; c < n; c += T.size) {
        const double *col = A + c * LDA;
        double s0 = 0.0, s1 = 0.0;
        for (int r = 0; r + 1 < n; r += 2) { s0 += col[r] * col[r]; s1 += col[r + 1] * col[r + 1]; }
        if (n & 1) s0 += col[n - 1] * col[n - 1];
        nrm2[c] = s0 + s1;
      }
      T.sync();
      for (int st = 0; st < np - 1; st++) {
#ifndef AB_HOST_MODEL
        if constexpr (Team::size == 64) {
          // two warps per problem: warp w takes the pairs [w ppw, (w+1) ppw), two lanes per pair (row halves).  The two halves of a
          // pair sit 16 lanes apart: a half-warp then only holds equal halves of consecutive columns, whose 8-byte words fall
          // into distinct banks (odd LDA) -- with neighbouring lanes a third of the column traffic was bank conflicts
          const int ppw = (half + 1) / 2, w = T.rank >> 5, lane = T.rank & 31, h = lane >> 4;
          const int t = w * ppw + (lane & 15);
          const bool act = (lane & 15) < ppw && t < half;
          int a = 0, b = 0;
          if (act) { if (t == 0) { a = np - 1; b = st; } else { a = st + t; if (a >= np - 1) a -= np - 1; b = st - t; if (b < 0) b += np - 1; } }
          const bool run = act && a < n && b < n;
          const unsigned rmask = __ballot_sync(0xffffffffu, run);
          const int nh = (n + 1) / 2, r0 = h ? nh : 0, nr = h ? n - nh : nh;
          double rel = 0.0;
          // (the shuffle inside needs both lanes of a pair: pairs are either fully active or fully idle)
          if (run) {
            double na2 = nrm2[a], nb2 = nrm2[b];
            if (n <= 12) rel = jacobi_halfpair<6>(A + a * LDA, A + b * LDA, r0, nr, na2, nb2, h == 0, rmask);
            else if (n <= 24) rel = jacobi_halfpair<12>(A + a * LDA, A + b * LDA, r0, nr, na2, nb2, h == 0, rmask);
            else if (Team::maxhalf <= 18 || n <= 36) rel = jacobi_halfpair<18>(A + a * LDA, A + b * LDA, r0, nr, na2, nb2, h == 0, rmask);
            else rel = jacobi_halfpair<(Team::maxhalf > 18 ? Team::maxhalf : 18)>(A + a * LDA, A + b * LDA, r0, nr, na2, nb2, h == 0, rmask);
            __syncwarp(rmask);                                 // the partner lane has read the old norms (racecheck-clean ordering)
            if (h == 0) { nrm2[a] = na2; nrm2[b] = nb2; }
          }
          if (rel > offmax) offmax = rel;
          T.sync();
          continue;
        }
#endif
        for (int t = T.rank; t < half; t += T.size) {
          int a, b;
          if (t == 0) { a = np - 1; b = st; } else { a = st + t; if (a >= np - 1) a -= np - 1; b = st - t; if (b < 0) b += np - 1; }
          if (a < n && b < n) {
            double rel;
            if (n <= 12) rel = jacobi_pair<12>(A + a * LDA, A + b * LDA, n, nrm2[a], nrm2[b]);
            else if (n <= 24) rel = jacobi_pair<24>(A + a * LDA, A + b * LDA, n, nrm2[a], nrm2[b]);
            else if (n <= 36) rel = jacobi_pair<36>(A + a * LDA, A + b * LDA, n, nrm2[a], nrm2[b]);
            else rel = jacobi_pair<0>(A + a * LDA, A + b * LDA, n, nrm2[a], nrm2[b]);
            if (rel > offmax) offmax = rel;
          }
        }
        T.sync();
      }
      int dummy = 0;
      offmax = T.argmax(offmax, dummy);
      // quadratic convergence: a sweep whose largest pre-rotation |cos(a_p,a_q)| was below 1e-8 leaves ~1e-16 behind
#ifdef AB_HM_SWEEP_HIST
      hm_sweep_last = sweep + 1; hm_off_hist[sweep < 31 ? sweep : 31] = offmax;
#endif
      if (offmax < 1.0e-8) break;
    }
#ifdef AB_HM_SWEEP_HIST
    hm_sweep_hist[m < 15 ? m : 15][hm_sweep_last < 31 ? hm_sweep_last : 31]++;
#endif
  }
  // ---- singular values = eigenvalues k of the layer; y = C u in place in A, x = C^-T u through global memory (XNg)
  bool bad = false;
  const double *rdiag = gup;                               // 1 / C(r,r) = 1 / sqrt(pivot), left there by the Cholesky step
  for (int e = T.rank; e < n; e += T.size) {
    double *col = A + e * LDA;
    double s2a = 0.0, s2b = 0.0;
    for (int r = 0; r + 1 < n; r += 2) { s2a += col[r] * col[r]; s2b += col[r + 1] * col[r + 1]; }
    if (n & 1) s2a += col[n - 1] * col[n - 1];
    const double ke = sqrt(s2a + s2b);
    sig[e] = ke;
    if (!(ke > 0.0)) { bad = true; continue; }
    double *xg = XNg + (size_t)e * n;
    if (n <= 12) eig_backtransform<12>(B1, LDB, rdiag, col, 1.0 / ke, n, xg);
    else if (n <= 24) eig_backtransform<24>(B1, LDB, rdiag, col, 1.0 / ke, n, xg);
    else if (n <= 36) eig_backtransform<36>(B1, LDB, rdiag, col, 1.0 / ke, n, xg);
    else eig_backtransform<0>(B1, LDB, rdiag, col, 1.0 / ke, n, xg);
  }
  if (bad) Z.fail[p] = 1;
  T.sync();                                               // C is dead from here on: B1 becomes XNEG [e*LDB + r]
  // ---- X+ / X- (streams 1..N), whole-layer transmittance, multipliers
  for (int e = T.rank; e < n; e += T.size) {
    const double ke = sig[e];
    if (!(ke > 0.0)) continue;
    double *col = A + e * LDA, *xn = B1 + e * LDB;
    const double *xg = XNg + (size_t)e * n;
    double nrm = 0.0, nrm1 = 0.0;
    for (int r = 0; r + 1 < n; r += 2) { const double s = xg[r] * sr[r], s1 = xg[r + 1] * sr[r + 1]; nrm += s * s; nrm1 += s1 * s1; }
    if (n & 1) { const double s = xg[n - 1] * sr[n - 1]; nrm += s * s; }
    const double inr = 1.0 / sqrt(nrm + nrm1), ikr = inr / ke;
    for (int r = 0; r < n; r++) {
      const double sg = xg[r] * sr[r] * inr, df = col[r] * sr[r] * ikr;
      col[r] = 0.5 * (sg + df);            // XPOS(i,o,e)
      xn[r] = 0.5 * (sg - df);             // XNEG(i,o,e)
    }
    Z.keig[q * n + e] = ke;
    const double hlp = ke * deltau, tde = hlp > MAX_TAU_PATH ? 0.0 : exp(-hlp);
    Z.tde[q * n + e] = tde;
  }
  T.sync();
  for (int idx = T.rank; idx < n * n; idx += T.size) { XPg[idx] = A[(idx / n) * LDA + (idx % n)]; XNg[idx] = B1[(idx / n) * LDB + (idx % n)]; }
  // ---- multi-geometry calls: user-stream solutions through per-eigenvector moment sums (the reference's HELP_QFUNC formulation,
  //      VM/vlidort_solutions.f90:1167-1508).  With H(l,j) = a_j [[P,0,0],[0,R,T],[0,T,R]](l, mu_j), H' = diag(1,1,-1) H and B_l the Greek
  //      block of the layer, z_l(e) = B_l sum_j [H(l,j) X+(j,e) + (-1)^(l-m) H'(l,j) X-(j,e)] does not depend on the viewing direction;
  //      a view then costs ~10 flops per (l, e) instead of building its 12 n operator matrix and multiplying it with X+, X-.
  if (G > 1 && Z.zws != nullptr) {
    double *zw = Z.zws + q * (size_t)(NM + 1) * n * 3;
    const int nl = NM + 1 - m;
    for (int idx = T.rank; idx < nl * n; idx += T.size) {
      const int l = m + idx / n, e = idx % n;
      double w0 = 0.0, w1 = 0.0, w2 = 0.0;
      if (layscat && sig[e] > 0.0) {
        const double ps = ((l - m) & 1) ? -1.0 : 1.0;
        const double *xp_ = A + e * LDA, *xn_ = B1 + e * LDB;
        for (int j = 0; j < N; j++) {
          const double a5 = 0.5 * QD.w[j];
          const double *pj_ = pq + ((size_t)l * N + j) * 3;
          if (S == 1) { w0 += a5 * pj_[0] * (xp_[j] + ps * xn_[j]); }
          else {
            const double Pj = a5 * pj_[0], Rj = a5 * pj_[1], Tj = a5 * pj_[2];
            const double x0 = xp_[S * j], x1 = xp_[S * j + 1], x2 = xp_[S * j + 2], y0 = xn_[S * j], y1 = xn_[S * j + 1], y2 = xn_[S * j + 2];
            w0 += Pj * (x0 + ps * y0);
            w1 += Rj * (x1 + ps * y1) + Tj * (x2 + ps * y2);
            w2 += Tj * (x1 - ps * y1) + Rj * (x2 - ps * y2);
          }
        }
      }
      const double *o4 = og + ((size_t)l * L + k) * 4;
      double *z = zw + ((size_t)(l - m) * n + e) * 3;
      if (S == 1) { z[0] = o4[0] * w0; z[1] = 0.0; z[2] = 0.0; }
      else { z[0] = o4[0] * w0 + o4[1] * w1; z[1] = o4[1] * w0 + o4[2] * w1; z[2] = o4[3] * w2; }
    }
    T.sync();
    for (int e = T.rank; e < n; e += T.size) {
      const double ke = sig[e];
      const double hlp = ke * deltau, tde = hlp > MAX_TAU_PATH ? 0.0 : exp(-hlp);
      for (int g = 0; g < G; g++) {
        const size_t qv = q * G + g;
        const size_t v = (size_t)view0(X, p) + g;
        if (ke > 0.0) {
          const double *scv = X.scal + v * 16;
          const double usec = scv[2], udel = X.tduserm[pixL(X, v) + k];
          const double rho_p = usec + ke, rho_m = usec - ke;
          const bool hs = fabs(rho_m) < TAYLOR_SMALL;
          Z.hm1[qv * n + e] = hs ? taylor_series_1(rho_m, deltau, udel, usec) : usec * (tde - udel) / rho_m;
          Z.hm2[qv * n + e] = usec * (1.0 - tde * udel) / rho_p;
        }
        const double *piu = Z.piu_g ? Z.piu_g + v * 3 * (NM + 1) : nullptr;
        double dn[3] = {0.0, 0.0, 0.0}, up[3] = {0.0, 0.0, 0.0};
        if (piu != nullptr)
          for (int l = m; l <= NM; l++) {
            const double ps = ((l - m) & 1) ? -1.0 : 1.0;
            const double *z = zw + ((size_t)(l - m) * n + e) * 3, *pu = piu + 3 * l;
            const double Pu = pu[0], Ru = pu[1], Tu = pu[2], z0 = z[0], z1 = z[1], z2 = z[2];
            dn[0] += ps * Pu * z0; up[0] += Pu * z0;
            if (S > 1) {
              dn[1] += ps * (Ru * z1 - Tu * z2); dn[2] += ps * (Ru * z2 - Tu * z1);
              up[1] += Ru * z1 + Tu * z2;        up[2] += -(Tu * z1 + Ru * z2);
            }
          }
#pragma unroll
        for (int o = 0; o < S; o++) { Z.uh_dn[(qv * n + e) * S + o] = dn[o]; Z.uh_up[(qv * n + e) * S + o] = up[o]; }
      }
    }
    T.sync();
  } else
  // ---- per view geometry: whole-layer multipliers and user-stream solutions
  for (int g = 0; g < G; g++) {
  const size_t qv = q * G + g;
  const double *scv = X.scal + ((size_t)view0(X, p) + g) * 16;
  const double usec = scv[2], udel = X.tduserm[pixL(X, view0(X, p) + g) + k];
  // the same for all 72 layers of the pixel-view: precomputed per Fourier term (a serial recurrence with square roots and divisions)
  const double *piu = Z.piu_g ? Z.piu_g + ((size_t)view0(X, p) + g) * 3 * (NM + 1) : piu_s;
  if (!Z.piu_g && T.rank == 0) pi_prt(m, NM, scv[1], piu_s);
  for (int e = T.rank; e < n; e += T.size) {
    const double ke = sig[e];
    if (!(ke > 0.0)) continue;
    const double hlp = ke * deltau, tde = hlp > MAX_TAU_PATH ? 0.0 : exp(-hlp);
    // HMULT (whole layer, upwelling user stream): vlidort_solutions.f90:1715-1745, zeta :1290-1305
    const double rho_p = usec + ke, rho_m = usec - ke;
    const bool hs = fabs(rho_m) < TAYLOR_SMALL;
    Z.hm1[qv * n + e] = hs ? taylor_series_1(rho_m, deltau, udel, usec) : usec * (tde - udel) / rho_m;
    Z.hm2[qv * n + e] = usec * (1.0 - tde * udel) / rho_p;
  }
  T.sync();
  // ---- user-stream operators: UHOM_UP/DN(o,e) = sum_r KK[..][r] X+-(r,e); item = (stream j, up/down)
  for (int idx = T.rank; idx < 2 * N; idx += T.size) {
    const int j = idx >> 1, ud = idx & 1;                                  // ud = 0: UHOM_UPDN, 1: UHOM_UPUP
    double kp[S * S], kn[S * S];                                           // [o][c]: coefficient of X+-(S*j+c) in component o
#pragma unroll
    for (int t = 0; t < S * S; t++) { kp[t] = 0.0; kn[t] = 0.0; }
    if (layscat) {
      const double a5 = 0.5 * QD.w[j];
      for (int l = m; l <= NM; l++) {
        const double ps = ((l - m) & 1) ? -1.0 : 1.0, f = ud ? 1.0 : ps;
        const double *o4 = og + ((size_t)l * L + k) * 4;
        const double *pj_ = pq + ((size_t)l * N + j) * 3, *pu = piu + 3 * l;
        if (S == 1) {
          const double g = f * pu[0] * o4[0] * a5 * pj_[0];
          kp[0] += g; kn[0] += ps * g;
        } else {
          const double Pj = a5 * pj_[0], Rj = a5 * pj_[1], Tj = a5 * pj_[2];
          const double b11 = o4[0], b12 = o4[1], b22 = o4[2], b33 = o4[3], Pu = pu[0], Ru = pu[1], Tu = pu[2];
          // hs = H X+ + ps H' X-,  H = [[Pj,0,0],[0,Rj,Tj],[0,Tj,Rj]],  H' = diag(1,1,-1) H ;  g = B hs ;  out = Uud g
          // rows of Uud: dn: ps*[Pu,0,0],[0,Ru,-Tu],[0,-Tu,Ru] ;  up: [Pu,0,0],[0,Ru,Tu],[0,-Tu,-Ru]
          const double s3 = ud ? 1.0 : -1.0, s4 = ud ? -1.0 : 1.0;
          // M = Uud * B  (3x3)
          const double m00 = f * Pu * b11, m01 = f * Pu * b12;
          const double m10 = f * Ru * b12, m11 = f * Ru * b22, m12 = f * s3 * Tu * b33;
          const double m20 = -f * Tu * b12, m21 = -f * Tu * b22, m22 = f * s4 * Ru * b33;
          // times H (for X+) and H' (for X-, third row of H negated, times ps)
          const double h00 = m00 * Pj, h01 = m01 * Rj, h02 = m01 * Tj;
          const double h10 = m10 * Pj, h11 = m11 * Rj, h12 = m11 * Tj, g11 = m12 * Tj, g12 = m12 * Rj;
          const double h20 = m20 * Pj, h21 = m21 * Rj, h22 = m21 * Tj, g21 = m22 * Tj, g22 = m22 * Rj;
          kp[0] += h00; kp[1] += h01; kp[2] += h02;
          kp[3] += h10; kp[4] += h11 + g11; kp[5] += h12 + g12;
          kp[6] += h20; kp[7] += h21 + g21; kp[8] += h22 + g22;
          kn[0] += ps * h00; kn[1] += ps * h01; kn[2] += ps * h02;
          kn[3] += ps * h10; kn[4] += ps * (h11 - g11); kn[5] += ps * (h12 - g12);
          kn[6] += ps * h20; kn[7] += ps * (h21 - g21); kn[8] += ps * (h22 - g22);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < S; o++)
#pragma unroll
      for (int c = 0; c < S; c++) {
        KK[((2 * ud + 0) * S + o) * n + S * j + c] = kp[o * S + c];
        KK[((2 * ud + 1) * S + o) * n + S * j + c] = kn[o * S + c];
      }
  }
  T.sync();
  for (int e = T.rank; e < n; e += T.size) {
    const double *col = A + e * LDA, *xn = B1 + e * LDB;
    double dn[S], up[S];
#pragma unroll
    for (int o = 0; o < S; o++) { dn[o] = 0.0; up[o] = 0.0; }
    for (int r = 0; r < n; r++) {
      const double xp_ = col[r], xn_ = xn[r];
#pragma unroll
      for (int o = 0; o < S; o++) {
        dn[o] += KK[(0 * S + o) * n + r] * xp_ + KK[(1 * S + o) * n + r] * xn_;
        up[o] += KK[(2 * S + o) * n + r] * xp_ + KK[(3 * S + o) * n + r] * xn_;
      }
    }
#pragma unroll
    for (int o = 0; o < S; o++) { Z.uh_dn[(qv * n + e) * S + o] = dn[o]; Z.uh_up[(qv * n + e) * S + o] = up[o]; }
  }
  T.sync();
  }   // views
  // ---- Green's function particular integral (solar beam)
  const bool active = layscat && (k + 1 <= ib[0]);
  double *WU = Z.wupper + q * (2 * n), *WL = Z.wlower + q * (2 * n);
  if (!active) {
    for (int r = T.rank; r < 2 * n; r += T.size) { WU[r] = 0.0; WL[r] = 0.0; }
    for (int e = T.rank; e < n; e += T.size) { Z.aterm[q * n + e] = 0.0; Z.bterm[q * n + e] = 0.0; Z.gam_m[q * n + e] = 1.0; Z.gam_p[q * n + e] = 1.0; }
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
      const double *col = A + e * LDA, *xnc = B1 + e * LDB;
      double nrm = 0.0, s1 = 0.0, s2 = 0.0;
      for (int i = 0; i < N; i++) {
        double t1 = 0.0, t2 = 0.0, h1 = 0.0, h2 = 0.0;
#pragma unroll
        for (int o = 0; o < S; o++) {
          const double xp = col[S * i + o], xn = xnc[S * i + o];
          t1 += xp * xp; t2 += xn * xn;
          h1 += dpi[S * i + o] * xp + dmi[S * i + o] * xn;
          h2 += dmi[S * i + o] * xp + dpi[S * i + o] * xn;
        }
        nrm += QD.x[i] * QD.w[i] * (t1 - t2);
        s1 += QD.w[i] * h1; s2 += QD.w[i] * h2;
      }
      const double at = s1 / nrm, bt = s2 / nrm, ke = sig[e], zdel = ke * deltau > MAX_TAU_PATH ? 0.0 : exp(-ke * deltau);
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
      const double xp = A[e * LDA + r], xn = B1[e * LDB + r];
      a += gup[e] * xn; b += gup[e] * xp; c += gdn[e] * xn; d += gdn[e] * xp;
    }
    const double sg = o > 1 ? -1.0 : 1.0;
    WU[r] = a; WU[n + r] = sg * b; WL[n + r] = sg * c; WL[r] = d;
  }
}

}  // namespace vlk
