// Stage B, kernel K3, forward elimination for the GPU (BlockTeam of 384 threads) -- warp-specialised version.
// Same mathematics as bvp_pixel's forward phase in vl_bvp.cuh (window of rows eliminated layer by layer with partial
// pivoting, panels of 8 columns), organised around the measured critical path, the pivot chain of the panel:
//   * panel warp PW (warp 0): owns the chain.  Lane l owns the window rows l, l+32, ...; the 8 columns of the panel
//     live in registers, pivot search = integer redux + ballot, pivot row and reciprocal travel by shuffles, no
//     block barrier anywhere.  PW also applies the previous panel to its own 8 columns (left-looking by one panel),
//     so it never waits for the update warps unless they fall two panels behind.
//   * update warps UW (9 warps, three per remaining scheduler): pivot rows of the panel over the trailing columns
//     (triangular solve, one column per thread), then the trailing update W -= L U on 8x8 tiles with FP64 tensor-core
//     MMAs (mma.sync m8n8k4.f64), multipliers stored negated so that D = A B + C.
//   * loader warps LW (2 warps, sharing PW's scheduler): spill the pivot rows of the panel before last to HBM
//     (logical column order, for bvp_backward), recycle their slots and stage the rows of the next layer interface
//     in them while the current layer is still being eliminated.
// The two halves of the window (unknowns of layer k / layer k+1) swap roles from layer to layer instead of being
// shifted.  Named barriers 1..7 carry the producer/consumer hand-offs (bar.arrive on the producer side).
#pragma once
#include "vl_bvp.cuh"

#ifndef AB_HOST_MODEL
namespace vlk {

#if defined(AB200_BVP_TIMING)
#define VLT_RESET { vlt_t = clock64(); }
#else
#define VLT_RESET
#endif

constexpr int B3_THREADS = 384;
constexpr int B3_NUW = 9;          // update warps
constexpr int B3_UWT = 32 * B3_NUW;
constexpr int B3_LWT = 64;         // loader threads
constexpr int B3_LPS = 12;         // row stride of the multiplier panels (A fragments conflict-free)
enum { B3_PR0 = 1, B3_PR1 = 2, B3_TD0 = 3, B3_TD1 = 4, B3_UB = 5, B3_RR = 6, B3_LB = 7 };

HDFN int b3_rows(int n) { return (3 * n + 16 + 7) & ~7; }
HDFN int b3_cs(int n) { int c = 4 * n + 2; while ((c & 15) != 8) c++; return c; }       // 16-byte C pairs conflict-free
HDFN int b3_ubs(int n) { int c = 4 * n + 8; while ((c & 15) != 4) c++; return c; }      // B fragments conflict-free
HDFN bool b3_supported(int n) { return b3_rows(n) <= 128; }
// doubles of the forward phase: window | 2 multiplier panels | 2 U panels | 2 LU blocks | U block of the look-ahead
HDFN size_t b3_fwd_doubles(int n) {
  return (size_t)b3_rows(n) * b3_cs(n) + 2 * (size_t)b3_rows(n) * B3_LPS + 2 * (size_t)BVP_NB * b3_ubs(n) + 2 * 64 + 64;
}
HDFN size_t b3_big_doubles(int n) {
  const size_t fwd = b3_fwd_doubles(n), bwd = (size_t)2 * (2 * n) * bvp_cols(n);
  return ((fwd > bwd ? fwd : bwd) + 1) & ~(size_t)1;
}
HDFN size_t b3_smem_doubles(int n) { return b3_big_doubles(n) + 8 * n + 16; }
HDFN size_t b3_smem_ints(int n) { return (size_t)3 * b3_rows(n) + 64; }

DEVFN void b3_sync(int id, int cnt) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(cnt) : "memory"); }
DEVFN void b3_arrive(int id, int cnt) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(cnt) : "memory"); }
DEVFN void b3_mma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct B3Ctl {             // integer control block in shared memory
  int pivrow[4][BVP_NB];   // window slots of the pivot rows of the panel (ring of 4 over the global panel number: the loader
                           // still reads panel g-1 while the panel warp may already be writing panel g+2)
  unsigned rowbits[2][4];  // rows with non-zero multipliers in the panel
  unsigned tmask[2];       // 8-row tiles with at least one such row
  int iflag;               // a pivot search found no usable pivot
  int fqn, cur, nst, st_r0, pad_;   // loader: free slots, staged rows of the next layer, rows of this event
};

struct B3Ctx {
  int S, N, n, n2, L, RS, Cs, Cc, UbS, np;
  double *W, *Lp, *Ub, *LU, *Uloc;
  B3Ctl *ctl;
  int *rowst, *fq, *stslot;
  size_t q0;
  double *Ug;
};

// physical column of logical column c of the window of a layer with parity `odd`
DEVFN int b3_pc(int c, int n2, bool odd) { return (!odd || c >= 2 * n2) ? c : (c < n2 ? c + n2 : c - n2); }

// ---- rows of the interface below layer k (k < L-1) or the surface rows (k = L-1), computed by `nthr` threads into staged
// slots: rows r2 = r0 .. r0+cnt-1 -> slots slot[0..cnt).  Same expressions as bvp_pixel's "load the new rows".
template <int S, class FBF, class FBF0>
DEVFN void b3_load_rows(const B3Ctx &C, const VLSol &Z, const VLQuad &QD, int k, int r0, int cnt, const int *slot, int tid, int nthr,
                        bool include_surface, bool lambertian, bool localbeam, double SFAC, double albedo, double attn, int m, FBF BFq, FBF0 BF0q) {
  const int n = C.n, n2 = C.n2, N = C.N, Cs = C.Cs;
  const bool odd = k & 1;
  const size_t qa = C.q0 + k, qb = qa + 1;
  double *W = C.W;
  if (k < C.L - 1) {
    for (int idx = tid; idx < cnt * n2; idx += nthr) {
      const int rr = idx % cnt, ce = idx / cnt, r2 = r0 + rr;       // ce: column inside a half, e = ce % n
      const int e = ce < n ? ce : ce - n;
      // full-stream values XPOS/XNEG(r2, e) of layers qa, qb from the stored downwelling halves
      const int rlo = r2 < n ? r2 : r2 - n;
      const double sg = (r2 >= n && (rlo % S) > 1) ? -1.0 : 1.0;
      const double *pa = (r2 < n) ? Z.xp : Z.xn, *na = (r2 < n) ? Z.xn : Z.xp;
      double *row = W + (size_t)slot[rr] * Cs;
      if (ce < n) {
        const double xa = sg * pa[(qa * n + e) * n + rlo], xb = sg * pa[(qb * n + e) * n + rlo];
        row[b3_pc(e, n2, odd)] = Z.tde[qa * n + e] * xa;
        row[b3_pc(n2 + e, n2, odd)] = -xb;
      } else {
        const double xa = sg * na[(qa * n + e) * n + rlo], xb = sg * na[(qb * n + e) * n + rlo];
        row[b3_pc(n + e, n2, odd)] = xa;
        row[b3_pc(n2 + n + e, n2, odd)] = -Z.tde[qb * n + e] * xb;
      }
    }
    for (int rr = tid; rr < cnt; rr += nthr) {
      const int r2 = r0 + rr;
      W[(size_t)slot[rr] * Cs + 4 * n] = Z.wupper[qb * n2 + r2] - Z.wlower[qa * n2 + r2];
    }
  } else {
    for (int idx = tid; idx < cnt * n; idx += nthr) {
      const int rr = idx % cnt, e = idx / cnt, r = r0 + rr, i = r / S, o1 = r % S;
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
      // XPOS(i+N) = +-XNEG-half, XNEG(i+N) = +-XPOS-half
      const double sg = ((r % S) > 1) ? -1.0 : 1.0;
      const double xpu = sg * Z.xn[(qa * n + e) * n + r], xnu = sg * Z.xp[(qa * n + e) * n + r];
      double *row = W + (size_t)slot[rr] * Cs;
      row[b3_pc(e, n2, odd)] = Z.tde[qa * n + e] * (xpu - rp);
      row[b3_pc(n + e, n2, odd)] = xnu - rm;
      row[b3_pc(n2 + e, n2, odd)] = 0.0;
      row[b3_pc(n2 + n + e, n2, odd)] = 0.0;
    }
    for (int rr = tid; rr < cnt; rr += nthr) {
      const int r = r0 + rr, i = r / S, o1 = r % S;
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
          else v += attn * BF0q(o1, i);
        }
      }
      W[(size_t)slot[rr] * Cs + 4 * n] = v;
    }
  }
}

// ---- panel warp
template <int RU>
DEVFN void b3_panel_warp(const B3Ctx &C) {
  const int lane = threadIdx.x & 31;
  const int n2 = C.n2, L = C.L, RS = C.RS, Cs = C.Cs, np = C.np;
  double *AB_RESTRICT W = C.W;
  B3Ctl *ctl = C.ctl;
  unsigned actm = 0;                       // bit u: row lane + 32 u is an active pivot candidate
  bool failed = false;
  int g = 0;
  VLT_DECL
  for (int k = 0; k < L; k++) {
    const int offL = (k & 1) ? n2 : 0;     // physical start of the left half (the columns of layer k)
    for (int P = 0; P < np; P++, g++) {
      const int jc0 = P * BVP_NB, nb = (n2 - jc0 < BVP_NB) ? n2 - jc0 : BVP_NB;
      const int par = g & 1;
      if (P == 0) {
        b3_sync(B3_RR, 32 + B3_LWT);
#pragma unroll
        for (int u = 0; u < RU; u++) { const int r = lane + 32 * u; if (r < RS && C.rowst[r] == 2) { actm |= 1u << u; C.rowst[r] = 0; } }
        VLT(1)
      }
      if (g >= 2) b3_sync(par ? B3_TD1 : B3_TD0, 32 + B3_UWT);
      VLT(2)
      if (!failed) {
        double pv[RU][BVP_NB];
        const int pc0 = offL + jc0;
#pragma unroll
        for (int u = 0; u < RU; u++) {
          const int r = lane + 32 * u;
          const bool a = (actm >> u) & 1u;
          const double2 *src = reinterpret_cast<const double2 *>(W + (size_t)(r < RS ? r : 0) * Cs + pc0);
#pragma unroll
          for (int j2 = 0; j2 < BVP_NB / 2; j2++) {
            double2 v = make_double2(0.0, 0.0);
            if (a) v = src[j2];
            pv[u][2 * j2] = v.x; pv[u][2 * j2 + 1] = v.y;
          }
        }
        if (g > 0) {
          // the previous panel, applied to these columns: first its pivot rows (triangular solve, one column per lane) ...
          const int jcp = (P > 0) ? jc0 - BVP_NB : (np - 1) * BVP_NB, nbp = (P > 0) ? BVP_NB : n2 - jcp;
          (void)jcp;
          const double *LUp = C.LU + (par ^ 1) * 64;
          const int *pivp = ctl->pivrow[(g - 1) & 3];
          if (lane < BVP_NB) {
            double uv[BVP_NB];
#pragma unroll
            for (int j = 0; j < BVP_NB; j++) {
              double v = 0.0;
              if (j < nbp && lane < nb) {
                double *w = W + (size_t)pivp[j] * Cs + pc0 + lane;
                v = *w;
#pragma unroll
                for (int jj = 0; jj < BVP_NB; jj++) if (jj < j) v -= LUp[j * BVP_NB + jj] * uv[jj];
                *w = v;
              }
              uv[j] = v;
              C.Uloc[j * BVP_NB + lane] = v;
            }
          }
          __syncwarp();
          // ... then the multipliers of the previous panel (stored negated) times those rows
          const double *LpP = C.Lp + (size_t)(par ^ 1) * RS * B3_LPS;
#pragma unroll
          for (int u = 0; u < RU; u++) {
            const int r = lane + 32 * u;
            if ((actm >> u) & 1u) {
#pragma unroll
              for (int j = 0; j < BVP_NB; j++) {
                const double l = LpP[r * B3_LPS + j];
#pragma unroll
                for (int c2 = 0; c2 < BVP_NB / 2; c2++) {
                  const double2 uu = *reinterpret_cast<const double2 *>(C.Uloc + j * BVP_NB + 2 * c2);
                  pv[u][2 * c2] += l * uu.x; pv[u][2 * c2 + 1] += l * uu.y;
                }
              }
            }
          }
        }
        VLT(3)
        double *LpN = C.Lp + (size_t)par * RS * B3_LPS;
        unsigned nzm = 0;
#pragma unroll
        for (int j = 0; j < BVP_NB; j++) {
          if (j < nb && !failed) {
            // pivot search: largest |value| among the active rows, lowest slot on exact ties
            double best = -1.0, vb = 1.0; int bu = 0;
#pragma unroll
            for (int u = 0; u < RU; u++) { const double a = fabs(pv[u][j]); if (((actm >> u) & 1u) && a > best) { best = a; bu = u; vb = pv[u][j]; } }
            const double rloc = 1.0 / vb;                                        // off the chain: overlaps the reductions
            const unsigned long long key = best < 0.0 ? 0ull : (unsigned long long)__double_as_longlong(best);
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            bool cand = (hi == mh) && (best >= 0.0);
            unsigned bal = __ballot_sync(0xffffffffu, cand);
            if (__popc(bal) > 1) {
              const unsigned ml = __reduce_max_sync(0xffffffffu, cand ? lo : 0u);
              cand = cand && (lo == ml);
              bal = __ballot_sync(0xffffffffu, cand);
              if (__popc(bal) > 1) {
                const unsigned idx = cand ? (unsigned)(lane + 32 * bu) : 0x7fffffffu;
                const unsigned mi = __reduce_min_sync(0xffffffffu, idx);
                cand = cand && (idx == mi);
                bal = __ballot_sync(0xffffffffu, cand);
              }
            }
            const int owner = __ffs(bal) - 1;
            const double bestv = __shfl_sync(0xffffffffu, best, owner < 0 ? 0 : owner);
            if (owner < 0 || !(bestv > 0.0)) failed = true;
            else {
              const int ustar = __shfl_sync(0xffffffffu, bu, owner);             // uniform
              const bool mine = (lane == owner);
              double prow[BVP_NB];
#pragma unroll
              for (int jj = 0; jj < BVP_NB; jj++) prow[jj] = 0.0;
#pragma unroll
              for (int u = 0; u < RU; u++)
                if (u == ustar) {                                                // uniform branch: plain register copies
#pragma unroll
                  for (int jj = 0; jj < BVP_NB; jj++) prow[jj] = pv[u][jj];
                  if (mine) actm &= ~(1u << u);
                }
#pragma unroll
              for (int jj = 0; jj < BVP_NB; jj++) if (jj >= j) prow[jj] = __shfl_sync(0xffffffffu, prow[jj], owner);
              const double ipiv = __shfl_sync(0xffffffffu, rloc, owner);
              if (mine) {                                                        // this row becomes pivot row j of the panel
                const int row = lane + 32 * ustar;
                ctl->pivrow[g & 3][j] = row;
#pragma unroll
                for (int jj = 0; jj < BVP_NB; jj++) if (jj >= j && jj < nb) W[(size_t)row * Cs + pc0 + jj] = prow[jj];
              }
#pragma unroll
              for (int u = 0; u < RU; u++) {
                const int r = lane + 32 * u;
                double l = 0.0;
                if ((actm >> u) & 1u) {
                  l = pv[u][j] * ipiv;
                  if (l != 0.0) nzm |= 1u << u;
#pragma unroll
                  for (int jj = 0; jj < BVP_NB; jj++) if (jj > j) pv[u][jj] -= l * prow[jj];
                }
                if (r < RS) LpN[r * B3_LPS + j] = -l;
              }
            }
          } else {
#pragma unroll
            for (int u = 0; u < RU; u++) { const int r = lane + 32 * u; if (r < RS) LpN[r * B3_LPS + j] = 0.0; }
          }
        }
        VLT(4)
        __syncwarp();
        if (!failed) {
          // multipliers of the pivot rows inside the panel -> LU block; then clear those rows of the multiplier panel
          double *LUn = C.LU + par * 64;
          const int *pivn = ctl->pivrow[g & 3];
          for (int idx = lane; idx < 64; idx += 32) {
            const int j = idx / BVP_NB, jj = idx % BVP_NB;
            LUn[idx] = (j < nb && jj < j) ? -LpN[pivn[j] * B3_LPS + jj] : 0.0;
          }
          __syncwarp();
          for (int idx = lane; idx < 64; idx += 32) {
            const int j = idx / BVP_NB, jj = idx % BVP_NB;
            if (j < nb) LpN[pivn[j] * B3_LPS + jj] = 0.0;
          }
          unsigned tm = 0;
#pragma unroll
          for (int u = 0; u < RU; u++) {
            const bool a = (actm >> u) & 1u;
            const unsigned b = __ballot_sync(0xffffffffu, a && ((nzm >> u) & 1u));
            if (lane == 0) ctl->rowbits[par][u] = b;
#pragma unroll
            for (int q = 0; q < 4; q++) if ((b >> (8 * q)) & 0xffu) tm |= 1u << (4 * u + q);
            // the eliminated entries of the rows that stay: zero (this half becomes the right half of the next layer)
            if (a) {
              double2 *dst = reinterpret_cast<double2 *>(W + (size_t)(lane + 32 * u) * Cs + pc0);
#pragma unroll
              for (int j2 = 0; j2 < BVP_NB / 2; j2++) if (2 * j2 < nb) dst[j2] = make_double2(0.0, 0.0);
            }
          }
#pragma unroll
          for (int u = RU; u < 4; u++) if (lane == 0) ctl->rowbits[par][u] = 0u;
          if (lane == 0) ctl->tmask[par] = tm;
        } else if (lane == 0) ctl->iflag = 1;
        VLT(5)
      }
      __syncwarp();
      b3_arrive(par ? B3_PR1 : B3_PR0, B3_THREADS);
    }
  }
  // drain the trailing-done barriers of the last two panels
  if (g >= 2) b3_sync((g & 1) ? B3_TD1 : B3_TD0, 32 + B3_UWT);
  if (g >= 1) b3_sync(((g - 1) & 1) ? B3_TD1 : B3_TD0, 32 + B3_UWT);
  VLT(6)
}

// ---- update warps
DEVFN void b3_update_warps(const B3Ctx &C, int uwi) {
  const int lane = threadIdx.x & 31, grp = lane >> 2, tig = lane & 3;
  const int ut = uwi * 32 + lane;
  const int n2 = C.n2, L = C.L, RS = C.RS, Cs = C.Cs, np = C.np, UbS = C.UbS;
  double *AB_RESTRICT W = C.W;
  B3Ctl *ctl = C.ctl;
  int g = 0;
  for (int k = 0; k < L; k++) {
    const bool odd = k & 1, last = (k == L - 1);
    for (int P = 0; P < np; P++, g++) {
      const int jc0 = P * BVP_NB, nb = (n2 - jc0 < BVP_NB) ? n2 - jc0 : BVP_NB;
      const int par = g & 1;
      int nbn;                                         // width of the next panel (its columns belong to the panel warp)
      if (P + 1 < np) nbn = (n2 - jc0 - nb < BVP_NB) ? n2 - jc0 - nb : BVP_NB; else nbn = last ? 0 : (n2 < BVP_NB ? n2 : BVP_NB);
      const int cstart = jc0 + nb + nbn;
      // trailing columns (logical): [cstart, 2n) in tiles that stop at 2n, then [max(cstart,2n), 4n] (rhs included), or the rhs
      // alone in the last layer
      const int nlc = cstart < n2 ? n2 - cstart : 0, nleft = (nlc + 7) >> 3;
      const int rs = last ? 2 * n2 : (cstart > n2 ? cstart : n2);
      const int nrc = 2 * n2 + 1 - rs, nright = (nrc + 7) >> 3;
      b3_sync(par ? B3_PR1 : B3_PR0, B3_THREADS);
      const bool ok = (ctl->iflag == 0);
      double *UbC = C.Ub + (size_t)par * BVP_NB * UbS;
      if (ok) {
        const double *LUn = C.LU + par * 64;
        const int *pivn = ctl->pivrow[g & 3];
        for (int idx = ut; idx < nlc + nrc; idx += B3_UWT) {
          const int c = idx < nlc ? cstart + idx : rs + (idx - nlc);
          const int pc = b3_pc(c, n2, odd);
          double uv[BVP_NB];
#pragma unroll
          for (int j = 0; j < BVP_NB; j++) {
            double v = 0.0;
            if (j < nb) {
              double *w = W + (size_t)pivn[j] * Cs + pc;
              v = *w;
#pragma unroll
              for (int jj = 0; jj < BVP_NB; jj++) if (jj < j) v -= LUn[j * BVP_NB + jj] * uv[jj];
              *w = v;
            }
            uv[j] = v;
            UbC[j * UbS + c] = v;
          }
        }
      }
      b3_sync(B3_UB, B3_UWT);
      if (ok) {
        const double *LpC = C.Lp + (size_t)par * RS * B3_LPS;
        const unsigned mask = ctl->tmask[par];
        const unsigned rb0 = ctl->rowbits[par][0], rb1 = ctl->rowbits[par][1], rb2 = ctl->rowbits[par][2], rb3 = ctl->rowbits[par][3];
        const int nact = __popc(mask), ntc = nleft + nright, total = nact * ntc;
        constexpr int UN = 3;
        for (int i0 = uwi; i0 < total; i0 += UN * B3_NUW) {
          double2 cv[UN]; double a0[UN], a1[UN], b0[UN], b1[UN]; double2 *cp[UN]; bool st[UN];
#pragma unroll
          for (int q = 0; q < UN; q++) {
            const int i = i0 + q * B3_NUW;
            st[q] = false; cv[q] = make_double2(0.0, 0.0); a0[q] = a1[q] = b0[q] = b1[q] = 0.0; cp[q] = nullptr;
            if (i < total) {
              const int ct = i / nact, rt = __fns(mask, 0, (i - ct * nact) + 1);
              const int c0 = ct < nleft ? cstart + 8 * ct : rs + 8 * (ct - nleft);
              const int cend = ct < nleft ? n2 : 2 * n2 + 1;
              const int r = 8 * rt + grp;
              const unsigned rb = (r < 32) ? rb0 : (r < 64 ? rb1 : (r < 96 ? rb2 : rb3));
              const int cl = c0 + 2 * tig;
              st[q] = ((rb >> (r & 31)) & 1u) && cl < cend;
              a0[q] = LpC[r * B3_LPS + tig]; a1[q] = LpC[r * B3_LPS + 4 + tig];
              b0[q] = UbC[tig * UbS + c0 + grp]; b1[q] = UbC[(4 + tig) * UbS + c0 + grp];
              cp[q] = reinterpret_cast<double2 *>(W + (size_t)r * Cs + b3_pc(cl, n2, odd));
              if (st[q]) cv[q] = *cp[q];
            }
          }
#pragma unroll
          for (int q = 0; q < UN; q++) b3_mma(cv[q].x, cv[q].y, a0[q], b0[q]);
#pragma unroll
          for (int q = 0; q < UN; q++) b3_mma(cv[q].x, cv[q].y, a1[q], b1[q]);
#pragma unroll
          for (int q = 0; q < UN; q++) if (st[q]) *cp[q] = cv[q];
        }
      }
      b3_arrive(par ? B3_TD1 : B3_TD0, 32 + B3_UWT);
    }
  }
}

// ---- loader warps
template <int S, class FBF, class FBF0>
DEVFN void b3_loader_warps(const B3Ctx &C, const VLSol &Z, const VLQuad &QD, int lwi, bool include_surface, bool lambertian, bool localbeam,
                           double SFAC, double albedo, double attn, int m, FBF BFq, FBF0 BF0q) {
  const int lane = threadIdx.x & 31, lt = lwi * 32 + lane;
  const int n = C.n, n2 = C.n2, L = C.L, Cs = C.Cs, Cc = C.Cc, np = C.np;
  B3Ctl *ctl = C.ctl;
  int g = 0;
  for (int k = 0; k < L; k++) {
    bool arrived = false;
    const int nrows_next = (k + 1 < L - 1) ? n2 : n;       // rows that join the window at the start of layer k+1
    for (int P = 0; P < np; P++, g++) {
      const int par = g & 1;
      b3_sync(par ? B3_PR1 : B3_PR0, B3_THREADS);
      const bool ok = (ctl->iflag == 0);
      if (P == 0 && lt == 0) ctl->cur = 0;
      {
        // pivot rows of the previous panel are final: spill them (logical column order), recycle their slots
        const int kp = (P > 0) ? k : k - 1, jcp = (P > 0) ? (P - 1) * BVP_NB : (np - 1) * BVP_NB;
        const int nbp = (n2 - jcp < BVP_NB) ? n2 - jcp : BVP_NB;
        const bool oddp = kp & 1;
        const int *pivp = ctl->pivrow[(g - 1) & 3];
        if (g >= 1 && ok)
          for (int j = lwi; j < nbp; j += 2) {
            const double *src = C.W + (size_t)pivp[j] * Cs;
            double *dst = C.Ug + ((size_t)kp * n2 + jcp + j) * Cc;
            for (int c = lane; c < Cc; c += 32) dst[c] = src[b3_pc(c, n2, oddp)];
          }
        b3_sync(B3_LB, B3_LWT);
        if (g >= 1 && ok && lt == 0) { int f = ctl->fqn; for (int j = 0; j < nbp; j++) C.fq[f++] = pivp[j]; ctl->fqn = f; }
      }
      if (lt == 0) {
        int nst = 0;
        if (ok && k + 1 < L) {
          nst = nrows_next - ctl->cur;
          if (nst > ctl->fqn) nst = ctl->fqn;
          int f = ctl->fqn;
          for (int i = 0; i < nst; i++) C.stslot[i] = C.fq[--f];
          ctl->fqn = f;
          ctl->st_r0 = ctl->cur; ctl->cur += nst;
        }
        ctl->nst = nst;
      }
      b3_sync(B3_LB, B3_LWT);
      const int nst = ctl->nst;
      if (nst > 0) {
        b3_load_rows<S>(C, Z, QD, k + 1, ctl->st_r0, nst, C.stslot, lt, B3_LWT, include_surface, lambertian, localbeam, SFAC, albedo, attn, m, BFq, BF0q);
        for (int i = lt; i < nst; i += B3_LWT) C.rowst[C.stslot[i]] = 2;
      }
      const bool done = (k + 1 < L) && (!ok || ctl->cur == nrows_next);
      b3_sync(B3_LB, B3_LWT);
      if (done && !arrived) { b3_arrive(B3_RR, 32 + B3_LWT); arrived = true; }
    }
    // (with >= 16 spare slots every row is staged before the last event of the layer; see the header comment)
    if (k + 1 < L && !arrived) b3_arrive(B3_RR, 32 + B3_LWT);
  }
}

// ---- one pixel: forward elimination (this file) + bvp_backward
template <int S>
DEVFN void bvp_pixel_v3(const BlockTeam &T, const VLPix &X, const VLSol &Z, const VLQuad &QD, const VLSurf &SF, const VLConv &CV,
                        int lambertian, double accuracy, int m, int p, double *scr, int *iscr, double *Ug) {
  const int N = QD.N, n = S * N, n2 = 2 * n, L = X.L;
  B3Ctx C;
  C.S = S; C.N = N; C.n = n; C.n2 = n2; C.L = L; C.RS = b3_rows(n); C.Cs = b3_cs(n); C.Cc = bvp_cols(n); C.UbS = b3_ubs(n);
  C.np = (n2 + BVP_NB - 1) / BVP_NB;
  C.W = scr; C.Lp = C.W + (size_t)C.RS * C.Cs; C.Ub = C.Lp + 2 * (size_t)C.RS * B3_LPS; C.LU = C.Ub + 2 * (size_t)BVP_NB * C.UbS; C.Uloc = C.LU + 128;
  double *xcur = scr + b3_big_doubles(n), *xnext = xcur + n2, *red = xnext + n2;
  C.ctl = reinterpret_cast<B3Ctl *>(iscr);
  C.rowst = iscr + 64; C.fq = C.rowst + C.RS; C.stslot = C.fq + C.RS;
  C.q0 = (size_t)p * L; C.Ug = Ug;
  const double *sc = X.scal + (size_t)p * 16;
  const int *ib = X.ibuf + (size_t)p * 4;
  int *cst = CV.state + (size_t)p * 4;
  const double cos_sza = sc[0], albedo = sc[4], FLUXF = sc[5];
  const bool refl_direct = ib[2] != 0;
  bool include_surface = true;
  if (lambertian) { if (m != 0) include_surface = false; else if (albedo == 0.0) include_surface = false; }
  const double SFAC = (m == 0) ? 2.0 : 1.0, DELTA_FACTOR = (m == 0) ? 1.0 : 2.0;
  const bool localbeam = include_surface && refl_direct;
  const double *BF = nullptr, *BF0 = nullptr;
  const int nq = SF.nq;
  if (!lambertian) {
    BF = SF.brdf_f + ((size_t)p * SF.nmfour + m) * SF.nq * N * N;
    BF0 = SF.brdf_f_0 + ((size_t)p * SF.nmfour + m) * SF.nq * N;
  }
  auto BFq = [=](int o1, int o2, int i, int jj) -> double {
    return nq == 16 ? BF[((size_t)(4 * o1 + o2) * N + i) * N + jj] : ((o1 | o2) == 0 ? BF[(size_t)i * N + jj] : 0.0);
  };
  auto BF0q = [=](int o1, int i) -> double { return nq == 16 ? BF0[(size_t)(4 * o1) * N + i] : (o1 == 0 ? BF0[i] : 0.0); };
  double attn = 0.0;
  if (localbeam) attn = FLUXF * (4.0 * cos_sza / DELTA_FACTOR) / PI4 * sc[3];
  const size_t q0 = C.q0;
  const int Cs = C.Cs, RS = C.RS;
  VLT_DECL

  // ---- prologue (all threads): TOA rows -> slots 0..n-1, rows of the first interface -> slots n.., rest free
  if (T.rank == 0) { C.ctl->iflag = 0; C.ctl->cur = 0; C.ctl->nst = 0; }
  for (int r = T.rank; r < RS; r += T.size) C.rowst[r] = 0;
  for (int idx = T.rank; idx < n * n; idx += T.size) {
    const int e = idx / n, r = idx % n;
    double *row = C.W + (size_t)r * Cs;
    row[e] = Z.xp[(q0 * n + e) * n + r];
    row[n + e] = Z.xn[(q0 * n + e) * n + r] * Z.tde[q0 * n + e];
    row[n2 + e] = 0.0; row[n2 + n + e] = 0.0;
  }
  for (int r = T.rank; r < n; r += T.size) C.W[(size_t)r * Cs + 4 * n] = -Z.wupper[q0 * n2 + r];
  const int nfirst = (L > 1) ? n2 : n;
  for (int t = T.rank; t < nfirst; t += T.size) C.stslot[t] = n + t;
  T.sync();
  b3_load_rows<S>(C, Z, QD, 0, 0, nfirst, C.stslot, T.rank, T.size, include_surface, lambertian != 0, localbeam, SFAC, albedo, attn, m, BFq, BF0q);
  for (int r = T.rank; r < n + nfirst; r += T.size) C.rowst[r] = 2;
  if (T.rank == 0) { int f = 0; for (int r = RS - 1; r >= n + nfirst; r--) C.fq[f++] = r; C.ctl->fqn = f; }
  T.sync();
  VLT(0)

  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    const int RU = (RS + 31) / 32;
    if (RU == 1) b3_panel_warp<1>(C); else if (RU == 2) b3_panel_warp<2>(C); else if (RU == 3) b3_panel_warp<3>(C); else b3_panel_warp<4>(C);
  } else if (warp == 4 || warp == 8) {
    b3_arrive(B3_RR, 32 + B3_LWT);                    // the rows of layer 0 are in place
    b3_loader_warps<S>(C, Z, QD, warp == 4 ? 0 : 1, include_surface, lambertian != 0, localbeam, SFAC, albedo, attn, m, BFq, BF0q);
  } else {
    b3_update_warps(C, warp < 4 ? warp - 1 : (warp < 8 ? warp - 2 : warp - 3));
  }
  T.sync();
  VLT_RESET
  if (C.ctl->iflag) { if (T.rank == 0) { cst[0] = 0; cst[3] = 4; } T.sync(); return; }
  // pivot rows of the very last panel
  {
    const int jcp = (C.np - 1) * BVP_NB, nbp = n2 - jcp, G = L * C.np;
    const int *pivp = C.ctl->pivrow[(G - 1) & 3];
    const bool oddp = (L - 1) & 1;
    const int lane = threadIdx.x & 31;
    for (int j = warp; j < nbp; j += B3_THREADS / 32) {
      const double *src = C.W + (size_t)pivp[j] * Cs;
      double *dst = Ug + ((size_t)(L - 1) * n2 + jcp + j) * C.Cc;
      for (int c = lane; c < C.Cc; c += 32) dst[c] = src[b3_pc(c, n2, oddp)];
    }
  }
  T.sync();
  VLT_RESET
  VLT(7)
  bvp_backward<S>(T, X, Z, QD, SF, CV, lambertian, accuracy, m, p, scr, xcur, xnext, red, Ug);
}

}  // namespace vlk
#endif
