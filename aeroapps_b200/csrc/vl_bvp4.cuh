// Stage B, kernel K3, forward elimination for the GPU -- register-resident window ("v4").
// Same mathematics as bvp_pixel's forward phase in vl_bvp.cuh (a window of at most 3n rows x (4n+1) columns eliminated
// layer by layer with partial pivoting over every active row of the window, right-hand side carried as a column), but
// the window never lives in shared memory:
//   * 12 warps = 4 row blocks x 3 column groups.  Lane l of row block rw owns window slot 32 rw + l; column group cg
//     owns the columns c with c % 3 == cg of each half (left half = unknowns of layer k, right half = layer k+1), so a
//     thread keeps 2H+1 matrix entries of ONE row in registers (H = 24: n <= 36) and the rank-1 update of a pivot step
//     is pure register DFMA against the pivot row, which is read as 16-byte shared-memory broadcasts.
//   * a pivot step: the column group that owns column j+1 updates it first and publishes it (128 doubles), barrier,
//     every warp redundantly finds the next pivot (integer redux on the IEEE bit pattern, lowest slot on ties), the
//     three owner threads of that row publish it after their update of step j, barrier.  Two block barriers per
//     column and no shared-memory traffic for the matrix itself.
//   * the loop over local column indices is unrolled, so every register index is static and the update of step j
//     only touches the columns that are still alive; at the end of a layer the right half is moved to the left half
//     (24 register moves) instead of shifting a shared-memory window.
//   * the eigen-solutions of the layers k, k+1 (and k+2, prefetched with cp.async) are staged in shared memory;
//     new interface rows are formed from there straight into the registers of the free slots.
//   * pivot rows go to HBM in logical column order (all warps copy a slice each) for bvp_backward.
// Shared-memory use of the forward phase is ~70 KB (n = 36); the backward phase (two layers of pivot rows) keeps the
// 170 KB footprint of the kernel.
#pragma once
#include "vl_bvp.cuh"

#ifndef AB_HOST_MODEL
namespace vlk {

constexpr int B4_THREADS = 384;
constexpr int B4_H = 24;                                            // local columns per half and group
HDFN bool b4_supported(int n) { return n >= 6 && 2 * n <= 3 * B4_H; }
HDFN int b4_xs(int n) { return (2 * n * n + 5 * n + 1) & ~1; }     // doubles of one staged layer: xp, xn, tde, wupper, wlower
HDFN size_t b4_fwd_doubles(int n) { return (size_t)6 * (2 * B4_H + 2) + 256 + 8 + 3 * (size_t)b4_xs(n); }
HDFN size_t b4_big_doubles(int n) { const size_t a = b4_fwd_doubles(n), b = bvp_big_doubles(n); return ((a > b ? a : b) + 1) & ~(size_t)1; }
HDFN size_t b4_smem_doubles(int n) { return b4_big_doubles(n) + 8 * n + 16; }

DEVFN void b4_cp8(double *dst, const double *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(vl_smem_u32(dst)), "l"(src) : "memory");
}
DEVFN void b4_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DEVFN void b4_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// eigen-solutions of problem q -> shared memory (asynchronous; completes at the next b4_wait_all + barrier)
DEVFN void b4_stage(double *dst, const VLSol &Z, size_t q, int n, int tid) {
  const int nn = n * n, n2 = 2 * n;
  const double *xp = Z.xp + q * nn, *xn = Z.xn + q * nn;
  for (int i = tid; i < nn; i += B4_THREADS) { b4_cp8(dst + i, xp + i); b4_cp8(dst + nn + i, xn + i); }
  for (int i = tid; i < n; i += B4_THREADS) b4_cp8(dst + 2 * nn + i, Z.tde + q * n + i);
  for (int i = tid; i < n2; i += B4_THREADS) {
    b4_cp8(dst + 2 * nn + n + i, Z.wupper + q * n2 + i);
    b4_cp8(dst + 2 * nn + 3 * n + i, Z.wlower + q * n2 + i);
  }
}

// next pivot, stage 1 (the four warps that own the column, one value per lane; inactive slots hold 0): the lane with the
// largest |value| of the warp, lowest lane on exact ties, leaves (value, slot) in the warp's candidate entry
DEVFN void b4_warp_cand(double v, int slot, int lane, double *cv, int *cs) {
  const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(v));
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  bool cand = (hi == mh);
  unsigned bal = __ballot_sync(0xffffffffu, cand);
  if (__popc(bal) > 1) {
    const unsigned ml = __reduce_max_sync(0xffffffffu, cand ? lo : 0u);
    cand = cand && (lo == ml);
    bal = __ballot_sync(0xffffffffu, cand);
  }
  if (lane == __ffs(bal) - 1) { *cv = v; *cs = slot; }
}
// stage 2 (every thread, identical result): the best of the four candidates, lowest row block (= lowest slot) on ties
DEVFN void b4_select(const double *cv, const int *cs, int &s, double &pv) {
  const double2 a = *reinterpret_cast<const double2 *>(cv), b = *reinterpret_cast<const double2 *>(cv + 2);
  const int4 sl = *reinterpret_cast<const int4 *>(cs);
  double best = fabs(a.x); pv = a.x; s = sl.x;
  if (fabs(a.y) > best) { best = fabs(a.y); pv = a.y; s = sl.y; }
  if (fabs(b.x) > best) { best = fabs(b.x); pv = b.x; s = sl.z; }
  if (fabs(b.y) > best) { best = fabs(b.y); pv = b.y; s = sl.w; }
}

template <int S>
DEVFN void bvp_pixel_v4(const BlockTeam &T, const VLPix &X, const VLSol &Z, const VLQuad &QD, const VLSurf &SF, const VLConv &CV,
                        int lambertian, double accuracy, int m, int p, double *scr, int *iscr, double *Ug) {
  constexpr int H = B4_H, NLC = 2 * H + 1, PBS = NLC + 1;
  const int N = QD.N, n = S * N, n2 = 2 * n, nn = n * n, L = X.L, Cc = bvp_cols(n);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rw = warp / 3, cg = warp - 3 * rw, slot = 32 * rw + lane;
  double *PB = scr;                        // [2][3][PBS] pivot rows, layout [group][local column]
  double *cb = PB + 6 * PBS;               // [2][128] published column
  double *candv = cb + 256;                // [2][4] pivot candidates of the four row blocks (value)
  double *Xs = candv + 8;                  // [3][XS] staged layers
  const int XS = b4_xs(n);
  double *xcur = scr + b4_big_doubles(n), *xnext = xcur + n2, *red = xnext + n2;
  int *cnt = iscr;                         // [4] free slots per row block
  int *cands = iscr + 8;                   // [2][4] pivot candidates (slot), 16-byte aligned
  // slice of a pivot row this thread copies to HBM: logical column sp_c <- offset sp_src inside a pivot-row buffer
  // (group row sp_row, physical index sp_phys at the start of a layer; -1: nothing to copy)
  int sp_c = warp * 13 + lane, sp_row = 0, sp_phys = -1;
  if (lane < 13 && sp_c < Cc) {
    if (sp_c < n2) { sp_row = (sp_c % 3) * PBS; sp_phys = sp_c / 3; }
    else if (sp_c < 2 * n2) { sp_row = ((sp_c - n2) % 3) * PBS; sp_phys = H + (sp_c - n2) / 3; }
    else { sp_row = 0; sp_phys = 2 * H; }
  } else sp_c = 0;
  const int cgo = cg * PBS;
  double *cbp = cb + slot;
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
  double attn = 0.0;
  if (localbeam) attn = FLUXF * (4.0 * cos_sza / DELTA_FACTOR) / PI4 * sc[3];
  const size_t q0 = (size_t)p * L;
  VLT_DECL

  double w[NLC];
#pragma unroll
  for (int lc = 0; lc < NLC; lc++) w[lc] = 0.0;
  bool active = false, failed = false;

  b4_stage(Xs, Z, q0, n, tid);
  if (L > 1) b4_stage(Xs + XS, Z, q0 + 1, n, tid);
  b4_commit(); b4_wait_all();
  __syncthreads();
  VLT(0)

  for (int k = 0; k < L; k++) {
    const bool last = (k == L - 1);
    if (k + 2 < L) { b4_stage(Xs + (size_t)((k + 2) % 3) * XS, Z, q0 + k + 2, n, tid); b4_commit(); }
    const double *Xa = Xs + (size_t)(k % 3) * XS, *Xb = Xs + (size_t)((k + 1) % 3) * XS;
    const double *xpA = Xa, *xnA = Xa + nn, *tdeA = Xa + 2 * nn, *wlowA = Xa + 2 * nn + 3 * n;
    const double *xpB = Xb, *xnB = Xb + nn, *tdeB = Xb + 2 * nn, *wupB = Xb + 2 * nn + n;
    double *RP = const_cast<double *>(Xb), *RM = RP + nn;       // last layer only: surface reflection of the solutions, [e][r]
    if (last) {
      for (int idx = tid; idx < nn; idx += B4_THREADS) {
        const int e = idx / n, r = idx - e * n, i = r / S, o1 = r - S * i;
        double rp = 0.0, rm = 0.0;
        if (include_surface) {
          if (lambertian) {
            if (o1 == 0) {
              for (int jj = 0; jj < N; jj++) { const double a = QD.x[jj] * QD.w[jj]; rp += a * xpA[e * n + S * jj]; rm += a * xnA[e * n + S * jj]; }
              rp *= SFAC * albedo; rm *= SFAC * albedo;
            }
          } else {
            for (int jj = 0; jj < N; jj++)
              for (int o2 = 0; o2 < S; o2++) {
                const double a = QD.x[jj] * QD.w[jj] * BFq(o1, o2, i, jj);
                rp += a * xpA[e * n + S * jj + o2]; rm += a * xnA[e * n + S * jj + o2];
              }
            rp *= SFAC; rm *= SFAC;
          }
        }
        RP[idx] = rp; RM[idx] = rm;
      }
    }
    // ---- new rows -> free slots (rank of the slot among the free ones = row number)
    const unsigned fm = __ballot_sync(0xffffffffu, !active);
    if (cg == 0 && lane == 0) cnt[rw] = __popc(fm);
    __syncthreads();
    if (!active) {
      int fidx = __popc(fm & ((1u << lane) - 1u));
      for (int u = 0; u < rw; u++) fidx += cnt[u];
      const int ntoa = (k == 0) ? n : 0, nint = last ? n : n2;
#pragma unroll
      for (int lc = 0; lc < NLC; lc++) w[lc] = 0.0;
      if (fidx < ntoa) {
        // TOA rows: no downwelling diffuse radiation
        const int r = fidx;
#pragma unroll
        for (int lc = 0; lc < H; lc++) {
          const int c = 3 * lc + cg;
          if (c < n2) w[lc] = (c < n) ? xpA[c * n + r] : xnA[(c - n) * n + r] * tdeA[c - n];
        }
        if (cg == 0) w[2 * H] = -Xa[2 * nn + n + r];                 // -WUPPER(r) of layer 0
        active = true;
      } else if (fidx < ntoa + nint) {
        const int r2 = fidx - ntoa;
        if (!last) {
          // continuity at the interface below layer k: full-stream XPOS/XNEG from the stored downwelling halves
          const int rlo = r2 < n ? r2 : r2 - n;
          const double sg = (r2 >= n && (rlo % S) > 1) ? -1.0 : 1.0;
          const double *paA = (r2 < n) ? xpA : xnA, *naA = (r2 < n) ? xnA : xpA;
          const double *paB = (r2 < n) ? xpB : xnB, *naB = (r2 < n) ? xnB : xpB;
#pragma unroll
          for (int lc = 0; lc < H; lc++) {
            const int c = 3 * lc + cg;
            if (c < n) {
              w[lc] = tdeA[c] * (sg * paA[c * n + rlo]);
              w[H + lc] = -(sg * paB[c * n + rlo]);
            } else if (c < n2) {
              const int e = c - n;
              w[lc] = sg * naA[e * n + rlo];
              w[H + lc] = -tdeB[e] * (sg * naB[e * n + rlo]);
            }
          }
          if (cg == 0) w[2 * H] = wupB[r2] - wlowA[r2];
        } else {
          // surface rows (upwelling streams): X(i+N) - R2_HOM ; rhs = -WLOWER(i+N) + R2_BEAM + direct beam
          const int r = r2, i = r / S, o1 = r - S * i;
          const double sg = (o1 > 1) ? -1.0 : 1.0;
#pragma unroll
          for (int lc = 0; lc < H; lc++) {
            const int c = 3 * lc + cg;
            if (c < n) w[lc] = tdeA[c] * (sg * xnA[c * n + r] - RP[c * n + r]);
            else if (c < n2) w[lc] = sg * xpA[(c - n) * n + r] - RM[(c - n) * n + r];
          }
          if (cg == 0) {
            double v = -wlowA[n + r];
            if (include_surface) {
              double rb = 0.0;
              if (lambertian) {
                if (o1 == 0) { for (int jj = 0; jj < N; jj++) rb += QD.x[jj] * QD.w[jj] * wlowA[S * jj]; rb *= SFAC * albedo; }
              } else {
                for (int jj = 0; jj < N; jj++) for (int o2 = 0; o2 < S; o2++) rb += QD.x[jj] * QD.w[jj] * BFq(o1, o2, i, jj) * wlowA[S * jj + o2];
                rb *= SFAC;
              }
              v += rb;
              if (localbeam) {
                if (lambertian) { if (m == 0 && o1 == 0) v += attn * albedo; }
                else v += attn * (nq == 16 ? BF0[(size_t)(4 * o1) * N + i] : (o1 == 0 ? BF0[i] : 0.0));
              }
            }
            w[2 * H] = v;
          }
        }
        active = true;
      }
    }
    VLT(1)
    // ---- first pivot of the layer (column 0 = group 0, local 0)
    int s; double pv;
    if (cg == 0) {
      const double v = active ? w[0] : 0.0;
      cb[slot] = v;
      b4_warp_cand(v, slot, lane, candv + rw, cands + rw);
    }
    __syncthreads();
    b4_select(candv, cands, s, pv);
    if (!(fabs(pv) > 0.0)) failed = true;
    if (!failed && slot == s) {
      double *pb = PB + (size_t)cg * PBS;
#pragma unroll
      for (int lc = 0; lc < NLC; lc++) pb[lc] = w[lc];
    }
    __syncthreads();
    if (failed) break;
    double rpiv = 1.0 / pv;
    VLT(2)
    // ---- eliminate the 2n columns of layer k in three blocks of 24 steps; inside a block every index is static
    //      (step t = 3 jl + g works on the column with physical register index jl of group g), after a block the
    //      registers are rotated by 8 positions, so one copy of the block's code serves the whole layer
    double *ugp = Ug + (size_t)k * n2 * Cc + sp_c;
#pragma unroll 1
    for (int b = 0; b < 3; b++) {
      const int spo = (sp_phys >= 8 * b) ? sp_row + sp_phys - 8 * b : -1;
#pragma unroll
      for (int jl = 0; jl < 8; jl++) {
#pragma unroll
        for (int g = 0; g < 3; g++) {
          const int t = 3 * jl + g, par = t & 1, j = 24 * b + t;
          if (j < n2 && !failed) {
            const bool has_next = (j + 1 < n2);
            const int gn = (g + 1) % 3, nl = (g < 2) ? jl : jl + 1;      // group / physical index of column j+1
            const bool is_gn = has_next && (cg == gn);
            const double *pb = PB + par * 3 * PBS + cgo;
            double l = 0.0;
            if (slot == s) active = false;
            else if (active) l = cbp[par * 128] * rpiv;
            if (spo >= 0) *ugp = PB[par * 3 * PBS + spo];                  // pivot row j -> HBM, logical column order
            ugp += Cc;
            VLT(3)
            int s2 = 0; double pv2 = 1.0, rpiv2 = 1.0;
            // phase 0: the owners of column j+1 update it, publish it and their candidate for the next pivot while the
            // other warps do their rank-1 update; barrier; phase 1: everyone picks the pivot among the four candidates,
            // the owners of column j+1 catch up with their rank-1 update
#pragma unroll 1
            for (int ph = 0; ph < 2; ph++) {
              if (ph == 0) {
                if (is_gn) {
                  w[nl] -= l * pb[nl];
                  const double v = active ? w[nl] : 0.0;
                  cbp[(par ^ 1) * 128] = v;
                  b4_warp_cand(v, slot, lane, candv + (par ^ 1) * 4 + rw, cands + (par ^ 1) * 4 + rw);
                }
              } else {
                VLT(4)
                __syncthreads();                                             // A
                VLT(5)
                if (has_next) {
                  b4_select(candv + (par ^ 1) * 4, cands + (par ^ 1) * 4, s2, pv2);
                  if (!(fabs(pv2) > 0.0)) failed = true;
                  rpiv2 = 1.0 / pv2;
                }
              }
              if ((ph == 1) == is_gn) {
                // rank-1 update of the live columns (column j+1 is already done; the dead tail of a block is all zero)
                if (__any_sync(0xffffffffu, l != 0.0)) {
                  const bool u0 = (cg > g + 1), u1 = !(has_next && g == 2 && cg == 0);
#pragma unroll
                  for (int lc = (jl & ~1); lc < NLC; lc += 2) {
                    if ((lc < 34 || b < 2) && (lc < 42 || b < 1)) {
                      const double2 p2 = *reinterpret_cast<const double2 *>(pb + lc);
                      if (lc >= jl) { if (lc == jl) { if (u0) w[lc] -= l * p2.x; } else if (lc == jl + 1) { if (u1) w[lc] -= l * p2.x; } else w[lc] -= l * p2.x; }
                      if (lc + 1 < NLC) { if (lc + 1 == jl) { if (u0) w[lc + 1] -= l * p2.y; } else if (lc + 1 == jl + 1) { if (u1) w[lc + 1] -= l * p2.y; } else w[lc + 1] -= l * p2.y; }
                    }
                  }
                }
              }
            }
            VLT(6)
            // the owners of the next pivot row publish it (live columns only; the last step of a block writes the layout
            // of the next block, i.e. shifted by the rotation)
            if (has_next && !failed && slot == s2) {
              double *pn = PB + (par ^ 1) * 3 * PBS + cgo;
              if (t == 23) {
#pragma unroll
                for (int lc = 8; lc < NLC; lc += 2) {
                  if (lc + 1 < NLC) *reinterpret_cast<double2 *>(pn + lc - 8) = make_double2(w[lc], w[lc + 1]);
                  else pn[lc - 8] = w[lc];
                }
#pragma unroll
                for (int lc = NLC - 8; lc < NLC; lc++) pn[lc] = 0.0;
              } else {
#pragma unroll
                for (int lc = (jl & ~1); lc < NLC; lc += 2) {
                  if (lc + 1 < NLC) *reinterpret_cast<double2 *>(pn + lc) = make_double2(w[lc], w[lc + 1]);
                  else pn[lc] = w[lc];
                }
              }
            }
            VLT(11)
            __syncthreads();                                                 // B
            VLT(12)
            s = s2; rpiv = rpiv2;
          }
        }
      }
      // rotate: the 8 eliminated columns of every group leave, 8 zero columns enter at the end
#pragma unroll
      for (int lc = 0; lc < NLC - 8; lc++) w[lc] = w[lc + 8];
#pragma unroll
      for (int lc = NLC - 8; lc < NLC; lc++) w[lc] = 0.0;
    }
    if (failed) break;
    // ---- the three rotations have moved the unknowns of layer k+1 to the left half; the rhs goes back to its place
    w[2 * H] = w[H]; w[H] = 0.0;
    b4_wait_all();
    __syncthreads();
    VLT(7)
  }
  __syncthreads();
  if (failed) { b4_wait_all(); if (tid == 0) { cst[0] = 0; cst[3] = 4; } __syncthreads(); return; }
  bvp_backward<S>(T, X, Z, QD, SF, CV, lambertian, accuracy, m, p, scr, xcur, xnext, red, Ug);
}

}  // namespace vlk
#endif
