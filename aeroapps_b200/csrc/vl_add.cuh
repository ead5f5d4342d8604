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
//   up-sweep, k = L .. 1:   Rs_k = R_k + Tm_k Rs_{k+1} E_k Tm_k,   E_k = (1 - R_k Rs_{k+1})^-1     (reflection of everything below;
//                           E_k applied as the product prod_j (1 + (R_k Rs_{k+1})^(2^j)), see the loop)
//                           Ss_k = sg_up + Tm_k (Ss_{k+1} + Rs_{k+1} y_k),   [Y_k | y_k] = E_k [Tm_k | R_k Ss_{k+1} + sg_dn]
//                           Rs_{L+1}, Ss_{L+1} = the surface (Lambertian / BRDF Fourier component, BVP_SURFACE_SETUP_HOM/_BEAM)
//   down-sweep, k = 1 .. L: I_dn(bot_k) = Y_k I_dn(top_k) + y_k,   u(bot_k) = Rs_{k+1} I_dn(bot_k) + Ss_{k+1},  I_dn(top_1) = 0
//                           a + b = S+ (in_dn + in_u),  a - b = S- (in_dn - in_u)                 (LCON, MCON of the layer)
// Work: 6 - 8 n x n products per layer for a typical column (12 - 16 n^3 flops against the 37 n^3 of the block-bidiagonal LU), all
// of them register-tiled products out of shared memory with no pivot-column chain (the reference's LU has 5184 chained columns per
// pixel and Fourier term), and 66 KB of shared memory per pixel instead of 125 KB, so three pixels share an SM.  The layer source terms are accumulated top-down with the running user-stream
// transmittance (the reference recurses bottom-up: same sum, different association).
#pragma once
#include "vl_bvp.cuh"
#include "vl_layerops.cuh"

namespace vlk {

// team scratch (doubles): six n x ld matrix arrays (ld = mat_ld(n)) used as Rs + five rotating arrays (R_k, two Q, two Y) in the up-sweep and
// as two copy slots in the down-sweep; 14 vectors
HDFN size_t add_smem_doubles(int n, bool) { return 6 * (size_t)n * mat_ld(n) + 14 * n + 2; }
// per-team global scratch (doubles): per layer Y, Rs Y (n x ld each), y, Rs y + Ss (n each) -- one contiguous block per layer
HDFN size_t add_layer_doubles(int n) { return 2 * (size_t)n * mat_ld(n) + 2 * n; }
HDFN size_t add_work_doubles(int n, int L) { return (size_t)L * add_layer_doubles(n); }

// an eigenproblem or the boundary-value solve failed: every view of the pixel leaves the Fourier loop with status 4
template <class Team>
DEVFN void add_fail_pixel(const Team &T, const VLPix &X, const VLConv &CV, int p) {
  for (int g = T.rank; g < X.G; g += T.size) { int *c = CV.state + (size_t)(view0(X, p) + g) * 4; c[0] = 0; c[3] = 4; }
  if (T.rank == 0) CV.col_active[p] = 0;
}

template <int S, class Team>
DEVFN void add_pixel(const Team &T, const VLPix &X, const VLSol &Z, const VLOps &O, const VLQuad &QD, const VLSurf &SF, const VLConv &CV,
                     int lambertian, double accuracy, int m, int p, double *scr, int *piv, double *Wg) {
  const int N = QD.N, n = S * N, L = X.L, NM = X.NM, ld = mat_ld(n);
  const size_t nl = (size_t)n * ld;
  double *Rs = scr, *pool = scr + nl;                         // pool: five matrix arrays whose roles rotate from layer to layer
  double *Ss = scr + 6 * nl, *Sn = Ss + n, *su = Sn + n, *sd = su + n, *sun = sd + n, *sdn = sun + n, *yv = sdn + n, *zz = yv + n, *xv = zz + n,
         *xb = xv + n, *ub = xb + n, *inp = ub + n, *inm = inp + n;
  const int G = X.G, v0 = view0(X, p);                        // p = pixel; its views are v0 .. v0 + G - 1
  const double *sc = X.scal + (size_t)v0 * 16;                // sun, flux, albedo: the same for every view of the pixel
  const int *ib = X.ibuf + (size_t)v0 * 4;
  const double cos_sza = sc[0], albedo = sc[4], FLUXF = sc[5];
  const int cutoff = ib[0];
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
  double attn = 0.0;
  if (localbeam) attn = FLUXF * (4.0 * cos_sza / DELTA_FACTOR) / PI4 * sc[3];
  const size_t q0 = (size_t)p * L;
  const size_t wl = add_layer_doubles(n), ob = ops_block_doubles(n);

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

  // ---- up-sweep.  E_k = (1 - Q)^-1, Q = R_k Rs, is never formed: reflectances are contractions (spectral radius of Q < 1), so
  //      E_k [Tm_k | v] = prod_j (1 + Q^(2^j)) [Tm_k | v]   -- one product per factor and one squaring per level, no pivoting,
  //      no pivot-column chain; the loop stops when the next factor is 1 to working precision (2 - 3 levels for the optically
  //      thin layers that make up most of a column, 6 - 7 for a thick cloud over a bright surface).
  //      Copy channel 0 brings the next layer's operators [R | Tm] into two free arrays while the current layer finishes; channel 1
  //      brings Tm_k back for the last product (the doubling overwrites the copy it started from); channel 2 the source vectors.
  double *Rk = pool, *Yc = pool + nl, *Qc = pool + 2 * nl, *Qn = pool + 3 * nl, *Yn = pool + 4 * nl;
  {
    const double *blk = O.rt + (q0 + L - 1) * ob;
    T.copy_begin(0, Rk, blk, nl);
    T.copy_wait(0);
    T.copy_begin(0, Yc, blk + nl, nl);
    T.copy_wait(0);
    T.copy_begin(2, su, blk + 2 * nl, 2 * (size_t)n);         // su, sd are adjacent
    T.copy_wait(2);
  }
  T.sync();
  for (int k = L - 1; k >= 0; k--) {
    const double *blk = O.rt + (q0 + k) * ob;
    double *yc = yv, *yn = zz;
    // Q = R_k Rs,   y = R_k Ss + sg_dn
    double qmax = 0.0;
    team_mm(T, n, n, n, Rk, ld, Rs, ld, [&](int r, int c, double acc) { Qc[(size_t)r * ld + c] = acc; qmax = fmax(qmax, fabs(acc)); });
    for (int r = T.rank; r < n; r += T.size) {
      const double *Rr = Rk + (size_t)r * ld;
      double v = sd[r];
      for (int c = 0; c < n; c++) v += Rr[c] * Ss[c];
      yc[r] = v;
    }
    { int dummy = 0; qmax = T.argmax(qmax, dummy) * (double)n; }          // >= the max-row-sum norm of Q; the collective also orders Q, y
    T.sync();
    bool okq = true;
    for (int lev = 0; qmax > 1.0e-17; lev++) {
      if (lev > 14 || !(qmax < 1.0e30)) { okq = false; break; }           // not a contraction (or NaN input): fail the pixel
      // [Y | y] <- (1 + Q) [Y | y]
      team_mm(T, n, n, n, Qc, ld, Yc, ld, [&](int r, int c, double acc) { Yn[(size_t)r * ld + c] = Yc[(size_t)r * ld + c] + acc; });
      for (int r = T.rank; r < n; r += T.size) {
        const double *Qr = Qc + (size_t)r * ld;
        double v = yc[r];
        for (int c = 0; c < n; c++) v += Qr[c] * yc[c];
        yn[r] = v;
      }
      { double *t = Yc; Yc = Yn; Yn = t; t = yc; yc = yn; yn = t; }
      if (qmax * qmax <= 1.0e-17) { T.sync(); break; }                    // the next factor 1 + Q^2 is 1 to working precision
      double q2 = 0.0;
      team_mm(T, n, n, n, Qc, ld, Qc, ld, [&](int r, int c, double acc) { Qn[(size_t)r * ld + c] = acc; q2 = fmax(q2, fabs(acc)); });
      { double *t = Qc; Qc = Qn; Qn = t; }
      { int dummy = 0; qmax = T.argmax(q2, dummy) * (double)n; }
      T.sync();
    }
    if (!okq) { add_fail_pixel(T, X, CV, p); return; }
    // Tm_k comes back into the array of the last Q (channel 1) while  Zb = Rs Y,  zb = Rs y  are formed;  Y and y go to the team's
    // global scratch for the down-sweep as they are (row-major, padded)
    double *Wk = Wg + (size_t)k * wl, *Zb = Qn, *Tk = Qc, *zb = yn;
    T.copy_begin(1, Tk, blk + nl, nl);
    if (k > 0) T.copy_begin(3, Yn, O.rt + (q0 + k - 1) * ob + nl, nl);    // the spare Y array: Tm of the next layer (channel 3)
    for (int idx = T.rank; idx < (int)nl; idx += T.size) Wk[idx] = Yc[idx];
    team_mm(T, n, n, n, Rs, ld, Yc, ld, [&](int r, int c, double acc) { Zb[(size_t)r * ld + c] = acc; });
    for (int r = T.rank; r < n; r += T.size) {
      const double *Rr = Rs + (size_t)r * ld;
      double v = 0.0;
      for (int c = 0; c < n; c++) v += Rr[c] * yc[c];
      zb[r] = v;
      Wk[2 * nl + r] = yc[r];
      Wk[2 * nl + n + r] = v + Ss[r];                          // u(bot_k) = (Rs Y) I_dn(top_k) + this
    }
    // Y is in the scratch: its array takes the next layer's R (channel 0) while this layer finishes
    if (k > 0) {
      const double *nb = O.rt + (q0 + k - 1) * ob;
      T.copy_begin(0, Yc, nb, nl);                             // (the barrier inside also orders Zb, zb)
      T.copy_begin(2, sun, nb + 2 * nl, 2 * (size_t)n);        // sun, sdn are adjacent
    } else T.sync();
    for (int idx = T.rank; idx < (int)nl; idx += T.size) { const int c = idx % ld; Wk[nl + idx] = c < n ? Zb[idx] : 0.0; }
    T.copy_wait(1);
    for (int r = T.rank; r < n; r += T.size) {
      const double *Tr = Tk + (size_t)r * ld;
      double v = su[r];
      for (int c = 0; c < n; c++) v += Tr[c] * (Ss[c] + zb[c]);
      Sn[r] = v;
    }
    // Rs <- R_k + Tm_k (Rs Y)       (nobody reads the old Rs any more)
    team_mm(T, n, n, n, Tk, ld, Zb, ld, [&](int r, int c, double acc) { Rs[(size_t)r * ld + c] = Rk[(size_t)r * ld + c] + acc; });
    T.sync();
    for (int r = T.rank; r < n; r += T.size) Ss[r] = Sn[r];
    if (k > 0) {
      // rotate the arrays: R_{k-1} sits in the old Y array, Tm_{k-1} in the spare one; the other three are free
      double *nR = Yc, *nY = Yn, *f1 = Rk, *f2 = Qc, *f3 = Qn;
      T.copy_wait(0); T.copy_wait(3); T.copy_wait(2);
      for (int r = T.rank; r < 2 * n; r += T.size) su[r] = sun[r];        // (su, sd) <- (sun, sdn)
      Rk = nR; Yc = nY; Qc = f1; Qn = f2; Yn = f3;
    }
    T.sync();
  }

  // ---- down-sweep, pass 1: the interface fields, a chain over the layers, I_dn(top_1) = 0:
  //          I_dn(bot_k) = Y_k I_dn(top_k) + y_k,   u(bot_k) = (Rs Y)_k I_dn(top_k) + (Rs y + Ss)_k
  //      each layer's block [Y | Rs Y | y | .] comes back from the team's scratch through two alternating copy slots; the layer's
  //      incoming fields minus the particular solution, in+ and in-, replace (y, .) in the scratch for pass 2
  const size_t slotd = (wl + 1) & ~(size_t)1;
  double *slot0 = pool, *slot1 = pool + slotd;                // 2 * slotd <= 5 nl
  for (int r = T.rank; r < n; r += T.size) xv[r] = 0.0;       // no diffuse downwelling at TOA
  T.copy_begin(0, slot0, Wg, wl);
  for (int k = 0; k < L; k++) {
    const size_t q = q0 + k;
    double *Wk = Wg + (size_t)k * wl;
    const double *sl = (k & 1) ? slot1 : slot0;
    if (k + 1 < L) T.copy_begin((k + 1) & 1, (k & 1) ? slot0 : slot1, Wg + (size_t)(k + 1) * wl, wl); else T.sync();
    T.copy_wait(k & 1);
    for (int item = T.rank; item < 2 * n; item += T.size) {
      const int which = item / n, r = item % n;
      const double *Mr = sl + (which ? nl : 0) + (size_t)r * ld;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int c = 0;
      for (; c + 3 < n; c += 4) { s0 += Mr[c] * xv[c]; s1 += Mr[c + 1] * xv[c + 1]; s2 += Mr[c + 2] * xv[c + 2]; s3 += Mr[c + 3] * xv[c + 3]; }
      for (; c < n; c++) s0 += Mr[c] * xv[c];
      const double v = (s0 + s1) + (s2 + s3) + sl[2 * nl + (which ? n : 0) + r];
      if (which) ub[r] = v; else xb[r] = v;
    }
    T.sync();
    const double *WU = Z.wupper + q * 2 * n, *WL = Z.wlower + q * 2 * n;
    for (int r = T.rank; r < n; r += T.size) {
      const double sg = ((r % S) > 1) ? -1.0 : 1.0;
      const double hd = xv[r] - WU[r], hu = ub[r] - sg * WL[n + r];
      Wk[2 * nl + r] = hd + hu; Wk[2 * nl + n + r] = hd - hu;
      xv[r] = xb[r];
    }
    T.sync();
  }
  // ---- pass 2a: LCON, MCON of every layer, all layers at once:  a + b = S+ in+,  a - b = S- in-   (results over the dead Y block)
  for (int item = T.rank; item < L * n; item += T.size) {
    const int k = item / n, e = item % n;
    if (!X.layscat[((size_t)p * (NM + 1) + m) * L + k]) continue;
    const size_t q = q0 + k;
    const double *SpT = O.spT + q * n * n, *SmT = O.smT + q * n * n;
    double *Wk = Wg + (size_t)k * wl;
    const double *ip = Wk + 2 * nl, *im = ip + n;
    double s0 = 0.0, s1 = 0.0, d0 = 0.0, d1 = 0.0;
    int r = 0;
    for (; r + 1 < n; r += 2) {
      s0 += SpT[(size_t)r * n + e] * ip[r]; s1 += SpT[(size_t)(r + 1) * n + e] * ip[r + 1];
      d0 += SmT[(size_t)r * n + e] * im[r]; d1 += SmT[(size_t)(r + 1) * n + e] * im[r + 1];
    }
    if (r < n) { s0 += SpT[(size_t)r * n + e] * ip[r]; d0 += SmT[(size_t)r * n + e] * im[r]; }
    const double ab = s0 + s1, amb = d0 + d1;
    Wk[e] = 0.5 * (ab + amb); Wk[n + e] = 0.5 * (ab - amb);              // LCON(e), MCON(e)
  }
  T.sync();
  // ---- pass 2b: layer source terms (WHOLELAYER_STERM_UP) accumulated top-down per view: slot = (view g, partition ep of the
  //      eigensolutions); every slot walks down the layers with its running user-stream transmittance
  const int EP = (T.size / G) > 0 ? T.size / G : 1, nslot = G * EP;
  double *red = Rs, *trn = red + (size_t)nslot * S;           // the matrix arrays are free now
  for (int slot = T.rank; slot < nslot; slot += T.size) {
    const int vg = slot % G, ep = slot / G, pv = v0 + vg;
    double acc[S], trans = 1.0;
#pragma unroll
    for (int o = 0; o < S; o++) acc[o] = 0.0;
    const bool vact = CV.state[(size_t)pv * 4] == 1;          // views that have converged keep their sums
    for (int k = 0; k < L; k++) {
      const size_t q = q0 + k, qv = q * G + vg;
      const double udel = X.tduserm[pixL(X, pv) + k];
      if (vact && ep < n && X.layscat[((size_t)p * (NM + 1) + m) * L + k]) {
        const double *Wk = Wg + (size_t)k * wl;
        const bool beam = (k + 1 <= cutoff);
        const double wdel = X.tdmubar[pixL(X, p) + k], itr = X.itrans[pixL(X, p) + k], itw = -itr * wdel, dl = X.deltau[pixL(X, p) + k];
        const double emu = X.emult_up[pixL(X, pv) + k], sgp = X.sigma_p[pixL(X, pv) + k], itu = X.itrans_userm[pixL(X, pv) + k];
        double part[S];
#pragma unroll
        for (int o = 0; o < S; o++) part[o] = 0.0;
        for (int e = ep; e < n; e += EP) {
          const double lc = Wk[e], mc = Wk[n + e];
          const double h1 = Z.hm1[qv * n + e], h2 = Z.hm2[qv * n + e];
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
          for (int o = 0; o < S; o++) part[o] += cd * Z.uh_dn[(qv * n + e) * S + o] + cu * Z.uh_up[(qv * n + e) * S + o];
        }
#pragma unroll
        for (int o = 0; o < S; o++) acc[o] += trans * part[o];
      }
      trans *= udel;
    }
#pragma unroll
    for (int o = 0; o < S; o++) red[(size_t)slot * S + o] = acc[o];
    trn[slot] = trans;
  }
  if (include_surface) for (int r = T.rank; r < n; r += T.size) { const int i = r / S; inp[r] = QD.x[i] * QD.w[i] * xv[r]; }
  T.sync();
  // ---- per view: Fourier summation and convergence (VLIDORT_CONVERGE_OBSGEO); FLUX_MULTIPLIER = DELTA_FACTOR
  for (int vg = T.rank; vg < G; vg += T.size) {
    const int pv = v0 + vg;
    int *cst = CV.state + (size_t)pv * 4;
    if (cst[0] != 1) continue;
    double cum[S], boa[S];
#pragma unroll
    for (int o = 0; o < S; o++) { cum[o] = 0.0; boa[o] = 0.0; }
    for (int t = 0; t < EP; t++) {
#pragma unroll
      for (int o = 0; o < S; o++) cum[o] += red[(size_t)(vg + G * t) * S + o];
    }
    const double trans = trn[vg];
    if (include_surface) {
      if (lambertian) {
        double rr = 0.0;
        for (int i = 0; i < N; i++) rr += inp[S * i];
        boa[0] = SFAC * albedo * rr;
      } else {
        const double *UBF = SF.user_brdf_f + ((size_t)pv * SF.nmfour + m) * SF.nq * N;
#pragma unroll
        for (int o1 = 0; o1 < S; o1++) {
          double rr = 0.0;
          for (int jj = 0; jj < N; jj++)
            for (int o2 = 0; o2 < S; o2++)
              rr += inp[S * jj + o2] * (nq == 16 ? UBF[(size_t)(4 * o1 + o2) * N + jj] : ((o1 | o2) == 0 ? UBF[jj] : 0.0));
          boa[o1] = SFAC * rr;
        }
      }
    }
    const double *scv = X.scal + (size_t)pv * 16;
    double *tot = CV.total + (size_t)pv * 4;
    double sf[S];
#pragma unroll
    for (int o = 0; o < S; o++) sf[o] = DELTA_FACTOR * (cum[o] + trans * boa[o]);
    if (m == 0) {
#pragma unroll
      for (int o = 0; o < S; o++) tot[o] = sf[o] + scv[6 + o] + scv[9 + o];
      cst[1] = 0; cst[2] = 0;
      if (NM == 0) cst[0] = 0;
    } else {
      const double arg = scv[12] * (double)m * DEG_TO_RAD, cf = cos(arg), sf_ = sin(arg);
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
  T.sync();
  // the pixel stays in the Fourier loop while any of its views does
  if (T.rank == 0) {
    int any = 0;
    for (int g = 0; g < G; g++) any |= (CV.state[(size_t)(v0 + g) * 4] == 1);
    CV.col_active[p] = any;
  }
}

}  // namespace vlk
