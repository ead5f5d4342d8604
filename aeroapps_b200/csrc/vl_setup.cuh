// Stage B, kernel K1: per pixel-channel set-up.  One Team (a CTA on the GPU) per pixel.
//   glue     V283/VLIDORT_ScatMod.F90:284-617,975-1247  (aerosol (+) Rayleigh mixing, float32 literals kept)
//   check    vlidort_inputs.f90:3396-3460,3692-3821      (-> per-pixel status 3 == rc 3 of the glue)
//   Chapman  vlidort_geometry.f90:119-700
//   delta-M  vlidort_miscsetups.f90:370-812 ; QSPREP :813-1013 ; PREPTRANS :1014-1295 ; EMULT :3187-3424
//   FO       vlidort_vfo_interface.f90:128-728 + vlidort_focode (see oracle/vl_fo.h for the line map)
#pragma once
#include "vl_common.cuh"

namespace vlk {

struct VLIn {                 // device pointers in the reference's Fortran layouts
  int km, nch, nobs, nMom, nPol, S, N, nfine;
  int planpar, lambertian, scalar_glue;
  int nadir_from_obs;         // first obs index that runs with FOCORR_NADIR (reference state quirk), nobs if none
  int ngeom;                  // view geometries per observation (>= 1): pixel-view pv = (j*nch + i)*ngeom + g
  double MISSING, earth_radius;
  const double *ROT;          // (km,nobs,nch)
  const double *depol;        // (nch)
  const double *alpha;        // (km,nobs,nch) or null
  const double *tau, *ssa;    // (km,nch,nobs)
  const double *pmom;         // (km,nch,nobs,nMom,nPol)
  // cloud variants (VLIDORT_Run_Vector_Cloud, V283/VLIDORT_ScatMod.F90:1865-2220): ice and liquid cloud optical depth and single
  // scattering albedo (km,nch,nobs) and the P11 moments (km,nch,nobs,nMom) -- the reference mixes in the first element of
  // pmomI / pmomL only (:2125-2131); all six null for the clear-sky entry points
  const double *tauI, *ssaI, *pmomI, *tauL, *ssaL, *pmomL;
  const double *ze;           // (km+1,nobs) [m]
  const double *albedo;       // (nobs,nch)   (Lambertian)
  const double *sza;          // (nobs): one sun per observation
  const double *raa, *vza;    // (nobs,ngeom)
  const double *flux_factor;  // (nch,nobs) or null (=1)
  const double *exactdb;      // (npix*ngeom,16) exact direct-bounce BRDF per pixel-view (non-Lambertian) or null
  const unsigned char *skip;  // (npix*ngeom) 1 = pixel-view not run (MISSING inputs), 2 = not run, outputs -500 (RTLS BR < 0); may be null
};

// per-pixel state produced by K1 and consumed by K2/K3 (structure of arrays).  A pixel c = (observation, channel) owns G view
// geometries that share the sun, the atmosphere and the surface; everything that does not depend on the viewing direction is
// stored per pixel [c], the rest per pixel-view [v] with v = c*G + g.  The single-geometry entry points have G = 1 (v == c).
struct VLPix {
  int L, NM, S, npix, G;
  double *deltau;      // [c][L]   delta-M scaled
  double *og;          // [c][(NM+1)][L][4]  omega' * (g11,g12,g22,g33)'
  double *avsec, *itrans, *tdmubar;                      // [c][L]   solar beam
  double *tduserm, *emult_up, *sigma_p, *itrans_userm;   // [v][L]   user stream
  double *scal;        // [v][16]: 0 cos_sza 1 ustream 2 usec 3 trans_solar_beam 4 albedo 5 flux 6.. fo_ss(3) 9.. fo_db(3) 12 raa
  int *ibuf;           // [v][4]: 0 beam_cutoff (1-based) 1 status 2 do_reflected_directbeam
  unsigned char *layscat;   // [c][(NM+1)][L]
};

DEVFN size_t pixL(const VLPix &X, int p) { return (size_t)p * X.L; }                 // p: pixel for the [c] arrays, pixel-view for the [v] arrays
DEVFN double *og_ptr(const VLPix &X, int p) { return X.og + (size_t)p * (X.NM + 1) * X.L * 4; }
DEVFN int view0(const VLPix &X, int c) { return c * X.G; }                           // first pixel-view of pixel c

// Sum_k ext[k-1]*sunpath(k) for a direct (FindSunPaths_D) or tangent (FindSunPaths_T) solar path that starts at
// radius Radstart inside layer N (1-based), local solar zenith thstart.  radii[0..L].
DEVFN double sun_tau(bool direct, bool zeroSunBOA, double Radstart, const double *radii, double thstart, double sthstart,
                     int N, int L, const double *ext) {
  const int N1 = N - 1;
  double sum = 0.0;
  if (N == 0 && direct) return 0.0;
  if (direct) {
    if (zeroSunBOA) {
      sum = ext[N - 1] * (radii[N1] - Radstart);
      for (int k = N1; k >= 1; k--) sum += ext[k - 1] * (radii[k - 1] - radii[k]);
      return sum;
    }
    double sth0 = sthstart, th0 = thstart;
    double sth1 = sth0 * Radstart / radii[N1], th1 = asin(sth1), ks1 = th0 - th1;
    sum = ext[N - 1] * (sin(ks1) * Radstart / sth1);
    sth0 = sth1; th0 = th1;
    for (int k = N1; k >= 1; k--) {
      sth1 = sth0 * radii[k] / radii[k - 1]; th1 = asin(sth1); ks1 = th0 - th1;
      sum += ext[k - 1] * (sin(ks1) * radii[k] / sth1);
      sth0 = sth1; th0 = th1;
    }
    return sum;
  }
  // tangent path
  const double tanr = sthstart * Radstart;
  int k = N1 < 0 ? 0 : N1; bool trawl = true;
  while (trawl) { trawl = (k <= L) && (radii[k] > tanr); k++; }
  const int nt = k - 1;
  double sth0, th0, sth1, th1, ks1, pN = 0.0;
  if (nt > N) {
    th0 = PIE - thstart; sth0 = sthstart;
    sth1 = sth0 * Radstart / radii[N]; th1 = asin(sth1); ks1 = th0 - th1;
    pN = 2.0 * sin(ks1) * Radstart / sth1;
    sth0 = sth1; th0 = th1;
    for (k = N + 1; k <= nt - 1; k++) {
      sth1 = sth0 * radii[k - 1] / radii[k]; th1 = asin(sth1); ks1 = th0 - th1;
      sum += ext[k - 1] * (2.0 * sin(ks1) * radii[k] / sth0);
      sth0 = sth1; th0 = th1;
    }
    ks1 = 0.5 * PIE - th0;
    if (nt >= 1 && nt <= L) sum += ext[nt - 1] * (2.0 * sin(ks1) * radii[nt - 1]);
  } else if (nt == N) {
    pN = -2.0 * Radstart * cos(thstart);
  }
  if (N >= 1) {
    th0 = PIE - thstart; sth0 = sthstart;
    sth1 = sth0 * Radstart / radii[N1]; th1 = asin(sth1); ks1 = th0 - th1;
    pN += sin(ks1) * Radstart / sth1;
    sum += ext[N - 1] * pN;
    sth0 = sth1; th0 = th1;
    for (k = N1; k >= 1; k--) {
      sth1 = sth0 * radii[k] / radii[k - 1]; th1 = asin(sth1); ks1 = th0 - th1;
      sum += ext[k - 1] * (sin(ks1) * radii[k] / sth1);
      sth0 = sth1; th0 = th1;
    }
  }
  return sum;
}

DEVFN void find_sun(bool doNadir, bool overhead, double Radius, const double *SD, double CumAngle, double theta_boa_R,
                    double &theta, double &stheta, double &ctheta, bool &DirSun) {
  if (doNadir) { DirSun = true; theta = theta_boa_R; ctheta = cos(theta_boa_R); stheta = sin(theta_boa_R); return; }
  if (overhead) { DirSun = true; ctheta = cos(CumAngle); stheta = sin(CumAngle); theta = CumAngle; return; }
  const double b = -Radius * sin(CumAngle) * SD[0] + Radius * cos(CumAngle) * SD[2];
  ctheta = -b / Radius;
  DirSun = (ctheta >= 0.0);
  stheta = sqrt(1.0 - ctheta * ctheta);
  theta = acos(ctheta);
}

// 3-point Gauss-Legendre abscissae/weights on (0,1) are not hard-wired: nfine is an input (reference: 3)
HDFN void gauss01(int n, double *x, double *w) {           // GETQUAD2 (FO_geometry_Generic.f90:331-426) on (0,1)
  const int N2 = (n + 1) / 2; const double NR = n;
  for (int M = 1; M <= N2; M++) {
    const double XX = PIE * ((double)M - 0.25) / (NR + 0.5), s = sin(XX);
    double X = (1.0 - (NR - 1.0) / (8.0 * NR * NR * NR) - 1.0 / (384.0 * NR * NR * NR * NR) * (39.0 - 28.0 / (s * s))) * cos(XX);
    double LP = 1.0, LPM1 = 0.0, D = 1.0;
    for (int it = 0; it < 100; it++) {
      LPM1 = 0.0; LP = 1.0;
      for (int I = 1; I <= n; I++) { const double IR = I, LPM2 = LPM1; LPM1 = LP; LP = ((2.0 * IR - 1.0) * X * LPM1 - (IR - 1.0) * LPM2) / IR; }
      D = NR * (X * LP - LPM1) / (X * X - 1.0);
      const double XOLD = X; X = XOLD - LP / D;
      if (fabs(X - XOLD) <= 1.0e-13) break;
    }
    x[M - 1] = 0.5 - 0.5 * X; w[M - 1] = 1.0 / ((1.0 - X * X) * D * D);
    x[n - M] = 0.5 + 0.5 * X; w[n - M] = w[M - 1];
  }
}

// scratch (doubles): 12*(L+1) + 2*(nMom+1) + 2*nfine
template <class Team>
DEVFN void setup_pixel(const Team &T, const VLIn &in, const VLPix &X, int p, double *scr) {
  // p = pixel-view; the views of a pixel all write the same values to the pixel's [c] arrays (every view needs them here anyway)
  const int L = in.km, S = in.S, N = in.N, nMom = in.nMom, NM = X.NM;
  const int G = in.ngeom > 0 ? in.ngeom : 1, pc = p / G, g = p % G;
  const int j = pc / in.nch, i = pc % in.nch;
  int *ib = X.ibuf + (size_t)p * 4;
  double *sc = X.scal + (size_t)p * 16;
  if (in.skip && in.skip[p]) { if (T.rank == 0) ib[1] = (in.skip[p] == 2) ? 98 : 99; return; }
  // scratch carve-up
  double *H = scr;                    // heights [km], 0..L
  double *din = H + (L + 1);          // deltau input
  double *win = din + (L + 1);        // omega input
  double *fac1 = win + (L + 1);
  double *trunc = fac1 + (L + 1);
  double *tslant = trunc + (L + 1);   // 0..L
  double *ext = tslant + (L + 1);
  double *atten = ext + (L + 1);      // 0..L
  double *stau = atten + (L + 1);     // 0..L
  double *ltr = stau + (L + 1);       // lostrans 1..L
  double *mult = ltr + (L + 1);
  double *flag = mult + (L + 1);      // flag[0] = bad-input counter
  double *gs1 = flag + (L + 1);
  double *gs2 = gs1 + (nMom + 1);
  double *xf = gs2 + (nMom + 1);
  double *wf = xf + (in.nfine > 0 ? in.nfine : 1);

  const double sza = in.sza[j], vza = in.vza[j + (size_t)in.nobs * g];
  double raa = in.raa[j + (size_t)in.nobs * g];
  const double depol = in.depol[i];
  const double FLUX = in.flux_factor ? in.flux_factor[i + (size_t)in.nch * j] : 1.0;
  const double albedo = in.lambertian ? in.albedo[j + (size_t)in.nobs * i] : 0.0;
  const double cos_sza = cos(sza * DEG_TO_RAD), ustream = cos(DEG_TO_RAD * vza), usec = 1.0 / ustream;
  const double *pm = in.pmom + (size_t)L * (i + (size_t)in.nch * j);       // + k + stride_m*(m + nMom*pol)
  const bool cloud = in.tauI != nullptr;
  const double *pmL = cloud ? in.pmomL + (size_t)L * (i + (size_t)in.nch * j) : nullptr;      // P11 only: + k + stride_m*m
  const double *pmI = cloud ? in.pmomI + (size_t)L * (i + (size_t)in.nch * j) : nullptr;
  const size_t stride_m = (size_t)L * in.nch * in.nobs;
  const double f32_1em3 = (double)1.0e-3f, w_clamp = (double)0.99999f;
  // Rayleigh Greek constants (VLIDORT_ScatMod.F90:456-485); SQRT(6.) is single precision there
  const double gammamom2 = -(double)2.4494898f * (1.0 - depol) / (2.0 + depol);   // sqrtf(6.f)
  const double alphamom2 = 6.0 * (1.0 - depol) / (2.0 + depol);
  const double raysmom2 = (1.0 - depol) / (2.0 + depol);

  if (T.rank == 0) flag[0] = 0.0;
  for (int k = T.rank; k <= L; k += T.size) H[k] = in.ze[k + (size_t)(L + 1) * j] * f32_1em3;
  T.sync();
  // ---- glue: layer optical properties
  double *ogp = og_ptr(X, pc);
  for (int k = T.rank; k < L; k += T.size) {
    const double ray = in.ROT[k + (size_t)L * (j + (size_t)in.nobs * i)];
    const double alp = in.alpha ? in.alpha[k + (size_t)L * (j + (size_t)in.nobs * i)] : 0.0;
    const double tau = in.tau[k + (size_t)L * (i + (size_t)in.nch * j)], ssa = in.ssa[k + (size_t)L * (i + (size_t)in.nch * j)];
    const size_t okc = k + (size_t)L * (i + (size_t)in.nch * j);
    const double tI = cloud ? in.tauI[okc] : 0.0, wI = cloud ? in.ssaI[okc] : 0.0, tL = cloud ? in.tauL[okc] : 0.0, wL = cloud ? in.ssaL[okc] : 0.0;
    const double tau_ext = cloud ? alp + ray + tau + tL + tI : alp + ray + tau;
    const double tau_scat = cloud ? ray + ssa * tau + tL * wL + tI * wI : ray + ssa * tau;
    double w = tau_scat / tau_ext;
    if (w > w_clamp) w = w_clamp;
    din[k] = tau_ext; win[k] = w;
    const double aerswt = ssa * tau / tau_scat, rayswt = ray / tau_scat;
    const double clLswt = wL * tL / tau_scat, clIswt = wI * tI / tau_scat;
    // input checks (VLIDORT_CHECK_INPUT_OPTICAL); greek(0,k,1) is forced to 1 by the glue
    if (!(tau_ext > 0.0) || !(w <= 1.0 - OMEGA_SMALLNUM) || !(w >= OMEGA_SMALLNUM)) flag[0] = 1.0;   // benign race: any writer sets 1
    // delta-M truncation factor from moment 2N of the mixed input
    const int nm1 = NM + 1;
    double g2n = 0.0;
    if (nm1 <= nMom) {
      const double aer = (nm1 <= nMom - 1) ? pm[k + stride_m * (size_t)(nm1 + (size_t)nMom * 0)] : 0.0;
      g2n = aer * aerswt + (nm1 == 2 ? raysmom2 * rayswt : 0.0);
      if (cloud && nm1 <= nMom - 1) g2n += pmL[k + stride_m * (size_t)nm1] * clLswt + pmI[k + stride_m * (size_t)nm1] * clIswt;
      if (nm1 == 0) g2n = 1.0;
    }
    const double fdel = g2n / (double)(2 * nm1 + 1), fac2 = 1.0 - fdel;
    fac1[k] = 1.0 - fdel * w; trunc[k] = fdel;
    double dsc = tau_ext * fac1[k], wsc = w * fac2 / fac1[k];
    if (wsc > 0.999999999) wsc = 0.999999999; else if (wsc < 1.0e-9) wsc = 1.0e-9;
    X.deltau[pixL(X, pc) + k] = dsc;
    unsigned char *ls = X.layscat + (size_t)pc * (NM + 1) * L;
    int nzero_tail = 0;                          // number of zero g11 input moments in l = m..NM, built from the top
    for (int l = NM; l >= 0; l--) {
      double a1 = 0, a2 = 0, a5 = 0, a3 = 0;     // aerosol P11, P12, P22, P33 moments (index l+1 in pmom)
      if (l <= nMom - 1) {
        a1 = pm[k + stride_m * (size_t)(l + (size_t)nMom * 0)];
        if (S > 1 && !in.scalar_glue) {
          a2 = pm[k + stride_m * (size_t)(l + (size_t)nMom * 1)];
          a3 = pm[k + stride_m * (size_t)(l + (size_t)nMom * 2)];
          a5 = pm[k + stride_m * (size_t)(l + (size_t)nMom * 4)];
        }
      }
      double g11 = a1 * aerswt, g12 = a2 * aerswt, g22 = a5 * aerswt, g33 = a3 * aerswt;
      if (cloud && l <= nMom - 1) g11 += pmL[k + stride_m * (size_t)l] * clLswt + pmI[k + stride_m * (size_t)l] * clIswt;
      if (l == 2) { g11 += raysmom2 * rayswt; if (S > 1) { g12 += gammamom2 * rayswt; g22 += alphamom2 * rayswt; } }
      if (l == 0) g11 = 1.0;                      // greekmat_total_input(0,i,1) = 1.0
      if (g11 == 0.0) nzero_tail++;
      if (l >= 3) ls[(size_t)l * L + k] = (nzero_tail < NM - l + 1) ? 1 : 0; else ls[(size_t)l * L + k] = 1;
      // delta-M (vlidort_miscsetups.f90:480-525) + singularity buster on g(1) (:624-656)
      const double fd = fdel * (double)(2 * l + 1);
      double s11 = (g11 - fd) / fac2, s12 = g12 / fac2, s22 = (g22 - fd) / fac2, s33 = (g33 - fd) / fac2;
      if (l == 0) s11 = 1.0;
      if (l == 1) {
        double g = s11 / 3.0;
        if (g > 0.999999999) g = 0.999999999; else if (g < -0.999999999) g = -0.999999999;
        else if (g >= 0.0 && g < 1.0e-9) g = 1.0e-9; else if (g < 0.0 && g > -1.0e-9) g = -1.0e-9;
        s11 = 3.0 * g;
      }
      double *o4 = ogp + ((size_t)l * L + k) * 4;
      o4[0] = wsc * s11; o4[1] = wsc * s12; o4[2] = wsc * s22; o4[3] = wsc * s33;
    }
  }
  T.sync();
  // angle checks (vlidort_inputs.f90:3396-3460)
  const bool bad_angles = !(sza >= 0.0 && sza < 90.0) || !(raa >= 0.0 && raa <= 360.0) || !(vza >= 0.0 && vza <= 90.0) ||
                          (in.lambertian && !(albedo >= 0.0 && albedo <= 1.0));
  if (flag[0] != 0.0 || bad_angles) { if (T.rank == 0) ib[1] = 3; return; }

  // ---- Chapman factors folded into the slant optical depth to each level (scaled): tslant[n], n=1..L
  const double mu_toa = cos_sza, sn_toa = sqrt(1.0 - mu_toa * mu_toa);
  for (int nn = T.rank + 1; nn <= L; nn += T.size) {
    double t = 0.0;
    if (in.planpar) { for (int k = 1; k <= nn; k++) t += din[k - 1] * fac1[k - 1] / mu_toa; }
    else if (sn_toa == 0.0) { for (int k = 1; k <= nn; k++) t += din[k - 1] * fac1[k - 1]; }
    else {
      const double r0 = H[0] + in.earth_radius;
      double sinth1 = sn_toa * (H[nn] + in.earth_radius) / r0, sth1 = asin(sinth1), re_upper = r0;
      for (int k = 1; k <= nn; k++) {
        const double delz = H[k - 1] - H[k], re_lower = re_upper - delz;
        const double sinth2 = re_upper * sinth1 / re_lower, sth2 = asin(sinth2);
        const double dist = re_upper * sin(sth2 - sth1) / sinth2;
        t += din[k - 1] * (dist / delz) * fac1[k - 1];
        re_upper = re_lower; sinth1 = sinth2; sth1 = sth2;
      }
    }
    tslant[nn] = t;
  }
  if (T.rank == 0) tslant[0] = 0.0;
  T.sync();
  // ---- QSPREP (sequential in the layer index)
  double *avsec = X.avsec + pixL(X, pc), *itrans = X.itrans + pixL(X, pc), *tdmubar = X.tdmubar + pixL(X, pc);
  double *tduserm = X.tduserm + pixL(X, p), *emult = X.emult_up + pixL(X, p), *sigp = X.sigma_p + pixL(X, p);
  double *itu = X.itrans_userm + pixL(X, p);
  const double *dsc = X.deltau + pixL(X, pc);
  if (T.rank == 0) {
    int cutoff = L;
    if (in.planpar) {
      const double sec0 = 1.0 / cos_sza;
      double tg = 0.0;                            // TAUGRID(n-1)
      for (int nn = 1; nn <= L; nn++) {
        const double tgn = tg + dsc[nn - 1];
        tslant[nn] = tgn * sec0;
        if (nn <= cutoff) {
          if (tslant[nn] > MAX_TAU_PATH) cutoff = nn;
          avsec[nn - 1] = sec0; itrans[nn - 1] = exp(-tg * sec0);
        } else { avsec[nn - 1] = 0.0; itrans[nn - 1] = 0.0; }
        tg = tgn;
      }
    } else {
      double s_t_1 = 0.0, s_t_0 = 1.0;
      for (int nn = 1; nn <= L; nn++) {
        if (nn <= cutoff) {
          if (tslant[nn] > MAX_TAU_PATH) cutoff = nn; else s_t_1 = exp(-tslant[nn]);
          avsec[nn - 1] = (tslant[nn] - tslant[nn - 1]) / dsc[nn - 1];
          itrans[nn - 1] = s_t_0; s_t_0 = s_t_1;
        } else { avsec[nn - 1] = 0.0; itrans[nn - 1] = 0.0; }
      }
    }
    const double tau_solar = tslant[L];
    ib[0] = cutoff; ib[1] = 0;
    if (tau_solar < MAX_TAU_PATH) { sc[3] = exp(-tau_solar); ib[2] = 1; } else { sc[3] = 0.0; ib[2] = 0; }
    sc[0] = cos_sza; sc[1] = ustream; sc[2] = usec; sc[4] = albedo; sc[5] = FLUX; sc[12] = raa;
  }
  T.sync();
  // ---- PREPTRANS + EMULT (obsgeo, upwelling)
  {
    const int cutoff = ib[0];
    for (int k = T.rank; k < L; k += T.size) {
      double tm = 0.0;
      if (k + 1 <= cutoff) { const double sp = dsc[k] * avsec[k]; tm = sp > MAX_TAU_PATH ? 0.0 : exp(-sp); }
      tdmubar[k] = tm;
      itu[k] = itrans[k] * usec;
      const double spu = dsc[k] * usec, tu = spu > MAX_TAU_PATH ? 0.0 : exp(-spu);
      tduserm[k] = tu;
      sigp[k] = avsec[k] + usec;
      emult[k] = (k + 1 <= cutoff) ? itu[k] * (1.0 - tm * tu) / sigp[k] : 0.0;
    }
  }
  // ---- FO: exact single scatter + direct bounce at TOA
  const int mode = in.planpar ? 0 : ((sza == 0.0 || j >= in.nadir_from_obs) ? 1 : 2);
  if (raa < 0.0) raa = -raa;
  if (raa > 360.0) raa = 360.0 - raa - 360.0;
  const double alpha_boa_R = vza * DEG_TO_RAD, theta_boa_R = sza * DEG_TO_RAD;
  double calpha_boa, salpha_boa, stheta_boa, ctheta_boa;
  if (vza == 90.0) { calpha_boa = 0.0; salpha_boa = 1.0; } else { calpha_boa = cos(alpha_boa_R); salpha_boa = sin(alpha_boa_R); }
  stheta_boa = sin(theta_boa_R); ctheta_boa = cos(theta_boa_R);
  const double cphi_boa = cos(raa * DEG_TO_RAD), sphi_boa = sin(raa * DEG_TO_RAD);
  const bool overhead = (sza == 0.0);
  const bool doNadir = (mode == 2) && (vza == 0.0);
  double cosscat;
  if (overhead) { cosscat = -calpha_boa; if (calpha_boa == 0.0) cosscat = calpha_boa; }
  else cosscat = -calpha_boa * ctheta_boa + salpha_boa * stheta_boa * cphi_boa;
  const double Raycon = salpha_boa * (in.earth_radius + H[L]);
  double SD[3] = {0, 0, 0};
  if (!overhead) { SD[0] = -stheta_boa * cphi_boa; SD[1] = -stheta_boa * sphi_boa; SD[2] = -ctheta_boa; }
  // radii in place of heights from here on: keep H (heights) and build radii in 'mult' temporarily
  double *radii = mult;                           // 0..L (the layer multipliers themselves go to tslant, free by now)
  for (int k = T.rank; k <= L; k += T.size) radii[k] = in.earth_radius + H[k];
  for (int k = T.rank; k < L; k += T.size) ext[k] = dsc[k] / (H[k] - H[k + 1]);
  if (T.rank == 0 && mode == 2 && in.nfine > 0) gauss01(in.nfine, xf, wf);
  // generalized spherical functions at the scattering angle: sequential recurrence (one lane), l = 0..nMom
  if (T.rank == (T.size > 1 ? 1 : 0)) {
    const double mus = cosscat;
    double p00n = 0.0, p00o = 1.0, p02n = 0.0, p02o = 0.0, sql41 = 0.0;
    const double QROOT6 = -0.25 * sqrt(6.0);
    for (int l = 0; l <= nMom; l++) {
      // (LOLD,LNEW) swap of the reference expressed with explicit old/new scalars
      if (l == 0) { gs1[0] = 1.0; gs2[0] = 0.0; p00o = 1.0; p00n = 0.0; }
      else if (l == 1) { gs1[1] = mus; gs2[1] = 0.0; p00n = p00o; p00o = mus; }
      else {
        const double dl = l, g1 = (2.0 * dl - 1.0) / dl, g2 = (dl - 1.0) / dl;
        const double v = g1 * mus * p00o - g2 * p00n; p00n = p00o; p00o = v; gs1[l] = v;
        if (l == 2) { p02o = QROOT6 * (1.0 - mus) * (1.0 + mus); p02n = 0.0; sql41 = 0.0; gs2[2] = p02o; }
        else {
          const double sql4 = sql41; sql41 = sqrt(dl * dl - 4.0);
          const double g3 = (2.0 * dl - 1.0) / sql41, g4 = sql4 / sql41;
          const double u = g3 * mus * p02o - g4 * p02n; p02n = p02o; p02o = u; gs2[l] = u;
        }
      }
    }
  }
  T.sync();
  // level attenuations (and LOS angles for the enhanced-PS mode)
  for (int nn = T.rank; nn <= L; nn += T.size) {
    double sumd = 0.0;
    if (mode == 0) { for (int k = 1; k <= nn; k++) sumd += ext[k - 1] * (H[k - 1] - H[k]) / ctheta_boa; }
    else if (mode == 1) { if (nn >= 1) sumd = sun_tau(true, overhead, radii[nn], radii, theta_boa_R, stheta_boa, nn, L, ext); }
    else {
      double th, st, ct; bool ds;
      if (nn == L) { th = theta_boa_R; st = stheta_boa; ct = ctheta_boa; ds = true; }
      else {
        const double alpha_n = doNadir ? 0.0 : asin(Raycon / radii[nn]);
        find_sun(doNadir, overhead, radii[nn], SD, alpha_boa_R - alpha_n, theta_boa_R, th, st, ct, ds);
      }
      const bool zsb = overhead && (nn == L || doNadir);
      if (nn >= 1) sumd = sun_tau(ds, zsb, radii[nn], radii, th, st, nn, L, ext);   // level 0 is not used in this mode
    }
    stau[nn] = sumd; atten[nn] = sumd < 88.0 ? exp(-sumd) : 0.0;
  }
  T.sync();
  for (int nn = T.rank + 1; nn <= L; nn += T.size) {
    double m_ = 0.0, lt = 0.0;
    if (mode != 2) {
      if (calpha_boa > 0.0) {
        const double lostau = dsc[nn - 1] / calpha_boa;
        if (lostau < 88.0) lt = exp(-lostau);
        const double f1 = atten[nn - 1] - atten[nn] * lt, f2 = 1.0 + (stau[nn] - stau[nn - 1]) / lostau;
        m_ = f1 / f2;
      }
    } else {
      double LosW, rcn = 0.0;
      if (doNadir) LosW = radii[nn - 1] - radii[nn];
      else {
        const double a0 = asin(Raycon / radii[nn - 1]), a1 = (nn == L) ? alpha_boa_R : asin(Raycon / radii[nn]);
        const double c0 = cos(a0), c1 = (nn == L) ? calpha_boa : cos(a1);
        LosW = radii[nn - 1] * c0 - radii[nn] * c1; rcn = radii[nn] * c1;
      }
      const double kn = ext[nn - 1], lostau = kn * LosW;
      if (lostau < 88.0) lt = exp(-lostau);
      double sum = 0.0;
      for (int jf = 0; jf < in.nfine; jf++) {
        const double tf = xf[jf] * LosW, af = wf[jf] * LosW;
        double rf, alf;
        if (doNadir) { rf = radii[nn] + tf; alf = 0.0; }
        else { const double tanf = Raycon / (rcn + tf); alf = atan(tanf); rf = Raycon / sin(alf); }
        double th, st, ct; bool ds;
        find_sun(doNadir, overhead, rf, SD, alpha_boa_R - alf, theta_boa_R, th, st, ct, ds);
        const bool zsb = overhead && (nn == L || doNadir);
        const double sumd = sun_tau(ds, zsb, rf, radii, th, st, nn, L, ext);
        const double att = sumd < 88.0 ? exp(-sumd) : 0.0;
        sum += att * exp(-kn * (LosW - tf)) * af;
      }
      m_ = sum * kn;
    }
    ltr[nn] = lt; tslant[nn] = m_;                 // tslant reused as the layer multiplier array
  }
  T.sync();
  // exact scattering per layer and the upward recursion (FO_VectorSS_RTCalcs_I.f90:419-431,697-760)
  double *ex0 = stau;                               // reuse: ex0[n], and two more arrays for Q,U parts
  double *ex1 = H, *ex2 = din;                      // heights / input deltau no longer needed
  {
    // rotation angles (FO_VectorSS_Spherfuncs.f90:205-260)
    const double SMALL = 1.0e-12, mus = cosscat;
    const double CTH = cos(DEG_TO_RAD * sza), STH = sin(DEG_TO_RAD * sza), CAL = cos(DEG_TO_RAD * vza), SAL = sin(DEG_TO_RAD * vza);
    const double CPHI = cos(DEG_TO_RAD * raa);
    const double hs = 1.0 - mus * mus, sinscat = hs <= 0.0 ? SMALL : sqrt(hs);
    double cs2;
    if (fabs(sinscat) <= SMALL) cs2 = 0.0;
    else cs2 = (SAL == 0.0) ? -CPHI : (-CTH - CAL * mus) / sinscat / SAL;
    (void)STH;
    cs2 = fmin(1.0, fmax(-1.0, cs2));
    const double ss2 = fabs(cs2 - 1.0) < SMALL ? 0.0 : sqrt(1.0 - cs2 * cs2);
    const double C2 = 2.0 * cs2 * cs2 - 1.0;
    const double S2 = (raa <= 180.0) ? 2.0 * cs2 * ss2 : -2.0 * cs2 * ss2;
    T.sync();
    for (int nn = T.rank + 1; nn <= L; nn += T.size) {
      const int k = nn - 1;
      const double ray = in.ROT[k + (size_t)L * (j + (size_t)in.nobs * i)];
      const double tau = in.tau[k + (size_t)L * (i + (size_t)in.nch * j)], ssa = in.ssa[k + (size_t)L * (i + (size_t)in.nch * j)];
      const size_t okc = k + (size_t)L * (i + (size_t)in.nch * j);
      const double tI = cloud ? in.tauI[okc] : 0.0, wI = cloud ? in.ssaI[okc] : 0.0, tL = cloud ? in.tauL[okc] : 0.0, wL = cloud ? in.ssaL[okc] : 0.0;
      const double tau_scat = cloud ? ray + ssa * tau + tL * wL + tI * wI : ray + ssa * tau;
      const double aerswt = ssa * tau / tau_scat, rayswt = ray / tau_scat;
      double f1 = 0.0, f2 = 0.0, fL = 0.0, fI = 0.0;
      for (int l = 0; l <= nMom - 1; l++) {
        f1 += gs1[l] * pm[k + stride_m * (size_t)(l + (size_t)nMom * 0)];
        if (S > 1 && !in.scalar_glue) f2 += gs2[l] * pm[k + stride_m * (size_t)(l + (size_t)nMom * 1)];
        if (cloud && l >= 1) { fL += gs1[l] * pmL[k + stride_m * (size_t)l]; fI += gs1[l] * pmI[k + stride_m * (size_t)l]; }
      }
      // l = 0 term of the mixed input is forced to 1; Rayleigh adds l = 2 (and nothing else for 11/12)
      const double a0 = (nMom >= 1) ? pm[k + stride_m * (size_t)(0)] : 0.0;
      f1 = (f1 - gs1[0] * a0) * aerswt + gs1[0] * 1.0;
      if (cloud) f1 += fL * (wL * tL / tau_scat) + fI * (wI * tI / tau_scat);
      f2 = f2 * aerswt;
      if (nMom >= 2) { f1 += gs1[2] * raysmom2 * rayswt; if (S > 1) f2 += gs2[2] * gammamom2 * rayswt; }
      const double tms = win[k] / (1.0 - trunc[k] * win[k]);
      ex0[nn] = f1 * tms * tslant[nn];
      if (S > 1) { ex1[nn] = tms * (-f2 * C2) * tslant[nn]; ex2[nn] = tms * (f2 * S2) * tslant[nn]; }
    }
  }
  T.sync();
  if (T.rank == 0) {
    double cum[3] = {0, 0, 0}, db[3] = {0, 0, 0};
    const double M4 = 4.0 * ctheta_boa;
    if (in.lambertian) db[0] = M4 * albedo * atten[L];
    else for (int o = 0; o < S; o++) db[o] = M4 * in.exactdb[(size_t)p * 16 + 4 * o] * atten[L];
    for (int nn = L; nn >= 1; nn--) {
      const double lt = ltr[nn];
      db[0] *= lt; cum[0] = lt * cum[0] + ex0[nn];
      if (S > 1) { db[1] *= lt; db[2] *= lt; cum[1] = lt * cum[1] + ex1[nn]; cum[2] = lt * cum[2] + ex2[nn]; }
    }
    const double fl = FLUX / PI4;
    for (int o = 0; o < 3; o++) { sc[6 + o] = fl * cum[o]; sc[9 + o] = fl * db[o]; }
  }
  T.sync();
}

}  // namespace vlk
