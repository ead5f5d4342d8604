/* TEST INFRASTRUCTURE ONLY -- First-Order (exact single scatter + direct bounce) part of the CPU oracle.
 * Observation geometry, one geometry, upwelling, TOA output only (what AeroApps uses).
 * Restates (relative to /root/reference/src/Components/rtms/RTSI/VLIDORT2p8p3/vlidort_focode/):
 *   FO_WPgeometry_Routines.f90:126-311    ObsGeom_PlanPar_WP
 *   FO_WPgeometry_Routines.f90:776-991    Obsgeom_RegularPS_WP
 *   FO_WPgeometry_Routines.f90:1511-1971  LosOut_EnhancedPS_Initial_WP / _Quadrature_WP
 *   FO_WPgeometry_Routines.f90:2409-2859  SolarIn_EnhancedPS_ObsGeom_SunPaths_WP
 *   FO_geometry_Generic.f90:105-327       FindSun, FindSunPaths_D, FindSunPaths_T
 *   FO_VectorSS_Spherfuncs.f90:93-330     generalized spherical functions + rotation angles
 *   FO_VectorSS_RTCalcs_I.f90:116-785     SSV_Integral_I_UP
 */
#pragma once
#include "vl_numerics.h"

namespace vl {

struct FOGeom {
  int L = 0, nfine = 0;
  bool doNadir = false;
  double Mu0 = 0, Mu1 = 0, cosscat = 0, Raycon = 0;
  std::vector<double> radii, LosW, alpha, sina, cosa;           /* 0..L */
  std::vector<double> sunpaths;                                  /* (L+1) x L : [n][k-1] */
  std::vector<int> ntraverse;                                    /* 0..L */
  std::vector<double> xfine, wfine, radiifine, alphafine;        /* (L+1) x nfine : [n][j] */
  std::vector<double> sunpathsfine;                              /* (L+1) x nfine x L */
  std::vector<int> ntraversefine;                                /* (L+1) x nfine */
  void init(int L_, int nf) {
    L = L_; nfine = nf;
    radii.assign(L + 1, 0); LosW.assign(L + 1, 0); alpha.assign(L + 1, 0); sina.assign(L + 1, 0); cosa.assign(L + 1, 0);
    sunpaths.assign((size_t)(L + 1) * L, 0); ntraverse.assign(L + 1, 0);
    int nf1 = std::max(nf, 1);
    xfine.assign((size_t)(L + 1) * nf1, 0); wfine = xfine; radiifine = xfine; alphafine = xfine;
    sunpathsfine.assign((size_t)(L + 1) * nf1 * L, 0); ntraversefine.assign((size_t)(L + 1) * nf1, 0);
  }
  double *sp(int n) { return &sunpaths[(size_t)n * L]; }                       /* sunpaths(n,1:L) */
  double *spf(int n, int j) { return &sunpathsfine[((size_t)n * std::max(nfine, 1) + j) * L]; }
};

/* FindSunPaths_D: sunpaths[k-1], k=1..N */
inline void find_sunpaths_D(bool zeroSunBOA, double Radstart, const double *radii, double thstart, double sthstart, int N,
                            double *sunpaths) {
  int N1 = N - 1;
  if (zeroSunBOA) {
    sunpaths[N - 1] = radii[N1] - Radstart;
    for (int k = N1; k >= 1; k--) sunpaths[k - 1] = radii[k - 1] - radii[k];
    return;
  }
  double sth0 = sthstart, th0 = thstart;
  double sth1 = sth0 * Radstart / radii[N1];
  double th1 = std::asin(sth1), ks1 = th0 - th1;
  sunpaths[N - 1] = std::sin(ks1) * Radstart / sth1;
  sth0 = sth1; th0 = th1;
  for (int k = N1; k >= 1; k--) {
    sth1 = sth0 * radii[k] / radii[k - 1];
    th1 = std::asin(sth1); ks1 = th0 - th1;
    sunpaths[k - 1] = std::sin(ks1) * radii[k] / sth1;
    sth0 = sth1; th0 = th1;
  }
}

/* FindSunPaths_T (tangent-height paths).  NT returned. */
inline void find_sunpaths_T(int maxlayers, double Pie, double Radstart, const double *radii, double thstart, double sthstart,
                            int N, double *sunpaths, int *NT) {
  int N1 = N - 1, nt;
  double tanr = sthstart * Radstart;
  int k = N1; bool trawl = true;
  while (k >= N1 && trawl) { trawl = (k <= maxlayers) && (radii[k] > tanr); k = k + 1; }
  nt = k - 1;
  *NT = nt;
  double sth0, th0, sth1, th1, ks1;
  if (nt > N) {
    th0 = Pie - thstart; sth0 = sthstart;
    sth1 = sth0 * Radstart / radii[N];
    th1 = std::asin(sth1); ks1 = th0 - th1;
    sunpaths[N - 1] = 2.0 * std::sin(ks1) * Radstart / sth1;
    sth0 = sth1; th0 = th1;
    for (k = N + 1; k <= nt - 1; k++) {
      sth1 = sth0 * radii[k - 1] / radii[k];
      th1 = std::asin(sth1); ks1 = th0 - th1;
      sunpaths[k - 1] = 2.0 * std::sin(ks1) * radii[k] / sth0;
      sth0 = sth1; th0 = th1;
    }
    sth1 = 1.0; ks1 = 0.5 * Pie - th0;
    sunpaths[nt - 1] = 2.0 * std::sin(ks1) * radii[nt - 1];
  } else if (nt == N) {
    sunpaths[N - 1] = -2.0 * Radstart * std::cos(thstart);
  }
  th0 = Pie - thstart; sth0 = sthstart;
  sth1 = sth0 * Radstart / radii[N1];
  th1 = std::asin(sth1); ks1 = th0 - th1;
  sunpaths[N - 1] = sunpaths[N - 1] + std::sin(ks1) * Radstart / sth1;
  sth0 = sth1; th0 = th1;
  for (k = N1; k >= 1; k--) {
    sth1 = sth0 * radii[k] / radii[k - 1];
    th1 = std::asin(sth1); ks1 = th0 - th1;
    sunpaths[k - 1] = std::sin(ks1) * radii[k] / sth1;
    sth0 = sth1; th0 = th1;
  }
}

inline void find_sun(bool doNadir, bool overhead, double Radius, const double *SolarDirection, double CumAngle,
                     double theta_boa_R, double *theta, double *stheta, double *ctheta, bool *DirSun) {
  if (doNadir) { *DirSun = true; *theta = theta_boa_R; *ctheta = std::cos(theta_boa_R); *stheta = std::sin(theta_boa_R); return; }
  if (overhead) { *DirSun = true; *ctheta = std::cos(CumAngle); *stheta = std::sin(CumAngle); *theta = CumAngle; return; }
  double px0 = -Radius * std::sin(CumAngle), px2 = Radius * std::cos(CumAngle);
  double b = px0 * SolarDirection[0] + 0.0 * SolarDirection[1] + px2 * SolarDirection[2];
  *ctheta = -b / Radius;
  *DirSun = (*ctheta >= 0.0);
  *stheta = std::sqrt(1.0 - (*ctheta) * (*ctheta));
  *theta = std::acos(*ctheta);
}

/* mode: 0 plane-parallel, 1 regular PS, 2 enhanced PS.  heights in km (0..L), angles in degrees (BOA). */
inline void fo_geometry(int mode, int L, int nfine, double eradius, const double *heights, double sza, double vza,
                        double raa, FOGeom &G) {
  const double dtr = DEG_TO_RAD, vsign = 1.0;
  G.init(L, mode == 2 ? nfine : 0);
  if (raa < 0.0) raa = -raa;                                   /* FO_SSWPgeometry_master.f90 angle fix-ups */
  if (raa > 360.0) raa = 360.0 - raa - 360.0;
  std::vector<double> diffhts(L + 1, 0.0);
  for (int n = 1; n <= L; n++) diffhts[n] = heights[n - 1] - heights[n];
  double alpha_boa_R = vza * dtr, calpha_boa, salpha_boa;
  if (vza == 90.0) { calpha_boa = 0.0; salpha_boa = 1.0; } else { calpha_boa = std::cos(alpha_boa_R); salpha_boa = std::sin(alpha_boa_R); }
  double theta_boa_R = sza * dtr, stheta_boa, ctheta_boa;
  if (mode != 0 && sza == 90.0) { ctheta_boa = 0.0; stheta_boa = 1.0; } else { stheta_boa = std::sin(theta_boa_R); ctheta_boa = std::cos(theta_boa_R); }
  double cphi_boa = std::cos(raa * dtr), sphi_boa = std::sin(raa * dtr);
  bool overhead = (sza == 0.0);
  if (mode == 0 || mode == 1) {
    for (int n = 0; n <= L; n++) G.radii[n] = eradius + heights[n];
    G.Mu0 = ctheta_boa; G.Mu1 = calpha_boa;
    if (mode == 1) G.Raycon = salpha_boa * G.radii[L];
    if (overhead) {
      double term2 = calpha_boa;
      G.cosscat = -vsign * term2; if (term2 == 0.0) G.cosscat = term2;
    } else {
      double term1 = salpha_boa * stheta_boa * cphi_boa, term2 = calpha_boa * ctheta_boa;
      G.cosscat = -vsign * term2 + term1;
    }
    for (int n = 0; n <= L; n++) G.alpha[n] = alpha_boa_R;
    double utheta_boa = 1.0 / ctheta_boa;
    for (int n = 1; n <= L; n++) {
      G.ntraverse[n] = n;
      G.LosW[n] = diffhts[n] / calpha_boa;
      if (mode == 0) for (int k = 1; k <= n; k++) G.sp(n)[k - 1] = diffhts[k] * utheta_boa;
      else find_sunpaths_D(overhead, G.radii[n], G.radii.data(), theta_boa_R, stheta_boa, n, G.sp(n));
    }
    return;
  }
  /* ---- enhanced PS: LosOut_EnhancedPS_Initial_WP */
  for (int n = 0; n <= L; n++) G.radii[n] = eradius + heights[n];
  G.doNadir = (vza == 0.0);
  if (G.doNadir) {
    for (int n = L; n >= 1; n--) G.LosW[n] = G.radii[n - 1] - G.radii[n];
  } else {
    double calpha1 = calpha_boa;
    G.cosa[L] = calpha1; G.sina[L] = salpha_boa; G.alpha[L] = alpha_boa_R;
    G.Raycon = salpha_boa * G.radii[L];
    for (int n = L - 1; n >= 0; n--) {
      int n1 = n + 1;
      G.sina[n] = G.Raycon / G.radii[n]; G.alpha[n] = std::asin(G.sina[n]);
      double calpha = std::cos(G.alpha[n]); G.cosa[n] = calpha;
      G.LosW[n1] = G.radii[n] * calpha - G.radii[n1] * calpha1;
      calpha1 = calpha;
    }
  }
  /* LosOut_EnhancedPS_Quadrature_WP (upwelling) */
  std::vector<double> tfine(nfine), afine(nfine);
  for (int n = L; n >= 1; n--) {
    if (G.doNadir) {
      double difh = G.radii[n - 1] - G.radii[n];
      getquad2(0.0, difh, nfine, tfine.data(), afine.data());
      for (int j = 0; j < nfine; j++) {
        G.radiifine[(size_t)n * nfine + j] = G.radii[n] + tfine[j];
        G.xfine[(size_t)n * nfine + j] = tfine[j]; G.wfine[(size_t)n * nfine + j] = afine[j];
      }
    } else {
      double rcn = G.radii[n] * G.cosa[n];
      getquad2(0.0, G.LosW[n], nfine, tfine.data(), afine.data());
      for (int j = 0; j < nfine; j++) {
        double tanfine = G.Raycon / (rcn + tfine[j]);
        double af = std::atan(tanfine), sf = std::sin(af);
        G.alphafine[(size_t)n * nfine + j] = af;
        G.radiifine[(size_t)n * nfine + j] = G.Raycon / sf;
        G.xfine[(size_t)n * nfine + j] = tfine[j]; G.wfine[(size_t)n * nfine + j] = afine[j];
      }
    }
  }
  /* SolarIn_EnhancedPS_ObsGeom_SunPaths_WP */
  for (int n = 1; n <= L; n++) { G.ntraverse[n] = n; for (int j = 0; j < nfine; j++) G.ntraversefine[(size_t)n * nfine + j] = n; }
  if (vza == 90.0) { calpha_boa = 0.0; salpha_boa = 1.0; } else { salpha_boa = std::sin(G.alpha[L]); calpha_boa = std::cos(G.alpha[L]); }
  G.Mu1 = calpha_boa; G.Mu0 = ctheta_boa;
  double SolarDirection[3] = {0, 0, 0};
  if (!overhead) { SolarDirection[0] = -stheta_boa * cphi_boa * vsign; SolarDirection[1] = -stheta_boa * sphi_boa; SolarDirection[2] = -ctheta_boa; }
  if (overhead) { double term2 = calpha_boa; G.cosscat = -vsign * term2; if (term2 == 0.0) G.cosscat = term2; }
  else { double term1 = salpha_boa * stheta_boa * cphi_boa, term2 = calpha_boa * ctheta_boa; G.cosscat = -vsign * term2 + term1; }
  std::vector<double> thetaf(nfine), sthetaf(nfine), cthetaf(nfine);
  std::vector<char> dsf(nfine);
  for (int n = L; n >= 1; n--) {
    bool zeroSunBOA = overhead && (n == L || G.doNadir);
    double Radstart = G.radii[n], theta_n, stheta, ctheta; bool DirectSun;
    if (n == L) { theta_n = theta_boa_R; stheta = stheta_boa; ctheta = ctheta_boa; DirectSun = true; }
    else {
      double CumAngle = G.alpha[L] - G.alpha[n];
      find_sun(G.doNadir, overhead, G.radii[n], SolarDirection, CumAngle, theta_boa_R, &theta_n, &stheta, &ctheta, &DirectSun);
    }
    for (int j = 0; j < nfine; j++) {
      double CumAngle = G.alpha[L] - G.alphafine[(size_t)n * nfine + j];
      bool ds;
      find_sun(G.doNadir, overhead, G.radiifine[(size_t)n * nfine + j], SolarDirection, CumAngle, theta_boa_R, &thetaf[j],
               &sthetaf[j], &cthetaf[j], &ds);
      dsf[j] = ds;
    }
    if (DirectSun) find_sunpaths_D(zeroSunBOA, Radstart, G.radii.data(), theta_n, stheta, n, G.sp(n));
    else find_sunpaths_T(L, PIE, Radstart, G.radii.data(), theta_n, stheta, n, G.sp(n), &G.ntraverse[n]);
    for (int j = 0; j < nfine; j++) {
      if (dsf[j]) find_sunpaths_D(zeroSunBOA, G.radiifine[(size_t)n * nfine + j], G.radii.data(), thetaf[j], sthetaf[j], n, G.spf(n, j));
      else find_sunpaths_T(L, PIE, G.radiifine[(size_t)n * nfine + j], G.radii.data(), thetaf[j], sthetaf[j], n, G.spf(n, j),
                           &G.ntraversefine[(size_t)n * nfine + j]);
    }
  }
  /* level 0 (TOA): only a tangent path contributes a sunpath */
  {
    double CumAngle = G.alpha[L] - G.alpha[0], th, st, ct; bool ds;
    find_sun(G.doNadir, overhead, G.radii[0], SolarDirection, CumAngle, theta_boa_R, &th, &st, &ct, &ds);
    if (!ds) find_sunpaths_T(L, PIE, G.radii[0], G.radii.data(), th, st, 1, G.sp(0), &G.ntraverse[0]);
  }
}

/* FO_VectorSS_spherfuncs, observation geometry, sunlight: genspher[l][0..1], rotations[4] */
inline void fo_spherfuncs(int nmoments, int nstokes, double sza, double vza, double raa, double cosscat,
                          std::vector<double> &gs1, std::vector<double> &gs2, double rot[4]) {
  const double vsign = 1.0, SMALL = 1.0e-12;
  gs1.assign(nmoments + 1, 0.0); gs2.assign(nmoments + 1, 0.0);
  std::vector<double> G1(nmoments + 2, 0), G2(nmoments + 2, 0), G3(nmoments + 2, 0), G4(nmoments + 2, 0);
  double SQL41 = 0.0;
  for (int L = 2; L <= nmoments; L++) {
    double DL = (double)L, DL1 = (double)(L - 1), DL21 = (double)(2 * L - 1);
    G1[L] = DL21 / DL; G2[L] = DL1 / DL;
    if (nstokes > 1) {
      if (L == 2) SQL41 = 0.0;
      else { double SQL4 = SQL41, DL41 = DL * DL - 4.0; SQL41 = std::sqrt(DL41); G3[L] = DL21 / SQL41; G4[L] = SQL4 / SQL41; }
    }
  }
  const double QROOT6 = -0.25 * std::sqrt(6.0);
  double mus = cosscat;
  double THETA_R = DEG_TO_RAD * sza, CTHETA = std::cos(THETA_R), STHETA = std::sin(THETA_R);
  double ALPHA_R = DEG_TO_RAD * vza, CALPHA = std::cos(ALPHA_R), SALPHA = std::sin(ALPHA_R);
  double CPHI = std::cos(DEG_TO_RAD * raa);
  double HELP_SINSCAT = 1.0 - mus * mus, SINSCAT = HELP_SINSCAT <= 0.0 ? SMALL : std::sqrt(HELP_SINSCAT);
  double CSIG1, CSIG2;
  if (std::fabs(SINSCAT) <= SMALL) { CSIG1 = 0.0; CSIG2 = 0.0; }
  else {
    CSIG1 = (STHETA == 0.0) ? -CPHI : (vsign * CALPHA + CTHETA * mus) / SINSCAT / STHETA;
    CSIG2 = (SALPHA == 0.0) ? -CPHI : (-CTHETA - vsign * CALPHA * mus) / SINSCAT / SALPHA;
  }
  CSIG1 = std::min(1.0, std::max(-1.0, CSIG1)); CSIG2 = std::min(1.0, std::max(-1.0, CSIG2));
  double CSIG1_2 = 2.0 * CSIG1, CSIG2_2 = 2.0 * CSIG2;
  double SSIG1 = std::fabs(CSIG1 - 1.0) < SMALL ? 0.0 : std::sqrt(1.0 - CSIG1 * CSIG1);
  double SSIG2 = std::fabs(CSIG2 - 1.0) < SMALL ? 0.0 : std::sqrt(1.0 - CSIG2 * CSIG2);
  double C1 = CSIG1_2 * CSIG1 - 1.0, C2 = CSIG2_2 * CSIG2 - 1.0, S1, S2;
  if (raa <= 180.0) { S1 = CSIG1_2 * SSIG1; S2 = CSIG2_2 * SSIG2; } else { S1 = -CSIG1_2 * SSIG1; S2 = -CSIG2_2 * SSIG2; }
  rot[0] = C1; rot[1] = S1; rot[2] = C2; rot[3] = S2;
  if (nstokes == 1) {
    gs1[0] = 1.0; if (nmoments >= 1) gs1[1] = mus;
    for (int L = 2; L <= nmoments; L++) gs1[L] = G1[L] * gs1[L - 1] * mus - G2[L] * gs1[L - 2];
    return;
  }
  double P00[2] = {0, 0}, P02[2] = {0, 0};
  int LNEW = 0, LOLD = 1;
  for (int L = 0; L <= nmoments; L++) {
    if (L == 0) { P00[LOLD] = 1.0; P00[LNEW] = 0.0; P02[LOLD] = 0.0; P02[LNEW] = 0.0; }
    else if (L == 1) P00[LOLD] = mus;
    else P00[LOLD] = G1[L] * mus * P00[LNEW] - G2[L] * P00[LOLD];
    gs1[L] = P00[LOLD];
    if (L == 2) { double MUM1 = 1.0 - mus, MUP1 = 1.0 + mus; P02[LOLD] = QROOT6 * MUM1 * MUP1; P02[LNEW] = 0.0; gs2[L] = P02[LOLD]; }
    else if (L > 2) { P02[LOLD] = G3[L] * mus * P02[LNEW] - G4[L] * P02[LOLD]; gs2[L] = P02[LOLD]; }
    std::swap(LNEW, LOLD);
  }
}

/* SSV_Integral_I_UP for TOA output (user level 0), sunlight.  greek(n,l,k): l-major input array.
 * out: stokes_ss[ns], stokes_db[ns] (already multiplied by flux = FLUX_FACTOR/4pi). */
struct FOOptical {
  int L, nstokes, nmoms_in;
  bool do_deltam, lambertian;
  double flux;                          /* FLUX_FACTOR / PI4 */
  const double *extinction;             /* L (index n-1): scaled deltau / layer thickness */
  const double *deltaus, *omega, *truncfac;
  const double *greek;                  /* [(l*L + (n-1))*16 + k] */
  double albedo; const double *reflec16;/* exact BRDF (16, Mueller index 4*(o1-1)+o2-1) when !lambertian */
};
inline void fo_ssv_integral_up(int mode, FOGeom &G, const FOOptical &O, const std::vector<double> &gs1,
                               const std::vector<double> &gs2, const double rot[4], double *stokes_ss, double *stokes_db) {
  const double cutoff = 88.0;
  const int L = O.L, ns = O.nstokes, nf = G.nfine;
  std::vector<double> tms(L + 1), exact(3 * (L + 1), 0.0), atten(L + 1, 0.0), suntau(L + 1, 0.0), lostrans(L + 1, 0.0), mult(L + 1, 0.0);
  for (int n = 1; n <= L; n++) tms[n] = O.do_deltam ? O.omega[n - 1] / (1.0 - O.truncfac[n - 1] * O.omega[n - 1]) : O.omega[n - 1];
  for (int n = 1; n <= L; n++) {
    double f1 = 0.0, f2 = 0.0;
    for (int l = 0; l <= O.nmoms_in; l++) {
      const double *g = O.greek + ((size_t)l * L + (n - 1)) * 16;
      f1 += gs1[l] * g[0];
      if (ns > 1) f2 += gs2[l] * g[1];
    }
    exact[3 * n + 0] = f1 * tms[n];
    if (ns > 1) { exact[3 * n + 1] = tms[n] * (-f2 * rot[2]); exact[3 * n + 2] = tms[n] * (f2 * rot[3]); }
  }
  for (int n = 0; n <= L; n++) {
    int nt = G.ntraverse[n]; double sumd = 0.0;
    for (int k = 1; k <= nt; k++) sumd += O.extinction[k - 1] * G.sp(n)[k - 1];
    suntau[n] = sumd; if (sumd < cutoff) atten[n] = std::exp(-sumd);
  }
  if (mode != 2) {
    for (int n = L; n >= 1; n--) {
      if (G.Mu1 > 0.0) {
        double lostau = O.deltaus[n - 1] / G.Mu1;
        if (lostau < cutoff) lostrans[n] = std::exp(-lostau);
        double factor1 = atten[n - 1] - atten[n] * lostrans[n];
        double factor2 = 1.0 + (suntau[n] - suntau[n - 1]) / lostau;
        mult[n] = factor1 / factor2;
      }
    }
  } else {
    for (int n = L; n >= 1; n--) {
      double kn = O.extinction[n - 1], lostau = kn * G.LosW[n];
      if (lostau < cutoff) lostrans[n] = std::exp(-lostau);
      double sum = 0.0;
      for (int j = 0; j < nf; j++) {
        int nt = G.ntraversefine[(size_t)n * nf + j]; double sumd = 0.0;
        for (int k = 1; k <= nt; k++) sumd += O.extinction[k - 1] * G.spf(n, j)[k - 1];
        double af = sumd < cutoff ? std::exp(-sumd) : 0.0;
        double dj = G.LosW[n] - G.xfine[(size_t)n * nf + j], tran = std::exp(-kn * dj);
        sum += af * tran * G.wfine[(size_t)n * nf + j];
      }
      mult[n] = sum * kn;
    }
  }
  double cum[3] = {0, 0, 0}, db[3] = {0, 0, 0};
  const double M4 = 4.0 * G.Mu0;
  if (O.lambertian) db[0] = M4 * O.albedo * atten[L] * 1.0;
  else for (int o1 = 0; o1 < ns; o1++) db[o1] = M4 * O.reflec16[4 * o1 + 0] * atten[L];   /* fvec = (1,0,0,0) */
  for (int n = L; n >= 1; n--)
    for (int o1 = 0; o1 < ns; o1++) {
      db[o1] = lostrans[n] * db[o1];
      cum[o1] = lostrans[n] * cum[o1] + exact[3 * n + o1] * mult[n];
    }
  for (int o1 = 0; o1 < ns; o1++) { stokes_ss[o1] = O.flux * cum[o1]; stokes_db[o1] = O.flux * db[o1]; }
}

}  // namespace vl
