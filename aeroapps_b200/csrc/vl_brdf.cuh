// Stage B, kernel K0: per-pixel BRDF supplement for the Cox-Munk, GISS Cox-Munk and MODIS RTLS wrappers.
//   surface set-up   V283/VLIDORT_SurfaceMod.F90:586-814 (GISS Cox-Munk) 816-1042 (Cox-Munk) 1376-1596 (LANDMODIS)
//   VBRDF_MAINMASTER vsup/vbrdf/vbrdf_sup_masters.f90:1559-3295   (kernel loop, WSA check :2953-2998)
//   VBRDF_MAKER / VBRDF_FOURIER        vsup/vbrdf/vbrdf_sup_routines.f90:153-690, 2073-2491
//   kernels                            vsup/vbrdf/vbrdf_sup_kernels.f90:201-335, 478-585, 1018-1495
// The reference evaluates the kernels on the azimuth quadrature into BRDFUNC arrays in one routine and projects
// them onto cos/sin(m phi) in a second one, per kernel.  Here one Team (a CTA) owns a pixel: the kernel values of a
// batch of (incident, reflected) stream pairs are staged in shared memory and projected straight away onto all
// Fourier components with host-built cos/sin tables, so the BRDFUNC arrays never exist in HBM.  Only the elements
// the solver reads are stored: nq = 1 (scalar kernels: element (1,1)) or 16 (GISS Cox-Munk, 3x3 block).
#pragma once
#include "vl_setup.cuh"
#include <vector>
#include <cstring>

namespace vlk {

constexpr int BRDF_NSCAL = 24;       // SCALING_NSTREAMS: white-sky-albedo quadrature (vbrdf_sup_masters.f90:2238)
constexpr int BRDF_MAXNB = 100;      // azimuth quadrature streams (MAXSTREAMS_BRDF)

// OCI variants (V283/VLIDORT_SurfaceMod.F90:82-538): same kernels, slope variance sigma2 = 0.00534 W instead of 0.003 + 0.00512 W
enum { SURF_LAMBERT = 0, SURF_CX = 2, SURF_GISSCX = 3, SURF_MODIS = 4, SURF_OCICX = 5, SURF_OCIGISSCX = 6, SURF_BPDF = 7,
       SURF_MODIS_BPDF = 8, SURF_LAMBERT_BPDF = 9 };
HDFN bool surf_is_giss(int kind) { return kind == SURF_GISSCX || kind == SURF_OCIGISSCX; }
HDFN bool surf_is_oci(int kind) { return kind == SURF_OCICX || kind == SURF_OCIGISSCX; }
HDFN bool surf_has_bpdf(int kind) { return kind == SURF_BPDF || kind == SURF_MODIS_BPDF || kind == SURF_LAMBERT_BPDF; }
HDFN bool surf_has_rtls(int kind) { return kind == SURF_MODIS || kind == SURF_MODIS_BPDF; }
HDFN bool surf_is_land(int kind) { return kind == SURF_MODIS || surf_has_bpdf(kind); }      // NSTREAMS_BRDF = 50, BR < 0 -> -500, no wind inputs
HDFN bool surf_is_vec(int kind) { return surf_is_giss(kind) || surf_has_bpdf(kind); }        // kernel with a full Mueller block (index >= 10)
HDFN bool brdf_vec_which(int which) { return which == 10 || which == 14; }

struct VLBrdf {
  int kind, N, NM, NB, nq, half_range, nch, nobs, bns;   // bns: BRDF nstokes (1 or 3)
  int G;                                    // view geometries per observation (0 or 1: single geometry)
  const double *sza;                        // (nobs)
  const double *vza, *raa;                  // (nobs,G)
  const double *u10m, *v10m, *mr;           // (nobs) (nobs) (nch)      Cox-Munk family
  const double *kernel_wt, *param;          // (3,nch,nobs) (2,nch,nobs) MODIS RTLS; param (3,nch,nobs) = BPDFparam for SURF_BPDF
  const double *bpdf, *albedo;              // SURF_MODIS_BPDF / SURF_LAMBERT_BPDF: BPDFparam (3,nch,nobs); SURF_LAMBERT_BPDF: albedo (nobs,nch)
  const double *xb, *ab, *cxb, *sxb;        // [NB] azimuth quadrature (angles, weights, cos, sin)
  const double *caz, *saz;                  // [NM+1][NB]  weight*cos(m x), weight*sin(m x)
  const double *sx, *ss, *ssw;              // [24] scaling quadrature: cosines, sines, cosine*weight
  double qx[32], qs[32];                    // stream cosines / sines
  // outputs
  double *exactdb;                          // [npix*G][16]   per pixel-view
  double *brdf_f, *brdf_f_0;                // chunk-local, per pixel: [pc][2N][nq][N][N], [pc][2N][nq][N]
  double *user_brdf_f;                      // chunk-local, per pixel-view: [pc*G + g][2N][nq][N]
  int *wsa_fail;                            // [npix]
};

// pixel-independent tables (host): azimuth quadrature of the supplement (BRDF_QUADRATURE_GAUSSIAN, half range for the
// scalar kernels, mirrored full range for kernel indices >= 10), weight*cos/sin(m x) for VBRDF_FOURIER, the 24-point
// scaling quadrature of the white-sky albedo and the stream cosines GETQUAD2(0,1,N).  B's table pointers are set to
// offsets inside h (rebased by the caller after an upload).
inline void brdf_host_tables(VLBrdf &B, int kind, int S, int scalar_glue, int N, int NM, int nch, int nobs, std::vector<double> &h) {
  memset(&B, 0, sizeof B);
  B.kind = kind; B.N = N; B.NM = NM; B.nch = nch; B.nobs = nobs;
  B.half_range = !surf_is_vec(kind);
  const int nsb = surf_is_land(kind) ? 50 : 100;                        // NSTREAMS_BRDF of the surface glue
  B.NB = B.half_range ? nsb / 2 : nsb;
  B.bns = (surf_is_vec(kind) && !scalar_glue && S == 3) ? 3 : 1;
  B.nq = (B.bns == 3) ? 16 : 1;
  const int NB = B.NB;
  h.assign((size_t)4 * NB + (size_t)2 * (NM + 1) * NB + 3 * BRDF_NSCAL, 0.0);
  double *xb = h.data(), *ab = xb + NB, *cxb = ab + NB, *sxb = cxb + NB, *caz = sxb + NB, *saz = caz + (size_t)(NM + 1) * NB;
  double *sx = saz + (size_t)(NM + 1) * NB, *ss = sx + BRDF_NSCAL, *ssw = ss + BRDF_NSCAL;
  if (B.half_range) gauss01(NB, xb, ab);
  else { gauss01(NB / 2, xb, ab); for (int b = 0; b < NB / 2; b++) { xb[b + NB / 2] = -xb[b]; ab[b + NB / 2] = ab[b]; } }
  for (int b = 0; b < NB; b++) { xb[b] = PIE * xb[b]; cxb[b] = cos(xb[b]); sxb[b] = sin(xb[b]); }
  for (int m = 0; m <= NM; m++)
    for (int b = 0; b < NB; b++) {
      caz[(size_t)m * NB + b] = (m == 0) ? ab[b] : ab[b] * cos(m * xb[b]);
      saz[(size_t)m * NB + b] = (m == 0) ? 0.0 : ab[b] * sin(m * xb[b]);
    }
  double sw[BRDF_NSCAL], qw[32];
  gauss01(BRDF_NSCAL, sx, sw);
  for (int t = 0; t < BRDF_NSCAL; t++) { ss[t] = sqrt(1.0 - sx[t] * sx[t]); ssw[t] = sx[t] * sw[t]; }
  gauss01(N, B.qx, qw);                                                   // also for N = 1 (0.5), unlike the solver's own stream
  for (int t = 0; t < N; t++) B.qs[t] = sqrt(1.0 - B.qx[t] * B.qx[t]);
  B.xb = xb; B.ab = ab; B.cxb = cxb; B.sxb = sxb; B.caz = caz; B.saz = saz; B.sx = sx; B.ss = ss; B.ssw = ssw;
}
inline void brdf_rebase_tables(VLBrdf &B, const double *from, const double *to) {
  const double **pp[] = {&B.xb, &B.ab, &B.cxb, &B.sxb, &B.caz, &B.saz, &B.sx, &B.ss, &B.ssw};
  for (auto q : pp) *q = to + (*q - from);
}

HDFN double derfc_e(double x) {                                                  // vbrdf_sup_aux.f90:266-286
  const double z = fabs(x), t = 1.0 / (1.0 + 0.5 * z);
  double r = t * exp(-z * z - 1.26551223 + t * (1.00002368 + t * (.37409196 + t * (.09678418 + t * (-.18628806 + t * (.27886807 + t *
             (-1.13520398 + t * (1.48851587 + t * (-.82215223 + t * .17087277)))))))));
  if (x < 0.0) r = 2.0 - r;
  return r;
}

// scalar kernels: value of element (1,1).  (XI,SXI) incident, (XJ,SXJ) reflected, as the reference's dummies.
HDFN double brdf_scalar_kernel(int which, const double *pars, double XI, double SXI, double XJ, double SXJ, double cphi, double sphi) {
  const double PIO2 = 0.5 * PIE, PIO4 = 0.25 * PIE;
  if (which == 1) return 1.0;
  if (which == 3) {                                                              // Ross-thick
    const double ds1 = XI * XJ, ds2 = SXI * SXJ, ds3 = XI + XJ;
    double cksi = ds1 - ds2 * cphi; if (cksi > 1.0) cksi = 1.0;
    const double sksi = sqrt(1.0 - cksi * cksi), ksi = acos(cksi);
    return ((PIO2 - ksi) * cksi + sksi) / ds3 - PIO4;
  }
  if (which == 4) {                                                              // Li-sparse
    const double ckphi = -cphi;
    const double t_inc = pars[1] * (SXI / XI); double ang = atan(t_inc); const double x_inc = cos(ang), sx_inc = sin(ang);
    const double t_ref = pars[1] * (SXJ / XJ); ang = atan(t_ref); const double x_ref = cos(ang), sx_ref = sin(ang);
    const double cksi = x_inc * x_ref + sx_inc * sx_ref * ckphi;
    const double P = (1.0 + cksi) / x_ref / x_inc, A = 1.0 / x_inc, B = 1.0 / x_ref, R = A + B;
    const double dt1 = t_ref * t_ref + t_inc * t_inc, dt2 = t_inc * t_ref, dt2sq = dt2 * dt2;
    const double delta = sqrt(dt1 - 2.0 * dt2 * ckphi), dsq = delta * delta;
    const double H = sqrt(dsq + sphi * sphi * dt2sq), cost = pars[0] * H / R;
    double Q;
    if (cost > 1.0) Q = 1.0;
    else { const double T = acos(cost), sint = sqrt(1.0 - cost * cost); Q = 1.0 - ((T - sint * cost) / PIE); }
    return 0.5 * P - Q * R;
  }
  if (which == 9) {                                                              // Cox-Munk
    double Z = XI * XJ - SXI * SXJ * cphi; if (Z > 1.0) Z = 1.0;
    const double Z1 = acos(Z), Z2 = cos(Z1 * 0.5), Z2m1 = Z2 * Z2 - 1.0;
    const double H1 = pars[1] * Z2, H2 = sqrt(pars[1] + Z2m1);
    const double RP = (H1 - H2) / (H1 + H2), RL = (Z2 - H2) / (Z2 + H2), XMP = 0.5 * (RP * RP + RL * RL);
    double A = 2.0 * Z2, B = (XI + XJ) / A; if (B > 1.0) B = 1.0;
    A = PIO2 - asin(B);
    const double TA = tan(A), arg = TA * TA / pars[0];
    double v = 0.0;
    if (arg < 88.0) {
      const double prob = exp(-arg), fac1 = prob / pars[0], fac2 = 0.25 / XI / pow(B, 4.0);
      v = XMP * fac1 * fac2 / XJ;
    }
    if (pars[2] == 0.0) return v;
    const double S1 = sqrt(pars[0] / PIE), S3 = 1.0 / sqrt(pars[0]), S2 = S3 * S3;
    double dcot = XI / sqrt(1.0 - XI * XI), T1 = exp(-dcot * dcot * S2), T2 = derfc_e(dcot * S3);
    const double shadowi = 0.5 * (S1 * T1 / dcot - T2);
    dcot = XJ / sqrt(1.0 - XJ * XJ); T1 = exp(-dcot * dcot * S2); T2 = derfc_e(dcot * S3);
    const double shadowr = 0.5 * (S1 * T1 / dcot - T2);
    return v * (1.0 / (1.0 + shadowi + shadowr));
  }
  return 0.0;
}

// GISS Cox-Munk: 3x3 block; K[3*o1+o2] (o1 = row).  bns = 1 evaluates the same code, the caller keeps K[0] only.
HDFN void brdf_giss_kernel(const double *pars, double XI, double SXI, double XJ, double SXJ, double cphi, double sphi, double *K) {
  double xi = XI, xj = XJ;
  const double sigma2 = pars[0];
  if (fabs(xi - 1.0) < 1e-9) xi = 0.999999999999;
  if (fabs(xj - 1.0) < 1e-9) xj = 0.999999999999;
  const double VI1 = SXI, VI3 = -xi;                                             // ckphi_inc = 1, skphi_inc = 0 -> VI2 = 0
  const double VR1 = SXJ * cphi, VR2 = SXJ * sphi, VR3 = xj;
  const double U1 = VI1 - VR1, U2 = -VR2, U3 = VI3 - VR3;
  const double FACT1 = U1 * U1 + U2 * U2 + U3 * U3, FACTOR = sqrt(1.0 / FACT1);
  const double CN2 = pars[1];
  const double XI1 = FACTOR * (U1 * VI1 + U3 * VI3);
  const double CXI2 = sqrt(1.0 - (1.0 - XI1 * XI1) / (CN2 * CN2));
  double C1 = XI1, C2 = CN2 * CXI2; const double CRPER = (C1 - C2) / (C1 + C2);
  C1 = CN2 * XI1; C2 = CXI2; const double CRPAR = (C1 - C2) / (C1 + C2);
  const double TI1 = -xi, TI3 = -SXI;                                            // TI2 = -xi*skphi_inc = -0
  const double TR1 = xj * cphi, TR2 = xj * sphi, TR3 = -SXJ;
  const double PR1 = -sphi, PR2 = cphi;                                          // PI = (0,1,0)
  const double PIKR = VR2, PRKI = PR1 * VI1;                                     // PR2*VI2 = 0, PR3 = 0
  const double TIKR = TI1 * VR1 + TI3 * VR3, TRKI = TR1 * VI1 + TR3 * VI3;
  const double E1 = PIKR * PRKI, E2 = TIKR * TRKI, E3 = TIKR * PRKI, E4 = PIKR * TRKI;
  double CF11 = E1 * CRPER + E2 * CRPAR, CF12 = -E3 * CRPER + E4 * CRPAR, CF21 = -E4 * CRPER + E3 * CRPAR, CF22 = E2 * CRPER + E1 * CRPAR;
  const double VP1 = -VI3 * VR2, VP2 = VI3 * VR1 - VI1 * VR3, VP3 = VI1 * VR2;
  double DMOD = VP1 * VP1 + VP2 * VP2 + VP3 * VP3; DMOD = DMOD * DMOD;
  if (DMOD == 0.0) { CF11 = CRPAR; CF22 = CRPER; DMOD = 1.0; }
  const double RDZ2 = U3 * U3, RDZ4 = RDZ2 * RDZ2;
  const double DCOEFF_0 = 1.0 / (8.0 * xi * xj * DMOD * RDZ4 * sigma2);
  const double arg = (U1 * U1 + U2 * U2) / (2.0 * sigma2 * RDZ2), DEX = arg > 88.0 ? 0.0 : exp(-arg);
  const double DCOEFF = DCOEFF_0 * FACT1 * FACT1 * DEX, AF = 0.5 * DCOEFF;
  const double AF11 = CF11 * CF11, AF12 = CF12 * CF12, AF21 = CF21 * CF21, AF22 = CF22 * CF22;
  K[0] = (AF11 + AF12 + AF21 + AF22) * AF; K[1] = (AF11 - AF12 + AF21 - AF22) * AF;
  K[3] = (AF11 - AF22 + AF12 - AF21) * AF; K[4] = (AF11 - AF12 - AF21 + AF22) * AF;
  const double CTTTP = CF11 * CF12, CTTPT = CF11 * CF21, CTTPP = CF11 * CF22, CTPPT = CF12 * CF21, CTPPP = CF12 * CF22, CPTPP = CF21 * CF22;
  K[2] = -(CTTTP + CPTPP) * DCOEFF; K[5] = -(CTTTP - CPTPP) * DCOEFF; K[6] = -(CTTPT + CTPPP) * DCOEFF;
  K[7] = -(CTTPT - CTPPP) * DCOEFF; K[8] = (CTTPP + CTPPT) * DCOEFF;
  if (pars[2] == 0.0) return;
  const double S1 = sqrt(2.0 * sigma2 / PIE), S3 = 1.0 / sqrt(2.0 * sigma2), S2 = S3 * S3;
  double shadowi = 0.0, shadowr = 0.0;
  if (xi != 1.0) { const double dcot = xi / sqrt(1.0 - xi * xi), T1 = exp(-dcot * dcot * S2), T2 = derfc_e(dcot * S3); shadowi = 0.5 * (S1 * T1 / dcot - T2); }
  if (xj != 1.0) { const double dcot = xj / sqrt(1.0 - xj * xj), T1 = exp(-dcot * dcot * S2), T2 = derfc_e(dcot * S3); shadowr = 0.5 * (S1 * T1 / dcot - T2); }
  const double shadow = 1.0 / (1.0 + shadowi + shadowr);
#pragma unroll
  for (int t = 0; t < 9; t++) K[t] *= shadow;
}

// BPDF-NDVI (Maignan et al. 2009; BPDFNDVI_VFUNCTION, VB/vbrdf_sup_kernels.f90:3017-3198, with FRESNEL_VECTOR :3560-3702): Fresnel
// reflection matrix of a facet of refractive index pars[0], scaled by C exp(-NDVI) exp(-tan(gamma)) / (4 (mu_i + mu_r));
// pars = (refractive index, NDVI, C).  3x3 block K[3*o1+o2] like the GISS kernel.
HDFN void brdf_bpdf_kernel(const double *pars, double XI, double SXI, double XJ, double SXJ, double cphi, double sphi, double *K) {
  double xi = XI, xj = XJ;
  if (fabs(xi - 1.0) < 1e-9) xi = 0.999999999999;
  if (fabs(xj - 1.0) < 1e-9) xj = 0.999999999999;
  double Z = xi * xj - SXI * SXJ * cphi;
  if (Z > 1.0) Z = 1.0;
  const double Z1 = acos(Z), Z2 = cos(Z1 * 0.5);
  double NDVI = pars[1];
  if (NDVI > 1.0 || NDVI < -1.0) NDVI = 0.0;
  const double Fp0 = 0.25 * exp(-NDVI) / (xi + xj);
  const double cgamma = Z2, sgamma = sqrt(1.0 - cgamma * cgamma), atten = exp(-sgamma / cgamma);
  double HF = pars[2] * Fp0 * atten;
  // FRESNEL_VECTOR (CN1 = 1, CN2 = pars[0]; incident azimuth 0)
  const double VI1 = SXI, VI3 = -xi;
  const double VR1 = SXJ * cphi, VR2 = SXJ * sphi, VR3 = xj;
  const double U1 = VI1 - VR1, U2 = -VR2, U3 = VI3 - VR3;
  const double FACT1 = U1 * U1 + U2 * U2 + U3 * U3, FACTOR = sqrt(1.0 / FACT1);
  const double CN2 = pars[0];
  const double XI1 = FACTOR * (U1 * VI1 + U3 * VI3);
  const double CXI2 = sqrt(1.0 - (1.0 - XI1 * XI1) / (CN2 * CN2));
  double C1 = XI1, C2 = CN2 * CXI2; const double CRPER = (C1 - C2) / (C1 + C2);
  C1 = CN2 * XI1; C2 = CXI2; const double CRPAR = (C1 - C2) / (C1 + C2);
  const double TI1 = -xi, TI3 = -SXI;
  const double TR1 = xj * cphi, TR3 = -SXJ;
  const double PR1 = -sphi;
  const double PIKR = VR2, PRKI = PR1 * VI1;
  const double TIKR = TI1 * VR1 + TI3 * VR3, TRKI = TR1 * VI1 + TR3 * VI3;
  const double E1 = PIKR * PRKI, E2 = TIKR * TRKI, E3 = TIKR * PRKI, E4 = PIKR * TRKI;
  double CF11 = E1 * CRPER + E2 * CRPAR, CF12 = -E3 * CRPER + E4 * CRPAR, CF21 = -E4 * CRPER + E3 * CRPAR, CF22 = E2 * CRPER + E1 * CRPAR;
  const double VP1 = -VI3 * VR2, VP2 = VI3 * VR1 - VI1 * VR3, VP3 = VI1 * VR2;
  double DMOD = VP1 * VP1 + VP2 * VP2 + VP3 * VP3; DMOD = DMOD * DMOD;
  if (DMOD == 0.0) { CF11 = CRPAR; CF22 = CRPER; DMOD = 1.0; }
  HF = HF / DMOD;
  const double AF11 = CF11 * CF11, AF12 = CF12 * CF12, AF21 = CF21 * CF21, AF22 = CF22 * CF22, AF = 0.5 * HF;
  K[0] = (AF11 + AF12 + AF21 + AF22) * AF; K[1] = (AF11 - AF12 + AF21 - AF22) * AF;
  K[3] = (AF11 - AF22 + AF12 - AF21) * AF; K[4] = (AF11 - AF12 - AF21 + AF22) * AF;
  const double CTTTP = CF11 * CF12, CTTPT = CF11 * CF21, CTTPP = CF11 * CF22, CTPPT = CF12 * CF21, CTPPP = CF12 * CF22, CPTPP = CF21 * CF22;
  K[2] = -(CTTTP + CPTPP) * HF; K[5] = -(CTTTP - CPTPP) * HF; K[6] = -(CTTPT + CTPPP) * HF;
  K[7] = -(CTTPT - CTPPP) * HF; K[8] = (CTTPP + CTPPT) * HF;
}
// kernels with a Mueller block (which = 10 GISS Cox-Munk, 14 BPDF-NDVI)
HDFN void brdf_vec_kernel(int which, const double *pars, double XI, double SXI, double XJ, double SXJ, double cphi, double sphi, double *K) {
  if (which == 10) brdf_giss_kernel(pars, XI, SXI, XJ, SXJ, cphi, sphi, K);
  else brdf_bpdf_kernel(pars, XI, SXI, XJ, SXJ, cphi, sphi, K);
}

struct SurfPars { int nk, which[4], lam[4]; double fac[4], pars[4][3]; };

// surface parameters of pixel (obs j, channel i) as the glue sets them
HDFN void surf_params(const VLBrdf &B, int j, int i, SurfPars &P) {
  for (int k = 0; k < 4; k++) { P.lam[k] = 0; P.fac[k] = 0.0; P.which[k] = 0; for (int t = 0; t < 3; t++) P.pars[k][t] = 0.0; }
  if (B.kind == SURF_LAMBERT_BPDF) {                                           // VLIDORT_LAMBERTIAN_BPDF, V283/VLIDORT_SurfaceMod.F90:2065-2275
    const double *bp = B.bpdf + 3 * (i + (size_t)B.nch * j);
    P.nk = 2; P.which[0] = 1; P.lam[0] = 1; P.fac[0] = B.albedo[j + (size_t)B.nobs * i];
    P.which[1] = 14; P.fac[1] = 1.0; P.pars[1][0] = bp[0]; P.pars[1][1] = bp[1]; P.pars[1][2] = bp[2];
    return;
  }
  if (B.kind == SURF_MODIS_BPDF) {                                             // VLIDORT_LANDMODIS_BPDF, :1826-2060: RTLS + BPDF-NDVI
    const double *kw = B.kernel_wt + 3 * (i + (size_t)B.nch * j), *pr = B.param + 2 * (i + (size_t)B.nch * j), *bp = B.bpdf + 3 * (i + (size_t)B.nch * j);
    P.nk = 4; P.which[0] = 1; P.which[1] = 3; P.which[2] = 4; P.which[3] = 14; P.lam[0] = 1;
    P.fac[0] = kw[0]; P.fac[1] = kw[2]; P.fac[2] = kw[1]; P.fac[3] = 1.0;
    P.pars[2][0] = pr[0]; P.pars[2][1] = pr[1];
    P.pars[3][0] = bp[0]; P.pars[3][1] = bp[1]; P.pars[3][2] = bp[2];
    return;
  }
  if (B.kind == SURF_MODIS) {
    const double *kw = B.kernel_wt + 3 * (i + (size_t)B.nch * j), *pr = B.param + 2 * (i + (size_t)B.nch * j);
    P.nk = 3; P.which[0] = 1; P.which[1] = 3; P.which[2] = 4; P.lam[0] = 1;
    P.fac[0] = kw[0]; P.fac[1] = kw[2]; P.fac[2] = kw[1];                       // (fiso, fvol, fgeo)
    P.pars[2][0] = pr[0]; P.pars[2][1] = pr[1];
    return;
  }
  if (B.kind == SURF_BPDF) {                                                   // VLIDORT_BPDF, V283/VLIDORT_SurfaceMod.F90:2279-2480
    const double *pr = B.param + 3 * (i + (size_t)B.nch * j);
    P.nk = 1; P.which[0] = 14; P.fac[0] = 1.0;
    P.pars[0][0] = pr[0]; P.pars[0][1] = pr[1]; P.pars[0][2] = pr[2];
    return;
  }
  double W = sqrt(B.u10m[j] * B.u10m[j] + B.v10m[j] * B.v10m[j]);
  if (W < 0) W = 0;
  if (W > 50) W = 50;
  const double sigma2 = surf_is_oci(B.kind) ? (double)0.00534f * W : (double)0.003f + (double)0.00512f * W;   // single-precision literals of the glue
  P.nk = 1; P.fac[0] = 1.0; P.pars[0][2] = 1.0;                                 // shadow effect on
  if (!surf_is_giss(B.kind)) { P.which[0] = 9; P.pars[0][0] = sigma2; P.pars[0][1] = B.mr[i] * B.mr[i]; }
  else { P.which[0] = 10; P.pars[0][0] = 0.5 * sigma2; P.pars[0][1] = B.mr[i]; }
}

// element list of the 3x3 block in the 16-element Mueller layout, and the cos/sin type of each element
// (COSSIN_MASK of vbrdf_sup_routines.f90:2151: 1 cos, 2 -sin, 3 +sin)
HDFN int brdf_q16(int t) { return 4 * (t / 3) + (t % 3); }
HDFN int brdf_cossin(int t) { return (t == 2 || t == 5) ? 2 : ((t == 6 || t == 7) ? 3 : 1); }

// exact direct-bounce BRDF of one pixel-view p = (j*nch + i)*G + g (thread-level)
HDFN void brdf_exact_pixel(const VLBrdf &B, int p) {
  const int G = B.G > 0 ? B.G : 1, pc = p / G, g = p % G;
  const int j = pc / B.nch, i = pc % B.nch;
  SurfPars P; surf_params(B, j, i, P);
  const double mu0 = cos(B.sza[j] * DEG_TO_RAD), smu0 = sqrt(1.0 - mu0 * mu0);
  const double muv = cos(B.vza[j + (size_t)B.nobs * g] * DEG_TO_RAD), smuv = sqrt(1.0 - muv * muv);
  const double phi = B.raa[j + (size_t)B.nobs * g] * DEG_TO_RAD, cphi = cos(phi), sphi = sin(phi);
  double ex[16];
  for (int q = 0; q < 16; q++) ex[q] = 0.0;
  for (int k = 0; k < P.nk; k++) {
    if (brdf_vec_which(P.which[k])) {
      double K[9]; brdf_vec_kernel(P.which[k], P.pars[k], mu0, smu0, muv, smuv, cphi, sphi, K);
      if (B.bns == 1) ex[0] += P.fac[k] * K[0];
      else for (int t = 0; t < 9; t++) ex[brdf_q16(t)] += P.fac[k] * K[t];
    } else ex[0] += P.fac[k] * brdf_scalar_kernel(P.which[k], P.pars[k], mu0, smu0, muv, smuv, cphi, sphi);
  }
  for (int q = 0; q < 16; q++) B.exactdb[(size_t)p * 16 + q] = ex[q];
}

HDFN size_t brdf_out_f(const VLBrdf &B) { return (size_t)2 * B.N * B.nq * B.N * B.N; }
HDFN size_t brdf_out_f0(const VLBrdf &B) { return (size_t)2 * B.N * B.nq * B.N; }

// Fourier components of one pixel (all its views).  pc = chunk-local pixel, p = global pixel.  scr: nscr doubles of team scratch
// (>= T.size + batch storage), batch storage = EB * nqs * NB with EB >= 1.
template <class Team>
DEVFN void brdf_pixel(const Team &T, const VLBrdf &B, int p, int pc, double *scr, int nscr) {
  const int N = B.N, NB = B.NB, NM = B.NM, j = p / B.nch, i = p % B.nch, G = B.G > 0 ? B.G : 1;
  SurfPars P; surf_params(B, j, i, P);
  const double mu0 = cos(B.sza[j] * DEG_TO_RAD), smu0 = sqrt(1.0 - mu0 * mu0);
  double *OF = B.brdf_f + (size_t)pc * brdf_out_f(B), *OF0 = B.brdf_f_0 + (size_t)pc * brdf_out_f0(B), *OU = B.user_brdf_f + (size_t)pc * G * brdf_out_f0(B);
  for (size_t t = T.rank; t < brdf_out_f(B); t += T.size) OF[t] = 0.0;
  for (size_t t = T.rank; t < brdf_out_f0(B); t += T.size) OF0[t] = 0.0;
  for (size_t t = T.rank; t < (size_t)G * brdf_out_f0(B); t += T.size) OU[t] = 0.0;
  double *red = scr, *vals = scr + T.size;
  const int nqs = (B.nq == 16) ? 9 : 1;
  const int EB = (nscr - T.size) / (nqs * NB);
  const int NE = N * N + N + G * N;                  // entries: (i,j) stream pairs, sun -> stream i, stream i -> user view g
  const double help_w = B.half_range ? 1.0 : 0.5;
  // ---- white-sky albedo check over all kernels (serious status -> the last kernel contributes no Fourier terms)
  double total_wsa = 0.0;
  for (int k = 0; k < P.nk; k++) {
    double wsa_k = 1.0;
    if (!P.lam[k]) {
      double part = 0.0;
      const int tot = BRDF_NSCAL * BRDF_NSCAL * NB;
      for (int idx = T.rank; idx < tot; idx += T.size) {
        const int b = idx % NB, ij = idx / NB, jj = ij % BRDF_NSCAL, ii = ij / BRDF_NSCAL;
        double v;
        if (brdf_vec_which(P.which[k])) { double K[9]; brdf_vec_kernel(P.which[k], P.pars[k], B.sx[jj], B.ss[jj], B.sx[ii], B.ss[ii], B.cxb[b], B.sxb[b], K); v = K[0]; }
        else v = brdf_scalar_kernel(P.which[k], P.pars[k], B.sx[jj], B.ss[jj], B.sx[ii], B.ss[ii], B.cxb[b], B.sxb[b]);
        part += v * B.ab[b] * B.ssw[jj] * B.ssw[ii];
      }
      T.sync();
      red[T.rank] = part;
      T.sync();
      double s = 0.0;
      for (int t = 0; t < T.size; t++) s += red[t];
      wsa_k = help_w * s * 4.0 / PIE;
    }
    total_wsa += P.fac[k] * wsa_k;
  }
  const bool wsa_bad = (total_wsa <= 0.0 || total_wsa > 1.0);
  if (T.rank == 0) B.wsa_fail[p] = wsa_bad ? 1 : 0;
  T.sync();
  // ---- kernels
  for (int k = 0; k < P.nk; k++) {
    if (wsa_bad && k == P.nk - 1) break;
    const double f = P.fac[k];
    if (P.lam[k]) {
      for (int t = T.rank; t < N * N; t += T.size) OF[t] += f;                   // m = 0, element (1,1)
      for (int t = T.rank; t < N; t += T.size) OF0[t] += f;
      for (int t = T.rank; t < G * N; t += T.size) OU[(size_t)(t / N) * brdf_out_f0(B) + t % N] += f;
      T.sync();
      continue;
    }
    const int kq = brdf_vec_which(P.which[k]) ? nqs : 1;       // Mueller elements this kernel fills: a scalar kernel only has (1,1)
    for (int e0 = 0; e0 < NE; e0 += EB) {
      const int ne = (NE - e0 < EB) ? NE - e0 : EB;
      // kernel values on the azimuth quadrature
      for (int idx = T.rank; idx < ne * NB; idx += T.size) {
        const int b = idx % NB, el = idx / NB, e = e0 + el;
        double XI, SXI, XJ, SXJ;                                                // incident, reflected
        if (e < N * N) { const int ir = e / N, jc = e % N; XI = B.qx[jc]; SXI = B.qs[jc]; XJ = B.qx[ir]; SXJ = B.qs[ir]; }
        else if (e < N * N + N) { const int ir = e - N * N; XI = mu0; SXI = smu0; XJ = B.qx[ir]; SXJ = B.qs[ir]; }
        else {
          const int u = e - N * N - N, jc = u % N, g = u / N;
          const double muv = cos(B.vza[j + (size_t)B.nobs * g] * DEG_TO_RAD);
          XI = B.qx[jc]; SXI = B.qs[jc]; XJ = muv; SXJ = sqrt(1.0 - muv * muv);
        }
        if (brdf_vec_which(P.which[k])) {
          double K[9]; brdf_vec_kernel(P.which[k], P.pars[k], XI, SXI, XJ, SXJ, B.cxb[b], B.sxb[b], K);
          if (kq == 1) vals[(size_t)el * NB + b] = K[0];
          else for (int t = 0; t < 9; t++) vals[((size_t)el * 9 + t) * NB + b] = K[t];
        } else vals[(size_t)el * NB + b] = brdf_scalar_kernel(P.which[k], P.pars[k], XI, SXI, XJ, SXJ, B.cxb[b], B.sxb[b]);
      }
      T.sync();
      // projection on the Fourier components
      const int ncombo = ne * kq * (NM + 1);
      for (int idx = T.rank; idx < ncombo; idx += T.size) {
        const int el = idx % ne, r = idx / ne, t = r % kq, m = r / kq, e = e0 + el;
        const double *v = vals + ((size_t)el * kq + t) * NB;
        const int cs = brdf_cossin(t);
        const double *tab = (cs == 1 ? B.caz : B.saz) + (size_t)m * NB;
        double s = 0.0;
        for (int b = 0; b < NB; b++) s += v[b] * tab[b];
        const double delfac = (m == 0) ? 1.0 : 2.0, help = B.half_range ? delfac : 0.5 * delfac;
        s = (cs == 2 ? -help : help) * s;
        const int q = (B.nq == 16) ? brdf_q16(t) : 0;
        if (e < N * N) OF[((size_t)m * B.nq + q) * N * N + e] += f * s;
        else if (e < N * N + N) OF0[((size_t)m * B.nq + q) * N + (e - N * N)] += f * s;
        else { const int u = e - N * N - N; OU[(size_t)(u / N) * brdf_out_f0(B) + ((size_t)m * B.nq + q) * N + u % N] += f * s; }
      }
      T.sync();
    }
  }
}

}  // namespace vlk
