/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the VBRDF supplement as the AeroApps surface glue uses it
 * (observation geometry, one geometry, solar source, no emission, no MSR correction, no albedo scaling).
 * Restates (relative to /root/reference/src/Components/rtms/RTSI/VLIDORT2p8p3/vsup/vbrdf/):
 *   vbrdf_sup_masters.f90:1559-3295   VBRDF_MAINMASTER (kernel loop, WSA check :2953-2998, Fourier loop :3040-3180)
 *   vbrdf_sup_routines.f90:153-690    VBRDF_MAKER ; :1962-2069 SCALING_FOURIER_ZERO ; :2073-2491 VBRDF_FOURIER
 *   vbrdf_sup_aux.f90:266-286,356-461 derfc_e, BRDF_QUADRATURE_GAUSSIAN
 *   vbrdf_sup_kernels.f90:201-236 (Lambertian) 287-335 (Ross-thick) 478-585 (Li-sparse)
 *                         1018-1144 (Cox-Munk) 1145-1495 (GISS Cox-Munk)
 * Pinned against V283/vlidort_v_test/saved_results/gfortran/nstokes3/results_brdf_self_tester.res
 * (tests/test_vbrdf_oracle_golden.py).
 */
#include "vl_numerics.h"
#include "vlidort_oracle.h"
#include <cstdio>

using namespace vl;

namespace {

double derfc_e(double x) {
  double z = std::fabs(x), t = 1.0 / (1.0 + 0.5 * z);
  double r = t * std::exp(-z * z - 1.26551223 + t * (1.00002368 + t * (.37409196 + t * (.09678418 + t * (-.18628806 + t * (.27886807 + t *
             (-1.13520398 + t * (1.48851587 + t * (-.82215223 + t * .17087277)))))))));
  if (x < 0.0) r = 2.0 - r;
  return r;
}

/* kernel(16) for incident (xj, sxj) -> reflected (xi, sxi) at azimuth phi.  NB the Fortran dummies are
 * (XI,SXI,XJ,SXJ) with the FIRST pair being the incident direction of the caller. */
void kernel_eval(int which, const double *pars, double xinc, double sxinc, double xref, double sxref, double phi, double cphi,
                 double sphi, double *K) {
  for (int q = 0; q < 16; q++) K[q] = 0.0;
  const double PIO2 = 0.5 * PIE, PIO4 = 0.25 * PIE;
  const double XI = xinc, SXI = sxinc, XJ = xref, SXJ = sxref;
  if (which == 1) { K[0] = 1.0; return; }                             /* LAMBERTIAN_VFUNCTION */
  if (which == 2) {                                                   /* ROSSTHIN_VFUNCTION (:237; golden coverage only) */
    double ckphi = -cphi, ds1 = XI * XJ, ds2 = SXI * SXJ;
    double cksi = ds1 + ds2 * ckphi; if (cksi > 1.0) cksi = 1.0;
    double sksi = std::sqrt(1.0 - cksi * cksi), ksi = std::acos(cksi);
    K[0] = ((PIO2 - ksi) * cksi + sksi) / ds1 - PIO2;
    return;
  }
  if (which == 3) {                                                   /* ROSSTHICK_VFUNCTION */
    double ckphi = -cphi, ds1 = XI * XJ, ds2 = SXI * SXJ, ds3 = XI + XJ;
    double cksi = ds1 + ds2 * ckphi; if (cksi > 1.0) cksi = 1.0;
    double sksi = std::sqrt(1.0 - cksi * cksi), ksi = std::acos(cksi);
    K[0] = ((PIO2 - ksi) * cksi + sksi) / ds3 - PIO4;
    return;
  }
  if (which == 4 || which == 5) {                                     /* LISPARSE_VFUNCTION / LIDENSE_VFUNCTION (:586) */
    double ckphi = -cphi;
    double t_inc = pars[1] * (SXI / XI), ang = std::atan(t_inc), x_inc = std::cos(ang), sx_inc = std::sin(ang);
    double t_ref = pars[1] * (SXJ / XJ); ang = std::atan(t_ref); double x_ref = std::cos(ang), sx_ref = std::sin(ang);
    double cksi = x_inc * x_ref + sx_inc * sx_ref * ckphi;
    double P = (1.0 + cksi) / x_ref / x_inc, A = 1.0 / x_inc, B = 1.0 / x_ref, R = A + B;
    double dt1 = t_ref * t_ref + t_inc * t_inc, dt2 = t_inc * t_ref, dt2sq = dt2 * dt2;
    double delta = std::sqrt(dt1 - 2.0 * dt2 * ckphi), dsq = delta * delta;
    double H = std::sqrt(dsq + sphi * sphi * dt2sq), cost = pars[0] * H / R, Q;
    if (cost > 1.0) Q = 1.0;
    else { double T = std::acos(cost), sint = std::sqrt(1.0 - cost * cost); Q = 1.0 - ((T - sint * cost) / PIE); }
    K[0] = which == 4 ? 0.5 * P - Q * R : P / Q / R - 2.0;
    return;
  }
  if (which == 9) {                                                   /* COXMUNK_VFUNCTION */
    const double CRITEXP = 88.0;
    double ckphi_neg = -cphi;
    double Z = XI * XJ + SXI * SXJ * ckphi_neg; if (Z > 1.0) Z = 1.0;
    double Z1 = std::acos(Z), Z2 = std::cos(Z1 * 0.5), Z2m1 = Z2 * Z2 - 1.0;
    double H1 = pars[1] * Z2, H2 = std::sqrt(pars[1] + Z2m1);
    double RP = (H1 - H2) / (H1 + H2), RL = (Z2 - H2) / (Z2 + H2), XMP = 0.5 * (RP * RP + RL * RL);
    double A = 2.0 * Z2, B = (XI + XJ) / A; if (B > 1.0) B = 1.0;
    A = PIO2 - std::asin(B);
    double TA = std::tan(A), arg = TA * TA / pars[0];
    if (arg < CRITEXP) {
      double prob = std::exp(-arg), fac1 = prob / pars[0], fac2 = 0.25 / XI / std::pow(B, 4.0);
      K[0] = XMP * fac1 * fac2 / XJ;
    }
    if (pars[2] == 0.0) return;
    double S1 = std::sqrt(pars[0] / PIE), S3 = 1.0 / std::sqrt(pars[0]), S2 = S3 * S3;
    double dcot = XI / std::sqrt(1.0 - XI * XI), T1 = std::exp(-dcot * dcot * S2), T2 = derfc_e(dcot * S3);
    double shadowi = 0.5 * (S1 * T1 / dcot - T2);
    dcot = XJ / std::sqrt(1.0 - XJ * XJ); T1 = std::exp(-dcot * dcot * S2); T2 = derfc_e(dcot * S3);
    double shadowr = 0.5 * (S1 * T1 / dcot - T2);
    K[0] *= 1.0 / (1.0 + shadowi + shadowr);
    return;
  }
  if (which == 14) {                                                  /* BPDFNDVI_VFUNCTION (VB/vbrdf_sup_kernels.f90:3017-3198) */
    double xi = XI, xj = XJ;
    const double CN1 = 1.0, CN2 = pars[0];
    if (std::fabs(xi - 1.0) < 1e-9) xi = 0.999999999999;
    if (std::fabs(xj - 1.0) < 1e-9) xj = 0.999999999999;
    double Z = xi * xj - SXI * SXJ * cphi; if (Z > 1.0) Z = 1.0;
    double Z1 = std::acos(Z), Z2 = std::cos(Z1 * 0.5);
    double NDVI = pars[1];
    if (NDVI > 1.0 || NDVI < -1.0) NDVI = 0.0;
    double DEXPNDVI = std::exp(-NDVI);
    double Fp0 = 0.25 * DEXPNDVI / (xi + xj);
    double cgamma = Z2, sgamma = std::sqrt(1.0 - cgamma * cgamma), atten = std::exp(-sgamma / cgamma);
    double C = pars[2];
    double HFUNCTION = C * Fp0 * atten;
    /* FRESNEL_VECTOR (:3560-3702) */
    const double ckphi_inc = 1.0, skphi_inc = 0.0, ckphi_ref = cphi, skphi_ref = sphi;
    double VI1 = SXI * ckphi_inc, VI2 = SXI * skphi_inc, VI3 = -xi;
    double VR1 = SXJ * ckphi_ref, VR2 = SXJ * skphi_ref, VR3 = xj;
    double U1 = VI1 - VR1, U2 = VI2 - VR2, U3 = VI3 - VR3;
    double FACT1 = U1 * U1 + U2 * U2 + U3 * U3, FACTOR = std::sqrt(1.0 / FACT1);
    double XI1 = FACTOR * (U1 * VI1 + U2 * VI2 + U3 * VI3);
    double CXI2 = 1.0 - (1.0 - XI1 * XI1) * CN1 * CN1 / (CN2 * CN2); CXI2 = std::sqrt(CXI2);
    double C1 = CN1 * XI1, C2 = CN2 * CXI2, CRPER = (C1 - C2) / (C1 + C2);
    C1 = CN2 * XI1; C2 = CN1 * CXI2; double CRPAR = (C1 - C2) / (C1 + C2);
    double TI1 = -xi * ckphi_inc, TI2 = -xi * skphi_inc, TI3 = -SXI;
    double TR1 = xj * ckphi_ref, TR2 = xj * skphi_ref, TR3 = -SXJ;
    double PI1 = -skphi_inc, PII2 = ckphi_inc, PI3 = 0.0, PR1 = -skphi_ref, PR2 = ckphi_ref, PR3 = 0.0;
    double PIKR = PI1 * VR1 + PII2 * VR2 + PI3 * VR3, PRKI = PR1 * VI1 + PR2 * VI2 + PR3 * VI3;
    double TIKR = TI1 * VR1 + TI2 * VR2 + TI3 * VR3, TRKI = TR1 * VI1 + TR2 * VI2 + TR3 * VI3;
    double E1 = PIKR * PRKI, E2 = TIKR * TRKI, E3 = TIKR * PRKI, E4 = PIKR * TRKI;
    double CF11 = E1 * CRPER + E2 * CRPAR, CF12 = -E3 * CRPER + E4 * CRPAR, CF21 = -E4 * CRPER + E3 * CRPAR, CF22 = E2 * CRPER + E1 * CRPAR;
    double VP1 = VI2 * VR3 - VI3 * VR2, VP2 = VI3 * VR1 - VI1 * VR3, VP3 = VI1 * VR2 - VI2 * VR1;
    double DMOD = VP1 * VP1 + VP2 * VP2 + VP3 * VP3; DMOD = DMOD * DMOD;
    if (DMOD == 0.0) { CF11 = CRPAR; CF22 = CRPER; DMOD = 1.0; }
    HFUNCTION = HFUNCTION / DMOD;
    double AF11 = std::fabs(CF11), AF12 = std::fabs(CF12), AF21 = std::fabs(CF21), AF22 = std::fabs(CF22);
    AF11 = AF11 * AF11; AF12 = AF12 * AF12; AF21 = AF21 * AF21; AF22 = AF22 * AF22;
    double F = 0.5 * HFUNCTION;
    K[0] = (AF11 + AF12 + AF21 + AF22) * F; K[1] = (AF11 - AF12 + AF21 - AF22) * F;
    K[4] = (AF11 - AF22 + AF12 - AF21) * F; K[5] = (AF11 - AF12 - AF21 + AF22) * F;
    double CTTTP = CF11 * CF12, CTTPT = CF11 * CF21, CTTPP = CF11 * CF22, CTPPT = CF12 * CF21, CTPPP = CF12 * CF22, CPTPP = CF21 * CF22;
    F = HFUNCTION;
    K[2] = -(CTTTP + CPTPP) * F; K[6] = -(CTTTP - CPTPP) * F; K[8] = -(CTTPT + CTPPP) * F;
    K[9] = -(CTTPT - CTPPP) * F; K[10] = (CTTPP + CTPPT) * F; K[15] = (CTTPP - CTPPT) * F;
    return;
  }
  if (which == 10) {                                                  /* GISSCOXMUNK_VFUNCTION */
    const double CRITEXP = 88.0;
    double xi = XI, xj = XJ;
    const double sigma2 = pars[0];
    if (std::fabs(xi - 1.0) < 1e-9) xi = 0.999999999999;
    if (std::fabs(xj - 1.0) < 1e-9) xj = 0.999999999999;
    const double ckphi_inc = 1.0, skphi_inc = 0.0, ckphi_ref = cphi, skphi_ref = sphi;
    double VI1 = SXI * ckphi_inc, VI2 = SXI * skphi_inc, VI3 = -xi;
    double VR1 = SXJ * ckphi_ref, VR2 = SXJ * skphi_ref, VR3 = xj;
    double U1 = VI1 - VR1, U2 = VI2 - VR2, U3 = VI3 - VR3;
    double FACT1 = U1 * U1 + U2 * U2 + U3 * U3, FACTOR = std::sqrt(1.0 / FACT1);
    double CN1 = 1.0, CN2 = pars[1];
    double XI1 = FACTOR * (U1 * VI1 + U2 * VI2 + U3 * VI3);
    double CXI2 = std::sqrt(1.0 - (1.0 - XI1 * XI1) * CN1 * CN1 / (CN2 * CN2));
    double C1 = CN1 * XI1, C2 = CN2 * CXI2, CRPER = (C1 - C2) / (C1 + C2);
    C1 = CN2 * XI1; C2 = CN1 * CXI2; double CRPAR = (C1 - C2) / (C1 + C2);
    double TI1 = -xi * ckphi_inc, TI2 = -xi * skphi_inc, TI3 = -SXI;
    double TR1 = xj * ckphi_ref, TR2 = xj * skphi_ref, TR3 = -SXJ;
    double PI1 = -skphi_inc, PII2 = ckphi_inc, PI3 = 0.0, PR1 = -skphi_ref, PR2 = ckphi_ref, PR3 = 0.0;
    double PIKR = PI1 * VR1 + PII2 * VR2 + PI3 * VR3, PRKI = PR1 * VI1 + PR2 * VI2 + PR3 * VI3;
    double TIKR = TI1 * VR1 + TI2 * VR2 + TI3 * VR3, TRKI = TR1 * VI1 + TR2 * VI2 + TR3 * VI3;
    double E1 = PIKR * PRKI, E2 = TIKR * TRKI, E3 = TIKR * PRKI, E4 = PIKR * TRKI;
    double CF11 = E1 * CRPER + E2 * CRPAR, CF12 = -E3 * CRPER + E4 * CRPAR, CF21 = -E4 * CRPER + E3 * CRPAR, CF22 = E2 * CRPER + E1 * CRPAR;
    double VP1 = VI2 * VR3 - VI3 * VR2, VP2 = VI3 * VR1 - VI1 * VR3, VP3 = VI1 * VR2 - VI2 * VR1;
    double DMOD = VP1 * VP1 + VP2 * VP2 + VP3 * VP3; DMOD = DMOD * DMOD;
    if (DMOD == 0.0) { CF11 = CRPAR; CF22 = CRPER; DMOD = 1.0; }
    double RDZ2 = U3 * U3, RDZ4 = RDZ2 * RDZ2;
    double DCOEFF_0 = 1.0 / (8.0 * xi * xj * DMOD * RDZ4 * sigma2);
    double arg = (U1 * U1 + U2 * U2) / (2.0 * sigma2 * RDZ2), DEX = arg > CRITEXP ? 0.0 : std::exp(-arg);
    double DCOEFF = DCOEFF_0 * FACT1 * FACT1 * DEX, AF = 0.5 * DCOEFF;
    double AF11 = CF11 * CF11, AF12 = CF12 * CF12, AF21 = CF21 * CF21, AF22 = CF22 * CF22;
    K[0] = (AF11 + AF12 + AF21 + AF22) * AF; K[1] = (AF11 - AF12 + AF21 - AF22) * AF;
    K[4] = (AF11 - AF22 + AF12 - AF21) * AF; K[5] = (AF11 - AF12 - AF21 + AF22) * AF;
    double CTTTP = CF11 * CF12, CTTPT = CF11 * CF21, CTTPP = CF11 * CF22, CTPPT = CF12 * CF21, CTPPP = CF12 * CF22, CPTPP = CF21 * CF22;
    K[2] = -(CTTTP + CPTPP) * DCOEFF; K[6] = -(CTTTP - CPTPP) * DCOEFF; K[8] = -(CTTPT + CTPPP) * DCOEFF;
    K[9] = -(CTTPT - CTPPP) * DCOEFF; K[10] = (CTTPP + CTPPT) * DCOEFF; K[15] = (CTTPP - CTPPT) * DCOEFF;
    if (pars[2] == 0.0) return;
    double S1 = std::sqrt(2.0 * sigma2 / PIE), S3 = 1.0 / std::sqrt(2.0 * sigma2), S2 = S3 * S3, shadowi = 0.0, shadowr = 0.0;
    if (xi != 1.0) { double dcot = xi / std::sqrt(1.0 - xi * xi), T1 = std::exp(-dcot * dcot * S2), T2 = derfc_e(dcot * S3); shadowi = 0.5 * (S1 * T1 / dcot - T2); }
    if (xj != 1.0) { double dcot = xj / std::sqrt(1.0 - xj * xj), T1 = std::exp(-dcot * dcot * S2), T2 = derfc_e(dcot * S3); shadowr = 0.5 * (S1 * T1 / dcot - T2); }
    double shadow = 1.0 / (1.0 + shadowi + shadowr);
    static const int KM[10] = {0, 1, 2, 4, 5, 6, 8, 9, 10, 15};
    for (int t = 0; t < 10; t++) K[KM[t]] *= shadow;
    return;
  }
}

}  // namespace

/* outputs: brdf_f [2N][16][N][N], brdf_f_0 [2N][16][N], user_brdf_f [2N][16][N], user_brdf_f_0 [2N][16], exactdb[16]
 * returns 0, or 4 when the white-sky-albedo check fails (then only exactdb is filled, like the reference) */
extern "C" int orc_vbrdf(int nstokes, int nstreams, int nkernels, const int *which, const double *factors,
                         const double *pars /*[k][4]*/, const int *lambertian_flag, int nstreams_brdf_in, int do_shadow,
                         double sza, double vza, double raa, double *brdf_f, double *brdf_f_0, double *user_brdf_f,
                         double *user_brdf_f_0, double *exactdb, double *wsa_out) {
  const int N = nstreams, NMOM = 2 * N - 1, NSQ = 16;
  std::fill(brdf_f, brdf_f + (size_t)2 * N * 16 * N * N, 0.0); std::fill(brdf_f_0, brdf_f_0 + (size_t)2 * N * 16 * N, 0.0);
  std::fill(user_brdf_f, user_brdf_f + (size_t)2 * N * 16 * N, 0.0); std::fill(user_brdf_f_0, user_brdf_f_0 + (size_t)2 * N * 16, 0.0);
  std::fill(exactdb, exactdb + 16, 0.0);
  bool half_range = true, modis_alone = false;
  for (int k = 0; k < nkernels; k++) if (which[k] >= 10 && which[k] != 17) half_range = false;
  if (nkernels == 1 && ((which[0] >= 2 && which[0] <= 5) || which[0] == 7)) modis_alone = true;
  const int NB = half_range ? nstreams_brdf_in / 2 : nstreams_brdf_in, NBH = NB / 2;
  const bool check_albedo = !modis_alone;
  const bool local_wsa = true || check_albedo;                       /* DO_WSABSA_OUTPUT = .true. in the glue */
  std::vector<double> qx(N), qw(N), qs(N);
  getquad2(0.0, 1.0, N, qx.data(), qw.data());
  for (int i = 0; i < N; i++) qs[i] = std::sqrt(1.0 - qx[i] * qx[i]);
  const int NSC = 24;
  std::vector<double> sx(NSC), sw(NSC), ss(NSC), ssw(NSC);
  getquad2(0.0, 1.0, NSC, sx.data(), sw.data());
  for (int i = 0; i < NSC; i++) { ss[i] = std::sqrt(1.0 - sx[i] * sx[i]); ssw[i] = sx[i] * sw[i]; }
  const double mu0 = std::cos(sza * DEG_TO_RAD), smu0 = std::sqrt(1.0 - mu0 * mu0);
  const double muv = std::cos(vza * DEG_TO_RAD), smuv = std::sqrt(1.0 - muv * muv);
  const double phiang = raa * DEG_TO_RAD, cphi = std::cos(phiang), sphi = std::sin(phiang);
  std::vector<double> xb(NB), ab(NB), cxb(NB), sxb(NB);
  if (half_range) getquad2(0.0, 1.0, NB, xb.data(), ab.data());
  else { getquad2(0.0, 1.0, NBH, xb.data(), ab.data()); for (int i = 0; i < NBH; i++) { xb[i + NBH] = -xb[i]; ab[i + NBH] = ab[i]; } }
  for (int i = 0; i < NB; i++) { xb[i] = PIE * xb[i]; cxb[i] = std::cos(xb[i]); sxb[i] = std::sin(xb[i]); }
  static const int QMASK3[9] = {0, 1, 2, 4, 5, 6, 8, 9, 10};
  static const int COSSIN[16] = {1, 1, 2, 0, 1, 1, 2, 0, 3, 3, 1, 0, 0, 0, 0, 1};
  double total_wsa = 0.0;
  std::vector<double> BF((size_t)16 * N * N * NB), BF0((size_t)16 * N * NB), UBF((size_t)16 * N * NB), UBF0((size_t)16 * NB), SCB((size_t)NSC * NSC * NB);
  double K[16];
  for (int k = 0; k < nkernels; k++) {
    double P[4] = {pars[4 * k], pars[4 * k + 1], pars[4 * k + 2], pars[4 * k + 3]};
    if (which[k] == 9 || which[k] == 10) P[2] = do_shadow ? 1.0 : 0.0;
    /* exact direct-bounce BRDF */
    kernel_eval(which[k], P, mu0, smu0, muv, smuv, phiang, cphi, sphi, K);
    for (int q = 0; q < 16; q++) exactdb[q] += factors[k] * K[q];
    /* white-sky albedo of this kernel (nssq = 1) */
    double wsa_k = 1.0;
    if (local_wsa) {
      if (!lambertian_flag[k]) {
        for (int i = 0; i < NSC; i++) for (int j = 0; j < NSC; j++) for (int b = 0; b < NB; b++) {
          kernel_eval(which[k], P, sx[j], ss[j], sx[i], ss[i], xb[b], cxb[b], sxb[b], K);
          SCB[((size_t)i * NSC + j) * NB + b] = K[0];
        }
        const double help = half_range ? 1.0 : 0.5;
        double help_a = 0.0;
        for (int i = 0; i < NSC; i++) {
          double sum = 0.0;
          for (int j = 0; j < NSC; j++) { double f = 0.0; for (int b = 0; b < NB; b++) f += SCB[((size_t)i * NSC + j) * NB + b] * ab[b]; sum += help * f * ssw[j]; }
          help_a += sum * ssw[i];
        }
        wsa_k = help_a * 4.0 / PIE;
      }
      total_wsa += factors[k] * wsa_k;
      if (k == nkernels - 1 && check_albedo) {
        if (wsa_out) *wsa_out = total_wsa;
        if (total_wsa <= 0.0 || total_wsa > 1.0) return 4;
      }
    }
    /* kernel values on the quadrature (VBRDF_MAKER) */
    if (!lambertian_flag[k]) {
      for (int i = 0; i < N; i++) for (int b = 0; b < NB; b++) {
        kernel_eval(which[k], P, mu0, smu0, qx[i], qs[i], xb[b], cxb[b], sxb[b], K);
        for (int q = 0; q < 16; q++) BF0[((size_t)q * N + i) * NB + b] = K[q];
        kernel_eval(which[k], P, qx[i], qs[i], muv, smuv, xb[b], cxb[b], sxb[b], K);         /* incident stream j=i, reflected user */
        for (int q = 0; q < 16; q++) UBF[((size_t)q * N + i) * NB + b] = K[q];
        for (int j = 0; j < N; j++) {
          kernel_eval(which[k], P, qx[j], qs[j], qx[i], qs[i], xb[b], cxb[b], sxb[b], K);
          for (int q = 0; q < 16; q++) BF[(((size_t)q * N + i) * N + j) * NB + b] = K[q];
        }
      }
      for (int b = 0; b < NB; b++) { kernel_eval(which[k], P, mu0, smu0, muv, smuv, xb[b], cxb[b], sxb[b], K); for (int q = 0; q < 16; q++) UBF0[(size_t)q * NB + b] = K[q]; }
    }
    /* Fourier components (VBRDF_FOURIER) */
    for (int M = 0; M <= NMOM; M++) {
      const double delfac = M == 0 ? 1.0 : 2.0, help = half_range ? delfac : 0.5 * delfac;
      std::vector<double> caz(NB), saz(NB);
      for (int b = 0; b < NB; b++) { caz[b] = M == 0 ? ab[b] : ab[b] * std::cos(M * xb[b]); saz[b] = M == 0 ? 0.0 : ab[b] * std::sin(M * xb[b]); }
      auto proj = [&](const double *f, int q) -> double {
        double s = 0.0;
        if (COSSIN[q] == 1) { for (int b = 0; b < NB; b++) s += f[b] * caz[b]; return help * s; }
        if (COSSIN[q] == 2) { for (int b = 0; b < NB; b++) s += f[b] * saz[b]; return -help * s; }
        if (COSSIN[q] == 3) { for (int b = 0; b < NB; b++) s += f[b] * saz[b]; return help * s; }
        return 0.0;
      };
      double *OF = brdf_f + (size_t)M * 16 * N * N, *OF0 = brdf_f_0 + (size_t)M * 16 * N, *OU = user_brdf_f + (size_t)M * 16 * N, *OU0 = user_brdf_f_0 + (size_t)M * 16;
      if (lambertian_flag[k]) {
        if (M == 0) {
          for (int i = 0; i < N; i++) { OF0[i] += factors[k]; OU[i] += factors[k]; for (int j = 0; j < N; j++) OF[(size_t)i * N + j] += factors[k]; }
          OU0[0] += factors[k];
        }
        continue;
      }
      for (int t = 0; t < nstokes * nstokes; t++) {
        const int q = nstokes == 1 ? 0 : QMASK3[(t / nstokes) * 3 + (t % nstokes)];
        for (int i = 0; i < N; i++) {
          OF0[(size_t)q * N + i] += factors[k] * proj(&BF0[((size_t)q * N + i) * NB], q);
          OU[(size_t)q * N + i] += factors[k] * proj(&UBF[((size_t)q * N + i) * NB], q);
          for (int j = 0; j < N; j++) OF[((size_t)q * N + i) * N + j] += factors[k] * proj(&BF[(((size_t)q * N + i) * N + j) * NB], q);
        }
        OU0[q] += factors[k] * proj(&UBF0[(size_t)q * NB], q);
      }
    }
  }
  (void)NSQ;
  return 0;
}
