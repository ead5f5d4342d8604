/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's geometry calculators, used by tests/ as the
 * parity checker of aeroapps_b200/csrc/geo_angles.cu; never linked or imported by the product.
 * PARITY UNPINNED: the reference holds no known-answer values for these routines (rtms/geo_vlidort/utAngles.py only prints), and no
 * Fortran compiler exists here to run them; every function follows the cited Fortran statement by statement, in the same operation
 * order, in IEEE double with no FMA contraction (oracle/Makefile builds this file with -ffp-contract=off).
 *   JULDAY          rtms/geo_vlidort/GeoAngles_py.F90:56-186   (= leo_vlidort/LidarAngles_py.F90:51-138, stn_vlidort/stnAngles_py.F90:51-138)
 *   solar_angles    GeoAngles_py.F90:305-457                   (= LidarAngles_py.F90:300-452, stnAngles_py.F90:179-331)
 *   view_angles     GeoAngles_py.F90:499-560                   (= LidarAngles_py.F90:494-584 view_angles_from_target)
 *   view_angles_from_space   LidarAngles_py.F90:620-717
 *   satAngles / satAngles_fromSpace / solarAngles / sunAngles  GeoAngles_py.F90:682-728, LidarAngles_py.F90:901-1056, stnAngles_py.F90:338-377 */
#include <math.h>
#include <stdlib.h>

static double orc_julday(int month, int day, int year, int hour, int minute, double second) {
  int bc = year < 0 ? 1 : 0;                 /* L_YEAR = YEAR + bc is computed by the reference and never used */
  (void)bc;
  int inJanFeb = month <= 2 ? 1 : 0;
  int jy = year - inJanFeb;
  int jm = month + 1 + 12 * inJanFeb;
  int jul = (int)floor(365.25 * jy) + (int)floor(30.6001 * jm) + day + 1720995;
  if (jul >= 2299171) {
    int ja = (int)floor(0.01 * jy);
    jul = jul + 2 - ja + (int)floor(0.25 * ja);
  }
  double eps = 2.2204460e-16;
  if (abs(jul) >= 1) eps = eps * abs(jul);
  return (double)jul + (double)hour / 24.0 - 0.5 + (double)minute / 1440.0 + second / 86400.0 + eps;
}

static void orc_solar_angles(int yr, int mo, int day, int hr, int min, double sec, double tz, double lat, double lon, double *az_out,
                             double *el_out) {
  double d2r = acos(-1.0) / 180.0;
  double r2d = 180.0 / acos(-1.0);
  double julianday = orc_julday(mo, day, yr, hr, min, sec);
  double t = (julianday - 2451545.0) / 36525.0 - tz / 24.0;
  double meanlon = 280.46646 + t * (36000.76983 + 0.0003032 * t);
  while (meanlon > 360.0) meanlon = meanlon - 360.0;
  while (meanlon < 0.0) meanlon = meanlon + 360.0;
  double mean_anom = 357.52911 + t * (35999.05029 - 0.0001537 * t);
  double mrad = d2r * mean_anom;
  double sinm = sin(mrad);
  double sin2m = sin(2.0 * mrad);
  double sin3m = sin(3.0 * mrad);
  double c = sinm * (1.914602 - t * (0.004817 + 0.000014 * t)) + sin2m * (0.019993 - 0.000101 * t) + sin3m * 0.000289;
  double truelon = meanlon + c;
  double omega = 125.04 - 1934.136 * t;
  double app_lon = truelon - 0.00569 - 0.00478 * sin(d2r * omega);
  double e = 0.016708634 - t * (0.000042037 + 0.0000001267 * t);
  double seconds = 21.448 - t * (46.8150 + t * (0.00059 - t * (0.001813)));
  double eps0 = 23.0 + (26.0 + (seconds / 60.0)) / 60.0;
  double epsilon = eps0 + 0.00256 * cos(d2r * omega);
  double sint = sin(d2r * epsilon) * sin(d2r * app_lon);
  double decl = r2d * asin(sint);
  double m = mean_anom;
  double lz = meanlon;
  double tn = tan(d2r * epsilon / 2.0);
  double y = tn * tn;
  double sin2lz = sin(2.0 * d2r * lz);
  sinm = sin(d2r * m);
  double cos2lz = cos(2.0 * d2r * lz);
  double sin4lz = sin(4.0 * d2r * lz);
  sin2m = sin(2.0 * d2r * m);
  double eqtime = y * sin2lz - 2.0 * e * sinm + 4.0 * e * y * sinm * cos2lz - 0.5 * y * y * sin4lz - 1.25 * e * e * sin2m;
  eqtime = r2d * eqtime * 4.0;
  double utc = (double)hr + ((double)min + sec / 60.0) / 60.0 - tz;
  double ha = 180.0 - 15.0 * utc - lon - eqtime / 4.0;
  double declr = d2r * decl;
  double latr = d2r * lat;
  double har = d2r * ha;
  double elr = asin(sin(latr) * sin(declr) + cos(latr) * cos(declr) * cos(har));
  double yaz = -sin(har) * cos(declr);
  double xaz = cos(har) * cos(declr) * sin(latr) - sin(declr) * cos(latr);
  double azr = atan2(-yaz, -xaz);
  double az = r2d * azr;
  if (az < 0.0) az = az + 360.0;
  *az_out = az;
  *el_out = r2d * elr;
}

static void orc_view_angles(double latp, double lonp, double latss, double lonss, double h, int from_space, double *a1, double *a2) {
  double d2r = acos(-1.0) / 180.0;
  double re = 6378.0;
  double srho = re / (re + h);
  double deltaL = fabs(lonss - lonp);
  double cdeltaL = cos(d2r * deltaL);
  double clatss = cos(d2r * latss);
  double slatss = sin(d2r * latss);
  double clatp = cos(d2r * latp);
  double slatp = sin(d2r * latp);
  double clambda = slatss * slatp + clatss * clatp * cdeltaL;
  double slambda = sin(acos(clambda));
  double cphiv;
  if (from_space) cphiv = (slatp - clambda * slatss) / (slambda * clatss);
  else cphiv = (slatss - slatp * clambda) / (clatp * slambda);
  if (cphiv > 1.0) cphiv = 1.0;
  if (cphiv < -1.0) cphiv = -1.0;
  double phiv = acos(cphiv);
  double taneta = srho * slambda / (1.0 - srho * clambda);
  double eta = atan(taneta);
  double ceps = sin(eta) / srho;
  double eps = acos(ceps);
  double r2d = 180.0 / acos(-1.0);
  *a1 = r2d * phiv;
  if (from_space) { if (lonp - lonss < 0.0) *a1 = 360.0 - *a1; }
  else { if (lonp - lonss > 0.0) *a1 = 360.0 - *a1; }
  *a2 = from_space ? r2d * eta : r2d * eps;
  double lambda = r2d * acos(clambda);
  double lambda0 = r2d * acos(srho);
  if (lambda > lambda0) *a2 = -*a2;
}

/* mode 0: satAngles (view from target), 1: satAngles_fromSpace, 2: solarAngles / sunAngles.  latss / lonss NULL: scalars latss0 / lonss0 */
void orc_angles(int n, int mode, const int *yr, const int *mo, const int *day, const int *hr, const int *min, const double *sec,
                const double *latp, const double *lonp, const double *latss, const double *lonss, double latss0, double lonss0, double tz,
                double hgtss, double *vaa, double *vza, double *saa, double *sza) {
  for (int i = 0; i < n; i++) {
    if (mode != 2) {
      double a1, a2;
      orc_view_angles(latp[i], lonp[i], latss ? latss[i] : latss0, lonss ? lonss[i] : lonss0, hgtss, mode == 1, &a1, &a2);
      vaa[i] = a1;
      vza[i] = mode == 1 ? a2 : 90.0 - a2;
    }
    double az, el;
    orc_solar_angles(yr[i], mo[i], day[i], hr[i], min[i], sec[i], tz, latp[i], lonp[i], &az, &el);
    saa[i] = az;
    sza[i] = 90.0 - el;
  }
}
