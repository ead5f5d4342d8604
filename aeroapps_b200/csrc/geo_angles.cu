// Geometry calculators feeding the hot path (include/aeroapps_b200_geom.h): solar and viewing angles per point.
// One thread per point; the point's five int32 date fields and three to five doubles are read once, four doubles written: 92 - 108 B
// of traffic against ~30 FP64 transcendentals, so the kernel is bound by the FP64 pipe, not by HBM.
// Built with -fmad=false: the formulas are ill-conditioned near the sub-satellite point (acos of a cosine near 1), where a fused
// multiply-add in the argument would move the last digits of the reference's answer.
#include <cuda_runtime.h>
#include <cmath>
#include <string>
#include "common.h"
#include "../../include/aeroapps_b200.h"

namespace {

// JULDAY (GeoAngles_py.F90:56-186, identical copies in LidarAngles_py.F90:51 and stnAngles_py.F90:51): Julian day number with the
// time of day, Gregorian switch at JUL >= 2299171, plus the relative epsilon the IDL original adds
__device__ double julday(int month, int day, int year, int hour, int minute, double second) {
  const int in_jan_feb = month <= 2 ? 1 : 0;
  const int jy = year - in_jan_feb, jm = month + 1 + 12 * in_jan_feb;
  int jul = (int)floor(365.25 * jy) + (int)floor(30.6001 * jm) + day + 1720995;
  if (jul >= 2299171) { const int ja = (int)floor(0.01 * jy); jul = jul + 2 - ja + (int)floor(0.25 * ja); }
  double eps = 2.2204460e-16;
  if (abs(jul) >= 1) eps = eps * abs(jul);
  return (double)jul + (double)hour / 24.0 - 0.5 + (double)minute / 1440.0 + second / 86400.0 + eps;
}

// solar_angles (GeoAngles_py.F90:305-457): NOAA low-accuracy solar position; az measured from north, el = elevation [degrees]
__device__ void solar_angles(int yr, int mo, int day, int hr, int mn, double sec, double tz, double lat, double lon, double &az, double &el) {
  const double d2r = acos(-1.0) / 180.0, r2d = 180.0 / acos(-1.0);
  const double jd = julday(mo, day, yr, hr, mn, sec);
  const double t = (jd - 2451545.0) / 36525.0 - tz / 24.0;
  double meanlon = 280.46646 + t * (36000.76983 + 0.0003032 * t);
  while (meanlon > 360.0) meanlon = meanlon - 360.0;
  while (meanlon < 0.0) meanlon = meanlon + 360.0;
  const double mean_anom = 357.52911 + t * (35999.05029 - 0.0001537 * t);
  const double mrad = d2r * mean_anom;
  double sinm = sin(mrad), sin2m = sin(2.0 * mrad);
  const double sin3m = sin(3.0 * mrad);
  const double c = sinm * (1.914602 - t * (0.004817 + 0.000014 * t)) + sin2m * (0.019993 - 0.000101 * t) + sin3m * 0.000289;
  const double truelon = meanlon + c;
  const double omega = 125.04 - 1934.136 * t;
  const double app_lon = truelon - 0.00569 - 0.00478 * sin(d2r * omega);
  const double e = 0.016708634 - t * (0.000042037 + 0.0000001267 * t);
  const double seconds = 21.448 - t * (46.8150 + t * (0.00059 - t * (0.001813)));
  const double eps0 = 23.0 + (26.0 + (seconds / 60.0)) / 60.0;
  const double epsilon = eps0 + 0.00256 * cos(d2r * omega);
  const double sint = sin(d2r * epsilon) * sin(d2r * app_lon);
  const double decl = r2d * asin(sint);
  const double m = mean_anom, lz = meanlon;
  const double th = tan(d2r * epsilon / 2.0), y = th * th;
  const double sin2lz = sin(2.0 * d2r * lz);
  sinm = sin(d2r * m);
  const double cos2lz = cos(2.0 * d2r * lz), sin4lz = sin(4.0 * d2r * lz);
  sin2m = sin(2.0 * d2r * m);
  double eqtime = y * sin2lz - 2.0 * e * sinm + 4.0 * e * y * sinm * cos2lz - 0.5 * y * y * sin4lz - 1.25 * e * e * sin2m;
  eqtime = r2d * eqtime * 4.0;
  const double utc = (double)hr + ((double)mn + sec / 60.0) / 60.0 - tz;
  const double ha = 180.0 - 15.0 * utc - lon - eqtime / 4.0;
  const double declr = d2r * decl, latr = d2r * lat, har = d2r * ha;
  const double elr = asin(sin(latr) * sin(declr) + cos(latr) * cos(declr) * cos(har));
  const double yaz = -sin(har) * cos(declr);
  const double xaz = cos(har) * cos(declr) * sin(latr) - sin(declr) * cos(latr);
  az = r2d * atan2(-yaz, -xaz);
  if (az < 0.0) az = az + 360.0;
  el = r2d * elr;
}

// view_angles (GeoAngles_py.F90:499-560) == view_angles_from_target (LidarAngles_py.F90:494-584): azimuth of the sub-satellite point
// seen from the target and elevation of the spacecraft there; view_angles_from_space (:620-717): azimuth of the target seen from the
// sub-satellite point and the nadir angle eta at the spacecraft.  Second output negated when the spacecraft is below the horizon.
__device__ void view_angles(double latp, double lonp, double latss, double lonss, double h, bool from_space, double &a1, double &a2) {
  const double d2r = acos(-1.0) / 180.0, r2d = 180.0 / acos(-1.0);
  const double re = 6378.0, srho = re / (re + h);
  const double deltaL = fabs(lonss - lonp);
  const double cdeltaL = cos(d2r * deltaL);
  const double clatss = cos(d2r * latss), slatss = sin(d2r * latss), clatp = cos(d2r * latp), slatp = sin(d2r * latp);
  const double clambda = slatss * slatp + clatss * clatp * cdeltaL;
  const double slambda = sin(acos(clambda));
  double cphiv = from_space ? (slatp - clambda * slatss) / (slambda * clatss) : (slatss - slatp * clambda) / (clatp * slambda);
  if (cphiv > 1.0) cphiv = 1.0;
  if (cphiv < -1.0) cphiv = -1.0;
  const double phiv = acos(cphiv);
  const double taneta = srho * slambda / (1.0 - srho * clambda);
  const double eta = atan(taneta);
  const double ceps = sin(eta) / srho;
  const double eps = acos(ceps);
  a1 = r2d * phiv;
  if (from_space) { if (lonp - lonss < 0.0) a1 = 360.0 - a1; } else { if (lonp - lonss > 0.0) a1 = 360.0 - a1; }
  a2 = from_space ? r2d * eta : r2d * eps;
  const double lambda = r2d * acos(clambda), lambda0 = r2d * acos(srho);
  if (lambda > lambda0) a2 = -a2;
}

// mode 0: view from target (view_zenith = 90 - elevation), 1: from space (nadir angle), 2: solar angles only.
// latss / lonss: per-point arrays, or nullptr with the scalars latss0 / lonss0 (geostationary platform)
__global__ void k_angles(int n, int mode, const int *__restrict__ yr, const int *__restrict__ mo, const int *__restrict__ day,
                         const int *__restrict__ hr, const int *__restrict__ mn, const double *__restrict__ sec,
                         const double *__restrict__ latp, const double *__restrict__ lonp, const double *__restrict__ latss,
                         const double *__restrict__ lonss, double latss0, double lonss0, double tz, double hgtss,
                         double *__restrict__ vaa, double *__restrict__ vza, double *__restrict__ saa, double *__restrict__ sza) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double la = latp[i], lo = lonp[i];
    if (mode != 2) {
      double a1, a2;
      view_angles(la, lo, latss ? latss[i] : latss0, lonss ? lonss[i] : lonss0, hgtss, mode == 1, a1, a2);
      vaa[i] = a1;
      vza[i] = mode == 1 ? a2 : 90.0 - a2;
    }
    double az, el;
    solar_angles(yr[i], mo[i], day[i], hr[i], mn[i], sec[i], tz, la, lo, az, el);
    saa[i] = az;
    sza[i] = 90.0 - el;
  }
}

int launch(int n, int mode, const int *yr, const int *mo, const int *day, const int *hr, const int *mn, const double *sec, const double *latp,
           const double *lonp, const double *latss, const double *lonss, double latss0, double lonss0, double tz, double hgtss, double *vaa,
           double *vza, double *saa, double *sza, cudaStream_t st) {
  const int threads = 128;
  int blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;                    // grid-stride beyond 16 resident CTAs per SM
  k_angles<<<blocks, threads, 0, st>>>(n, mode, yr, mo, day, hr, mn, sec, latp, lonp, latss, lonss, latss0, lonss0, tz, hgtss, vaa, vza, saa, sza);
  AB_CUDA_OK(cudaGetLastError());
  return 0;
}

// host-pointer front: one staging allocation, inputs up, kernel, outputs down
int host_call(int n, int mode, const int *yr, const int *mo, const int *day, const int *hr, const int *mn, const double *sec, const double *latp,
              const double *lonp, const double *latss, const double *lonss, double latss0, double lonss0, double tz, double hgtss, double *vaa,
              double *vza, double *saa, double *sza) {
  if (n < 0) { ab200_set_error("angles: n < 0"); return AB200_ERR_ARG; }
  if (n == 0) return 0;
  if (!yr || !mo || !day || !hr || !mn || !sec || !latp || !lonp || !saa || !sza || (mode != 2 && (!vaa || !vza))) {
    ab200_set_error("angles: NULL array argument"); return AB200_ERR_ARG;
  }
  const size_t N = (size_t)n, nd_in = 3 + (latss ? 1 : 0) + (lonss ? 1 : 0), nd_out = mode == 2 ? 2 : 4;
  char *d = nullptr;
  AB_CUDA_OK(cudaMalloc(&d, N * (8 * (nd_in + nd_out) + 4 * 5) + 256));
  double *dd = reinterpret_cast<double *>(d);
  double *d_sec = dd, *d_lat = dd + N, *d_lon = dd + 2 * N, *p = dd + 3 * N;
  double *d_latss = nullptr, *d_lonss = nullptr;
  if (latss) { d_latss = p; p += N; }
  if (lonss) { d_lonss = p; p += N; }
  double *d_out = p; p += nd_out * N;
  int *di = reinterpret_cast<int *>(p);
  cudaStream_t st = 0;
  cudaError_t e = cudaSuccess;
  auto up = [&](void *dst, const void *src, size_t bytes) { if (e == cudaSuccess) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st); };
  up(d_sec, sec, 8 * N); up(d_lat, latp, 8 * N); up(d_lon, lonp, 8 * N);
  if (latss) up(d_latss, latss, 8 * N);
  if (lonss) up(d_lonss, lonss, 8 * N);
  up(di, yr, 4 * N); up(di + N, mo, 4 * N); up(di + 2 * N, day, 4 * N); up(di + 3 * N, hr, 4 * N); up(di + 4 * N, mn, 4 * N);
  int rc = 0;
  if (e == cudaSuccess) {
    double *o_saa = mode == 2 ? d_out : d_out + 2 * N, *o_sza = o_saa + N;
    rc = launch(n, mode, di, di + N, di + 2 * N, di + 3 * N, di + 4 * N, d_sec, d_lat, d_lon, d_latss, d_lonss, latss0, lonss0, tz, hgtss,
                d_out, d_out + N, o_saa, o_sza, st);
    auto down = [&](void *dst, const void *src) { if (e == cudaSuccess) e = cudaMemcpyAsync(dst, src, 8 * N, cudaMemcpyDeviceToHost, st); };
    if (rc == 0) {
      if (mode != 2) { down(vaa, d_out); down(vza, d_out + N); }
      down(saa, o_saa); down(sza, o_sza);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  cudaFree(d);
  if (e != cudaSuccess) { ab200_set_error(std::string("angles: ") + cudaGetErrorString(e)); return AB200_ERR_CUDA; }
  return rc;
}

}  // namespace

extern "C" int ab200_geo_satangles(int n, const int *yr, const int *mo, const int *day, const int *hr, const int *min, const double *sec,
                                   const double *latp, const double *lonp, double lonss, double *view_azimuth, double *view_zenith,
                                   double *solar_azimuth, double *solar_zenith) {
  return host_call(n, 0, yr, mo, day, hr, min, sec, latp, lonp, nullptr, nullptr, 0.0, lonss, 0.0, 35786.0, view_azimuth, view_zenith,
                   solar_azimuth, solar_zenith);
}
extern "C" int ab200_lidar_satangles(int n, const int *yr, const int *mo, const int *day, const int *hr, const int *min, const double *sec,
                                     const double *latp, const double *lonp, const double *latss, const double *lonss, double tz, double hgtss,
                                     double *view_azimuth, double *view_zenith, double *solar_azimuth, double *solar_zenith) {
  if (n > 0 && (!latss || !lonss)) { ab200_set_error("satangles: NULL latss / lonss"); return AB200_ERR_ARG; }
  return host_call(n, 0, yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, 0.0, 0.0, tz, hgtss, view_azimuth, view_zenith, solar_azimuth,
                   solar_zenith);
}
extern "C" int ab200_lidar_satangles_fromspace(int n, const int *yr, const int *mo, const int *day, const int *hr, const int *min,
                                               const double *sec, const double *latp, const double *lonp, const double *latss,
                                               const double *lonss, double tz, double hgtss, double *view_azimuth, double *view_angle,
                                               double *solar_azimuth, double *solar_zenith) {
  if (n > 0 && (!latss || !lonss)) { ab200_set_error("satangles_fromspace: NULL latss / lonss"); return AB200_ERR_ARG; }
  return host_call(n, 1, yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, 0.0, 0.0, tz, hgtss, view_azimuth, view_angle, solar_azimuth,
                   solar_zenith);
}
extern "C" int ab200_solarangles(int n, const int *yr, const int *mo, const int *day, const int *hr, const int *min, const double *sec,
                                 const double *latp, const double *lonp, double tz, double *solar_azimuth, double *solar_zenith) {
  return host_call(n, 2, yr, mo, day, hr, min, sec, latp, lonp, nullptr, nullptr, 0.0, 0.0, tz, 0.0, nullptr, nullptr, solar_azimuth, solar_zenith);
}
extern "C" int ab200_satangles_dev(int n, const int *yr, const int *mo, const int *day, const int *hr, const int *min, const double *sec,
                                   const double *latp, const double *lonp, const double *latss, const double *lonss, double tz, double hgtss,
                                   int from_space, double *view_azimuth, double *view_zenith, double *solar_azimuth, double *solar_zenith,
                                   void *stream) {
  if (n < 0) { ab200_set_error("angles: n < 0"); return AB200_ERR_ARG; }
  if (n == 0) return 0;
  return launch(n, from_space ? 1 : 0, yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, 0.0, 0.0, tz, hgtss, view_azimuth, view_zenith,
                solar_azimuth, solar_zenith, (cudaStream_t)stream);
}
