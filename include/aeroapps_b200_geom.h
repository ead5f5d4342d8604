/* aeroapps_b200 -- geometry calculators feeding the hot path: per-pixel solar and viewing angles from
 * latitude / longitude / UTC time, the arrays the Stage-B entry points take as solar_zenith, sensor_zenith and relat_azymuth.
 * C ABI of the three f2py modules GeoAngles_ (Components/rtms/geo_vlidort/GeoAngles_py.F90), LidarAngles_
 * (rtms/leo_vlidort/LidarAngles_py.F90) and stnAngles_ (rtms/stn_vlidort/stnAngles_py.F90).  All arrays have n elements: integer*4
 * yr, mo, day, hr, min; real*8 sec, latp, lonp [degrees] and the outputs [degrees].  Host pointers; n = 0 is a no-op.
 * Return value: 0 or AB200_ERR_* (ab200_last_error()).  The reference prints a warning when the spacecraft is below the target's
 * horizon and returns the elevation negated; the warning is not printed here, the negation is kept. */
#ifndef AEROAPPS_B200_GEOM_H
#define AEROAPPS_B200_GEOM_H
#ifdef __cplusplus
extern "C" {
#endif

/* GeoAngles_.satangles (GeoAngles_py.F90:682-728): geostationary platform, tz = 0, satellite at latitude 0, height 35786 km
 * (module variables :5-7), sub-satellite longitude lonss.  view_zenith = 90 - elevation of the spacecraft at the target. */
int ab200_geo_satangles(int n, const int *yr, const int *mo, const int *day, const int *hr, const int *min, const double *sec,
                        const double *latp, const double *lonp, double lonss, double *view_azimuth, double *view_zenith,
                        double *solar_azimuth, double *solar_zenith);
/* LidarAngles_.satangles (LidarAngles_py.F90:901-952): moving platform, latss/lonss per point, time zone tz [h], height hgtss [km] */
int ab200_lidar_satangles(int n, const int *yr, const int *mo, const int *day, const int *hr, const int *min, const double *sec,
                          const double *latp, const double *lonp, const double *latss, const double *lonss, double tz, double hgtss,
                          double *view_azimuth, double *view_zenith, double *solar_azimuth, double *solar_zenith);
/* LidarAngles_.satangles_fromspace (:956-1007): azimuth of the target seen from the sub-satellite point and the nadir angle eta
 * at the spacecraft (view_angle), solar angles at the target */
int ab200_lidar_satangles_fromspace(int n, const int *yr, const int *mo, const int *day, const int *hr, const int *min,
                                    const double *sec, const double *latp, const double *lonp, const double *latss,
                                    const double *lonss, double tz, double hgtss, double *view_azimuth, double *view_angle,
                                    double *solar_azimuth, double *solar_zenith);
/* LidarAngles_.solarangles (:1011-1056) == stnAngles_.sunangles (stnAngles_py.F90:338-377): solar azimuth and zenith at the target */
int ab200_solarangles(int n, const int *yr, const int *mo, const int *day, const int *hr, const int *min, const double *sec,
                      const double *latp, const double *lonp, double tz, double *solar_azimuth, double *solar_zenith);
/* device-pointer form of ab200_lidar_satangles for pipelines that keep the geometry in HBM (stream = cudaStream_t);
 * from_space != 0 selects the satangles_fromspace outputs */
int ab200_satangles_dev(int n, const int *yr, const int *mo, const int *day, const int *hr, const int *min, const double *sec,
                        const double *latp, const double *lonp, const double *latss, const double *lonss, double tz, double hgtss,
                        int from_space, double *view_azimuth, double *view_zenith, double *solar_azimuth, double *solar_zenith,
                        void *stream);

#ifdef __cplusplus
}
#endif
#endif
