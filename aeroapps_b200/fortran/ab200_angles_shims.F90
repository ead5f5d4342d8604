! ISO_C_BINDING shims for the geometry calculators: the reference's external subroutines (the bodies the f2py modules GeoAngles_,
! LidarAngles_ and stnAngles_ wrap), re-created on top of libaeroapps_b200.so.  Each of the three reference files defines its own
! `satAngles`; build the shim that belongs to the f2py module being replaced (preprocessor symbol AB200_GEO / AB200_LIDAR /
!  AB200_STN).
!   satAngles (geostationary)            Components/rtms/geo_vlidort/GeoAngles_py.F90:682-728
!   satAngles, satAngles_fromSpace, solarAngles   Components/rtms/leo_vlidort/LidarAngles_py.F90:901-1056
!   sunAngles                            Components/rtms/stn_vlidort/stnAngles_py.F90:338-377
! NOT compiled in this repository's CI (no Fortran compiler in the build image or on the GPU box).
module ab200_angles_c_api
   use iso_c_binding
   implicit none
   interface
      integer(c_int) function ab200_geo_satangles(n, yr, mo, day, hr, min, sec, latp, lonp, lonss, vaa, vza, saa, sza) bind(C, &
            name="ab200_geo_satangles")
         import :: c_int, c_double
         integer(c_int), value :: n
         integer(c_int), intent(in) :: yr(*), mo(*), day(*), hr(*), min(*)
         real(c_double), intent(in) :: sec(*), latp(*), lonp(*)
         real(c_double), value :: lonss
         real(c_double), intent(out) :: vaa(*), vza(*), saa(*), sza(*)
      end function
      integer(c_int) function ab200_lidar_satangles(n, yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, tz, hgtss, vaa, vza, &
            saa, sza) &
                                                    bind(C, name="ab200_lidar_satangles")
         import :: c_int, c_double
         integer(c_int), value :: n
         integer(c_int), intent(in) :: yr(*), mo(*), day(*), hr(*), min(*)
         real(c_double), intent(in) :: sec(*), latp(*), lonp(*), latss(*), lonss(*)
         real(c_double), value :: tz, hgtss
         real(c_double), intent(out) :: vaa(*), vza(*), saa(*), sza(*)
      end function
      integer(c_int) function ab200_lidar_satangles_fromspace(n, yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, tz, hgtss, &
            vaa, va, saa, sza) &
                                                              bind(C, name="ab200_lidar_satangles_fromspace")
         import :: c_int, c_double
         integer(c_int), value :: n
         integer(c_int), intent(in) :: yr(*), mo(*), day(*), hr(*), min(*)
         real(c_double), intent(in) :: sec(*), latp(*), lonp(*), latss(*), lonss(*)
         real(c_double), value :: tz, hgtss
         real(c_double), intent(out) :: vaa(*), va(*), saa(*), sza(*)
      end function
      integer(c_int) function ab200_solarangles(n, yr, mo, day, hr, min, sec, latp, lonp, tz, saa, sza) bind(C, &
            name="ab200_solarangles")
         import :: c_int, c_double
         integer(c_int), value :: n
         integer(c_int), intent(in) :: yr(*), mo(*), day(*), hr(*), min(*)
         real(c_double), intent(in) :: sec(*), latp(*), lonp(*)
         real(c_double), value :: tz
         real(c_double), intent(out) :: saa(*), sza(*)
      end function
   end interface
end module ab200_angles_c_api

#ifdef AB200_GEO
subroutine satAngles(n, yr, mo, day, hr, min, sec, latp, lonp, lonss, view_azimuth, view_zenith, solar_azimuth, solar_zenith)
   use ab200_angles_c_api
   implicit none
   integer(kind=4), intent(in) :: n, yr(n), mo(n), day(n), hr(n), min(n)
   real(kind=8),    intent(in) :: sec(n), lonp(n), latp(n), lonss
   real(kind=8), intent(out) :: view_azimuth(n), view_zenith(n), solar_azimuth(n), solar_zenith(n)
   integer :: st
   st = ab200_geo_satangles(n, yr, mo, day, hr, min, sec, latp, lonp, lonss, view_azimuth, view_zenith, solar_azimuth, solar_zenith)
end subroutine satAngles
#endif

#ifdef AB200_LIDAR
subroutine satAngles(n, yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, tz, hgtss, view_azimuth, view_zenith, &
      solar_azimuth, solar_zenith)
   use ab200_angles_c_api
   implicit none
   integer(kind=4), intent(in) :: n, yr(n), mo(n), day(n), hr(n), min(n)
   real(kind=8),    intent(in) :: sec(n), lonp(n), latp(n), lonss(n), latss(n), tz, hgtss
   real(kind=8), intent(out) :: view_azimuth(n), view_zenith(n), solar_azimuth(n), solar_zenith(n)
   integer :: st
   st = ab200_lidar_satangles(n, yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, tz, hgtss, view_azimuth, view_zenith, &
         solar_azimuth, solar_zenith)
end subroutine satAngles
subroutine satAngles_fromSpace(n, yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, tz, hgtss, view_azimuth, view_angle, &
      solar_azimuth, solar_zenith)
   use ab200_angles_c_api
   implicit none
   integer(kind=4), intent(in) :: n, yr(n), mo(n), day(n), hr(n), min(n)
   real(kind=8),    intent(in) :: sec(n), lonp(n), latp(n), lonss(n), latss(n), tz, hgtss
   real(kind=8), intent(out) :: view_azimuth(n), view_angle(n), solar_azimuth(n), solar_zenith(n)
   integer :: st
   st = ab200_lidar_satangles_fromspace(n, yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, tz, hgtss, view_azimuth, &
         view_angle, solar_azimuth, &
                                        solar_zenith)
end subroutine satAngles_fromSpace
subroutine solarAngles(n, yr, mo, day, hr, min, sec, latp, lonp, tz, solar_azimuth, solar_zenith)
   use ab200_angles_c_api
   implicit none
   integer(kind=4), intent(in) :: n, yr(n), mo(n), day(n), hr(n), min(n)
   real(kind=8),    intent(in) :: sec(n), lonp(n), latp(n), tz
   real(kind=8), intent(out) :: solar_azimuth(n), solar_zenith(n)
   integer :: st
   st = ab200_solarangles(n, yr, mo, day, hr, min, sec, latp, lonp, tz, solar_azimuth, solar_zenith)
end subroutine solarAngles
#endif

#ifdef AB200_STN
subroutine sunAngles(n, yr, mo, day, hr, min, sec, latp, lonp, tz, solar_azimuth, solar_zenith)
   use ab200_angles_c_api
   implicit none
   integer(kind=4), intent(in) :: n, yr(n), mo(n), day(n), hr(n), min(n)
   real(kind=8),    intent(in) :: sec(n), lonp(n), latp(n), tz
   real(kind=8), intent(out) :: solar_azimuth(n), solar_zenith(n)
   integer :: st
   st = ab200_solarangles(n, yr, mo, day, hr, min, sec, latp, lonp, tz, solar_azimuth, solar_zenith)
end subroutine sunAngles
#endif
