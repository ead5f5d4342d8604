! ISO_C_BINDING shims: the reference's module procedures, re-created on top of libaeroapps_b200.so.
! Drop these modules into the AeroApps build in place of the bodies of
!   Components/rtms/geo_vlidort/VLIDORT_ROT.F90, VLIDORT_LAMBERT.F90, VLIDORT_BRDF_CX.F90, VLIDORT_BRDF_MODIS.F90
! (module and procedure names, argument lists and intents are unchanged, so geo_vlidort.F90, leo_vlidort_cloud.F90 and
! the f2py pass-throughs of leo_vlidort/VLIDORT_POLAR_py.F90 compile and link untouched), and link with
!   -L<repo>/aeroapps_b200 -laeroapps_b200 -lcudart.
! NOT compiled in this repository's CI: the build image has no Fortran compiler (see DESIGN.md section 2).
module ab200_c_api
   use iso_c_binding
   implicit none
   interface
      integer(c_int) function ab200_rot_calc(km, nch, nobs, channels, pe, he, te, MISSING, verbose, ROT, depol, rc) bind(C, name="ab200_rot_calc")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), pe(*), he(*), te(*)
         real(c_double), intent(out) :: ROT(*), depol(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_scalar_lambert(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, g, &
                                                           pmom, pe, he, te, albedo, sza, raa, vza, MISSING, verbose, rad, refl, rc) &
                                                           bind(C, name="ab200_vlidort_scalar_lambert")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), g(*), pmom(*), pe(*), he(*), te(*), albedo(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_lambert(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, alpha, tau, ssa, &
                                                           pmom, pe, he, te, albedo, sza, raa, vza, flux_factor, MISSING, verbose, rad, refl, Q, U, rc) &
                                                           bind(C, name="ab200_vlidort_vector_lambert")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), albedo(*), sza(*), raa(*), vza(*), flux_factor(*)
         real(c_double), intent(out) :: rad(*), refl(*), Q(*), U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_cx(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, &
                                                      U10m, V10m, mr, sza, raa, vza, MISSING, verbose, rad, refl, BR, Q, U, BR_Q, BR_U, rc) &
                                                      bind(C, name="ab200_vlidort_vector_cx")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_gisscx(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, &
                                                      U10m, V10m, mr, sza, raa, vza, MISSING, verbose, rad, refl, BR, Q, U, BR_Q, BR_U, rc) &
                                                      bind(C, name="ab200_vlidort_vector_gisscx")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_ocicx(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, &
                                                      U10m, V10m, mr, sza, raa, vza, MISSING, verbose, rad, refl, BR, Q, U, BR_Q, BR_U, rc) &
                                                      bind(C, name="ab200_vlidort_vector_ocicx")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_ocigisscx(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, alpha, tau, ssa, pmom, &
                                                             pe, he, te, U10m, V10m, mr, sza, raa, vza, flux_factor, MISSING, verbose, rad, refl, BR, Q, U, &
                                                             BR_Q, BR_U, rc) bind(C, name="ab200_vlidort_vector_ocigisscx")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*)
         real(c_double), intent(in)  :: sza(*), raa(*), vza(*), flux_factor(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_brdf_modis(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, nkernel, nparam, ROT, depol, &
                                                              tau, ssa, pmom, pe, he, te, kernel_wt, param, sza, raa, vza, MISSING, verbose, rad, refl, &
                                                              BR, Q, U, BR_Q, BR_U, rc) bind(C, name="ab200_vlidort_vector_brdf_modis")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, nkernel, nparam, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), kernel_wt(*), param(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
   end interface
end module ab200_c_api

module VLIDORT_ROT                      ! replaces geo_vlidort/VLIDORT_ROT.F90:19-86
   use ab200_c_api
   implicit none
   public :: VLIDORT_ROT_CALC
contains
   subroutine VLIDORT_ROT_CALC(km, nch, nobs, channels, pe, he, te, MISSING, verbose, ROT, depol_ratio, rc)
      integer, intent(in)  :: km, nch, nobs, verbose
      real*8,  intent(in)  :: channels(nch), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), MISSING
      real*8,  intent(out) :: ROT(km,nobs,nch), depol_ratio(nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_rot_calc(km, nch, nobs, channels, pe, he, te, MISSING, verbose, ROT, depol_ratio, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_ROT_CALC
end module VLIDORT_ROT

module VLIDORT_LAMBERT                  ! replaces geo_vlidort/VLIDORT_LAMBERT.F90:34-342 (SingleGeom forms)
   use ab200_c_api
   implicit none
   public :: VLIDORT_Scalar_Lambert, VLIDORT_Vector_Lambert
contains
   subroutine VLIDORT_Scalar_Lambert(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, g, pmom, pe, he, te, albedo, &
                                     solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL, reflectance_VL, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs)
      real*8,  intent(in)  :: pmom(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), albedo(nobs,nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL(nobs,nch), reflectance_VL(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_scalar_lambert(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, tau, ssa, g, pmom, &
                                         pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL, reflectance_VL, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_Lambert

   subroutine VLIDORT_Vector_Lambert(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, alpha, tau, ssa, pmom, pe, he, te, albedo, &
                                     solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, radiance_VL, reflectance_VL, Q, U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs)
      real*8,  intent(in)  :: pmom(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), albedo(nobs,nch), flux_factor(nch,nobs)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL(nobs,nch), reflectance_VL(nobs,nch), Q(nobs,nch), U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_lambert(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, alpha, tau, ssa, pmom, &
                                         pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, &
                                         radiance_VL, reflectance_VL, Q, U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_Lambert
end module VLIDORT_LAMBERT

module VLIDORT_BRDF_CX                  ! replaces geo_vlidort/VLIDORT_BRDF_CX.F90:587-762 (note the output order Q,U,BRDF,...)
   use ab200_c_api
   implicit none
   public :: VLIDORT_Vector_CX, VLIDORT_Vector_GissCX, VLIDORT_Vector_OCICX, VLIDORT_Vector_OCIGissCX
contains
   subroutine VLIDORT_Vector_CX(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, U10m, V10m, mr, &
                                solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, &
                                Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BRDF(nobs,nch), BRDF_Q(nobs,nch), BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_cx(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, &
                                    U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                    radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_CX
   subroutine VLIDORT_Vector_GissCX(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, U10m, V10m, mr, &
                                solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, &
                                Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BRDF(nobs,nch), BRDF_Q(nobs,nch), BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_gisscx(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, &
                                    U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                    radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_GissCX
   subroutine VLIDORT_Vector_OCICX(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, U10m, V10m, mr, &
                                solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, &
                                Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BRDF(nobs,nch), BRDF_Q(nobs,nch), BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_ocicx(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, &
                                    U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                    radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_OCICX
   subroutine VLIDORT_Vector_OCIGissCX(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, alpha, tau, ssa, pmom, pe, he, te, &
                                       U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, &
                                       radiance_VL_SURF, reflectance_VL_SURF, Q, U, BRDF, BRDF_Q, BRDF_U, rc)    ! VLIDORT_BRDF_CX.F90:54-236
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs)
      real*8,  intent(in)  :: pmom(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BRDF(nobs,nch), BRDF_Q(nobs,nch), BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_ocigisscx(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, alpha, tau, ssa, pmom, &
                                           pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, &
                                           radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_OCIGissCX
end module VLIDORT_BRDF_CX

module VLIDORT_BRDF_MODIS               ! replaces geo_vlidort/VLIDORT_BRDF_MODIS.F90:206-410
   use ab200_c_api
   implicit none
   public :: VLIDORT_Vector_LandMODIS
contains
   subroutine VLIDORT_Vector_LandMODIS(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, &
                                       kernel_wt, param, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                       radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), kernel_wt(3,nch,nobs), param(2,nch,nobs)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BR(nobs,nch), Q(nobs,nch), U(nobs,nch), BR_Q(nobs,nch), BR_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_brdf_modis(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, 3, 2, ROT, depol, tau, ssa, pmom, &
                                            pe, he, te, kernel_wt, param, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                            radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_LandMODIS
end module VLIDORT_BRDF_MODIS
