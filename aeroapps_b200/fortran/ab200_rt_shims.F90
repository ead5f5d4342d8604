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
      integer(c_int) function ab200_rot_calc(km, nch, nobs, channels, pe, he, te, MISSING, verbose, ROT, depol, rc) bind(C, &
            name="ab200_rot_calc")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), pe(*), he(*), te(*)
         real(c_double), intent(out) :: ROT(*), depol(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_scalar_lambert(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
            depol, tau, ssa, g, &
                                                           pmom, pe, he, te, albedo, sza, raa, vza, MISSING, verbose, rad, refl, &
                                                                 rc) &
                                                           bind(C, name="ab200_vlidort_scalar_lambert")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), g(*), pmom(*), pe(*), he(*), te(*), &
               albedo(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_lambert(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
            depol, alpha, tau, ssa, &
                                                           pmom, pe, he, te, albedo, sza, raa, vza, flux_factor, MISSING, &
                                                                 verbose, rad, refl, Q, U, rc) &
                                                           bind(C, name="ab200_vlidort_vector_lambert")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), &
               albedo(*), sza(*), raa(*), vza(*), flux_factor(*)
         real(c_double), intent(out) :: rad(*), refl(*), Q(*), U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_cx(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, &
            tau, ssa, pmom, pe, he, te, &
                                                      U10m, V10m, mr, sza, raa, vza, MISSING, verbose, rad, refl, BR, Q, U, &
                                                            BR_Q, BR_U, rc) &
                                                      bind(C, name="ab200_vlidort_vector_cx")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), U10m(*), &
               V10m(*), mr(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_gisscx(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
            depol, tau, ssa, pmom, pe, he, te, &
                                                      U10m, V10m, mr, sza, raa, vza, MISSING, verbose, rad, refl, BR, Q, U, &
                                                            BR_Q, BR_U, rc) &
                                                      bind(C, name="ab200_vlidort_vector_gisscx")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), U10m(*), &
               V10m(*), mr(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_ocicx(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
            depol, tau, ssa, pmom, pe, he, te, &
                                                      U10m, V10m, mr, sza, raa, vza, MISSING, verbose, rad, refl, BR, Q, U, &
                                                            BR_Q, BR_U, rc) &
                                                      bind(C, name="ab200_vlidort_vector_ocicx")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), U10m(*), &
               V10m(*), mr(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_ocigisscx(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
            depol, alpha, tau, ssa, pmom, &
                                                             pe, he, te, U10m, V10m, mr, sza, raa, vza, flux_factor, MISSING, &
                                                                   verbose, rad, refl, BR, Q, U, &
                                                             BR_Q, BR_U, rc) bind(C, name="ab200_vlidort_vector_ocigisscx")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), &
               U10m(*), V10m(*), mr(*)
         real(c_double), intent(in)  :: sza(*), raa(*), vza(*), flux_factor(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_brdf_modis(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, &
            nkernel, nparam, ROT, depol, &
                                                              tau, ssa, pmom, pe, he, te, kernel_wt, param, sza, raa, vza, &
                                                                    MISSING, verbose, rad, refl, &
                                                              BR, Q, U, BR_Q, BR_U, rc) bind(C, &
                                                                    name="ab200_vlidort_vector_brdf_modis")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, nkernel, nparam, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), &
               kernel_wt(*), param(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_scalar_cx(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, &
            tau, ssa, g, pmom, pe, he, te, &
                                                      U10m, V10m, mr, sza, raa, vza, MISSING, verbose, rad, refl, BRDF, &
                                                            rc) bind(C, name="ab200_vlidort_scalar_cx")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), g(*), pmom(*), pe(*), he(*), te(*), &
               U10m(*), V10m(*), mr(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BRDF(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_scalar_gisscx(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
            depol, tau, ssa, g, pmom, pe, he, te, &
                                                          U10m, V10m, mr, sza, raa, vza, MISSING, verbose, rad, refl, BRDF, &
                                                                rc) bind(C, name="ab200_vlidort_scalar_gisscx")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), g(*), pmom(*), pe(*), he(*), te(*), &
               U10m(*), V10m(*), mr(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BRDF(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_scalar_brdf_modis(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, &
            ROT, depol, tau, ssa, g, pmom, &
                                                              pe, he, te, kernel_wt, param, sza, raa, vza, MISSING, verbose, &
                                                                    rad, refl, BR, rc) &
                                                              bind(C, name="ab200_vlidort_scalar_brdf_modis")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), g(*), pmom(*), pe(*), he(*), te(*), &
               kernel_wt(*), param(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*)
         integer(c_int), intent(out) :: rc
      end function
      ! multi-geometry forms: angles (nobs,ngeom), outputs (nobs,nch,ngeom)
      integer(c_int) function ab200_vlidort_vector_lambert_multigeom(km, nch, nobs, ngeom, channels, nstreams, plane_parallel, &
            nMom, nPol, ROT, depol, alpha, &
                                                                     tau, ssa, pmom, pe, he, te, albedo, sza, raa, vza, &
                                                                           flux_factor, MISSING, verbose, &
                                                                     rad, refl, Q, U, rc) bind(C, &
                                                                           name="ab200_vlidort_vector_lambert_multigeom")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, ngeom, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), &
               albedo(*), sza(*), raa(*), vza(*), flux_factor(*)
         real(c_double), intent(out) :: rad(*), refl(*), Q(*), U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_cx_multigeom(km, nch, nobs, ngeom, channels, nstreams, plane_parallel, nMom, &
            nPol, ROT, depol, tau, ssa, pmom, &
                                                                pe, he, te, U10m, V10m, mr, sza, raa, vza, MISSING, verbose, &
                                                                      rad, refl, BR, Q, U, BR_Q, BR_U, rc) &
                                                                bind(C, name="ab200_vlidort_vector_cx_multigeom")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, ngeom, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), U10m(*), &
               V10m(*), mr(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_brdf_modis_multigeom(km, nch, nobs, ngeom, channels, nstreams, &
            plane_parallel, nMom, nPol, nkernel, nparam, &
                                                                        ROT, depol, tau, ssa, pmom, pe, he, te, kernel_wt, &
                                                                              param, sza, raa, vza, MISSING, &
                                                                        verbose, rad, refl, BR, Q, U, BR_Q, BR_U, rc) &
                                                                        bind(C, name="ab200_vlidort_vector_brdf_modis_multigeom")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, ngeom, nstreams, plane_parallel, nMom, nPol, nkernel, nparam, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), &
               kernel_wt(*), param(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_gisscx_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, &
            ROT, depol, tau, ssa, pmom, &
                                                                tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, &
                                                                      mr, sza, raa, vza, &
                                                                MISSING, verbose, rad, refl, BR, Q, U, BR_Q, BR_U, rc) &
                                                                bind(C, name="ab200_vlidort_vector_gisscx_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), pmomI(*), &
               tauL(*), ssaL(*), pmomL(*)
         real(c_double), intent(in)  :: pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_brdf_modis_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, &
            nPol, nkernel, nparam, ROT, depol, &
                                                                    alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, &
                                                                          pe, he, te, kernel_wt, param, &
                                                                    sza, raa, vza, flux_factor, MISSING, verbose, rad, refl, BR, &
                                                                          Q, U, BR_Q, BR_U, rc) &
                                                                    bind(C, name="ab200_vlidort_vector_brdf_modis_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, nkernel, nparam, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), &
               pmomI(*), tauL(*), ssaL(*), pmomL(*)
         real(c_double), intent(in)  :: pe(*), he(*), te(*), kernel_wt(*), param(*), sza(*), raa(*), vza(*), flux_factor(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_bpdf(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, nparam, &
            ROT, depol, tau, ssa, pmom, &
                                                        pe, he, te, BPDFparam, sza, raa, vza, MISSING, verbose, rad, refl, BR, &
                                                              Q, U, BR_Q, BR_U, rc) &
                                                        bind(C, name="ab200_vlidort_vector_bpdf")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, nparam, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), &
               BPDFparam(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_brdf_modis_bpdf(km, nch, nobs, channels, nstreams, plane_parallel, nMom, &
            nPol, nkernel, nparam, nparam_bpdf, &
                                                                   ROT, depol, tau, ssa, pmom, pe, he, te, kernel_wt, RTLSparam, &
                                                                         BPDFparam, sza, raa, vza, &
                                                                   MISSING, verbose, rad, refl, BR, Q, U, BR_Q, BR_U, rc) &
                                                                   bind(C, name="ab200_vlidort_vector_brdf_modis_bpdf")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, nkernel, nparam, nparam_bpdf, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), &
               kernel_wt(*), RTLSparam(*), BPDFparam(*)
         real(c_double), intent(in)  :: sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_lambert_bpdf(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, &
            nparam, ROT, depol, tau, ssa, &
                                                                pmom, pe, he, te, albedo, BPDFparam, sza, raa, vza, MISSING, &
                                                                      verbose, rad, refl, BR, Q, U, &
                                                                BR_Q, BR_U, rc) bind(C, name="ab200_vlidort_vector_lambert_bpdf")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, nparam, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), albedo(*), &
               BPDFparam(*), sza(*), raa(*), vza(*)
         real(c_double), intent(out) :: rad(*), refl(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      ! ---- cloud forms added last (entries of include/aeroapps_b200_rt.h, same order of arguments)
      integer(c_int) function ab200_vlidort_scalar_lambert_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, &
            ROT, depol, alpha, &
            tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, albedo, solar_zenith, relat_azymuth, &
                  sensor_zenith, &
            flux_factor, MISSING, verbose, radiance_VL, reflectance_VL, rc) &
            bind(C, name="ab200_vlidort_scalar_lambert_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), g(*), pmom(*), tauI(*), &
               ssaI(*), gI(*), pmomI(*), &
            tauL(*), ssaL(*), gL(*), pmomL(*), pe(*), he(*), te(*), albedo(*), solar_zenith(*), relat_azymuth(*), &
                  sensor_zenith(*), flux_factor(*)
         real(c_double), intent(out) :: radiance_VL(*), reflectance_VL(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_lambert_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, &
            ROT, depol, alpha, &
            tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, albedo, solar_zenith, relat_azymuth, &
                  sensor_zenith, flux_factor, &
            MISSING, verbose, radiance_VL, reflectance_VL, Q, U, rc) &
            bind(C, name="ab200_vlidort_vector_lambert_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), &
               pmomI(*), tauL(*), &
            ssaL(*), pmomL(*), pe(*), he(*), te(*), albedo(*), solar_zenith(*), relat_azymuth(*), sensor_zenith(*), flux_factor(*)
         real(c_double), intent(out) :: radiance_VL(*), reflectance_VL(*), Q(*), U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_lambert_cloud_multigeom(km, nch, nobs, ngeom, channels, nstreams, &
            plane_parallel, nMom, nPol, ROT, &
            depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, albedo, solar_zenith, relat_azymuth, &
                  sensor_zenith, &
            MISSING, verbose, radiance_VL, reflectance_VL, Q, U, rc) &
            bind(C, name="ab200_vlidort_vector_lambert_cloud_multigeom")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, ngeom, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), &
               pmomI(*), tauL(*), &
            ssaL(*), pmomL(*), pe(*), he(*), te(*), albedo(*), solar_zenith(*), relat_azymuth(*), sensor_zenith(*)
         real(c_double), intent(out) :: radiance_VL(*), reflectance_VL(*), Q(*), U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_cx_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
            depol, tau, ssa, &
            pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, &
                  MISSING, verbose, &
            radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc) &
            bind(C, name="ab200_vlidort_vector_cx_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), pmomI(*), &
               tauL(*), ssaL(*), &
            pmomL(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), solar_zenith(*), relat_azymuth(*), sensor_zenith(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BRDF(*), Q(*), U(*), BRDF_Q(*), BRDF_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_cx_cloud_multigeom(km, nch, nobs, ngeom, channels, nstreams, plane_parallel, &
            nMom, nPol, ROT, &
            depol, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, &
                  relat_azymuth, sensor_zenith, &
            MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc) &
            bind(C, name="ab200_vlidort_vector_cx_cloud_multigeom")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, ngeom, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), pmomI(*), &
               tauL(*), ssaL(*), &
            pmomL(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), solar_zenith(*), relat_azymuth(*), sensor_zenith(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BRDF(*), Q(*), U(*), BRDF_Q(*), BRDF_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_ocigisscx_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, &
            nPol, ROT, depol, alpha, &
            tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, &
                  sensor_zenith, &
            flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc) &
            bind(C, name="ab200_vlidort_vector_ocigisscx_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), &
               pmomI(*), tauL(*), &
            ssaL(*), pmomL(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), solar_zenith(*), relat_azymuth(*), &
                  sensor_zenith(*), flux_factor(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BRDF(*), Q(*), U(*), BRDF_Q(*), BRDF_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_scalar_ocigisscx_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, &
            nPol, ROT, depol, alpha, &
            tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, &
                  relat_azymuth, sensor_zenith, &
            flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc) &
            bind(C, name="ab200_vlidort_scalar_ocigisscx_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), g(*), pmom(*), tauI(*), &
               ssaI(*), gI(*), pmomI(*), &
            tauL(*), ssaL(*), gL(*), pmomL(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), solar_zenith(*), relat_azymuth(*), &
                  sensor_zenith(*), &
            flux_factor(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BRDF(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_scalar_gisscx_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, &
            ROT, depol, tau, ssa, &
            g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, &
                  sensor_zenith, MISSING, &
            verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc) &
            bind(C, name="ab200_vlidort_scalar_gisscx_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), g(*), pmom(*), tauI(*), ssaI(*), gI(*), &
               pmomI(*), tauL(*), &
            ssaL(*), gL(*), pmomL(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), solar_zenith(*), relat_azymuth(*), &
                  sensor_zenith(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BRDF(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_brdf_modis_cloud_multigeom(km, nch, nobs, ngeom, channels, nstreams, &
            plane_parallel, nMom, nPol, &
            n0, n1, ROT, depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, kernel_wt, param, &
                  solar_zenith, &
            relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, &
                  BR_U, rc) &
            bind(C, name="ab200_vlidort_vector_brdf_modis_cloud_multigeom")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, ngeom, nstreams, plane_parallel, nMom, nPol, n0, n1, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), &
               pmomI(*), tauL(*), &
            ssaL(*), pmomL(*), pe(*), he(*), te(*), kernel_wt(*), param(*), solar_zenith(*), relat_azymuth(*), sensor_zenith(*), &
                  flux_factor(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_scalar_brdf_modis_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, &
            nPol, ROT, depol, alpha, &
            tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, kernel_wt, param, solar_zenith, &
                  relat_azymuth, &
            sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BR, rc) &
            bind(C, name="ab200_vlidort_scalar_brdf_modis_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), g(*), pmom(*), tauI(*), &
               ssaI(*), gI(*), pmomI(*), &
            tauL(*), ssaL(*), gL(*), pmomL(*), pe(*), he(*), te(*), kernel_wt(*), param(*), solar_zenith(*), relat_azymuth(*), &
                  sensor_zenith(*), &
            flux_factor(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BR(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_brdf_modis_bpdf_cloud(km, nch, nobs, channels, nstreams, plane_parallel, &
            nMom, nPol, n0, n1, n2, &
            ROT, depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, kernel_wt, RTLSparam, &
                  BPDFparam, solar_zenith, &
            relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, &
                  BR_U, rc) &
            bind(C, name="ab200_vlidort_vector_brdf_modis_bpdf_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, n0, n1, n2, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), &
               pmomI(*), tauL(*), &
            ssaL(*), pmomL(*), pe(*), he(*), te(*), kernel_wt(*), RTLSparam(*), BPDFparam(*), solar_zenith(*), relat_azymuth(*), &
                  sensor_zenith(*), &
            flux_factor(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_lambert_bpdf_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, &
            nPol, n0, ROT, depol, &
            tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, albedo, BPDFparam, solar_zenith, relat_azymuth, &
                  sensor_zenith, MISSING, &
            verbose, radiance_VL, reflectance_VL, BR, Q, U, BR_Q, BR_U, rc) &
            bind(C, name="ab200_vlidort_vector_lambert_bpdf_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, n0, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), pmomI(*), &
               tauL(*), ssaL(*), &
            pmomL(*), pe(*), he(*), te(*), albedo(*), BPDFparam(*), solar_zenith(*), relat_azymuth(*), sensor_zenith(*)
         real(c_double), intent(out) :: radiance_VL(*), reflectance_VL(*), BR(*), Q(*), U(*), BR_Q(*), BR_U(*)
         integer(c_int), intent(out) :: rc
      end function
      ! ---- NOBM water-leaving forms
      integer(c_int) function ab200_vlidort_vector_cx_nobm(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
            depol, tau, ssa, pmom, &
            pe, he, te, U10m, V10m, mr, sleave, sleave_iso, sleave_adjust, solar_zenith, relat_azymuth, sensor_zenith, MISSING, &
                  verbose, &
            radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc, ADJUSTED_SLEAVE) &
            bind(C, name="ab200_vlidort_vector_cx_nobm")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, sleave_iso, sleave_adjust, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), pe(*), he(*), te(*), U10m(*), &
               V10m(*), mr(*), &
            sleave(*), solar_zenith(*), relat_azymuth(*), sensor_zenith(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BRDF(*), Q(*), U(*), BRDF_Q(*), BRDF_U(*), &
               ADJUSTED_SLEAVE(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_scalar_gisscx_nobm_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, &
            nPol, ROT, depol, tau, &
            ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, sleave, sleave_iso, &
                  sleave_adjust, solar_zenith, &
            relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc, ADJUSTED_SLEAVE) &
            bind(C, name="ab200_vlidort_scalar_gisscx_nobm_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, sleave_iso, sleave_adjust, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), g(*), pmom(*), tauI(*), ssaI(*), gI(*), &
               pmomI(*), tauL(*), &
            ssaL(*), gL(*), pmomL(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), sleave(*), solar_zenith(*), &
                  relat_azymuth(*), sensor_zenith(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BRDF(*), ADJUSTED_SLEAVE(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_scalar_ocigisscx_nobm_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, &
            nPol, ROT, depol, &
            alpha, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, sleave, &
                  sleave_iso, sleave_adjust, &
            solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, &
                  BRDF, rc, &
            ADJUSTED_SLEAVE) &
            bind(C, name="ab200_vlidort_scalar_ocigisscx_nobm_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, sleave_iso, sleave_adjust, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), g(*), pmom(*), tauI(*), &
               ssaI(*), gI(*), pmomI(*), &
            tauL(*), ssaL(*), gL(*), pmomL(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), sleave(*), solar_zenith(*), &
                  relat_azymuth(*), &
            sensor_zenith(*), flux_factor(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BRDF(*), ADJUSTED_SLEAVE(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_gisscx_nobm_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, &
            nPol, ROT, depol, tau, &
            ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, sleave, sleave_iso, sleave_adjust, &
                  solar_zenith, &
            relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, &
                  rc, ADJUSTED_SLEAVE) &
            bind(C, name="ab200_vlidort_vector_gisscx_nobm_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, sleave_iso, sleave_adjust, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), pmomI(*), &
               tauL(*), ssaL(*), &
            pmomL(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), sleave(*), solar_zenith(*), relat_azymuth(*), sensor_zenith(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BRDF(*), Q(*), U(*), BRDF_Q(*), BRDF_U(*), &
               ADJUSTED_SLEAVE(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_vlidort_vector_ocigisscx_nobm_cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, &
            nPol, ROT, depol, &
            alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, sleave, sleave_iso, &
                  sleave_adjust, solar_zenith, &
            relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, &
                  BRDF_Q, BRDF_U, rc, &
            ADJUSTED_SLEAVE) &
            bind(C, name="ab200_vlidort_vector_ocigisscx_nobm_cloud")
         import :: c_int, c_double
         integer(c_int), value :: km, nch, nobs, nstreams, plane_parallel, nMom, nPol, sleave_iso, sleave_adjust, verbose
         real(c_double), value :: MISSING
         real(c_double), intent(in)  :: channels(*), ROT(*), depol(*), alpha(*), tau(*), ssa(*), pmom(*), tauI(*), ssaI(*), &
               pmomI(*), tauL(*), &
            ssaL(*), pmomL(*), pe(*), he(*), te(*), U10m(*), V10m(*), mr(*), sleave(*), solar_zenith(*), relat_azymuth(*), &
                  sensor_zenith(*), &
            flux_factor(*)
         real(c_double), intent(out) :: radiance_VL_SURF(*), reflectance_VL_SURF(*), BRDF(*), Q(*), U(*), BRDF_Q(*), BRDF_U(*), &
               ADJUSTED_SLEAVE(*)
         integer(c_int), intent(out) :: rc
      end function
   end interface
contains
   ! the optional switches of the reference's vector wrappers (GV/VLIDORT_LAMBERT.F90:176,350; GV/VLIDORT_BRDF_MODIS.F90:210,417):
   ! DO_BOA (bottom-of-atmosphere output) and DO_2OS_CORRECTION are outside this library's path -> refuse loudly (rc = -2)
   logical function ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      ab200_unsupported_switch = .false.
      if (present(DO_2OS_CORRECTION)) ab200_unsupported_switch = ab200_unsupported_switch .or. DO_2OS_CORRECTION
      if (present(DO_BOA)) ab200_unsupported_switch = ab200_unsupported_switch .or. DO_BOA
   end function ab200_unsupported_switch
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

module VLIDORT_LAMBERT                  ! replaces geo_vlidort/VLIDORT_LAMBERT.F90:34-1075
   use ab200_c_api
   implicit none
   public :: VLIDORT_Scalar_Lambert, VLIDORT_Vector_Lambert, VLIDORT_Scalar_Lambert_Cloud, VLIDORT_Vector_Lambert_Cloud
   interface VLIDORT_Vector_Lambert      ! VLIDORT_LAMBERT.F90:14-17
      module procedure VLIDORT_Vector_Lambert_SingleGeom
      module procedure VLIDORT_Vector_Lambert_MultiGeom
   end interface
   interface VLIDORT_Vector_Lambert_Cloud   ! VLIDORT_LAMBERT.F90:19-22
      module procedure VLIDORT_Vector_Lambert_Cloud_SingleGeom
      module procedure VLIDORT_Vector_Lambert_Cloud_MultiGeom
   end interface
contains
   subroutine VLIDORT_Scalar_Lambert(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, g, &
         pmom, pe, he, te, albedo, &
                                     solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL, reflectance_VL, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs)
      real*8,  intent(in)  :: pmom(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), albedo(nobs,nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL(nobs,nch), reflectance_VL(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_scalar_lambert(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, &
            tau, ssa, g, pmom, &
                                         pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                               radiance_VL, reflectance_VL, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_Lambert

   subroutine VLIDORT_Vector_Lambert_SingleGeom(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, &
         alpha, tau, ssa, pmom, pe, he, te, albedo, &
                                     solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, radiance_VL, &
                                           reflectance_VL, Q, U, rc, &
                                     DO_2OS_CORRECTION, DO_BOA)
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs)
      real*8,  intent(in)  :: pmom(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), albedo(nobs,nch), &
            flux_factor(nch,nobs)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL(nobs,nch), reflectance_VL(nobs,nch), Q(nobs,nch), U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      if (ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_lambert(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, &
            alpha, tau, ssa, pmom, &
                                         pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, &
                                               verbose, &
                                         radiance_VL, reflectance_VL, Q, U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_Lambert_SingleGeom

   ! VLIDORT_LAMBERT.F90:346-520. ngeom is not capped at MAX_USER_OBSGEOMS; albedo(:,:,1) is what the reference uses for every
   !  geometry
   subroutine VLIDORT_Vector_Lambert_MultiGeom(km, nch, nobs, ngeom, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, &
         alpha, tau, ssa, pmom, pe, he, te, &
                                     albedo, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, &
                                           radiance_VL, reflectance_VL, Q, U, rc, &
                                     DO_2OS_CORRECTION, DO_BOA)
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      integer, intent(in)  :: km, nch, nobs, ngeom, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs)
      real*8,  intent(in)  :: pmom(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), albedo(nobs,nch,ngeom), &
            flux_factor(nch,nobs)
      real*8,  intent(in)  :: solar_zenith(nobs,ngeom), relat_azymuth(nobs,ngeom), sensor_zenith(nobs,ngeom), MISSING
      real*8,  intent(out) :: radiance_VL(nobs,nch,ngeom), reflectance_VL(nobs,nch,ngeom), Q(nobs,nch,ngeom), U(nobs,nch,ngeom)
      integer, intent(out) :: rc
      integer :: ier
      if (ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)) then
         rc = -2; return
      end if
      ! albedo(:,:,1) is contiguous at the start of albedo(nobs,nch,ngeom): passed as is
      ier = ab200_vlidort_vector_lambert_multigeom(km, nch, nobs, ngeom, channels, nstreams, merge(1, 0, plane_parallel), nMom, &
            nPol, ROT, depol, alpha, tau, ssa, &
                                                   pmom, pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, &
                                                         flux_factor, MISSING, verbose, &
                                                   radiance_VL, reflectance_VL, Q, U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_Lambert_MultiGeom
   ! VLIDORT_LAMBERT.F90:524-686: clear-sky scalar runner with alpha and the solar flux; the cloud arrays do not enter
   subroutine VLIDORT_Scalar_Lambert_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, alpha, &
         tau, ssa, g, pmom, tauI, &
            ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, &
                  MISSING, verbose, &
            radiance_VL, reflectance_VL, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), &
            ssa(km,nch,nobs), g(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), gI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), &
                  tauL(km,nch,nobs), &
            ssaL(km,nch,nobs), gL(km,nch,nobs), pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), &
                  albedo(nobs,nch), &
            solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL(nobs,nch), reflectance_VL(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_scalar_lambert_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, &
            depol, alpha, tau, &
            ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, albedo, solar_zenith, relat_azymuth, &
                  sensor_zenith, flux_factor, &
            MISSING, verbose, radiance_VL, reflectance_VL, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_Lambert_Cloud
   ! VLIDORT_LAMBERT.F90:688-879 (f2py VECTOR_LAMBERT_CLOUD): clear-sky runner (:843)
   subroutine VLIDORT_Vector_Lambert_Cloud_SingleGeom(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, &
         alpha, tau, ssa, &
            pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, &
                  flux_factor, MISSING, &
            verbose, radiance_VL, reflectance_VL, Q, U, rc, DO_2OS_CORRECTION, DO_BOA)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), &
                  ssaL(km,nch,nobs), &
            pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), albedo(nobs,nch), solar_zenith(nobs), &
                  relat_azymuth(nobs), &
            sensor_zenith(nobs), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL(nobs,nch), reflectance_VL(nobs,nch), Q(nobs,nch), U(nobs,nch)
      integer, intent(out) :: rc
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      integer :: ier
      if (ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_lambert_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, &
            depol, alpha, tau, &
            ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, &
                  flux_factor, MISSING, &
            verbose, radiance_VL, reflectance_VL, Q, U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_Lambert_Cloud_SingleGeom
   ! VLIDORT_LAMBERT.F90:881-1075: clear-sky runner (:1036), no flux_factor; albedo(:,:,1) is passed as is
   subroutine VLIDORT_Vector_Lambert_Cloud_MultiGeom(km, nch, nobs, ngeom, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
         depol, alpha, tau, &
            ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, &
                  MISSING, verbose, &
            radiance_VL, reflectance_VL, Q, U, rc, DO_2OS_CORRECTION, DO_BOA)
      integer, intent(in)  :: km, nch, nobs, ngeom, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), &
                  ssaL(km,nch,nobs), &
            pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), albedo(nobs,nch,ngeom), &
                  solar_zenith(nobs,ngeom), &
            relat_azymuth(nobs,ngeom), sensor_zenith(nobs,ngeom), MISSING
      real*8,  intent(out) :: radiance_VL(nobs,nch,ngeom), reflectance_VL(nobs,nch,ngeom), Q(nobs,nch,ngeom), U(nobs,nch,ngeom)
      integer, intent(out) :: rc
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      integer :: ier
      if (ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_lambert_cloud_multigeom(km, nch, nobs, ngeom, channels, nstreams, merge(1, 0, plane_parallel), &
            nMom, nPol, ROT, &
            depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, albedo, solar_zenith, relat_azymuth, &
                  sensor_zenith, &
            MISSING, verbose, radiance_VL, reflectance_VL, Q, U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_Lambert_Cloud_MultiGeom
end module VLIDORT_LAMBERT

module VLIDORT_BRDF_CX                  ! replaces geo_vlidort/VLIDORT_BRDF_CX.F90:587-762 (note the output order Q,U,BRDF,...)
   use ab200_c_api
   implicit none
   public :: VLIDORT_Vector_CX, VLIDORT_Vector_GissCX, VLIDORT_Vector_OCICX, VLIDORT_Vector_OCIGissCX, VLIDORT_Scalar_CX, &
         VLIDORT_Scalar_GissCX
   public :: VLIDORT_Vector_CX_NOBM, VLIDORT_Scalar_GissCX_NOBM_Cloud, VLIDORT_Scalar_OCIGissCX_NOBM_Cloud, &
         VLIDORT_Vector_GissCX_NOBM_Cloud, &
             VLIDORT_Vector_OCIGissCx_NOBM_Cloud
   public :: VLIDORT_Vector_CX_Cloud, VLIDORT_Vector_GissCx_Cloud, VLIDORT_Vector_OCIGissCX_Cloud, VLIDORT_Scalar_GissCX_Cloud, &
         VLIDORT_Scalar_OCIGissCX_Cloud
   interface VLIDORT_Vector_CX_Cloud     ! VLIDORT_BRDF_CX.F90:33-36
      module procedure VLIDORT_Vector_CX_Cloud_SingleGeom
      module procedure VLIDORT_Vector_CX_Cloud_MultiGeom
   end interface
   interface VLIDORT_Vector_CX           ! VLIDORT_BRDF_CX.F90:27-30
      module procedure VLIDORT_Vector_CX_SingleGeom
      module procedure VLIDORT_Vector_CX_MultiGeom
   end interface
contains
   ! VLIDORT_BRDF_CX.F90:420-584 / 1553-1717: scalar forms; BRDF(1,nobs,nch)
   subroutine VLIDORT_Scalar_CX(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, g, pmom, &
         pe, he, te, U10m, V10m, mr, &
                                solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, &
                                      reflectance_VL_SURF, BRDF, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BRDF(1,nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_scalar_cx(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, tau, &
            ssa, g, pmom, pe, he, te, &
                                    U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                          radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_CX
   subroutine VLIDORT_Scalar_GissCX(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, g, &
         pmom, pe, he, te, U10m, V10m, mr, &
                                solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, &
                                      reflectance_VL_SURF, BRDF, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BRDF(1,nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_scalar_gisscx(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, &
            tau, ssa, g, pmom, pe, he, te, &
                                        U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                              radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_GissCX
   ! VLIDORT_BRDF_CX.F90:766-947 (note the output order Q, U, BRDF, ... of the Fortran interface)
   subroutine VLIDORT_Vector_CX_MultiGeom(km, nch, nobs, ngeom, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, &
         ssa, pmom, pe, he, te, U10m, V10m, mr, &
                                solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, &
                                      reflectance_VL_SURF, &
                                Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, ngeom, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs,ngeom), relat_azymuth(nobs,ngeom), sensor_zenith(nobs,ngeom), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch,ngeom), reflectance_VL_SURF(nobs,nch,ngeom), Q(nobs,nch,ngeom), &
            U(nobs,nch,ngeom)
      real*8,  intent(out) :: BRDF(nobs,nch,ngeom), BRDF_Q(nobs,nch,ngeom), BRDF_U(nobs,nch,ngeom)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_cx_multigeom(km, nch, nobs, ngeom, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, &
            ROT, depol, tau, ssa, pmom, pe, he, te, &
                                              U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                              radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_CX_MultiGeom
   subroutine VLIDORT_Vector_CX_SingleGeom(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, &
         pmom, pe, he, te, U10m, V10m, mr, &
                                solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, &
                                      reflectance_VL_SURF, &
                                Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BRDF(nobs,nch), BRDF_Q(nobs,nch), BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_cx(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, tau, &
            ssa, pmom, pe, he, te, &
                                    U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                    radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_CX_SingleGeom
   subroutine VLIDORT_Vector_GissCX(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, &
         pe, he, te, U10m, V10m, mr, &
                                solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, &
                                      reflectance_VL_SURF, &
                                Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BRDF(nobs,nch), BRDF_Q(nobs,nch), BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_gisscx(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, &
            tau, ssa, pmom, pe, he, te, &
                                    U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                    radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_GissCX
   subroutine VLIDORT_Vector_OCICX(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, &
         pe, he, te, U10m, V10m, mr, &
                                solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, &
                                      reflectance_VL_SURF, &
                                Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BRDF(nobs,nch), BRDF_Q(nobs,nch), BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_ocicx(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, &
            tau, ssa, pmom, pe, he, te, &
                                    U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                    radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_OCICX
   subroutine VLIDORT_Vector_OCIGissCX(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, alpha, tau, &
         ssa, pmom, pe, he, te, &
                                       U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, &
                                       ! VLIDORT_BRDF_CX.F90:54-236
                                       radiance_VL_SURF, reflectance_VL_SURF, Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs)
      real*8,  intent(in)  :: pmom(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), &
            mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BRDF(nobs,nch), BRDF_Q(nobs,nch), BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_ocigisscx(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, &
            depol, alpha, tau, ssa, pmom, &
                                           pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, &
                                                 MISSING, verbose, &
                                           radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_OCIGissCX
   subroutine VLIDORT_Vector_GissCx_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, &
         pmom, tauI, ssaI, pmomI, &
                                          tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, &
                                                sensor_zenith, MISSING, verbose, &
                                          ! VLIDORT_BRDF_CX.F90:2276-2467
                                          radiance_VL_SURF, reflectance_VL_SURF, Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), &
            ssaL(km,nch,nobs)
      real*8,  intent(in)  :: pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), &
            mr(nch)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BRDF(nobs,nch), BRDF_Q(nobs,nch), BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_gisscx_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, &
            depol, tau, ssa, pmom, &
                                              tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, &
                                                    relat_azymuth, &
                                              sensor_zenith, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, &
                                                    U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_GissCx_Cloud
   ! VLIDORT_BRDF_CX.F90:951-1144 (f2py VECTOR_CX_CLOUD): clear-sky runner (:1118)
   subroutine VLIDORT_Vector_CX_Cloud_SingleGeom(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, &
         ssa, pmom, tauI, &
            ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, &
                  verbose, &
            radiance_VL_SURF, reflectance_VL_SURF, Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), &
            tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), ssaL(km,nch,nobs), &
                  pmomL(km,nch,nobs,nMom,nPol), &
            pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch), solar_zenith(nobs), relat_azymuth(nobs), &
            sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch), &
            BRDF(nobs,nch), BRDF_Q(nobs,nch), &
            BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_cx_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, &
            depol, tau, ssa, pmom, &
            tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, &
                  MISSING, verbose, &
            radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_CX_Cloud_SingleGeom
   ! VLIDORT_BRDF_CX.F90:1147-1348: clear-sky runner (:1320)
   subroutine VLIDORT_Vector_CX_Cloud_MultiGeom(km, nch, nobs, ngeom, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
         depol, tau, ssa, pmom, &
            tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, &
                  MISSING, verbose, &
            radiance_VL_SURF, reflectance_VL_SURF, Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, ngeom, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), &
            tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), ssaL(km,nch,nobs), &
                  pmomL(km,nch,nobs,nMom,nPol), &
            pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch), solar_zenith(nobs,ngeom), &
                  relat_azymuth(nobs,ngeom), &
            sensor_zenith(nobs,ngeom), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch,ngeom), reflectance_VL_SURF(nobs,nch,ngeom), Q(nobs,nch,ngeom), &
            U(nobs,nch,ngeom), &
            BRDF(nobs,nch,ngeom), BRDF_Q(nobs,nch,ngeom), BRDF_U(nobs,nch,ngeom)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_cx_cloud_multigeom(km, nch, nobs, ngeom, channels, nstreams, merge(1, 0, plane_parallel), nMom, &
            nPol, ROT, depol, &
            tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, &
                  sensor_zenith, MISSING, &
            verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_CX_Cloud_MultiGeom
   ! VLIDORT_BRDF_CX.F90:2468-2668: cloud runner (:2642)
   subroutine VLIDORT_Vector_OCIGissCX_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, alpha, &
         tau, ssa, pmom, tauI, &
            ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, &
                  MISSING, verbose, &
            radiance_VL_SURF, reflectance_VL_SURF, Q, U, BRDF, BRDF_Q, BRDF_U, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), &
                  ssaL(km,nch,nobs), &
            pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch), &
                  solar_zenith(nobs), &
            relat_azymuth(nobs), sensor_zenith(nobs), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch), &
            BRDF(nobs,nch), BRDF_Q(nobs,nch), &
            BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_vector_ocigisscx_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, &
            ROT, depol, alpha, tau, &
            ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, &
                  sensor_zenith, flux_factor, &
            MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_OCIGissCX_Cloud
   ! VLIDORT_BRDF_CX.F90:2084-2273: scalar cloud runner (:2251)
   subroutine VLIDORT_Scalar_OCIGissCX_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, alpha, &
         tau, ssa, g, pmom, &
            tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, &
                  sensor_zenith, flux_factor, &
            MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), &
            ssa(km,nch,nobs), g(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), gI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), &
                  tauL(km,nch,nobs), &
            ssaL(km,nch,nobs), gL(km,nch,nobs), pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), &
                  U10m(nobs), V10m(nobs), &
            mr(nch), solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BRDF(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_scalar_ocigisscx_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, &
            ROT, depol, alpha, tau, &
            ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, &
                  sensor_zenith, &
            flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_OCIGissCX_Cloud
   ! VLIDORT_BRDF_CX.F90:1898-2083: scalar cloud runner (:2058)
   subroutine VLIDORT_Scalar_GissCX_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, &
         g, pmom, tauI, ssaI, &
            gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, &
                  verbose, &
            radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), gI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), &
                  tauL(km,nch,nobs), &
            ssaL(km,nch,nobs), gL(km,nch,nobs), pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), &
                  U10m(nobs), V10m(nobs), &
            mr(nch), solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BRDF(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_scalar_gisscx_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, &
            depol, tau, ssa, g, &
            pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, &
                  sensor_zenith, MISSING, &
            verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_GissCX_Cloud
   ! VLIDORT_BRDF_CX.F90:1351-1550.  sleave_adjust = .true. (VLIDORT's transmittance iteration) is refused: rc = -2
   subroutine VLIDORT_Vector_CX_NOBM(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, &
         pe, he, te, U10m, &
            V10m, mr, sleave, sleave_iso, sleave_adjust, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                  radiance_VL_SURF, &
            reflectance_VL_SURF, Q, U, BRDF, BRDF_Q, BRDF_U, rc, ADJUSTED_SLEAVE)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel, sleave_iso, sleave_adjust
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), &
            pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch), sleave(nch,nobs), solar_zenith(nobs), &
                  relat_azymuth(nobs), &
            sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch), &
            BRDF(nobs,nch), BRDF_Q(nobs,nch), &
            BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      real*8,  optional, intent(out) :: ADJUSTED_SLEAVE(nobs,nch)
      real*8, allocatable :: adj(:,:)
      integer :: ier
      allocate(adj(nobs,nch))
      ier = ab200_vlidort_vector_cx_nobm(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, depol, &
            tau, ssa, pmom, pe, &
            he, te, U10m, V10m, mr, sleave, merge(1, 0, sleave_iso), merge(1, 0, sleave_adjust), solar_zenith, relat_azymuth, &
                  sensor_zenith, MISSING, &
            verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc, adj)
      if (present(ADJUSTED_SLEAVE)) ADJUSTED_SLEAVE = adj
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_CX_NOBM
   ! VLIDORT_BRDF_CX.F90:2671-2880.  sleave_adjust = .true. (VLIDORT's transmittance iteration) is refused: rc = -2
   subroutine VLIDORT_Scalar_GissCX_NOBM_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, &
         ssa, g, pmom, tauI, &
            ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, sleave, sleave_iso, sleave_adjust, solar_zenith, &
                  relat_azymuth, &
            sensor_zenith, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc, ADJUSTED_SLEAVE)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel, sleave_iso, sleave_adjust
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), gI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), &
                  tauL(km,nch,nobs), &
            ssaL(km,nch,nobs), gL(km,nch,nobs), pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), &
                  U10m(nobs), V10m(nobs), &
            mr(nch), sleave(nch,nobs), solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BRDF(nobs,nch)
      integer, intent(out) :: rc
      real*8,  optional, intent(out) :: ADJUSTED_SLEAVE(nobs,nch)
      real*8, allocatable :: adj(:,:)
      integer :: ier
      allocate(adj(nobs,nch))
      ier = ab200_vlidort_scalar_gisscx_nobm_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, &
            ROT, depol, tau, ssa, &
            g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, sleave, merge(1, 0, sleave_iso), &
            merge(1, 0, sleave_adjust), solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, &
                  reflectance_VL_SURF, BRDF, &
            rc, adj)
      if (present(ADJUSTED_SLEAVE)) ADJUSTED_SLEAVE = adj
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_GissCX_NOBM_Cloud
   ! VLIDORT_BRDF_CX.F90:2883-3099.  sleave_adjust = .true. (VLIDORT's transmittance iteration) is refused: rc = -2
   subroutine VLIDORT_Scalar_OCIGissCX_NOBM_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, &
         alpha, tau, ssa, g, &
            pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, sleave, sleave_iso, sleave_adjust, &
                  solar_zenith, &
            relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, rc, &
                  ADJUSTED_SLEAVE)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel, sleave_iso, sleave_adjust
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), &
            ssa(km,nch,nobs), g(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), gI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), &
                  tauL(km,nch,nobs), &
            ssaL(km,nch,nobs), gL(km,nch,nobs), pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), &
                  U10m(nobs), V10m(nobs), &
            mr(nch), sleave(nch,nobs), solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BRDF(nobs,nch)
      integer, intent(out) :: rc
      real*8,  optional, intent(out) :: ADJUSTED_SLEAVE(nobs,nch)
      real*8, allocatable :: adj(:,:)
      integer :: ier
      allocate(adj(nobs,nch))
      ier = ab200_vlidort_scalar_ocigisscx_nobm_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, &
            nPol, ROT, depol, alpha, &
            tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, U10m, V10m, mr, sleave, merge(1, 0, &
                  sleave_iso), &
            merge(1, 0, sleave_adjust), solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, &
                  radiance_VL_SURF, &
            reflectance_VL_SURF, BRDF, rc, adj)
      if (present(ADJUSTED_SLEAVE)) ADJUSTED_SLEAVE = adj
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_OCIGissCX_NOBM_Cloud
   ! VLIDORT_BRDF_CX.F90:3102-3317.  sleave_adjust = .true. (VLIDORT's transmittance iteration) is refused: rc = -2
   subroutine VLIDORT_Vector_GissCX_NOBM_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, &
         ssa, pmom, tauI, ssaI, &
            pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, sleave, sleave_iso, sleave_adjust, solar_zenith, &
                  relat_azymuth, sensor_zenith, &
            MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, Q, U, BRDF, BRDF_Q, BRDF_U, rc, ADJUSTED_SLEAVE)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel, sleave_iso, sleave_adjust
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), &
            tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), ssaL(km,nch,nobs), &
                  pmomL(km,nch,nobs,nMom,nPol), &
            pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch), sleave(nch,nobs), solar_zenith(nobs), &
                  relat_azymuth(nobs), &
            sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch), &
            BRDF(nobs,nch), BRDF_Q(nobs,nch), &
            BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      real*8,  optional, intent(out) :: ADJUSTED_SLEAVE(nobs,nch)
      real*8, allocatable :: adj(:,:)
      integer :: ier
      allocate(adj(nobs,nch))
      ier = ab200_vlidort_vector_gisscx_nobm_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, &
            ROT, depol, tau, ssa, &
            pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, sleave, merge(1, 0, sleave_iso), merge(1, 0, &
                  sleave_adjust), &
            solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BRDF, Q, U, &
                  BRDF_Q, BRDF_U, rc, adj)
      if (present(ADJUSTED_SLEAVE)) ADJUSTED_SLEAVE = adj
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_GissCX_NOBM_Cloud
   ! VLIDORT_BRDF_CX.F90:3320-3538 (f2py VECTOR_OCIGissCX_NOBM_Cloud). sleave_adjust = .true. (VLIDORT's transmittance iteration)
   !  is refused: rc = -2
   subroutine VLIDORT_Vector_OCIGissCx_NOBM_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, &
         alpha, tau, ssa, pmom, &
            tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, sleave, sleave_iso, sleave_adjust, solar_zenith, &
                  relat_azymuth, &
            sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, Q, U, BRDF, BRDF_Q, BRDF_U, rc, &
                  ADJUSTED_SLEAVE)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel, sleave_iso, sleave_adjust
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), &
                  ssaL(km,nch,nobs), &
            pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), U10m(nobs), V10m(nobs), mr(nch), &
                  sleave(nch,nobs), &
            solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), Q(nobs,nch), U(nobs,nch), &
            BRDF(nobs,nch), BRDF_Q(nobs,nch), &
            BRDF_U(nobs,nch)
      integer, intent(out) :: rc
      real*8,  optional, intent(out) :: ADJUSTED_SLEAVE(nobs,nch)
      real*8, allocatable :: adj(:,:)
      integer :: ier
      allocate(adj(nobs,nch))
      ier = ab200_vlidort_vector_ocigisscx_nobm_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, &
            nPol, ROT, depol, alpha, &
            tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m, mr, sleave, merge(1, 0, sleave_iso), &
            merge(1, 0, sleave_adjust), solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, &
                  radiance_VL_SURF, &
            reflectance_VL_SURF, BRDF, Q, U, BRDF_Q, BRDF_U, rc, adj)
      if (present(ADJUSTED_SLEAVE)) ADJUSTED_SLEAVE = adj
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_OCIGissCx_NOBM_Cloud
end module VLIDORT_BRDF_CX

module VLIDORT_BRDF_MODIS               ! replaces geo_vlidort/VLIDORT_BRDF_MODIS.F90:206-410
   use ab200_c_api
   implicit none
   public :: VLIDORT_Vector_LandMODIS, VLIDORT_Scalar_LandMODIS, VLIDORT_Vector_LandMODIS_Cloud, VLIDORT_Scalar_LandMODIS_Cloud
   interface VLIDORT_Vector_LandMODIS_Cloud   ! VLIDORT_BRDF_MODIS.F90:19-22
      module procedure VLIDORT_Vector_LandMODIS_Cloud_SingleGeom
      module procedure VLIDORT_Vector_LandMODIS_Cloud_MultiGeom
   end interface
   interface VLIDORT_Vector_LandMODIS    ! VLIDORT_BRDF_MODIS.F90:15-18
      module procedure VLIDORT_Vector_LandMODIS_SingleGeom
      module procedure VLIDORT_Vector_LandMODIS_MultiGeom
   end interface VLIDORT_Vector_LandMODIS
contains
   ! VLIDORT_BRDF_MODIS.F90:35-202
   subroutine VLIDORT_Scalar_LandMODIS(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, g, &
         pmom, pe, he, te, &
                                       kernel_wt, param, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                       radiance_VL_SURF, reflectance_VL_SURF, BR, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), kernel_wt(3,nch,nobs), param(2,nch,nobs)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BR(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_scalar_brdf_modis(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, ROT, &
            depol, tau, ssa, g, pmom, &
                                            pe, he, te, kernel_wt, param, solar_zenith, relat_azymuth, sensor_zenith, MISSING, &
                                                  verbose, &
                                            radiance_VL_SURF, reflectance_VL_SURF, BR, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_LandMODIS
   ! VLIDORT_BRDF_MODIS.F90:413-...: kernel_wt / param of the first geometry are used (one surface per observation)
   subroutine VLIDORT_Vector_LandMODIS_MultiGeom(km, nch, nobs, ngeom, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
         depol, tau, ssa, pmom, pe, he, te, &
                                       kernel_wt, param, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                       radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc, DO_2OS_CORRECTION, DO_BOA)
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      integer, intent(in)  :: km, nch, nobs, ngeom, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), kernel_wt(3,nch,nobs), param(2,nch,nobs)
      real*8,  intent(in)  :: solar_zenith(nobs,ngeom), relat_azymuth(nobs,ngeom), sensor_zenith(nobs,ngeom), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch,ngeom), reflectance_VL_SURF(nobs,nch,ngeom), BR(nobs,nch,ngeom), &
            Q(nobs,nch,ngeom), U(nobs,nch,ngeom)
      real*8,  intent(out) :: BR_Q(nobs,nch,ngeom), BR_U(nobs,nch,ngeom)
      integer, intent(out) :: rc
      integer :: ier
      if (ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_brdf_modis_multigeom(km, nch, nobs, ngeom, channels, nstreams, merge(1, 0, plane_parallel), &
            nMom, nPol, 3, 2, ROT, depol, tau, ssa, &
                                                      pmom, pe, he, te, kernel_wt, param, solar_zenith, relat_azymuth, &
                                                            sensor_zenith, MISSING, verbose, &
                                                      radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_LandMODIS_MultiGeom
   subroutine VLIDORT_Vector_LandMODIS_SingleGeom(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, &
         tau, ssa, pmom, pe, he, te, &
                                       kernel_wt, param, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, &
                                       radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc, DO_2OS_CORRECTION, DO_BOA)
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), kernel_wt(3,nch,nobs), param(2,nch,nobs)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BR(nobs,nch), Q(nobs,nch), U(nobs,nch), &
            BR_Q(nobs,nch), BR_U(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      if (ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_brdf_modis(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, 3, 2, &
            ROT, depol, tau, ssa, pmom, &
                                            pe, he, te, kernel_wt, param, solar_zenith, relat_azymuth, sensor_zenith, MISSING, &
                                                  verbose, &
                                            radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_LandMODIS_SingleGeom
   ! VLIDORT_Vector_LandMODIS_Cloud_SingleGeom (VLIDORT_BRDF_MODIS.F90:825-1053): the specific the f2py VECTOR_BRDF_MODIS_CLOUD
   !  reaches;
   ! the optional DO_2OS_CORRECTION / DO_BOA switches are refused like in the clear-sky forms
   subroutine VLIDORT_Vector_LandMODIS_Cloud_SingleGeom(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, &
         depol, alpha, tau, ssa, pmom, &
                                             tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, kernel_wt, param, solar_zenith, &
                                                   relat_azymuth, &
                                             sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, &
                                                   reflectance_VL_SURF, BR, Q, U, BR_Q, &
                                             BR_U, rc, DO_2OS_CORRECTION, DO_BOA)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs)
      real*8,  intent(in)  :: pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: tauL(km,nch,nobs), ssaL(km,nch,nobs), pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), &
            te(km+1,nobs)
      real*8,  intent(in)  :: kernel_wt(3,nch,nobs), param(2,nch,nobs), solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs)
      real*8,  intent(in)  :: flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BR(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BR_Q(nobs,nch), BR_U(nobs,nch)
      integer, intent(out) :: rc
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      integer :: ier
      if (ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_brdf_modis_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, 3, &
            2, ROT, depol, alpha, &
                                                  tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, kernel_wt, &
                                                        param, solar_zenith, &
                                                  relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, &
                                                        reflectance_VL_SURF, &
                                                  BR, Q, U, BR_Q, BR_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_LandMODIS_Cloud_SingleGeom
   ! VLIDORT_BRDF_MODIS.F90:1055-1268: cloud runner (:1258) for all geometries
   subroutine VLIDORT_Vector_LandMODIS_Cloud_MultiGeom(km, nch, nobs, ngeom, channels, nstreams, plane_parallel, nMom, nPol, &
         ROT, depol, alpha, tau, &
            ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, kernel_wt, param, solar_zenith, relat_azymuth, &
                  sensor_zenith, flux_factor, &
            MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc, DO_2OS_CORRECTION, DO_BOA)
      integer, intent(in)  :: km, nch, nobs, ngeom, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), &
                  ssaL(km,nch,nobs), &
            pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), kernel_wt(3,nch,nobs), param(2,nch,nobs), &
            solar_zenith(nobs,ngeom), relat_azymuth(nobs,ngeom), sensor_zenith(nobs,ngeom), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch,ngeom), reflectance_VL_SURF(nobs,nch,ngeom), BR(nobs,nch,ngeom), &
            Q(nobs,nch,ngeom), &
            U(nobs,nch,ngeom), BR_Q(nobs,nch,ngeom), BR_U(nobs,nch,ngeom)
      integer, intent(out) :: rc
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      integer :: ier
      if (ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_brdf_modis_cloud_multigeom(km, nch, nobs, ngeom, channels, nstreams, merge(1, 0, &
            plane_parallel), nMom, nPol, 3, 2, &
            ROT, depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, kernel_wt, param, solar_zenith, &
                  relat_azymuth, &
            sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_LandMODIS_Cloud_MultiGeom
   ! VLIDORT_BRDF_MODIS.F90:625-823: scalar cloud runner (:797)
   subroutine VLIDORT_Scalar_LandMODIS_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, alpha, &
         tau, ssa, g, pmom, &
            tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, kernel_wt, param, solar_zenith, relat_azymuth, &
                  sensor_zenith, flux_factor, &
            MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BR, rc)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), &
            ssa(km,nch,nobs), g(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), gI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), &
                  tauL(km,nch,nobs), &
            ssaL(km,nch,nobs), gL(km,nch,nobs), pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), &
                  kernel_wt(3,nch,nobs), &
            param(2,nch,nobs), solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BR(nobs,nch)
      integer, intent(out) :: rc
      integer :: ier
      ier = ab200_vlidort_scalar_brdf_modis_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, &
            ROT, depol, alpha, tau, &
            ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, kernel_wt, param, solar_zenith, &
                  relat_azymuth, sensor_zenith, &
            flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BR, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Scalar_LandMODIS_Cloud
end module VLIDORT_BRDF_MODIS

! replaces geo_vlidort/VLIDORT_BRDF_MODIS_BPDF.F90:23-230,459-643 (+ the _BPDF_Cloud form)
module VLIDORT_BRDF_MODIS_BPDF
   use ab200_c_api
   implicit none
contains
   subroutine VLIDORT_Vector_LandMODIS_BPDF(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, &
         pmom, pe, he, te, &
                                            kernel_wt, RTLSparam, BPDFparam, solar_zenith, relat_azymuth, sensor_zenith, &
                                                  MISSING, verbose, &
                                            radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc, DO_BOA)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), kernel_wt(3,nch,nobs), RTLSparam(2,nch,nobs), &
            BPDFparam(3,nch,nobs)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BR(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BR_Q(nobs,nch), BR_U(nobs,nch)
      integer, intent(out) :: rc
      logical, optional, intent(in) :: DO_BOA
      integer :: ier
      if (ab200_unsupported_switch(DO_BOA = DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_brdf_modis_bpdf(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, 3, &
            2, 3, ROT, depol, tau, ssa, &
                                                 pmom, pe, he, te, kernel_wt, RTLSparam, BPDFparam, solar_zenith, relat_azymuth, &
                                                       sensor_zenith, MISSING, &
                                                 verbose, radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_LandMODIS_BPDF
   subroutine VLIDORT_Vector_BPDF(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, pe, &
         he, te, BPDFparam, &
                                  solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, &
                                        reflectance_VL_SURF, BR, Q, U, &
                                  BR_Q, BR_U, rc, DO_BOA)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), BPDFparam(3,nch,nobs)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BR(nobs,nch), Q(nobs,nch), U(nobs,nch)
      real*8,  intent(out) :: BR_Q(nobs,nch), BR_U(nobs,nch)
      integer, intent(out) :: rc
      logical, optional, intent(in) :: DO_BOA
      integer :: ier
      if (ab200_unsupported_switch(DO_BOA = DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_bpdf(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, 3, ROT, depol, &
            tau, ssa, pmom, pe, he, te, &
                                      BPDFparam, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL_SURF, &
                                            reflectance_VL_SURF, &
                                      BR, Q, U, BR_Q, BR_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_BPDF
   ! VLIDORT_BRDF_MODIS_BPDF.F90:232-455: place holder, clear-sky runner (:427) with alpha and the solar flux
   subroutine VLIDORT_Vector_LandMODIS_BPDF_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, &
         alpha, tau, ssa, pmom, &
            tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, kernel_wt, RTLSparam, BPDFparam, solar_zenith, relat_azymuth, &
                  sensor_zenith, &
            flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, BR_U, rc, DO_BOA)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), alpha(km,nobs,nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), &
                  ssaL(km,nch,nobs), &
            pmomL(km,nch,nobs,nMom,nPol), pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), kernel_wt(3,nch,nobs), &
                  RTLSparam(2,nch,nobs), &
            BPDFparam(3,nch,nobs), solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), flux_factor(nch,nobs), MISSING
      real*8,  intent(out) :: radiance_VL_SURF(nobs,nch), reflectance_VL_SURF(nobs,nch), BR(nobs,nch), Q(nobs,nch), U(nobs,nch), &
            BR_Q(nobs,nch), &
            BR_U(nobs,nch)
      integer, intent(out) :: rc
      logical, optional, intent(in) :: DO_BOA
      integer :: ier
      if (ab200_unsupported_switch(DO_BOA=DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_brdf_modis_bpdf_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, &
            nPol, 3, 2, 3, ROT, &
            depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, kernel_wt, RTLSparam, BPDFparam, &
                  solar_zenith, &
            relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose, radiance_VL_SURF, reflectance_VL_SURF, BR, Q, U, BR_Q, &
                  BR_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_LandMODIS_BPDF_Cloud
end module VLIDORT_BRDF_MODIS_BPDF

module VLIDORT_LAMBERT_BPDF                ! replaces geo_vlidort/VLIDORT_LAMBERT_BPDF.F90:23-223 (+ the _BPDF_Cloud form)
   use ab200_c_api
   implicit none
contains
   subroutine VLIDORT_Vector_Lambert_BPDF(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, &
         pmom, pe, he, te, albedo, &
                                          BPDFparam, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose, radiance_VL, &
                                                reflectance_VL, BR, Q, U, &
                                          BR_Q, BR_U, rc, DO_2OS_CORRECTION, DO_BOA)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol)
      real*8,  intent(in)  :: pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), albedo(nobs,nch), BPDFparam(3,nch,nobs)
      real*8,  intent(in)  :: solar_zenith(nobs), relat_azymuth(nobs), sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL(nobs,nch), reflectance_VL(nobs,nch), BR(nobs,nch), Q(nobs,nch), U(nobs,nch), &
            BR_Q(nobs,nch), BR_U(nobs,nch)
      integer, intent(out) :: rc
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      integer :: ier
      if (ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_lambert_bpdf(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, 3, &
            ROT, depol, tau, ssa, pmom, &
                                              pe, he, te, albedo, BPDFparam, solar_zenith, relat_azymuth, sensor_zenith, &
                                                    MISSING, verbose, radiance_VL, &
                                              reflectance_VL, BR, Q, U, BR_Q, BR_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_Lambert_BPDF
   ! VLIDORT_LAMBERT_BPDF.F90:228-440: place holder, clear-sky runner (:409)
   subroutine VLIDORT_Vector_Lambert_BPDF_Cloud(km, nch, nobs, channels, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, &
         ssa, pmom, tauI, &
            ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, albedo, BPDFparam, solar_zenith, relat_azymuth, sensor_zenith, MISSING, &
                  verbose, radiance_VL, &
            reflectance_VL, BR, Q, U, BR_Q, BR_U, rc, DO_2OS_CORRECTION, DO_BOA)
      integer, intent(in)  :: km, nch, nobs, nstreams, nMom, nPol, verbose
      logical, intent(in)  :: plane_parallel
      real*8,  intent(in)  :: channels(nch), ROT(km,nobs,nch), depol(nch), tau(km,nch,nobs), ssa(km,nch,nobs), &
            pmom(km,nch,nobs,nMom,nPol), &
            tauI(km,nch,nobs), ssaI(km,nch,nobs), pmomI(km,nch,nobs,nMom,nPol), tauL(km,nch,nobs), ssaL(km,nch,nobs), &
                  pmomL(km,nch,nobs,nMom,nPol), &
            pe(km+1,nobs), he(km+1,nobs), te(km+1,nobs), albedo(nobs,nch), BPDFparam(3,nch,nobs), solar_zenith(nobs), &
                  relat_azymuth(nobs), &
            sensor_zenith(nobs), MISSING
      real*8,  intent(out) :: radiance_VL(nobs,nch), reflectance_VL(nobs,nch), BR(nobs,nch), Q(nobs,nch), U(nobs,nch), &
            BR_Q(nobs,nch), BR_U(nobs,nch)
      integer, intent(out) :: rc
      logical, optional, intent(in) :: DO_2OS_CORRECTION, DO_BOA
      integer :: ier
      if (ab200_unsupported_switch(DO_2OS_CORRECTION, DO_BOA)) then
         rc = -2; return
      end if
      ier = ab200_vlidort_vector_lambert_bpdf_cloud(km, nch, nobs, channels, nstreams, merge(1, 0, plane_parallel), nMom, nPol, &
            3, ROT, depol, tau, &
            ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, albedo, BPDFparam, solar_zenith, relat_azymuth, &
                  sensor_zenith, MISSING, &
            verbose, radiance_VL, reflectance_VL, BR, Q, U, BR_Q, BR_U, rc)
      if (ier /= 0 .and. rc == 0) rc = ier
   end subroutine VLIDORT_Vector_Lambert_BPDF_Cloud
end module VLIDORT_LAMBERT_BPDF
