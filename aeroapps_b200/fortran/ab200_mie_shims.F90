! ISO_C_BINDING shims for Stage A: the reference's external subroutines, re-created on top of libaeroapps_b200.so.
!   VLIDORT_getAOPscalar / VLIDORT_getAOPvector   replace the bodies in Components/rtms/geo_vlidort/VLIDORT_MieMod.F90:10,124
!   getAOPscalar / getAOPvector / getExt / getInt / getEdgeVars   replace the bodies in GMAO_Shared/Chem_Base/MieObs_py.F90:27-770
! Names, argument lists and intents are the reference's, so geo_vlidort.F90, leo_vlidort_cloud.F90 and the f2py module MieObs_
! compile and link untouched. The NetCDF optics-file reader and the rc parser stay the reference's (Chem_MieCreate): the shim
!  takes the
! type(Chem_Mie) it returns, uploads the interpolated float32 tables once per (rcfile, channels, nMom) through
! ab200_mie_load_table_f32 and keeps the GPU handle in a small cache, so a caller that creates the tables per call (MieObs_py does)
! pays the upload once.  Link with -L<repo>/aeroapps_b200 -laeroapps_b200 -lcudart.
! NOT compiled in this repository's CI: neither the build image nor the GPU box has a Fortran compiler (DESIGN.md section 2);
! the C entry points underneath are exercised with the same Fortran-ordered buffers by tests/cabi/test_cabi.c and
!  tests/test_mie_gpu.py.
module ab200_mie_c_api
   use iso_c_binding
   implicit none
   interface
      integer(c_int) function ab200_mie_create(nch, channels_m, nmom, h) bind(C, name="ab200_mie_create")
         import :: c_int, c_float, c_ptr
         integer(c_int), value :: nch, nmom
         real(c_float), intent(in) :: channels_m(*)
         type(c_ptr), intent(out) :: h
      end function
      integer(c_int) function ab200_mie_destroy(h) bind(C, name="ab200_mie_destroy")
         import :: c_int, c_ptr
         type(c_ptr), value :: h
      end function
      integer(c_int) function ab200_mie_load_table_f32(h, species, nrh, nbin, npol, rh, reff, gf, rhop, rhod, vol, area, bext, &
            bsca, bbck, g, &
                                                       refr, refi, pback, pmom, rhi, rha) bind(C, name="ab200_mie_load_table_f32")
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         integer(c_int), value :: species, nrh, nbin, npol
         type(c_ptr), value :: rh, reff, gf, rhop, rhod, vol, area, bext, bsca, bbck, g, refr, refi, pback, pmom, rhi, rha
      end function
      integer(c_int) function ab200_get_aop_scalar(h, km, nobs, nch, nq, channels, vname, verbose, qm, rh, tau, ssa, g, rc) &
                                                   bind(C, name="ab200_get_aop_scalar")
         import :: c_int, c_float, c_ptr, c_char
         type(c_ptr), value :: h
         integer(c_int), value :: km, nobs, nch, nq, verbose
         real(c_float), intent(in) :: channels(*), qm(*), rh(*)
         character(kind=c_char), intent(in) :: vname(*)
         real(c_float), intent(out) :: tau(*), ssa(*), g(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_get_aop_vector(h, km, nobs, nch, nq, channels, vname, verbose, qm, rh, nmom, npol, tau, ssa, &
            g, pmom, rc) &
                                                   bind(C, name="ab200_get_aop_vector")
         import :: c_int, c_float, c_ptr, c_char
         type(c_ptr), value :: h
         integer(c_int), value :: km, nobs, nch, nq, verbose, nmom, npol
         real(c_float), intent(in) :: channels(*), qm(*), rh(*)
         character(kind=c_char), intent(in) :: vname(*)
         real(c_float), intent(out) :: tau(*), ssa(*), g(*), pmom(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_get_ext(h, km, nobs, nch, nq, channels, vname, verbose, qc, qm, rh, ext, sca, bsc, absc_sfc, &
            absc_toa, depol, rc) &
                                            bind(C, name="ab200_get_ext")
         import :: c_int, c_float, c_ptr, c_char
         type(c_ptr), value :: h
         integer(c_int), value :: km, nobs, nch, nq, verbose
         real(c_float), intent(in) :: channels(*), qc(*), qm(*), rh(*)
         character(kind=c_char), intent(in) :: vname(*)
         real(c_float), intent(out) :: ext(*), sca(*), bsc(*), absc_sfc(*), absc_toa(*), depol(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_get_int(h, km, nobs, nch, nq, channels, vname, verbose, qc, qm, rh, vol, area, refr, refi, &
            reff, rc) &
                                            bind(C, name="ab200_get_int")
         import :: c_int, c_float, c_ptr, c_char
         type(c_ptr), value :: h
         integer(c_int), value :: km, nobs, nch, nq, verbose
         real(c_float), intent(in) :: channels(*), qc(*), qm(*), rh(*)
         character(kind=c_char), intent(in) :: vname(*)
         real(c_float), intent(out) :: vol(*), area(*), refr(*), refi(*), reff(*)
         integer(c_int), intent(out) :: rc
      end function
      integer(c_int) function ab200_get_edge_vars(km, nobs, airdens, delp, ptop, pe, ze, te) bind(C, name="ab200_get_edge_vars")
         import :: c_int, c_float
         integer(c_int), value :: km, nobs
         real(c_float), value :: ptop
         real(c_float), intent(in) :: airdens(*), delp(*)
         real(c_float), intent(out) :: pe(*), ze(*), te(*)
      end function
   end interface
end module ab200_mie_c_api

! GPU-handle cache: one entry per (rcfile, channel list, nMom) seen so far
module ab200_mie_handles
   use iso_c_binding
   use Chem_MieMod
   use Chem_MieTableMod
   use ab200_mie_c_api
   implicit none
   private
   public :: ab200_handle_for, ab200_handle_for_rc
   integer, parameter :: MAXH = 16
   type entry_t
      type(c_ptr) :: h = c_null_ptr
      character(len=255) :: rcfile = ''
      integer :: nch = 0, nMom = -1
      real :: channels(64) = 0.
   end type
   type(entry_t), save :: cache(MAXH)
   integer, save :: nused = 0
contains
   ! species order of libaeroapps_b200 == member order of type(Chem_Mie) (Chem_MieMod.F90:65-99): DU SS BC OC BRC SU NI
   subroutine upload(h, species, t, wantMom, rc)
      type(c_ptr), intent(in) :: h
      integer, intent(in) :: species, wantMom
      type(Chem_MieTable), pointer :: t
      integer, intent(out) :: rc
      type(c_ptr) :: p_pback, p_pmom, p_gf, p_rhop, p_rhod, p_refr, p_refi
      integer :: npol
      rc = 0
      if (.not. associated(t)) return
      p_pback = c_null_ptr; p_pmom = c_null_ptr; p_gf = c_null_ptr; p_rhop = c_null_ptr; p_rhod = c_null_ptr
      p_refr = c_null_ptr; p_refi = c_null_ptr
      npol = max(t%nPol, 1)
      if (associated(t%pback)) then
         p_pback = c_loc(t%pback(1,1,1,1)); npol = size(t%pback, 4)
      end if
      if (wantMom > 0 .and. associated(t%pmom)) p_pmom = c_loc(t%pmom(1,1,1,1,1))
      if (associated(t%gf))   p_gf   = c_loc(t%gf(1,1))
      if (associated(t%rhop)) p_rhop = c_loc(t%rhop(1,1))
      if (associated(t%rhod)) p_rhod = c_loc(t%rhod(1,1))
      if (associated(t%refr)) p_refr = c_loc(t%refr(1,1,1))
      if (associated(t%refi)) p_refi = c_loc(t%refi(1,1,1))
      rc = ab200_mie_load_table_f32(h, species, t%nrh, t%nbin, npol, c_loc(t%rh(1)), c_loc(t%reff(1,1)), p_gf, p_rhop, p_rhod, &
                                    c_loc(t%vol(1,1)), c_loc(t%area(1,1)), c_loc(t%bext(1,1,1)), c_loc(t%bsca(1,1,1)), &
                                          c_loc(t%bbck(1,1,1)), &
                                    c_loc(t%g(1,1,1)), p_refr, p_refi, p_pback, p_pmom, c_loc(t%rhi(1)), c_loc(t%rha(1)))
   end subroutine

   ! the GPU handle of an existing type(Chem_Mie) (VLIDORT_MieMod callers create it once per run)
   function ab200_handle_for(mt, rc) result(h)
      type(Chem_Mie), intent(inout), target :: mt
      integer, intent(out) :: rc
      type(c_ptr) :: h
      integer :: i, n
      rc = 0
      n = min(mt%nch, 64)
      do i = 1, nused
         if (cache(i)%rcfile == mt%rcfile .and. cache(i)%nch == mt%nch .and. cache(i)%nMom == mt%nMom) then
            if (all(cache(i)%channels(1:n) == mt%channels(1:n))) then
               h = cache(i)%h; return
            end if
         end if
      end do
      rc = ab200_mie_create(mt%nch, mt%channels, mt%nMom, h)
      if (rc /= 0) return
      call upload(h, 0, mt%mie_DU,  mt%nMom, rc); if (rc /= 0) return
      call upload(h, 1, mt%mie_SS,  mt%nMom, rc); if (rc /= 0) return
      call upload(h, 2, mt%mie_BC,  mt%nMom, rc); if (rc /= 0) return
      call upload(h, 3, mt%mie_OC,  mt%nMom, rc); if (rc /= 0) return
      call upload(h, 4, mt%mie_BRC, mt%nMom, rc); if (rc /= 0) return
      call upload(h, 5, mt%mie_SU,  mt%nMom, rc); if (rc /= 0) return
      call upload(h, 6, mt%mie_NI,  mt%nMom, rc); if (rc /= 0) return
      if (nused == MAXH) then                       ! recycle the oldest entry
         i = ab200_mie_destroy(cache(1)%h)
         cache(1:MAXH-1) = cache(2:MAXH); nused = MAXH - 1
      end if
      nused = nused + 1
      cache(nused)%h = h; cache(nused)%rcfile = mt%rcfile; cache(nused)%nch = mt%nch; cache(nused)%nMom = mt%nMom
      cache(nused)%channels = 0.; cache(nused)%channels(1:n) = mt%channels(1:n)
   end function

   ! MieObs_py.F90 entry points get the rc file name: the tables are read by the reference's Chem_MieCreate the first time only
   function ab200_handle_for_rc(rcfile, rc) result(h)
      character(len=*), intent(in) :: rcfile
      integer, intent(out) :: rc
      type(c_ptr) :: h
      type(Chem_Mie) :: mt
      integer :: i
      rc = 0
      do i = 1, nused
         if (trim(cache(i)%rcfile) == trim(rcfile)) then
            h = cache(i)%h; return
         end if
      end do
      mt = Chem_MieCreate(rcfile, rc)
      if (rc /= 0) then
         print *, 'Cannot create Mie tables from '//trim(rcfile)
         h = c_null_ptr; return
      end if
      mt%rcfile = rcfile
      h = ab200_handle_for(mt, rc)
      call Chem_MieDestroy(mt, i)
   end function
end module ab200_mie_handles

! ------------------------------------------------------------------ Components/rtms/geo_vlidort/VLIDORT_MieMod.F90
subroutine VLIDORT_getAOPscalar(mieTables, km, nobs, nch, nq, channels, vname, verbose, qm, rh, tau, ssa, g, rc)
   use Chem_MieMod
   use ab200_mie_handles
   use ab200_mie_c_api
   use iso_c_binding
   implicit none
   type(Chem_Mie), intent(inout) :: mieTables
   integer, intent(in)  :: km, nobs, nch, nq, verbose
   real,    intent(in)  :: channels(nch)
   character, intent(in) :: vname(nq,16)
   real,    intent(in)  :: qm(km,nq,nobs), rh(km,nobs)
   real,    intent(out) :: tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs)
   integer, intent(out) :: rc
   type(c_ptr) :: h
   integer :: st
   h = ab200_handle_for(mieTables, rc)
   if (rc /= 0) return
   st = ab200_get_aop_scalar(h, km, nobs, nch, nq, channels, vname, verbose, qm, rh, tau, ssa, g, rc)
   if (st /= 0 .and. rc == 0) rc = st
end subroutine VLIDORT_getAOPscalar

subroutine VLIDORT_getAOPvector(mieTables, km, nobs, nch, nq, channels, vname, verbose, qm, rh, nMom, nPol, tau, ssa, g, pmom, rc)
   use Chem_MieMod
   use ab200_mie_handles
   use ab200_mie_c_api
   use iso_c_binding
   implicit none
   type(Chem_Mie), intent(inout) :: mieTables
   integer, intent(in)  :: km, nobs, nch, nq, verbose, nMom, nPol
   real,    intent(in)  :: channels(nch)
   character, intent(in) :: vname(nq,16)
   real,    intent(in)  :: qm(km,nq,nobs), rh(km,nobs)
   real,    intent(out) :: tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs), pmom(km,nch,nobs,nMom,nPol)
   integer, intent(out) :: rc
   type(c_ptr) :: h
   integer :: st
   h = ab200_handle_for(mieTables, rc)
   if (rc /= 0) return
   st = ab200_get_aop_vector(h, km, nobs, nch, nq, channels, vname, verbose, qm, rh, nMom, nPol, tau, ssa, g, pmom, rc)
   if (st /= 0 .and. rc == 0) rc = st
end subroutine VLIDORT_getAOPvector

! ------------------------------------------------------------------ GMAO_Shared/Chem_Base/MieObs_py.F90 (the f2py module MieObs_)
subroutine getEdgeVars(km, nobs, airdens, delp, ptop, pe, ze, te)
   use ab200_mie_c_api
   implicit none
   integer, intent(in)  :: km, nobs
   real,    intent(in)  :: airdens(km,nobs), delp(km,nobs), ptop
   real,    intent(out) :: pe(km+1,nobs), ze(km+1,nobs), te(km+1,nobs)
   integer :: st
   st = ab200_get_edge_vars(km, nobs, airdens, delp, ptop, pe, ze, te)
end subroutine getEdgeVars

subroutine getAOPscalar(km, nobs, nch, nq, rcfile, channels, vname, verbose, qm, rh, tau, ssa, g, rc)
   use ab200_mie_handles
   use ab200_mie_c_api
   use iso_c_binding
   implicit none
   integer, intent(in)  :: km, nobs, nch, nq, verbose
   character(len=*), intent(in) :: rcfile
   real,    intent(in)  :: channels(nch)
   character, intent(in) :: vname(nq,16)
   real,    intent(in)  :: qm(km,nq,nobs), rh(km,nobs)
   real,    intent(out) :: tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs)
   integer, intent(out) :: rc
   type(c_ptr) :: h
   integer :: st
   h = ab200_handle_for_rc(rcfile, rc)
   if (rc /= 0) return
   st = ab200_get_aop_scalar(h, km, nobs, nch, nq, channels, vname, verbose, qm, rh, tau, ssa, g, rc)
   if (st /= 0 .and. rc == 0) rc = st
end subroutine getAOPscalar

subroutine getAOPvector(km, nobs, nch, nq, rcfile, channels, vname, verbose, qm, rh, nMom, nPol, tau, ssa, g, pmom, rc)
   use ab200_mie_handles
   use ab200_mie_c_api
   use iso_c_binding
   implicit none
   integer, intent(in)  :: km, nobs, nch, nq, verbose, nMom, nPol
   character(len=*), intent(in) :: rcfile
   real,    intent(in)  :: channels(nch)
   character, intent(in) :: vname(nq,16)
   real,    intent(in)  :: qm(km,nq,nobs), rh(km,nobs)
   real,    intent(out) :: tau(km,nch,nobs), ssa(km,nch,nobs), g(km,nch,nobs), pmom(km,nch,nobs,nMom,nPol)
   integer, intent(out) :: rc
   type(c_ptr) :: h
   integer :: st
   h = ab200_handle_for_rc(rcfile, rc)
   if (rc /= 0) return
   st = ab200_get_aop_vector(h, km, nobs, nch, nq, channels, vname, verbose, qm, rh, nMom, nPol, tau, ssa, g, pmom, rc)
   if (st /= 0 .and. rc == 0) rc = st
end subroutine getAOPvector

subroutine getExt(km, nobs, nch, nq, rcfile, channels, vname, verbose, qc, qm, rh, ext, sca, bsc, absc_SFC, absc_TOA, depol, rc)
   use ab200_mie_handles
   use ab200_mie_c_api
   use iso_c_binding
   implicit none
   integer, intent(in)  :: km, nobs, nch, nq, verbose
   character(len=*), intent(in) :: rcfile
   real,    intent(in)  :: channels(nch)
   character, intent(in) :: vname(nq,16)
   real,    intent(in)  :: qc(km,nq,nobs), qm(km,nq,nobs), rh(km,nobs)
   real,    intent(out) :: ext(km,nch,nobs), sca(km,nch,nobs), bsc(km,nch,nobs), absc_SFC(km,nch,nobs), absc_TOA(km,nch,nobs), &
         depol(km,nch,nobs)
   integer, intent(out) :: rc
   type(c_ptr) :: h
   integer :: st
   h = ab200_handle_for_rc(rcfile, rc)
   if (rc /= 0) return
   st = ab200_get_ext(h, km, nobs, nch, nq, channels, vname, verbose, qc, qm, rh, ext, sca, bsc, absc_SFC, absc_TOA, depol, rc)
   if (st /= 0 .and. rc == 0) rc = st
end subroutine getExt

subroutine getInt(km, nobs, nch, nq, rcfile, channels, vname, verbose, qc, qm, rh, vol, area, refr, refi, reff, rc)
   use ab200_mie_handles
   use ab200_mie_c_api
   use iso_c_binding
   implicit none
   integer, intent(in)  :: km, nobs, nch, nq, verbose
   character(len=*), intent(in) :: rcfile
   real,    intent(in)  :: channels(nch)
   character, intent(in) :: vname(nq,16)
   real,    intent(in)  :: qc(km,nq,nobs), qm(km,nq,nobs), rh(km,nobs)
   real,    intent(out) :: vol(km,nch,nobs), area(km,nch,nobs), refr(km,nch,nobs), refi(km,nch,nobs), reff(km,nch,nobs)
   integer, intent(out) :: rc
   type(c_ptr) :: h
   integer :: st
   h = ab200_handle_for_rc(rcfile, rc)
   if (rc /= 0) return
   st = ab200_get_int(h, km, nobs, nch, nq, channels, vname, verbose, qc, qm, rh, vol, area, refr, refi, reff, rc)
   if (st /= 0 .and. rc == 0) rc = st
end subroutine getInt
