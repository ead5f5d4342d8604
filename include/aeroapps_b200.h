/* aeroapps_b200 -- C ABI of the B200-native Mie-AOP + VLIDORT hot path.
 *
 * Every entry point replaces one reference Fortran/f2py interface (cited per function, paths relative to
 * the AeroApps tree).  Arrays use the reference's Fortran (column-major) layouts, `float` where the
 * reference has `real`, `double` where it has `real*8`; sizes are plain ints; the caller owns all
 * buffers.  `*_dev` variants take DEVICE pointers and a CUDA stream (void* = cudaStream_t) and do no
 * host<->device copies; the plain variants take HOST pointers, exactly like the Fortran they replace.
 * Return value: 0 = OK, <0 = library failure (see ab200_last_error()), >0 = the reference's own rc.
 */
#ifndef AEROAPPS_B200_H
#define AEROAPPS_B200_H
#ifdef __cplusplus
extern "C" {
#endif

const char *ab200_last_error(void);
int ab200_version(void);
/* device id the library works on (cudaSetDevice is the caller's job; one process per GPU) */
int ab200_device_count(void);

/* ------------------------------------------------------------------ Stage A: Mie tables
 * Species order of Chem_Mie (src/Shared/GMAO_Shared/Chem_Base/Chem_MieMod.F90:65-99). */
enum { AB200_DU = 0, AB200_SS = 1, AB200_BC = 2, AB200_OC = 3, AB200_BRC = 4, AB200_SU = 5, AB200_NI = 6,
       AB200_NSPECIES = 7 };

typedef struct ab200_mie ab200_mie;

/* Chem_MieCreateFromRC (Chem_MieMod.F90:124-411) minus the rc/NetCDF parsing, which stays with the caller:
 * nch/channels_m = n_channels/r_channels [m], nmom = n_moments. Installs the default tracer registry
 * (du001-5, ss001-5, BCphobic/philic, OCphobic/philic, BRCphobic/philic, SO4, NO3an1-3). */
int ab200_mie_create(int nch, const float *channels_m, int nmom, ab200_mie **out);
int ab200_mie_destroy(ab200_mie *h);                       /* Chem_MieDestroy (Chem_MieMod.F90:660-707) */

/* Chem_MieTableRead (Chem_MieTableMod.F90:238-719) on raw optics-file arrays (double, layout of
 * nf_get_var_double = Fortran (lambda,rh,radius[,nMom,nPol])): lambda-interpolation (polint), derived
 * vol/area and the 991-entry RH hash are computed on the GPU.  pback/pmom/gf/rhop may be NULL
 * (reference defaults 1 / absent / -999).  rc 99 when the file holds fewer moments than nmom. */
int ab200_mie_load_table(ab200_mie *h, int species, int nlam, int nrh, int nbin, int nmom_file, int npol,
                         const double *lambda, const double *rh, const double *reff,
                         const double *bext, const double *bsca, const double *bbck, const double *g,
                         const double *refr, const double *refi, const double *pback, const double *pmom,
                         const double *gf, const double *rhop);
/* The same for a table the caller has already read and interpolated to the handle's channels: the contents of the reference's
 * type(Chem_MieTable) (Chem_MieTableMod.F90:55-83) after Chem_MieTableRead, float32, in that type's array orders: rh(nrh);
 * reff,gf,rhop,rhod,vol,area (nrh,nbin); bext,bsca,bbck,g,refr,refi (nrh,nch,nbin); pback (nrh,nch,nbin,nPol);
 * pmom (nrh,nch,nbin,nMom,nPol) or NULL; rhi(991) (1-based), rha(991).  Used by the Fortran shims of VLIDORT_getAOPscalar/vector
 * (GV/VLIDORT_MieMod.F90:10,124), which keep the reference's NetCDF reader.  rc 99 when the handle wants moments and pmom is NULL. */
int ab200_mie_load_table_f32(ab200_mie *h, int species, int nrh, int nbin, int npol, const float *rh, const float *reff,
                             const float *gf, const float *rhop, const float *rhod, const float *vol, const float *area,
                             const float *bext, const float *bsca, const float *bbck, const float *g, const float *refr,
                             const float *refi, const float *pback, const float *pmom, const int *rhi, const float *rha);
/* registry: tracer name -> (species table, 1-based bin); replaces Chem_Registry + the hard-wired map
 * of Chem_MieMod.F90:343-400 */
int ab200_mie_register_tracer(ab200_mie *h, const char *name, int species, int bin);
/* Chem_MieQueryIdx (Chem_MieMod.F90:720-777): returns the 1-based bin or -1; *species gets the table */
int ab200_mie_query_idx(ab200_mie *h, const char *qname, int *species);
/* debug/parity access to the interpolated float32 tables: var in {"rh","bext","bsca","bbck","g","refr",
 * "refi","pback","pmom","reff","gf","rhop","rhod","vol","area","rha"} or "rhi" (int32). Returns element
 * count (copies min(count,cap)). */
long ab200_mie_get_table(ab200_mie *h, int species, const char *var, void *dst, long cap);

/* getAOPscalar (Chem_Base/MieObs_py.F90:87-214) == VLIDORT_getAOPscalar
 * (Components/rtms/geo_vlidort/VLIDORT_MieMod.F90:10-120) with the tables passed in.
 * channels [nm] (nch); vname = character vname(nq,16) column-major; qm(km,nq,nobs), rh(km,nobs);
 * out tau,ssa,g (km,nch,nobs).  rc: 0, 99 (channel not in tables). */
int ab200_get_aop_scalar(ab200_mie *h, int km, int nobs, int nch, int nq, const float *channels,
                         const char *vname, int verbose, const float *qm, const float *rh,
                         float *tau, float *ssa, float *g, int *rc);
/* getAOPvector (MieObs_py.F90:218-456) == VLIDORT_getAOPvector (VLIDORT_MieMod.F90:124-336);
 * extra out pmom(km,nch,nobs,nMom,nPol).  rc: 0, 99, 666 (too few moments), 777 (nPol mismatch). */
int ab200_get_aop_vector(ab200_mie *h, int km, int nobs, int nch, int nq, const float *channels,
                         const char *vname, int verbose, const float *qm, const float *rh,
                         int nmom, int npol, float *tau, float *ssa, float *g, float *pmom, int *rc);
/* getExt (MieObs_py.F90:460-638) */
int ab200_get_ext(ab200_mie *h, int km, int nobs, int nch, int nq, const float *channels, const char *vname,
                  int verbose, const float *qc, const float *qm, const float *rh,
                  float *ext, float *sca, float *bsc, float *absc_sfc, float *absc_toa, float *depol, int *rc);
/* getInt (MieObs_py.F90:642-770) */
int ab200_get_int(ab200_mie *h, int km, int nobs, int nch, int nq, const float *channels, const char *vname,
                  int verbose, const float *qc, const float *qm, const float *rh,
                  float *vol, float *area, float *refr, float *refi, float *reff, int *rc);
/* getEdgeVars (MieObs_py.F90:27-83): airdens,delp (km,nobs) -> pe,ze,te (km+1,nobs) */
int ab200_get_edge_vars(int km, int nobs, const float *airdens, const float *delp, float ptop,
                        float *pe, float *ze, float *te);

/* device-pointer variants (inputs already resident in HBM; launch on `stream`) */
int ab200_get_aop_scalar_dev(ab200_mie *h, int km, int nobs, int nch, int nq, const float *channels_host,
                             const char *vname_host, const float *qm_dev, const float *rh_dev,
                             float *tau_dev, float *ssa_dev, float *g_dev, void *stream, int *rc);
int ab200_get_aop_vector_dev(ab200_mie *h, int km, int nobs, int nch, int nq, const float *channels_host,
                             const char *vname_host, const float *qm_dev, const float *rh_dev,
                             int nmom, int npol, float *tau_dev, float *ssa_dev, float *g_dev,
                             float *pmom_dev, void *stream, int *rc);

int ab200_get_edge_vars_dev(int km, int nobs, const float *airdens_dev, const float *delp_dev, float ptop, float *pe_dev,
                            float *ze_dev, float *te_dev, void *stream);

#include "aeroapps_b200_rt.h"
#include "aeroapps_b200_geom.h"

#ifdef __cplusplus
}
#endif
#endif
