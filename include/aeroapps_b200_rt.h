/* Stage B (VLIDORT) part of the C ABI -- included by aeroapps_b200.h.
 * Every entry point keeps the argument list of the Fortran module procedure / f2py subroutine it replaces
 * (scalars by value, arrays by pointer in the Fortran layout, `logical` as int, rc through a pointer).
 * Per-pixel failures do not set rc: the outputs of that (obs, channel) become MISSING, as in the reference.
 */
#ifndef AEROAPPS_B200_RT_H
#define AEROAPPS_B200_RT_H

/* VLIDORT_ROT_CALC (Components/rtms/geo_vlidort/VLIDORT_ROT.F90:19-86; f2py ROT_CALC,
 * Components/rtms/leo_vlidort/VLIDORT_POLAR_py.F90:9): pe,he,te (km+1,nobs) -> ROT (km,nobs,nch), depol (nch) */
int ab200_rot_calc(int km, int nch, int nobs, const double *channels, const double *pe, const double *he,
                   const double *te, double MISSING, int verbose, double *ROT, double *depol, int *rc);

/* VLIDORT_Scalar_Lambert (geo_vlidort/VLIDORT_LAMBERT.F90:34-168; f2py SCALAR_LAMBERT, VLIDORT_POLAR_py.F90:371) */
int ab200_vlidort_scalar_lambert(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                 int nMom, int nPol, const double *ROT, const double *depol, const double *tau,
                                 const double *ssa, const double *g, const double *pmom, const double *pe,
                                 const double *he, const double *te, const double *albedo, const double *solar_zenith,
                                 const double *relat_azymuth, const double *sensor_zenith, double MISSING, int verbose,
                                 double *radiance_VL, double *reflectance_VL, int *rc);

/* VLIDORT_Vector_Lambert_SingleGeom (VLIDORT_LAMBERT.F90:172-342; f2py VECTOR_LAMBERT, VLIDORT_POLAR_py.F90:457) */
int ab200_vlidort_vector_lambert(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                 int nMom, int nPol, const double *ROT, const double *depol, const double *alpha,
                                 const double *tau, const double *ssa, const double *pmom, const double *pe,
                                 const double *he, const double *te, const double *albedo, const double *solar_zenith,
                                 const double *relat_azymuth, const double *sensor_zenith, const double *flux_factor,
                                 double MISSING, int verbose, double *radiance_VL, double *reflectance_VL, double *Q,
                                 double *U, int *rc);

/* device-pointer variant (inputs/outputs resident in HBM, launched on `stream` = cudaStream_t) */
int ab200_vlidort_vector_lambert_dev(int km, int nch, int nobs, int nstreams, int plane_parallel, int nMom, int nPol,
                                     const double *ROT, const double *depol, const double *alpha, const double *tau,
                                     const double *ssa, const double *pmom, const double *he, const double *albedo,
                                     const double *sza, const double *raa, const double *vza, const double *flux_factor,
                                     double MISSING, double *radiance, double *reflectance, double *Q, double *U,
                                     void *stream, int *rc);
/* algorithmic FLOP counters (SURVEY.md 8(d) formula) and kernel-launch counter since the last reset */
int ab200_vlidort_stats2(double *flops_eigen, double *flops_bvp, long *n_launches);
/* ms_ops: share of ms_bvp spent building the layer interaction operators; flops_exec: flops the default (interaction-principle)
 * boundary-value kernels execute, while flops_bvp is the band-LU count of SURVEY.md 8(d) */
int ab200_vlidort_stats3(double *ms_ops, double *flops_exec);

/* ---- BRDF surfaces.  Surface set-up: RTSI/VLIDORT2p8p3/VLIDORT_SurfaceMod.F90 (CoxMunk :816-1042, GissCoxMunk :586-814,
 * LANDMODIS :1376-1596) + the VBRDF supplement (vsup/vbrdf).  Outputs (nobs,nch).  BR = exact direct-bounce BRDF
 * element (1,1); BR_Q, BR_U = elements (1,2), (1,3) for the GISS and OCI wrappers (BS_DBOUNCE_BRDFUNC(2..3,1,1,1)), 0 for
 * Vector_CX and LandMODIS as in the reference.  MODIS RTLS: BR < 0 -> every output of that pixel is -500. */

/* VLIDORT_Vector_CX_SingleGeom (geo_vlidort/VLIDORT_BRDF_CX.F90:587-762; f2py VECTOR_CX, VLIDORT_POLAR_py.F90:830-902).
 * U10m,V10m (nobs), mr (nch). */
int ab200_vlidort_vector_cx(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                            int nPol, const double *ROT, const double *depol, const double *tau, const double *ssa,
                            const double *pmom, const double *pe, const double *he, const double *te, const double *U10m,
                            const double *V10m, const double *mr, const double *solar_zenith, const double *relat_azymuth,
                            const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                            double *reflectance_VL_SURF, double *BR, double *Q, double *U, double *BR_Q, double *BR_U,
                            int *rc);
/* VLIDORT_Vector_GissCX (VLIDORT_BRDF_CX.F90:1720-1894; f2py VECTOR_GissCX, VLIDORT_POLAR_py.F90:673-745) */
int ab200_vlidort_vector_gisscx(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                int nMom, int nPol, const double *ROT, const double *depol, const double *tau,
                                const double *ssa, const double *pmom, const double *pe, const double *he, const double *te,
                                const double *U10m, const double *V10m, const double *mr, const double *solar_zenith,
                                const double *relat_azymuth, const double *sensor_zenith, double MISSING, int verbose,
                                double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q, double *U,
                                double *BR_Q, double *BR_U, int *rc);
/* VLIDORT_Vector_OCICX (VLIDORT_BRDF_CX.F90:241-416; f2py VECTOR_OCICX, VLIDORT_POLAR_py.F90:903-974): Cox-Munk with the OCI
 * slope variance sigma2 = 0.00534 W (VLIDORT_SurfaceMod.F90:82-308) */
int ab200_vlidort_vector_ocicx(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                               int nPol, const double *ROT, const double *depol, const double *tau, const double *ssa,
                               const double *pmom, const double *pe, const double *he, const double *te, const double *U10m,
                               const double *V10m, const double *mr, const double *solar_zenith, const double *relat_azymuth,
                               const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                               double *reflectance_VL_SURF, double *BR, double *Q, double *U, double *BR_Q, double *BR_U,
                               int *rc);
/* VLIDORT_Vector_OCIGissCX (VLIDORT_BRDF_CX.F90:54-236; f2py VECTOR_OCIGissCX, VLIDORT_POLAR_py.F90:976-1053): GISS Cox-Munk
 * with the OCI slope variance (VLIDORT_SurfaceMod.F90:310-538), trace-gas absorption alpha (km,nobs,nch) and
 * flux_factor (nch,nobs) */
int ab200_vlidort_vector_ocigisscx(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                   int nMom, int nPol, const double *ROT, const double *depol, const double *alpha,
                                   const double *tau, const double *ssa, const double *pmom, const double *pe,
                                   const double *he, const double *te, const double *U10m, const double *V10m,
                                   const double *mr, const double *solar_zenith, const double *relat_azymuth,
                                   const double *sensor_zenith, const double *flux_factor, double MISSING, int verbose,
                                   double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q, double *U,
                                   double *BR_Q, double *BR_U, int *rc);
/* VLIDORT_Scalar_GissCX (VLIDORT_BRDF_CX.F90:1553-1716); BRDF (1,nobs,nch) */
int ab200_vlidort_scalar_gisscx(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                int nMom, int nPol, const double *ROT, const double *depol, const double *tau,
                                const double *ssa, const double *g, const double *pmom, const double *pe, const double *he,
                                const double *te, const double *U10m, const double *V10m, const double *mr,
                                const double *solar_zenith, const double *relat_azymuth, const double *sensor_zenith,
                                double MISSING, int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF,
                                double *BRDF, int *rc);
/* VLIDORT_Scalar_CX (VLIDORT_BRDF_CX.F90:420-584); BRDF (1,nobs,nch) */
int ab200_vlidort_scalar_cx(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                            int nPol, const double *ROT, const double *depol, const double *tau, const double *ssa,
                            const double *g, const double *pmom, const double *pe, const double *he, const double *te,
                            const double *U10m, const double *V10m, const double *mr, const double *solar_zenith,
                            const double *relat_azymuth, const double *sensor_zenith, double MISSING, int verbose,
                            double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BRDF, int *rc);
/* VLIDORT_Vector_LandMODIS_SingleGeom (geo_vlidort/VLIDORT_BRDF_MODIS.F90:206-410; f2py VECTOR_BRDF_MODIS,
 * VLIDORT_POLAR_py.F90:137-213).  kernel_wt (nkernel=3,nch,nobs) = fiso,fgeo,fvol; param (nparam=2,nch,nobs). */
int ab200_vlidort_vector_brdf_modis(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                    int nMom, int nPol, int nkernel, int nparam, const double *ROT, const double *depol,
                                    const double *tau, const double *ssa, const double *pmom, const double *pe,
                                    const double *he, const double *te, const double *kernel_wt, const double *param,
                                    const double *solar_zenith, const double *relat_azymuth, const double *sensor_zenith,
                                    double MISSING, int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF,
                                    double *BR, double *Q, double *U, double *BR_Q, double *BR_U, int *rc);
/* VLIDORT_Vector_BPDF (GV/VLIDORT_BRDF_MODIS_BPDF.F90:459-643; f2py VECTOR_BPDF, LV/VLIDORT_POLAR_py.F90:295-368): polarised land surface,
 * one BPDF-NDVI kernel (VLIDORT_BPDF, V283/VLIDORT_SurfaceMod.F90:2279; kernel index 14, 50 azimuth streams).  BPDFparam (nparam = 3,nch,nobs) =
 * refractive index, NDVI, scaling factor C.  BR_Q, BR_U = elements (1,2), (1,3) of the exact direct-bounce BRDF; BR < 0 -> all outputs -500. */
int ab200_vlidort_vector_bpdf(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom, int nPol,
                              int nparam, const double *ROT, const double *depol, const double *tau, const double *ssa,
                              const double *pmom, const double *pe, const double *he, const double *te, const double *BPDFparam,
                              const double *solar_zenith, const double *relat_azymuth, const double *sensor_zenith, double MISSING,
                              int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q, double *U,
                              double *BR_Q, double *BR_U, int *rc);
/* VLIDORT_Vector_LandMODIS_BPDF (GV/VLIDORT_BRDF_MODIS_BPDF.F90:23-230; f2py VECTOR_BRDF_MODIS_BPDF, LV/VLIDORT_POLAR_py.F90:214-292): the MODIS
 * RTLS kernels (kernel_wt (3,nch,nobs), RTLSparam (2,nch,nobs)) plus the BPDF-NDVI kernel (BPDFparam (3,nch,nobs)); VLIDORT_LANDMODIS_BPDF,
 * V283/VLIDORT_SurfaceMod.F90:1826-2060 */
int ab200_vlidort_vector_brdf_modis_bpdf(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                         int nPol, int nkernel, int nparam, int nparam_bpdf, const double *ROT, const double *depol,
                                         const double *tau, const double *ssa, const double *pmom, const double *pe, const double *he,
                                         const double *te, const double *kernel_wt, const double *RTLSparam, const double *BPDFparam,
                                         const double *solar_zenith, const double *relat_azymuth, const double *sensor_zenith,
                                         double MISSING, int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR,
                                         double *Q, double *U, double *BR_Q, double *BR_U, int *rc);
/* VLIDORT_Vector_Lambert_BPDF (GV/VLIDORT_LAMBERT_BPDF.F90:23-223; f2py VECTOR_LAMBERT_BPDF, LV/VLIDORT_POLAR_py.F90:528-603): Lambertian
 * albedo (nobs,nch) plus the BPDF-NDVI kernel; VLIDORT_LAMBERTIAN_BPDF, V283/VLIDORT_SurfaceMod.F90:2065-2275.  A MISSING albedo makes
 * the pixel MISSING. */
int ab200_vlidort_vector_lambert_bpdf(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                      int nPol, int nparam, const double *ROT, const double *depol, const double *tau, const double *ssa,
                                      const double *pmom, const double *pe, const double *he, const double *te, const double *albedo,
                                      const double *BPDFparam, const double *solar_zenith, const double *relat_azymuth,
                                      const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                                      double *reflectance_VL_SURF, double *BR, double *Q, double *U, double *BR_Q, double *BR_U, int *rc);
/* VLIDORT_Scalar_LandMODIS (VLIDORT_BRDF_MODIS.F90:35-202) */
int ab200_vlidort_scalar_brdf_modis(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                    int nMom, int nPol, const double *ROT, const double *depol, const double *tau,
                                    const double *ssa, const double *g, const double *pmom, const double *pe,
                                    const double *he, const double *te, const double *kernel_wt, const double *param,
                                    const double *solar_zenith, const double *relat_azymuth, const double *sensor_zenith,
                                    double MISSING, int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF,
                                    double *BR, int *rc);
/* ---- cloud variants: the entry points whose reference wrapper calls VLIDORT_Run_Vector_Cloud
 * (V283/VLIDORT_ScatMod.F90:1865-2220): ice (tauI, ssaI, pmomI) and liquid (tauL, ssaL, pmomL) cloud layers are added to the extinction
 * and scattering optical depth and their P11 moments to the Greek matrix (only the first element of pmomI / pmomL is used, :2125-2131);
 * the Rayleigh depolarisation ratio is recomputed from the wavelength (:2031-2032), the depol argument is not used.
 * tauI, ssaI, tauL, ssaL (km,nch,nobs); pmomI, pmomL (km,nch,nobs,nMom,nPol).
 * VLIDORT_Vector_GissCx_Cloud (GV/VLIDORT_BRDF_CX.F90:2276-2467) */
int ab200_vlidort_vector_gisscx_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                      int nPol, const double *ROT, const double *depol, const double *tau, const double *ssa,
                                      const double *pmom, const double *tauI, const double *ssaI, const double *pmomI,
                                      const double *tauL, const double *ssaL, const double *pmomL, const double *pe, const double *he,
                                      const double *te, const double *U10m, const double *V10m, const double *mr,
                                      const double *solar_zenith, const double *relat_azymuth, const double *sensor_zenith,
                                      double MISSING, int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR,
                                      double *Q, double *U, double *BR_Q, double *BR_U, int *rc);
/* VLIDORT_Vector_LandMODIS_Cloud_SingleGeom (GV/VLIDORT_BRDF_MODIS.F90:825-1053; f2py VECTOR_BRDF_MODIS_CLOUD,
 * LV/VLIDORT_POLAR_py.F90:43-134): with trace-gas absorption alpha (km,nobs,nch) and solar flux flux_factor (nch,nobs) */
int ab200_vlidort_vector_brdf_modis_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                          int nPol, int nkernel, int nparam, const double *ROT, const double *depol, const double *alpha,
                                          const double *tau, const double *ssa, const double *pmom, const double *tauI,
                                          const double *ssaI, const double *pmomI, const double *tauL, const double *ssaL,
                                          const double *pmomL, const double *pe, const double *he, const double *te,
                                          const double *kernel_wt, const double *param, const double *solar_zenith,
                                          const double *relat_azymuth, const double *sensor_zenith, const double *flux_factor,
                                          double MISSING, int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR,
                                          double *Q, double *U, double *BR_Q, double *BR_U, int *rc);
/* VLIDORT_Vector_Lambert_Cloud_SingleGeom (GV/VLIDORT_LAMBERT.F90:688-879; f2py VECTOR_LAMBERT_CLOUD, LV/VLIDORT_POLAR_py.F90:371) and
 * VLIDORT_Vector_CX_Cloud_SingleGeom (GV/VLIDORT_BRDF_CX.F90:951-1144; f2py VECTOR_CX_CLOUD, :746).  Both reference wrappers accept the
 * cloud arrays and then call the clear-sky runner VLIDORT_Run_Vector (VLIDORT_LAMBERT.F90:843, VLIDORT_BRDF_CX.F90:1118): the cloud
 * arrays do not enter the result there, nor here (they may be NULL). */
int ab200_vlidort_vector_lambert_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                       int nPol, const double *ROT, const double *depol, const double *alpha, const double *tau,
                                       const double *ssa, const double *pmom, const double *tauI, const double *ssaI, const double *pmomI,
                                       const double *tauL, const double *ssaL, const double *pmomL, const double *pe, const double *he,
                                       const double *te, const double *albedo, const double *solar_zenith, const double *relat_azymuth,
                                       const double *sensor_zenith, const double *flux_factor, double MISSING, int verbose,
                                       double *radiance_VL, double *reflectance_VL, double *Q, double *U, int *rc);
int ab200_vlidort_vector_cx_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom, int nPol,
                                  const double *ROT, const double *depol, const double *tau, const double *ssa, const double *pmom,
                                  const double *tauI, const double *ssaI, const double *pmomI, const double *tauL, const double *ssaL,
                                  const double *pmomL, const double *pe, const double *he, const double *te, const double *U10m,
                                  const double *V10m, const double *mr, const double *solar_zenith, const double *relat_azymuth,
                                  const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                                  double *reflectance_VL_SURF, double *BR, double *Q, double *U, double *BR_Q, double *BR_U, int *rc);
/* scalar forms: VLIDORT_Scalar_GissCX_Cloud (GV/VLIDORT_BRDF_CX.F90:1898-2083), VLIDORT_Scalar_LandMODIS_Cloud
 * (GV/VLIDORT_BRDF_MODIS.F90:625-823) -> VLIDORT_Run_Scalar_Cloud (V283/VLIDORT_ScatMod.F90:2596-2891).  g, gI, gL (km,nch,nobs) are part of
 * the reference signatures and, as there, do not enter the result */
int ab200_vlidort_scalar_gisscx_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                      int nPol, const double *ROT, const double *depol, const double *tau, const double *ssa,
                                      const double *g, const double *pmom, const double *tauI, const double *ssaI, const double *gI,
                                      const double *pmomI, const double *tauL, const double *ssaL, const double *gL, const double *pmomL,
                                      const double *pe, const double *he, const double *te, const double *U10m, const double *V10m,
                                      const double *mr, const double *solar_zenith, const double *relat_azymuth,
                                      const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                                      double *reflectance_VL_SURF, double *BRDF, int *rc);
int ab200_vlidort_scalar_brdf_modis_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                          int nPol, const double *ROT, const double *depol, const double *alpha, const double *tau,
                                          const double *ssa, const double *g, const double *pmom, const double *tauI, const double *ssaI,
                                          const double *gI, const double *pmomI, const double *tauL, const double *ssaL, const double *gL,
                                          const double *pmomL, const double *pe, const double *he, const double *te,
                                          const double *kernel_wt, const double *param, const double *solar_zenith,
                                          const double *relat_azymuth, const double *sensor_zenith, const double *flux_factor,
                                          double MISSING, int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR,
                                          int *rc);
/* ab200_vlidort_vector_cx with every array (inputs and outputs) resident in device memory; stream = cudaStream_t */
int ab200_vlidort_vector_cx_dev(int km, int nch, int nobs, int nstreams, int plane_parallel, int nMom, int nPol,
                                const double *ROT, const double *depol, const double *tau, const double *ssa,
                                const double *pmom, const double *he, const double *U10m, const double *V10m,
                                const double *mr, const double *sza, const double *raa, const double *vza, double MISSING,
                                double *radiance, double *reflectance, double *BR, double *Q, double *U, void *stream,
                                int *rc);


/* ab200_rot_calc with every array resident in device memory */
int ab200_rot_calc_dev(int km, int nch, int nobs, const double *channels, const double *pe, const double *he, const double *te,
                       double *ROT, double *depol, void *stream);

/* Chained column pipeline (four stages back to back on one stream, not one fused kernel): MieObs_.getEdgeVars -> ROT_CALC -> MieObs_.getAOPvector -> VECTOR_CX, the call
 * sequence of leo_vlidort / polar_vlidort (LV/polar_vlidort.py:600-950, LV/py_leo_vlidort/vlidort.py:400-580), in one call.
 * Host inputs are the model state of the Mie side (qm (km,nq,nobs), rh / airdens / delp (km,nobs) float32, tracer names), the
 * wavelength list channels [nm] float64 (the Mie channel match gets its float32 copy, as f2py makes for MieObs_) and the geometry / wind / refractive index of the RT side (float64); edge variables, Rayleigh
 * optical thickness, tau, ssa and the phase-matrix moments never leave HBM.  Results are bit-identical to the four calls. */
int ab200_pipeline_vector_cx(ab200_mie *h, int km, int nobs, int nch, int nq, const double *channels, const char *vname,
                             const float *qm, const float *rh, const float *airdens, const float *delp, float ptop,
                             int nstreams, int plane_parallel, int nMom, int nPol, const double *U10m, const double *V10m,
                             const double *mr, const double *solar_zenith, const double *relat_azymuth,
                             const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                             double *reflectance_VL_SURF, double *BR, double *Q, double *U, double *BR_Q, double *BR_U, int *rc);

/* ---- multi-geometry forms: the generic interfaces VLIDORT_Vector_Lambert / VLIDORT_Vector_CX / VLIDORT_Vector_LandMODIS over
 * SingleGeom / MultiGeom (GV/VLIDORT_LAMBERT.F90:14-17,346-350; GV/VLIDORT_BRDF_CX.F90:27-30,766; GV/VLIDORT_BRDF_MODIS.F90:413).
 * solar_zenith, relat_azymuth, sensor_zenith (nobs,ngeom); outputs (nobs,nch,ngeom); albedo (nobs,nch) = the reference's
 * albedo(:,:,1) (VLIDORT_SurfaceLamb_MultiGeom uses the first geometry's albedo, V283/VLIDORT_SurfaceMod.F90:564-582).  ngeom is not
 * capped at the reference's MAX_USER_OBSGEOMS = 2.  One sun per observation (LV/polar_vlidort.py:935-950: 60 view angles per column):
 * eigenproblems and boundary-value solve once per pixel and Fourier term, per-view post-processing and Fourier convergence;
 * geometries with their own suns are solved one geometry at a time. */
int ab200_vlidort_vector_lambert_multigeom(int km, int nch, int nobs, int ngeom, const double *channels, int nstreams,
                                           int plane_parallel, int nMom, int nPol, const double *ROT, const double *depol,
                                           const double *alpha, const double *tau, const double *ssa, const double *pmom,
                                           const double *pe, const double *he, const double *te, const double *albedo,
                                           const double *solar_zenith, const double *relat_azymuth, const double *sensor_zenith,
                                           const double *flux_factor, double MISSING, int verbose, double *radiance_VL,
                                           double *reflectance_VL, double *Q, double *U, int *rc);
int ab200_vlidort_vector_cx_multigeom(int km, int nch, int nobs, int ngeom, const double *channels, int nstreams, int plane_parallel,
                                      int nMom, int nPol, const double *ROT, const double *depol, const double *tau, const double *ssa,
                                      const double *pmom, const double *pe, const double *he, const double *te, const double *U10m,
                                      const double *V10m, const double *mr, const double *solar_zenith, const double *relat_azymuth,
                                      const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                                      double *reflectance_VL_SURF, double *BR, double *Q, double *U, double *BR_Q, double *BR_U, int *rc);
int ab200_vlidort_vector_brdf_modis_multigeom(int km, int nch, int nobs, int ngeom, const double *channels, int nstreams,
                                              int plane_parallel, int nMom, int nPol, int nkernel, int nparam, const double *ROT,
                                              const double *depol, const double *tau, const double *ssa, const double *pmom,
                                              const double *pe, const double *he, const double *te, const double *kernel_wt,
                                              const double *param, const double *solar_zenith, const double *relat_azymuth,
                                              const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                                              double *reflectance_VL_SURF, double *BR, double *Q, double *U, double *BR_Q,
                                              double *BR_U, int *rc);
/* ---- the remaining cloud forms of the reference modules (Fortran module procedures without an f2py wrapper).
 * VLIDORT_Scalar_Lambert_Cloud (GV/VLIDORT_LAMBERT.F90:524-686): clear-sky scalar runner (:663) with alpha (km,nobs,nch) and flux_factor
 * (nch,nobs); the cloud arrays and g, gI, gL do not enter (they may be NULL) */
int ab200_vlidort_scalar_lambert_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                       int nPol, const double *ROT, const double *depol, const double *alpha, const double *tau,
                                       const double *ssa, const double *g, const double *pmom, const double *tauI,
                                       const double *ssaI, const double *gI, const double *pmomI, const double *tauL,
                                       const double *ssaL, const double *gL, const double *pmomL, const double *pe,
                                       const double *he, const double *te, const double *albedo, const double *solar_zenith,
                                       const double *relat_azymuth, const double *sensor_zenith, const double *flux_factor,
                                       double MISSING, int verbose, double *radiance_VL, double *reflectance_VL, int *rc);
/* VLIDORT_Vector_Lambert_Cloud_MultiGeom (GV/VLIDORT_LAMBERT.F90:881-1075) and VLIDORT_Vector_CX_Cloud_MultiGeom (GV/VLIDORT_BRDF_CX.F90:
 * 1147-1348): multi-geometry forms that take the cloud arrays and call the clear-sky runner (:1036, :1320); shapes as for the
 * _multigeom entries above, albedo (nobs,nch) = the reference's albedo(:,:,1) */
int ab200_vlidort_vector_lambert_cloud_multigeom(int km, int nch, int nobs, int ngeom, const double *channels, int nstreams,
                                                 int plane_parallel, int nMom, int nPol, const double *ROT, const double *depol,
                                                 const double *alpha, const double *tau, const double *ssa, const double *pmom,
                                                 const double *tauI, const double *ssaI, const double *pmomI, const double *tauL,
                                                 const double *ssaL, const double *pmomL, const double *pe, const double *he,
                                                 const double *te, const double *albedo, const double *solar_zenith,
                                                 const double *relat_azymuth, const double *sensor_zenith, double MISSING,
                                                 int verbose, double *radiance_VL, double *reflectance_VL, double *Q, double *U,
                                                 int *rc);
int ab200_vlidort_vector_cx_cloud_multigeom(int km, int nch, int nobs, int ngeom, const double *channels, int nstreams,
                                            int plane_parallel, int nMom, int nPol, const double *ROT, const double *depol,
                                            const double *tau, const double *ssa, const double *pmom, const double *tauI,
                                            const double *ssaI, const double *pmomI, const double *tauL, const double *ssaL,
                                            const double *pmomL, const double *pe, const double *he, const double *te,
                                            const double *U10m, const double *V10m, const double *mr, const double *solar_zenith,
                                            const double *relat_azymuth, const double *sensor_zenith, double MISSING, int verbose,
                                            double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q, double *U,
                                            double *BR_Q, double *BR_U, int *rc);
/* VLIDORT_Vector_LandMODIS_Cloud_MultiGeom (GV/VLIDORT_BRDF_MODIS.F90:1055-1268): cloud runner, all geometries of an observation */
int ab200_vlidort_vector_brdf_modis_cloud_multigeom(int km, int nch, int nobs, int ngeom, const double *channels, int nstreams,
                                                    int plane_parallel, int nMom, int nPol, int nkernel, int nparam,
                                                    const double *ROT, const double *depol, const double *alpha, const double *tau,
                                                    const double *ssa, const double *pmom, const double *tauI, const double *ssaI,
                                                    const double *pmomI, const double *tauL, const double *ssaL,
                                                    const double *pmomL, const double *pe, const double *he, const double *te,
                                                    const double *kernel_wt, const double *param, const double *solar_zenith,
                                                    const double *relat_azymuth, const double *sensor_zenith,
                                                    const double *flux_factor, double MISSING, int verbose,
                                                    double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q,
                                                    double *U, double *BR_Q, double *BR_U, int *rc);
/* VLIDORT_Vector_OCIGissCX_Cloud (GV/VLIDORT_BRDF_CX.F90:2468-2668) and VLIDORT_Scalar_OCIGissCX_Cloud (:2084-2273): OCI GISS Cox-Munk
 * set-up, trace-gas absorption, solar flux, cloud runners */
int ab200_vlidort_vector_ocigisscx_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                         int nPol, const double *ROT, const double *depol, const double *alpha, const double *tau,
                                         const double *ssa, const double *pmom, const double *tauI, const double *ssaI,
                                         const double *pmomI, const double *tauL, const double *ssaL, const double *pmomL,
                                         const double *pe, const double *he, const double *te, const double *U10m,
                                         const double *V10m, const double *mr, const double *solar_zenith,
                                         const double *relat_azymuth, const double *sensor_zenith, const double *flux_factor,
                                         double MISSING, int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF,
                                         double *BR, double *Q, double *U, double *BR_Q, double *BR_U, int *rc);
int ab200_vlidort_scalar_ocigisscx_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                         int nPol, const double *ROT, const double *depol, const double *alpha, const double *tau,
                                         const double *ssa, const double *g, const double *pmom, const double *tauI,
                                         const double *ssaI, const double *gI, const double *pmomI, const double *tauL,
                                         const double *ssaL, const double *gL, const double *pmomL, const double *pe,
                                         const double *he, const double *te, const double *U10m, const double *V10m,
                                         const double *mr, const double *solar_zenith, const double *relat_azymuth,
                                         const double *sensor_zenith, const double *flux_factor, double MISSING, int verbose,
                                         double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BRDF, int *rc);
/* VLIDORT_Vector_LandMODIS_BPDF_Cloud (GV/VLIDORT_BRDF_MODIS_BPDF.F90:232-455) and VLIDORT_Vector_Lambert_BPDF_Cloud
 * (GV/VLIDORT_LAMBERT_BPDF.F90:228-440): "place holder" wrappers that call the clear-sky runner (:427, :409); the first differs from
 * VLIDORT_Vector_LandMODIS_BPDF by alpha and flux_factor; the cloud arrays do not enter (they may be NULL) */
int ab200_vlidort_vector_brdf_modis_bpdf_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                               int nMom, int nPol, int nkernel, int nparam, int nparam_bpdf, const double *ROT,
                                               const double *depol, const double *alpha, const double *tau, const double *ssa,
                                               const double *pmom, const double *tauI, const double *ssaI, const double *pmomI,
                                               const double *tauL, const double *ssaL, const double *pmomL, const double *pe,
                                               const double *he, const double *te, const double *kernel_wt, const double *RTLSparam,
                                               const double *BPDFparam, const double *solar_zenith, const double *relat_azymuth,
                                               const double *sensor_zenith, const double *flux_factor, double MISSING, int verbose,
                                               double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q,
                                               double *U, double *BR_Q, double *BR_U, int *rc);
int ab200_vlidort_vector_lambert_bpdf_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                            int nMom, int nPol, int nparam, const double *ROT, const double *depol,
                                            const double *tau, const double *ssa, const double *pmom, const double *tauI,
                                            const double *ssaI, const double *pmomI, const double *tauL, const double *ssaL,
                                            const double *pmomL, const double *pe, const double *he, const double *te,
                                            const double *albedo, const double *BPDFparam, const double *solar_zenith,
                                            const double *relat_azymuth, const double *sensor_zenith, double MISSING, int verbose,
                                            double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q, double *U,
                                            double *BR_Q, double *BR_U, int *rc);
/* ---- NOBM water-leaving forms: VLIDORT_Vector_CX_NOBM, VLIDORT_Scalar_GissCX_NOBM_Cloud, VLIDORT_Scalar_OCIGissCX_NOBM_Cloud,
 * VLIDORT_Vector_GissCX_NOBM_Cloud, VLIDORT_Vector_OCIGissCx_NOBM_Cloud (GV/VLIDORT_BRDF_CX.F90:1351,2671,2883,3102,3320; the last one is
 * f2py VECTOR_OCIGissCX_NOBM_Cloud, LV/VLIDORT_POLAR_py.F90:1055).  sleave (nch,nobs) = sun-normalised water-leaving radiance;
 * sleave_iso / sleave_adjust = the reference's logicals; ADJUSTED_SLEAVE (nobs,nch) may be NULL (optional in the reference).
 * sleave_adjust = 0: VLIDORT 2.8.3 scales every water-leaving term by TRANS_ATMOS_FINAL, which stays 0 without the transmittance
 *   iteration or DO_EXTERNAL_WLEAVE (VM/vlidort_bvproblem.f90:1500-1513,1563-1571; VM/vlidort_masters_V2p8p3.f90:2493-2499) -- the reference's
 *   answer is the one without water-leaving and ADJUSTED_SLEAVE = 0; these entries return exactly that (the sibling entry's result).
 * sleave_adjust = 1 (BVP_ADJUSTED_BACKSUB, VM/vlidort_bvproblem.f90:2487-2910) is NOT supported: AB200_ERR_ARG, rc = -2. */
int ab200_vlidort_vector_cx_nobm(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom, int nPol,
                                 const double *ROT, const double *depol, const double *tau, const double *ssa, const double *pmom,
                                 const double *pe, const double *he, const double *te, const double *U10m, const double *V10m,
                                 const double *mr, const double *sleave, int sleave_iso, int sleave_adjust,
                                 const double *solar_zenith, const double *relat_azymuth, const double *sensor_zenith, double MISSING,
                                 int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q, double *U,
                                 double *BR_Q, double *BR_U, int *rc, double *ADJUSTED_SLEAVE);
int ab200_vlidort_vector_gisscx_nobm_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                           int nMom, int nPol, const double *ROT, const double *depol, const double *tau,
                                           const double *ssa, const double *pmom, const double *tauI, const double *ssaI,
                                           const double *pmomI, const double *tauL, const double *ssaL, const double *pmomL,
                                           const double *pe, const double *he, const double *te, const double *U10m,
                                           const double *V10m, const double *mr, const double *sleave, int sleave_iso,
                                           int sleave_adjust, const double *solar_zenith, const double *relat_azymuth,
                                           const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                                           double *reflectance_VL_SURF, double *BR, double *Q, double *U, double *BR_Q, double *BR_U,
                                           int *rc, double *ADJUSTED_SLEAVE);
int ab200_vlidort_vector_ocigisscx_nobm_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                              int nMom, int nPol, const double *ROT, const double *depol, const double *alpha,
                                              const double *tau, const double *ssa, const double *pmom, const double *tauI,
                                              const double *ssaI, const double *pmomI, const double *tauL, const double *ssaL,
                                              const double *pmomL, const double *pe, const double *he, const double *te,
                                              const double *U10m, const double *V10m, const double *mr, const double *sleave,
                                              int sleave_iso, int sleave_adjust, const double *solar_zenith,
                                              const double *relat_azymuth, const double *sensor_zenith, const double *flux_factor,
                                              double MISSING, int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF,
                                              double *BR, double *Q, double *U, double *BR_Q, double *BR_U, int *rc,
                                              double *ADJUSTED_SLEAVE);
int ab200_vlidort_scalar_gisscx_nobm_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                           int nMom, int nPol, const double *ROT, const double *depol, const double *tau,
                                           const double *ssa, const double *g, const double *pmom, const double *tauI,
                                           const double *ssaI, const double *gI, const double *pmomI, const double *tauL,
                                           const double *ssaL, const double *gL, const double *pmomL, const double *pe,
                                           const double *he, const double *te, const double *U10m, const double *V10m,
                                           const double *mr, const double *sleave, int sleave_iso, int sleave_adjust,
                                           const double *solar_zenith, const double *relat_azymuth, const double *sensor_zenith,
                                           double MISSING, int verbose, double *radiance_VL_SURF, double *reflectance_VL_SURF,
                                           double *BRDF, int *rc, double *ADJUSTED_SLEAVE);
int ab200_vlidort_scalar_ocigisscx_nobm_cloud(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                              int nMom, int nPol, const double *ROT, const double *depol, const double *alpha,
                                              const double *tau, const double *ssa, const double *g, const double *pmom,
                                              const double *tauI, const double *ssaI, const double *gI, const double *pmomI,
                                              const double *tauL, const double *ssaL, const double *gL, const double *pmomL,
                                              const double *pe, const double *he, const double *te, const double *U10m,
                                              const double *V10m, const double *mr, const double *sleave, int sleave_iso,
                                              int sleave_adjust, const double *solar_zenith, const double *relat_azymuth,
                                              const double *sensor_zenith, const double *flux_factor, double MISSING, int verbose,
                                              double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BRDF, int *rc,
                                              double *ADJUSTED_SLEAVE);

/* diagnostics of the calling thread's most recent VLIDORT call: Fourier count and status per pixel p = j*nch + i (per pixel-view
 * (j*nch + i)*ngeom + g for the multi-geometry forms); returns their number */
long ab200_vlidort_last_diag(int *nfourier_out, int *status_out, long cap);
/* accumulated device timings (CUDA events) of the three kernels and problem counts; reset != 0 clears them */
int ab200_vlidort_stats(double *ms_setup, double *ms_eigen, double *ms_bvp, long *n_eigen, long *n_bvp, long *npix,
                        int reset);
/* Multi-GPU inside the library (the reference spreads pixels over MPI ranks, GV/geo_vlidort.F90:482-495): the host-pointer Stage-B
 * entries split their observations into one contiguous slab per device, one host thread per device, each device copying its slab
 * of the results straight into the caller's arrays.  n <= 0: every visible device; default AB200_DEVICES, else 1 (rank-per-GPU
 * mains keep one device per process).  Returns the number of devices that will be used. */
int ab200_set_devices(int n);
/* bytes the host-pointer Stage-B entry points have copied host -> device and device -> host since the last reset (slab arrays; of pmom only
 * the elements and moments the solve reads are uploaded: P11, P12 fully, P33 and P22 up to moment 2 nstreams - 1) */
int ab200_vlidort_io_bytes(double *h2d_bytes, double *d2h_bytes, int reset);
/* self-test of the CTA linear algebra of the boundary-value kernels (tests/test_linalg_gpu.py): E = A^-1 (register-resident
 * Gauss-Jordan with partial pivoting) and P = A E (16-byte-load product) for one n x n row-major matrix, one CTA of `threads`
 * threads.  Returns 0, or 1 when the matrix is singular to working precision. */
int ab200_selftest_linalg(int n, int threads, const double *A, double *E, double *P);
/* frees the grow-only device workspace */
int ab200_vlidort_release(void);
#endif
