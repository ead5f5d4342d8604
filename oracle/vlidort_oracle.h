/* TEST INFRASTRUCTURE ONLY -- C interface of the Stage-B CPU oracle (see vlidort_oracle.cpp). */
#ifndef VLIDORT_ORACLE_H
#define VLIDORT_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
  int nstokes, nstreams, nlayers, nmoments_input, nfine;
  int do_plane_parallel, do_focorr, do_focorr_nadir, do_focorr_outgoing, do_deltam, do_lambertian;
  int max_fourier_override;      /* <0: VLIDORT's convergence test; >=0: run m = 0..value, no test */
  double accuracy, asymtx_tolerance, flux_factor, earth_radius, albedo, sza, vza, raa;
  const double *height_grid;     /* 0..L [km] */
  const double *deltau, *omega;  /* L */
  const double *greekmat;        /* [(l*L + n)*16 + k], l = 0..nmoments_input, k = GREEKMAT index - 1 */
  const double *brdf_f;          /* [m][q][i][j]  m = 0..2N-1, q = Mueller index 0..15 */
  const double *brdf_f_0;        /* [m][q][i] */
  const double *user_brdf_f;     /* [m][q][j] */
  const double *user_brdf_f_0;   /* [m][q] */
  const double *exactdb;         /* [q] */
} vl_input;
typedef struct {
  double stokes[4], stokes_ss[4], stokes_db[4];
  int nfourier, status;
} vl_output;
int vl_master(const vl_input *in, vl_output *out);
#ifdef __cplusplus
}
#endif
#endif

#ifdef __cplusplus
extern "C" {
#endif
/* AeroApps glue + drop-in wrapper level (V283/VLIDORT_ScatMod.F90, GV/VLIDORT_LAMBERT.F90 ...), Fortran layouts.
 * surface: 0 Lambertian (albedo(nobs,nch)); 1 BRDF Fourier components supplied per pixel p = j*nch+i
 *          (brdf_f [p][2N][16][N][N], brdf_f_0 [p][2N][16][N], user_brdf_f [p][2N][16][N], user_brdf_f_0 [p][2N][16],
 *           exactdb [p][16]); skip[p] != 0 marks pixels whose surface set-up failed.
 * outputs (nobs,nch): radiance, reflectance, Q, U (Q,U may be NULL for scalar); nfourier/status (npix) optional. */
int orc_vlidort_run(int nstokes, int scalar_glue, int km, int nch, int nobs, int nstreams, int plane_parallel,
                    int nMom, int nPol, const double *ROT, const double *depol, const double *alpha,
                    const double *tau, const double *ssa, const double *pmom, const double *he,
                    const double *albedo, const double *sza, const double *raa, const double *vza,
                    const double *flux_factor, double MISSING, int surface, const double *brdf_f,
                    const double *brdf_f_0, const double *user_brdf_f, const double *user_brdf_f_0,
                    const double *exactdb, const unsigned char *skip, int obs_begin, int obs_end,
                    double *radiance, double *reflectance, double *Q, double *U, int *nfourier, int *status);
/* same, with the surface set up from the BRDF wrappers' own inputs: kind 2 Cox-Munk (u10m,v10m (nobs), mr (nch)),
 * 3 GISS Cox-Munk (same), 5 / 6 their OCI variants, 4 MODIS RTLS (kernel_wt (3,nch,nobs) = fiso,fgeo,fvol; param (2,nch,nobs)).
 * BR/BR_Q/BR_U (nobs,nch) receive the wrappers' surface outputs. */
typedef struct {
  int kind;
  const double *u10m, *v10m, *mr, *kernel_wt, *param;
  double *BR, *BR_Q, *BR_U;
  const double *bpdf;            /* kinds 8, 9: BPDFparam (3,nch,nobs) (kind 7 passes it as param) */
} orc_surface;
int orc_vlidort_run_surf(int nstokes, int scalar_glue, int km, int nch, int nobs, int nstreams, int plane_parallel,
                    int nMom, int nPol, const double *ROT, const double *depol, const double *alpha,
                    const double *tau, const double *ssa, const double *pmom, const double *he,
                    const double *albedo, const double *sza, const double *raa, const double *vza,
                    const double *flux_factor, double MISSING, int surface, const double *brdf_f,
                    const double *brdf_f_0, const double *user_brdf_f, const double *user_brdf_f_0,
                    const double *exactdb, const unsigned char *skip, int obs_begin, int obs_end,
                    double *radiance, double *reflectance, double *Q, double *U, int *nfourier, int *status,
                    const orc_surface *sf);
void orc_set_asymtx_tolerance(double tol);   /* default 1e-12 = the reference glue's value */
/* cloud inputs of VLIDORT_Run_Vector_Cloud for the following runs; all NULL clears them (see vlidort_oracle.cpp) */
void orc_set_cloud(const double *tauI, const double *ssaI, const double *pmomI, const double *tauL, const double *ssaL,
                   const double *pmomL, const double *channels);
/* VLIDORT_ROT_CALC (GV/VLIDORT_ROT.F90:19-86 -> V283/VLIDORT_ScatMod.F90:143-281) */
void orc_rot_calc(int km, int nch, int nobs, const double *channels, const double *pe, const double *he,
                  const double *te, double *ROT, double *depol);
#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
extern "C" {
#endif
/* VBRDF supplement as the AeroApps surface glue calls it; see vbrdf_oracle.cpp */
int orc_vbrdf(int nstokes, int nstreams, int nkernels, const int *which, const double *factors, const double *pars,
              const int *lambertian_flag, int nstreams_brdf_in, int do_shadow, double sza, double vza, double raa,
              double *brdf_f, double *brdf_f_0, double *user_brdf_f, double *user_brdf_f_0, double *exactdb,
              double *wsa_out);
#ifdef __cplusplus
}
#endif
