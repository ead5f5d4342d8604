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
