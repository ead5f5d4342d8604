/* Plain-C caller of the C ABI (include/aeroapps_b200.h), compiled with gcc -std=c99 and linked against libaeroapps_b200.so only: what a
 * cgo / Fortran / C maintainer binds.  No Python objects cross the boundary: every array is a malloc'ed buffer in the reference's
 * Fortran (column-major) order, vname is character vname(nq,16).
 *
 *   test_cabi <scene.bin> <out.bin>
 * scene.bin (written by tests/test_cabi.py): int32 km,nch,nobs,nstreams,plane_parallel,nMom,nPol ; then float64 arrays
 *   channels(nch) pe(km+1,nobs) he(km+1,nobs) te(km+1,nobs) tau(km,nch,nobs) ssa(km,nch,nobs) g(km,nch,nobs) pmom(km,nch,nobs,nMom,nPol)
 *   albedo(nobs,nch) U10m(nobs) V10m(nobs) mr(nch) sza(nobs) raa(nobs) vza(nobs)
 * out.bin: float64 ROT(km,nobs,nch) depol(nch) ; vector_cx: rad refl BR Q U BR_Q BR_U (nobs,nch each) ; scalar_lambert: rad refl ;
 *   then int32 rc_rot rc_cx rc_sl, int32 nfourier(nobs*nch) status(nobs*nch) of the LAST call (scalar_lambert)
 * Reference interfaces: VLIDORT_ROT.F90:19, VLIDORT_BRDF_CX.F90:588, VLIDORT_LAMBERT.F90:34. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "aeroapps_b200.h"

static double *rd(FILE *f, long n) {
  double *p = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (!p || fread(p, sizeof(double), (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read (%ld doubles)\n", n); exit(2); }
  return p;
}
static double *mk(long n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }
static void wr(FILE *f, const double *p, long n) { if (fwrite(p, sizeof(double), (size_t)n, f) != (size_t)n) exit(3); }

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: test_cabi scene.bin out.bin\n"); return 1; }
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  int hd[7];
  if (fread(hd, sizeof(int), 7, f) != 7) return 2;
  const int km = hd[0], nch = hd[1], nobs = hd[2], nstreams = hd[3], pp = hd[4], nMom = hd[5], nPol = hd[6];
  const long kn = (long)km * nch * nobs, on = (long)nobs * nch;
  double *channels = rd(f, nch), *pe = rd(f, (long)(km + 1) * nobs), *he = rd(f, (long)(km + 1) * nobs), *te = rd(f, (long)(km + 1) * nobs);
  double *tau = rd(f, kn), *ssa = rd(f, kn), *g = rd(f, kn), *pmom = rd(f, kn * nMom * nPol);
  double *albedo = rd(f, on), *u10 = rd(f, nobs), *v10 = rd(f, nobs), *mr = rd(f, nch), *sza = rd(f, nobs), *raa = rd(f, nobs), *vza = rd(f, nobs);
  fclose(f);

  if (ab200_version() < 100) { fprintf(stderr, "bad version\n"); return 4; }
  double *ROT = mk((long)km * nobs * nch), *depol = mk(nch);
  int rc_rot = -1, rc_cx = -1, rc_sl = -1, st;
  st = ab200_rot_calc(km, nch, nobs, channels, pe, he, te, 1e15, 0, ROT, depol, &rc_rot);
  if (st) { fprintf(stderr, "rot_calc: %d %s\n", st, ab200_last_error()); return 5; }
  double *o[9];
  for (int i = 0; i < 9; ++i) o[i] = mk(on);
  st = ab200_vlidort_vector_cx(km, nch, nobs, channels, nstreams, pp, nMom, nPol, ROT, depol, tau, ssa, pmom, pe, he, te, u10, v10, mr, sza, raa, vza,
                               1e15, 0, o[0], o[1], o[2], o[3], o[4], o[5], o[6], &rc_cx);
  if (st) { fprintf(stderr, "vector_cx: %d %s\n", st, ab200_last_error()); return 6; }
  st = ab200_vlidort_scalar_lambert(km, nch, nobs, channels, nstreams, pp, nMom, nPol, ROT, depol, tau, ssa, g, pmom, pe, he, te, albedo, sza, raa,
                                    vza, 1e15, 0, o[7], o[8], &rc_sl);
  if (st) { fprintf(stderr, "scalar_lambert: %d %s\n", st, ab200_last_error()); return 7; }
  int *nf = (int *)calloc((size_t)on, sizeof(int)), *stt = (int *)calloc((size_t)on, sizeof(int));
  if (ab200_vlidort_last_diag(nf, stt, on) != on) { fprintf(stderr, "last_diag count\n"); return 8; }
  /* argument errors must come back as a status, not a crash (a NULL array, a negative size) */
  int rc_bad = 0;
  if (ab200_rot_calc(km, nch, nobs, channels, NULL, he, te, 1e15, 0, ROT, depol, &rc_bad) == 0) { fprintf(stderr, "NULL pe accepted\n"); return 9; }
  if (ab200_vlidort_scalar_lambert(-1, nch, nobs, channels, nstreams, pp, nMom, nPol, ROT, depol, tau, ssa, g, pmom, pe, he, te, albedo, sza, raa,
                                   vza, 1e15, 0, o[7], o[8], &rc_bad) == 0) { fprintf(stderr, "km=-1 accepted\n"); return 9; }

  f = fopen(argv[2], "wb");
  if (!f) { perror(argv[2]); return 1; }
  wr(f, ROT, (long)km * nobs * nch); wr(f, depol, nch);
  for (int i = 0; i < 9; ++i) wr(f, o[i], on);
  int rcs[3] = {rc_rot, rc_cx, rc_sl};
  fwrite(rcs, sizeof(int), 3, f); fwrite(nf, sizeof(int), (size_t)on, f); fwrite(stt, sizeof(int), (size_t)on, f);
  fclose(f);
  ab200_vlidort_release();
  printf("test_cabi ok: km=%d nch=%d nobs=%d nstreams=%d\n", km, nch, nobs, nstreams);
  return 0;
}
