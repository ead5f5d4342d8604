/* TEST INFRASTRUCTURE ONLY -- CPU oracle for Stage A (Mie AOP).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  The product path (aeroapps_b200/csrc) never links it.
 */
#ifndef MIE_ORACLE_H
#define MIE_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NRHHASH 991

/* A Chem_MieTable after Chem_MieTableRead (CB/Chem_MieTableMod.F90:55-83):
 * all arrays float32, Fortran order with irh fastest. */
typedef struct {
  int nch, nrh, nbin, nmom, npol;
  float *rh;     /* (nrh) */
  float *bext, *bsca, *bbck, *g, *refr, *refi; /* (nrh,nch,nbin) */
  float *pback;  /* (nrh,nch,nbin,npol) */
  float *pmom;   /* (nrh,nch,nbin,nmom,npol) or NULL */
  float *reff, *gf, *rhop, *rhod, *vol, *area; /* (nrh,nbin) */
  int   rhi[ORC_NRHHASH];
  float rha[ORC_NRHHASH];
} orc_mie_table;

/* raw optics file contents, double, Fortran order as read by nf_get_var_double */
typedef struct {
  int nlam, nrh, nbin, nmom, npol;
  const double *lambda;  /* (nlam) [m] */
  const double *rh;      /* (nrh) */
  const double *reff;    /* (nrh,nbin) */
  const double *bext, *bsca, *bbck, *g, *refr, *refi; /* (nlam,nrh,nbin) */
  const double *pback;   /* (nlam,nrh,nbin,npol) or NULL -> 1 */
  const double *pmom;    /* (nlam,nrh,nbin,nmom,npol) or NULL */
  const double *gf, *rhop; /* (nrh,nbin), NULL -> -999 */
} orc_mie_raw;

orc_mie_table *orc_mie_table_read(const orc_mie_raw *raw, int nch, const float *channels_m, int nmom_keep);
void orc_mie_table_free(orc_mie_table *t);

/* tab[iq]/bin[iq]: table pointer and 1-based bin of tracer iq, tab[iq]==NULL -> tracer skipped
 * (Chem_MieQueryIdx returned -1).  chidx[n] = 1-based channel index in the tables. */
int orc_get_aop_scalar(int km, int nobs, int nch, int nq, orc_mie_table *const *tab, const int *bin,
                       const int *chidx, const float *qm, const float *rh,
                       float *tau, float *ssa, float *g);
int orc_get_aop_vector(int km, int nobs, int nch, int nq, orc_mie_table *const *tab, const int *bin,
                       const int *chidx, const float *qm, const float *rh, int nmom, int npol,
                       float *tau, float *ssa, float *g, float *pmom);
int orc_get_ext(int km, int nobs, int nch, int nq, orc_mie_table *const *tab, const int *bin,
                const int *chidx, const float *qc, const float *qm, const float *rh,
                float *ext, float *sca, float *bsc, float *absc_sfc, float *absc_toa, float *depol);
int orc_get_int(int km, int nobs, int nch, int nq, orc_mie_table *const *tab, const int *bin,
                const int *chidx, const float *qc, const float *qm, const float *rh,
                float *vol, float *area, float *refr, float *refi, float *reff);
void orc_get_edge_vars(int km, int nobs, const float *airdens, const float *delp, float ptop,
                       float *pe, float *ze, float *te);
#ifdef __cplusplus
}
#endif
#endif
