/* TEST INFRASTRUCTURE ONLY -- CPU oracle for Stage A (Mie AOP), a plain-C restatement of the
 * reference Fortran.  Build with -ffp-contract=off (no FMA): float32 results are meant to be
 * bit-identical to gfortran's for the same inputs.
 *
 * PARITY UNPINNED: the reference holds no golden vectors or asserting tests for Stage A
 * (CB/ut_Mie.F90 has no assertions), and no Fortran compiler exists in the build
 * image, so this restatement is reviewed line-by-line against the cited Fortran only.
 *
 * Follows (paths relative to /root/reference/src/Shared/GMAO_Shared/Chem_Base/):
 *   Chem_MieTableMod.F90:238-719  Chem_MieTableRead + polint
 *   Chem_MieMod.F90:789-965       Chem_MieQueryByInt
 *   MieObs_py.F90:27-83,87-214,218-456,460-638,642-770
 */
#include "mie_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* polint: CB/Chem_MieTableMod.F90:671-717.  x,y real*8; xWant,yWant,slope,dx real*4. */
static float polint(const double *x, const double *y, long ystride, int n, float xWant) {
  int i = 0, j;
  if ((double)xWant < x[0]) i = 1;
  if ((double)xWant > x[n - 1]) i = n;
  if (i == 0)
    for (j = 1; j <= n; j++)
      if ((double)xWant >= x[j - 1]) i = j;
  float slope, dx;
  if (i == n) slope = 0.f;
  else slope = (float)((y[(long)i * ystride] - y[(long)(i - 1) * ystride]) / (x[i] - x[i - 1]));
  dx = (float)((double)xWant - x[i - 1]);
  float sd = slope * dx;                       /* real*4 product */
  return (float)(y[(long)(i - 1) * ystride] + (double)sd);
}

/* Chem_MieTableRead: CB/Chem_MieTableMod.F90:238-669 (the NetCDF part replaced by orc_mie_raw) */
orc_mie_table *orc_mie_table_read(const orc_mie_raw *raw, int nch, const float *channels, int nmom_keep) {
  orc_mie_table *t = (orc_mie_table *)calloc(1, sizeof(*t));
  const int nrh = raw->nrh, nbin = raw->nbin, nlam = raw->nlam, npol = raw->npol;
  t->nch = nch; t->nrh = nrh; t->nbin = nbin; t->nmom = nmom_keep; t->npol = npol;
  size_t n3 = (size_t)nrh * nch * nbin, n2 = (size_t)nrh * nbin;
  t->rh = (float *)malloc(sizeof(float) * nrh);
  t->bext = (float *)malloc(sizeof(float) * n3); t->bsca = (float *)malloc(sizeof(float) * n3);
  t->bbck = (float *)malloc(sizeof(float) * n3); t->g = (float *)malloc(sizeof(float) * n3);
  t->refr = (float *)malloc(sizeof(float) * n3); t->refi = (float *)malloc(sizeof(float) * n3);
  t->pback = (float *)malloc(sizeof(float) * n3 * npol);
  t->pmom = nmom_keep > 0 ? (float *)malloc(sizeof(float) * n3 * nmom_keep * npol) : NULL;
  t->reff = (float *)malloc(sizeof(float) * n2); t->gf = (float *)malloc(sizeof(float) * n2);
  t->rhop = (float *)malloc(sizeof(float) * n2); t->rhod = (float *)malloc(sizeof(float) * n2);
  t->vol = (float *)malloc(sizeof(float) * n2); t->area = (float *)malloc(sizeof(float) * n2);

  for (int i = 0; i < nrh; i++) t->rh[i] = (float)raw->rh[i];
  for (int j = 0; j < nbin; j++)
    for (int i = 0; i < nrh; i++) {
      size_t ij = i + (size_t)nrh * j;
      double gf = raw->gf ? raw->gf[ij] : (double)-999.f;
      double rhop = raw->rhop ? raw->rhop[ij] : (double)-999.f;
      double rhod = raw->rhop ? raw->rhop[0 + (size_t)nrh * j] : (double)-999.f; /* :455-465 */
      double reff = raw->reff[ij];
      double vol = gf * gf * gf / rhod;                    /* :487 */
      double area = vol / ((double)(4.f / 3.f) * reff);    /* :491, 4./3. is real*4 */
      t->reff[ij] = (float)reff; t->gf[ij] = (float)gf; t->rhop[ij] = (float)rhop;
      t->rhod[ij] = (float)rhod; t->vol[ij] = (float)vol; t->area[ij] = (float)area;
    }
  /* linear interpolation to the requested channels: :558-589 */
  for (int j = 0; j < nbin; j++)
    for (int i = 0; i < nrh; i++)
      for (int n = 0; n < nch; n++) {
        size_t src = (size_t)nlam * (i + (size_t)nrh * j);       /* (:,i,j) */
        size_t dst = i + (size_t)nrh * (n + (size_t)nch * j);    /* (i,n,j) */
        float w = channels[n];
        t->bext[dst] = polint(raw->lambda, raw->bext + src, 1, nlam, w);
        t->bsca[dst] = polint(raw->lambda, raw->bsca + src, 1, nlam, w);
        t->bbck[dst] = polint(raw->lambda, raw->bbck + src, 1, nlam, w);
        t->g[dst] = polint(raw->lambda, raw->g + src, 1, nlam, w);
        t->refr[dst] = polint(raw->lambda, raw->refr + src, 1, nlam, w);
        if (raw->refi) {
          /* refi_table = abs(refi_table) :482 */
          double tmp[64]; double *yy = nlam <= 64 ? tmp : (double *)malloc(sizeof(double) * nlam);
          for (int l = 0; l < nlam; l++) yy[l] = fabs(raw->refi[src + l]);
          t->refi[dst] = polint(raw->lambda, yy, 1, nlam, w);
          if (yy != tmp) free(yy);
        } else t->refi[dst] = -999.f;
        for (int p = 0; p < npol; p++) {
          size_t d4 = dst + n3 * p;
          if (raw->pback) t->pback[d4] = polint(raw->lambda, raw->pback + src + (size_t)nlam * n2 * p, 1, nlam, w);
          else t->pback[d4] = 1.f;
        }
        if (nmom_keep > 0)
          for (int m = 0; m < nmom_keep; m++)
            for (int p = 0; p < npol; p++) {
              size_t s5 = src + (size_t)nlam * n2 * (m + (size_t)raw->nmom * p);
              size_t d5 = dst + n3 * (m + (size_t)nmom_keep * p);
              t->pmom[d5] = polint(raw->lambda, raw->pmom + s5, 1, nlam, w);
            }
      }
  /* RH hash: :612-627 */
  for (int j = 1; j <= ORC_NRHHASH; j++) {
    t->rhi[j - 1] = 1; t->rha[j - 1] = 0.f;   /* reference leaves these undefined if rh(1)>0 */
    for (int i = nrh; i >= 1; i--) {
      if ((j - 1) >= (int)(t->rh[i - 1] * 1000.f)) {
        int ip1 = i + 1;
        t->rhi[j - 1] = i;
        if (ip1 > nrh) t->rha[j - 1] = 0.f;
        else {
          float a = (float)(j - 1) / 1000.f - t->rh[i - 1];
          float b = t->rh[ip1 - 1] - t->rh[i - 1];
          t->rha[j - 1] = a / b;
        }
        break;
      }
    }
  }
  return t;
}

void orc_mie_table_free(orc_mie_table *t) {
  if (!t) return;
  free(t->rh); free(t->bext); free(t->bsca); free(t->bbck); free(t->g); free(t->refr); free(t->refi);
  free(t->pback); free(t->pmom); free(t->reff); free(t->gf); free(t->rhop); free(t->rhod);
  free(t->vol); free(t->area); free(t);
}

/* RH snap of Chem_MieQueryByInt: CB/Chem_MieMod.F90:860-867 (the second rhUse= overwrites the first) */
typedef struct { int irh, irhp1; float arh; } rhsnap;
static rhsnap snap(const orc_mie_table *T, float rh) {
  rhsnap s;
  float rhUse = rh < 0.99f ? rh : 0.99f;     /* min(rh,0.99) */
  int isnap = (int)((rhUse + 0.001f) * 1000.f);
  if (isnap < 1) isnap = 1;
  s.arh = T->rha[isnap - 1];
  s.irh = T->rhi[isnap - 1];
  s.irhp1 = s.irh + 1;
  if (s.irhp1 > T->nrh) s.irhp1 = T->nrh;
  return s;
}
#define I3(T, ch, b) ((size_t)(T)->nrh * ((ch) - 1 + (size_t)(T)->nch * ((b) - 1)))
static inline float lerp_tab(const float *a, size_t base, rhsnap s) {
  float t0 = a[base + s.irh - 1] * (1.f - s.arh);
  float t1 = a[base + s.irhp1 - 1] * s.arh;
  return t0 + t1;
}

/* getAOPscalar: CB/MieObs_py.F90:87-214 */
int orc_get_aop_scalar(int km, int nobs, int nch, int nq, orc_mie_table *const *tab, const int *bin,
                       const int *chidx, const float *qm, const float *rh,
                       float *tau, float *ssa, float *g) {
  size_t nout = (size_t)km * nch * nobs;
  memset(tau, 0, nout * 4); memset(ssa, 0, nout * 4); memset(g, 0, nout * 4);
  for (int iq = 0; iq < nq; iq++) {
    const orc_mie_table *T = tab[iq];
    if (!T) continue;
    for (int i = 0; i < nobs; i++)
      for (int n = 0; n < nch; n++)
        for (int k = 0; k < km; k++) {
          rhsnap s = snap(T, rh[k + (size_t)km * i]);
          size_t b3 = I3(T, chidx[n], bin[iq]);
          float bext = lerp_tab(T->bext, b3, s);
          float bsca = lerp_tab(T->bsca, b3, s);
          float gas = lerp_tab(T->g, b3, s);
          float tau_ = bext * qm[k + (size_t)km * (iq + (size_t)nq * i)];
          float ssa_ = bsca / bext;
          size_t o = k + (size_t)km * (n + (size_t)nch * i);
          tau[o] = tau[o] + tau_;
          float st = ssa_ * tau_;
          ssa[o] = ssa[o] + st;
          float gs = gas * ssa_; gs = gs * tau_;
          g[o] = g[o] + gs;
        }
  }
  for (size_t o = 0; o < nout; o++) if (tau[o] > 0.f) ssa[o] = ssa[o] / tau[o];
  for (size_t o = 0; o < nout; o++) { float d = ssa[o] * tau[o]; if (d > 0.f) g[o] = g[o] / d; }
  return 0;
}

/* getAOPvector: CB/MieObs_py.F90:218-456 */
int orc_get_aop_vector(int km, int nobs, int nch, int nq, orc_mie_table *const *tab, const int *bin,
                       const int *chidx, const float *qm, const float *rh, int nmom, int npol,
                       float *tau, float *ssa, float *g, float *pmom) {
  size_t nout = (size_t)km * nch * nobs;
  memset(tau, 0, nout * 4); memset(ssa, 0, nout * 4); memset(g, 0, nout * 4);
  memset(pmom, 0, nout * 4 * nmom * npol);
  for (int iq = 0; iq < nq; iq++) {
    const orc_mie_table *T = tab[iq];
    if (!T) continue;
    if (nmom > T->nmom) return 666;                               /* :369-373 */
    int npol_, sph;
    if (npol <= T->npol) { npol_ = npol; sph = 0; }               /* :377-388 */
    else if (npol == 6 && T->npol == 4) { npol_ = 4; sph = 1; }
    else return 777;
    size_t n3 = (size_t)T->nrh * T->nch * T->nbin;
    for (int i = 0; i < nobs; i++)
      for (int n = 0; n < nch; n++)
        for (int k = 0; k < km; k++) {
          rhsnap s = snap(T, rh[k + (size_t)km * i]);
          size_t b3 = I3(T, chidx[n], bin[iq]);
          float bext = lerp_tab(T->bext, b3, s);
          float bsca = lerp_tab(T->bsca, b3, s);
          float gas = lerp_tab(T->g, b3, s);
          float tau_ = bext * qm[k + (size_t)km * (iq + (size_t)nq * i)];
          float ssa_ = bsca / bext;
          size_t o = k + (size_t)km * (n + (size_t)nch * i);
          tau[o] = tau[o] + tau_;
          float st = ssa_ * tau_;
          ssa[o] = ssa[o] + st;
          float gs = gas * ssa_; gs = gs * tau_;
          g[o] = g[o] + gs;
          for (int p = 0; p < npol_; p++)
            for (int m = 0; m < nmom; m++) {
              float pm = lerp_tab(T->pmom, b3 + n3 * (m + (size_t)T->nmom * p), s);
              float w = pm * ssa_; w = w * tau_;
              size_t op = o + nout * (m + (size_t)nmom * p);
              pmom[op] = pmom[op] + w;
            }
          if (sph)                                                 /* :419-422 */
            for (int m = 0; m < nmom; m++) {
              pmom[o + nout * (m + (size_t)nmom * 4)] = pmom[o + nout * (m + (size_t)nmom * 0)];
              pmom[o + nout * (m + (size_t)nmom * 5)] = pmom[o + nout * (m + (size_t)nmom * 2)];
            }
        }
  }
  for (size_t o = 0; o < nout; o++) if (tau[o] > 0.f) ssa[o] = ssa[o] / tau[o];
  for (size_t o = 0; o < nout; o++) { float d = ssa[o] * tau[o]; if (d > 0.f) g[o] = g[o] / d; }
  for (size_t o = 0; o < nout; o++) {
    float d = ssa[o] * tau[o];
    if (d > 0.f)
      for (int p = 0; p < npol; p++)
        for (int m = 0; m < nmom; m++) {
          size_t op = o + nout * (m + (size_t)nmom * p);
          pmom[op] = pmom[op] / d;
        }
  }
  return 0;
}

/* getExt: CB/MieObs_py.F90:460-638 */
int orc_get_ext(int km, int nobs, int nch, int nq, orc_mie_table *const *tab, const int *bin,
                const int *chidx, const float *qc, const float *qm, const float *rh,
                float *ext, float *sca, float *bsc, float *absc_sfc, float *absc_toa, float *depol) {
  size_t nout = (size_t)km * nch * nobs;
  float *tau = (float *)calloc(nout, 4), *depol_ = (float *)calloc(nout, 4);
  memset(ext, 0, nout * 4); memset(sca, 0, nout * 4); memset(bsc, 0, nout * 4);
  memset(absc_sfc, 0, nout * 4); memset(absc_toa, 0, nout * 4); memset(depol, 0, nout * 4);
  for (int iq = 0; iq < nq; iq++) {
    const orc_mie_table *T = tab[iq];
    if (!T) continue;
    size_t n3 = (size_t)T->nrh * T->nch * T->nbin;
    for (int i = 0; i < nobs; i++)
      for (int n = 0; n < nch; n++)
        for (int k = 0; k < km; k++) {
          rhsnap s = snap(T, rh[k + (size_t)km * i]);
          size_t b3 = I3(T, chidx[n], bin[iq]);
          float bext = lerp_tab(T->bext, b3, s);
          float bsca = lerp_tab(T->bsca, b3, s);
          float bbck = lerp_tab(T->bbck, b3, s);
          float qcv = qc[k + (size_t)km * (iq + (size_t)nq * i)];
          float qmv = qm[k + (size_t)km * (iq + (size_t)nq * i)];
          float ext_ = bext * qcv;
          float ssa_ = bsca / bext;
          size_t o = k + (size_t)km * (n + (size_t)nch * i);
          ext[o] = ext[o] + ext_;
          float t1 = ssa_ * ext_; sca[o] = sca[o] + t1;
          float t2 = bbck * qcv; bsc[o] = bsc[o] + t2;
          float tau_ = bext * qmv;
          float p11 = lerp_tab(T->pback, b3 + n3 * 0, s);
          /* pback(...,5); out of bounds in the reference when the table has nPol=4 -> defined here as P22:=P11 */
          float p22 = lerp_tab(T->pback, b3 + (T->npol >= 5 ? n3 * 4 : 0), s);
          tau[o] = tau[o] + tau_;
          float d1 = (p11 - p22) * ssa_; d1 = d1 * tau_;
          float d2 = (p11 + p22) * ssa_; d2 = d2 * tau_;
          depol[o] = depol[o] + d1;
          depol_[o] = depol_[o] + d2;
        }
  }
  for (size_t o = 0; o < nout; o++) depol[o] = depol_[o] > 0.f ? depol[o] / depol_[o] : 0.f;
  for (size_t o = 0; o < nout; o++) { bsc[o] *= 1e03f; ext[o] *= 1e03f; sca[o] *= 1e03f; }
  for (int i = 0; i < nobs; i++)
    for (int n = 0; n < nch; n++) {
      size_t c = (size_t)km * (n + (size_t)nch * i);
      absc_toa[c] = bsc[c] * expf(-tau[c]);                       /* :606 */
      for (int k = 2; k <= km; k++) {
        float taulev = 0.f;
        for (int l = 1; l <= k - 1; l++) taulev = taulev + tau[c + l - 1];
        taulev = taulev + 0.5f * tau[c + k - 1];
        absc_toa[c + k - 1] = bsc[c + k - 1] * expf(-2.f * taulev);
      }
      absc_sfc[c + km - 1] = bsc[c + km - 1] * expf(-tau[c + km - 1]);
      for (int k = km - 1; k >= 1; k--) {
        float taulev = 0.f;
        for (int l = km; l >= k + 1; l--) taulev = taulev + tau[c + l - 1];
        taulev = taulev + 0.5f * tau[c + k - 1];
        absc_sfc[c + k - 1] = bsc[c + k - 1] * expf(-2.f * taulev);
      }
    }
  free(tau); free(depol_);
  return 0;
}

/* getInt: CB/MieObs_py.F90:642-770 */
int orc_get_int(int km, int nobs, int nch, int nq, orc_mie_table *const *tab, const int *bin,
                const int *chidx, const float *qc, const float *qm, const float *rh,
                float *vol, float *area, float *refr, float *refi, float *reff) {
  (void)qm;
  size_t nout = (size_t)km * nch * nobs;
  memset(vol, 0, nout * 4); memset(area, 0, nout * 4); memset(refr, 0, nout * 4);
  memset(refi, 0, nout * 4); memset(reff, 0, nout * 4);
  for (int iq = 0; iq < nq; iq++) {
    const orc_mie_table *T = tab[iq];
    if (!T) continue;
    for (int i = 0; i < nobs; i++)
      for (int n = 0; n < nch; n++)
        for (int k = 0; k < km; k++) {
          rhsnap s = snap(T, rh[k + (size_t)km * i]);
          size_t b3 = I3(T, chidx[n], bin[iq]);
          size_t b2 = (size_t)T->nrh * (bin[iq] - 1);
          float vol_ = lerp_tab(T->vol, b2, s);
          float area_ = lerp_tab(T->area, b2, s);
          float refr_ = lerp_tab(T->refr, b3, s);
          float refi_ = lerp_tab(T->refi, b3, s);
          float qcv = qc[k + (size_t)km * (iq + (size_t)nq * i)];
          size_t o = k + (size_t)km * (n + (size_t)nch * i);
          float a = vol_ * qcv; vol[o] = vol[o] + a;
          float b = area_ * qcv; area[o] = area[o] + b;
          float c = refr_ * vol_; c = c * qcv; refr[o] = refr[o] + c;
          float d = refi_ * vol_; d = d * qcv; refi[o] = refi[o] + d;
        }
  }
  for (size_t o = 0; o < nout; o++) { float t = (3.f / 4.f) * vol[o]; reff[o] = t / area[o]; }
  for (size_t o = 0; o < nout; o++) if (vol[o] > 0.f) { refr[o] = refr[o] / vol[o]; refi[o] = refi[o] / vol[o]; }
  return 0;
}

/* getEdgeVars: CB/MieObs_py.F90:27-83 */
void orc_get_edge_vars(int km, int nobs, const float *airdens, const float *delp, float ptop,
                       float *pe, float *ze, float *te) {
  const float grav = 9.80616f, Rgas = 287.f, kappa = 2.0f / 7.0f;
  float *pm = (float *)malloc(4 * km), *tm = (float *)malloc(4 * km);
  for (int n = 0; n < nobs; n++) {
    float *pen = pe + (size_t)(km + 1) * n, *zen = ze + (size_t)(km + 1) * n, *ten = te + (size_t)(km + 1) * n;
    const float *ad = airdens + (size_t)km * n, *dp = delp + (size_t)km * n;
    pen[0] = ptop;
    for (int k = 0; k < km; k++) pen[k + 1] = pen[k] + dp[k];
    zen[km] = 0.f;
    for (int k = km - 1; k >= 0; k--) { float d = ad[k] * grav; zen[k] = zen[k + 1] + dp[k] / d; }
    for (int k = 0; k < km; k++) { pm[k] = (pen[k] + pen[k + 1]) / 2.0f; float d = ad[k] * Rgas; tm[k] = pm[k] / d; }
    ten[0] = tm[0];
    for (int k = 1; k < km; k++) {
      float alpha = logf(pen[k] / pm[k - 1]) / logf(pm[k] / pm[k - 1]);
      float d = alpha * (tm[k] - tm[k - 1]);
      ten[k] = tm[k - 1] + d;
    }
    ten[km] = tm[km - 1] * powf(pen[km] / pm[km - 1], kappa);
  }
  free(pm); free(tm);
}
