// TEST-ONLY build of the product's Stage-B kernel bodies (aeroapps_b200/csrc/vl_*.cuh) with a one-lane
// SerialTeam, so that their mathematics can be compared with the CPU oracle in the GPU-less build container.
// Not part of the shipped library; the product has no CPU path.
#define AB_HOST_MODEL 1
#include "../../aeroapps_b200/csrc/vl_add.cuh"
#include "../../aeroapps_b200/csrc/vl_brdf.cuh"
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
using namespace vlk;

// ngeom view geometries per observation: raa, vza (nobs,ngeom); stokes / nfourier / status / BR per pixel-view v = (j*nch + i)*ngeom + g
extern "C" int hm_vlidort(int ngeom, int S, int N, int km, int nch, int nobs, int nMom, int nPol, int planpar, int lambertian,
                          int scalar_glue, int nfine, int nadir_from_obs, double earth_radius, double accuracy,
                          const double *ROT, const double *depol, const double *alpha, const double *tau,
                          const double *ssa, const double *pmom, const double *ze, const double *albedo,
                          const double *sza, const double *raa, const double *vza, const double *flux_factor,
                          const double *brdf_f, const double *brdf_f_0, const double *user_brdf_f,
                          const double *user_brdf_f_0, const double *exactdb, int max_fourier_override,
                          int surf_kind, const double *u10m, const double *v10m, const double *mr, const double *kernel_wt,
                          const double *param, double *BR /*[npix]*/, int *wsa_fail /*[npix]*/,
                          double *stokes /*[npix][4]*/, int *nfourier /*[npix]*/, int *status /*[npix]*/) {
  const int G = ngeom > 1 ? ngeom : 1;
  const int npix = nobs * nch, nview = npix * G, L = km, n = S * N, NM = std::min(2 * N - 1, nMom);
  VLIn in; memset(&in, 0, sizeof in);
  in.ngeom = G;
  in.km = km; in.nch = nch; in.nobs = nobs; in.nMom = nMom; in.nPol = nPol; in.S = S; in.N = N; in.nfine = nfine;
  in.planpar = planpar; in.lambertian = lambertian; in.scalar_glue = scalar_glue; in.nadir_from_obs = nadir_from_obs;
  in.earth_radius = earth_radius; in.ROT = ROT; in.depol = depol; in.alpha = alpha; in.tau = tau; in.ssa = ssa; in.pmom = pmom;
  in.ze = ze; in.albedo = albedo; in.sza = sza; in.raa = raa; in.vza = vza; in.flux_factor = flux_factor; in.exactdb = exactdb;
  // BRDF wrappers: the same per-pixel bodies the CUDA kernels run, one pixel at a time
  VLBrdf B; std::vector<double> btab, bex, bf, bf0, buf, bscr_b;
  if (surf_kind != 0) {
    brdf_host_tables(B, surf_kind, S, scalar_glue, N, NM, nch, nobs, btab);
    B.G = G;
    B.sza = sza; B.vza = vza; B.raa = raa; B.u10m = u10m; B.v10m = v10m; B.mr = mr; B.kernel_wt = kernel_wt; B.param = param;
    bex.assign((size_t)nview * 16, 0.0); B.exactdb = bex.data(); B.wsa_fail = wsa_fail;
    bf.assign((size_t)npix * brdf_out_f(B), 0.0); bf0.assign((size_t)npix * brdf_out_f0(B), 0.0); buf.assign((size_t)nview * brdf_out_f0(B), 0.0);
    B.brdf_f = bf.data(); B.brdf_f_0 = bf0.data(); B.user_brdf_f = buf.data();
    const int nscr = 1 + 7 * ((B.nq == 16) ? 9 : 1) * B.NB;        // small batches: exercises the batching logic
    bscr_b.assign(nscr, 0.0);
    SerialTeam Tb;
    for (int v = 0; v < nview; v++) { brdf_exact_pixel(B, v); BR[v] = bex[(size_t)v * 16]; }
    for (int p = 0; p < npix; p++) brdf_pixel(Tb, B, p, p, bscr_b.data(), nscr);
    in.exactdb = bex.data();
  }
  VLPix X; X.L = L; X.NM = NM; X.S = S; X.npix = npix; X.G = G;
  std::vector<double> deltau((size_t)npix * L), og((size_t)npix * (NM + 1) * L * 4), a1((size_t)npix * L), a2(a1), a3(a1), a4((size_t)nview * L), a5(a4), a6(a4), a7(a4), scal((size_t)nview * 16, 0.0);
  std::vector<int> ibuf((size_t)nview * 4, 0);
  std::vector<unsigned char> layscat((size_t)npix * (NM + 1) * L, 1);
  X.deltau = deltau.data(); X.og = og.data(); X.avsec = a1.data(); X.itrans = a2.data(); X.tdmubar = a3.data(); X.tduserm = a4.data();
  X.emult_up = a5.data(); X.sigma_p = a6.data(); X.itrans_userm = a7.data(); X.scal = scal.data(); X.ibuf = ibuf.data(); X.layscat = layscat.data();
  SerialTeam T;
  std::vector<double> scr(12 * (L + 1) + 2 * (nMom + 1) + 2 * std::max(nfine, 1) + 16);
  for (int v = 0; v < nview; v++) setup_pixel(T, in, X, v, scr.data());
  VLQuad QD; QD.N = N;
  if (N == 1) { QD.x[0] = std::sqrt(1.0 / 3.0); QD.w[0] = 1.0; } else gauss01(N, QD.x, QD.w);
  // one pixel at a time to keep the workspace small
  VLSol Z; Z.n = n; Z.S = S; Z.N = N; Z.piu_g = nullptr; Z.pi0_g = nullptr; Z.zws = nullptr;
  std::vector<double> keig((size_t)L * n), tde(keig), xp((size_t)L * n * n), xn(xp), uhd((size_t)L * G * n * S), uhu(uhd), hm1((size_t)L * G * n), hm2(hm1), at(keig), bt(keig), gm(keig), gp(keig), wu((size_t)L * 2 * n), wl(wu);
  std::vector<int> fail(npix, 0), colact(npix, 0);
  std::vector<double> total((size_t)nview * 4, 0.0); std::vector<int> cstate((size_t)nview * 4, 0);
  VLConv CV; CV.total = total.data(); CV.state = cstate.data(); CV.col_active = colact.data();
  VLSurf SFc; SFc.brdf_f = brdf_f; SFc.brdf_f_0 = brdf_f_0; SFc.user_brdf_f = user_brdf_f; SFc.user_brdf_f_0 = user_brdf_f_0; SFc.nmfour = 2 * N; SFc.nq = 16;
  if (surf_kind != 0) { SFc.brdf_f = bf.data(); SFc.brdf_f_0 = bf0.data(); SFc.user_brdf_f = buf.data(); SFc.user_brdf_f_0 = nullptr; SFc.nq = B.nq; }
  std::vector<double> pq((size_t)(NM + 1) * N * 3), tmp(3 * (NM + 2));
  std::vector<double> escr(eigen_scratch_doubles(n, NM) + 64), bscr(bvp_smem_doubles(n) + 64), Ug(bvp_work_doubles(n, L));
  std::vector<int> iscr(13 * n + 32);
  // default K3: interaction-principle solver (vl_layerops.cuh + vl_add.cuh); AB_HM_LEGACY_BVP=1: the windowed band elimination
  const bool legacy = (getenv("AB_HM_LEGACY_BVP") != nullptr || n > 64) && G == 1;
  VLOps O;
  std::vector<double> ort((size_t)L * ops_block_doubles(n)), osp((size_t)L * n * n), osm(osp);
  std::vector<double> oscr(ops_smem_doubles(n, false) + 64), ascr(add_smem_doubles(n, false) + 64), awork(add_work_doubles(n, L));
  for (int p = 0; p < npix; p++) {
    int any = 0;
    for (int g = 0; g < G; g++) {
      const size_t v = (size_t)p * G + g;
      status[v] = ibuf[v * 4 + 1]; nfourier[v] = 0;
      for (int o = 0; o < 4; o++) stokes[v * 4 + o] = 0.0;
      cstate[v * 4] = status[v] == 0 ? 1 : 0;
      any |= status[v] == 0;
    }
    if (!any) continue;
    // solution arrays are indexed with q = p*L + k inside the bodies: shift the base pointers for this pixel
    const size_t q0 = (size_t)p * L;
    Z.keig = keig.data() - q0 * n; Z.tde = tde.data() - q0 * n; Z.xp = xp.data() - q0 * n * n; Z.xn = xn.data() - q0 * n * n;
    Z.uh_dn = uhd.data() - q0 * G * n * S; Z.uh_up = uhu.data() - q0 * G * n * S; Z.hm1 = hm1.data() - q0 * G * n; Z.hm2 = hm2.data() - q0 * G * n;
    Z.aterm = at.data() - q0 * n; Z.bterm = bt.data() - q0 * n; Z.gam_m = gm.data() - q0 * n; Z.gam_p = gp.data() - q0 * n;
    Z.wupper = wu.data() - q0 * 2 * n; Z.wlower = wl.data() - q0 * 2 * n; Z.fail = fail.data();
    colact[p] = 1;
    bool failed = false;
    for (int m = 0; m <= NM && colact[p] == 1; m++) {
      if (max_fourier_override >= 0 && m > max_fourier_override) break;
      for (int i = 0; i < N; i++) { pi_prt(m, NM, QD.x[i], tmp.data()); for (int l = m; l <= NM; l++) for (int c = 0; c < 3; c++) pq[((size_t)l * N + i) * 3 + c] = tmp[3 * l + c]; }
      for (int k = 0; k < L; k++) {
        if (S == 1) eigen_problem<1>(T, X, Z, QD, pq.data(), m, p, k, escr.data());
        else eigen_problem<3>(T, X, Z, QD, pq.data(), m, p, k, escr.data());
      }
      if (fail[p]) { failed = true; break; }
      const double acc = (max_fourier_override >= 0) ? -1.0 : accuracy;
      if (legacy) {
        if (S == 1) bvp_pixel<1>(T, X, Z, QD, SFc, CV, lambertian, acc, m, p, bscr.data(), iscr.data(), Ug.data());
        else bvp_pixel<3>(T, X, Z, QD, SFc, CV, lambertian, acc, m, p, bscr.data(), iscr.data(), Ug.data());
        colact[p] = cstate[(size_t)p * 4];
      } else {
        O.rt = ort.data() - q0 * ops_block_doubles(n); O.spT = osp.data() - q0 * n * n; O.smT = osm.data() - q0 * n * n;
        for (int k = 0; k < L; k++) {
          if (S == 1) layer_ops_problem<1>(T, X, Z, O, p, k, oscr.data(), iscr.data());
          else layer_ops_problem<3>(T, X, Z, O, p, k, oscr.data(), iscr.data());
        }
        if (fail[p]) { failed = true; break; }
        if (S == 1) add_pixel<1>(T, X, Z, O, QD, SFc, CV, lambertian, acc, m, p, ascr.data(), iscr.data(), awork.data());
        else add_pixel<3>(T, X, Z, O, QD, SFc, CV, lambertian, acc, m, p, ascr.data(), iscr.data(), awork.data());
      }
      for (int g = 0; g < G; g++) { const size_t v = (size_t)p * G + g; if (cstate[v * 4 + 3]) failed = true; else if (status[v] == 0 && (cstate[v * 4] == 1 || cstate[v * 4 + 2] == m || max_fourier_override >= 0)) nfourier[v] = m; }
      if (failed) break;
    }
    for (int g = 0; g < G; g++) {
      const size_t v = (size_t)p * G + g;
      if (failed && status[v] == 0) status[v] = 4;
      if (max_fourier_override < 0 && status[v] == 0) nfourier[v] = cstate[v * 4 + 2];
      for (int o = 0; o < S; o++) stokes[v * 4 + o] = total[v * 4 + o];
    }
  }
  return 0;
}

#ifdef AB_HM_SWEEP_HIST
// diagnostic build only (-DAB_HM_SWEEP_HIST): Jacobi sweeps per eigenproblem, by Fourier term
extern "C" void hm_sweep_histogram(long *out) { for (int m = 0; m < 16; m++) for (int s = 0; s < 32; s++) out[m * 32 + s] = hm_sweep_hist[m][s]; }
#endif
