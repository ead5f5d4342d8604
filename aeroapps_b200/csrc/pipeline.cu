// Chained column pipeline: the four calls leo_vlidort / polar_vlidort make per scene
//   MieObs_.getEdgeVars -> VLIDORT_POLAR_.ROT_CALC -> MieObs_.getAOPvector -> VLIDORT_POLAR_.VECTOR_CX
// (LV/polar_vlidort.py:600-950, LV/py_leo_vlidort/vlidort.py:400-580) behind ONE C-ABI entry.  Only the model state
// (qm, rh, airdens, delp: ~100 B per voxel) and the geometry cross PCIe; the edge variables, the Rayleigh optical thickness,
// tau / ssa and the 7.2 KB per voxel of phase-matrix moments stay in HBM between the stages.  Same kernels, same values
// as the four separate calls (the float32 -> float64 promotions and the P12 / P34 sign flip the Python driver does on the
// host, vlidort.py:438-440, are one elementwise kernel here), hence bit-identical results.
#include "../../include/aeroapps_b200.h"
#include "common.h"
#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace {
__global__ void k_f2d(const float *__restrict__ src, double *__restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (double)src[i];
}
// pmom (km,nch,nobs,nMom,nPol) float -> double, P12 (p = 1) and P34 (p = 3) negated
__global__ void k_pmom_f2d(const float *__restrict__ src, double *__restrict__ dst, size_t nvm, int nPol) {
  const size_t n = nvm * nPol;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / nvm);
    const double v = (double)__ldcs(src + i);
    dst[i] = (p == 1 || p == 3) ? -v : v;
  }
}
size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

extern "C" int ab200_pipeline_vector_cx(ab200_mie *h, int km, int nobs, int nch, int nq, const double *channels, const char *vname,
                                        const float *qm, const float *rh, const float *airdens, const float *delp, float ptop,
                                        int nstreams, int plane_parallel, int nMom, int nPol, const double *U10m,
                                        const double *V10m, const double *mr, const double *solar_zenith,
                                        const double *relat_azymuth, const double *sensor_zenith, double MISSING, int verbose,
                                        double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q, double *U,
                                        double *BR_Q, double *BR_U, int *rc) {
  (void)verbose;
  if (!h) { ab200_set_error("pipeline: null Mie handle"); return AB200_ERR_ARG; }
  if (rc) *rc = 0;
  if (km < 1 || nobs < 0 || nch < 1 || nq < 1) { ab200_set_error("pipeline: bad dimensions"); return AB200_ERR_ARG; }
  if (!channels || !vname || !qm || !rh || !airdens || !delp || !U10m || !V10m || !mr || !solar_zenith || !relat_azymuth || !sensor_zenith ||
      !radiance_VL_SURF || !reflectance_VL_SURF) { ab200_set_error("pipeline: NULL array argument"); return AB200_ERR_ARG; }
  const size_t npix = (size_t)nobs * nch;
  if (npix == 0) return 0;
  const size_t nvox = (size_t)km * nch * nobs, ncol = (size_t)km * nobs, ne = (size_t)(km + 1) * nobs, nrot = (size_t)km * nobs * nch;
  const size_t npm = nvox * (size_t)nMom * (size_t)nPol;
  // one grow-only device buffer: float section, then double section
  size_t off = 0;
  auto takef = [&](size_t n) { size_t o = off; off = al(off + 4 * n); return o; };
  auto taked = [&](size_t n) { size_t o = off; off = al(off + 8 * n); return o; };
  const size_t o_qm = takef(ncol * nq), o_rh = takef(ncol), o_ad = takef(ncol), o_dp = takef(ncol), o_pe = takef(ne), o_ze = takef(ne),
               o_te = takef(ne), o_tau = takef(nvox), o_ssa = takef(nvox), o_g = takef(nvox), o_pm = takef(npm);
  const size_t o_ped = taked(ne), o_zed = taked(ne), o_ted = taked(ne), o_ch = taked(nch), o_rot = taked(nrot), o_dep = taked(nch),
               o_taud = taked(nvox), o_ssad = taked(nvox), o_pmd = taked(npm), o_u = taked(nobs), o_v = taked(nobs), o_mr = taked(nch),
               o_sza = taked(nobs), o_raa = taked(nobs), o_vza = taked(nobs), o_out = taked(5 * npix);
  // one grow-only staging buffer per device, serialised per device (calls on different devices run concurrently)
  struct Ctx { std::mutex mu; DevBuf buf; };
  static std::mutex reg_mu;
  static std::map<int, std::unique_ptr<Ctx>> reg;
  Ctx *ctx;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(reg_mu);
    auto &u = reg[dev];
    if (!u) u.reset(new Ctx);
    ctx = u.get();
  }
  std::lock_guard<std::mutex> lock(ctx->mu);
  AB_CUDA_OK(ctx->buf.reserve(off + 256));
  char *b = ctx->buf.as<char>();
  auto F = [&](size_t o) { return reinterpret_cast<float *>(b + o); };
  auto D = [&](size_t o) { return reinterpret_cast<double *>(b + o); };
  cudaStream_t st = 0;
  AB_CUDA_OK(cudaMemcpyAsync(F(o_qm), qm, 4 * ncol * nq, cudaMemcpyHostToDevice, st));
  AB_CUDA_OK(cudaMemcpyAsync(F(o_rh), rh, 4 * ncol, cudaMemcpyHostToDevice, st));
  AB_CUDA_OK(cudaMemcpyAsync(F(o_ad), airdens, 4 * ncol, cudaMemcpyHostToDevice, st));
  AB_CUDA_OK(cudaMemcpyAsync(F(o_dp), delp, 4 * ncol, cudaMemcpyHostToDevice, st));
  // the driver's one wavelength list [nm], real*8: ROT_CALC and VLIDORT see it as it is, the Mie channel match gets the real*4
  // copy f2py makes for MieObs_ (MieObs_py.F90:232)
  std::vector<float> chf(nch);
  for (int i = 0; i < nch; i++) chf[i] = (float)channels[i];
  AB_CUDA_OK(cudaMemcpyAsync(D(o_ch), channels, 8 * (size_t)nch, cudaMemcpyHostToDevice, st));
  AB_CUDA_OK(cudaMemcpyAsync(D(o_u), U10m, 8 * (size_t)nobs, cudaMemcpyHostToDevice, st));
  AB_CUDA_OK(cudaMemcpyAsync(D(o_v), V10m, 8 * (size_t)nobs, cudaMemcpyHostToDevice, st));
  AB_CUDA_OK(cudaMemcpyAsync(D(o_mr), mr, 8 * (size_t)nch, cudaMemcpyHostToDevice, st));
  AB_CUDA_OK(cudaMemcpyAsync(D(o_sza), solar_zenith, 8 * (size_t)nobs, cudaMemcpyHostToDevice, st));
  AB_CUDA_OK(cudaMemcpyAsync(D(o_raa), relat_azymuth, 8 * (size_t)nobs, cudaMemcpyHostToDevice, st));
  AB_CUDA_OK(cudaMemcpyAsync(D(o_vza), sensor_zenith, 8 * (size_t)nobs, cudaMemcpyHostToDevice, st));
  // stage 1: hydrostatic edge variables (float32, MieObs_py.F90:27-83), promoted to float64 for the RT side
  if (int e = ab200_get_edge_vars_dev(km, nobs, F(o_ad), F(o_dp), ptop, F(o_pe), F(o_ze), F(o_te), st)) return e;
  k_f2d<<<296, 256, 0, st>>>(F(o_pe), D(o_ped), ne);
  k_f2d<<<296, 256, 0, st>>>(F(o_ze), D(o_zed), ne);
  k_f2d<<<296, 256, 0, st>>>(F(o_te), D(o_ted), ne);
  // stage 2: Rayleigh optical thickness and depolarisation (GV/VLIDORT_ROT.F90)
  if (int e = ab200_rot_calc_dev(km, nch, nobs, D(o_ch), D(o_ped), D(o_zed), D(o_ted), D(o_rot), D(o_dep), st)) return e;
  // stage 3: Mie AOPs with the phase-matrix moments (MieObs_py.F90:218-456)
  int rc_ = 0;
  if (int e = ab200_get_aop_vector_dev(h, km, nobs, nch, nq, chf.data(), vname, F(o_qm), F(o_rh), nMom, nPol, F(o_tau), F(o_ssa),
                                       F(o_g), F(o_pm), st, &rc_)) { if (rc) *rc = rc_; return e; }
  if (rc_) { if (rc) *rc = rc_; return rc_; }
  k_f2d<<<296, 256, 0, st>>>(F(o_tau), D(o_taud), nvox);
  k_f2d<<<296, 256, 0, st>>>(F(o_ssa), D(o_ssad), nvox);
  k_pmom_f2d<<<148 * 8, 256, 0, st>>>(F(o_pm), D(o_pmd), nvox * (size_t)nMom, nPol);
  AB_CUDA_OK(cudaGetLastError());
  // stage 4: vector VLIDORT over the Cox-Munk surface (GV/VLIDORT_BRDF_CX.F90:587-762)
  double *o = D(o_out);
  if (int e = ab200_vlidort_vector_cx_dev(km, nch, nobs, nstreams, plane_parallel, nMom, nPol, D(o_rot), D(o_dep), D(o_taud), D(o_ssad),
                                          D(o_pmd), D(o_zed), D(o_u), D(o_v), D(o_mr), D(o_sza), D(o_raa), D(o_vza), MISSING, o,
                                          o + npix, o + 2 * npix, o + 3 * npix, o + 4 * npix, st, &rc_)) { if (rc) *rc = rc_; return e; }
  if (rc) *rc = rc_;
  if (rc_) return rc_;
  AB_CUDA_OK(cudaMemcpyAsync(radiance_VL_SURF, o, 8 * npix, cudaMemcpyDeviceToHost, st));
  AB_CUDA_OK(cudaMemcpyAsync(reflectance_VL_SURF, o + npix, 8 * npix, cudaMemcpyDeviceToHost, st));
  if (BR) AB_CUDA_OK(cudaMemcpyAsync(BR, o + 2 * npix, 8 * npix, cudaMemcpyDeviceToHost, st));
  if (Q) AB_CUDA_OK(cudaMemcpyAsync(Q, o + 3 * npix, 8 * npix, cudaMemcpyDeviceToHost, st));
  if (U) AB_CUDA_OK(cudaMemcpyAsync(U, o + 4 * npix, 8 * npix, cudaMemcpyDeviceToHost, st));
  AB_CUDA_OK(cudaStreamSynchronize(st));
  if (BR_Q) for (size_t i = 0; i < npix; i++) BR_Q[i] = 0.0;       // VLIDORT_Vector_CX: BRDF_Q = BRDF_U = 0 (:733-734)
  if (BR_U) for (size_t i = 0; i < npix; i++) BR_U[i] = 0.0;
  return 0;
}
