// Stage B -- VLIDORT drop-in entry points on B200: kernel wrappers, chunked host driver, C ABI.
// Kernels: k_vl_setup (K1, CTA per pixel), k_vl_eigen (K2, warp per pixel-layer), k_vl_bvp (K3, persistent CTAs,
// one pixel at a time), k_vl_finish.  The Fourier loop runs on the host; pixels that have converged drop out
// of the device-side active list, so later components only launch work for the pixels that still need it.
#include "../../include/aeroapps_b200.h"
#include "common.h"
#include "vl_add.cuh"
#ifdef AB200_BVP_VARIANTS            // opt-in build (make EXTRA=-DAB200_BVP_VARIANTS): the two experimental forward eliminations of round 1
#include "vl_bvp3.cuh"
#include "vl_bvp4.cuh"
#endif
#include "vl_brdf.cuh"
#include <algorithm>
#include <vector>
#include <cmath>

using namespace vlk;

struct VLStats { double ms_setup, ms_eigen, ms_bvp, flops_eigen, flops_bvp; long n_eigen_problems, n_bvp_problems, n_launches; long npix; double ms_ops, flops_exec; };
static VLStats g_stats;
static std::vector<int> g_last_nfourier, g_last_status;

// ------------------------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(128) k_vl_setup(VLIn in, VLPix X, int p0) {
  extern __shared__ double smem_d[];
  __shared__ double sd[64];
  __shared__ int si[64];
  BlockTeam T(sd, si);
  const int pl = blockIdx.x;                 // pixel index inside the chunk
  // the inputs are indexed by the global pixel, the per-pixel state by the chunk-local pixel
  VLPix Xl = X;
  VLIn inl = in;
  // shift input pixel numbering: setup_pixel derives (j,i) from p, outputs are addressed with the same p
  // -> give it a state view whose arrays start p0 pixels "before" the chunk
  const int L = X.L, NM = X.NM;
  Xl.deltau -= (size_t)p0 * L; Xl.og -= (size_t)p0 * (NM + 1) * L * 4; Xl.avsec -= (size_t)p0 * L; Xl.itrans -= (size_t)p0 * L;
  Xl.tdmubar -= (size_t)p0 * L; Xl.tduserm -= (size_t)p0 * L; Xl.emult_up -= (size_t)p0 * L; Xl.sigma_p -= (size_t)p0 * L;
  Xl.itrans_userm -= (size_t)p0 * L; Xl.scal -= (size_t)p0 * 16; Xl.ibuf -= (size_t)p0 * 4; Xl.layscat -= (size_t)p0 * (NM + 1) * L;
  setup_pixel(T, inl, Xl, p0 + pl, smem_d);
}

template <int S>
__global__ void __launch_bounds__(256) k_vl_eigen(VLPix X, VLSol Z, VLQuad QD, const double *__restrict__ pq, int m, const int *__restrict__ active,
                           int nactive, int scr_doubles) {
  extern __shared__ double smem_d[];
  const int warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const long prob = (long)blockIdx.x * wpc + warp;
  if (prob >= (long)nactive * X.L) return;
  const int p = active[prob / X.L], k = (int)(prob % X.L);
  WarpTeam T;
  eigen_problem<S>(T, X, Z, QD, pq, m, p, k, smem_d + (size_t)warp * scr_doubles);
}

// W lanes per (pixel, layer) problem (SubWarpTeam): 32 / W problems per warp for small nstokes*nstreams (the scalar 8-stream
// configuration keeps 4 of 32 lanes busy with one problem per warp)
template <int S, int W>
__global__ void __launch_bounds__(256) k_vl_eigen_sub(VLPix X, VLSol Z, VLQuad QD, const double *__restrict__ pq, int m,
                                                      const int *__restrict__ active, int nactive, int scr_doubles) {
  extern __shared__ double smem_d[];
  constexpr int SUBS = 32 / W;
  const int slot = (threadIdx.x >> 5) * SUBS + ((threadIdx.x & 31) / W), spc = (blockDim.x >> 5) * SUBS;
  const long prob = (long)blockIdx.x * spc + slot;
  if (prob >= (long)nactive * X.L) return;
  const int p = active[prob / X.L], k = (int)(prob % X.L);
  SubWarpTeam<W> T;
  eigen_problem<S>(T, X, Z, QD, pq, m, p, k, smem_d + (size_t)slot * scr_doubles);
}

// two warps per (pixel, layer) problem (DuoTeam): same shared memory per problem, twice the warps per SM
template <int S, int MAXHALF = 18>
__global__ void __launch_bounds__(256, 2) k_vl_eigen_duo(VLPix X, VLSol Z, VLQuad QD, const double *__restrict__ pq, int m,
                                                        const int *__restrict__ active, int nactive, int scr_doubles) {
  extern __shared__ double smem_d[];
  const int duo = threadIdx.x >> 6, ppc = blockDim.x >> 6;
  const long prob = (long)blockIdx.x * ppc + duo;
  if (prob >= (long)nactive * X.L) return;                  // both warps of a duo leave together: no barrier is left short
  const int p = active[prob / X.L], k = (int)(prob % X.L);
  double *scr = smem_d + (size_t)duo * (scr_doubles + 4);
  DuoTeamT<MAXHALF> T(scr + scr_doubles);
  eigen_problem<S>(T, X, Z, QD, pq, m, p, k, scr);
}

// VAR: forward-elimination variant, one kernel each (own register allocation): 0 = blocked look-ahead (vl_bvp.cuh, the default),
// 1 = the same with the window in a per-CTA HBM scratch ("large n"), 3 = warp-specialised (vl_bvp3.cuh), 4 = register-resident
// window (vl_bvp4.cuh)
template <int S, int VAR>
__global__ void __launch_bounds__(384) k_vl_bvp(VLPix X, VLSol Z, VLQuad QD, VLSurf SF, VLConv CV, int lambertian, double accuracy,
                                                int m, const int *__restrict__ active, int nactive, int *__restrict__ next_list,
                                                int *__restrict__ next_count, double *__restrict__ work, size_t work_per_cta, int v3,
                                                int smem_doubles, double *__restrict__ wglob) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double sd[64];
  __shared__ int si[64];
  const int n = S * QD.N;
  (void)v3;
#ifdef AB200_BVP_VARIANTS
  int *iscr = reinterpret_cast<int *>(smem_d + (VAR == 4 ? b4_smem_doubles(n) : (VAR == 3 ? b3_smem_doubles(n) : (VAR == 1 ? bvp_smem_doubles_g(n) : bvp_smem_doubles(n)))));
#else
  int *iscr = reinterpret_cast<int *>(smem_d + (VAR == 1 ? bvp_smem_doubles_g(n) : bvp_smem_doubles(n)));
#endif
  double *wg = (VAR == 1) ? wglob + (size_t)blockIdx.x * bvp_window_doubles(n) : nullptr;
  __shared__ unsigned long long mbar;
  for (int i = threadIdx.x; i < smem_doubles; i += blockDim.x) smem_d[i] = 0.0;      // no stale NaN patterns in never-written padding
  BlockTeam T(sd, si, &mbar);
  double *Ug = work + (size_t)blockIdx.x * work_per_cta;
  for (int a = blockIdx.x; a < nactive; a += gridDim.x) {
    const int p = active[a];
    if (Z.fail[p]) {
      if (threadIdx.x == 0) { CV.state[(size_t)p * 4] = 0; CV.state[(size_t)p * 4 + 3] = 4; }
    } else {
#ifdef AB200_BVP_VARIANTS
      if constexpr (VAR == 4) bvp_pixel_v4<S>(T, X, Z, QD, SF, CV, lambertian, accuracy, m, p, smem_d, iscr, Ug);
      else if constexpr (VAR == 3) bvp_pixel_v3<S>(T, X, Z, QD, SF, CV, lambertian, accuracy, m, p, smem_d, iscr, Ug);
      else
#endif
      if constexpr (VAR == 1) bvp_pixel<S, BlockTeam, true>(T, X, Z, QD, SF, CV, lambertian, accuracy, m, p, smem_d, iscr, Ug, wg);
      else bvp_pixel<S>(T, X, Z, QD, SF, CV, lambertian, accuracy, m, p, smem_d, iscr, Ug);
    }
    __syncthreads();
    if (threadIdx.x == 0 && CV.state[(size_t)p * 4] == 1) { const int idx = atomicAdd(next_count, 1); next_list[idx] = p; }
    __syncthreads();
  }
}

// K3 for small problems (3 nstokes nstreams <= 64 window rows, e.g. the scalar 8-stream configuration): one WARP per pixel with its
// own window in shared memory, several pixels per CTA -- the 384-thread kernel above leaves the SM almost idle on a 24 x 33 window
template <int S>
__global__ void __launch_bounds__(256) k_vl_bvp_warp(VLPix X, VLSol Z, VLQuad QD, VLSurf SF, VLConv CV, int lambertian, double accuracy,
                                                     int m, const int *__restrict__ active, int nactive, int *__restrict__ next_list,
                                                     int *__restrict__ next_count, double *__restrict__ work, size_t work_per_warp,
                                                     int smem_doubles_per_warp) {
  extern __shared__ __align__(16) double smem_d[];
  const int n = S * QD.N, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  double *scr = smem_d + (size_t)warp * smem_doubles_per_warp;
  for (int i = lane; i < smem_doubles_per_warp; i += 32) scr[i] = 0.0;
  __syncwarp();
  int *iscr = reinterpret_cast<int *>(scr + bvp_smem_doubles(n));
  WarpTeam T;
  const int gw = blockIdx.x * wpc + warp;
  double *Ug = work + (size_t)gw * work_per_warp;
  for (int a = gw; a < nactive; a += gridDim.x * wpc) {
    const int p = active[a];
    if (Z.fail[p]) {
      if (lane == 0) { CV.state[(size_t)p * 4] = 0; CV.state[(size_t)p * 4 + 3] = 4; }
    } else {
      bvp_pixel<S>(T, X, Z, QD, SF, CV, lambertian, accuracy, m, p, scr, iscr, Ug);
    }
    __syncwarp();
    if (lane == 0 && CV.state[(size_t)p * 4] == 1) { const int idx = atomicAdd(next_count, 1); next_list[idx] = p; }
    __syncwarp();
  }
}

// K2b: layer interaction operators (vl_layerops.cuh), one CTA per (pixel, layer) problem
template <int S>
__global__ void __launch_bounds__(128) k_vl_layer_ops(VLPix X, VLSol Z, VLOps O, const int *__restrict__ active, int nactive) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double sd[64];
  __shared__ int si[64];
  const long prob = blockIdx.x;
  const int p = active[prob / X.L], k = (int)(prob % X.L);
  BlockTeam T(sd, si);
  layer_ops_problem<S>(T, X, Z, O, p, k, smem_d, reinterpret_cast<int *>(smem_d + ops_smem_doubles(Z.n)));
}
// small problems: one warp per (pixel, layer) problem, WPC problems per CTA
template <int S>
__global__ void __launch_bounds__(256) k_vl_layer_ops_warp(VLPix X, VLSol Z, VLOps O, const int *__restrict__ active, int nactive, int per_warp) {
  extern __shared__ __align__(16) double smem_d[];
  const int warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const long prob = (long)blockIdx.x * wpc + warp;
  if (prob >= (long)nactive * X.L) return;
  const int p = active[prob / X.L], k = (int)(prob % X.L);
  double *scr = smem_d + (size_t)warp * per_warp;
  WarpTeam T;
  layer_ops_problem<S>(T, X, Z, O, p, k, scr, reinterpret_cast<int *>(scr + ops_smem_doubles(Z.n)));
}

// K3 (default): interaction-principle solver (vl_add.cuh), persistent CTAs of 128 threads, three per SM at nstokes*nstreams = 36
template <int S>
__global__ void __launch_bounds__(128, 3) k_vl_add(VLPix X, VLSol Z, VLOps O, VLQuad QD, VLSurf SF, VLConv CV, int lambertian, double accuracy,
                                                   int m, const int *__restrict__ active, int nactive, int *__restrict__ next_list,
                                                   int *__restrict__ next_count, double *__restrict__ work, size_t work_per_cta) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double sd[64];
  __shared__ int si[64];
  BlockTeam T(sd, si);
  int *piv = reinterpret_cast<int *>(smem_d + add_smem_doubles(Z.n));
  double *Wg = work + (size_t)blockIdx.x * work_per_cta;
  for (int a = blockIdx.x; a < nactive; a += gridDim.x) {
    const int p = active[a];
    if (Z.fail[p]) {
      if (threadIdx.x == 0) { CV.state[(size_t)p * 4] = 0; CV.state[(size_t)p * 4 + 3] = 4; }
    } else {
      add_pixel<S>(T, X, Z, O, QD, SF, CV, lambertian, accuracy, m, p, smem_d, piv, Wg);
    }
    __syncthreads();
    if (threadIdx.x == 0 && CV.state[(size_t)p * 4] == 1) { const int idx = atomicAdd(next_count, 1); next_list[idx] = p; }
    __syncthreads();
  }
}
// small problems: one warp per pixel
template <int S>
__global__ void __launch_bounds__(256) k_vl_add_warp(VLPix X, VLSol Z, VLOps O, VLQuad QD, VLSurf SF, VLConv CV, int lambertian, double accuracy,
                                                     int m, const int *__restrict__ active, int nactive, int *__restrict__ next_list,
                                                     int *__restrict__ next_count, double *__restrict__ work, size_t work_per_warp, int per_warp) {
  extern __shared__ __align__(16) double smem_d[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  double *scr = smem_d + (size_t)warp * per_warp;
  int *piv = reinterpret_cast<int *>(scr + add_smem_doubles(Z.n));
  WarpTeam T;
  const int gw = blockIdx.x * wpc + warp;
  double *Wg = work + (size_t)gw * work_per_warp;
  for (int a = gw; a < nactive; a += gridDim.x * wpc) {
    const int p = active[a];
    if (Z.fail[p]) {
      if (lane == 0) { CV.state[(size_t)p * 4] = 0; CV.state[(size_t)p * 4 + 3] = 4; }
    } else {
      add_pixel<S>(T, X, Z, O, QD, SF, CV, lambertian, accuracy, m, p, scr, piv, Wg);
    }
    __syncwarp();
    if (lane == 0 && CV.state[(size_t)p * 4] == 1) { const int idx = atomicAdd(next_count, 1); next_list[idx] = p; }
    __syncwarp();
  }
}

// K0a: exact direct-bounce BRDF of every pixel + the wrappers' BR outputs (nobs,nch)
__global__ void k_vl_brdf_exact(VLBrdf B, int npix, const unsigned char *__restrict__ skip, double *BR, double *BR_Q, double *BR_U) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const int j = p / B.nch, i = p % B.nch;
  const size_t o = j + (size_t)B.nobs * i;
  if (skip[p]) { for (int q = 0; q < 16; q++) B.exactdb[(size_t)p * 16 + q] = 0.0; BR[o] = 0.0; if (BR_Q) { BR_Q[o] = 0.0; BR_U[o] = 0.0; } return; }
  brdf_exact_pixel(B, p);
  double br = B.exactdb[(size_t)p * 16];
  const bool neg = (B.kind == SURF_MODIS && br < 0);
  if (neg) br = -500.0;
  BR[o] = br;
  // BRDF_Q / BRDF_U: elements (1,2) and (1,3) of the exact direct-bounce BRDF for the GISS and the OCI wrappers
  // (BS_DBOUNCE_BRDFUNC(2,1,1,1), (3,1,1,1): VLIDORT_BRDF_CX.F90:199-201,388-389,1865-1866), 0 for Vector_CX (:733-734) and
  // LandMODIS (VLIDORT_BRDF_MODIS.F90:365-366)
  if (BR_Q) {
    const bool pol = surf_is_giss(B.kind) || B.kind == SURF_OCICX;
    BR_Q[o] = neg ? -500.0 : (pol ? B.exactdb[(size_t)p * 16 + 1] : 0.0);
    BR_U[o] = neg ? -500.0 : (pol ? B.exactdb[(size_t)p * 16 + 2] : 0.0);
  }
}

// K0b: BRDF Fourier components of the pixels of one chunk, one CTA per pixel
__global__ void __launch_bounds__(256) k_vl_brdf_fourier(VLBrdf B, int p0, const unsigned char *__restrict__ skip, int nscr) {
  extern __shared__ double smem_d[];
  __shared__ double sd[64];
  __shared__ int si[64];
  const int pc = blockIdx.x, p = p0 + pc;
  if (skip[p]) return;
  BlockTeam T(sd, si);
  brdf_pixel(T, B, p, pc, smem_d, nscr);
}

// outputs in the drop-in layout (nobs,nch); reflectance uses the single-precision pi of the glue
// (VLIDORT_ScatMod.F90:344,596)
__global__ void k_vl_finish(VLPix X, VLConv CV, int p0, int npix_chunk, int nch, int nobs, const double *__restrict__ sza,
                            double MISSING, int S, double *radiance, double *reflectance, double *Q, double *U,
                            int *nfourier, int *status) {
  const int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= npix_chunk) return;
  const int p = p0 + pl, j = p / nch, i = p % nch;
  const size_t o = j + (size_t)nobs * i;
  const int st0 = X.ibuf[(size_t)pl * 4 + 1];
  const int fail = CV.state[(size_t)pl * 4 + 3];
  const int st = st0 != 0 ? st0 : fail;
  if (status) status[p] = st;
  if (nfourier) nfourier[p] = CV.state[(size_t)pl * 4 + 2];
  if (st != 0) {
    const double fill = (st == 98) ? -500.0 : MISSING;     // 98: RTLS bidirectional reflectance < 0 (GV/VLIDORT_BRDF_MODIS.F90:368-377)
    radiance[o] = fill; reflectance[o] = fill;
    if (Q) Q[o] = fill;
    if (U) U[o] = fill;
    return;
  }
  const double pi_f = (double)3.14159274101257324f;      // 4.*atan(1.0) in real*4
  const double I = CV.total[(size_t)pl * 4];
  radiance[o] = I;
  reflectance[o] = (pi_f * I) / (cos(sza[j] * pi_f / 180.0) * X.scal[(size_t)pl * 16 + 5]);
  if (Q) Q[o] = S > 1 ? CV.total[(size_t)pl * 4 + 1] : 0.0;
  if (U) U[o] = S > 1 ? CV.total[(size_t)pl * 4 + 2] : 0.0;
}

__global__ void k_vl_init_active(VLPix X, VLConv CV, VLSol Z, int npix_chunk, int *list, int *count) {
  const int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= npix_chunk) return;
  const bool ok = X.ibuf[(size_t)pl * 4 + 1] == 0;
  CV.state[(size_t)pl * 4] = ok ? 1 : 0; CV.state[(size_t)pl * 4 + 1] = 0; CV.state[(size_t)pl * 4 + 2] = 0; CV.state[(size_t)pl * 4 + 3] = 0;
  for (int o = 0; o < 4; o++) CV.total[(size_t)pl * 4 + o] = 0.0;
  Z.fail[pl] = 0;
  if (ok) { const int idx = atomicAdd(count, 1); list[idx] = pl; }
}

// Rayleigh optical thickness (VLIDORT_ROT_CALC): one thread per (layer, obs, channel)
__device__ double f_air_dev(double XX2, double C) {
  const double depol1 = (double)1.034f + (double)3.17e-4f * XX2;
  const double depol2 = (double)1.096f + (double)1.385e-3f * XX2 + (double)1.448e-4f * XX2 * XX2;
  const float s3 = (78.084f + 20.946f) + 0.934f;
  return ((double)78.084f * depol1 + (double)20.946f * depol2 + (double)0.934f * 1.0 + C * (double)100.0f * (double)1.15f) /
         ((double)s3 + C * (double)100.0f);
}
__global__ void k_rot_calc(int km, int nch, int nobs, const double *__restrict__ channels, const double *__restrict__ pe,
                           const double *__restrict__ he, const double *__restrict__ te, double *__restrict__ ROT,
                           double *__restrict__ depol) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= (long)km * nobs * nch) return;
  const int k = (int)(t % km); const long r = t / km; const int j = (int)(r % nobs), i = (int)(r / nobs);
  const double Av = (double)6.0221367e23f, Ts = (double)288.15f, Ps = (double)1013.25f;
  const double pi = (double)3.14159274101257324f;
  const double Ns = (Av / (double)22.4141f) * ((double)273.15f / Ts) * (double)(1.0f / 1000.f);
  const double C = (double)(360.0f * 1e-6f);
  const double wmicron = channels[i] * (double)1.e-3f;
  const double XX = 1.0 / wmicron, XX2 = XX * XX, CM = wmicron / (double)10000.0f, pi3 = pi * pi * pi;
  double RI = (double)8060.51f + (double)2480990.0f / ((double)132.274f - XX2) + (double)17455.7f / ((double)39.32957f - XX2);
  RI = RI * (double)1e-8f + 1.0;
  RI = (RI - 1) * (1.0 + (double)0.54f * (C - (double)0.0003f)) + 1;
  const double RI2 = RI * RI, fa = f_air_dev(XX2, C);
  const double A = 24.0 * pi3 * fa * ((RI2 - 1) * (RI2 - 1)) / (Ns * Ns * CM * CM * CM * CM * (RI2 + 2) * (RI2 + 2));
  if (k == 0 && j == 0) depol[i] = 6.0 * (fa - 1.0) / (7.0 * fa + 3.0);
  const size_t e0 = k + (size_t)(km + 1) * j;
  double v0 = A * Ns * ((pe[e0] * (double)1.e-2f) / Ps) * (Ts / te[e0]); v0 = v0 * (double)1.e5f;
  double v1 = A * Ns * ((pe[e0 + 1] * (double)1.e-2f) / Ps) * (Ts / te[e0 + 1]); v1 = v1 * (double)1.e5f;
  const double difz = he[e0] * (double)1.e-3f - he[e0 + 1] * (double)1.e-3f;
  ROT[k + (size_t)km * (j + (size_t)nobs * i)] = difz * (v0 + v1) / 2.0;
}

// ------------------------------------------------------------------------------------------------ host driver
namespace {

struct DevArena {                      // one allocation carved into aligned pieces
  char *base = nullptr; size_t cap = 0, off = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) { off = 0; return cudaSuccess; }
    if (base) cudaFree(base);
    base = nullptr; cap = 0; off = 0;
    cudaError_t e = cudaMalloc((void **)&base, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  template <class T> T *take(size_t n) { off = (off + 255) / 256 * 256; T *p = reinterpret_cast<T *>(base + off); off += n * sizeof(T); return p; }
};
DevArena g_arena;          // grow-only; released by ab200_vlidort_release()

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct RunDesc {
  int S, N, km, nch, nobs, nMom, nPol, planpar, lambertian, scalar_glue;
  double MISSING;
  // device inputs (Fortran layouts)
  const double *ROT, *depol, *alpha, *tau, *ssa, *pmom, *he, *albedo, *sza, *raa, *vza, *flux_factor;
  // BRDF (device, per pixel) when !lambertian
  VLSurf surf; const double *exactdb;
  const VLBrdf *brdf;              // non-null: the Fourier components are computed per chunk from the wrappers' surface inputs
  // host copies of the per-obs angles and the per-pixel skip mask (host decides MISSING / state quirk)
  const std::vector<double> *h_sza; const std::vector<unsigned char> *h_skip;
  // device outputs (nobs,nch)
  double *radiance, *reflectance, *Q, *U; int *nfourier, *status;
  cudaStream_t stream;
};

int run_vlidort(const RunDesc &R) {
  const int S = R.S, N = R.N, L = R.km, n = S * N, NM = std::min(2 * N - 1, R.nMom), npix = R.nobs * R.nch;
  if (npix == 0) return 0;
  if (N < 1 || N > 32 || L < 1 || L > MAXLAY || (S != 1 && S != 3) || 2 * N > MAXMOM) { ab200_set_error("vlidort: unsupported nstreams/nlayers/nstokes"); return AB200_ERR_ARG; }
  cudaStream_t st = R.stream;
  // ---- sizes
  const size_t per_pix_state = (size_t)L * 8 * 8 + (size_t)(NM + 1) * L * 4 * 8 + 16 * 8 + 4 * 4 + (size_t)(NM + 1) * L;
  // K3 variant: interaction-principle solver (vl_add.cuh) for nstokes*nstreams <= 64; AB200_VL_BVP_LEGACY=1 (or larger problems)
  // selects the windowed band elimination of vl_bvp.cuh
  bool use_add = n <= 64;
  if (const char *e = getenv("AB200_VL_BVP_LEGACY")) use_add = use_add && atoi(e) == 0;
  const size_t per_pix_sol = (size_t)L * ((size_t)2 * n * n + (6 + 2 * S) * n + 2 * n + 2 * 2 * n + (use_add ? 4 * (size_t)n * n + 2 * n : 0)) * 8;
  size_t per_pix_brdf = 0; int brdf_nscr = 0; size_t brdf_smem = 0;
  if (R.brdf) {
    const VLBrdf &B0 = *R.brdf;
    per_pix_brdf = (brdf_out_f(B0) + 2 * brdf_out_f0(B0)) * 8;
    const int nqs = (B0.nq == 16) ? 9 : 1, NE = N * N + 2 * N;
    const size_t want = (size_t)NE * nqs * B0.NB, cap = 20000;
    brdf_nscr = 256 + (int)std::min(want, std::max<size_t>(cap, (size_t)nqs * B0.NB));
    brdf_smem = (size_t)brdf_nscr * 8;
    AB_CUDA_OK(cudaFuncSetAttribute((const void *)k_vl_brdf_fourier, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)brdf_smem));
  }
  const size_t per_pix = per_pix_state + per_pix_sol + 4 * 8 + 4 * 4 + 4 * 3 + per_pix_brdf;
  double ws_gb = 24.0;
  if (const char *e = getenv("AB200_VL_WORKSPACE_GB")) ws_gb = atof(e);
  size_t freeb = 0, totalb = 0;
  cudaMemGetInfo(&freeb, &totalb);
  size_t budget = (size_t)(ws_gb * 1e9);
  budget = std::min(budget, (size_t)(0.6 * (double)(freeb + g_arena.cap)));
  // K3 geometry
  const int bvp_threads = 384;
  // warp-specialised forward elimination (vl_bvp3.cuh): opt-in with AB200_VL_BVP_V3=1 while it is slower than the blocked
  // look-ahead one (measured 2426 vs 2893 pixels/s on configs[2]); needs its window (3n+16 rows) to fit 128 row slots
  int use_v3 = 0;
#ifdef AB200_BVP_VARIANTS
  if (const char *e = getenv("AB200_VL_BVP_V3"))
    use_v3 = (atoi(e) != 0 && b3_supported(n) && (b3_smem_doubles(n) * 8 + b3_smem_ints(n) * 4 <= 227 * 1024)) ? 3 : 0;
  // register-resident forward elimination (vl_bvp4.cuh): opt-in with AB200_VL_BVP_V4=1 (n > 24) or =2 (every supported n).
  // Parity green; measured 1316 ms vs 1189 ms of K3 time per 3 bench steps, so the blocked look-ahead one stays the default
  {
    int v4 = 0;
    if (const char *e = getenv("AB200_VL_BVP_V4")) v4 = atoi(e);
    if (!use_v3 && v4 && b4_supported(n) && (v4 >= 2 || n > 24)) use_v3 = 4;
  }
#endif
  // default variant: when the 3n x (4n+1) window does not fit the 227 KB of shared memory (e.g. nstokes 3 x 16 streams) it moves
  // to a per-CTA HBM scratch ("large n" layout, vl_bvp.cuh)
  const bool big_n = !use_v3 && (bvp_smem_doubles(n) * 8 + (size_t)(13 * n + 32) * 4 > 227 * 1024);
#ifdef AB200_BVP_VARIANTS
  const size_t bvp_smem_d = use_v3 == 4 ? b4_smem_doubles(n) : (use_v3 ? b3_smem_doubles(n) : (big_n ? bvp_smem_doubles_g(n) : bvp_smem_doubles(n)));
  const size_t bvp_smem = bvp_smem_d * 8 + (use_v3 == 3 ? b3_smem_ints(n) : (size_t)(13 * n + 32)) * 4;
#else
  const size_t bvp_smem_d = big_n ? bvp_smem_doubles_g(n) : bvp_smem_doubles(n);
  const size_t bvp_smem = bvp_smem_d * 8 + (size_t)(13 * n + 32) * 4;
#endif
  int dev = 0, nsm = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  int bvp_per_sm = 1;
  const int var = use_v3 ? use_v3 : (big_n ? 1 : 0);
  using bvp_fn = void (*)(VLPix, VLSol, VLQuad, VLSurf, VLConv, int, double, int, const int *, int, int *, int *, double *, size_t, int, int, double *);
#ifdef AB200_BVP_VARIANTS
  const bvp_fn kbf = (S == 1) ? (var == 4 ? k_vl_bvp<1, 4> : var == 3 ? k_vl_bvp<1, 3> : var == 1 ? k_vl_bvp<1, 1> : k_vl_bvp<1, 0>)
                              : (var == 4 ? k_vl_bvp<3, 4> : var == 3 ? k_vl_bvp<3, 3> : var == 1 ? k_vl_bvp<3, 1> : k_vl_bvp<3, 0>);
#else
  const bvp_fn kbf = (S == 1) ? (var == 1 ? k_vl_bvp<1, 1> : k_vl_bvp<1, 0>) : (var == 1 ? k_vl_bvp<3, 1> : k_vl_bvp<3, 0>);
#endif
  auto kb = (const void *)kbf;
  auto ke = (S == 1) ? (const void *)k_vl_eigen<1> : (const void *)k_vl_eigen<3>;
  if (!use_add) {
    if (bvp_smem > 227 * 1024) { ab200_set_error("vlidort: nstokes*nstreams too large for the shared-memory BVP window"); return AB200_ERR_ARG; }
    AB_CUDA_OK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bvp_smem));
    AB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bvp_per_sm, kb, bvp_threads, bvp_smem));
    bvp_per_sm = std::max(1, std::min(bvp_per_sm, 8));
  }
  // small problems (3n <= 64 window rows): one warp per pixel, 4 pixels per CTA (k_vl_bvp_warp); AB200_VL_BVP_WARP=0 switches it off
  bool warp_bvp = !use_add && !use_v3 && !big_n && 3 * n <= 64;
  if (const char *e = getenv("AB200_VL_BVP_WARP")) warp_bvp = warp_bvp && atoi(e) != 0;
  const int wb_wpc = 4;
  const size_t wb_smem_d = (bvp_smem_doubles(n) + (size_t)(13 * n + 32 + 1) / 2 + 2 + 1) & ~(size_t)1;     // per warp, doubles (ints behind)
  const size_t wb_smem = wb_smem_d * 8 * wb_wpc;
  auto kw = (S == 1) ? (const void *)k_vl_bvp_warp<1> : (const void *)k_vl_bvp_warp<3>;
  int wb_per_sm = 1;
  if (warp_bvp) {
    AB_CUDA_OK(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wb_smem));
    AB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wb_per_sm, kw, 32 * wb_wpc, wb_smem));
    wb_per_sm = std::max(1, std::min(wb_per_sm, 16));
  }
  // interaction-principle solver: K2b (one CTA or warp per layer problem) + K3 (persistent CTAs, or one warp per pixel for n <= 16)
  const bool warp_add = use_add && n <= 16;
  const int add_wpc = 8;
  const size_t ops_pw = (ops_smem_doubles(n) + (size_t)(n + 1) / 2 + 1 + 1) & ~(size_t)1, add_pw = (add_smem_doubles(n) + (size_t)(n + 1) / 2 + 1 + 1) & ~(size_t)1;
  const size_t ops_smem = warp_add ? ops_pw * 8 * add_wpc : ops_smem_doubles(n) * 8 + (size_t)n * 4 + 16;
  const size_t add_smem = warp_add ? add_pw * 8 * add_wpc : add_smem_doubles(n) * 8 + (size_t)n * 4 + 16;
  int add_per_sm = 1;
  if (use_add) {
    auto ko = warp_add ? ((S == 1) ? (const void *)k_vl_layer_ops_warp<1> : (const void *)k_vl_layer_ops_warp<3>)
                       : ((S == 1) ? (const void *)k_vl_layer_ops<1> : (const void *)k_vl_layer_ops<3>);
    auto ka = warp_add ? ((S == 1) ? (const void *)k_vl_add_warp<1> : (const void *)k_vl_add_warp<3>)
                       : ((S == 1) ? (const void *)k_vl_add<1> : (const void *)k_vl_add<3>);
    if (ops_smem > 227 * 1024 || add_smem > 227 * 1024) { ab200_set_error("vlidort: nstokes*nstreams too large for the shared-memory solver"); return AB200_ERR_ARG; }
    AB_CUDA_OK(cudaFuncSetAttribute(ko, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ops_smem));
    AB_CUDA_OK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)add_smem));
    AB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&add_per_sm, ka, warp_add ? 32 * add_wpc : 128, add_smem));
    add_per_sm = std::max(1, std::min(add_per_sm, 8));
  }
  // (work buffers: one per concurrently running solver = CTA of the block kernel, or warp of the warp kernel)
  const int bvp_grid_max = use_add ? (warp_add ? nsm * add_per_sm * add_wpc : nsm * add_per_sm) : (warp_bvp ? nsm * wb_per_sm * wb_wpc : nsm * bvp_per_sm);
  const size_t work_per_cta = use_add ? add_work_doubles(n, L) : bvp_work_doubles(n, L);
  const size_t bvp_work_bytes = (size_t)bvp_grid_max * (work_per_cta + (big_n && !use_add ? bvp_window_doubles(n) : 0)) * 8;
  // K2 geometry
  const int scr_d = eigen_scratch_doubles(n, NM);
  // warps (= problems) resident per SM within the 227 KB opt-in limit; two CTAs per SM when possible: problems need a
  // different number of Jacobi sweeps and a CTA only retires when its slowest warp does (measured: 2 x 4 warps beats 1 x 8)
  int wpc = (int)std::min<size_t>(8, (226 * 1024) / ((size_t)scr_d * 8));
  if (wpc >= 4) wpc /= 2;
  if (const char *e = getenv("AB200_VL_EIGEN_WARPS")) wpc = std::max(1, std::min(wpc, atoi(e)));      // tuning knob
  if (wpc < 1) { ab200_set_error("vlidort: nstokes*nstreams too large for the shared-memory eigen scratch"); return AB200_ERR_ARG; }
  // two warps per problem (k_vl_eigen_duo) when the register-resident half-pair Jacobi applies (24 <= n <= 48); AB200_VL_EIGEN_DUO=0 switches it off
  bool duo = n <= 48 && n >= 24;                            // (the row halves of a pair live in registers: 24 rows at most)
  if (const char *e = getenv("AB200_VL_EIGEN_DUO")) duo = duo && atoi(e) != 0;
  const size_t eig_smem = duo ? (size_t)wpc * (scr_d + 4) * 8 : (size_t)wpc * scr_d * 8;
  const bool duo_big = duo && n > 36;                       // own kernel: the 24-row register path must not cost the n <= 36 kernel registers
  if (duo) ke = duo_big ? ((S == 1) ? (const void *)k_vl_eigen_duo<1, 24> : (const void *)k_vl_eigen_duo<3, 24>)
                        : ((S == 1) ? (const void *)k_vl_eigen_duo<1> : (const void *)k_vl_eigen_duo<3>);
  // sub-warp teams for small problems: 8 lanes per problem up to n = 8, 16 up to n = 16 (AB200_VL_EIGEN_SUB=0 switches them off)
  int subw = (n <= 8) ? 8 : (n <= 16 ? 16 : 32);
  if (const char *e = getenv("AB200_VL_EIGEN_SUB")) { if (atoi(e) == 0) subw = 32; }
  const int subs = 32 / subw;
  size_t eig_smem_sub = 0;
  int sub_warps = 4;
  if (subw < 32) {
    ke = (S == 1) ? (subw == 8 ? (const void *)k_vl_eigen_sub<1, 8> : (const void *)k_vl_eigen_sub<1, 16>)
                  : (subw == 8 ? (const void *)k_vl_eigen_sub<3, 8> : (const void *)k_vl_eigen_sub<3, 16>);
    eig_smem_sub = (size_t)sub_warps * subs * scr_d * 8;
  }
  AB_CUDA_OK(cudaFuncSetAttribute(ke, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(subw < 32 ? eig_smem_sub : eig_smem)));
  const size_t setup_smem = (size_t)(12 * (L + 1) + 2 * (R.nMom + 1) + 2 * 4 + 16) * 8;

  size_t chunk = npix;
  if (budget > bvp_work_bytes + (64 << 20)) chunk = std::min<size_t>(npix, (budget - bvp_work_bytes - (64 << 20)) / per_pix);
  else chunk = 1;
  chunk = std::max<size_t>(1, chunk);
  const size_t total_bytes = chunk * per_pix + chunk * 64 + bvp_work_bytes + (size_t)(NM + 1) * (NM + 1) * N * 3 * 8 + (size_t)npix + (64 << 20);
  AB_CUDA_OK(g_arena.reserve(total_bytes));
  DevArena &A = g_arena;
  VLPix X; X.L = L; X.NM = NM; X.S = S; X.npix = (int)chunk;
  X.deltau = A.take<double>(chunk * L); X.og = A.take<double>(chunk * (NM + 1) * L * 4);
  X.avsec = A.take<double>(chunk * L); X.itrans = A.take<double>(chunk * L); X.tdmubar = A.take<double>(chunk * L);
  X.tduserm = A.take<double>(chunk * L); X.emult_up = A.take<double>(chunk * L); X.sigma_p = A.take<double>(chunk * L);
  X.itrans_userm = A.take<double>(chunk * L); X.scal = A.take<double>(chunk * 16); X.ibuf = A.take<int>(chunk * 4);
  X.layscat = A.take<unsigned char>(chunk * (NM + 1) * L);
  VLSol Z; Z.n = n; Z.S = S; Z.N = N;
  const size_t nq = chunk * L;
  Z.keig = A.take<double>(nq * n); Z.tde = A.take<double>(nq * n); Z.xp = A.take<double>(nq * n * n); Z.xn = A.take<double>(nq * n * n);
  Z.uh_dn = A.take<double>(nq * n * S); Z.uh_up = A.take<double>(nq * n * S); Z.hm1 = A.take<double>(nq * n); Z.hm2 = A.take<double>(nq * n);
  Z.aterm = A.take<double>(nq * n); Z.bterm = A.take<double>(nq * n); Z.gam_m = A.take<double>(nq * n); Z.gam_p = A.take<double>(nq * n);
  Z.wupper = A.take<double>(nq * 2 * n); Z.wlower = A.take<double>(nq * 2 * n); Z.fail = A.take<int>(chunk);
  VLOps O; memset(&O, 0, sizeof O);
  if (use_add) {
    O.rk = A.take<double>(nq * n * n); O.tk = A.take<double>(nq * n * n); O.spT = A.take<double>(nq * n * n); O.smT = A.take<double>(nq * n * n);
    O.sgu = A.take<double>(nq * n); O.sgd = A.take<double>(nq * n);
  }
  VLConv CV; CV.total = A.take<double>(chunk * 4); CV.state = A.take<int>(chunk * 4);
  int *listA = A.take<int>(chunk), *listB = A.take<int>(chunk), *counts = A.take<int>(2 * (NM + 4));
  double *d_pq = A.take<double>((size_t)(NM + 1) * N * 3 * (NM + 1));
  double *work = A.take<double>((size_t)bvp_grid_max * work_per_cta);
  double *wglob = (big_n && !use_add) ? A.take<double>((size_t)bvp_grid_max * bvp_window_doubles(n)) : nullptr;
  unsigned char *d_skip = A.take<unsigned char>(npix);
  VLBrdf Bd; VLSurf SFb; memset(&SFb, 0, sizeof SFb);
  if (R.brdf) {
    Bd = *R.brdf;
    Bd.brdf_f = A.take<double>(chunk * brdf_out_f(Bd)); Bd.brdf_f_0 = A.take<double>(chunk * brdf_out_f0(Bd)); Bd.user_brdf_f = A.take<double>(chunk * brdf_out_f0(Bd));
    SFb.brdf_f = Bd.brdf_f; SFb.brdf_f_0 = Bd.brdf_f_0; SFb.user_brdf_f = Bd.user_brdf_f; SFb.user_brdf_f_0 = nullptr; SFb.nmfour = 2 * N; SFb.nq = Bd.nq;
  }
  if (A.off > A.cap) { ab200_set_error("vlidort: internal workspace accounting error"); return AB200_ERR_ARG; }

  // ---- quadrature and the Pi tables at the quadrature cosines for every Fourier component (pixel independent)
  VLQuad QD; memset(&QD, 0, sizeof QD); QD.N = N;
  if (N == 1) { QD.x[0] = std::sqrt(1.0 / 3.0); QD.w[0] = 1.0; } else gauss01(N, QD.x, QD.w);
  {
    std::vector<double> h_pq((size_t)(NM + 1) * N * 3 * (NM + 1), 0.0), tmp(3 * (NM + 2));
    for (int m = 0; m <= NM; m++)
      for (int i = 0; i < N; i++) {
        pi_prt(m, NM, QD.x[i], tmp.data());
        for (int l = m; l <= NM; l++) for (int c = 0; c < 3; c++) h_pq[(size_t)m * (NM + 1) * N * 3 + ((size_t)l * N + i) * 3 + c] = tmp[3 * l + c];
      }
    AB_CUDA_OK(cudaMemcpyAsync(d_pq, h_pq.data(), h_pq.size() * 8, cudaMemcpyHostToDevice, st));
    AB_CUDA_OK(cudaStreamSynchronize(st));
  }
  // ---- skip mask and the FOCORR_NADIR state quirk (first obs with solar_zenith == 0.0 that is actually run)
  int nadir_from = R.nobs;
  {
    const std::vector<unsigned char> &sk = *R.h_skip;
    for (int j = 0; j < R.nobs && nadir_from == R.nobs; j++)
      if ((*R.h_sza)[j] == 0.0) for (int i = 0; i < R.nch; i++) if (!sk[(size_t)j * R.nch + i]) { nadir_from = j; break; }
    AB_CUDA_OK(cudaMemcpyAsync(d_skip, sk.data(), npix, cudaMemcpyHostToDevice, st));
  }
  VLIn in; memset(&in, 0, sizeof in);
  in.km = L; in.nch = R.nch; in.nobs = R.nobs; in.nMom = R.nMom; in.nPol = R.nPol; in.S = S; in.N = N; in.nfine = 3;
  in.planpar = R.planpar; in.lambertian = R.lambertian; in.scalar_glue = R.scalar_glue; in.nadir_from_obs = nadir_from;
  in.MISSING = R.MISSING; in.earth_radius = 6371.0;
  in.ROT = R.ROT; in.depol = R.depol; in.alpha = R.alpha; in.tau = R.tau; in.ssa = R.ssa; in.pmom = R.pmom; in.ze = R.he;
  in.albedo = R.albedo; in.sza = R.sza; in.raa = R.raa; in.vza = R.vza; in.flux_factor = R.flux_factor; in.exactdb = R.exactdb;
  in.skip = d_skip;
  const double accuracy = (double)0.0001f;
  cudaEvent_t ev[4];
  for (auto &e : ev) cudaEventCreate(&e);

  for (size_t p0 = 0; p0 < (size_t)npix; p0 += chunk) {
    const int nc = (int)std::min<size_t>(chunk, npix - p0);
    AB_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int) * 2 * (NM + 4), st));
    cudaEventRecord(ev[0], st);
    if (R.brdf) { k_vl_brdf_fourier<<<nc, 256, brdf_smem, st>>>(Bd, (int)p0, d_skip, brdf_nscr); g_stats.n_launches += 1; }
    k_vl_setup<<<nc, 128, setup_smem, st>>>(in, X, (int)p0);
    k_vl_init_active<<<(nc + 255) / 256, 256, 0, st>>>(X, CV, Z, nc, listA, counts);
    cudaEventRecord(ev[1], st);
    AB_CUDA_OK(cudaGetLastError());
    int nactive = 0;
    AB_CUDA_OK(cudaMemcpyAsync(&nactive, counts, sizeof(int), cudaMemcpyDeviceToHost, st));
    AB_CUDA_OK(cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, ev[0], ev[1]); g_stats.ms_setup += ms;
    // BRDF pointers are indexed with the global pixel inside bvp_pixel via SF; shift them to the chunk
    VLSurf SFc = R.surf;
    if (R.brdf) SFc = SFb;                                   // chunk-local arrays, rewritten for every chunk
    else if (!R.lambertian) {
      SFc.brdf_f += (size_t)p0 * SFc.nmfour * SFc.nq * N * N; SFc.brdf_f_0 += (size_t)p0 * SFc.nmfour * SFc.nq * N;
      SFc.user_brdf_f += (size_t)p0 * SFc.nmfour * SFc.nq * N;
    }
    int *cur = listA, *nxt = listB;
    for (int m = 0; m <= NM && nactive > 0; m++) {
      const long nprob = (long)nactive * L;
      const int egrid = (int)((nprob + wpc - 1) / wpc);
      const double *pqm = d_pq + (size_t)m * (NM + 1) * N * 3;
      cudaEventRecord(ev[0], st);
      if (subw < 32) {
        const int spc = sub_warps * subs, sgrid = (int)((nprob + spc - 1) / spc);
        if (S == 1) { if (subw == 8) k_vl_eigen_sub<1, 8><<<sgrid, 32 * sub_warps, eig_smem_sub, st>>>(X, Z, QD, pqm, m, cur, nactive, scr_d);
                      else k_vl_eigen_sub<1, 16><<<sgrid, 32 * sub_warps, eig_smem_sub, st>>>(X, Z, QD, pqm, m, cur, nactive, scr_d); }
        else { if (subw == 8) k_vl_eigen_sub<3, 8><<<sgrid, 32 * sub_warps, eig_smem_sub, st>>>(X, Z, QD, pqm, m, cur, nactive, scr_d);
               else k_vl_eigen_sub<3, 16><<<sgrid, 32 * sub_warps, eig_smem_sub, st>>>(X, Z, QD, pqm, m, cur, nactive, scr_d); }
      } else if (duo_big) {
        if (S == 1) k_vl_eigen_duo<1, 24><<<egrid, 64 * wpc, eig_smem, st>>>(X, Z, QD, pqm, m, cur, nactive, scr_d);
        else k_vl_eigen_duo<3, 24><<<egrid, 64 * wpc, eig_smem, st>>>(X, Z, QD, pqm, m, cur, nactive, scr_d);
      } else if (duo) {
        if (S == 1) k_vl_eigen_duo<1><<<egrid, 64 * wpc, eig_smem, st>>>(X, Z, QD, pqm, m, cur, nactive, scr_d);
        else k_vl_eigen_duo<3><<<egrid, 64 * wpc, eig_smem, st>>>(X, Z, QD, pqm, m, cur, nactive, scr_d);
      } else if (S == 1) k_vl_eigen<1><<<egrid, 32 * wpc, eig_smem, st>>>(X, Z, QD, pqm, m, cur, nactive, scr_d);
      else k_vl_eigen<3><<<egrid, 32 * wpc, eig_smem, st>>>(X, Z, QD, pqm, m, cur, nactive, scr_d);
      cudaEventRecord(ev[1], st);
      const int bgrid = std::min(nactive, bvp_grid_max);
      int *cnt = counts + 1 + m;
      if (use_add) {
        if (warp_add) {
          const int ogrid = (int)((nprob + add_wpc - 1) / add_wpc), agrid = std::min((nactive + add_wpc - 1) / add_wpc, nsm * add_per_sm);
          if (S == 1) k_vl_layer_ops_warp<1><<<ogrid, 32 * add_wpc, ops_smem, st>>>(X, Z, O, cur, nactive, (int)ops_pw);
          else k_vl_layer_ops_warp<3><<<ogrid, 32 * add_wpc, ops_smem, st>>>(X, Z, O, cur, nactive, (int)ops_pw);
          cudaEventRecord(ev[3], st);
          if (S == 1) k_vl_add_warp<1><<<agrid, 32 * add_wpc, add_smem, st>>>(X, Z, O, QD, SFc, CV, R.lambertian, accuracy, m, cur, nactive, nxt, cnt, work, work_per_cta, (int)add_pw);
          else k_vl_add_warp<3><<<agrid, 32 * add_wpc, add_smem, st>>>(X, Z, O, QD, SFc, CV, R.lambertian, accuracy, m, cur, nactive, nxt, cnt, work, work_per_cta, (int)add_pw);
        } else {
          if (S == 1) k_vl_layer_ops<1><<<(unsigned)nprob, 128, ops_smem, st>>>(X, Z, O, cur, nactive);
          else k_vl_layer_ops<3><<<(unsigned)nprob, 128, ops_smem, st>>>(X, Z, O, cur, nactive);
          cudaEventRecord(ev[3], st);
          if (S == 1) k_vl_add<1><<<bgrid, 128, add_smem, st>>>(X, Z, O, QD, SFc, CV, R.lambertian, accuracy, m, cur, nactive, nxt, cnt, work, work_per_cta);
          else k_vl_add<3><<<bgrid, 128, add_smem, st>>>(X, Z, O, QD, SFc, CV, R.lambertian, accuracy, m, cur, nactive, nxt, cnt, work, work_per_cta);
        }
        g_stats.n_launches += 1;
        // executed flops of the two kernels: 2 x (3 n^3 elimination + 2 n^3 product) per layer, and 3 n^3 + 3 x 2 n^3 per layer
        g_stats.flops_exec += (double)nprob * 19.0 * (double)n * n * n;
      } else if (warp_bvp) {
        const int wgrid = std::min((nactive + wb_wpc - 1) / wb_wpc, nsm * wb_per_sm);
        if (S == 1) k_vl_bvp_warp<1><<<wgrid, 32 * wb_wpc, wb_smem, st>>>(X, Z, QD, SFc, CV, R.lambertian, accuracy, m, cur, nactive, nxt, cnt, work, work_per_cta, (int)wb_smem_d);
        else k_vl_bvp_warp<3><<<wgrid, 32 * wb_wpc, wb_smem, st>>>(X, Z, QD, SFc, CV, R.lambertian, accuracy, m, cur, nactive, nxt, cnt, work, work_per_cta, (int)wb_smem_d);
      } else
      kbf<<<bgrid, bvp_threads, bvp_smem, st>>>(X, Z, QD, SFc, CV, R.lambertian, accuracy, m, cur, nactive, nxt, cnt, work, work_per_cta, use_v3, (int)bvp_smem_d, wglob);
      cudaEventRecord(ev[2], st);
      AB_CUDA_OK(cudaGetLastError());
      g_stats.n_eigen_problems += nprob; g_stats.n_bvp_problems += nactive; g_stats.n_launches += 2;
      {  // algorithmic flops of SURVEY.md 8(d): eigen part (2+25+2) n^3 + 4 S^3 N^2 (2N-m); BVP part (14/3)(2n)^3 + 12 (2n)^2
        const double n3 = (double)n * n * n;
        g_stats.flops_eigen += (double)nprob * (29.0 * n3 + 4.0 * S * S * S * (double)N * N * (2 * N - m));
        g_stats.flops_bvp += (double)nprob * ((14.0 / 3.0) * 8.0 * n3 + 12.0 * 4.0 * n * n);
      }
      AB_CUDA_OK(cudaMemcpyAsync(&nactive, cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
      AB_CUDA_OK(cudaStreamSynchronize(st));
      float m1 = 0, m2 = 0, m3 = 0; cudaEventElapsedTime(&m1, ev[0], ev[1]); cudaEventElapsedTime(&m2, ev[1], ev[2]);
      if (use_add) { cudaEventElapsedTime(&m3, ev[1], ev[3]); g_stats.ms_ops += m3; }
      g_stats.ms_eigen += m1; g_stats.ms_bvp += m2;
      std::swap(cur, nxt);
    }
    k_vl_finish<<<(nc + 255) / 256, 256, 0, st>>>(X, CV, (int)p0, nc, R.nch, R.nobs, R.sza, R.MISSING, S, R.radiance, R.reflectance, R.Q, R.U,
                                                  R.nfourier, R.status);
    AB_CUDA_OK(cudaGetLastError());
    AB_CUDA_OK(cudaStreamSynchronize(st));
  }
  for (auto &e : ev) cudaEventDestroy(e);
  g_stats.npix += npix; g_stats.n_launches += 3 * ((npix + chunk - 1) / chunk);
  return 0;
}

bool is_missing_h(double x, double MISSING) { return std::fabs(x / MISSING - 1.0) < (double)0.001f; }

// host-pointer front end shared by the Lambertian entry points
int lambert_host(int S, int scalar_glue, int km, int nch, int nobs, int nstreams, int plane_parallel, int nMom, int nPol,
                 const double *ROT, const double *depol, const double *alpha, const double *tau, const double *ssa,
                 const double *pmom, const double *he, const double *albedo, const double *sza, const double *raa,
                 const double *vza, const double *flux_factor, double MISSING, double *radiance, double *reflectance,
                 double *Q, double *U, int *rc) {
  *rc = 0;
  // VLIDORT_Init dimension checks (V283/VLIDORT_Mod.F90:366-374; GEOS-sized maxima of vlidort_pars.f90:136-199)
  if (nstreams > 25) { *rc = 2; return 2; }
  if (km > 72) { *rc = 3; return 3; }
  if (nMom > 1000) { *rc = 5; return 5; }
  const size_t npix = (size_t)nobs * nch;
  if (npix == 0) return 0;
  const size_t n_rot = (size_t)km * nobs * nch, n_pm = n_rot * nMom * nPol, n_he = (size_t)(km + 1) * nobs;
  size_t bytes = 8 * (n_rot * (alpha ? 4 : 3) + n_pm + n_he + npix * (1 + (flux_factor ? 1 : 0)) + 3 * (size_t)nobs + nch + 4 * npix) + 8 * npix + 4096 * 16;
  static DevBuf io;
  AB_CUDA_OK(io.reserve(bytes));
  char *b = io.as<char>(); size_t off = 0;
  auto put = [&](const double *src, size_t n) -> double * { off = align256(off); double *d = (double *)(b + off); off += n * 8; if (src) cudaMemcpyAsync(d, src, n * 8, cudaMemcpyHostToDevice, 0); return d; };
  RunDesc R; memset(&R, 0, sizeof R);
  R.S = S; R.N = nstreams; R.km = km; R.nch = nch; R.nobs = nobs; R.nMom = nMom; R.nPol = nPol; R.planpar = plane_parallel;
  R.lambertian = 1; R.scalar_glue = scalar_glue; R.MISSING = MISSING;
  R.ROT = put(ROT, n_rot); R.depol = put(depol, nch); R.alpha = alpha ? put(alpha, n_rot) : nullptr; R.tau = put(tau, n_rot); R.ssa = put(ssa, n_rot);
  R.pmom = put(pmom, n_pm); R.he = put(he, n_he); R.albedo = put(albedo, npix); R.sza = put(sza, nobs); R.raa = put(raa, nobs); R.vza = put(vza, nobs);
  R.flux_factor = flux_factor ? put(flux_factor, npix) : nullptr;
  R.radiance = put(nullptr, npix); R.reflectance = put(nullptr, npix); R.Q = put(nullptr, npix); R.U = put(nullptr, npix);
  off = align256(off); R.nfourier = (int *)(b + off); off += 4 * npix; off = align256(off); R.status = (int *)(b + off); off += 4 * npix;
  std::vector<double> h_sza(sza, sza + nobs);
  std::vector<unsigned char> skip(npix, 0);
  for (int j = 0; j < nobs; j++) {
    const bool ma = is_missing_h(sza[j], MISSING) || is_missing_h(vza[j], MISSING) || is_missing_h(raa[j], MISSING);
    for (int i = 0; i < nch; i++) skip[(size_t)j * nch + i] = (ma || is_missing_h(albedo[j + (size_t)nobs * i], MISSING)) ? 1 : 0;
  }
  R.h_sza = &h_sza; R.h_skip = &skip; R.stream = 0;
  int e = run_vlidort(R);
  if (e) return e;
  AB_CUDA_OK(cudaMemcpyAsync(radiance, R.radiance, 8 * npix, cudaMemcpyDeviceToHost, 0));
  AB_CUDA_OK(cudaMemcpyAsync(reflectance, R.reflectance, 8 * npix, cudaMemcpyDeviceToHost, 0));
  if (Q) AB_CUDA_OK(cudaMemcpyAsync(Q, R.Q, 8 * npix, cudaMemcpyDeviceToHost, 0));
  if (U) AB_CUDA_OK(cudaMemcpyAsync(U, R.U, 8 * npix, cudaMemcpyDeviceToHost, 0));
  g_last_nfourier.resize(npix); g_last_status.resize(npix);
  AB_CUDA_OK(cudaMemcpyAsync(g_last_nfourier.data(), R.nfourier, 4 * npix, cudaMemcpyDeviceToHost, 0));
  AB_CUDA_OK(cudaMemcpyAsync(g_last_status.data(), R.status, 4 * npix, cudaMemcpyDeviceToHost, 0));
  AB_CUDA_OK(cudaStreamSynchronize(0));
  return 0;
}


// ------------------------------------------------------------------------------------------------ BRDF wrappers
struct SurfArgs { int kind; const double *u10m, *v10m, *mr, *kernel_wt, *param; double *BR, *BR_Q, *BR_U; const double *alpha = nullptr, *flux_factor = nullptr; };

// pixel-independent BRDF tables -> device
int build_brdf_tables(VLBrdf &B, int kind, int S, int scalar_glue, int N, int NM, int nch, int nobs, cudaStream_t st) {
  std::vector<double> h;
  brdf_host_tables(B, kind, S, scalar_glue, N, NM, nch, nobs, h);
  static DevBuf tab;
  AB_CUDA_OK(tab.reserve(h.size() * 8));
  AB_CUDA_OK(cudaMemcpyAsync(tab.as<double>(), h.data(), h.size() * 8, cudaMemcpyHostToDevice, st));
  AB_CUDA_OK(cudaStreamSynchronize(st));
  brdf_rebase_tables(B, h.data(), tab.as<double>());
  return 0;
}

// front end of the Cox-Munk / GISS Cox-Munk / MODIS RTLS wrappers; dev: every array pointer is a device pointer
int brdf_front(bool dev, int S, int scalar_glue, int km, int nch, int nobs, int nstreams, int plane_parallel, int nMom, int nPol,
               const double *ROT, const double *depol, const double *tau, const double *ssa, const double *pmom,
               const double *he, const double *sza, const double *raa, const double *vza, double MISSING, const SurfArgs &sa,
               double *radiance, double *reflectance, double *Q, double *U, cudaStream_t st, int *rc) {
  *rc = 0;
  if (nstreams > 25) { *rc = 2; return 2; }
  if (km > 72) { *rc = 3; return 3; }
  if (nMom > 1000) { *rc = 5; return 5; }
  const size_t npix = (size_t)nobs * nch;
  if (npix == 0) return 0;
  const bool cx = (sa.kind != SURF_MODIS);
  const size_t n_rot = (size_t)km * nobs * nch, n_pm = n_rot * nMom * nPol, n_he = (size_t)(km + 1) * nobs;
  const size_t bytes = 8 * (n_rot * 4 + n_pm + n_he + 5 * (size_t)nobs + 2 * nch + 6 * npix + 7 * npix + 16 * npix) + 16 * npix + 4096 * 32;
  static DevBuf io;
  AB_CUDA_OK(io.reserve(dev ? 8 * (16 * npix + 7 * npix) + 16 * npix + 4096 * 8 : bytes));
  char *b = io.as<char>(); size_t off = 0;
  auto put = [&](const double *src, size_t n) -> double * {
    if (dev && src) return const_cast<double *>(src);
    off = align256(off); double *d = (double *)(b + off); off += n * 8;
    if (src) cudaMemcpyAsync(d, src, n * 8, cudaMemcpyHostToDevice, st);
    return d;
  };
  RunDesc R; memset(&R, 0, sizeof R);
  R.S = S; R.N = nstreams; R.km = km; R.nch = nch; R.nobs = nobs; R.nMom = nMom; R.nPol = nPol; R.planpar = plane_parallel;
  R.lambertian = 0; R.scalar_glue = scalar_glue; R.MISSING = MISSING;
  R.ROT = put(ROT, n_rot); R.depol = put(depol, nch); R.tau = put(tau, n_rot); R.ssa = put(ssa, n_rot);
  R.pmom = put(pmom, n_pm); R.he = put(he, n_he); R.sza = put(sza, nobs); R.raa = put(raa, nobs); R.vza = put(vza, nobs);
  R.alpha = sa.alpha ? put(sa.alpha, n_rot) : nullptr; R.flux_factor = sa.flux_factor ? put(sa.flux_factor, npix) : nullptr;
  const double *d_u = cx ? put(sa.u10m, nobs) : nullptr, *d_v = cx ? put(sa.v10m, nobs) : nullptr, *d_mr = cx ? put(sa.mr, nch) : nullptr;
  const double *d_kw = cx ? nullptr : put(sa.kernel_wt, 3 * npix), *d_pr = cx ? nullptr : put(sa.param, 2 * npix);
  double *d_exact = put(nullptr, 16 * npix);
  double *d_BR, *d_BRQ, *d_BRU;
  if (dev) {
    R.radiance = radiance; R.reflectance = reflectance; R.Q = Q; R.U = U; d_BR = sa.BR; d_BRQ = sa.BR_Q; d_BRU = sa.BR_U;
  } else {
    R.radiance = put(nullptr, npix); R.reflectance = put(nullptr, npix); R.Q = put(nullptr, npix); R.U = put(nullptr, npix);
    d_BR = put(nullptr, npix); d_BRQ = put(nullptr, npix); d_BRU = put(nullptr, npix);
  }
  if (!d_BR) d_BR = put(nullptr, npix);
  off = align256(off); R.nfourier = (int *)(b + off); off += 4 * npix; off = align256(off); R.status = (int *)(b + off); off += 4 * npix;
  off = align256(off); unsigned char *d_skip0 = (unsigned char *)(b + off); off += npix;
  // host copies for the MISSING screening
  std::vector<double> h_sza(nobs), h_vza(nobs), h_raa(nobs), h_u, h_v, h_mr;
  auto fetch = [&](std::vector<double> &hv, const double *src, size_t n) {
    hv.resize(n);
    if (dev) cudaMemcpyAsync(hv.data(), src, 8 * n, cudaMemcpyDeviceToHost, st); else std::copy(src, src + n, hv.begin());
  };
  fetch(h_sza, sza, nobs); fetch(h_vza, vza, nobs); fetch(h_raa, raa, nobs);
  if (cx) { fetch(h_u, sa.u10m, nobs); fetch(h_v, sa.v10m, nobs); fetch(h_mr, sa.mr, nch); }
  if (dev) AB_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<unsigned char> skip(npix, 0);
  for (int j = 0; j < nobs; j++) {
    const bool ma = is_missing_h(h_sza[j], MISSING) || is_missing_h(h_vza[j], MISSING) || is_missing_h(h_raa[j], MISSING);
    const bool mw = cx && (is_missing_h(h_u[j], MISSING) || is_missing_h(h_v[j], MISSING));
    for (int i = 0; i < nch; i++) skip[(size_t)j * nch + i] = (ma || mw || (cx && is_missing_h(h_mr[i], MISSING))) ? 1 : 0;
  }
  AB_CUDA_OK(cudaMemcpyAsync(d_skip0, skip.data(), npix, cudaMemcpyHostToDevice, st));
  VLBrdf B;
  const int NM = std::min(2 * nstreams - 1, nMom);
  if (int e = build_brdf_tables(B, sa.kind, S, scalar_glue, nstreams, NM, nch, nobs, st)) return e;
  B.sza = R.sza; B.vza = R.vza; B.raa = R.raa; B.u10m = d_u; B.v10m = d_v; B.mr = d_mr; B.kernel_wt = d_kw; B.param = d_pr;
  B.exactdb = d_exact;
  off = align256(off); B.wsa_fail = (int *)(b + off); off += 4 * npix;
  k_vl_brdf_exact<<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(B, (int)npix, d_skip0, d_BR, d_BRQ, d_BRU);
  AB_CUDA_OK(cudaGetLastError());
  g_stats.n_launches += 1;
  if (sa.kind == SURF_MODIS) {                         // BR < 0 -> every output -500, pixel not run
    std::vector<double> h_br(npix);
    AB_CUDA_OK(cudaMemcpyAsync(h_br.data(), d_BR, 8 * npix, cudaMemcpyDeviceToHost, st));
    AB_CUDA_OK(cudaStreamSynchronize(st));
    for (int j = 0; j < nobs; j++) for (int i = 0; i < nch; i++) if (!skip[(size_t)j * nch + i] && h_br[j + (size_t)nobs * i] < 0) skip[(size_t)j * nch + i] = 2;
  }
  R.exactdb = d_exact; R.brdf = &B;
  R.h_sza = &h_sza; R.h_skip = &skip; R.stream = st;
  if (int e = run_vlidort(R)) return e;
  g_last_nfourier.resize(npix); g_last_status.resize(npix);
  AB_CUDA_OK(cudaMemcpyAsync(g_last_nfourier.data(), R.nfourier, 4 * npix, cudaMemcpyDeviceToHost, st));
  AB_CUDA_OK(cudaMemcpyAsync(g_last_status.data(), R.status, 4 * npix, cudaMemcpyDeviceToHost, st));
  if (!dev) {
    AB_CUDA_OK(cudaMemcpyAsync(radiance, R.radiance, 8 * npix, cudaMemcpyDeviceToHost, st));
    AB_CUDA_OK(cudaMemcpyAsync(reflectance, R.reflectance, 8 * npix, cudaMemcpyDeviceToHost, st));
    if (Q) AB_CUDA_OK(cudaMemcpyAsync(Q, R.Q, 8 * npix, cudaMemcpyDeviceToHost, st));
    if (U) AB_CUDA_OK(cudaMemcpyAsync(U, R.U, 8 * npix, cudaMemcpyDeviceToHost, st));
    if (sa.BR) AB_CUDA_OK(cudaMemcpyAsync(sa.BR, d_BR, 8 * npix, cudaMemcpyDeviceToHost, st));
    if (sa.BR_Q) AB_CUDA_OK(cudaMemcpyAsync(sa.BR_Q, d_BRQ, 8 * npix, cudaMemcpyDeviceToHost, st));
    if (sa.BR_U) AB_CUDA_OK(cudaMemcpyAsync(sa.BR_U, d_BRU, 8 * npix, cudaMemcpyDeviceToHost, st));
  }
  AB_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int ab200_vlidort_scalar_lambert(int km, int nch, int nobs, const double *channels, int nstreams,
                                            int plane_parallel, int nMom, int nPol, const double *ROT, const double *depol,
                                            const double *tau, const double *ssa, const double *g, const double *pmom,
                                            const double *pe, const double *he, const double *te, const double *albedo,
                                            const double *solar_zenith, const double *relat_azymuth,
                                            const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL,
                                            double *reflectance_VL, int *rc) {
  (void)channels; (void)g; (void)pe; (void)te; (void)verbose;
  return lambert_host(1, 1, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, nullptr, tau, ssa, pmom, he, albedo,
                      solar_zenith, relat_azymuth, sensor_zenith, nullptr, MISSING, radiance_VL, reflectance_VL, nullptr, nullptr, rc);
}

extern "C" int ab200_vlidort_vector_lambert(int km, int nch, int nobs, const double *channels, int nstreams,
                                            int plane_parallel, int nMom, int nPol, const double *ROT, const double *depol,
                                            const double *alpha, const double *tau, const double *ssa, const double *pmom,
                                            const double *pe, const double *he, const double *te, const double *albedo,
                                            const double *solar_zenith, const double *relat_azymuth,
                                            const double *sensor_zenith, const double *flux_factor, double MISSING,
                                            int verbose, double *radiance_VL, double *reflectance_VL, double *Q, double *U,
                                            int *rc) {
  (void)channels; (void)pe; (void)te; (void)verbose;
  return lambert_host(3, 0, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, alpha, tau, ssa, pmom, he, albedo,
                      solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, radiance_VL, reflectance_VL, Q, U, rc);
}

extern "C" int ab200_rot_calc(int km, int nch, int nobs, const double *channels, const double *pe, const double *he,
                              const double *te, double MISSING, int verbose, double *ROT, double *depol, int *rc) {
  (void)MISSING; (void)verbose;
  *rc = 0;
  if (km > 72) { *rc = 3; return 3; }
  const size_t ne = (size_t)(km + 1) * nobs, nr = (size_t)km * nobs * nch;
  if (nr == 0) return 0;
  static DevBuf io;
  AB_CUDA_OK(io.reserve(8 * (3 * ne + nr + 2 * nch) + 4096));
  double *d = io.as<double>();
  double *d_pe = d, *d_he = d + ne, *d_te = d + 2 * ne, *d_ch = d + 3 * ne, *d_dp = d_ch + nch, *d_rot = d_dp + nch;
  cudaMemcpyAsync(d_pe, pe, 8 * ne, cudaMemcpyHostToDevice, 0); cudaMemcpyAsync(d_he, he, 8 * ne, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(d_te, te, 8 * ne, cudaMemcpyHostToDevice, 0); cudaMemcpyAsync(d_ch, channels, 8 * nch, cudaMemcpyHostToDevice, 0);
  k_rot_calc<<<(unsigned)((nr + 255) / 256), 256>>>(km, nch, nobs, d_ch, d_pe, d_he, d_te, d_rot, d_dp);
  AB_CUDA_OK(cudaGetLastError());
  cudaMemcpyAsync(ROT, d_rot, 8 * nr, cudaMemcpyDeviceToHost, 0); cudaMemcpyAsync(depol, d_dp, 8 * nch, cudaMemcpyDeviceToHost, 0);
  AB_CUDA_OK(cudaStreamSynchronize(0));
  return 0;
}

// device-pointer variant of ab200_rot_calc (every array resident in HBM, kernel launched on `stream`)
extern "C" int ab200_rot_calc_dev(int km, int nch, int nobs, const double *channels, const double *pe, const double *he,
                                  const double *te, double *ROT, double *depol, void *stream) {
  if (km > 72) return 3;
  const size_t nr = (size_t)km * nobs * nch;
  if (nr == 0) return 0;
  k_rot_calc<<<(unsigned)((nr + 255) / 256), 256, 0, (cudaStream_t)stream>>>(km, nch, nobs, channels, pe, he, te, ROT, depol);
  AB_CUDA_OK(cudaGetLastError());
  return 0;
}

// diagnostics of the most recent call: per-pixel Fourier count (FOURIER_SAVED) and status, p = j*nch + i
extern "C" long ab200_vlidort_last_diag(int *nfourier_out, int *status_out, long cap) {
  const long n = (long)g_last_nfourier.size();
  for (long k = 0; k < n && k < cap; k++) { if (nfourier_out) nfourier_out[k] = g_last_nfourier[k]; if (status_out) status_out[k] = g_last_status[k]; }
  return n;
}

extern "C" int ab200_vlidort_stats2(double *flops_eigen, double *flops_bvp, long *n_launches) {
  if (flops_eigen) *flops_eigen = g_stats.flops_eigen;
  if (flops_bvp) *flops_bvp = g_stats.flops_bvp;
  if (n_launches) *n_launches = g_stats.n_launches;
  return 0;
}

// ms_ops: the share of ms_bvp spent in K2b (layer interaction operators); flops_exec: flops the interaction-principle kernels
// actually execute (flops_bvp stays the SURVEY 8(d) band-LU count the roofline is quoted on)
extern "C" int ab200_vlidort_stats3(double *ms_ops, double *flops_exec) {
  if (ms_ops) *ms_ops = g_stats.ms_ops;
  if (flops_exec) *flops_exec = g_stats.flops_exec;
  return 0;
}

// device-pointer variant of ab200_vlidort_vector_lambert: every array already resident in HBM (same layouts)
extern "C" int ab200_vlidort_vector_lambert_dev(int km, int nch, int nobs, int nstreams, int plane_parallel, int nMom, int nPol,
                                                const double *ROT, const double *depol, const double *alpha,
                                                const double *tau, const double *ssa, const double *pmom, const double *he,
                                                const double *albedo, const double *sza, const double *raa, const double *vza,
                                                const double *flux_factor, double MISSING, double *radiance,
                                                double *reflectance, double *Q, double *U, void *stream, int *rc) {
  *rc = 0;
  if (nstreams > 25) { *rc = 2; return 2; }
  if (km > 72) { *rc = 3; return 3; }
  if (nMom > 1000) { *rc = 5; return 5; }
  const size_t npix = (size_t)nobs * nch;
  if (npix == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<double> h_sza(nobs), h_vza(nobs), h_raa(nobs), h_alb(npix);
  AB_CUDA_OK(cudaMemcpyAsync(h_sza.data(), sza, 8 * (size_t)nobs, cudaMemcpyDeviceToHost, st));
  AB_CUDA_OK(cudaMemcpyAsync(h_vza.data(), vza, 8 * (size_t)nobs, cudaMemcpyDeviceToHost, st));
  AB_CUDA_OK(cudaMemcpyAsync(h_raa.data(), raa, 8 * (size_t)nobs, cudaMemcpyDeviceToHost, st));
  AB_CUDA_OK(cudaMemcpyAsync(h_alb.data(), albedo, 8 * npix, cudaMemcpyDeviceToHost, st));
  AB_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<unsigned char> skip(npix, 0);
  for (int j = 0; j < nobs; j++) {
    const bool ma = is_missing_h(h_sza[j], MISSING) || is_missing_h(h_vza[j], MISSING) || is_missing_h(h_raa[j], MISSING);
    for (int i = 0; i < nch; i++) skip[(size_t)j * nch + i] = (ma || is_missing_h(h_alb[j + (size_t)nobs * i], MISSING)) ? 1 : 0;
  }
  static DevBuf diag;
  AB_CUDA_OK(diag.reserve(8 * npix + 512));
  RunDesc R; memset(&R, 0, sizeof R);
  R.S = 3; R.N = nstreams; R.km = km; R.nch = nch; R.nobs = nobs; R.nMom = nMom; R.nPol = nPol; R.planpar = plane_parallel;
  R.lambertian = 1; R.scalar_glue = 0; R.MISSING = MISSING;
  R.ROT = ROT; R.depol = depol; R.alpha = alpha; R.tau = tau; R.ssa = ssa; R.pmom = pmom; R.he = he; R.albedo = albedo;
  R.sza = sza; R.raa = raa; R.vza = vza; R.flux_factor = flux_factor;
  R.radiance = radiance; R.reflectance = reflectance; R.Q = Q; R.U = U;
  R.nfourier = diag.as<int>(); R.status = diag.as<int>() + npix;
  R.h_sza = &h_sza; R.h_skip = &skip; R.stream = st;
  return run_vlidort(R);
}

// ---- BRDF surfaces (argument lists follow LV/VLIDORT_POLAR_py.F90:137-213 VECTOR_BRDF_MODIS, :673-745 VECTOR_GissCX,
//      :830-902 VECTOR_CX; scalar forms GV/VLIDORT_BRDF_CX.F90:420-584, GV/VLIDORT_BRDF_MODIS.F90:35-202)
#define AB_CX_ARGS int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom, int nPol, \
  const double *ROT, const double *depol, const double *tau, const double *ssa, const double *pmom, const double *pe,   \
  const double *he, const double *te, const double *U10m, const double *V10m, const double *mr,                         \
  const double *solar_zenith, const double *relat_azymuth, const double *sensor_zenith, double MISSING, int verbose
extern "C" int ab200_vlidort_vector_cx(AB_CX_ARGS, double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q,
                                       double *U, double *BR_Q, double *BR_U, int *rc) {
  (void)channels; (void)pe; (void)te; (void)verbose;
  SurfArgs sa = {SURF_CX, U10m, V10m, mr, nullptr, nullptr, BR, BR_Q, BR_U};
  return brdf_front(false, 3, 0, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, he, solar_zenith,
                    relat_azymuth, sensor_zenith, MISSING, sa, radiance_VL_SURF, reflectance_VL_SURF, Q, U, 0, rc);
}
extern "C" int ab200_vlidort_vector_gisscx(AB_CX_ARGS, double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR,
                                           double *Q, double *U, double *BR_Q, double *BR_U, int *rc) {
  (void)channels; (void)pe; (void)te; (void)verbose;
  SurfArgs sa = {SURF_GISSCX, U10m, V10m, mr, nullptr, nullptr, BR, BR_Q, BR_U};
  return brdf_front(false, 3, 0, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, he, solar_zenith,
                    relat_azymuth, sensor_zenith, MISSING, sa, radiance_VL_SURF, reflectance_VL_SURF, Q, U, 0, rc);
}
extern "C" int ab200_vlidort_vector_ocicx(AB_CX_ARGS, double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q,
                                          double *U, double *BR_Q, double *BR_U, int *rc) {
  (void)channels; (void)pe; (void)te; (void)verbose;
  SurfArgs sa = {SURF_OCICX, U10m, V10m, mr, nullptr, nullptr, BR, BR_Q, BR_U};
  return brdf_front(false, 3, 0, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, he, solar_zenith,
                    relat_azymuth, sensor_zenith, MISSING, sa, radiance_VL_SURF, reflectance_VL_SURF, Q, U, 0, rc);
}
extern "C" int ab200_vlidort_vector_ocigisscx(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                              int nMom, int nPol, const double *ROT, const double *depol, const double *alpha,
                                              const double *tau, const double *ssa, const double *pmom, const double *pe,
                                              const double *he, const double *te, const double *U10m, const double *V10m,
                                              const double *mr, const double *solar_zenith, const double *relat_azymuth,
                                              const double *sensor_zenith, const double *flux_factor, double MISSING, int verbose,
                                              double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BR, double *Q,
                                              double *U, double *BR_Q, double *BR_U, int *rc) {
  (void)channels; (void)pe; (void)te; (void)verbose;
  SurfArgs sa = {SURF_OCIGISSCX, U10m, V10m, mr, nullptr, nullptr, BR, BR_Q, BR_U, alpha, flux_factor};
  return brdf_front(false, 3, 0, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, he, solar_zenith,
                    relat_azymuth, sensor_zenith, MISSING, sa, radiance_VL_SURF, reflectance_VL_SURF, Q, U, 0, rc);
}
extern "C" int ab200_vlidort_scalar_gisscx(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                           int nMom, int nPol, const double *ROT, const double *depol, const double *tau,
                                           const double *ssa, const double *g, const double *pmom, const double *pe,
                                           const double *he, const double *te, const double *U10m, const double *V10m,
                                           const double *mr, const double *solar_zenith, const double *relat_azymuth,
                                           const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                                           double *reflectance_VL_SURF, double *BRDF, int *rc) {
  (void)channels; (void)g; (void)pe; (void)te; (void)verbose;
  SurfArgs sa = {SURF_GISSCX, U10m, V10m, mr, nullptr, nullptr, BRDF, nullptr, nullptr};
  return brdf_front(false, 1, 1, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, he, solar_zenith,
                    relat_azymuth, sensor_zenith, MISSING, sa, radiance_VL_SURF, reflectance_VL_SURF, nullptr, nullptr, 0, rc);
}
extern "C" int ab200_vlidort_scalar_cx(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel, int nMom,
                                       int nPol, const double *ROT, const double *depol, const double *tau, const double *ssa,
                                       const double *g, const double *pmom, const double *pe, const double *he, const double *te,
                                       const double *U10m, const double *V10m, const double *mr, const double *solar_zenith,
                                       const double *relat_azymuth, const double *sensor_zenith, double MISSING, int verbose,
                                       double *radiance_VL_SURF, double *reflectance_VL_SURF, double *BRDF, int *rc) {
  (void)channels; (void)g; (void)pe; (void)te; (void)verbose;
  SurfArgs sa = {SURF_CX, U10m, V10m, mr, nullptr, nullptr, BRDF, nullptr, nullptr};
  return brdf_front(false, 1, 1, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, he, solar_zenith,
                    relat_azymuth, sensor_zenith, MISSING, sa, radiance_VL_SURF, reflectance_VL_SURF, nullptr, nullptr, 0, rc);
}
extern "C" int ab200_vlidort_vector_brdf_modis(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                               int nMom, int nPol, int nkernel, int nparam, const double *ROT, const double *depol,
                                               const double *tau, const double *ssa, const double *pmom, const double *pe,
                                               const double *he, const double *te, const double *kernel_wt, const double *param,
                                               const double *solar_zenith, const double *relat_azymuth,
                                               const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                                               double *reflectance_VL_SURF, double *BR, double *Q, double *U, double *BR_Q,
                                               double *BR_U, int *rc) {
  (void)channels; (void)pe; (void)te; (void)verbose;
  if (nkernel != 3 || nparam != 2) { ab200_set_error("vector_brdf_modis: nkernel must be 3 and nparam 2 (GV/VLIDORT_BRDF_MODIS.F90:224-225)"); return AB200_ERR_ARG; }
  SurfArgs sa = {SURF_MODIS, nullptr, nullptr, nullptr, kernel_wt, param, BR, BR_Q, BR_U};
  return brdf_front(false, 3, 0, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, he, solar_zenith,
                    relat_azymuth, sensor_zenith, MISSING, sa, radiance_VL_SURF, reflectance_VL_SURF, Q, U, 0, rc);
}
extern "C" int ab200_vlidort_scalar_brdf_modis(int km, int nch, int nobs, const double *channels, int nstreams, int plane_parallel,
                                               int nMom, int nPol, const double *ROT, const double *depol, const double *tau,
                                               const double *ssa, const double *g, const double *pmom, const double *pe,
                                               const double *he, const double *te, const double *kernel_wt, const double *param,
                                               const double *solar_zenith, const double *relat_azymuth,
                                               const double *sensor_zenith, double MISSING, int verbose, double *radiance_VL_SURF,
                                               double *reflectance_VL_SURF, double *BR, int *rc) {
  (void)channels; (void)g; (void)pe; (void)te; (void)verbose;
  SurfArgs sa = {SURF_MODIS, nullptr, nullptr, nullptr, kernel_wt, param, BR, nullptr, nullptr};
  return brdf_front(false, 1, 1, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, he, solar_zenith,
                    relat_azymuth, sensor_zenith, MISSING, sa, radiance_VL_SURF, reflectance_VL_SURF, nullptr, nullptr, 0, rc);
}
// device-pointer variant of ab200_vlidort_vector_cx (every array resident in HBM, outputs too)
extern "C" int ab200_vlidort_vector_cx_dev(int km, int nch, int nobs, int nstreams, int plane_parallel, int nMom, int nPol,
                                           const double *ROT, const double *depol, const double *tau, const double *ssa,
                                           const double *pmom, const double *he, const double *U10m, const double *V10m,
                                           const double *mr, const double *sza, const double *raa, const double *vza,
                                           double MISSING, double *radiance, double *reflectance, double *BR, double *Q, double *U,
                                           void *stream, int *rc) {
  SurfArgs sa = {SURF_CX, U10m, V10m, mr, nullptr, nullptr, BR, nullptr, nullptr};
  return brdf_front(true, 3, 0, km, nch, nobs, nstreams, plane_parallel, nMom, nPol, ROT, depol, tau, ssa, pmom, he, sza, raa, vza,
                    MISSING, sa, radiance, reflectance, Q, U, (cudaStream_t)stream, rc);
}

extern "C" int ab200_vlidort_stats(double *ms_setup, double *ms_eigen, double *ms_bvp, long *n_eigen, long *n_bvp, long *npix,
                                   int reset) {
  if (ms_setup) *ms_setup = g_stats.ms_setup;
  if (ms_eigen) *ms_eigen = g_stats.ms_eigen;
  if (ms_bvp) *ms_bvp = g_stats.ms_bvp;
  if (n_eigen) *n_eigen = g_stats.n_eigen_problems;
  if (n_bvp) *n_bvp = g_stats.n_bvp_problems;
  if (npix) *npix = g_stats.npix;
  if (reset) memset(&g_stats, 0, sizeof g_stats);
  return 0;
}

extern "C" int ab200_vlidort_release(void) {
  if (g_arena.base) cudaFree(g_arena.base);
  g_arena.base = nullptr; g_arena.cap = 0; g_arena.off = 0;
  return 0;
}

#ifdef AB200_BVP_TIMING
// debug build only: summed clock64 deltas per K3 phase (thread 0 of every CTA); reading clears them
extern "C" int ab200_vlidort_debug_timing(unsigned long long *out16) {
  unsigned long long z[16] = {0};
  AB_CUDA_OK(cudaMemcpyFromSymbol(out16, vlk::g_bvp_tim, sizeof z));
  AB_CUDA_OK(cudaMemcpyToSymbol(vlk::g_bvp_tim, z, sizeof z));
  return 0;
}
#endif
