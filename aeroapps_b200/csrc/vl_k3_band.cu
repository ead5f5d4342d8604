// Stage B, kernel K3 in its band form (vl_bvp.cuh: windowed elimination of the reference's band matrix with the pivot candidates DGBTRF
// sees) -- kernel wrappers and launch plan.  Used for nstokes*nstreams > 64 and with AB200_VL_BVP_LEGACY=1 (A/B against vl_add.cuh).
#include "vl_launch.h"
#ifdef AB200_BVP_VARIANTS            // opt-in build (make EXTRA=-DAB200_BVP_VARIANTS): the two experimental forward eliminations of round 1
#include "vl_bvp3.cuh"
#include "vl_bvp4.cuh"
#endif
#include <algorithm>

using namespace vlk;

// VAR: 0 = blocked look-ahead elimination, window in shared memory; 1 = the same with the window in a per-CTA HBM scratch ("large n");
// 3 = warp-specialised (vl_bvp3.cuh), 4 = register-resident window (vl_bvp4.cuh) -- the last two only with AB200_BVP_VARIANTS
template <int S, int VAR>
__global__ void __launch_bounds__(384) k_vl_bvp(VLPix X, VLSol Z, VLQuad QD, VLSurf SF, VLConv CV, int lambertian, double accuracy,
                                                int m, const int *__restrict__ active, int nactive, int *__restrict__ next_list,
                                                int *__restrict__ next_count, double *__restrict__ work, size_t work_per_cta,
                                                int smem_doubles, double *__restrict__ wglob) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double sd[64];
  __shared__ int si[64];
  const int n = S * QD.N;
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

// small problems (3 nstokes nstreams <= 64 window rows): one WARP per pixel with its own window in shared memory, several pixels per CTA
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

namespace vlk {

using band_fn = void (*)(VLPix, VLSol, VLQuad, VLSurf, VLConv, int, double, int, const int *, int, int *, int *, double *, size_t, int, double *);
static band_fn band_kernel(int S, int var) {
#ifdef AB200_BVP_VARIANTS
  if (var == 4) return S == 1 ? k_vl_bvp<1, 4> : k_vl_bvp<3, 4>;
  if (var == 3) return S == 1 ? k_vl_bvp<1, 3> : k_vl_bvp<3, 3>;
#endif
  if (var == 1) return S == 1 ? k_vl_bvp<1, 1> : k_vl_bvp<3, 1>;
  return S == 1 ? k_vl_bvp<1, 0> : k_vl_bvp<3, 0>;
}

int vl_band_plan(BandPlan &P, int S, int N, int L, int nsm) {
  const int n = S * N;
  P.S = S; P.n = n; P.L = L; P.var = 0;
#ifdef AB200_BVP_VARIANTS
  if (const char *e = getenv("AB200_VL_BVP_V3"))
    if (atoi(e) != 0 && b3_supported(n) && (b3_smem_doubles(n) * 8 + b3_smem_ints(n) * 4 <= 227 * 1024)) P.var = 3;
  if (const char *e = getenv("AB200_VL_BVP_V4")) { const int v4 = atoi(e); if (!P.var && v4 && b4_supported(n) && (v4 >= 2 || n > 24)) P.var = 4; }
#endif
  // when the 3n x (4n+1) window does not fit the 227 KB of shared memory it moves to a per-CTA HBM scratch ("large n" layout)
  const bool big_n = !P.var && (bvp_smem_doubles(n) * 8 + (size_t)(13 * n + 32) * 4 > 227 * 1024);
  if (big_n) P.var = 1;
  size_t smem_d = big_n ? bvp_smem_doubles_g(n) : bvp_smem_doubles(n), ints = (size_t)(13 * n + 32);
#ifdef AB200_BVP_VARIANTS
  if (P.var == 4) smem_d = b4_smem_doubles(n);
  if (P.var == 3) { smem_d = b3_smem_doubles(n); ints = b3_smem_ints(n); }
#endif
  P.smem_d = (int)smem_d; P.smem = smem_d * 8 + ints * 4;
  P.warp = !P.var && 3 * n <= 64;
  if (const char *e = getenv("AB200_VL_BVP_WARP")) P.warp = P.warp && atoi(e) != 0;
  P.wpc = 4;
  if (P.warp) {
    const size_t per_warp = (bvp_smem_doubles(n) + (size_t)(13 * n + 32 + 1) / 2 + 2 + 1) & ~(size_t)1;     // doubles (ints behind)
    P.smem_d = (int)per_warp; P.smem = per_warp * 8 * P.wpc;
    auto kw = (S == 1) ? (const void *)k_vl_bvp_warp<1> : (const void *)k_vl_bvp_warp<3>;
    AB_CUDA_OK(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
    AB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&P.per_sm, kw, 32 * P.wpc, P.smem));
    P.per_sm = std::max(1, std::min(P.per_sm, 16));
    P.grid_max = nsm * P.per_sm * P.wpc;
  } else {
    if (P.smem > 227 * 1024) {
      ab200_set_error("vlidort: nstokes*nstreams too large (> 64: vector calls take up to 21 streams, scalar calls the reference's 25): the "
                      "boundary-value working set of a pixel does not fit the shared memory of an SM");
      return AB200_ERR_ARG;
    }
    auto kb = (const void *)band_kernel(S, P.var);
    AB_CUDA_OK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
    AB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&P.per_sm, kb, 384, P.smem));
    P.per_sm = std::max(1, std::min(P.per_sm, 8));
    P.grid_max = nsm * P.per_sm;
  }
  P.work_per_solver = bvp_work_doubles(n, L);
  P.window = big_n ? bvp_window_doubles(n) : 0;
  return 0;
}

void vl_band_launch(const BandPlan &P, const VLPix &X, const VLSol &Z, const VLQuad &QD, const VLSurf &SF, const VLConv &CV, int lambertian,
                    double accuracy, int m, const int *active, int nactive, int *next_list, int *next_count, double *work, double *wglob,
                    cudaStream_t st) {
  if (P.warp) {
    const int grid = std::min((nactive + P.wpc - 1) / P.wpc, P.grid_max / P.wpc);
    if (P.S == 1) k_vl_bvp_warp<1><<<grid, 32 * P.wpc, P.smem, st>>>(X, Z, QD, SF, CV, lambertian, accuracy, m, active, nactive, next_list, next_count, work, P.work_per_solver, P.smem_d);
    else k_vl_bvp_warp<3><<<grid, 32 * P.wpc, P.smem, st>>>(X, Z, QD, SF, CV, lambertian, accuracy, m, active, nactive, next_list, next_count, work, P.work_per_solver, P.smem_d);
  } else {
    const int grid = std::min(nactive, P.grid_max);
    band_kernel(P.S, P.var)<<<grid, 384, P.smem, st>>>(X, Z, QD, SF, CV, lambertian, accuracy, m, active, nactive, next_list, next_count, work, P.work_per_solver, P.smem_d, wglob);
  }
}

}  // namespace vlk
