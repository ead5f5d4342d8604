// Stage B, kernel K2: per-layer discrete-ordinate eigenproblems (vl_eigen.cuh) -- kernel wrappers and launch plan
#include "vl_launch.h"
#include <algorithm>

using namespace vlk;

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

// two warps per (pixel, layer) problem (DuoTeam): same shared memory per problem, twice the warps per SM.  The 24-row instantiation
// (36 < n <= 48) runs two problems per CTA (shared memory), so it may use up to 255 registers: no spills of the 48 column values per lane
template <int S, int MAXHALF = 18>
__global__ void __launch_bounds__(MAXHALF > 18 ? 128 : 256, 2) k_vl_eigen_duo(VLPix X, VLSol Z, VLQuad QD, const double *__restrict__ pq, int m,
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

namespace vlk {

static const void *eigen_fn(int S, int kind) {
  switch (kind) {
    case 1: return S == 1 ? (const void *)k_vl_eigen_sub<1, 8> : (const void *)k_vl_eigen_sub<3, 8>;
    case 2: return S == 1 ? (const void *)k_vl_eigen_sub<1, 16> : (const void *)k_vl_eigen_sub<3, 16>;
    case 3: return S == 1 ? (const void *)k_vl_eigen_duo<1, 18> : (const void *)k_vl_eigen_duo<3, 18>;
    case 4: return S == 1 ? (const void *)k_vl_eigen_duo<1, 24> : (const void *)k_vl_eigen_duo<3, 24>;
    default: return S == 1 ? (const void *)k_vl_eigen<1> : (const void *)k_vl_eigen<3>;
  }
}

int vl_eigen_plan(EigenPlan &P, int S, int N, int NM) {
  const int n = S * N;
  P.S = S; P.n = n;
  P.scr_d = eigen_scratch_doubles(n, NM);
  // warps (= problems) resident per SM within the 227 KB opt-in limit; two CTAs per SM when possible: problems need a
  // different number of Jacobi sweeps and a CTA only retires when its slowest warp does (measured: 2 x 4 warps beats 1 x 8)
  int wpc = (int)std::min<size_t>(8, (226 * 1024) / ((size_t)P.scr_d * 8));
  if (wpc >= 4) wpc /= 2;
  if (const char *e = getenv("AB200_VL_EIGEN_WARPS")) wpc = std::max(1, std::min(wpc, atoi(e)));      // tuning knob
  if (wpc < 1) { ab200_set_error("vlidort: nstokes*nstreams too large for the shared-memory eigen scratch"); return AB200_ERR_ARG; }
  // two warps per problem when the register-resident half-pair Jacobi applies (24 <= n <= 48); AB200_VL_EIGEN_DUO=0 switches it off
  bool duo = n <= 48 && n >= 24;
  if (const char *e = getenv("AB200_VL_EIGEN_DUO")) duo = duo && atoi(e) != 0;
  // sub-warp teams for small problems: 8 lanes per problem up to n = 8, 16 up to n = 16 (AB200_VL_EIGEN_SUB=0 switches them off)
  int subw = (n <= 8) ? 8 : (n <= 16 ? 16 : 32);
  if (const char *e = getenv("AB200_VL_EIGEN_SUB")) { if (atoi(e) == 0) subw = 32; }
  if (subw < 32) {
    const int sub_warps = 4, subs = 32 / subw;
    P.kind = subw == 8 ? 1 : 2; P.threads = 32 * sub_warps; P.probs_per_cta = sub_warps * subs;
    P.smem = (size_t)P.probs_per_cta * P.scr_d * 8;
  } else if (duo) {
    P.kind = n > 36 ? 4 : 3;
    if (P.kind == 4) wpc = std::min(wpc, 2);             // the 24-row kernel is compiled for at most 128 threads per CTA (255 registers)
    P.threads = 64 * wpc; P.probs_per_cta = wpc; P.smem = (size_t)wpc * (P.scr_d + 4) * 8;
  } else {
    P.kind = 0; P.threads = 32 * wpc; P.probs_per_cta = wpc; P.smem = (size_t)wpc * P.scr_d * 8;
  }
  P.wpc = wpc;
  AB_CUDA_OK(cudaFuncSetAttribute(eigen_fn(S, P.kind), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
  return 0;
}

void vl_eigen_launch(const EigenPlan &P, const VLPix &X, const VLSol &Z, const VLQuad &QD, const double *pq, int m, const int *active,
                     int nactive, cudaStream_t st) {
  const long nprob = (long)nactive * X.L;
  const int grid = (int)((nprob + P.probs_per_cta - 1) / P.probs_per_cta);
  const int S = P.S;
#define AB_EIG(K) K<<<grid, P.threads, P.smem, st>>>(X, Z, QD, pq, m, active, nactive, P.scr_d)
  switch (P.kind) {
    case 1: if (S == 1) AB_EIG((k_vl_eigen_sub<1, 8>)); else AB_EIG((k_vl_eigen_sub<3, 8>)); break;
    case 2: if (S == 1) AB_EIG((k_vl_eigen_sub<1, 16>)); else AB_EIG((k_vl_eigen_sub<3, 16>)); break;
    case 3: if (S == 1) AB_EIG((k_vl_eigen_duo<1, 18>)); else AB_EIG((k_vl_eigen_duo<3, 18>)); break;
    case 4: if (S == 1) AB_EIG((k_vl_eigen_duo<1, 24>)); else AB_EIG((k_vl_eigen_duo<3, 24>)); break;
    default: if (S == 1) AB_EIG((k_vl_eigen<1>)); else AB_EIG((k_vl_eigen<3>)); break;
  }
#undef AB_EIG
}

}  // namespace vlk
