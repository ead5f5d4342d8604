// Stage B, kernels K2b (layer interaction operators, vl_layerops.cuh) and K3 (interaction-principle boundary-value solver +
// post-processing + Fourier convergence, vl_add.cuh) -- kernel wrappers and launch plan
#include "vl_launch.h"
#include <algorithm>

using namespace vlk;

// K2b: one CTA per (pixel, layer) problem
// TPB: 128 (four CTAs per SM at nstokes*nstreams = 36), or 512 = upper bound of the larger CTAs that bigger problems are launched with
template <int S, int TPB>
__global__ void __launch_bounds__(TPB, TPB == 128 ? 4 : 1) k_vl_layer_ops(VLPix X, VLSol Z, VLOps O, const int *__restrict__ active, int nactive) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double sd[64];
  __shared__ int si[64];
  const long prob = blockIdx.x;
  const int p = active[prob / X.L], k = (int)(prob % X.L);
  BlockTeam T(sd, si);
  layer_ops_problem<S>(T, X, Z, O, p, k, smem_d, reinterpret_cast<int *>(smem_d + ops_smem_doubles(Z.n, true)));
}
// small problems: one warp per (pixel, layer) problem, several problems per CTA
template <int S>
__global__ void __launch_bounds__(256) k_vl_layer_ops_warp(VLPix X, VLSol Z, VLOps O, const int *__restrict__ active, int nactive, int per_warp) {
  extern __shared__ __align__(16) double smem_d[];
  const int warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const long prob = (long)blockIdx.x * wpc + warp;
  if (prob >= (long)nactive * X.L) return;
  const int p = active[prob / X.L], k = (int)(prob % X.L);
  double *scr = smem_d + (size_t)warp * per_warp;
  WarpTeam T;
  layer_ops_problem<S>(T, X, Z, O, p, k, scr, reinterpret_cast<int *>(scr + ops_smem_doubles(Z.n, false)));
}

// K3: persistent CTAs, one pixel at a time per CTA, several CTAs per SM
template <int S, int TPB>
__global__ void __launch_bounds__(TPB, TPB == 128 ? 3 : 1) k_vl_add(VLPix X, VLSol Z, VLOps O, VLQuad QD, VLSurf SF, VLConv CV, int lambertian, double accuracy,
                                                   int m, const int *__restrict__ active, int nactive, int *__restrict__ next_list,
                                                   int *__restrict__ next_count, double *__restrict__ work, size_t work_per_cta) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double sd[64];
  __shared__ int si[64];
  __shared__ unsigned long long chan[4];
  for (int i = threadIdx.x; i < (int)add_smem_doubles(Z.n, true); i += blockDim.x) smem_d[i] = 0.0;    // no stale NaN patterns in the padding
  BlockTeam T(sd, si, nullptr, chan);
  int *piv = reinterpret_cast<int *>(smem_d + add_smem_doubles(Z.n, true));
  double *Wg = work + (size_t)blockIdx.x * work_per_cta;
  for (int a = blockIdx.x; a < nactive; a += gridDim.x) {
    const int p = active[a];
    if (Z.fail[p]) add_fail_pixel(T, X, CV, p);
    else add_pixel<S>(T, X, Z, O, QD, SF, CV, lambertian, accuracy, m, p, smem_d, piv, Wg);
    __syncthreads();
    if (threadIdx.x == 0 && CV.col_active[p]) { const int idx = atomicAdd(next_count, 1); next_list[idx] = p; }
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
  int *piv = reinterpret_cast<int *>(scr + add_smem_doubles(Z.n, false));
  WarpTeam T;
  const int gw = blockIdx.x * wpc + warp;
  double *Wg = work + (size_t)gw * work_per_warp;
  for (int a = gw; a < nactive; a += gridDim.x * wpc) {
    const int p = active[a];
    if (Z.fail[p]) add_fail_pixel(T, X, CV, p);
    else add_pixel<S>(T, X, Z, O, QD, SF, CV, lambertian, accuracy, m, p, scr, piv, Wg);
    __syncwarp();
    if (lane == 0 && CV.col_active[p]) { const int idx = atomicAdd(next_count, 1); next_list[idx] = p; }
    __syncwarp();
  }
}

// ---- self-test of the CTA linear algebra (tests/test_linalg_gpu.py): E = A^-1 by team_inverse and P = A E by team_mm, one CTA
__global__ void k_vl_selftest_linalg(int n, const double *__restrict__ A, double *__restrict__ E, double *__restrict__ P, int *__restrict__ okflag) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double sd[64];
  __shared__ int si[64];
  BlockTeam T(sd, si);
  const int ld = mat_ld(n);
  double *As = smem_d, *Es = As + (size_t)n * ld, *wk = Es + (size_t)n * ld;
  int *piv = reinterpret_cast<int *>(wk + inv_work_doubles(n, true));
  for (int idx = T.rank; idx < n * ld; idx += T.size) { const int r = idx / ld, c = idx % ld; As[idx] = c < n ? A[(size_t)r * n + c] : 0.0; Es[idx] = As[idx]; }
  T.sync();
  const bool ok = team_inverse(T, n, Es, Es, ld, wk, piv);
  if (T.rank == 0) *okflag = ok ? 1 : 0;
  if (!ok) return;
  for (int idx = T.rank; idx < n * n; idx += T.size) E[idx] = Es[(size_t)(idx / n) * ld + idx % n];
  team_mm(T, n, n, n, As, ld, Es, ld, [&](int r, int c, double acc) { P[(size_t)r * n + c] = acc; });
}

extern "C" int ab200_selftest_linalg(int n, int threads, const double *A, double *E, double *P) {
  if (n < 1 || n > 64 || threads < 32 || threads > 512 || threads % 32) { ab200_set_error("selftest_linalg: bad arguments"); return AB200_ERR_ARG; }
  const size_t nn = (size_t)n * n, smem = (2 * (size_t)n * mat_ld(n) + inv_work_doubles(n, true)) * 8 + (size_t)n * 8 + 16;
  double *d = nullptr; int *dok = nullptr;
  AB_CUDA_OK(cudaMalloc((void **)&d, 3 * nn * 8 + 16));
  dok = reinterpret_cast<int *>(d + 3 * nn);
  AB_CUDA_OK(cudaMemcpy(d, A, nn * 8, cudaMemcpyHostToDevice));
  AB_CUDA_OK(cudaFuncSetAttribute((const void *)k_vl_selftest_linalg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_vl_selftest_linalg<<<1, threads, smem>>>(n, d, d + nn, d + 2 * nn, dok);
  AB_CUDA_OK(cudaGetLastError());
  int ok = 0;
  AB_CUDA_OK(cudaMemcpy(&ok, dok, 4, cudaMemcpyDeviceToHost));
  AB_CUDA_OK(cudaMemcpy(E, d + nn, nn * 8, cudaMemcpyDeviceToHost));
  AB_CUDA_OK(cudaMemcpy(P, d + 2 * nn, nn * 8, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return ok ? 0 : 1;
}

namespace vlk {

int vl_add_plan(AddPlan &P, int S, int N, int L, int nsm) {
  const int n = S * N;
  P.S = S; P.n = n; P.L = L;
  P.warp = n <= 16;
  P.wpc = 8;
  // CTA size: the register-resident inversion keeps ceil(n / (threads / n)) <= 13 elements of a matrix row per thread
  P.threads = 128;
  while (P.threads < 1024 && (n + P.threads / n - 1) / (P.threads / n) > 13) P.threads += 128;
  P.ops_pw = (ops_smem_doubles(n, false) + (size_t)n + 2) & ~(size_t)1;
  P.add_pw = (add_smem_doubles(n, false) + (size_t)n + 2) & ~(size_t)1;
  P.ops_smem = P.warp ? P.ops_pw * 8 * P.wpc : ops_smem_doubles(n, true) * 8 + (size_t)n * 16 + 16;      // + 4n ints (pivots of the inverse pair)
  P.add_smem = P.warp ? P.add_pw * 8 * P.wpc : add_smem_doubles(n, true) * 8 + (size_t)n * 8 + 16;
  if (P.ops_smem > 227 * 1024 || P.add_smem > 227 * 1024 || P.threads > 512) {
    ab200_set_error("vlidort: nstokes*nstreams too large for the shared-memory interaction-principle solver");
    return AB200_ERR_ARG;
  }
  const bool big = P.threads > 128;
  auto ko = P.warp ? ((S == 1) ? (const void *)k_vl_layer_ops_warp<1> : (const void *)k_vl_layer_ops_warp<3>)
                   : (big ? ((S == 1) ? (const void *)k_vl_layer_ops<1, 512> : (const void *)k_vl_layer_ops<3, 512>)
                          : ((S == 1) ? (const void *)k_vl_layer_ops<1, 128> : (const void *)k_vl_layer_ops<3, 128>));
  auto ka = P.warp ? ((S == 1) ? (const void *)k_vl_add_warp<1> : (const void *)k_vl_add_warp<3>)
                   : (big ? ((S == 1) ? (const void *)k_vl_add<1, 512> : (const void *)k_vl_add<3, 512>)
                          : ((S == 1) ? (const void *)k_vl_add<1, 128> : (const void *)k_vl_add<3, 128>));
  AB_CUDA_OK(cudaFuncSetAttribute(ko, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.ops_smem));
  AB_CUDA_OK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.add_smem));
  AB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&P.per_sm, ka, P.warp ? 32 * P.wpc : P.threads, P.add_smem));
  P.per_sm = std::max(1, std::min(P.per_sm, 8));
  P.grid_max = P.warp ? nsm * P.per_sm * P.wpc : nsm * P.per_sm;        // concurrently running solvers (CTAs, or warps)
  P.work_per_solver = add_work_doubles(n, L);
  return 0;
}

void vl_ops_launch(const AddPlan &P, const VLPix &X, const VLSol &Z, const VLOps &O, const int *active, int nactive, cudaStream_t st) {
  const long nprob = (long)nactive * X.L;
  if (P.warp) {
    const int grid = (int)((nprob + P.wpc - 1) / P.wpc);
    if (P.S == 1) k_vl_layer_ops_warp<1><<<grid, 32 * P.wpc, P.ops_smem, st>>>(X, Z, O, active, nactive, (int)P.ops_pw);
    else k_vl_layer_ops_warp<3><<<grid, 32 * P.wpc, P.ops_smem, st>>>(X, Z, O, active, nactive, (int)P.ops_pw);
  } else {
    if (P.threads > 128) {
      if (P.S == 1) k_vl_layer_ops<1, 512><<<(unsigned)nprob, P.threads, P.ops_smem, st>>>(X, Z, O, active, nactive);
      else k_vl_layer_ops<3, 512><<<(unsigned)nprob, P.threads, P.ops_smem, st>>>(X, Z, O, active, nactive);
    } else {
      if (P.S == 1) k_vl_layer_ops<1, 128><<<(unsigned)nprob, P.threads, P.ops_smem, st>>>(X, Z, O, active, nactive);
      else k_vl_layer_ops<3, 128><<<(unsigned)nprob, P.threads, P.ops_smem, st>>>(X, Z, O, active, nactive);
    }
  }
}

void vl_add_launch(const AddPlan &P, const VLPix &X, const VLSol &Z, const VLOps &O, const VLQuad &QD, const VLSurf &SF, const VLConv &CV,
                   int lambertian, double accuracy, int m, const int *active, int nactive, int *next_list, int *next_count, double *work,
                   cudaStream_t st) {
  if (P.warp) {
    const int grid = std::min((nactive + P.wpc - 1) / P.wpc, P.grid_max / P.wpc);
    if (P.S == 1) k_vl_add_warp<1><<<grid, 32 * P.wpc, P.add_smem, st>>>(X, Z, O, QD, SF, CV, lambertian, accuracy, m, active, nactive, next_list, next_count, work, P.work_per_solver, (int)P.add_pw);
    else k_vl_add_warp<3><<<grid, 32 * P.wpc, P.add_smem, st>>>(X, Z, O, QD, SF, CV, lambertian, accuracy, m, active, nactive, next_list, next_count, work, P.work_per_solver, (int)P.add_pw);
  } else {
    const int grid = std::min(nactive, P.grid_max);
#define AB_ADD(K) K<<<grid, P.threads, P.add_smem, st>>>(X, Z, O, QD, SF, CV, lambertian, accuracy, m, active, nactive, next_list, next_count, work, P.work_per_solver)
    if (P.threads > 128) { if (P.S == 1) AB_ADD((k_vl_add<1, 512>)); else AB_ADD((k_vl_add<3, 512>)); }
    else { if (P.S == 1) AB_ADD((k_vl_add<1, 128>)); else AB_ADD((k_vl_add<3, 128>)); }
#undef AB_ADD
  }
}

}  // namespace vlk
