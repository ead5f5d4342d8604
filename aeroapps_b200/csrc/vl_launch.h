// Stage B -- host-side launch plans of the three heavy kernel families.  Each family lives in its own translation unit
// (vl_k2_eigen.cu, vl_k3_add.cu, vl_k3_band.cu) so that the library builds in parallel; vlidort.cu holds the driver and the C ABI.
#pragma once
#include "../../include/aeroapps_b200.h"
#include "common.h"
#include "vl_add.cuh"

namespace vlk {

// K2: per-layer eigenproblems.  kind: 0 one warp per problem, 1 / 2 sub-warp teams of 8 / 16 lanes, 3 / 4 two warps per problem with
// 18- / 24-row register halves
struct EigenPlan {
  int S = 0, n = 0, kind = 0, wpc = 0, scr_d = 0, threads = 0, probs_per_cta = 0;
  size_t smem = 0;
};
int vl_eigen_plan(EigenPlan &P, int S, int N, int NM);
void vl_eigen_launch(const EigenPlan &P, const VLPix &X, const VLSol &Z, const VLQuad &QD, const double *pq, int m, const int *active,
                     int nactive, cudaStream_t st);

// K2b + K3: interaction-principle solver (vl_layerops.cuh, vl_add.cuh), nstokes*nstreams <= 64
struct AddPlan {
  int S = 0, n = 0, L = 0, warp = 0, wpc = 8, threads = 128, per_sm = 1, grid_max = 0;
  size_t ops_smem = 0, add_smem = 0, ops_pw = 0, add_pw = 0, work_per_solver = 0;
};
int vl_add_plan(AddPlan &P, int S, int N, int L, int nsm);
void vl_ops_launch(const AddPlan &P, const VLPix &X, const VLSol &Z, const VLOps &O, const int *active, int nactive, cudaStream_t st);
void vl_add_launch(const AddPlan &P, const VLPix &X, const VLSol &Z, const VLOps &O, const VLQuad &QD, const VLSurf &SF, const VLConv &CV,
                   int lambertian, double accuracy, int m, const int *active, int nactive, int *next_list, int *next_count, double *work,
                   cudaStream_t st);

// K3, band form: windowed elimination of the reference's band matrix (vl_bvp.cuh); AB200_VL_BVP_LEGACY=1 or nstokes*nstreams > 64
struct BandPlan {
  int S = 0, n = 0, L = 0, var = 0, warp = 0, wpc = 4, per_sm = 1, grid_max = 0, smem_d = 0;
  size_t smem = 0, work_per_solver = 0, window = 0;
};
int vl_band_plan(BandPlan &P, int S, int N, int L, int nsm);
void vl_band_launch(const BandPlan &P, const VLPix &X, const VLSol &Z, const VLQuad &QD, const VLSurf &SF, const VLConv &CV, int lambertian,
                    double accuracy, int m, const int *active, int nactive, int *next_list, int *next_count, double *work, double *wglob,
                    cudaStream_t st);

}  // namespace vlk
