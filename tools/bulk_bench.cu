// replica of the v4 rank-1 update: 12 warps, each step 18 broadcast LDS.128 + 37 independent DFMA with one multiplier.
// variants: 0 = straightforward, 1 = loads batched 8 x LDS.128 ahead of their FMAs, 2 = no shared loads (pure DFMA), 3 = LDS only
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void __launch_bounds__(384) k(double *out, long long *cyc, int iters) {
  __shared__ __align__(16) double pbs[2 * 3 * 50];
  for (int i = threadIdx.x; i < 300; i += blockDim.x) pbs[i] = 1e-3 * i;
  __syncthreads();
  const int warp = threadIdx.x >> 5, cg = warp % 3;
  double w[49];
#pragma unroll
  for (int i = 0; i < 49; i++) w[i] = i + threadIdx.x;
  double l = 1e-6 * threadIdx.x, sink = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const double *pb = pbs + ((it & 1) * 3 + cg) * 50;
    if (V == 0) {
#pragma unroll
      for (int lc = 12; lc < 49; lc += 2) {
        const double2 p2 = *reinterpret_cast<const double2 *>(pb + lc);
        w[lc] -= l * p2.x;
        if (lc + 1 < 49) w[lc + 1] -= l * p2.y;
      }
    } else if (V == 1) {
#pragma unroll
      for (int c0 = 12; c0 < 49; c0 += 16) {
        double2 p[8];
#pragma unroll
        for (int q = 0; q < 8; q++) if (c0 + 2 * q < 49) asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(p[q].x), "=d"(p[q].y) : "r"((unsigned)__cvta_generic_to_shared(pb + c0 + 2 * q)));
#pragma unroll
        for (int q = 0; q < 8; q++) if (c0 + 2 * q < 49) { w[c0 + 2 * q] -= l * p[q].x; if (c0 + 2 * q + 1 < 49) w[c0 + 2 * q + 1] -= l * p[q].y; }
      }
    } else if (V == 2) {
#pragma unroll
      for (int lc = 12; lc < 49; lc++) w[lc] -= l * 1.000001;
    } else {
#pragma unroll
      for (int lc = 12; lc < 49; lc += 2) { double a, b; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"((unsigned)__cvta_generic_to_shared(pb + lc))); sink += a; }
    }
    l += 1e-9;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / iters;
  double a = sink;
#pragma unroll
  for (int i = 0; i < 49; i++) a += w[i];
  out[threadIdx.x] = a;
}
int main() {
  double *out; long long *cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
  long long h;
  for (int v = 0; v < 4; v++) {
    for (int rep = 0; rep < 2; rep++) {
      if (v == 0) k<0><<<1, 384>>>(out, cyc, 5000); if (v == 1) k<1><<<1, 384>>>(out, cyc, 5000);
      if (v == 2) k<2><<<1, 384>>>(out, cyc, 5000); if (v == 3) k<3><<<1, 384>>>(out, cyc, 5000);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("variant %d: %lld cycles per step\n", v, h);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
