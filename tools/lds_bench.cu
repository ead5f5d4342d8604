// throughput of shared-memory loads where every lane reads the same address (broadcast): LDS.32 / .64 / .128, 12 warps
#include <cstdio>
#include <cuda_runtime.h>
template <int W> __global__ void k(double *out, long long *cyc, int iters, int bcast) {
  __shared__ __align__(16) double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int base = warp * 64 + (bcast ? 0 : lane * (W / 4 > 2 ? 2 : 1) * 2);   // doubles; distinct case: contiguous per lane
  unsigned acc[16];
#pragma unroll
  for (int u = 0; u < 16; u++) acc[u] = u;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const unsigned addr = sbase + (unsigned)(((base + u * 2 + (i & 1) * 32) & 2047) * 8);
      unsigned a0, a1, a2, a3;
      if (W == 4) { asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a0) : "r"(addr)); acc[u] ^= a0; }
      if (W == 8) { asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(addr)); acc[u] ^= a0; }
      if (W == 16) { asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr)); acc[u] ^= a0; }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  unsigned tt = 0;
  for (int u = 0; u < 16; u++) tt ^= acc[u];
  out[threadIdx.x] = tt;
}
int main() {
  double *out; long long *cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
  long long h;
  const int iters = 2000, nw = 12;
  for (int bc = 1; bc >= 0; bc--) {
    for (int W : {4, 8, 16}) {
      for (int rep = 0; rep < 2; rep++) {
        if (W == 4) k<4><<<1, 32 * nw>>>(out, cyc, iters, bc); if (W == 8) k<8><<<1, 32 * nw>>>(out, cyc, iters, bc); if (W == 16) k<16><<<1, 32 * nw>>>(out, cyc, iters, bc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%s LDS.%d: %.2f cycles per warp-instruction (SM-wide, 12 warps)\n", bc ? "broadcast" : "distinct ", W * 8, (double)h / (16.0 * iters * nw));
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
