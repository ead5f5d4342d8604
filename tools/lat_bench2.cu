// does a latency-bound FP64 warp slow down when other warps (other SMSPs / same SMSP) saturate the FP64 pipe?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int iters, int hogmask, int hogmode) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double b = 0.999999, c = 1e-9;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 0) {
    double a = 1.0 + lane;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    long long t1 = clock64();
    double p = a;
    long long t2 = clock64();
    for (int i = 0; i < iters; i++) { p = __shfl_sync(0xffffffffu, p, (lane + 1) & 31); p = fma(p, b, c); }
    long long t3 = clock64();
    double r[8]; for (int j = 0; j < 8; j++) r[j] = a + j;
    long long t4 = clock64();
    for (int i = 0; i < iters; i++) { for (int j = 0; j < 8; j++) r[j] = fma(r[j], b, c); }
    long long t5 = clock64();
    for (int j = 0; j < 8; j++) a += r[j];
    if (lane == 0) { cyc[0] = (t1 - t0) / (4 * iters); cyc[1] = (t3 - t2) / iters; cyc[2] = (t5 - t4) * 100 / (8 * iters); stop = 1; }
    out[lane] = a + p;
  } else if ((hogmask >> warp) & 1) {
    double e[16]; for (int j = 0; j < 16; j++) e[j] = j + threadIdx.x;
    while (!stop) {
      if (hogmode == 0) for (int j = 0; j < 16; j++) e[j] = fma(e[j], b, c);
      else for (int j = 0; j < 8; j++) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e[2*j]), "+d"(e[2*j+1]) : "d"(b), "d"(c));
    }
    double a = 0; for (int j = 0; j < 16; j++) a += e[j];
    out[threadIdx.x] = a;
  }
}
int main() {
  double *out; long long *cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 16);
  long long h[4];
  struct { const char *name; int mask, mode; } cases[] = {
    {"alone", 0, 0},
    {"DFMA hogs on other SMSPs (warps 1,2,3)", 0xE, 0}, {"DMMA hogs on other SMSPs (warps 1,2,3)", 0xE, 1},
    {"DFMA hog same SMSP (warp 4)", 0x10, 0}, {"DMMA hog same SMSP (warp 4)", 0x10, 1},
    {"DMMA hogs 9 warps other SMSPs (1-3,5-7,9-11)", 0xEEE, 1}, {"DFMA hogs 9 warps other SMSPs", 0xEEE, 0},
    {"DMMA hogs all 11 other warps", 0xFFE, 1}, {"DFMA hogs all 11 other warps", 0xFFE, 0}};
  for (auto &cs : cases) {
    for (int rep = 0; rep < 2; rep++) { k<<<1, 384>>>(out, cyc, 4000, cs.mask, cs.mode); cudaDeviceSynchronize(); }
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-50s DFMA-dep %lld  shfl+fma %lld  indepDFMA x100 %lld\n", cs.name, h[0], h[1], h[2]);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
