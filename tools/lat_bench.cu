// latency microbenchmarks (single warp, dependent chains) for the K3 panel chain: DFMA, DADD, 1/x, SHFL, REDUX, BALLOT, LDS
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int iters) {
  __shared__ double sm[1024];
  const int lane = threadIdx.x;
  for (int i = lane; i < 1024; i += 32) sm[i] = (double)((i * 7 + 1) % 1024);
  __syncwarp();
  double a = 1.0 + lane * 1e-9, b = 0.999999, c = 1e-9;
  long long t0, t1;
  // 0 DFMA chain
  t0 = clock64();
  for (int i = 0; i < iters; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  t1 = clock64(); if (lane == 0) cyc[0] = (t1 - t0) / (4 * iters);
  // 1 independent DFMA x8 (throughput of one warp)
  double r[8]; for (int j = 0; j < 8; j++) r[j] = a + j;
  t0 = clock64();
  for (int i = 0; i < iters; i++) { for (int j = 0; j < 8; j++) r[j] = fma(r[j], b, c); }
  t1 = clock64(); if (lane == 0) cyc[1] = (t1 - t0) * 100 / (8 * iters);
  for (int j = 0; j < 8; j++) a += r[j];
  // 2 double division chain
  t0 = clock64();
  for (int i = 0; i < iters; i++) { a = 1.0 / (a + 1.5); }
  t1 = clock64(); if (lane == 0) cyc[2] = (t1 - t0) / iters;
  // 3 shfl 64-bit chain
  t0 = clock64();
  for (int i = 0; i < iters; i++) { a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31); }
  t1 = clock64(); if (lane == 0) cyc[3] = (t1 - t0) / iters;
  // 4 redux max chain
  unsigned u = __double2hiint(a) + lane;
  t0 = clock64();
  for (int i = 0; i < iters; i++) { u = __reduce_max_sync(0xffffffffu, u + lane); }
  t1 = clock64(); if (lane == 0) cyc[4] = (t1 - t0) / iters;
  // 5 ballot chain
  t0 = clock64();
  for (int i = 0; i < iters; i++) { u = __ballot_sync(0xffffffffu, (u >> (lane & 7)) & 1u) + i; }
  t1 = clock64(); if (lane == 0) cyc[5] = (t1 - t0) / iters;
  // 6 LDS chain (64-bit)
  int idx = lane;
  t0 = clock64();
  for (int i = 0; i < iters; i++) { idx = (int)sm[idx & 1023]; }
  t1 = clock64(); if (lane == 0) cyc[6] = (t1 - t0) / iters;
  // 7 fabs-compare-select chain (argmax step)
  double best = -1.0; int bi = 0;
  t0 = clock64();
  for (int i = 0; i < iters; i++) { double v = fabs(a + i); if (v > best) { best = v * 0.5; bi = i; } }
  t1 = clock64(); if (lane == 0) cyc[7] = (t1 - t0) / iters;
  // 8 32-bit shfl chain
  t0 = clock64();
  for (int i = 0; i < iters; i++) { u = __shfl_sync(0xffffffffu, u, (lane + 1) & 31) + 1; }
  t1 = clock64(); if (lane == 0) cyc[8] = (t1 - t0) / iters;
  // 9 DMMA chain (dependent accumulators)
  double d0 = a, d1 = b;
  t0 = clock64();
  for (int i = 0; i < iters; i++) { asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(b), "d"(c)); }
  t1 = clock64(); if (lane == 0) cyc[9] = (t1 - t0) / iters;
  // 10 DMMA independent x8 (x100)
  double e[16]; for (int j = 0; j < 16; j++) e[j] = j;
  t0 = clock64();
  for (int i = 0; i < iters; i++) { for (int j = 0; j < 8; j++) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e[2*j]), "+d"(e[2*j+1]) : "d"(b), "d"(c)); }
  t1 = clock64(); if (lane == 0) cyc[10] = (t1 - t0) * 100 / (8 * iters);
  for (int j = 0; j < 16; j++) a += e[j];
  // 11 bar.sync named barrier round trip with 1 warp
  t0 = clock64();
  for (int i = 0; i < iters; i++) { asm volatile("bar.sync 1, 32;" ::: "memory"); }
  t1 = clock64(); if (lane == 0) cyc[11] = (t1 - t0) / iters;
  out[lane] = a + best + bi + u + idx + d0 + d1;
}
// multi-warp throughput: nw warps doing independent DFMA / DMMA; reports cycles*100 per warp-instruction per SMSP
__global__ void k2(double *out, long long *cyc, int iters, int mode) {
  double b = 0.999999, c = 1e-9;
  double e[16]; for (int j = 0; j < 16; j++) e[j] = j + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) for (int i = 0; i < iters; i++) { for (int j = 0; j < 16; j++) e[j] = fma(e[j], b, c); }
  else for (int i = 0; i < iters; i++) { for (int j = 0; j < 8; j++) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e[2*j]), "+d"(e[2*j+1]) : "d"(b), "d"(c)); }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  double a = 0; for (int j = 0; j < 16; j++) a += e[j];
  out[threadIdx.x] = a;
}
int main() {
  double *out; long long *cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 16);
  long long h[16];
  const char *names[] = {"DFMA dependent", "DFMA indep x100 per instr (1 warp)", "1/x double", "SHFL.64 chain", "REDUX.max chain", "BALLOT chain", "LDS.64 chain", "fabs-cmp-select", "SHFL.32 chain", "DMMA dependent", "DMMA indep x100 (1 warp)", "bar.sync 1 warp"};
  for (int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(out, cyc, 2000); cudaDeviceSynchronize(); }
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 12; i++) printf("%-40s %lld\n", names[i], h[i]);
  for (int mode = 0; mode < 2; mode++)
    for (int nw = 4; nw <= 16; nw *= 2) {
      const int iters = 2000;
      k2<<<1, 32 * nw>>>(out, cyc, iters, mode); cudaDeviceSynchronize();
      k2<<<1, 32 * nw>>>(out, cyc, iters, mode); cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
      const double flops = (mode == 0 ? 16.0 * 32 : 8.0 * 256) * nw * iters;   // FMAs
      printf("%s nw=%d: %.1f FMA/clk/SM\n", mode ? "DMMA" : "DFMA", nw, flops / (double)h[0]);
    }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
