// FP64 peak microbenchmark for the Stage-B roofline denominator (none in MEASURED_PEAKS.json):
// (1) DFMA: 8 independent FMA chains per thread; (2) DMMA mma.sync m8n8k4 f64.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_peak.cu -o tools/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, double a, double b, int iters) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_dmma(double *out, double a, double b, int iters) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double av = a + threadIdx.x * 1e-9, bv = b;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(av), "d"(bv));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(av), "d"(bv));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(av), "d"(bv));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(av), "d"(bv));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 256, blocks = sms * 8, iters = 2048;
  double *out; cudaMalloc(&out, sizeof(double) * threads * blocks);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best_fma = 1e30f, best_mma = 1e30f, ms;
  for (int r = 0; r < 6; r++) {
    cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, 0.999999, 1e-7, iters); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best_fma) best_fma = ms;
    cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, 0.999999, 1e-7, iters); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best_mma) best_mma = ms;
  }
  double fl_fma = 2.0 * 64 * iters * (double)threads * blocks;                // 64 FMAs per iteration per thread
  double fl_mma = 2.0 * 8 * 8 * 4 * 32.0 * iters * (double)(threads / 32) * blocks;  // 32 mma per iteration per warp
  printf("{\"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f}\n", sms,
         fl_fma / best_fma / 1e9, fl_mma / best_mma / 1e9, best_fma, best_mma);
  return 0;
}
