// Microbenchmark: what does the FP64 tensor pipe deliver with N warps per SM issuing independent
// DMMA.8x8x4 streams from registers (no memory traffic)?  Used to interpret the contraction kernel's
// sm__pipe_tensor_cycles_active and to pick warps-per-sub-partition.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC, int LDS_PER_16, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(double* out, int iters, long long* cyc) {
  __shared__ double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
  __syncthreads();
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (LDS_PER_16 > 0 && (i % (16 / LDS_PER_16)) == 0) {
        a = sm[(threadIdx.x + it * 32 + i * 64) & 2047];
      }
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NACC, int LDS, int WARPS>
void run(int ctas_per_sm, int iters) {
  const int warps = WARPS;
  int sms = 148;
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * warps * 32);
  cudaMalloc(&cyc, sizeof(long long) * sms * ctas_per_sm);
  k<NACC, LDS, WARPS><<<sms * ctas_per_sm, warps * 32>>>(out, 10, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<NACC, LDS, WARPS><<<sms * ctas_per_sm, warps * 32>>>(out, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148 * 8]; cudaMemcpy(h, cyc, sizeof(long long) * sms * ctas_per_sm, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms * ctas_per_sm; ++i) avg += h[i]; avg /= sms * ctas_per_sm;
  double dmma_per_sm = (double)iters * NACC * warps * ctas_per_sm;
  double fma_per_clk_sm = dmma_per_sm * 256.0 / avg;
  double tflops = 2.0 * 256.0 * dmma_per_sm * sms / (ms * 1e-3) / 1e12;
  printf("nacc=%2d lds/16=%d warps/CTA=%2d CTAs/SM=%d : %.2f FMA/clk/SM (of 64), %.2f TFLOP/s by events, %.0f cyc\n", NACC, LDS, warps,
         ctas_per_sm, fma_per_clk_sm, tflops, avg);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaError_t e = cudaGetLastError();
  run<16, 0, 4>(1, 4000); run<16, 0, 8>(1, 4000); run<16, 0, 16>(1, 2000);
  run<32, 0, 4>(1, 2000); run<32, 0, 8>(1, 2000); run<32, 0, 16>(1, 1000);
  run<64, 0, 4>(1, 1000); run<64, 0, 8>(1, 1000);
  run<32, 4, 8>(1, 2000); run<32, 8, 8>(1, 2000); run<64, 4, 8>(1, 1000); run<64, 8, 8>(1, 1000);
  run<32, 4, 16>(1, 1000); run<32, 8, 16>(1, 1000);
  run<32, 4, 4>(2, 2000); run<64, 4, 4>(2, 1000);
  e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
