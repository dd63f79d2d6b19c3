// Micro-benchmark (measurement only, not part of the library): HBM read bandwidth of a 1024 x 2048 f64 matrix pair (W, RW: 33.5 MB)
// under the access patterns the few-samples kernels could use.  One warp = 8 rows x `kc` columns, as in few_samples_fwd_kernel.
//   0: fragment order   -- lane (g, j) reads 16 B at row m0 + g, column k + 2j: one instruction = 8 rows x 64 B
//   1: row order        -- lane l reads 16 B at row m0 + 2u + l / 16, column k + 2 (l % 16): one instruction = 2 rows x 256 B
//   2: full-row order   -- lane l reads 16 B at column 2l of a 512-byte row segment: one instruction = 1 row x 512 B
//   3: linear           -- grid-stride over the whole matrix, 512 contiguous bytes per instruction
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_pattern bw_pattern.cu ; run: ./bw_pattern
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

template <int PAT>
__global__ void __launch_bounds__(128) k(const double* __restrict__ W, const double* __restrict__ RW, double* out, int rows, int cols, int kc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
  const int m0 = blockIdx.x * 32 + 8 * warp, kbase = blockIdx.y * kc;
  double acc = 0.0;
  if (PAT == 0) {
    const double* wp = W + (long long)(m0 + g) * cols + kbase + 2 * j;
    const double* rp = RW + (long long)(m0 + g) * cols + kbase + 2 * j;
    for (int c0 = 0; c0 < kc / 8; c0 += 4) {
      double2 w[4], r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { w[u] = ld2(wp + 8 * (c0 + u)); r[u] = ld2(rp + 8 * (c0 + u)); }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += w[u].x + w[u].y + r[u].x + r[u].y;
    }
  } else if (PAT == 1) {
    for (int c0 = 0; c0 < kc / 8; c0 += 4) {   // same 8 rows x 32 columns per group, read as 4 instructions of 2 rows x 256 B
      double2 w[4], r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long o = (long long)(m0 + 2 * u + (lane >> 4)) * cols + kbase + 8 * c0 + 2 * (lane & 15);
        w[u] = ld2(W + o); r[u] = ld2(RW + o);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += w[u].x + w[u].y + r[u].x + r[u].y;
    }
  } else if (PAT == 2) {
    for (int c0 = 0; c0 < kc / 64; ++c0) {     // 8 rows x 64 columns per group, one row (512 B) per instruction
      double2 w[8], r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long o = (long long)(m0 + u) * cols + kbase + 64 * c0 + 2 * lane;
        w[u] = ld2(W + o); r[u] = ld2(RW + o);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += w[u].x + w[u].y + r[u].x + r[u].y;
    }
  } else {
    const long long total = (long long)rows * cols / 2, stride = (long long)gridDim.x * gridDim.y * 128;
    long long i = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 128 + threadIdx.x;
    for (; i + 3 * stride < total; i += 4 * stride) {
      double2 w[4], r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { w[u] = ld2(W + 2 * (i + u * stride)); r[u] = ld2(RW + 2 * (i + u * stride)); }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += w[u].x + w[u].y + r[u].x + r[u].y;
    }
  }
  if (acc == 123.456) out[0] = acc;
}

int main() {
  const int rows = 1024, cols = 2048;
  double *W, *RW, *out, *flush;
  cudaMalloc(&W, sizeof(double) * rows * cols); cudaMalloc(&RW, sizeof(double) * rows * cols); cudaMalloc(&out, 8);
  cudaMalloc(&flush, 256 << 20);
  cudaMemset(W, 0, sizeof(double) * rows * cols); cudaMemset(RW, 0, sizeof(double) * rows * cols);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int kcs[] = {64, 128, 256, 512, 2048};
  for (int pat = 0; pat < 4; ++pat)
    for (int kc : kcs) {
      dim3 grid(rows / 32, cols / kc);
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemsetAsync(flush, rep, 256 << 20);  // evict W / RW from the 126 MB L2
        cudaEventRecord(e0);
        if (pat == 0) k<0><<<grid, 128>>>(W, RW, out, rows, cols, kc);
        if (pat == 1) k<1><<<grid, 128>>>(W, RW, out, rows, cols, kc);
        if (pat == 2) k<2><<<grid, 128>>>(W, RW, out, rows, cols, kc);
        if (pat == 3) k<3><<<grid, 128>>>(W, RW, out, rows, cols, kc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("pattern %d kc %4d grid %3d x %2d (%5d warps): %7.2f us  %6.2f TB/s\n", pat, kc, grid.x, grid.y, grid.x * grid.y * 4, best * 1e3,
             2.0 * 8 * rows * cols / (best * 1e-3) / 1e12);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
