// Shared launchers for the HBM-bound streaming kernels (elementwise.cu, mathops.cu).
//   ew_launch      one pass, coalesced 8-byte lanes, loads of 4 elements x all inputs in flight per thread,
//                  grid capped at 8 CTAs per SM
//   reduce2_launch two sums in one pass: per-block warp-shuffle tree -> ctx->scratch -> one folding
//                  block (deterministic: no atomics), results written or accumulated on the device
#pragma once
#include "common.h"

namespace af {

constexpr int kEwThreads = 256;
constexpr int kEwUnroll = 4;

// Two-phase streaming kernel: a thread first issues the loads of ALL inputs of its kEwUnroll elements
// (NIN x kEwUnroll independent 8-byte loads in flight), then computes, then stores.  The phases are explicit
// because the input and output vectors may alias (the backward kernels work in place), so the compiler cannot
// hoist the loads of element u+1 above the stores of element u on its own -- with one element in flight per
// thread `add` reached 0.73 of the measured HBM bandwidth, with this layout it streams at full rate.
// A NULL input reads as 0 (identically-zero R operand, variable.go:85-91); a NULL output is not stored.
template <int N> struct EwIn { const double* p[N]; };
template <int N> struct EwOut { double* p[N]; };

template <int NIN, int NOUT, typename F>
__global__ void __launch_bounds__(kEwThreads) ew_kernel(long long n, EwIn<NIN> in, EwOut<NOUT> out, F f) {
  const long long stride = (long long)gridDim.x * kEwThreads;
  long long i = (long long)blockIdx.x * kEwThreads + threadIdx.x;
  for (; i + (kEwUnroll - 1) * stride < n; i += kEwUnroll * stride) {
    double x[kEwUnroll][NIN], y[kEwUnroll][NOUT];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u)
#pragma unroll
      for (int k = 0; k < NIN; ++k) x[u][k] = in.p[k] ? in.p[k][i + u * stride] : 0.0;
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) f(i + u * stride, x[u], y[u]);
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u)
#pragma unroll
      for (int k = 0; k < NOUT; ++k)
        if (out.p[k]) out.p[k][i + u * stride] = y[u][k];
  }
  for (; i < n; i += stride) {
    double x[NIN], y[NOUT];
#pragma unroll
    for (int k = 0; k < NIN; ++k) x[k] = in.p[k] ? in.p[k][i] : 0.0;
    f(i, x, y);
#pragma unroll
    for (int k = 0; k < NOUT; ++k)
      if (out.p[k]) out.p[k][i] = y[k];
  }
}

// `bpe` = algorithmic bytes per element (each distinct vector read or written once);
// f(i, x, y): x[k] = value of input k at i, y[k] = value to store to output k
template <int NIN, int NOUT, typename F>
int ew_launch(af_ctx* ctx, long long n, const char* what, int bpe, EwIn<NIN> in, EwOut<NOUT> out, F f) {
  if (n <= 0) return AF_OK;
  long long blocks = (n + (long long)kEwThreads * kEwUnroll - 1) / ((long long)kEwThreads * kEwUnroll);
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  ew_kernel<NIN, NOUT><<<(unsigned)blocks, kEwThreads, 0, ctx->stream>>>(n, in, out, f);
  return after_launch(ctx, what, KIND_ELEMENTWISE, (double)n * bpe);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// stage 1: f(i) -> (a, b) summed (or max-combined when MAX) per block into part[0..grid) | part[grid..2grid)
template <bool MAX, typename F>
__global__ void __launch_bounds__(256) reduce2_partial_kernel(long long n, double* __restrict__ part, F f) {
  double a = 0.0, b = 0.0;
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    const double2 v = f(i);
    if (MAX) { a = fmax(a, v.x); b = fmax(b, v.y); }
    else { a += v.x; b += v.y; }
  }
  a = MAX ? warp_max(a) : warp_sum(a);
  b = MAX ? warp_max(b) : warp_sum(b);
  __shared__ double s0[8], s1[8];
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a; s1[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      if (MAX) { a = fmax(a, s0[w]); b = fmax(b, s1[w]); }
      else { a += s0[w]; b += s1[w]; }
    }
    part[blockIdx.x] = a;
    part[gridDim.x + blockIdx.x] = b;
  }
}

// stage 2: fold the partials; G(a, b) finishes on the device (writes the scalar outputs)
template <bool MAX, typename G>
__global__ void reduce2_final_kernel(const double* __restrict__ part, int nparts, G g) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 32) {
    if (MAX) { a = fmax(a, part[i]); b = fmax(b, part[nparts + i]); }
    else { a += part[i]; b += part[nparts + i]; }
  }
  a = MAX ? warp_max(a) : warp_sum(a);
  b = MAX ? warp_max(b) : warp_sum(b);
  if (threadIdx.x == 0) g(a, b);
}

template <bool MAX, typename F, typename G>
int reduce2_launch(af_ctx* ctx, long long n, const char* what, int bpe, F f, G g) {
  if (n < 0) return fail(AF_ESHAPE, "negative length");
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  const long long cap = ctx->scratch_doubles / 2;
  if (blocks > cap) blocks = cap;
  if (blocks > 4ll * ctx->sm_count) blocks = 4ll * ctx->sm_count;
  if (blocks < 1) blocks = 1;
  reduce2_partial_kernel<MAX><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, ctx->scratch, f);
  AF_TRY(after_launch(ctx, what, KIND_REDUCE, (double)n * bpe));
  reduce2_final_kernel<MAX><<<1, 32, 0, ctx->stream>>>(ctx->scratch, (int)blocks, g);
  return after_launch(ctx, "reduce2_final_kernel", KIND_REDUCE, 16.0 * blocks);
}

#define AF_CHECK_CTX(ctx) \
  if (!(ctx)) return af::fail(AF_EINVAL, "null context"); \
  af::DeviceGuard _guard((ctx)->device)
#define AF_NEED(p) \
  if (!(p)) return af::fail(AF_EINVAL, "null pointer argument: " #p)

}  // namespace af
