// Elementwise / reduction kernels of the path (math_funcs.go, arithmetic.go, lin_add.go).
// All are HBM-bound streaming kernels: one pass, coalesced 8-byte lanes, 4 independent elements
// in flight per thread, grid sized to a multiple of the SM count.  Each computes the value form
// and the R form in the same pass where the reference walks the vectors twice.
#include "common.h"
#include "ew.cuh"

namespace af {
namespace {

__device__ __forceinline__ double sigm(double x) { return 1.0 / (1.0 + exp(-x)); }

// column sums of an n x len row-major matrix, accumulated into out[len] (and outR from UR)
__global__ void __launch_bounds__(256) colsum_kernel(const double* __restrict__ U, const double* __restrict__ UR,
                                                     double* __restrict__ gb, double* __restrict__ rgb, long long len,
                                                     long long n, long long rows_per_block, double* __restrict__ part) {
  __shared__ double sm0[8][33];
  __shared__ double sm1[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long c = (long long)blockIdx.x * 32 + tx;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > n) r1 = n;
  double a0 = 0.0, a1 = 0.0;
  if (c < len) {
    long long r = r0 + ty;
    if (gb && rgb) {  // the common case, unrolled: eight independent loads in flight per thread
      const double* u = U + c;
      const double* ur = UR + c;
      for (; r + 24 < r1; r += 32) {
        const double x0 = u[r * len], x1 = u[(r + 8) * len], x2 = u[(r + 16) * len], x3 = u[(r + 24) * len];
        const double y0 = ur[r * len], y1 = ur[(r + 8) * len], y2 = ur[(r + 16) * len], y3 = ur[(r + 24) * len];
        a0 += (x0 + x1) + (x2 + x3);
        a1 += (y0 + y1) + (y2 + y3);
      }
    }
    for (; r < r1; r += 8) {
      if (gb) a0 += U[r * len + c];
      if (rgb) a1 += UR[r * len + c];
    }
  }
  sm0[ty][tx] = a0;
  sm1[ty][tx] = a1;
  __syncthreads();
  if (ty == 0 && c < len) {
#pragma unroll
    for (int y = 1; y < 8; ++y) { a0 += sm0[y][tx]; a1 += sm1[y][tx]; }
    if (part) {  // deterministic mode: block row y parks its partial sums, colsum_fold_kernel adds them up in y order
      part[(2ll * blockIdx.y) * len + c] = a0;
      part[(2ll * blockIdx.y + 1) * len + c] = a1;
    } else if (gridDim.y == 1) {
      if (gb) gb[c] += a0;
      if (rgb) rgb[c] += a1;
    } else {
      if (gb) atomicAdd(gb + c, a0);
      if (rgb) atomicAdd(rgb + c, a1);
    }
  }
}

// deterministic column sums, stage 2: out[c] += part[0][c] + part[1][c] + ... in block-row order
__global__ void __launch_bounds__(256) colsum_fold_kernel(const double* __restrict__ part, int nparts, double* __restrict__ gb,
                                                          double* __restrict__ rgb, long long len) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= len) return;
  double a0 = 0.0, a1 = 0.0;
  for (int y = 0; y < nparts; ++y) {
    a0 += part[(2ll * y) * len + c];
    a1 += part[(2ll * y + 1) * len + c];
  }
  if (gb) gb[c] += a0;
  if (rgb) rgb[c] += a1;
}

// stage 1 of SumAll: per-block partial sums (warp shuffle tree); stage 2: one block folds them
__global__ void __launch_bounds__(256) sum_partial_kernel(const double* __restrict__ x, const double* __restrict__ xr,
                                                          double* __restrict__ part, long long n) {
  double a = 0.0, b = 0.0;
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    a += x[i];
    if (xr) b += xr[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  __shared__ double s0[8], s1[8];
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a; s1[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { a += s0[w]; b += s1[w]; }
    part[blockIdx.x] = a;
    part[gridDim.x + blockIdx.x] = b;
  }
}
__global__ void sum_final_kernel(const double* __restrict__ part, int nparts, double* out, double* outr) {
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nparts; ++i) { a += part[i]; b += part[nparts + i]; }
    if (out) *out = a;
    if (outr) *outr = b;
  }
}

}  // namespace

int colsum_launch(af_ctx* ctx, const double* U, const double* UR, double* gb, double* rgb, long long len, long long n) {
  if (len <= 0 || n <= 0 || (!gb && !rgb)) return AF_OK;
  const long long colblocks = (len + 31) / 32;
  // 8 resident CTAs per SM: with two row loads in flight per thread that is what it takes to cover the HBM latency
  long long want = (8ll * ctx->sm_count + colblocks - 1) / colblocks;
  long long maxsplit = (n + 63) / 64;
  if (want > maxsplit) want = maxsplit;
  // deterministic mode: the partial sums of the block rows are parked in the context's scratch and folded in order
  const bool ordered = ctx->deterministic && want > 1;
  if (ordered) {
    const long long cap = ctx->scratch_doubles / (2 * len);
    if (want > cap) want = cap;
  }
  if (want < 1) want = 1;
  const long long rpb = (n + want - 1) / want;
  const long long ysplit = (n + rpb - 1) / rpb;
  dim3 grid((unsigned)colblocks, (unsigned)ysplit);
  double* part = (ordered && ysplit > 1) ? ctx->scratch : nullptr;
  colsum_kernel<<<grid, 256, 0, ctx->stream>>>(U, UR, gb, rgb, len, n, rpb, part);
  AF_TRY(after_launch(ctx, "colsum_kernel", KIND_REDUCE, 8.0 * len * n * ((gb ? 1 : 0) + (rgb ? 1 : 0))));
  if (!part) return AF_OK;
  colsum_fold_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(part, (int)ysplit, gb, rgb, len);
  return after_launch(ctx, "colsum_fold_kernel", KIND_REDUCE, 16.0 * len * ysplit);
}

}  // namespace af

using namespace af;

extern "C" {

int af_bias_fwd(af_ctx* ctx, const double* X, const double* b, double* Y, int64_t len, int64_t n) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(b); AF_NEED(Y);
  if (len < 0 || n < 0) return fail(AF_ESHAPE, "negative length");
  return ew_launch(ctx, len * n, "bias_fwd", 16, EwIn<1>{{X}}, EwOut<1>{{Y}},
                   [=] __device__(long long i, const double* x, double* y) { y[0] = x[0] + b[i % len]; });
}

int af_bias_fwd_r(af_ctx* ctx, const double* X, const double* RX, const double* b, const double* Rb, double* Y,
                  double* RY, int64_t len, int64_t n) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(b); AF_NEED(Y); AF_NEED(RY);
  if (len < 0 || n < 0) return fail(AF_ESHAPE, "negative length");
  return ew_launch(ctx, len * n, "bias_fwd_r", 32, EwIn<2>{{X, RX}}, EwOut<2>{{Y, RY}},
                   [=] __device__(long long i, const double* x, double* y) {
                     const long long c = i % len;
                     y[0] = x[0] + b[c];
                     y[1] = x[1] + (Rb ? Rb[c] : 0.0);
                   });
}

int af_bias_grad(af_ctx* ctx, const double* U, double* gb, int64_t len, int64_t n) {
  AF_CHECK_CTX(ctx); AF_NEED(U);
  return colsum_launch(ctx, U, nullptr, gb, nullptr, len, n);
}

int af_bias_grad_r(af_ctx* ctx, const double* U, const double* UR, double* gb, double* rgb, int64_t len, int64_t n) {
  AF_CHECK_CTX(ctx);
  if (gb) AF_NEED(U);
  if (rgb) AF_NEED(UR);
  return colsum_launch(ctx, U, UR, gb, rgb, len, n);
}

int af_sigmoid_fwd(af_ctx* ctx, const double* X, double* S, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(S);
  return ew_launch(ctx, len, "sigmoid_fwd", 16, EwIn<1>{{X}}, EwOut<1>{{S}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = sigm(x[0]); });
}

int af_sigmoid_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* S, double* RS, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(S); AF_NEED(RS);
  return ew_launch(ctx, len, "sigmoid_fwd_r", 32, EwIn<2>{{X, RX}}, EwOut<2>{{S, RS}},
                   [=] __device__(long long, const double* x, double* y) {
                     const double s = sigm(x[0]);
                     y[0] = s;
                     y[1] = s * (1 - s) * x[1];
                   });
}

int af_sigmoid_bwd(af_ctx* ctx, const double* S, double* U, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(S); AF_NEED(U);
  return ew_launch(ctx, len, "sigmoid_bwd", 24, EwIn<2>{{S, U}}, EwOut<1>{{U}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = x[1] * (x[0] * (1 - x[0])); });
}

int af_sigmoid_bwd_r(af_ctx* ctx, const double* S, const double* RS, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(S); AF_NEED(RS); AF_NEED(U); AF_NEED(UR);
  return ew_launch(ctx, len, "sigmoid_bwd_r", 48, EwIn<4>{{S, RS, U, UR}}, EwOut<2>{{U, UR}},
                   [=] __device__(long long, const double* v, double* y) {
                     const double x = v[0], xr = v[1], p = v[2], pr = v[3];
                     y[0] = x * (1 - x) * p;
                     y[1] = xr * (1 - x) * p - x * xr * p + x * (1 - x) * pr;
                   });
}

int af_exp_fwd(af_ctx* ctx, const double* X, double* E, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(E);
  return ew_launch(ctx, len, "exp_fwd", 16, EwIn<1>{{X}}, EwOut<1>{{E}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = exp(x[0]); });
}

int af_exp_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* E, double* RE, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(E); AF_NEED(RE);
  return ew_launch(ctx, len, "exp_fwd_r", 32, EwIn<2>{{X, RX}}, EwOut<2>{{E, RE}},
                   [=] __device__(long long, const double* x, double* y) {
                     const double e = exp(x[0]);
                     y[0] = e;
                     y[1] = e * x[1];
                   });
}

int af_exp_bwd(af_ctx* ctx, const double* E, double* U, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(E); AF_NEED(U);
  return ew_launch(ctx, len, "exp_bwd", 24, EwIn<2>{{E, U}}, EwOut<1>{{U}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = x[1] * x[0]; });
}

int af_exp_bwd_r(af_ctx* ctx, const double* E, const double* RE, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(E); AF_NEED(RE); AF_NEED(U); AF_NEED(UR);
  return ew_launch(ctx, len, "exp_bwd_r", 48, EwIn<4>{{E, RE, U, UR}}, EwOut<2>{{U, UR}},
                   [=] __device__(long long, const double* v, double* y) {
                     const double x = v[0], xr = v[1], u = v[2];
                     y[0] = u * x;
                     y[1] = v[3] * x + u * xr;
                   });
}

int af_add(af_ctx* ctx, const double* a, const double* b, double* out, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(a); AF_NEED(b); AF_NEED(out);
  return ew_launch(ctx, len, "add", 24, EwIn<2>{{a, b}}, EwOut<1>{{out}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = x[0] + x[1]; });
}

int af_sub(af_ctx* ctx, const double* a, const double* b, double* out, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(a); AF_NEED(b); AF_NEED(out);
  return ew_launch(ctx, len, "sub", 24, EwIn<2>{{a, b}}, EwOut<1>{{out}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = x[0] - x[1]; });
}

int af_mul(af_ctx* ctx, const double* a, const double* b, double* out, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(a); AF_NEED(b); AF_NEED(out);
  return ew_launch(ctx, len, "mul", 24, EwIn<2>{{a, b}}, EwOut<1>{{out}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = x[0] * x[1]; });
}

int af_mul_r(af_ctx* ctx, const double* a, const double* aR, const double* b, const double* bR, double* out,
             int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(a); AF_NEED(b); AF_NEED(out);
  return ew_launch(ctx, len, "mul_r", 40, EwIn<4>{{a, aR, b, bR}}, EwOut<1>{{out}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = x[0] * x[3] + x[1] * x[2]; });
}

int af_mul_bwd(af_ctx* ctx, const double* U, const double* UR, const double* o, const double* oR, double* D,
               double* DR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(U); AF_NEED(o); AF_NEED(D);
  return ew_launch(ctx, len, "mul_bwd", 56, EwIn<4>{{U, DR ? UR : nullptr, o, DR ? oR : nullptr}}, EwOut<2>{{D, DR}},
                   [=] __device__(long long, const double* x, double* y) {
                     y[1] = x[0] * x[3] + x[1] * x[2];
                     y[0] = x[0] * x[2];
                   });
}

int af_scale(af_ctx* ctx, double* x, double f, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(x);
  return ew_launch(ctx, len, "scale", 16, EwIn<1>{{x}}, EwOut<1>{{x}},
                   [=] __device__(long long, const double* v, double* y) { y[0] = v[0] * f; });
}

int af_axpy(af_ctx* ctx, double f, const double* x, double* y, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(x); AF_NEED(y);
  return ew_launch(ctx, len, "axpy", 24, EwIn<2>{{x, y}}, EwOut<1>{{y}},
                   [=] __device__(long long, const double* v, double* o) { o[0] = v[1] + f * v[0]; });
}

int af_square_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "square_fwd_r", 32, EwIn<2>{{X, RO ? RX : nullptr}}, EwOut<2>{{O, RO}},
                   [=] __device__(long long, const double* x, double* y) {
                     y[0] = x[0] * x[0];
                     y[1] = 2 * x[0] * x[1];
                   });
}

int af_square_bwd_r(af_ctx* ctx, const double* X, const double* RX, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(U);
  return ew_launch(ctx, len, "square_bwd_r", 48, EwIn<4>{{X, UR ? RX : nullptr, U, UR}}, EwOut<2>{{U, UR}},
                   [=] __device__(long long, const double* v, double* y) {
                     const double x = v[0], u = v[2];
                     y[1] = 2 * (v[3] * x + u * v[1]);
                     y[0] = u * (x * 2);
                   });
}

int af_sum_all(af_ctx* ctx, const double* x, const double* xR, double* out, double* outR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(x); AF_NEED(out);
  if (len < 0) return fail(AF_ESHAPE, "negative length");
  long long blocks = (len + 256 * 8 - 1) / (256 * 8);
  const long long cap = ctx->scratch_doubles / 2;
  if (blocks > cap) blocks = cap;
  if (blocks > 4ll * ctx->sm_count) blocks = 4ll * ctx->sm_count;
  if (blocks < 1) blocks = 1;
  sum_partial_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(x, (outR && xR) ? xR : nullptr, ctx->scratch, len);
  AF_TRY(after_launch(ctx, "sum_partial_kernel", KIND_REDUCE, 8.0 * len * ((outR && xR) ? 2 : 1)));
  sum_final_kernel<<<1, 32, 0, ctx->stream>>>(ctx->scratch, (int)blocks, out, outR);
  return after_launch(ctx, "sum_final_kernel", KIND_REDUCE, 16.0 * blocks);
}

int af_fill(af_ctx* ctx, double* x, double value, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(x);
  return ew_launch(ctx, len, "fill", 8, EwIn<1>{{nullptr}}, EwOut<1>{{x}},
                   [=] __device__(long long, const double*, double* y) { y[0] = value; });
}

}  // extern "C"
