// the elementwise functions, scalar-broadcast ops, cost heads and the parameter update that sit
// either side of the dense path (math_funcs.go: Log, LogSigmoid, Sin, Norm, Softmax; arithmetic.go: AddScaler, Div,
// ScaleFirst, AddFirst, Inverse, Pow, the log-domain sums; gradient.go: AddToVars).
//
// Same design as elementwise.cu: every kernel is one HBM pass, computes the value form and the R form together, and the
// backward kernels overwrite the upstreams in place as the reference does (result.go:24-29).  Scalars that the
// reference reads on the host (`scaler.Output()[0]`, `MaxAbs()`, the upstream of a SumAll) stay on the device and are
// passed BY POINTER, so a chain of these ops never synchronises the stream.
#include "common.h"
#include "ew.cuh"

namespace af {
namespace {

__device__ __forceinline__ double rd(const double* p, long long i) { return p ? p[i] : 0.0; }

// ---- row-wise softmax ----------------------------------------------------------------------------
// A group of G threads owns one row (G = 32: one warp, rows up to ~1k entries; G = 256: one block).
template <int G>
__device__ __forceinline__ double group_sum(double v, double* sm) {
  v = warp_sum(v);
  if (G == 32) return v;
  const int w = threadIdx.x >> 5;
  __syncthreads();  // sm may still be read by the previous reduction
  if ((threadIdx.x & 31) == 0) sm[w] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < G / 32; ++i) t += sm[i];
  return t;
}

// Y = e / sum(e), e = exp(x * invT);  RY = e*invR + eR*inv with eR = e*xR*invT, inv = 1/S, invR = -inv^2 * sum(eR)
// (math_funcs.go:487-509 = Scale, Exp, SumAll, Inverse, ScaleFirst; formulas in the reference's operand order)
template <int G>
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const double* __restrict__ X, const double* __restrict__ RX,
                                                          double* __restrict__ Y, double* __restrict__ RY, double invT,
                                                          long long len, long long n) {
  __shared__ double sm[8];
  const int lane = threadIdx.x % G;
  const long long row = (long long)blockIdx.x * (256 / G) + threadIdx.x / G;
  const bool live = row < n;  // G == 256: whole block shares `live`; G == 32: whole warp does
  if (G == 32 && !live) return;
  const double* x = X + row * len;
  const double* xr = RX ? RX + row * len : nullptr;
  double s = 0.0, sr = 0.0;
  if (live)
    for (long long k = lane; k < len; k += G) {
      const double e = exp(x[k] * invT);
      s += e;
      if (RY) sr += e * (rd(xr, k) * invT);
    }
  s = group_sum<G>(s, sm);
  if (RY) sr = group_sum<G>(sr, sm);
  if (!live) return;
  const double inv = 1.0 / s;
  const double invR = -(inv * inv) * sr;
  for (long long k = lane; k < len; k += G) {
    const double e = exp(x[k] * invT);
    Y[row * len + k] = e * inv;
    if (RY) RY[row * len + k] = e * invR + (e * (rd(xr, k) * invT)) * inv;
  }
}

// in place: U <- invT * y (u - <u,y>);  UR <- invT * [ yR (u - <u,y>) + y (uR - <uR,y> - <u,yR>) ]
// (the reference's chain ScaleFirst -> Inverse -> SumAll -> Pool -> Exp -> Scale, collapsed algebraically)
template <int G>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const double* __restrict__ Y, const double* __restrict__ RY,
                                                          double* __restrict__ U, double* __restrict__ UR, double invT,
                                                          long long len, long long n) {
  __shared__ double sm[8];
  const int lane = threadIdx.x % G;
  const long long row = (long long)blockIdx.x * (256 / G) + threadIdx.x / G;
  const bool live = row < n;
  if (G == 32 && !live) return;
  const long long o = row * len;
  double a = 0.0, b = 0.0;  // <u,y> ; <uR,y> + <u,yR>
  if (live)
    for (long long k = lane; k < len; k += G) {
      const double u = U[o + k], y = Y[o + k];
      a += u * y;
      if (UR) b += UR[o + k] * y + u * rd(RY, o + k);
    }
  a = group_sum<G>(a, sm);
  if (UR) b = group_sum<G>(b, sm);
  if (!live) return;
  for (long long k = lane; k < len; k += G) {
    const double u = U[o + k], y = Y[o + k];
    if (UR) UR[o + k] = invT * (rd(RY, o + k) * (u - a) + y * (UR[o + k] - b));
    U[o + k] = invT * (y * (u - a));
  }
}

template <typename K32, typename K256>
void row_launch(af_ctx* ctx, long long len, long long n, K32 k32, K256 k256) {
  if (len <= 1024) k32((unsigned)((n + 7) / 8));
  else k256((unsigned)n);
}

}  // namespace
}  // namespace af

using namespace af;

extern "C" {

// ---- Log (math_funcs.go:99-189) --------------------------------------------------------------------
int af_log_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "log_fwd_r", RO ? 32 : 16, [=] __device__(long long i) {
    const double x = X[i];
    O[i] = log(x);
    if (RO) RO[i] = rd(RX, i) / x;
  });
}

int af_log_bwd_r(af_ctx* ctx, const double* X, const double* RX, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(U);
  return ew_launch(ctx, len, "log_bwd_r", UR ? 48 : 24, [=] __device__(long long i) {
    const double x = X[i], u = U[i];
    if (UR) UR[i] = UR[i] / x - u * rd(RX, i) / (x * x);
    U[i] = UR ? u / x : u * (1 / x);  // math_funcs.go:150 multiplies by 1/x, :184 divides
  });
}

// ---- LogSigmoid (math_funcs.go:376-477) -------------------------------------------------------------
int af_logsigmoid_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "logsigmoid_fwd_r", RO ? 32 : 16, [=] __device__(long long i) {
    const double x = X[i];
    O[i] = x > 0 ? -log(1 + exp(-x)) : x - log(1 + exp(x));
    if (RO) RO[i] = rd(RX, i) / (1 + exp(x));
  });
}

int af_logsigmoid_bwd_r(af_ctx* ctx, const double* X, const double* RX, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(U);
  return ew_launch(ctx, len, "logsigmoid_bwd_r", UR ? 48 : 24, [=] __device__(long long i) {
    const double u = U[i], ex = exp(X[i]);
    if (UR) {
      const double s = 1 / (1 + ex);
      UR[i] = UR[i] * s - u * s * s * ex * rd(RX, i);
      U[i] = u * s;
    } else {
      U[i] = u / (1 + ex);
    }
  });
}

// ---- Sin (math_funcs.go:511-597); Cos = Sin(AddScaler(x, pi/2)) on the host side --------------------
int af_sin_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "sin_fwd_r", RO ? 32 : 16, [=] __device__(long long i) {
    double s, c;
    sincos(X[i], &s, &c);
    O[i] = s;
    if (RO) RO[i] = c * rd(RX, i);
  });
}

int af_sin_bwd_r(af_ctx* ctx, const double* X, const double* RX, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(U);
  return ew_launch(ctx, len, "sin_bwd_r", UR ? 48 : 24, [=] __device__(long long i) {
    double s, c;
    sincos(X[i], &s, &c);
    const double u = U[i];
    if (UR) UR[i] = UR[i] * c + u * (-s * rd(RX, i));
    U[i] = u * c;
  });
}

// ---- Tanh: extension named by the north star, same contract as Sigmoid (T, RT kept by the caller) ----
int af_tanh_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* T, double* RT, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(T);
  return ew_launch(ctx, len, "tanh_fwd_r", RT ? 32 : 16, [=] __device__(long long i) {
    const double t = tanh(X[i]);
    T[i] = t;
    if (RT) RT[i] = (1 - t * t) * rd(RX, i);
  });
}

int af_tanh_bwd_r(af_ctx* ctx, const double* T, const double* RT, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(T); AF_NEED(U);
  return ew_launch(ctx, len, "tanh_bwd_r", UR ? 48 : 24, [=] __device__(long long i) {
    const double t = T[i], u = U[i];
    if (UR) UR[i] = (1 - t * t) * UR[i] - 2 * t * rd(RT, i) * u;
    U[i] = (1 - t * t) * u;
  });
}

// ---- Inverse (arithmetic.go:772-867) ------------------------------------------------------------------
int af_inverse_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "inverse_fwd_r", RO ? 32 : 16, [=] __device__(long long i) {
    const double r = 1 / X[i];
    O[i] = r;
    if (RO) RO[i] = -(r * r) * rd(RX, i);
  });
}

// O = the forward output 1/x
int af_inverse_bwd_r(af_ctx* ctx, const double* O, const double* RX, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(O); AF_NEED(U);
  return ew_launch(ctx, len, "inverse_bwd_r", UR ? 48 : 24, [=] __device__(long long i) {
    const double o = O[i], u = U[i], sq = -(o * o);
    if (UR) UR[i] = sq * UR[i] + -2 * (sq * o) * rd(RX, i) * u;
    U[i] = sq * u;
  });
}

// ---- Pow (arithmetic.go:869-966) ----------------------------------------------------------------------
int af_pow_fwd_r(af_ctx* ctx, const double* X, const double* RX, double p, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "pow_fwd_r", RO ? 32 : 16, [=] __device__(long long i) {
    const double x = X[i];
    O[i] = pow(x, p);
    if (RO) RO[i] = p != 0 ? p * pow(x, p - 1) * rd(RX, i) : 0.0;
  });
}

int af_pow_bwd_r(af_ctx* ctx, const double* X, const double* RX, double p, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(U);
  if (p == 1) return AF_OK;  // arithmetic.go:898,952: upstream passes through untouched
  return ew_launch(ctx, len, "pow_bwd_r", UR ? 48 : 24, [=] __device__(long long i) {
    const double x = X[i], u = U[i];
    const double d1 = p * pow(x, p - 1);
    if (UR) UR[i] = UR[i] * d1 + u * (p * (p - 1) * pow(x, p - 2) * rd(RX, i));
    U[i] = u * d1;
  });
}

// ---- Div (arithmetic.go:378-423; DivR is MulR(a, InverseR(b)) and needs no kernel of its own) --------
int af_div(af_ctx* ctx, const double* A, const double* B, double* O, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(A); AF_NEED(B); AF_NEED(O);
  return ew_launch(ctx, len, "div", 24, [=] __device__(long long i) { O[i] = A[i] / B[i]; });
}

// DA = U / B (numerator's upstream), DB = U * (-O / B) (denominator's); either may be NULL
int af_div_bwd(af_ctx* ctx, const double* U, const double* O, const double* B, double* DA, double* DB, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(U); AF_NEED(O); AF_NEED(B);
  return ew_launch(ctx, len, "div_bwd", 40, [=] __device__(long long i) {
    const double u = U[i], b = B[i];
    if (DA) DA[i] = u / b;
    if (DB) DB[i] = u * (-O[i] / b);
  });
}

// ---- scalar-broadcast ops: AddScaler, ScaleFirst, AddFirst, SumAll backward ---------------------------
int af_add_scaler(af_ctx* ctx, const double* X, double f, double* O, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "add_scaler", 16, [=] __device__(long long i) { O[i] = X[i] + f; });
}

// O = X + sign * f[0] with f on the device (AddFirst, arithmetic.go:595-612; the "- maxVal" of the log-domain sums)
int af_add_dev(af_ctx* ctx, const double* X, const double* d_f, double sign, double* O, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(d_f); AF_NEED(O);
  return ew_launch(ctx, len, "add_dev", 16, [=] __device__(long long i) { O[i] = X[i] + sign * d_f[0]; });
}

// O = X * f[0];  RO = X * fR[0] + RX * f[0]   (ScaleFirst forward and, applied to the upstreams, its backward:
// arithmetic.go:503-510,544-553,583-592).  O may alias X and RO may alias RX.
int af_scale_dev(af_ctx* ctx, const double* X, const double* RX, const double* d_f, const double* d_fR, double* O,
                 double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(d_f); AF_NEED(O);
  return ew_launch(ctx, len, "scale_dev", RO ? 32 : 16, [=] __device__(long long i) {
    const double x = X[i], f = d_f[0];
    if (RO) RO[i] = x * (d_fR ? d_fR[0] : 0.0) + rd(RX, i) * f;
    O[i] = x * f;
  });
}

// x[i] = value[0]: SumAll's backward broadcast (arithmetic.go:987-995,1035-1047) without a host round trip
int af_fill_dev(af_ctx* ctx, double* x, const double* d_value, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(x); AF_NEED(d_value);
  return ew_launch(ctx, len, "fill_dev", 8, [=] __device__(long long i) { x[i] = d_value[0]; });
}

// out[0] = sum a.b (+ sum c.d when c, d are given): DotFast at arithmetic.go:523,571-573
int af_dot(af_ctx* ctx, const double* a, const double* b, const double* c, const double* d, double* out, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(a); AF_NEED(b); AF_NEED(out);
  if ((c == nullptr) != (d == nullptr)) return fail(AF_EINVAL, "af_dot: c and d must be given together");
  return reduce2_launch<false>(
      ctx, len, "dot_partial_kernel", c ? 32 : 16,
      [=] __device__(long long i) { return make_double2(a[i] * b[i], c ? c[i] * d[i] : 0.0); },
      [=] __device__(double s0, double s1) { out[0] = s0 + s1; });
}

// out[0] = max_i |x[i]| (linalg.Vector.MaxAbs at arithmetic.go:1057,1078); combine != 0 keeps max(out[0], ...)
int af_max_abs(af_ctx* ctx, const double* x, double* out, int combine, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(out);
  if (len > 0) AF_NEED(x);
  return reduce2_launch<true>(
      ctx, len, "max_abs_partial_kernel", 8, [=] __device__(long long i) { return make_double2(fabs(x[i]), 0.0); },
      [=] __device__(double m, double) { out[0] = combine ? fmax(out[0], m) : m; });
}

// ---- Norm (math_funcs.go:201-284) ----------------------------------------------------------------------
int af_norm_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* out, double* outR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(out);
  return reduce2_launch<false>(
      ctx, len, "norm_partial_kernel", outR ? 16 : 8,
      [=] __device__(long long i) {
        const double x = X[i];
        return make_double2(x * x, outR ? x * rd(RX, i) : 0.0);
      },
      [=] __device__(double ss, double xr) {
        const double o = sqrt(ss);
        out[0] = o;
        if (outR) outR[0] = (1 / o) * xr;
      });
}

// D = x * (u/o);  DR = x * (-u*ro/o^2 + uR/o) + xR * (u/o)   with o, ro, u, uR device scalars
int af_norm_bwd_r(af_ctx* ctx, const double* X, const double* RX, const double* d_o, const double* d_ro,
                  const double* d_u, const double* d_uR, double* D, double* DR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(d_o); AF_NEED(d_u); AF_NEED(D);
  if (DR) { AF_NEED(d_ro); AF_NEED(d_uR); }
  return ew_launch(ctx, len, "norm_bwd_r", DR ? 32 : 16, [=] __device__(long long i) {
    const double o = d_o[0], u = d_u[0], s = u / o, x = X[i];
    if (DR) DR[i] = x * (-u * d_ro[0] / (o * o) + d_uR[0] / o) + rd(RX, i) * s;
    D[i] = x * s;
  });
}

// ---- Softmax, one per row of an n x len packed batch (math_funcs.go:479-509 under RFuncBatcher) --------
int af_softmax_fwd_r(af_ctx* ctx, const double* X, const double* RX, double temperature, double* Y, double* RY,
                     int64_t len, int64_t n) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(Y);
  if (len < 0 || n < 0) return fail(AF_ESHAPE, "negative length");
  if (len == 0 || n == 0) return AF_OK;
  const double invT = (temperature != 0 && temperature != 1) ? 1 / temperature : 1.0;
  row_launch(ctx, len, n,
             [&](unsigned g) { softmax_fwd_kernel<32><<<g, 256, 0, ctx->stream>>>(X, RX, Y, RY, invT, len, n); },
             [&](unsigned g) { softmax_fwd_kernel<256><<<g, 256, 0, ctx->stream>>>(X, RX, Y, RY, invT, len, n); });
  return after_launch(ctx, "softmax_fwd_kernel", KIND_REDUCE, 8.0 * len * n * (RY ? 4 : 2));
}

int af_softmax_bwd_r(af_ctx* ctx, const double* Y, const double* RY, double temperature, double* U, double* UR,
                     int64_t len, int64_t n) {
  AF_CHECK_CTX(ctx); AF_NEED(Y); AF_NEED(U);
  if (len < 0 || n < 0) return fail(AF_ESHAPE, "negative length");
  if (len == 0 || n == 0) return AF_OK;
  const double invT = (temperature != 0 && temperature != 1) ? 1 / temperature : 1.0;
  row_launch(ctx, len, n,
             [&](unsigned g) { softmax_bwd_kernel<32><<<g, 256, 0, ctx->stream>>>(Y, RY, U, UR, invT, len, n); },
             [&](unsigned g) { softmax_bwd_kernel<256><<<g, 256, 0, ctx->stream>>>(Y, RY, U, UR, invT, len, n); });
  return after_launch(ctx, "softmax_bwd_kernel", KIND_REDUCE, 8.0 * len * n * (UR ? 6 : 3));
}

// ---- squared-error cost head: c = 0.5 * ||a - y||^2 = Scale(SquaredNorm(Sub(a, y)), .5) ----------------
// (math_funcs.go:191-199, arithmetic.go:105-182,436-493,696-770,972-1047 collapsed into one reduction and one
// elementwise pass)   cR = sum (a - y)(aR - yR)
int af_sqerr_fwd_r(af_ctx* ctx, const double* A, const double* RA, const double* Y, const double* RY, double* out,
                   double* outR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(A); AF_NEED(Y); AF_NEED(out);
  return reduce2_launch<false>(
      ctx, len, "sqerr_partial_kernel", outR ? 32 : 16,
      [=] __device__(long long i) {
        const double d = A[i] - Y[i];
        return make_double2(d * d, outR ? 2 * d * (rd(RA, i) - rd(RY, i)) : 0.0);
      },
      [=] __device__(double s, double sr) {
        out[0] = s * 0.5;
        if (outR) outR[0] = sr * 0.5;
      });
}

// upstream (u, uR: device scalars) on the cost -> upstream on `a`:  G = (.5u)(2d);  GR = 2((.5uR) d + (.5u) dR);
// the upstream on `y` is the negation (arithmetic.go:175-182).  sign = +1 for a, -1 for y.
int af_sqerr_bwd_r(af_ctx* ctx, const double* A, const double* RA, const double* Y, const double* RY, const double* d_u,
                   const double* d_uR, double sign, double* G, double* GR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(A); AF_NEED(Y); AF_NEED(d_u); AF_NEED(G);
  if (GR) AF_NEED(d_uR);
  return ew_launch(ctx, len, "sqerr_bwd_r", GR ? 48 : 24, [=] __device__(long long i) {
    const double d = A[i] - Y[i], u = d_u[0] * 0.5;
    if (GR) GR[i] = sign * (2 * ((d_uR[0] * 0.5) * d + u * (rd(RA, i) - rd(RY, i))));
    G[i] = sign * (u * (d * 2));
  });
}

}  // extern "C"
