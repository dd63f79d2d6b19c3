// CPU restatement of the C ABI (include/autofunc_b200.h) -- TEST INFRASTRUCTURE ONLY.
//
// Built by tests/ into oracle/_build/libautofunc_b200_ref.so so that the C++ host layer
// (include/autofunc_b200.hpp: graph logic, gating, accumulation, ownership, panic texts) can be exercised by the
// `-m "not gpu"` suite with the reference's own tests (tests/cpp/host_tests.cpp), where no GPU exists.  "Device"
// pointers are host pointers here.  Nothing in autofunc_b200/ links, loads or ships this file; the product library
// has no CPU path (af_ctx_create fails without an sm_100 device).  Scalar loops only, each formula citing the
// reference file:line it restates; parity status as in oracle/autofunc_ref.py (pinned on the reference's KATs and
// FullCheck at 1e-5; below that defined by the oracle).
//
// Covered: every entry point the host layer calls except the LSTM plan and recorded graphs (af_lstm_*, af_graph_*
// return AF_ENOTSUP).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "autofunc_b200.h"

struct af_ctx {
  int64_t launches = 0;
};

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& m) {
  g_err = m;
  return code;
}
#define CTX(c) \
  if (!(c)) return fail(AF_EINVAL, "null context"); \
  (c)->launches++
#define NEED(p) \
  if (!(p)) return fail(AF_EINVAL, "null pointer argument")
#define SHAPE(rows, cols, n) \
  if ((rows) < 0 || (cols) < 0 || (n) < 0) return fail(AF_ESHAPE, "negative dimension")

double sigmoid(double x) { return 1 / (1 + std::exp(-x)); }  // math_funcs.go:296
}  // namespace

struct af_mlp {
  af_ctx* ctx;
  std::vector<int64_t> sizes;
  int64_t n;
  std::vector<std::vector<double>> param, rvec, grad, rgrad;  // 2L each
  std::vector<bool> has_r;
  std::vector<std::vector<double>> act, ract;  // L+1
  std::vector<double> up, upr;
  int L() const { return (int)sizes.size() - 1; }
  int64_t plen(int i) const { return (i & 1) ? sizes[i / 2 + 1] : sizes[i / 2] * sizes[i / 2 + 1]; }
};

extern "C" {

const char* af_last_error(void) { return g_err.c_str(); }
int af_version(void) { return 1; }
int af_ctx_create(int, af_ctx** out) {
  if (!out) return fail(AF_EINVAL, "null argument");
  *out = new af_ctx();
  return AF_OK;
}
int af_ctx_destroy(af_ctx* c) {
  delete c;
  return AF_OK;
}
int af_sync(af_ctx* c) { return c ? AF_OK : fail(AF_EINVAL, "null context"); }
int64_t af_launch_count(af_ctx* c) { return c ? c->launches : -1; }

int af_malloc(af_ctx* c, int64_t n, double** out) {
  if (!c || !out) return fail(AF_EINVAL, "null argument");
  *out = (double*)std::malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
  return *out ? AF_OK : fail(AF_ENOMEM, "out of memory");
}
int af_free(af_ctx*, double* p) {
  std::free(p);
  return AF_OK;
}
int af_upload(af_ctx* c, double* d, const double* h, int64_t n) {
  if (!c) return fail(AF_EINVAL, "null context");
  if (n > 0) std::memcpy(d, h, (size_t)n * 8);
  return AF_OK;
}
int af_download(af_ctx* c, double* h, const double* d, int64_t n) { return af_upload(c, h, d, n); }
int af_zero(af_ctx* c, double* p, int64_t n) {
  CTX(c);
  if (n > 0) std::memset(p, 0, (size_t)n * 8);
  return AF_OK;
}
int af_copy(af_ctx* c, double* d, const double* s, int64_t n) {
  CTX(c);
  if (n > 0) std::memmove(d, s, (size_t)n * 8);
  return AF_OK;
}

// slices.go:13-119,130-206 on packed batches: rows x cols block copy between pitched buffers
int af_copy_2d(af_ctx* c, double* dst, int64_t ld_dst, const double* src, int64_t ld_src, int64_t rows, int64_t cols) {
  CTX(c);
  for (int64_t r = 0; r < rows; r++)
    if (cols > 0) std::memmove(dst + r * ld_dst, src + r * ld_src, (size_t)cols * 8);
  return AF_OK;
}

// ---- LinTran (lin_tran.go) ------------------------------------------------------------------------------------
int af_lintran_fwd_r(af_ctx* c, const double* W, const double* RW, const double* X, const double* RX, double* Y,
                     double* RY, int64_t rows, int64_t cols, int64_t n) {
  CTX(c);
  SHAPE(rows, cols, n);
  for (int64_t s = 0; s < n; s++)
    for (int64_t r = 0; r < rows; r++) {
      double y = 0, ry = 0;
      for (int64_t k = 0; k < cols; k++) {
        y += X[s * cols + k] * W[r * cols + k];                // lin_tran.go:79-117
        if (RW) ry += X[s * cols + k] * RW[r * cols + k];      // lin_tran.go:119-177
        if (RX) ry += RX[s * cols + k] * W[r * cols + k];
      }
      if (Y) Y[s * rows + r] = y;
      if (RY) RY[s * rows + r] = ry;
    }
  return AF_OK;
}
int af_lintran_fwd(af_ctx* c, const double* W, const double* X, double* Y, int64_t rows, int64_t cols, int64_t n) {
  return af_lintran_fwd_r(c, W, nullptr, X, nullptr, Y, nullptr, rows, cols, n);
}
int af_lintran_wgrad_r(af_ctx* c, const double* U, const double* UR, const double* X, const double* RX, double* gW,
                       double* rgW, int64_t rows, int64_t cols, int64_t n) {
  CTX(c);
  SHAPE(rows, cols, n);
  for (int64_t r = 0; r < rows; r++)
    for (int64_t k = 0; k < cols; k++) {
      double g = 0, rg = 0;
      for (int64_t s = 0; s < n; s++) {
        g += U[s * rows + r] * X[s * cols + k];                       // lin_tran.go:179-212
        if (rgW && RX) rg += U[s * rows + r] * RX[s * cols + k];      // lin_tran.go:214-270
        if (rgW) rg += UR[s * rows + r] * X[s * cols + k];
      }
      if (gW) gW[r * cols + k] += g;
      if (rgW) rgW[r * cols + k] += rg;
    }
  return AF_OK;
}
int af_lintran_wgrad(af_ctx* c, const double* U, const double* X, double* gW, int64_t rows, int64_t cols, int64_t n) {
  return af_lintran_wgrad_r(c, U, nullptr, X, nullptr, gW, nullptr, rows, cols, n);
}
int af_lintran_igrad_r(af_ctx* c, const double* U, const double* UR, const double* W, const double* RW, double* D,
                       double* DR, int64_t rows, int64_t cols, int64_t n) {
  CTX(c);
  SHAPE(rows, cols, n);
  for (int64_t s = 0; s < n; s++)
    for (int64_t k = 0; k < cols; k++) {
      double d = 0, dr = 0;
      for (int64_t r = 0; r < rows; r++) {
        d += U[s * rows + r] * W[r * cols + k];                     // lin_tran.go:272-306
        if (DR) dr += UR[s * rows + r] * W[r * cols + k];           // lin_tran.go:308-359
        if (DR && RW) dr += U[s * rows + r] * RW[r * cols + k];
      }
      if (D) D[s * cols + k] = d;
      if (DR) DR[s * cols + k] = dr;
    }
  return AF_OK;
}
int af_lintran_igrad(af_ctx* c, const double* U, const double* W, double* D, int64_t rows, int64_t cols, int64_t n) {
  return af_lintran_igrad_r(c, U, nullptr, W, nullptr, D, nullptr, rows, cols, n);
}
// linTranRResult.PropagateRGradient as one call (lin_tran.go:393-425): both gradients; accumulate_input: D += , DR +=
int af_lintran_backward_r(af_ctx* c, const double* U, const double* UR, const double* W, const double* RW, const double* X,
                          const double* RX, double* gW, double* rgW, double* D, double* DR, int64_t rows, int64_t cols, int64_t n,
                          int accumulate_input) {
  CTX(c);
  SHAPE(rows, cols, n);
  if (gW || rgW) {
    int rc = af_lintran_wgrad_r(c, U, UR, X, RX, gW, rgW, rows, cols, n);
    if (rc) return rc;
  }
  if (!D && !DR) return AF_OK;
  if (!accumulate_input) return af_lintran_igrad_r(c, U, UR, W, RW, D, DR, rows, cols, n);
  std::vector<double> d((size_t)(cols * n)), dr((size_t)(cols * n));
  int rc = af_lintran_igrad_r(c, U, UR, W, RW, D ? d.data() : nullptr, DR ? dr.data() : nullptr, rows, cols, n);
  if (rc) return rc;
  for (int64_t i = 0; i < cols * n; i++) {
    if (D) D[i] += d[i];
    if (DR) DR[i] += dr[i];
  }
  return AF_OK;
}
int af_lintran_backward(af_ctx* c, const double* U, const double* W, const double* X, double* gW, double* D, int64_t rows,
                        int64_t cols, int64_t n, int accumulate_input) {
  return af_lintran_backward_r(c, U, nullptr, W, nullptr, X, nullptr, gW, nullptr, D, nullptr, rows, cols, n, accumulate_input);
}

// ---- LinAdd, batched (lin_add.go:13-109) ----------------------------------------------------------------------
int af_bias_fwd_r(af_ctx* c, const double* X, const double* RX, const double* b, const double* Rb, double* Y,
                  double* RY, int64_t len, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < len * n; i++) {
    Y[i] = X[i] + b[i % len];
    if (RY) RY[i] = (RX ? RX[i] : 0) + (Rb ? Rb[i % len] : 0);
  }
  return AF_OK;
}
int af_bias_fwd(af_ctx* c, const double* X, const double* b, double* Y, int64_t len, int64_t n) {
  return af_bias_fwd_r(c, X, nullptr, b, nullptr, Y, nullptr, len, n);
}
int af_bias_grad_r(af_ctx* c, const double* U, const double* UR, double* gb, double* rgb, int64_t len, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < len * n; i++) {
    if (gb) gb[i % len] += U[i];
    if (rgb) rgb[i % len] += UR[i];
  }
  return AF_OK;
}
int af_bias_grad(af_ctx* c, const double* U, double* gb, int64_t len, int64_t n) {
  return af_bias_grad_r(c, U, nullptr, gb, nullptr, len, n);
}

// ---- elementwise (math_funcs.go) ---------------------------------------------------------------------------------
#define EW_FWD(name, body)                                                                                     \
  int name(af_ctx* c, const double* X, const double* RX, double* O, double* RO, int64_t len) {                 \
    CTX(c);                                                                                                    \
    for (int64_t i = 0; i < len; i++) {                                                                        \
      const double x = X[i], xr = RX ? RX[i] : 0;                                                              \
      double o, ro;                                                                                            \
      body;                                                                                                    \
      O[i] = o;                                                                                                \
      if (RO) RO[i] = ro;                                                                                      \
    }                                                                                                          \
    return AF_OK;                                                                                              \
  }
// saved s / sr: forward output or input (see the header); u, ur in place
#define EW_BWD(name, body)                                                                                     \
  int name(af_ctx* c, const double* S, const double* SR, double* U, double* UR, int64_t len) {                 \
    CTX(c);                                                                                                    \
    for (int64_t i = 0; i < len; i++) {                                                                        \
      const double s = S[i], sr = SR ? SR[i] : 0, u = U[i], ur = UR ? UR[i] : 0;                               \
      double nu, nur;                                                                                          \
      body;                                                                                                    \
      U[i] = nu;                                                                                               \
      if (UR) UR[i] = nur;                                                                                     \
    }                                                                                                          \
    return AF_OK;                                                                                              \
  }
// math_funcs.go:286-374
EW_FWD(af_sigmoid_fwd_r, o = sigmoid(x); ro = o * (1 - o) * xr)
EW_BWD(af_sigmoid_bwd_r, nu = s * (1 - s) * u; nur = sr * (1 - s) * u - s * sr * u + s * (1 - s) * ur)
int af_sigmoid_fwd(af_ctx* c, const double* X, double* S, int64_t len) { return af_sigmoid_fwd_r(c, X, nullptr, S, nullptr, len); }
int af_sigmoid_bwd(af_ctx* c, const double* S, double* U, int64_t len) { return af_sigmoid_bwd_r(c, S, nullptr, U, nullptr, len); }
// math_funcs.go:12-97
EW_FWD(af_exp_fwd_r, o = std::exp(x); ro = o * xr)
EW_BWD(af_exp_bwd_r, nu = u * s; nur = ur * s + u * sr)
int af_exp_fwd(af_ctx* c, const double* X, double* E, int64_t len) { return af_exp_fwd_r(c, X, nullptr, E, nullptr, len); }
int af_exp_bwd(af_ctx* c, const double* E, double* U, int64_t len) { return af_exp_bwd_r(c, E, nullptr, U, nullptr, len); }
// math_funcs.go:99-189
EW_FWD(af_log_fwd_r, o = std::log(x); ro = xr / x)
EW_BWD(af_log_bwd_r, nu = u / s; nur = ur / s - u * sr / (s * s))
// math_funcs.go:376-477
EW_FWD(af_logsigmoid_fwd_r, o = x < 0 ? x - std::log1p(std::exp(x)) : -std::log1p(std::exp(-x)); ro = xr / (1 + std::exp(x)))
EW_BWD(af_logsigmoid_bwd_r, const double e = std::exp(s); nu = u / (1 + e); nur = ur / (1 + e) - u * sr * e / ((1 + e) * (1 + e)))
// math_funcs.go:511-597
EW_FWD(af_sin_fwd_r, o = std::sin(x); ro = std::cos(x) * xr)
EW_BWD(af_sin_bwd_r, nu = u * std::cos(s); nur = ur * std::cos(s) - u * std::sin(s) * sr)
// extension (north star): tanh with Sigmoid's contract
EW_FWD(af_tanh_fwd_r, o = std::tanh(x); ro = (1 - o * o) * xr)
EW_BWD(af_tanh_bwd_r, nu = (1 - s * s) * u; nur = (1 - s * s) * ur - 2 * s * sr * u)
// arithmetic.go:696-770
EW_FWD(af_square_fwd_r, o = x * x; ro = 2 * x * xr)
EW_BWD(af_square_bwd_r, nu = 2 * s * u; nur = 2 * (ur * s + u * sr))

// ---- arithmetic.go ---------------------------------------------------------------------------------------------------
int af_add(af_ctx* c, const double* a, const double* b, double* o, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) o[i] = a[i] + b[i];
  return AF_OK;
}
int af_sub(af_ctx* c, const double* a, const double* b, double* o, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) o[i] = a[i] - b[i];
  return AF_OK;
}
int af_mul(af_ctx* c, const double* a, const double* b, double* o, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) o[i] = a[i] * b[i];
  return AF_OK;
}
int af_mul_r(af_ctx* c, const double* a, const double* aR, const double* b, const double* bR, double* o, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) o[i] = a[i] * (bR ? bR[i] : 0) + (aR ? aR[i] : 0) * b[i];  // arithmetic.go:325-329
  return AF_OK;
}
int af_mul_bwd(af_ctx* c, const double* U, const double* UR, const double* o, const double* oR, double* D, double* DR,
               int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) {  // arithmetic.go:291-293,358-362
    if (DR) DR[i] = U[i] * (oR ? oR[i] : 0) + (UR ? UR[i] : 0) * o[i];
    D[i] = U[i] * o[i];
  }
  return AF_OK;
}
int af_scale(af_ctx* c, double* x, double f, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) x[i] *= f;
  return AF_OK;
}
int af_axpy(af_ctx* c, double f, const double* x, double* y, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) y[i] += f * x[i];
  return AF_OK;
}
int af_add_scaler(af_ctx* c, const double* X, double f, double* O, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) O[i] = X[i] + f;
  return AF_OK;
}
int af_sum_all(af_ctx* c, const double* x, const double* xR, double* out, double* outR, int64_t n) {
  CTX(c);
  double s = 0, sr = 0;
  for (int64_t i = 0; i < n; i++) {
    s += x[i];
    if (xR) sr += xR[i];
  }
  *out = s;
  if (outR) *outR = sr;
  return AF_OK;
}
int af_fill_dev(af_ctx* c, double* x, const double* v, int64_t n) {
  CTX(c);
  const double f = *v;
  for (int64_t i = 0; i < n; i++) x[i] = f;
  return AF_OK;
}
int af_fill(af_ctx* c, double* x, double v, int64_t n) { return af_fill_dev(c, x, &v, n); }

// arithmetic.go:772-867.  The backward pass reads the forward OUTPUT o = 1/x and the input's R-vector.
EW_FWD(af_inverse_fwd_r, o = 1 / x; ro = -(o * o) * xr)
EW_BWD(af_inverse_bwd_r, const double sq = -(s * s); nu = sq * u; nur = sq * ur + -2 * (sq * s) * sr * u)
// arithmetic.go:869-966
int af_pow_fwd_r(af_ctx* c, const double* X, const double* RX, double p, double* O, double* RO, int64_t len) {
  CTX(c);
  for (int64_t i = 0; i < len; i++) {
    O[i] = std::pow(X[i], p);
    if (RO) RO[i] = p != 0 ? p * std::pow(X[i], p - 1) * (RX ? RX[i] : 0) : 0.0;  // :925-930
  }
  return AF_OK;
}
int af_pow_bwd_r(af_ctx* c, const double* X, const double* RX, double p, double* U, double* UR, int64_t len) {
  CTX(c);
  if (p == 1) return AF_OK;  // :898,952
  for (int64_t i = 0; i < len; i++) {
    const double u = U[i], d1 = p * std::pow(X[i], p - 1);
    if (UR) UR[i] = UR[i] * d1 + u * (p * (p - 1) * std::pow(X[i], p - 2) * (RX ? RX[i] : 0));  // :955-961
    U[i] = u * d1;
  }
  return AF_OK;
}
// arithmetic.go:378-423
int af_div(af_ctx* c, const double* A, const double* B, double* O, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) O[i] = A[i] / B[i];
  return AF_OK;
}
int af_div_bwd(af_ctx* c, const double* U, const double* O, const double* B, double* DA, double* DB, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) {
    if (DA) DA[i] = U[i] / B[i];              // :409-415
    if (DB) DB[i] = U[i] * (-O[i] / B[i]);    // :416-421
  }
  return AF_OK;
}
// arithmetic.go:595-612 (AddFirst) and the "- maxVal" / "+ maxVal" of :1056-1090
int af_add_dev(af_ctx* c, const double* X, const double* f, double sign, double* O, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) O[i] = X[i] + sign * f[0];
  return AF_OK;
}
// arithmetic.go:503-510,544-553 forward; :583-592 on the upstreams.  Outputs may alias inputs.
int af_scale_dev(af_ctx* c, const double* X, const double* RX, const double* f, const double* fR, double* O,
                 double* RO, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) {
    const double x = X[i], xr = RX ? RX[i] : 0;
    if (RO) RO[i] = x * (fR ? fR[0] : 0) + xr * f[0];
    O[i] = x * f[0];
  }
  return AF_OK;
}
// DotFast at arithmetic.go:523,571-573
int af_dot(af_ctx* c, const double* a, const double* b, const double* cc, const double* d, double* out, int64_t n) {
  CTX(c);
  if ((cc == nullptr) != (d == nullptr)) return fail(AF_EINVAL, "af_dot: c and d must be given together");
  double s = 0;
  for (int64_t i = 0; i < n; i++) s += a[i] * b[i] + (cc ? cc[i] * d[i] : 0);
  *out = s;
  return AF_OK;
}
// linalg.Vector.MaxAbs at arithmetic.go:1057,1078
int af_max_abs(af_ctx* c, const double* x, double* out, int combine, int64_t n) {
  CTX(c);
  double m = 0;
  for (int64_t i = 0; i < n; i++) m = std::fmax(m, std::fabs(x[i]));
  *out = combine ? std::fmax(*out, m) : m;
  return AF_OK;
}
// math_funcs.go:201-284
int af_norm_fwd_r(af_ctx* c, const double* X, const double* RX, double* out, double* outR, int64_t n) {
  CTX(c);
  double ss = 0, xr = 0;
  for (int64_t i = 0; i < n; i++) ss += X[i] * X[i], xr += X[i] * (RX ? RX[i] : 0);
  *out = std::sqrt(ss);
  if (outR) *outR = (1 / *out) * xr;  // :222-224
  return AF_OK;
}
int af_norm_bwd_r(af_ctx* c, const double* X, const double* RX, const double* o, const double* ro, const double* u,
                  const double* uR, double* D, double* DR, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) {
    const double s = u[0] / o[0];
    if (DR) DR[i] = X[i] * (-u[0] * ro[0] / (o[0] * o[0]) + uR[0] / o[0]) + (RX ? RX[i] : 0) * s;  // :267-284
    D[i] = X[i] * s;                                                                              // :238-247
  }
  return AF_OK;
}
// Scale(SquaredNorm(Sub(a, y)), .5): math_funcs.go:191-199, arithmetic.go:105-182,436-493,696-770,972-1047
int af_sqerr_fwd_r(af_ctx* c, const double* A, const double* RA, const double* Y, const double* RY, double* out,
                   double* outR, int64_t n) {
  CTX(c);
  double s = 0, sr = 0;
  for (int64_t i = 0; i < n; i++) {
    const double d = A[i] - Y[i];
    s += d * d;
    sr += 2 * d * ((RA ? RA[i] : 0) - (RY ? RY[i] : 0));
  }
  *out = 0.5 * s;
  if (outR) *outR = 0.5 * sr;
  return AF_OK;
}
int af_sqerr_bwd_r(af_ctx* c, const double* A, const double* RA, const double* Y, const double* RY, const double* u,
                   const double* uR, double sign, double* G, double* GR, int64_t n) {
  CTX(c);
  for (int64_t i = 0; i < n; i++) {
    const double d = A[i] - Y[i], dr = (RA ? RA[i] : 0) - (RY ? RY[i] : 0), hu = 0.5 * u[0];
    if (GR) GR[i] = sign * (2 * ((0.5 * uR[0]) * d + hu * dr));
    G[i] = sign * (hu * (2 * d));
  }
  return AF_OK;
}

// ---- Softmax rows (math_funcs.go:479-509; no max-subtraction, like the reference) -------------------------------
int af_softmax_fwd_r(af_ctx* c, const double* X, const double* RX, double t, double* Y, double* RY, int64_t len,
                     int64_t n) {
  CTX(c);
  const double it = (t != 0 && t != 1) ? 1 / t : 1.0;
  for (int64_t r = 0; r < n; r++) {
    const double* x = X + r * len;
    double z = 0, zr = 0;
    for (int64_t i = 0; i < len; i++) {
      const double e = std::exp(x[i] * it);
      z += e;
      if (RX) zr += e * RX[r * len + i] * it;
    }
    for (int64_t i = 0; i < len; i++) {
      const double y = std::exp(x[i] * it) / z;
      Y[r * len + i] = y;
      if (RY) RY[r * len + i] = y * ((RX ? RX[r * len + i] * it : 0) - zr / z);
    }
  }
  return AF_OK;
}
int af_softmax_bwd_r(af_ctx* c, const double* Y, const double* RY, double t, double* U, double* UR, int64_t len,
                     int64_t n) {
  CTX(c);
  const double it = (t != 0 && t != 1) ? 1 / t : 1.0;
  for (int64_t r = 0; r < n; r++) {
    const double *y = Y + r * len, *yr = RY ? RY + r * len : nullptr;
    double *u = U + r * len, *ur = UR ? UR + r * len : nullptr;
    double uy = 0, ury = 0, uyr = 0;
    for (int64_t i = 0; i < len; i++) {
      uy += u[i] * y[i];
      if (ur) ury += ur[i] * y[i], uyr += u[i] * (yr ? yr[i] : 0);
    }
    for (int64_t i = 0; i < len; i++) {
      if (ur) ur[i] = ((yr ? yr[i] : 0) * (u[i] - uy) + y[i] * (ur[i] - ury - uyr)) * it;
      u[i] = y[i] * (u[i] - uy) * it;
    }
  }
  return AF_OK;
}

// ---- fused dense layer (the same algebra composed) ------------------------------------------------------------------
int af_dense_fwd(af_ctx* c, const double* W, const double* RW, const double* b, const double* Rb, const double* X,
                 const double* RX, double* S, double* RS, int64_t rows, int64_t cols, int64_t n, int activation) {
  int rc = af_lintran_fwd_r(c, W, RW, X, RX, S, RS, rows, cols, n);
  if (rc) return rc;
  for (int64_t i = 0; i < rows * n; i++) {
    double pre = S[i] + (b ? b[i % rows] : 0), rpre = RS ? RS[i] + (Rb ? Rb[i % rows] : 0) : 0;  // lin_add.go:37-42
    double s = activation ? sigmoid(pre) : pre;
    S[i] = s;
    if (RS) RS[i] = activation ? s * (1 - s) * rpre : rpre;  // math_funcs.go:305-309
  }
  return AF_OK;
}
int af_dense_igrad(af_ctx* c, const double* U, const double* UR, const double* W, const double* RW,
                   const double* Sprev, const double* RSprev, double* D, double* DR, int64_t rows, int64_t cols,
                   int64_t n) {
  int rc = af_lintran_igrad_r(c, U, UR, W, RW, D, DR, rows, cols, n);
  if (rc || !Sprev) return rc;
  return af_sigmoid_bwd_r(c, Sprev, RSprev, D, DR, cols * n);
}

// ---- whole-network plan (bench/mlp.go:24-126), host-buffer form only --------------------------------------------------
int af_mlp_create(af_ctx* c, const int64_t* sizes, int n_layers, int64_t batch, af_mlp** out) {
  if (!c || !sizes || !out) return fail(AF_EINVAL, "null argument");
  if (n_layers < 1 || batch < 1) return fail(AF_ESHAPE, "MLP needs at least one layer and one sample");
  af_mlp* m = new af_mlp();
  m->ctx = c;
  m->sizes.assign(sizes, sizes + n_layers + 1);
  m->n = batch;
  const int P = 2 * n_layers;
  m->param.resize(P), m->rvec.resize(P), m->grad.resize(P), m->rgrad.resize(P), m->has_r.assign(P, false);
  for (int i = 0; i < P; i++) {
    m->param[i].assign((size_t)m->plen(i), 0.0), m->rvec[i].assign((size_t)m->plen(i), 0.0);
    m->grad[i].assign((size_t)m->plen(i), 0.0), m->rgrad[i].assign((size_t)m->plen(i), 0.0);
  }
  m->act.resize(n_layers + 1), m->ract.resize(n_layers + 1);
  for (int l = 0; l <= n_layers; l++) m->act[l].assign((size_t)(batch * sizes[l]), 0.0), m->ract[l] = m->act[l];
  m->up.assign((size_t)(batch * sizes[n_layers]), 0.0), m->upr = m->up;
  *out = m;
  return AF_OK;
}
int af_mlp_destroy(af_mlp* m) {
  delete m;
  return AF_OK;
}
int af_mlp_set_param(af_mlp* m, int i, const double* h, int64_t len) {
  if (!m || !h) return fail(AF_EINVAL, "null argument");
  if (i < 0 || i >= 2 * m->L() || len != m->plen(i)) return fail(AF_ESHAPE, "parameter length does not match");
  m->param[i].assign(h, h + len);
  return AF_OK;
}
int af_mlp_set_rvector(af_mlp* m, int i, const double* h, int64_t len) {
  if (!m) return fail(AF_EINVAL, "null argument");
  if (i < 0 || i >= 2 * m->L()) return fail(AF_ESHAPE, "parameter index out of range");
  m->has_r[i] = h != nullptr;
  if (!h) return AF_OK;
  if (len != m->plen(i)) return fail(AF_ESHAPE, "R-vector length does not match");
  m->rvec[i].assign(h, h + len);
  return AF_OK;
}
int af_mlp_step_host(af_mlp* m, int with_r, const double* h_X, const double* h_up, const double* h_upr, double* h_out,
                     double* h_rout, double* const* h_grads, double* const* h_rgrads) {
  if (!m || !h_X) return fail(AF_EINVAL, "null argument");
  af_ctx* c = m->ctx;
  const int L = m->L();
  const int64_t n = m->n;
  m->act[0].assign(h_X, h_X + n * m->sizes[0]);
  for (int i = 0; i < 2 * L; i++) {  // fresh Gradient maps (bench/mlp.go:33-35,47-49)
    std::fill(m->grad[i].begin(), m->grad[i].end(), 0.0);
    std::fill(m->rgrad[i].begin(), m->rgrad[i].end(), 0.0);
  }
  auto R = [&](int i) -> const double* { return with_r && m->has_r[i] ? m->rvec[i].data() : nullptr; };
  for (int l = 0; l < L; l++)
    af_dense_fwd(c, m->param[2 * l].data(), R(2 * l), m->param[2 * l + 1].data(), R(2 * l + 1), m->act[l].data(),
                 (with_r && l > 0) ? m->ract[l].data() : nullptr, m->act[l + 1].data(),
                 with_r ? m->ract[l + 1].data() : nullptr, m->sizes[l + 1], m->sizes[l], n, 1);
  const int64_t nout = n * m->sizes[L];
  if (h_out) std::memcpy(h_out, m->act[L].data(), (size_t)nout * 8);
  if (h_rout && with_r) std::memcpy(h_rout, m->ract[L].data(), (size_t)nout * 8);
  for (int64_t i = 0; i < nout; i++) {  // bench/mlp.go:118-126
    m->up[i] = h_up ? h_up[i] : 1.0;
    m->upr[i] = h_upr ? h_upr[i] : 0.0;
  }
  std::vector<double> u = m->up, ur = m->upr;
  for (int l = L - 1; l >= 0; l--) {
    const int64_t rows = m->sizes[l + 1], cols = m->sizes[l];
    af_sigmoid_bwd_r(c, m->act[l + 1].data(), with_r ? m->ract[l + 1].data() : nullptr, u.data(),
                     with_r ? ur.data() : nullptr, rows * n);
    af_bias_grad_r(c, u.data(), ur.data(), m->grad[2 * l + 1].data(), with_r ? m->rgrad[2 * l + 1].data() : nullptr,
                   rows, n);
    af_lintran_wgrad_r(c, u.data(), ur.data(), m->act[l].data(), (with_r && l > 0) ? m->ract[l].data() : nullptr,
                       m->grad[2 * l].data(), with_r ? m->rgrad[2 * l].data() : nullptr, rows, cols, n);
    if (l > 0) {
      std::vector<double> d((size_t)(cols * n)), dr((size_t)(cols * n));
      af_lintran_igrad_r(c, u.data(), ur.data(), m->param[2 * l].data(), R(2 * l), d.data(),
                         with_r ? dr.data() : nullptr, rows, cols, n);
      u.swap(d), ur.swap(dr);
    }
  }
  for (int i = 0; i < 2 * L; i++) {
    if (h_grads && h_grads[i]) std::memcpy(h_grads[i], m->grad[i].data(), m->grad[i].size() * 8);
    if (with_r && h_rgrads && h_rgrads[i]) std::memcpy(h_rgrads[i], m->rgrad[i].data(), m->rgrad[i].size() * 8);
  }
  return AF_OK;
}
int af_mlp_step_fresh(af_mlp* m, int with_r) {
  if (!m) return fail(AF_EINVAL, "null argument");
  std::vector<double> x = m->act[0];
  return af_mlp_step_host(m, with_r, x.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}
int af_mlp_add_to_vars(af_mlp* m, double scale) {  // gradient.go:40-52
  if (!m) return fail(AF_EINVAL, "null argument");
  for (size_t i = 0; i < m->param.size(); i++)
    for (size_t k = 0; k < m->param[i].size(); k++) m->param[i][k] += scale * m->grad[i][k];
  return AF_OK;
}
double* af_mlp_input_ptr(af_mlp* m) { return m->act[0].data(); }
double* af_mlp_output_ptr(af_mlp* m) { return m->act.back().data(); }
double* af_mlp_routput_ptr(af_mlp* m) { return m->ract.back().data(); }
double* af_mlp_grad_ptr(af_mlp* m, int i) { return m->grad[i].data(); }
double* af_mlp_rgrad_ptr(af_mlp* m, int i) { return m->rgrad[i].data(); }

// ---- recorded launch sequences: a CUDA-graph facility of the product, nothing to restate on a CPU -----------------
int af_graph_begin(af_ctx*) { return fail(AF_ENOTSUP, "the CPU restatement of the ABI records no graphs"); }
int af_graph_end(af_ctx*, af_graph**) { return fail(AF_ENOTSUP, "no graphs"); }
int af_graph_launch(af_graph*) { return fail(AF_ENOTSUP, "no graphs"); }
int64_t af_graph_launches(af_graph*) { return 0; }
int af_graph_destroy(af_graph*) { return AF_OK; }

// ---- LSTM plan: not restated here (oracle/autofunc_ref.py LSTMNet is the checker for it) ---------------------------
int af_lstm_create(af_ctx*, int64_t, int64_t, int64_t, int64_t, int64_t, af_lstm**) {
  return fail(AF_ENOTSUP, "the CPU restatement of the ABI has no LSTM plan");
}
int af_lstm_destroy(af_lstm*) { return AF_OK; }
int af_lstm_set_param(af_lstm*, int, const double*, int64_t) { return fail(AF_ENOTSUP, "no LSTM plan"); }
int af_lstm_set_rvector(af_lstm*, int, const double*, int64_t) { return fail(AF_ENOTSUP, "no LSTM plan"); }
int af_lstm_run_host(af_lstm*, int, const double*, const double*, const double*, double*, double*, double* const*,
                     double* const*) {
  return fail(AF_ENOTSUP, "no LSTM plan");
}

}  // extern "C"
