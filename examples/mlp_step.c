/* Plain C99 client of the C ABI (include/autofunc_b200.h): one bench/mlp.go iteration -- forward, Gradient and
 * RGradient of a small sigmoid MLP -- from HOST buffers, the call a cgo / JNI / ctypes binding makes.
 *
 * It checks itself without any other implementation, through properties the reference defines:
 *   - gradients accumulate: the second resident step doubles every accumulator (variable.go:45)
 *   - the R-operator is linear in the R-vector: scaling every R-vector by 2 doubles ROutput and RGradient
 *   - Gradient.AddToVars(0) leaves the parameters unchanged and AddToVars(s) moves them by s * grad
 *
 *   gcc -std=c99 -Iinclude examples/mlp_step.c -Lautofunc_b200 -lautofunc_b200 -Wl,-rpath,$PWD/autofunc_b200 -lm -o mlp_step
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "autofunc_b200.h"

#define CHECK(expr)                                                               \
  do {                                                                            \
    int rc_ = (expr);                                                             \
    if (rc_ != AF_OK) {                                                           \
      fprintf(stderr, "%s failed (%d): %s\n", #expr, rc_, af_last_error());       \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

static double unif(unsigned long long* s) { /* splitmix64 -> U[-1, 1) */
  unsigned long long z = (*s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  z ^= z >> 31;
  return (double)(z >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

static int close_to(const double* a, const double* b, double scale, long n, const char* what) {
  for (long i = 0; i < n; ++i) {
    double want = scale * b[i];
    if (fabs(a[i] - want) > 1e-12 + 1e-10 * fabs(want)) {
      fprintf(stderr, "%s: entry %ld is %.17g, expected %.17g\n", what, i, a[i], want);
      return 0;
    }
  }
  return 1;
}

int main(void) {
  enum { L = 3, P = 2 * L };
  const int64_t sizes[L + 1] = {64, 96, 32, 10};
  const int64_t n = 48;
  af_ctx* ctx = NULL;
  af_mlp* net = NULL;
  CHECK(af_ctx_create(0, &ctx));
  CHECK(af_mlp_create(ctx, sizes, L, n, &net));

  unsigned long long seed = 123;
  long plen[P];
  double *param[P], *rvec[P], *g1[P], *rg1[P], *g2[P], *rg2[P];
  for (int i = 0; i < P; ++i) {
    int l = i / 2;
    plen[i] = (i & 1) ? sizes[l + 1] : sizes[l] * sizes[l + 1];
    param[i] = malloc(sizeof(double) * plen[i]);
    rvec[i] = malloc(sizeof(double) * plen[i]);
    g1[i] = malloc(sizeof(double) * plen[i]); rg1[i] = malloc(sizeof(double) * plen[i]);
    g2[i] = malloc(sizeof(double) * plen[i]); rg2[i] = malloc(sizeof(double) * plen[i]);
    for (long j = 0; j < plen[i]; ++j) {
      param[i][j] = unif(&seed) / sqrt((double)sizes[l]);
      rvec[i][j] = unif(&seed);
    }
    CHECK(af_mlp_set_param(net, i, param[i], plen[i]));
    CHECK(af_mlp_set_rvector(net, i, rvec[i], plen[i]));
  }
  const long nin = n * sizes[0], nout = n * sizes[L];
  double* x = malloc(sizeof(double) * nin);
  double *out1 = malloc(sizeof(double) * nout), *rout1 = malloc(sizeof(double) * nout);
  double *out2 = malloc(sizeof(double) * nout), *rout2 = malloc(sizeof(double) * nout);
  for (long j = 0; j < nin; ++j) x[j] = unif(&seed);

  /* one iteration from host buffers: upstream NULL = outGrad 1, outRGrad 0 (bench/mlp.go:118-126) */
  CHECK(af_mlp_step_host(net, 1, x, NULL, NULL, out1, rout1, g1, rg1));

  /* gradients accumulate: a second resident step (upstreams refilled) doubles the accumulators */
  CHECK(af_fill(ctx, af_mlp_upstream_ptr(net), 1.0, nout));
  CHECK(af_zero(ctx, af_mlp_upstream_r_ptr(net), nout));
  CHECK(af_mlp_step(net, 1, 1));
  for (int i = 0; i < P; ++i) {
    CHECK(af_download(ctx, g2[i], af_mlp_grad_ptr(net, i), plen[i]));
    CHECK(af_download(ctx, rg2[i], af_mlp_rgrad_ptr(net, i), plen[i]));
    if (!close_to(g2[i], g1[i], 2.0, plen[i], "accumulated grad") || !close_to(rg2[i], rg1[i], 2.0, plen[i], "accumulated rgrad"))
      return 1;
  }

  /* the R-operator is linear in the R-vector */
  for (int i = 0; i < P; ++i) {
    for (long j = 0; j < plen[i]; ++j) rvec[i][j] *= 2.0;
    CHECK(af_mlp_set_rvector(net, i, rvec[i], plen[i]));
  }
  CHECK(af_mlp_step_host(net, 1, x, NULL, NULL, out2, rout2, g2, rg2));
  if (!close_to(out2, out1, 1.0, nout, "Output") || !close_to(rout2, rout1, 2.0, nout, "ROutput")) return 1;
  for (int i = 0; i < P; ++i)
    if (!close_to(g2[i], g1[i], 1.0, plen[i], "grad") || !close_to(rg2[i], rg1[i], 2.0, plen[i], "rgrad")) return 1;

  /* Gradient.AddToVars on the device (gradient.go:40-52) */
  CHECK(af_mlp_add_to_vars(net, 0.0));
  CHECK(af_mlp_get_param(net, 0, g2[0], plen[0]));
  if (!close_to(g2[0], param[0], 1.0, plen[0], "AddToVars(0)")) return 1;
  CHECK(af_mlp_add_to_vars(net, -0.5));
  CHECK(af_mlp_get_param(net, 0, g2[0], plen[0]));
  for (long j = 0; j < plen[0]; ++j) rg2[0][j] = param[0][j] - 0.5 * g1[0][j];
  if (!close_to(g2[0], rg2[0], 1.0, plen[0], "AddToVars(-0.5)")) return 1;

  printf("mlp_step ok: %lld kernel launches, output[0] = %.12f, routput[0] = %.12f\n",
         (long long)af_launch_count(ctx), out1[0], rout1[0]);
  CHECK(af_mlp_destroy(net));
  CHECK(af_ctx_destroy(ctx));
  return 0;
}
