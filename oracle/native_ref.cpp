// TEST / MEASUREMENT INFRASTRUCTURE -- never linked into or called by the product (autofunc_b200/).
//
// "oracle-native" of SURVEY 8(d) / BASELINE.md 3(1): the batched MLP step of bench/mlp.go:43-57 restated in C++ with the
// reference's own call structure -- every lin_tran.go BLAS call site is ONE separate Gemm call (lin_tran.go:113,
// 172-173, 210, 267-268, 302, 354-355), bias and sigmoid are separate passes (lin_add.go:37-42, math_funcs.go:305-309,
// 365-371) -- and a Gemm that is cache-blocked and parallel over tiles of C, which is what gonum's pure-Go native Dgemm
// does with goroutines (github.com/gonum/blas/native, unvendored and unpinned in the reference; recalled, not
// verifiable here).  It is the stand-in for `go test -bench . ./bench` (the reference's DEFAULT BLAS backend) next to
// the numpy/OpenBLAS oracle, which stands in for ./bench_cblas.  Compiled without -march flags: gonum's assembly
// kernels are SSE2, and the GPU box's host CPU may differ from the build container's.
//
// Checked against oracle/autofunc_ref.py at 1e-10 by tests/test_oracle.py; timed by bench.py --impl reference.
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// C (m x n, ldc) = beta * C + op(A) * op(B); row-major; ta / tb = transposed operand, as blas64.Gemm(tA, tB, 1, A, B, beta, C)
void gemm(bool ta, bool tb, int64_t m, int64_t n, int64_t k, const double* A, int64_t lda, const double* B, int64_t ldb,
          double beta, double* C, int64_t ldc) {
  constexpr int64_t BM = 64, BN = 64, BK = 256;
  const int64_t tm = (m + BM - 1) / BM, tn = (n + BN - 1) / BN;
#pragma omp parallel for schedule(dynamic) collapse(2)
  for (int64_t bi = 0; bi < tm; ++bi) {
    for (int64_t bj = 0; bj < tn; ++bj) {
      const int64_t i0 = bi * BM, i1 = std::min(m, i0 + BM), j0 = bj * BN, j1 = std::min(n, j0 + BN);
      double acc[BM][BN];
      for (int64_t i = i0; i < i1; ++i)
        for (int64_t j = j0; j < j1; ++j) acc[i - i0][j - j0] = beta == 0.0 ? 0.0 : beta * C[i * ldc + j];
      double pa[BM][BK], pb[BK][BN];  // packed panels: pa[i][p] = op(A)[i0+i][k0+p], pb[p][j] = op(B)[k0+p][j0+j]
      for (int64_t k0 = 0; k0 < k; k0 += BK) {
        const int64_t k1 = std::min(k, k0 + BK);
        for (int64_t i = i0; i < i1; ++i)
          for (int64_t p = k0; p < k1; ++p) pa[i - i0][p - k0] = ta ? A[p * lda + i] : A[i * lda + p];
        for (int64_t p = k0; p < k1; ++p)
          for (int64_t j = j0; j < j1; ++j) pb[p - k0][j - j0] = tb ? B[j * ldb + p] : B[p * ldb + j];
        for (int64_t i = 0; i < i1 - i0; ++i)
          for (int64_t p = 0; p < k1 - k0; ++p) {  // axpy form: the inner loop is unit-stride in both operands
            const double a = pa[i][p];
            const double* __restrict__ brow = pb[p];
            double* __restrict__ crow = acc[i];
            for (int64_t j = 0; j < j1 - j0; ++j) crow[j] += a * brow[j];
          }
      }
      for (int64_t i = i0; i < i1; ++i)
        for (int64_t j = j0; j < j1; ++j) C[i * ldc + j] = acc[i - i0][j - j0];
    }
  }
}

}  // namespace

extern "C" {

int native_ref_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// torchrun exports OMP_NUM_THREADS=1 to its ranks: the CPU leg sizes the pool explicitly (all host cores)
void native_ref_set_threads(int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
}

// One iteration of bench/mlp.go:43-57 on a packed batch of n samples: ApplyR through {LinTran, LinAdd, Sigmoid} x L, then
// PropagateRGradient with outGrad = 1, outRGrad = 0 (bench/mlp.go:118-126) and non-nil grad, into ZEROED accumulators.
// params / rvecs / grads / rgrads: 2L pointers in the order W0, b0, W1, b1, ... (bench/mlp.go:84-85); W_l is
// sizes[l+1] x sizes[l] row-major.  The input is a constant (not in the RVector): RX of layer 1 is a zero vector that is
// still multiplied, as the reference does (variable.go:85-91, lin_tran.go:172-173).
int native_mlp_step_r(const int64_t* sizes, int L, int64_t n, const double* const* params, const double* const* rvecs,
                      const double* x, double* out, double* rout, double* const* grads, double* const* rgrads) {
  std::vector<std::vector<double>> act(L + 1), ract(L + 1);
  act[0].assign(x, x + n * sizes[0]);
  ract[0].assign((size_t)(n * sizes[0]), 0.0);
  for (int l = 0; l < L; ++l) {
    const int64_t K = sizes[l], M = sizes[l + 1];
    const double *W = params[2 * l], *RW = rvecs[2 * l], *b = params[2 * l + 1], *Rb = rvecs[2 * l + 1];
    std::vector<double>&y = act[l + 1], &ry = ract[l + 1];
    y.assign((size_t)(n * M), 0.0);
    ry.assign((size_t)(n * M), 0.0);
    gemm(false, true, n, M, K, act[l].data(), K, W, K, 0.0, y.data(), M);    // lin_tran.go:113  Y = X W^T
    gemm(false, true, n, M, K, act[l].data(), K, RW, K, 0.0, ry.data(), M);  // :172  RY = X RW^T
    gemm(false, true, n, M, K, ract[l].data(), K, W, K, 1.0, ry.data(), M);  // :173  RY += RX W^T
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i)
      for (int64_t j = 0; j < M; ++j) {
        const double z = y[i * M + j] + b[j], rz = ry[i * M + j] + Rb[j];  // lin_add.go:37-42
        const double s = 1.0 / (1.0 + std::exp(-z));                       // math_funcs.go:305-309
        y[i * M + j] = s;
        ry[i * M + j] = s * (1 - s) * rz;
      }
  }
  const int64_t top = n * sizes[L];
  std::memcpy(out, act[L].data(), (size_t)top * 8);
  std::memcpy(rout, ract[L].data(), (size_t)top * 8);
  std::vector<double> u((size_t)top, 1.0), ur((size_t)top, 0.0);  // bench/mlp.go:118-126
  for (int l = L - 1; l >= 0; --l) {
    const int64_t K = sizes[l], M = sizes[l + 1];
    const double *W = params[2 * l], *RW = rvecs[2 * l];
    const std::vector<double>&s = act[l + 1], &rs = ract[l + 1];
#pragma omp parallel for
    for (int64_t e = 0; e < n * M; ++e) {  // math_funcs.go:365-371, operand order kept
      const double o = s[e], orr = rs[e], uu = u[e];
      ur[e] = orr * (1 - o) * uu - o * orr * uu + o * (1 - o) * ur[e];
      u[e] = o * (1 - o) * uu;
    }
    double *gb = grads[2 * l + 1], *rgb = rgrads[2 * l + 1];  // lin_add.go:94-104 per sample == column sums
    for (int64_t i = 0; i < n; ++i)
      for (int64_t j = 0; j < M; ++j) { gb[j] += u[i * M + j]; rgb[j] += ur[i * M + j]; }
    gemm(true, false, M, K, n, u.data(), M, act[l].data(), K, 1.0, grads[2 * l], K);     // :210  gW += U^T X
    gemm(true, false, M, K, n, u.data(), M, ract[l].data(), K, 1.0, rgrads[2 * l], K);   // :267  rgW += U^T RX
    gemm(true, false, M, K, n, ur.data(), M, act[l].data(), K, 1.0, rgrads[2 * l], K);   // :268  rgW += UR^T X
    if (l > 0) {  // the input of layer 1 is constant: lin_tran.go:420 prunes its input gradient
      std::vector<double> d((size_t)(n * K), 0.0), dr((size_t)(n * K), 0.0);
      gemm(false, false, n, K, M, u.data(), M, W, K, 0.0, d.data(), K);     // :302  D = U W
      gemm(false, false, n, K, M, ur.data(), M, W, K, 0.0, dr.data(), K);   // :354  DR = UR W
      gemm(false, false, n, K, M, u.data(), M, RW, K, 1.0, dr.data(), K);   // :355  DR += U RW
      u.swap(d);
      ur.swap(dr);
    }
  }
  return 0;
}

}  // extern "C"
