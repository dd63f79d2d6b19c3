// Small-batch LinTran kernels: the reference's own n == 1 branch (Gemv / Ger instead of Gemm,
// lin_tran.go:90-99, 138-152, 187-196, 223-241, 285-288, 329-334), generalised to n <= 8 samples (n <= 16 for the
// weight gradient; the other passes of 9..16 take the tensor-core kernel with transposed narrow tiles, which
// measured faster there).
//
// With a handful of samples the six contractions are HBM-bound streaming passes over the weight
// matrix, not tensor-core work: the algorithmic traffic is W, RW read once per forward and once per
// input-gradient pass, and gW, rgW read-modify-written once (8 passes of 8*M*K bytes per
// R-step).  Each kernel below makes exactly those passes with 16-byte coalesced accesses, keeps the
// n samples' partial sums in registers, and applies the fused elementwise step where one thread block
// owns a complete output (forward); the input gradient is split over row chunks and combined with
// red.global.add.f64, its elementwise step applied afterwards by a tiny pass.
#include <atomic>
#include "common.h"
#include "gemm_dmma.cuh"

namespace af {
namespace {

constexpr int kSmallThreads = 256;

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// ---- forward: `wpr` warps per weight row --------------------------------------------------------
// Y[i][m] = sum_k X[i][k] W[m][k] ;  RY[i][m] = sum_k X[i][k] RW[m][k] + RX[i][k] W[m][k]
// (Gemv N at lin_tran.go:99,151-152), then + bias, sigmoid and the R form in the same block.
// A block of 8 warps owns 8/wpr rows; the warps of a row split K, reduce by shuffle, then through
// shared memory, so that even a 512-row layer puts >= 256 blocks on the chip.
template <int NMAX>
__global__ void __launch_bounds__(kSmallThreads) small_fwd_kernel(const double* __restrict__ W, const double* __restrict__ RW,
                                                                  const double* __restrict__ X, const double* __restrict__ RX,
                                                                  int rows, int cols, int n, int wpr, const GemmArgs p) {
  __shared__ double part[8][2 * NMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_per_block = 8 / wpr;
  const int m = blockIdx.x * rows_per_block + warp / wpr;
  const int ws = warp % wpr;
  const bool live = m < rows;
  double y[NMAX], ry[NMAX];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) y[i] = ry[i] = 0.0;
  if (live) {
    const double* w = W + (long long)m * cols;
    const double* rw = RW ? RW + (long long)m * cols : nullptr;
#pragma unroll 2
    for (int k = 2 * lane + 64 * ws; k < cols; k += 64 * wpr) {
      const double2 wv = ld2(w + k);
      double2 rwv = make_double2(0.0, 0.0);
      if (rw) rwv = ld2(rw + k);
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
        if (i < n) {
          const double2 xv = ld2(X + (long long)i * cols + k);
          y[i] = fma(xv.x, wv.x, fma(xv.y, wv.y, y[i]));
          if (rw) ry[i] = fma(xv.x, rwv.x, fma(xv.y, rwv.y, ry[i]));
          if (RX) {
            const double2 rxv = ld2(RX + (long long)i * cols + k);
            ry[i] = fma(rxv.x, wv.x, fma(rxv.y, wv.y, ry[i]));
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      y[i] += __shfl_xor_sync(0xffffffffu, y[i], o);
      ry[i] += __shfl_xor_sync(0xffffffffu, ry[i], o);
    }
    if (lane == 0) { part[warp][2 * i] = y[i]; part[warp][2 * i + 1] = ry[i]; }
  }
  __syncthreads();
  // the first warp of each row group folds the group's partial sums; lane i finishes sample i
  if (live && ws == 0 && lane < n) {
    double a = 0.0, b = 0.0;
    for (int q = 0; q < wpr; ++q) { a += part[warp + q][2 * lane]; b += part[warp + q][2 * lane + 1]; }
    epi_elem<true>(p, lane, m, a, b);
  }
}

// ---- weight gradient: rank-n update ------------------------------------------------------------
// gW[m][k] += sum_i U[i][m] X[i][k] ; rgW[m][k] += sum_i U[i][m] RX[i][k] + UR[i][m] X[i][k]
// (Ger at lin_tran.go:196,240-241).  One pass: gW and rgW are read and written once.
template <int NMAX>
__global__ void __launch_bounds__(kSmallThreads) small_wgrad_kernel(const double* __restrict__ U, const double* __restrict__ UR,
                                                                    const double* __restrict__ X, const double* __restrict__ RX,
                                                                    double* __restrict__ gW, double* __restrict__ rgW, int rows,
                                                                    int cols, int n) {
  const int m = blockIdx.y;
  const int k = 2 * (blockIdx.x * kSmallThreads + threadIdx.x);
  if (k >= cols) return;
  double2 g = make_double2(0.0, 0.0), rg = make_double2(0.0, 0.0);
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    if (i < n) {
      const double u = U[(long long)i * rows + m];
      const double2 xv = ld2(X + (long long)i * cols + k);
      g.x = fma(u, xv.x, g.x); g.y = fma(u, xv.y, g.y);
      if (rgW) {
        if (RX) {
          const double2 rxv = ld2(RX + (long long)i * cols + k);
          rg.x = fma(u, rxv.x, rg.x); rg.y = fma(u, rxv.y, rg.y);
        }
        if (UR) {
          const double ur = UR[(long long)i * rows + m];
          rg.x = fma(ur, xv.x, rg.x); rg.y = fma(ur, xv.y, rg.y);
        }
      }
    }
  }
  const long long o = (long long)m * cols + k;
  if (gW) {
    double2 c = ld2(gW + o);
    c.x += g.x; c.y += g.y;
    *reinterpret_cast<double2*>(gW + o) = c;
  }
  if (rgW) {
    double2 c = ld2(rgW + o);
    c.x += rg.x; c.y += rg.y;
    *reinterpret_cast<double2*>(rgW + o) = c;
  }
}

// ---- input gradient: column-wise reduction over a chunk of weight rows --------------------------
// D[i][k] = sum_m U[i][m] W[m][k] ; DR[i][k] = sum_m UR[i][m] W[m][k] + U[i][m] RW[m][k]
// (Gemv T at lin_tran.go:288,333-334).  Thread = column pair, block row = chunk of `mchunk` rows;
// partial sums go to the zeroed outputs with red.global.add.f64.
template <int NMAX>
__global__ void __launch_bounds__(kSmallThreads) small_igrad_kernel(const double* __restrict__ U, const double* __restrict__ UR,
                                                                    const double* __restrict__ W, const double* __restrict__ RW,
                                                                    double* __restrict__ D, double* __restrict__ DR, int rows,
                                                                    int cols, int n, int mchunk) {
  extern __shared__ double su[];  // U and UR for this row chunk: [2][NMAX][mchunk]
  const int m0 = blockIdx.y * mchunk;
  const int mc = min(mchunk, rows - m0);
  for (int e = threadIdx.x; e < NMAX * mchunk; e += kSmallThreads) {
    const int i = e / mchunk, mm = e - i * mchunk;
    const bool ok = i < n && mm < mc;
    su[e] = ok ? U[(long long)i * rows + m0 + mm] : 0.0;
    su[NMAX * mchunk + e] = (ok && UR) ? UR[(long long)i * rows + m0 + mm] : 0.0;
  }
  __syncthreads();
  const int k = 2 * (blockIdx.x * kSmallThreads + threadIdx.x);
  if (k >= cols) return;
  double2 d[NMAX], dr[NMAX];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) d[i] = dr[i] = make_double2(0.0, 0.0);
  for (int mm = 0; mm < mc; ++mm) {
    const long long wo = (long long)(m0 + mm) * cols + k;
    const double2 wv = ld2(W + wo);
    double2 rwv = make_double2(0.0, 0.0);
    if (RW) rwv = ld2(RW + wo);
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
      if (i < n) {
        const double u = su[i * mchunk + mm];
        d[i].x = fma(u, wv.x, d[i].x); d[i].y = fma(u, wv.y, d[i].y);
        if (DR) {
          const double ur = su[NMAX * mchunk + i * mchunk + mm];
          dr[i].x = fma(ur, wv.x, fma(u, rwv.x, dr[i].x));
          dr[i].y = fma(ur, wv.y, fma(u, rwv.y, dr[i].y));
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    if (i < n) {
      const long long o = (long long)i * cols + k;
      if (D) { atomicAdd(D + o, d[i].x); atomicAdd(D + o + 1, d[i].y); }
      if (DR) { atomicAdd(DR + o, dr[i].x); atomicAdd(DR + o + 1, dr[i].y); }
    }
  }
}

// ---- input gradient through a layer with FEW OUTPUT ROWS (the 10-class head of the bench network) ----------
// D[i][k] = sum_{m < M} U[i][m] W[m][k] with M <= 16: a contraction of length M is no tensor-core work (one k-tile,
// 11 % pipe utilisation measured), it is a streaming pass over the n x K outputs.  Thread = (sample, column pair);
// W / RW (M x K, a few tens of KB) come from L1, the epilogue (sigmoid backward of the layer below, or plain store)
// is applied in registers, so HBM sees S, RS once and D, DR once.
constexpr int kFewRowsSpt = 4;  // samples per thread: every W / RW value read from shared memory is reused for four samples

// W / RW are staged once per (persistent) block in shared memory and the block walks the batch in groups of four samples,
// whose upstream rows U / UR (4 x M values) go through shared memory too: a thread's inner loop touches no global memory
// (first version: W / RW through L1 at 50 % hit rate plus 80 broadcast loads of U per thread -- ncu: long_scoreboard,
// 35 us for 67 MB).
template <bool DUAL>
__global__ void __launch_bounds__(256, 2) few_rows_igrad_kernel(const double* __restrict__ U, const double* __restrict__ UR,
                                                                const double* __restrict__ W, const double* __restrict__ RW,
                                                                int rows, int cols, long long n, const GemmArgs p) {
  extern __shared__ __align__(16) double few_smem[];
  double* ws = few_smem;                           // [rows][cols]
  double* rws = ws + (size_t)rows * cols;          // [rows][cols]
  double* us = rws + (size_t)rows * cols;          // [kFewRowsSpt][16]
  double* urs = us + kFewRowsSpt * 16;
  const bool has_rw = DUAL && RW != nullptr;
  for (int e = threadIdx.x * 2; e < rows * cols; e += 512) {
    *reinterpret_cast<double2*>(ws + e) = ld2(W + e);
    *reinterpret_cast<double2*>(rws + e) = has_rw ? ld2(RW + e) : make_double2(0.0, 0.0);
  }
  const int half = cols / 2;
  const bool with_colsum = p.epi == EPI_SIGMOID_BWD_COLSUM;  // (the host checked: vectorised path, one column pair per thread)
  const bool sig = p.epi == EPI_SIGMOID_BWD || with_colsum;
  const bool vec_ok = (p.ldo & 1) == 0 && (!sig || (p.lds & 1) == 0);
  const bool fused_sig = sig && vec_ok;
  double2 cs = make_double2(0.0, 0.0), csr = make_double2(0.0, 0.0);  // column sums of this thread's pair over the block's samples
  const long long groups = (n + kFewRowsSpt - 1) / kFewRowsSpt;
  for (long long ig = blockIdx.x; ig < groups; ig += gridDim.x) {
    const long long i0 = ig * kFewRowsSpt;
    __syncthreads();  // the previous group's upstream rows have been consumed (and, first trip, W / RW are staged)
    if (threadIdx.x < kFewRowsSpt * 16) {
      const int q = threadIdx.x >> 4, m = threadIdx.x & 15;
      const bool in = m < rows && i0 + q < n;
      us[threadIdx.x] = in ? U[(i0 + q) * rows + m] : 0.0;
      urs[threadIdx.x] = (DUAL && UR && in) ? UR[(i0 + q) * rows + m] : 0.0;
    }
    __syncthreads();
    for (int kp = threadIdx.x; kp < half; kp += 256) {
      const int k = 2 * kp;
      // the sigmoid-backward epilogue reads S / RS at the output position: issue those loads first
      double2 sx[kFewRowsSpt], rsx[kFewRowsSpt];
#pragma unroll
      for (int q = 0; q < kFewRowsSpt; ++q) {
        sx[q] = rsx[q] = make_double2(0.0, 0.0);
        if (fused_sig && i0 + q < n) {
          sx[q] = ld2(p.s + (i0 + q) * p.lds + k);
          if (DUAL && p.rs) rsx[q] = ld2(p.rs + (i0 + q) * p.lds + k);
        }
      }
      double2 d[kFewRowsSpt], dr[kFewRowsSpt];
#pragma unroll
      for (int q = 0; q < kFewRowsSpt; ++q) d[q] = dr[q] = make_double2(0.0, 0.0);
#pragma unroll 2
      for (int m = 0; m < rows; ++m) {
        const double2 wv = *reinterpret_cast<const double2*>(ws + m * cols + k);
        const double2 rwv = *reinterpret_cast<const double2*>(rws + m * cols + k);
#pragma unroll
        for (int q = 0; q < kFewRowsSpt; ++q) {
          const double um = us[q * 16 + m];
          d[q].x = fma(um, wv.x, d[q].x); d[q].y = fma(um, wv.y, d[q].y);
          if (DUAL) {
            const double urm = urs[q * 16 + m];
            dr[q].x = fma(urm, wv.x, fma(um, rwv.x, dr[q].x));
            dr[q].y = fma(urm, wv.y, fma(um, rwv.y, dr[q].y));
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kFewRowsSpt; ++q) {
        if (i0 + q >= n) continue;
        if (fused_sig) {
          // math_funcs.go:365-371, operand order kept: u' = x(1-x)u ; uR' = xR(1-x)u - x xR u + x(1-x)uR
          const long long o = (i0 + q) * p.ldo + k;
          const double x0 = sx[q].x, x1 = sx[q].y, u0 = d[q].x, u1 = d[q].y;
          const double2 v = make_double2(x0 * (1 - x0) * u0, x1 * (1 - x1) * u1);
          if (p.out0) *reinterpret_cast<double2*>(p.out0 + o) = v;
          cs.x += v.x; cs.y += v.y;
          if (DUAL && p.out1) {
            const double r0 = rsx[q].x, r1 = rsx[q].y;
            const double2 vr = make_double2(r0 * (1 - x0) * u0 - x0 * r0 * u0 + x0 * (1 - x0) * dr[q].x,
                                            r1 * (1 - x1) * u1 - x1 * r1 * u1 + x1 * (1 - x1) * dr[q].y);
            *reinterpret_cast<double2*>(p.out1 + o) = vr;
            csr.x += vr.x; csr.y += vr.y;
          }
        } else {
          epi_pair<DUAL>(p, vec_ok, i0 + q, k, d[q].x, d[q].y, dr[q].x, dr[q].y);
        }
      }
    }
  }
  if (with_colsum && (int)threadIdx.x < half) {  // lin_add.go:94-104 for the layer below: gb += colsum(D), rgb += colsum(DR)
    double* gb = const_cast<double*>(p.bias0);
    double* rgb = const_cast<double*>(p.bias1);
    if (gb) { atomicAdd(gb + 2 * threadIdx.x, cs.x); atomicAdd(gb + 2 * threadIdx.x + 1, cs.y); }
    if (DUAL && rgb) { atomicAdd(rgb + 2 * threadIdx.x, csr.x); atomicAdd(rgb + 2 * threadIdx.x + 1, csr.y); }
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// nmax: 8 for all three passes (ctx->small_wgrad_max for the weight gradient; 16 in round 1, when the rank-n update
// streamed faster than the tensor-core tiles up to there -- 29.6 vs 41.0 us on C1, n = 10 -- until the tiles' accumulate
// epilogue stopped serialising its read-modify-writes: 22.5 us; the other two passes measured 52 vs 27 and 44 vs 26 us)
bool small_n_applicable(af_ctx* ctx, long long rows, long long cols, long long n, std::initializer_list<const void*> ptrs,
                        int nmax) {
  if (ctx->gemm_path == 1 || ctx->small_n_off) return false;
  if (n < 1 || n > nmax || (cols & 1) || rows < 1 || cols < 2 || rows > (1 << 30) || cols > (1 << 30)) return false;
  for (const void* p : ptrs)
    if (p && !aligned16(p)) return false;
  return true;
}

// S (and RS) = epilogue(X W^T + ..., X RW^T + RX W^T + ...); `e` carries bias / activation like gemm_launch's
int small_fwd_launch(af_ctx* ctx, const double* W, const double* RW, const double* X, const double* RX, long long rows,
                     long long cols, long long n, const GemmEpilogue& e, bool dual) {
  GemmArgs a{};
  a.I = (int)n; a.J = (int)rows; a.Kd = (int)cols;
  a.epi = e.epi; a.out0 = e.out0; a.out1 = dual ? e.out1 : nullptr; a.ldo = e.ldo;
  a.bias0 = e.bias0; a.bias1 = e.bias1;
  int wpr = 1;  // warps per row: enough blocks to cover the chip twice
  while (wpr < 8 && (rows * wpr + 7) / 8 < 2ll * ctx->sm_count) wpr *= 2;
  const unsigned blocks = (unsigned)((rows + 8 / wpr - 1) / (8 / wpr));
  const double* rw = dual ? RW : nullptr;
  const double* rx = dual ? RX : nullptr;
  if (n <= 2) small_fwd_kernel<2><<<blocks, kSmallThreads, 0, ctx->stream>>>(W, rw, X, rx, (int)rows, (int)cols, (int)n, wpr, a);
  else if (n <= 8) small_fwd_kernel<8><<<blocks, kSmallThreads, 0, ctx->stream>>>(W, rw, X, rx, (int)rows, (int)cols, (int)n, wpr, a);
  else small_fwd_kernel<16><<<blocks, kSmallThreads, 0, ctx->stream>>>(W, rw, X, rx, (int)rows, (int)cols, (int)n, wpr, a);
  const int products = 1 + (dual ? (rw ? 1 : 0) + (rx ? 1 : 0) : 0);
  return after_launch(ctx, "small_fwd_kernel", KIND_SMALL_N, 8.0 * rows * cols * (rw ? 2 : 1) + 0.0 * products);
}

int small_wgrad_launch(af_ctx* ctx, const double* U, const double* UR, const double* X, const double* RX, double* gW,
                       double* rgW, long long rows, long long cols, long long n) {
  dim3 grid((unsigned)((cols / 2 + kSmallThreads - 1) / kSmallThreads), (unsigned)rows);
  if (n <= 2) small_wgrad_kernel<2><<<grid, kSmallThreads, 0, ctx->stream>>>(U, UR, X, RX, gW, rgW, (int)rows, (int)cols, (int)n);
  else if (n <= 8) small_wgrad_kernel<8><<<grid, kSmallThreads, 0, ctx->stream>>>(U, UR, X, RX, gW, rgW, (int)rows, (int)cols, (int)n);
  else small_wgrad_kernel<16><<<grid, kSmallThreads, 0, ctx->stream>>>(U, UR, X, RX, gW, rgW, (int)rows, (int)cols, (int)n);
  return after_launch(ctx, "small_wgrad_kernel", KIND_SMALL_N, 16.0 * rows * cols * ((gW ? 1 : 0) + (rgW ? 1 : 0)));
}

int gemm_deferred_epilogue(af_ctx* ctx, const GemmArgs& logical, bool dual);

int small_igrad_launch(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW, long long rows,
                       long long cols, long long n, const GemmEpilogue& e, bool dual) {
  double* D = e.out0;
  double* DR = dual ? e.out1 : nullptr;
  if (D) AF_CUDA(cudaMemsetAsync(D, 0, (size_t)(n * cols) * 8, ctx->stream));
  if (DR) AF_CUDA(cudaMemsetAsync(DR, 0, (size_t)(n * cols) * 8, ctx->stream));
  const long long xblocks = (cols / 2 + kSmallThreads - 1) / kSmallThreads;
  long long ychunks = (2ll * ctx->sm_count + xblocks - 1) / xblocks;
  if (ychunks > rows) ychunks = rows;
  if (ychunks < 1) ychunks = 1;
  int mchunk = (int)((rows + ychunks - 1) / ychunks);
  if (mchunk > 128) mchunk = 128;
  dim3 grid((unsigned)xblocks, (unsigned)((rows + mchunk - 1) / mchunk));
  const double* ur = dual ? UR : nullptr;
  const double* rw = dual ? RW : nullptr;
#define AF_SMALL_IGRAD(NM)                                                                                              \
  small_igrad_kernel<NM><<<grid, kSmallThreads, (size_t)2 * NM * mchunk * 8, ctx->stream>>>(U, ur, W, rw, D, DR, (int)rows, \
                                                                                            (int)cols, (int)n, mchunk)
  if (n <= 2) AF_SMALL_IGRAD(2);
  else if (n <= 8) AF_SMALL_IGRAD(8);
  else AF_SMALL_IGRAD(16);
#undef AF_SMALL_IGRAD
  AF_TRY(after_launch(ctx, "small_igrad_kernel", KIND_SMALL_N, 8.0 * rows * cols * (rw ? 2 : 1)));
  if (e.epi == EPI_STORE) return AF_OK;
  GemmArgs a{};
  a.I = (int)n; a.J = (int)cols; a.epi = e.epi; a.out0 = D; a.out1 = DR; a.ldo = e.ldo;
  a.s = e.s; a.rs = e.rs; a.lds = e.lds;
  return gemm_deferred_epilogue(ctx, a, dual);
}

bool few_rows_igrad_applicable(af_ctx* ctx, long long rows, long long cols, long long n, std::initializer_list<const void*> ptrs) {
  if (ctx->gemm_path == 1 || ctx->small_n_off) return false;
  if (rows < 1 || rows > 16 || n <= 16 || (cols & 1) || cols < 2 || cols > (1 << 30)) return false;
  if (2 * rows * cols * 8 > 100 * 1024) return false;  // W and RW staged in shared memory, two blocks per SM
  for (const void* p : ptrs)
    if (p && !aligned16(p)) return false;
  return true;
}

int few_rows_igrad_launch(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW, long long rows,
                          long long cols, long long n, const GemmEpilogue& e, bool dual) {
  GemmArgs a{};
  a.I = (int)n; a.J = (int)cols; a.Kd = (int)rows; a.epi = e.epi; a.out0 = e.out0; a.out1 = dual ? e.out1 : nullptr;
  a.ldo = e.ldo; a.s = e.s; a.rs = e.rs; a.lds = e.lds;
  // the bias gradient of the layer below (column sums of what this pass stores) rides along when every thread keeps ONE
  // column pair for the whole launch (cols <= 512) and the fused, vectorised sigmoid-backward path is taken
  if (e.epi == EPI_SIGMOID_BWD && e.colsum_done && (e.colsum0 || e.colsum1) && cols <= 512 && (e.ldo & 1) == 0 && (e.lds & 1) == 0) {
    a.epi = EPI_SIGMOID_BWD_COLSUM;
    a.bias0 = e.colsum0; a.bias1 = dual ? e.colsum1 : nullptr;
    *e.colsum_done = true;
  }
  long long blocks = (n + kFewRowsSpt - 1) / kFewRowsSpt;  // one group of four samples per block trip; persistent blocks
  if (blocks > 2ll * ctx->sm_count) blocks = 2ll * ctx->sm_count;
  const size_t smem = ((size_t)2 * rows * cols + 2 * kFewRowsSpt * 16) * 8;
  static std::atomic<unsigned long long> configured{0};
  if (!(configured.load(std::memory_order_acquire) >> (ctx->device & 63) & 1ull)) {
    AF_CUDA(cudaFuncSetAttribute(few_rows_igrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 102 * 1024));
    AF_CUDA(cudaFuncSetAttribute(few_rows_igrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 102 * 1024));
    configured.fetch_or(1ull << (ctx->device & 63), std::memory_order_release);
  }
  if (dual) few_rows_igrad_kernel<true><<<(unsigned)blocks, 256, smem, ctx->stream>>>(U, UR, W, RW, (int)rows, (int)cols, n, a);
  else few_rows_igrad_kernel<false><<<(unsigned)blocks, 256, smem, ctx->stream>>>(U, nullptr, W, nullptr, (int)rows, (int)cols, n, a);
  const double out_vectors = (dual ? 2.0 : 1.0) * (e.epi == EPI_SIGMOID_BWD ? 2.0 : 1.0);
  return after_launch(ctx, "few_rows_igrad_kernel", KIND_SMALL_N, 8.0 * n * cols * out_vectors);
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward through a layer with FEW OUTPUT ROWS and a large batch (the 10-class head of the bench network, the LSTM's
// 10-wide output layer over all T x B rows): S = act(X W^T + b), RS likewise (lin_tran.go:79-177, lin_add.go:13-50,
// math_funcs.go:288-315).  rows <= 16 is no tensor-core work -- as 64 x 16 tiles with split-K atomics, a zeroing pass
// and a deferred elementwise pass it took 21 + 6 us on the bench step -- it is one streaming pass over X / RX: a warp
// takes two samples, its lanes split the contraction in column pairs (W / RW, a few tens of KB, come from L1), the
// partial sums are combined with shuffles, and lane m finishes output m.
// With `bwd` the same lane also runs the TOP of the backward pass on its output (bench/mlp.go:52-54 calls
// PropagateRGradient right after ApplyR): the sigmoid backward of math_funcs.go:365-371 on the caller's upstream --
// or on outGrad = 1 / outRGrad = 0 (bench/mlp.go:118-126) generated in registers when `unit` -- written in place, and
// the bias gradient (lin_add.go:94-104), so the fill / zero / sigmoid-backward / column-sum launches of the head vanish.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kFewRowsMax = 16;

// W / RW are staged ONCE per block in shared memory (rows x cols doubles each; the bench head: 2 x 40 KB) and the blocks
// are persistent over the batch (two per SM): the first version read them through L1 with 226 registers and one block
// per SM -- 24 % of those loads missed, every warp stalled on each of them (ncu: long_scoreboard, 12 % warps active) and
// the kernel took 75 us for 33 MB.
template <bool DUAL>
__global__ void __launch_bounds__(256, 2) few_rows_fwd_kernel(const double* __restrict__ W, const double* __restrict__ RW,
                                                              const double* __restrict__ b, const double* __restrict__ Rb,
                                                              const double* __restrict__ X, const double* __restrict__ RX,
                                                              double* __restrict__ S, double* __restrict__ RS, int act, int bwd, int unit,
                                                              double* __restrict__ U, double* __restrict__ UR, double* __restrict__ gb,
                                                              double* __restrict__ rgb, int rows, int cols, long long n) {
  extern __shared__ __align__(16) double few_smem[];
  double* ws = few_smem;                     // [rows][cols]
  double* rws = few_smem + (size_t)rows * cols;  // [rows][cols] (DUAL with RW only)
  const bool has_rw = DUAL && RW != nullptr;
  for (int e = threadIdx.x * 2; e < rows * cols; e += 512) {
    *reinterpret_cast<double2*>(ws + e) = ld2(W + e);
    if (has_rw) *reinterpret_cast<double2*>(rws + e) = ld2(RW + e);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * 256) >> 5;
  const int half = cols / 2;
  const bool has_rx = DUAL && RX != nullptr;
  double gsum = 0.0, rgsum = 0.0;  // lane m: bias-gradient partial of output m over this warp's samples
  for (long long i = warp; i < n; i += nwarps) {
    double z[kFewRowsMax], rz[DUAL ? kFewRowsMax : 1];
#pragma unroll
    for (int m = 0; m < kFewRowsMax; ++m) {
      z[m] = 0.0;
      if (DUAL) rz[m] = 0.0;
    }
    const double* xr = X + i * cols;
    const double* rxr = has_rx ? RX + i * cols : nullptr;
    // two column pairs per trip: four independent global loads in flight before the first use
    for (int kp = lane; kp < half; kp += 64) {
      const int kq = kp + 32;
      const bool second = kq < half;
      const double2 x0 = ld2(xr + 2 * kp);
      const double2 x1 = second ? ld2(xr + 2 * kq) : make_double2(0.0, 0.0);
      double2 r0 = make_double2(0.0, 0.0), r1 = make_double2(0.0, 0.0);
      if (has_rx) {
        r0 = ld2(rxr + 2 * kp);
        if (second) r1 = ld2(rxr + 2 * kq);
      }
      const int kq_s = second ? kq : kp;  // (x1 = r1 = 0 then: the duplicate column pair contributes nothing)
#pragma unroll
      for (int m = 0; m < kFewRowsMax; ++m) {
        if (m < rows) {
          const double2 w0 = *reinterpret_cast<const double2*>(ws + m * cols + 2 * kp);
          const double2 w1 = *reinterpret_cast<const double2*>(ws + m * cols + 2 * kq_s);
          z[m] = fma(x1.y, w1.y, fma(x1.x, w1.x, fma(x0.y, w0.y, fma(x0.x, w0.x, z[m]))));
          if (DUAL) {
            double a = rz[m];
            if (has_rw) {
              const double2 q0 = *reinterpret_cast<const double2*>(rws + m * cols + 2 * kp);
              const double2 q1 = *reinterpret_cast<const double2*>(rws + m * cols + 2 * kq_s);
              a = fma(x1.y, q1.y, fma(x1.x, q1.x, fma(x0.y, q0.y, fma(x0.x, q0.x, a))));
            }
            if (has_rx) a = fma(r1.y, w1.y, fma(r1.x, w1.x, fma(r0.y, w0.y, fma(r0.x, w0.x, a))));
            rz[m] = a;
          }
        }
      }
    }
    // butterfly: afterwards every lane holds every total; lane m keeps output m (selected with predicated moves --
    // a lane-indexed read of the register arrays would put them in local memory)
    double mine = 0.0, rmine = 0.0;
#pragma unroll
    for (int m = 0; m < kFewRowsMax; ++m) {
      if (m < rows) {
        double v = z[m], rv = DUAL ? rz[m] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          v += __shfl_xor_sync(0xffffffffu, v, o);
          if (DUAL) rv += __shfl_xor_sync(0xffffffffu, rv, o);
        }
        if (lane == m) { mine = v; rmine = rv; }
      }
    }
    if (lane < rows) {
      const long long o = i * rows + lane;
      const double zb = mine + (b ? b[lane] : 0.0);      // lin_add.go:15-17
      const double s = act ? sigmoid_f64(zb) : zb;       // math_funcs.go:292
      double rs = 0.0;
      if (DUAL) {
        const double rzb = rmine + (Rb ? Rb[lane] : 0.0);  // lin_add.go:37-42
        rs = act ? s * (1 - s) * rzb : rzb;                // math_funcs.go:305-309
      }
      if (S) S[o] = s;
      if (DUAL && RS) RS[o] = rs;
      if (bwd) {
        // math_funcs.go:365-371, operand order kept: u' = x(1-x)u ; uR' = xR(1-x)u - x xR u + x(1-x)uR
        const double u = unit ? 1.0 : U[o];
        const double du = s * (1 - s) * u;
        U[o] = du;
        gsum += du;
        if (DUAL) {
          const double ur = (unit || !UR) ? 0.0 : UR[o];
          const double dur = rs * (1 - s) * u - s * rs * u + s * (1 - s) * ur;
          UR[o] = dur;
          rgsum += dur;
        }
      }
    }
  }
  if (bwd && (gb || (DUAL && rgb))) {  // lin_add.go:94-104 batched: gb += colsum(U'), rgb += colsum(UR')
    // one atomic per block and output: thousands of same-address float64 atomics would serialise in L2
    __shared__ double part[2][8][kFewRowsMax];
    const int w = threadIdx.x >> 5;
    if (lane < kFewRowsMax) { part[0][w][lane] = gsum; part[1][w][lane] = rgsum; }
    __syncthreads();
    if (threadIdx.x < rows) {
      double a = 0.0, r = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { a += part[0][q][threadIdx.x]; r += part[1][q][threadIdx.x]; }
      if (gb) atomicAdd(gb + threadIdx.x, a);
      if (DUAL && rgb) atomicAdd(rgb + threadIdx.x, r);
    }
  }
}

bool few_rows_fwd_applicable(af_ctx* ctx, long long rows, long long cols, long long n, std::initializer_list<const void*> ptrs) {
  if (ctx->gemm_path == 1 || ctx->small_n_off) return false;
  if (rows < 1 || rows > kFewRowsMax || n <= 16 || (cols & 1) || cols < 2 || cols > (1 << 30)) return false;
  if (2 * rows * cols * 8 > 100 * 1024) return false;  // W and RW staged in shared memory, two blocks per SM
  for (const void* p : ptrs)
    if (p && !aligned16(p)) return false;
  return true;
}

// S / RS = act(X W^T + b) as one streaming pass; `head` != NULL adds the top of the backward pass (see the kernel)
int few_rows_fwd_launch(af_ctx* ctx, const double* W, const double* RW, const double* b, const double* Rb, const double* X,
                        const double* RX, double* S, double* RS, int act, const FewRowsHead* head, long long rows, long long cols,
                        long long n, bool dual) {
  long long blocks = (n + 7) / 8;  // one warp per sample at a time; persistent blocks, two per SM
  if (blocks > 2ll * ctx->sm_count) blocks = 2ll * ctx->sm_count;
  const int bwd = head ? 1 : 0, unit = head && head->unit ? 1 : 0;
  double* U = head ? head->U : nullptr;
  double* UR = head ? head->UR : nullptr;
  double* gb = head ? head->gb : nullptr;
  double* rgb = head ? head->rgb : nullptr;
  const size_t smem = (size_t)rows * cols * 8 * ((dual && RW) ? 2 : 1);
  static std::atomic<unsigned long long> configured{0};
  if (!(configured.load(std::memory_order_acquire) >> (ctx->device & 63) & 1ull)) {
    AF_CUDA(cudaFuncSetAttribute(few_rows_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    AF_CUDA(cudaFuncSetAttribute(few_rows_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured.fetch_or(1ull << (ctx->device & 63), std::memory_order_release);
  }
  if (dual)
    few_rows_fwd_kernel<true><<<(unsigned)blocks, 256, smem, ctx->stream>>>(W, RW, b, Rb, X, RX, S, RS, act, bwd, unit, U, UR, gb, rgb,
                                                                            (int)rows, (int)cols, n);
  else
    few_rows_fwd_kernel<false><<<(unsigned)blocks, 256, smem, ctx->stream>>>(W, nullptr, b, nullptr, X, nullptr, S, nullptr, act, bwd,
                                                                             unit, U, nullptr, gb, nullptr, (int)rows, (int)cols, n);
  return after_launch(ctx, "few_rows_fwd_kernel", KIND_SMALL_N, 8.0 * n * cols * (dual && RX ? 2.0 : 1.0));
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward through a layer with very FEW INPUT COLUMNS and a large batch: Y = X W^T + b (RY = X RW^T + Rb; X has no
// R-vector), cols <= 32.  The LSTM's x projection (bench/lstm.go:140 -- the x part of concat(state, x), I = 30 wide, for
// all T x B rows at once, lstm.cu) is this shape: 32768 x 2048 outputs with a contraction of 30.  As tensor-core tiles that
// is two k-tiles per 64 x 64 tile, all prologue and epilogue (measured 0.94 ms for 1.07 GB of output = 1.1 TB/s); as a
// register-tiled streaming pass -- a block stages 64 rows of X and 64 rows of W / RW transposed in shared memory, every
// thread accumulates a 4 x 4 block of outputs for both products -- it is bound by the output stream.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kFewColsMax = 32;

template <bool DUAL>
__global__ void __launch_bounds__(256) few_cols_fwd_kernel(const double* __restrict__ X, long long ldx, const double* __restrict__ W,
                                                           const double* __restrict__ RW, long long ldw,
                                                           const double* __restrict__ b, const double* __restrict__ Rb,
                                                           double* __restrict__ Y, double* __restrict__ RY, long long ldy, long long n,
                                                           int rows, int K) {
  __shared__ double xs[kFewColsMax][64], ws[kFewColsMax][64], rws[DUAL ? kFewColsMax : 1][64];
  const long long r0 = (long long)blockIdx.y * 64;
  const int c0 = blockIdx.x * 64;
  for (int e = threadIdx.x; e < 64 * K; e += 256) {
    const int r = e / K, k = e - r * K;
    xs[k][r] = (r0 + r < n) ? X[(r0 + r) * ldx + k] : 0.0;
    const bool in = c0 + r < rows;
    ws[k][r] = in ? W[(long long)(c0 + r) * ldw + k] : 0.0;
    if (DUAL) rws[k][r] = (in && RW) ? RW[(long long)(c0 + r) * ldw + k] : 0.0;  // RW == NULL: zero R-vector (variable.go:85-91)
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4], racc[DUAL ? 4 : 1][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[i][j] = 0.0;
      if (DUAL) racc[i][j] = 0.0;
    }
  for (int k = 0; k < K; ++k) {
    const double4 xv = *reinterpret_cast<const double4*>(&xs[k][ty * 4]);
    const double4 wv = *reinterpret_cast<const double4*>(&ws[k][tx * 4]);
    const double x[4] = {xv.x, xv.y, xv.z, xv.w}, w[4] = {wv.x, wv.y, wv.z, wv.w};
    double rw[4] = {0.0, 0.0, 0.0, 0.0};
    if (DUAL) {
      const double4 rv = *reinterpret_cast<const double4*>(&rws[k][tx * 4]);
      rw[0] = rv.x; rw[1] = rv.y; rw[2] = rv.z; rw[3] = rv.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fma(x[i], w[j], acc[i][j]);
        if (DUAL) racc[DUAL ? i : 0][j] = fma(x[i], rw[j], racc[DUAL ? i : 0][j]);
      }
  }
  const int c = c0 + tx * 4;
  const bool vec = ((ldy & 1) == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0) && (!DUAL || (reinterpret_cast<uintptr_t>(RY) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = r0 + ty * 4 + i;
    if (r >= n) continue;
    double y[4], ry[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in = c + j < rows;
      y[j] = acc[i][j] + ((b && in) ? b[c + j] : 0.0);                                    // lin_add.go:15-17
      ry[j] = DUAL ? racc[DUAL ? i : 0][j] + ((Rb && in) ? Rb[c + j] : 0.0) : 0.0;        // lin_add.go:37-42
    }
    if (vec && c + 3 < rows) {
      *reinterpret_cast<double2*>(Y + r * ldy + c) = make_double2(y[0], y[1]);
      *reinterpret_cast<double2*>(Y + r * ldy + c + 2) = make_double2(y[2], y[3]);
      if (DUAL) {
        *reinterpret_cast<double2*>(RY + r * ldy + c) = make_double2(ry[0], ry[1]);
        *reinterpret_cast<double2*>(RY + r * ldy + c + 2) = make_double2(ry[2], ry[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (c + j >= rows) continue;
        Y[r * ldy + c + j] = y[j];
        if (DUAL) RY[r * ldy + c + j] = ry[j];
      }
    }
  }
}

bool few_cols_fwd_applicable(af_ctx* ctx, long long rows, long long cols, long long n) {
  return ctx->gemm_path != 1 && !ctx->small_n_off && cols >= 1 && cols <= kFewColsMax && n > 16 && rows >= 1 &&
         (n + 63) / 64 <= 65535 && rows < (1ll << 30);
}

// Y (n x rows, pitch ldy) = X (n x cols, pitch ldx) W^T (rows x cols, pitch ldw) + b; RY likewise with RW, Rb when RY != NULL
int few_cols_fwd_launch(af_ctx* ctx, const double* X, long long ldx, const double* W, const double* RW, long long ldw,
                        const double* b, const double* Rb, double* Y, double* RY, long long ldy, long long rows, long long cols,
                        long long n) {
  const dim3 grid((unsigned)((rows + 63) / 64), (unsigned)((n + 63) / 64));
  if (RY) few_cols_fwd_kernel<true><<<grid, 256, 0, ctx->stream>>>(X, ldx, W, RW, ldw, b, Rb, Y, RY, ldy, n, (int)rows, (int)cols);
  else few_cols_fwd_kernel<false><<<grid, 256, 0, ctx->stream>>>(X, ldx, W, nullptr, ldw, b, nullptr, Y, nullptr, ldy, n, (int)rows, (int)cols);
  return after_launch(ctx, "few_cols_fwd_kernel", KIND_SMALL_N, 8.0 * n * rows * (RY ? 2.0 : 1.0));
}

}  // namespace af
