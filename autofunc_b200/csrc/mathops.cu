// the elementwise functions, scalar-broadcast ops, cost heads and the parameter update that sit
// either side of the dense path (math_funcs.go: Log, LogSigmoid, Sin, Norm, Softmax; arithmetic.go: AddScaler, Div,
// ScaleFirst, AddFirst, Inverse, Pow, the log-domain sums; gradient.go: AddToVars).
//
// Same design as elementwise.cu: every kernel is one HBM pass, computes the value form and the R form together, and the
// backward kernels overwrite the upstreams in place as the reference does (result.go:24-29).  Scalars that the
// reference reads on the host (`scaler.Output()[0]`, `MaxAbs()`, the upstream of a SumAll) stay on the device and are
// passed BY POINTER, so a chain of these ops never synchronises the stream.
#include <atomic>
#include "common.h"
#include "ew.cuh"

#include <algorithm>

namespace af {
namespace {

__device__ __forceinline__ double rd(const double* p, long long i) { return p ? p[i] : 0.0; }  // NULL R operand = 0

// ---- row-wise softmax ----------------------------------------------------------------------------
// A group of G threads owns one row: G = 8 / 16 / 32 lanes of a warp for short rows (several rows per warp, so a
// 10-class head still issues full coalesced warps), G = 256 (one block) for long ones.  Each thread keeps its first
// kRowCache exponentials in registers, so rows up to G * kRowCache entries are read from HBM once and exp() is
// evaluated once; only the excess of longer rows is recomputed from L2.
constexpr int kRowCache = 8;  // forward: exponentials (and their R forms) per thread
constexpr int kBwdCache = 4;  // backward keeps four vectors per entry, so half as many entries (registers)

template <int G>
__device__ __forceinline__ double group_sum(double v, double* sm) {
#pragma unroll
  for (int o = (G < 32 ? G : 32) / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (G <= 32) return v;
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
  const bool live = row < n;  // a dead group still takes part in the shuffles / barriers (its sums are zero)
  const double* x = X + row * len;
  const double* xr = RX ? RX + row * len : nullptr;
  double ev[kRowCache], erv[kRowCache];
  double s = 0.0, sr = 0.0;
  if (live) {
#pragma unroll
    for (int j = 0; j < kRowCache; ++j) {
      const long long k = lane + (long long)j * G;
      ev[j] = erv[j] = 0.0;
      if (k < len) {
        ev[j] = exp(x[k] * invT);
        s += ev[j];
        if (RY) { erv[j] = ev[j] * (rd(xr, k) * invT); sr += erv[j]; }
      }
    }
    for (long long k = lane + (long long)kRowCache * G; k < len; k += G) {
      const double e = exp(x[k] * invT);
      s += e;
      if (RY) sr += e * (rd(xr, k) * invT);
    }
  }
  s = group_sum<G>(s, sm);
  if (RY) sr = group_sum<G>(sr, sm);
  if (!live) return;
  const double inv = 1.0 / s;
  const double invR = -(inv * inv) * sr;
#pragma unroll
  for (int j = 0; j < kRowCache; ++j) {
    const long long k = lane + (long long)j * G;
    if (k < len) {
      Y[row * len + k] = ev[j] * inv;
      if (RY) RY[row * len + k] = ev[j] * invR + erv[j] * inv;
    }
  }
  for (long long k = lane + (long long)kRowCache * G; k < len; k += G) {
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
  const long long o = row * len;
  double uv[kBwdCache], yv[kBwdCache], urv[kBwdCache], ryv[kBwdCache];
  double a = 0.0, b = 0.0;  // <u,y> ; <uR,y> + <u,yR>
  if (live) {
#pragma unroll
    for (int j = 0; j < kBwdCache; ++j) {
      const long long k = lane + (long long)j * G;
      uv[j] = yv[j] = urv[j] = ryv[j] = 0.0;
      if (k < len) {
        uv[j] = U[o + k]; yv[j] = Y[o + k];
        a += uv[j] * yv[j];
        if (UR) { urv[j] = UR[o + k]; ryv[j] = rd(RY, o + k); b += urv[j] * yv[j] + uv[j] * ryv[j]; }
      }
    }
    for (long long k = lane + (long long)kBwdCache * G; k < len; k += G) {
      const double u = U[o + k], y = Y[o + k];
      a += u * y;
      if (UR) b += UR[o + k] * y + u * rd(RY, o + k);
    }
  }
  a = group_sum<G>(a, sm);
  if (UR) b = group_sum<G>(b, sm);
  if (!live) return;
#pragma unroll
  for (int j = 0; j < kBwdCache; ++j) {
    const long long k = lane + (long long)j * G;
    if (k < len) {
      if (UR) UR[o + k] = invT * (ryv[j] * (uv[j] - a) + yv[j] * (urv[j] - b));
      U[o + k] = invT * (yv[j] * (uv[j] - a));
    }
  }
  for (long long k = lane + (long long)kBwdCache * G; k < len; k += G) {
    const double u = U[o + k], y = Y[o + k];
    if (UR) UR[o + k] = invT * (rd(RY, o + k) * (u - a) + y * (UR[o + k] - b));
    U[o + k] = invT * (y * (u - a));
  }
}

// ---- staged row-wise softmax: rows up to kSoftChunk entries ------------------------------------------
// A block stages a contiguous chunk of WHOLE rows (up to kSoftChunk doubles per vector) in shared memory with one
// coalesced cooperative load -- every thread has 8 independent loads per vector in flight, whatever the row length --
// then groups of G threads reduce and rescale their rows in place in shared memory, and the block stores the chunk
// back coalesced.  Each vector crosses HBM exactly once, exp() is evaluated once, and a 10-class head moves full
// 256-byte warp transactions instead of 80-byte rows.
constexpr int kSoftChunk = 2048;

template <int G>
__device__ __forceinline__ double staged_group_sum(double v, double* red, int row_in_pass) {
#pragma unroll
  for (int o = (G < 32 ? G : 32) / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (G <= 32) return v;
  // G == 256: one row per pass, the whole block reduces
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

__device__ __forceinline__ void stage_in(double* dst, const double* __restrict__ src, long long base, int elems) {
  // 8 x 256 = kSoftChunk: fully unrolled so the loads are all issued before the first use
#pragma unroll
  for (int u = 0; u < kSoftChunk / 256; ++u) {
    const int e = threadIdx.x + u * 256;
    if (e < elems) dst[e] = src ? src[base + e] : 0.0;
  }
}
__device__ __forceinline__ void stage_out(double* __restrict__ dst, const double* src, long long base, int elems) {
#pragma unroll
  for (int u = 0; u < kSoftChunk / 256; ++u) {
    const int e = threadIdx.x + u * 256;
    if (e < elems) dst[base + e] = src[e];
  }
}

template <int G>
__global__ void __launch_bounds__(256) softmax_fwd_staged_kernel(const double* __restrict__ X, const double* __restrict__ RX,
                                                                 double* __restrict__ Y, double* __restrict__ RY,
                                                                 double invT, int len, long long n, int rows_per_chunk) {
  extern __shared__ double sm_soft[];
  double* sx = sm_soft;                 // x -> e -> y
  double* sr = sm_soft + kSoftChunk;    // xR -> eR -> yR
  __shared__ double red[8];
  const long long row0 = (long long)blockIdx.x * rows_per_chunk;
  const int nrows = (int)min((long long)rows_per_chunk, n - row0);
  const int elems = nrows * len;
  const long long base = row0 * len;
  stage_in(sx, X, base, elems);
  if (RY) stage_in(sr, RX, base, elems);
  __syncthreads();
  constexpr int P = 256 / G;  // rows per pass
  const int lane = threadIdx.x % G, sub = threadIdx.x / G;
  for (int r0 = 0; r0 < nrows; r0 += P) {  // uniform trip count: dead groups still take part in the reductions
    const int r = r0 + sub;
    const bool live = r < nrows;
    double* x = sx + r * len;
    double* xr = sr + r * len;
    double s = 0.0, srsum = 0.0;
    if (live)
      for (int k = lane; k < len; k += G) {
        const double e = exp(x[k] * invT);
        x[k] = e;
        s += e;
        if (RY) { const double er = e * (xr[k] * invT); xr[k] = er; srsum += er; }
      }
    s = staged_group_sum<G>(s, red, sub);
    if (RY) srsum = staged_group_sum<G>(srsum, red, sub);
    if (live) {
      const double inv = 1.0 / s, invR = -(inv * inv) * srsum;
      for (int k = lane; k < len; k += G) {
        const double e = x[k];
        if (RY) xr[k] = e * invR + xr[k] * inv;
        x[k] = e * inv;
      }
    }
  }
  __syncthreads();
  stage_out(Y, sx, base, elems);
  if (RY) stage_out(RY, sr, base, elems);
}

template <int G>
__global__ void __launch_bounds__(256) softmax_bwd_staged_kernel(const double* __restrict__ Y, const double* __restrict__ RY,
                                                                 double* __restrict__ U, double* __restrict__ UR, double invT,
                                                                 int len, long long n, int rows_per_chunk) {
  extern __shared__ double sm_soft[];
  double* su = sm_soft;
  double* sy = sm_soft + kSoftChunk;
  double* sur = sm_soft + 2 * kSoftChunk;
  double* sry = sm_soft + 3 * kSoftChunk;
  __shared__ double red[8];
  const long long row0 = (long long)blockIdx.x * rows_per_chunk;
  const int nrows = (int)min((long long)rows_per_chunk, n - row0);
  const int elems = nrows * len;
  const long long base = row0 * len;
  stage_in(su, U, base, elems);
  stage_in(sy, Y, base, elems);
  if (UR) { stage_in(sur, UR, base, elems); stage_in(sry, RY, base, elems); }
  __syncthreads();
  constexpr int P = 256 / G;
  const int lane = threadIdx.x % G, sub = threadIdx.x / G;
  for (int r0 = 0; r0 < nrows; r0 += P) {
    const int r = r0 + sub;
    const bool live = r < nrows;
    const int o = r * len;
    double a = 0.0, b = 0.0;  // <u,y> ; <uR,y> + <u,yR>
    if (live)
      for (int k = lane; k < len; k += G) {
        const double u = su[o + k], y = sy[o + k];
        a += u * y;
        if (UR) b += sur[o + k] * y + u * sry[o + k];
      }
    a = staged_group_sum<G>(a, red, sub);
    if (UR) b = staged_group_sum<G>(b, red, sub);
    if (live)
      for (int k = lane; k < len; k += G) {
        const double u = su[o + k], y = sy[o + k];
        if (UR) sur[o + k] = invT * (sry[o + k] * (u - a) + y * (sur[o + k] - b));
        su[o + k] = invT * (y * (u - a));
      }
  }
  __syncthreads();
  stage_out(U, su, base, elems);
  if (UR) stage_out(UR, sur, base, elems);
}

// ---- short rows (len <= kShortRow: the 10-class head): ONE THREAD PER ROW for the sums only --------------------------
// The staged kernels above touch shared memory six times per entry (stage, read, write e, read e, write y, unstage), do
// all of a row's arithmetic in a 4- or 8-lane group between two block barriers, and reduce with shuffles; on rows of 10
// that serialised phase -- every exp() of the chunk waits for the chunk's last load -- and the shared-memory traffic cost
// as much as the HBM stream itself (0.64-0.70 of the copy rate).  Here the per-ENTRY work rides on the coalesced phases,
// spread over all 256 threads like in the elementwise kernels: the load phase evaluates e = exp(x invT) and eR as the
// loads arrive and parks them in shared memory at an ODD row pitch (len | 1: thread r reading entry k of row r is free
// of bank conflicts); one thread per row then only adds its row up (sequentially: no shuffles) and leaves the row's two
// scalars in shared memory; the store phase scales on the way out.  Every vector crosses HBM once, exp() is evaluated once.
constexpr int kShortRow = 32;  // (e * magic) >> 16 is exact while chunk * len <= 65536
constexpr int kShortChunk = 2048;     // doubles per staged vector (pitched), forward: 2 vectors = 32 KB per block
constexpr int kShortChunkBwd = 1024;  // backward: 4 vectors = 32 KB per block

struct ShortRowGeom {
  int len, pitch, rows_per_block, magic;  // e / len == (e * magic) >> 16 for e * len < 65536
};

// formulas and operand order of softmax_fwd_staged_kernel (math_funcs.go:487-509)
template <bool R>
__global__ void __launch_bounds__(256) softmax_fwd_short_kernel(const double* __restrict__ X, const double* __restrict__ RX,
                                                                double* __restrict__ Y, double* __restrict__ RY,
                                                                double invT, long long n, const ShortRowGeom g) {
  constexpr int PER = kShortChunk / 256;
  __shared__ double se[kShortChunk];             // e
  __shared__ double ser[R ? kShortChunk : 1];    // eR
  __shared__ double sinv[256], sinvr[R ? 256 : 1];
  const long long row0 = (long long)blockIdx.x * g.rows_per_block;
  const int nrows = (int)min((long long)g.rows_per_block, n - row0);
  const int elems = nrows * g.len;
  const long long base = row0 * g.len;
  int row[PER], at[PER];   // entry u of this thread: its row in the chunk and its pitched slot
  {
    double x[PER], xr[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int e = threadIdx.x + u * 256;
      row[u] = (e * g.magic) >> 16;
      at[u] = row[u] * g.pitch + (e - row[u] * g.len);
      x[u] = e < elems ? X[base + e] : 0.0;
      if (R) xr[u] = (e < elems && RX) ? RX[base + e] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < PER; ++u)
      if (threadIdx.x + u * 256 < elems) {
        const double ev = exp(x[u] * invT);
        se[at[u]] = ev;
        if (R) ser[at[u]] = ev * (xr[u] * invT);
      }
  }
  __syncthreads();
  if ((int)threadIdx.x < nrows) {
    const int o = threadIdx.x * g.pitch;
    double s = 0.0, srsum = 0.0;
#pragma unroll 4
    for (int k = 0; k < g.len; ++k) {
      s += se[o + k];
      if (R) srsum += ser[o + k];
    }
    const double inv = 1.0 / s;
    sinv[threadIdx.x] = inv;
    if (R) sinvr[threadIdx.x] = -(inv * inv) * srsum;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int e = threadIdx.x + u * 256;
    if (e < elems) {
      const double ev = se[at[u]], inv = sinv[row[u]];
      if (R) RY[base + e] = ev * sinvr[row[u]] + ser[at[u]] * inv;
      Y[base + e] = ev * inv;
    }
  }
}

template <bool R>
__global__ void __launch_bounds__(256) softmax_bwd_short_kernel(const double* __restrict__ Y, const double* __restrict__ RY,
                                                                double* __restrict__ U, double* __restrict__ UR, double invT,
                                                                long long n, const ShortRowGeom g) {
  constexpr int PER = kShortChunkBwd / 256;
  __shared__ double su[kShortChunkBwd], sy[kShortChunkBwd];
  __shared__ double sur[R ? kShortChunkBwd : 1], sry[R ? kShortChunkBwd : 1];
  __shared__ double sa[256], sb[R ? 256 : 1];
  const long long row0 = (long long)blockIdx.x * g.rows_per_block;
  const int nrows = (int)min((long long)g.rows_per_block, n - row0);
  const int elems = nrows * g.len;
  const long long base = row0 * g.len;
  int row[PER], at[PER];
  {
    double u_[PER], y_[PER], ur_[PER], ry_[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int e = threadIdx.x + u * 256;
      row[u] = (e * g.magic) >> 16;
      at[u] = row[u] * g.pitch + (e - row[u] * g.len);
      const bool ok = e < elems;
      u_[u] = ok ? U[base + e] : 0.0;
      y_[u] = ok ? Y[base + e] : 0.0;
      if (R) { ur_[u] = ok ? UR[base + e] : 0.0; ry_[u] = (ok && RY) ? RY[base + e] : 0.0; }
    }
#pragma unroll
    for (int u = 0; u < PER; ++u)
      if (threadIdx.x + u * 256 < elems) {
        su[at[u]] = u_[u]; sy[at[u]] = y_[u];
        if (R) { sur[at[u]] = ur_[u]; sry[at[u]] = ry_[u]; }
      }
  }
  __syncthreads();
  if ((int)threadIdx.x < nrows) {
    const int o = threadIdx.x * g.pitch;
    double a = 0.0, b = 0.0;  // <u,y> ; <uR,y> + <u,yR>
#pragma unroll 4
    for (int k = 0; k < g.len; ++k) {
      const double u = su[o + k], y = sy[o + k];
      a += u * y;
      if (R) b += sur[o + k] * y + u * sry[o + k];
    }
    sa[threadIdx.x] = a;
    if (R) sb[threadIdx.x] = b;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int e = threadIdx.x + u * 256;
    if (e < elems) {
      const double uv = su[at[u]], y = sy[at[u]], a = sa[row[u]];
      if (R) UR[base + e] = invT * (sry[at[u]] * (uv - a) + y * (sur[at[u]] - sb[row[u]]));
      U[base + e] = invT * (y * (uv - a));
    }
  }
}

inline ShortRowGeom short_row_geom(long long len, int chunk) {
  ShortRowGeom g;
  g.len = (int)len;
  g.pitch = (int)len | 1;
  g.rows_per_block = std::min(256, chunk / g.pitch);
  g.magic = (65536 + g.len - 1) / g.len;
  return g;
}

// group width for a row length: the smallest sub-warp that covers the row in <= 2 strides, a full warp up to
// 32 * kRowCache entries, one block beyond that
// one block pass per 256/G rows (measured: capping the grid and walking rows persistently is slower -- a block's rows
// then serialise on the load -> reduce -> store dependency; softmax_bwd rows of 1024: 4.5 vs 2.4 TB/s)
#define AF_ROW_GRID(G) (unsigned)((n + 256 / (G) - 1) / (256 / (G)))
#define AF_ROW_DISPATCH(KERNEL, ...)                                                              \
  do {                                                                                            \
    if (len <= 16) KERNEL<8><<<AF_ROW_GRID(8), 256, 0, ctx->stream>>>(__VA_ARGS__);               \
    else if (len <= 32) KERNEL<16><<<AF_ROW_GRID(16), 256, 0, ctx->stream>>>(__VA_ARGS__);        \
    else if (len <= 32 * kRowCache) KERNEL<32><<<AF_ROW_GRID(32), 256, 0, ctx->stream>>>(__VA_ARGS__); \
    else KERNEL<256><<<AF_ROW_GRID(256), 256, 0, ctx->stream>>>(__VA_ARGS__);                     \
  } while (0)

// group width for the staged kernels: 4 / 8 / 16 / 32 lanes per row up to 256 entries (about 2-8 entries per lane), the whole block beyond; the
// bwd variant needs 64 KB of dynamic shared memory (opt-in attribute, set once per instantiation)
#define AF_STAGED_LAUNCH(KERNEL, G, SMEM, ...)                                                     \
  do {                                                                                             \
    /* one bit per device (the attribute is per device and function), published after the call returns */ \
    static std::atomic<unsigned long long> configured{0};                                          \
    if (!(configured.load(std::memory_order_acquire) >> (ctx->device & 63) & 1ull)) {              \
      AF_CUDA(cudaFuncSetAttribute(KERNEL<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kSoftChunk * 8)); \
      configured.fetch_or(1ull << (ctx->device & 63), std::memory_order_release);                  \
    }                                                                                              \
    KERNEL<G><<<grid, 256, SMEM, ctx->stream>>>(__VA_ARGS__);                                      \
  } while (0)
#define AF_STAGED_DISPATCH(KERNEL, SMEM, ...)                                  \
  do {                                                                         \
    if (len <= 12) AF_STAGED_LAUNCH(KERNEL, 4, SMEM, __VA_ARGS__);             \
    else if (len <= 16) AF_STAGED_LAUNCH(KERNEL, 8, SMEM, __VA_ARGS__);        \
    else if (len <= 32) AF_STAGED_LAUNCH(KERNEL, 16, SMEM, __VA_ARGS__);       \
    else if (len <= 256) AF_STAGED_LAUNCH(KERNEL, 32, SMEM, __VA_ARGS__);      \
    else AF_STAGED_LAUNCH(KERNEL, 256, SMEM, __VA_ARGS__);                     \
  } while (0)

}  // namespace
}  // namespace af

using namespace af;

extern "C" {

// ---- Log (math_funcs.go:99-189) --------------------------------------------------------------------
int af_log_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "log_fwd_r", RO ? 32 : 16, EwIn<2>{{X, RO ? RX : nullptr}}, EwOut<2>{{O, RO}},
                   [=] __device__(long long, const double* x, double* y) {
                     y[0] = log(x[0]);
                     y[1] = x[1] / x[0];
                   });
}

int af_log_bwd_r(af_ctx* ctx, const double* X, const double* RX, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(U);
  const bool r = UR != nullptr;
  return ew_launch(ctx, len, "log_bwd_r", r ? 48 : 24, EwIn<4>{{X, r ? RX : nullptr, U, UR}}, EwOut<2>{{U, UR}},
                   [=] __device__(long long, const double* v, double* y) {
                     const double x = v[0], u = v[2];
                     y[1] = v[3] / x - u * v[1] / (x * x);
                     y[0] = r ? u / x : u * (1 / x);  // math_funcs.go:150 multiplies by 1/x, :184 divides
                   });
}

// ---- LogSigmoid (math_funcs.go:376-477) -------------------------------------------------------------
int af_logsigmoid_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  const bool r = RO != nullptr;
  return ew_launch(ctx, len, "logsigmoid_fwd_r", r ? 32 : 16, EwIn<2>{{X, r ? RX : nullptr}}, EwOut<2>{{O, RO}},
                   [=] __device__(long long, const double* v, double* y) {
                     const double x = v[0];
                     y[0] = x > 0 ? -log(1 + exp(-x)) : x - log(1 + exp(x));
                     if (r) y[1] = v[1] / (1 + exp(x));
                   });
}

int af_logsigmoid_bwd_r(af_ctx* ctx, const double* X, const double* RX, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(U);
  const bool r = UR != nullptr;
  return ew_launch(ctx, len, "logsigmoid_bwd_r", r ? 48 : 24, EwIn<4>{{X, r ? RX : nullptr, U, UR}}, EwOut<2>{{U, UR}},
                   [=] __device__(long long, const double* v, double* y) {
                     const double u = v[2], ex = exp(v[0]);
                     if (r) {
                       const double s = 1 / (1 + ex);
                       y[1] = v[3] * s - u * s * s * ex * v[1];
                       y[0] = u * s;
                     } else {
                       y[0] = u / (1 + ex);
                     }
                   });
}

// ---- Sin (math_funcs.go:511-597); Cos = Sin(AddScaler(x, pi/2)) on the host side --------------------
int af_sin_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "sin_fwd_r", RO ? 32 : 16, EwIn<2>{{X, RO ? RX : nullptr}}, EwOut<2>{{O, RO}},
                   [=] __device__(long long, const double* x, double* y) {
                     double s, c;
                     sincos(x[0], &s, &c);
                     y[0] = s;
                     y[1] = c * x[1];
                   });
}

int af_sin_bwd_r(af_ctx* ctx, const double* X, const double* RX, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(U);
  return ew_launch(ctx, len, "sin_bwd_r", UR ? 48 : 24, EwIn<4>{{X, UR ? RX : nullptr, U, UR}}, EwOut<2>{{U, UR}},
                   [=] __device__(long long, const double* v, double* y) {
                     double s, c;
                     sincos(v[0], &s, &c);
                     const double u = v[2];
                     y[1] = v[3] * c + u * (-s * v[1]);
                     y[0] = u * c;
                   });
}

// ---- Tanh: extension named by the north star, same contract as Sigmoid (T, RT kept by the caller) ----
int af_tanh_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* T, double* RT, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(T);
  return ew_launch(ctx, len, "tanh_fwd_r", RT ? 32 : 16, EwIn<2>{{X, RT ? RX : nullptr}}, EwOut<2>{{T, RT}},
                   [=] __device__(long long, const double* x, double* y) {
                     const double t = tanh(x[0]);
                     y[0] = t;
                     y[1] = (1 - t * t) * x[1];
                   });
}

int af_tanh_bwd_r(af_ctx* ctx, const double* T, const double* RT, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(T); AF_NEED(U);
  return ew_launch(ctx, len, "tanh_bwd_r", UR ? 48 : 24, EwIn<4>{{T, UR ? RT : nullptr, U, UR}}, EwOut<2>{{U, UR}},
                   [=] __device__(long long, const double* v, double* y) {
                     const double t = v[0], u = v[2];
                     y[1] = (1 - t * t) * v[3] - 2 * t * v[1] * u;
                     y[0] = (1 - t * t) * u;
                   });
}

// ---- Inverse (arithmetic.go:772-867) ------------------------------------------------------------------
int af_inverse_fwd_r(af_ctx* ctx, const double* X, const double* RX, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "inverse_fwd_r", RO ? 32 : 16, EwIn<2>{{X, RO ? RX : nullptr}}, EwOut<2>{{O, RO}},
                   [=] __device__(long long, const double* x, double* y) {
                     const double r = 1 / x[0];
                     y[0] = r;
                     y[1] = -(r * r) * x[1];
                   });
}

// O = the forward output 1/x
int af_inverse_bwd_r(af_ctx* ctx, const double* O, const double* RX, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(O); AF_NEED(U);
  return ew_launch(ctx, len, "inverse_bwd_r", UR ? 48 : 24, EwIn<4>{{O, UR ? RX : nullptr, U, UR}}, EwOut<2>{{U, UR}},
                   [=] __device__(long long, const double* v, double* y) {
                     const double o = v[0], u = v[2], sq = -(o * o);
                     y[1] = sq * v[3] + -2 * (sq * o) * v[1] * u;
                     y[0] = sq * u;
                   });
}

// ---- Pow (arithmetic.go:869-966) ----------------------------------------------------------------------
int af_pow_fwd_r(af_ctx* ctx, const double* X, const double* RX, double p, double* O, double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  const bool r = RO != nullptr;
  return ew_launch(ctx, len, "pow_fwd_r", r ? 32 : 16, EwIn<2>{{X, r ? RX : nullptr}}, EwOut<2>{{O, RO}},
                   [=] __device__(long long, const double* x, double* y) {
                     y[0] = pow(x[0], p);
                     if (r) y[1] = p != 0 ? p * pow(x[0], p - 1) * x[1] : 0.0;
                   });
}

int af_pow_bwd_r(af_ctx* ctx, const double* X, const double* RX, double p, double* U, double* UR, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(U);
  if (p == 1) return AF_OK;  // arithmetic.go:898,952: upstream passes through untouched
  const bool r = UR != nullptr;
  return ew_launch(ctx, len, "pow_bwd_r", r ? 48 : 24, EwIn<4>{{X, r ? RX : nullptr, U, UR}}, EwOut<2>{{U, UR}},
                   [=] __device__(long long, const double* v, double* y) {
                     const double x = v[0], u = v[2];
                     const double d1 = p * pow(x, p - 1);
                     if (r) y[1] = v[3] * d1 + u * (p * (p - 1) * pow(x, p - 2) * v[1]);
                     y[0] = u * d1;
                   });
}

// ---- Div (arithmetic.go:378-423; DivR is MulR(a, InverseR(b)) and needs no kernel of its own) --------
int af_div(af_ctx* ctx, const double* A, const double* B, double* O, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(A); AF_NEED(B); AF_NEED(O);
  return ew_launch(ctx, len, "div", 24, EwIn<2>{{A, B}}, EwOut<1>{{O}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = x[0] / x[1]; });
}

// DA = U / B (numerator's upstream), DB = U * (-O / B) (denominator's); either may be NULL
int af_div_bwd(af_ctx* ctx, const double* U, const double* O, const double* B, double* DA, double* DB, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(U); AF_NEED(O); AF_NEED(B);
  return ew_launch(ctx, len, "div_bwd", 40, EwIn<3>{{U, O, B}}, EwOut<2>{{DA, DB}},
                   [=] __device__(long long, const double* x, double* y) {
                     y[0] = x[0] / x[2];
                     y[1] = x[0] * (-x[1] / x[2]);
                   });
}

// ---- scalar-broadcast ops: AddScaler, ScaleFirst, AddFirst, SumAll backward ---------------------------
int af_add_scaler(af_ctx* ctx, const double* X, double f, double* O, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(O);
  return ew_launch(ctx, len, "add_scaler", 16, EwIn<1>{{X}}, EwOut<1>{{O}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = x[0] + f; });
}

// O = X + sign * f[0] with f on the device (AddFirst, arithmetic.go:595-612; the "- maxVal" of the log-domain sums)
int af_add_dev(af_ctx* ctx, const double* X, const double* d_f, double sign, double* O, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(d_f); AF_NEED(O);
  return ew_launch(ctx, len, "add_dev", 16, EwIn<1>{{X}}, EwOut<1>{{O}},
                   [=] __device__(long long, const double* x, double* y) { y[0] = x[0] + sign * d_f[0]; });
}

// O = X * f[0];  RO = X * fR[0] + RX * f[0]   (ScaleFirst forward and, applied to the upstreams, its backward:
// arithmetic.go:503-510,544-553,583-592).  O may alias X and RO may alias RX.
int af_scale_dev(af_ctx* ctx, const double* X, const double* RX, const double* d_f, const double* d_fR, double* O,
                 double* RO, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(d_f); AF_NEED(O);
  return ew_launch(ctx, len, "scale_dev", RO ? 32 : 16, EwIn<2>{{X, RO ? RX : nullptr}}, EwOut<2>{{O, RO}},
                   [=] __device__(long long, const double* x, double* y) {
                     const double f = d_f[0];
                     y[1] = x[0] * (d_fR ? d_fR[0] : 0.0) + x[1] * f;
                     y[0] = x[0] * f;
                   });
}

// x[i] = value[0]: SumAll's backward broadcast (arithmetic.go:987-995,1035-1047) without a host round trip
int af_fill_dev(af_ctx* ctx, double* x, const double* d_value, int64_t len) {
  AF_CHECK_CTX(ctx); AF_NEED(x); AF_NEED(d_value);
  return ew_launch(ctx, len, "fill_dev", 8, EwIn<1>{{nullptr}}, EwOut<1>{{x}},
                   [=] __device__(long long, const double*, double* y) { y[0] = d_value[0]; });
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
  return ew_launch(ctx, len, "norm_bwd_r", DR ? 32 : 16, EwIn<2>{{X, DR ? RX : nullptr}}, EwOut<2>{{D, DR}},
                   [=] __device__(long long, const double* x, double* y) {
                     const double o = d_o[0], u = d_u[0], s = u / o;
                     if (DR) y[1] = x[0] * (-u * d_ro[0] / (o * o) + d_uR[0] / o) + x[1] * s;
                     y[0] = x[0] * s;
                   });
}

// ---- Softmax, one per row of an n x len packed batch (math_funcs.go:479-509 under RFuncBatcher) --------
int af_softmax_fwd_r(af_ctx* ctx, const double* X, const double* RX, double temperature, double* Y, double* RY,
                     int64_t len, int64_t n) {
  AF_CHECK_CTX(ctx); AF_NEED(X); AF_NEED(Y);
  if (len < 0 || n < 0) return fail(AF_ESHAPE, "negative length");
  if (len == 0 || n == 0) return AF_OK;
  const double invT = (temperature != 0 && temperature != 1) ? 1 / temperature : 1.0;
  if (len <= kShortRow && !ctx->short_rows_off) {
    const ShortRowGeom g = short_row_geom(len, kShortChunk);
    const unsigned grid = (unsigned)((n + g.rows_per_block - 1) / g.rows_per_block);
    if (RY) softmax_fwd_short_kernel<true><<<grid, 256, 0, ctx->stream>>>(X, RX, Y, RY, invT, (long long)n, g);
    else softmax_fwd_short_kernel<false><<<grid, 256, 0, ctx->stream>>>(X, nullptr, Y, nullptr, invT, (long long)n, g);
    return after_launch(ctx, "softmax_fwd_short_kernel", KIND_REDUCE, 8.0 * len * n * (RY ? 4 : 2));
  }
  if (len <= kSoftChunk) {
    const int rpc = (int)(kSoftChunk / len);
    const unsigned grid = (unsigned)((n + rpc - 1) / rpc);
    const size_t smem = (size_t)(RY ? 2 : 1) * kSoftChunk * 8;
    AF_STAGED_DISPATCH(softmax_fwd_staged_kernel, smem, X, RX, Y, RY, invT, (int)len, (long long)n, rpc);
    return after_launch(ctx, "softmax_fwd_staged_kernel", KIND_REDUCE, 8.0 * len * n * (RY ? 4 : 2));
  }
  AF_ROW_DISPATCH(softmax_fwd_kernel, X, RX, Y, RY, invT, len, n);
  return after_launch(ctx, "softmax_fwd_kernel", KIND_REDUCE, 8.0 * len * n * (RY ? 4 : 2));
}

int af_softmax_bwd_r(af_ctx* ctx, const double* Y, const double* RY, double temperature, double* U, double* UR,
                     int64_t len, int64_t n) {
  AF_CHECK_CTX(ctx); AF_NEED(Y); AF_NEED(U);
  if (len < 0 || n < 0) return fail(AF_ESHAPE, "negative length");
  if (len == 0 || n == 0) return AF_OK;
  const double invT = (temperature != 0 && temperature != 1) ? 1 / temperature : 1.0;
  if (len <= kShortRow && !ctx->short_rows_off) {
    const ShortRowGeom g = short_row_geom(len, kShortChunkBwd);
    const unsigned grid = (unsigned)((n + g.rows_per_block - 1) / g.rows_per_block);
    if (UR) {
      softmax_bwd_short_kernel<true><<<grid, 256, 0, ctx->stream>>>(Y, RY, U, UR, invT, (long long)n, g);
    } else {
      softmax_bwd_short_kernel<false><<<grid, 256, 0, ctx->stream>>>(Y, nullptr, U, nullptr, invT, (long long)n, g);
    }
    return after_launch(ctx, "softmax_bwd_short_kernel", KIND_REDUCE, 8.0 * len * n * (UR ? 6 : 3));
  }
  if (len <= kSoftChunk) {
    const int rpc = (int)(kSoftChunk / len);
    const unsigned grid = (unsigned)((n + rpc - 1) / rpc);
    const size_t smem = (size_t)(UR ? 4 : 2) * kSoftChunk * 8;
    AF_STAGED_DISPATCH(softmax_bwd_staged_kernel, smem, Y, RY, U, UR, invT, (int)len, (long long)n, rpc);
    return after_launch(ctx, "softmax_bwd_staged_kernel", KIND_REDUCE, 8.0 * len * n * (UR ? 6 : 3));
  }
  AF_ROW_DISPATCH(softmax_bwd_kernel, Y, RY, U, UR, invT, len, n);
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
  return ew_launch(ctx, len, "sqerr_bwd_r", GR ? 48 : 24, EwIn<4>{{A, Y, GR ? RA : nullptr, GR ? RY : nullptr}},
                   EwOut<2>{{G, GR}}, [=] __device__(long long, const double* x, double* y) {
                     const double d = x[0] - x[1], u = d_u[0] * 0.5;
                     if (GR) y[1] = sign * (2 * ((d_uR[0] * 0.5) * d + u * (x[2] - x[3])));
                     y[0] = sign * (u * (d * 2));
                   });
}

}  // extern "C"
