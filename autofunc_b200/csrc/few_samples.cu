// LinTran with a FEW SAMPLES (n <= 16; the reference's own LinTranBenchmark runs n = 10, bench/lintran.go:12-16) as
// streaming tensor-core passes over the weight matrix.
//
// With a handful of samples every pass is bound by the HBM stream of the weight-sized operands, not by arithmetic:
//   forward  (lin_tran.go:64-77,119-177)    reads W, RW once;
//   backward (lin_tran.go:393-425: dataGradient(R) :179-270 + inputGradient(R) :272-359) reads W, RW once AND
//            read-modify-writes gW, rgW once -- ONE kernel for what linTranRResult.PropagateRGradient does, so that the
//            weight matrix is streamed once for both gradients instead of once per pass.
// The arithmetic rides along on the FP64 tensor pipe: the samples are the 8-wide M (forward, input gradient) or the
// 4-deep K (weight gradient) of DMMA.8x8x4 fragments, so a thread holds only a few accumulators and the rest of its
// registers carry loads in flight -- the streaming kernels of small_n.cu hold n partial sums per thread and ran at a third
// of the HBM rate for n = 10.  No operand passes through shared memory except the few-KB sample-side fragments.
//
// Work decomposition: one warp = one 8-row band (forward) or one 16-column strip (backward) of W; the other extent is
// split over blocks so that about one wave of blocks covers the chip.  Partial results of the split extent (forward: the
// K split; backward: the input gradient's row split) are parked in the context's scratch and summed IN SPLIT ORDER by the
// last block to arrive (ticket counter), so results are run-to-run deterministic and no output needs zeroing first.
#include <algorithm>
#include <atomic>
#include <type_traits>

#include "common.h"
#include "gemm_dmma.cuh"

namespace af {
namespace {

constexpr int kFsThreads = 128;   // 4 warps
constexpr int kFsMaxSamples = 16;

__device__ __forceinline__ double2 ld2g(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2g(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

constexpr int kRingStages = 3;
constexpr uint32_t kSlotBytes = 1024;                 // one 8 x 16 tile, or one 16 x 8 sample block
constexpr uint32_t kStageBytes = 6 * kSlotBytes;      // W RW gW rgW U UR
constexpr uint32_t kWarpRingBytes = kRingStages * kStageBytes;

// byte offset of element (r, c) of an 8-row x 16-column tile that TMA wrote with the 128-byte swizzle
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 128 + ((((c >> 1) ^ r) & 7) << 4) + ((c & 1) << 3)); }

// Last-block-folds protocol: every block of a fold group parks its partial result, fences, and takes a ticket; the block
// that draws the last ticket sums the group's partials in split order.  Returns true in every thread of that block.  The
// counter is handed back at zero, so the next launch (or graph replay) on the stream finds it ready.
__device__ __forceinline__ bool fold_ticket(unsigned* counter, unsigned splits) {
  __shared__ unsigned last;
  __syncthreads();  // this block's partial stores are issued
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(counter, 1u);
    last = (t == splits - 1) ? 1u : 0u;
    if (last) *counter = 0;
  }
  __syncthreads();
  if (last) __threadfence();
  return last != 0;
}

// ------------------------------------------------------------------------------------------------------------------
// forward:  Y[i][m] = sum_k X[i][k] W[m][k] ;  RY[i][m] = sum_k RX[i][k] W[m][k] + X[i][k] RW[m][k]
// Tile = 8 samples x 8 weight rows; the contraction runs over 8-column chunks of W, two DMMA k-steps per chunk: lane
// (g, j) loads the 16-byte pair W[m0 + g][k + 2j .. 2j + 1] -- the warp's load is 8 rows x 64 contiguous bytes -- and
// uses .x in k-step 0, .y in k-step 1; the X fragments are staged in shared memory in the same pairing.
// ------------------------------------------------------------------------------------------------------------------
struct FsFwdArgs {
  const double* W; const double* RW; const double* X; const double* RX;
  double* Y; double* RY;
  const double* b; const double* Rb;
  double* part; unsigned* tickets;
  int rows, cols, n, kc, ksplits, epi, dual, ppitch;
};

// (A variant that brought the W / RW tiles through warp-private TMA rings like few_samples_bwd_tma_kernel measured the
// same 18.5 us on C1: with n = 10 padded to two sample tiles the main loop is bound by the FP64 tensor pipe -- 12 DMMAs
// per 8 x 8 block of W, 5.4 us of pipe time -- and the rest is launch, fragment staging and the fold.)
template <int MT>
__global__ void __launch_bounds__(kFsThreads, 4) few_samples_fwd_kernel(const FsFwdArgs a) {
  extern __shared__ unsigned char fs_fwd_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
  const int chunks = a.kc >> 3;
  const int kbase = blockIdx.y * a.kc;
  const bool has_rx = a.dual && a.RX != nullptr, has_rw = a.dual && a.RW != nullptr;
  double2* sX = reinterpret_cast<double2*>(fs_fwd_raw);   // X fragments [chunk][mt][lane] | RX fragments
  double2* sRX = sX + chunks * MT * 32;

  const int m0 = blockIdx.x * 32 + 8 * warp;
  const int row = m0 + g;
  const bool row_ok = row < a.rows;
  const double* wp = a.W + (long long)row * a.cols + kbase + 2 * j;
  const double* rwp = has_rw ? a.RW + (long long)row * a.cols + kbase + 2 * j : nullptr;
  constexpr int UN = 4;  // chunks whose weight loads are issued together; two such groups are in flight per warp
  struct Group { double2 w[UN], rw[UN]; };
  auto load_group = [&](int c0, Group& t) {
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int c = c0 + u;
      const bool ok = row_ok && c < chunks && kbase + 8 * c + 2 * j < a.cols;
      t.w[u] = ok ? ld2g(wp + 8 * c) : make_double2(0.0, 0.0);
      t.rw[u] = (ok && has_rw) ? ld2g(rwp + 8 * c) : make_double2(0.0, 0.0);
    }
  };
  // the first weight loads leave before the X fragments are staged: the HBM round trip overlaps the staging
  Group cur;
  load_group(0, cur);

  // stage the X / RX fragments of this block's K range, four independent loads per thread at a time
  const int frag_total = chunks * MT * 32;
  for (int e0 = threadIdx.x; e0 < frag_total; e0 += 4 * kFsThreads) {
    double2 xv[4], rxv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = e0 + q * kFsThreads;
      const int l = e & 31, mt = (e >> 5) % MT, c = e / (32 * MT);
      const int i = 8 * mt + (l >> 2), k = kbase + 8 * c + 2 * (l & 3);
      const bool ok = e < frag_total && i < a.n && k < a.cols;
      xv[q] = ok ? ld2g(a.X + (long long)i * a.cols + k) : make_double2(0.0, 0.0);
      rxv[q] = (ok && has_rx) ? ld2g(a.RX + (long long)i * a.cols + k) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = e0 + q * kFsThreads;
      if (e < frag_total) {
        sX[e] = xv[q];
        if (has_rx) sRX[e] = rxv[q];
      }
    }
  }
  __syncthreads();

  // two accumulator sets (even / odd chunks) halve the dependent-DMMA chains; they are added once at the end
  double y[2][MT][2], ry[2][MT][2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int t = 0; t < MT; ++t) y[h][t][0] = y[h][t][1] = ry[h][t][0] = ry[h][t][1] = 0.0;

  auto compute_chunks = [&](int c0, Group& t, auto count) {
#pragma unroll
    for (int u = 0; u < decltype(count)::value; ++u) {
      const int c = c0 + u;
      if (c < chunks) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const double2 x = sX[(c * MT + mt) * 32 + lane];
          dmma_8x8x4(y[u & 1][mt][0], y[u & 1][mt][1], x.x, t.w[u].x);
          dmma_8x8x4(y[u & 1][mt][0], y[u & 1][mt][1], x.y, t.w[u].y);
          if (has_rw) {
            dmma_8x8x4(ry[u & 1][mt][0], ry[u & 1][mt][1], x.x, t.rw[u].x);
            dmma_8x8x4(ry[u & 1][mt][0], ry[u & 1][mt][1], x.y, t.rw[u].y);
          }
          if (has_rx) {
            const double2 rx = sRX[(c * MT + mt) * 32 + lane];
            dmma_8x8x4(ry[u & 1][mt][0], ry[u & 1][mt][1], rx.x, t.w[u].x);
            dmma_8x8x4(ry[u & 1][mt][0], ry[u & 1][mt][1], rx.y, t.w[u].y);
          }
        }
      }
    }
  };
  auto compute_group = [&](int c0, Group& t) { compute_chunks(c0, t, std::integral_constant<int, UN>{}); };
  // (no second group in flight: see the note in few_samples_bwd_kernel -- resident warps, not per-warp depth, set the rate)
  compute_group(0, cur);
  for (int c0 = UN; c0 < chunks; c0 += UN) {
    load_group(c0, cur);
    compute_group(c0, cur);
  }
#pragma unroll
  for (int t = 0; t < MT; ++t) {
    y[0][t][0] += y[1][t][0]; y[0][t][1] += y[1][t][1];
    ry[0][t][0] += ry[1][t][0]; ry[0][t][1] += ry[1][t][1];
  }

  // park the partial sums: part[split][which][sample][row], C fragment = (sample 8t + g, rows m0 + 2j, m0 + 2j + 1)
  const long long slab = (long long)(8 * MT) * a.ppitch;
  double* p0 = a.part + (long long)blockIdx.y * 2 * slab;
  if (m0 + 2 * j < a.rows) {
#pragma unroll
    for (int t = 0; t < MT; ++t) {
      const long long o = (long long)(8 * t + g) * a.ppitch + m0 + 2 * j;
      st2g(p0 + o, make_double2(y[0][t][0], y[0][t][1]));
      if (a.dual) st2g(p0 + slab + o, make_double2(ry[0][t][0], ry[0][t][1]));
    }
  }
  if (!fold_ticket(a.tickets + blockIdx.x, (unsigned)a.ksplits)) return;
  // the fold: element (sample i, row m) = sum over the K splits in split order.  Every thread's (up to four) elements are
  // folded together, eight splits' loads in flight at a time and no branches around the loads: the fold is a chain of L2
  // round trips on the kernel's critical path (the first version took one round trip per split and element: 9 us of 20)
  constexpr int ROUNDS = (8 * MT * 32) / kFsThreads;
  double v[ROUNDS], rv[ROUNDS];
  const double* pe[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int e = r * kFsThreads + threadIdx.x;
    const int i = min(e >> 5, a.n - 1), m = min(blockIdx.x * 32 + (e & 31), a.rows - 1);
    pe[r] = a.part + (long long)i * a.ppitch + m;
    v[r] = rv[r] = 0.0;
  }
  for (int s0 = 0; s0 < a.ksplits; s0 += 8) {
    double x[ROUNDS][8];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r)
#pragma unroll
      for (int q = 0; q < 8; ++q) x[r][q] = __ldcg(pe[r] + (long long)min(s0 + q, a.ksplits - 1) * 2 * slab);
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r)
#pragma unroll
      for (int q = 0; q < 8; ++q) v[r] += (s0 + q < a.ksplits) ? x[r][q] : 0.0;
    if (a.dual) {
#pragma unroll
      for (int r = 0; r < ROUNDS; ++r)
#pragma unroll
        for (int q = 0; q < 8; ++q) x[r][q] = __ldcg(pe[r] + (long long)min(s0 + q, a.ksplits - 1) * 2 * slab + slab);
#pragma unroll
      for (int r = 0; r < ROUNDS; ++r)
#pragma unroll
        for (int q = 0; q < 8; ++q) rv[r] += (s0 + q < a.ksplits) ? x[r][q] : 0.0;
    }
  }
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int e = r * kFsThreads + threadIdx.x;
    const int i = e >> 5, m = blockIdx.x * 32 + (e & 31);
    if (i >= a.n || m >= a.rows) continue;
    double y0 = v[r], y1 = rv[r];
    if (a.epi == EPI_BIAS || a.epi == EPI_BIAS_SIGMOID) {
      if (a.b) y0 += a.b[m];
      if (a.dual && a.Rb) y1 += a.Rb[m];
    }
    if (a.epi == EPI_BIAS_SIGMOID) {
      const double sg = sigmoid_f64(y0);
      y0 = sg;
      y1 = sg * (1 - sg) * y1;
    }
    const long long o = (long long)i * a.rows + m;
    if (a.Y) a.Y[o] = y0;
    if (a.dual && a.RY) a.RY[o] = y1;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward:  gW[m][k] += sum_i U[i][m] X[i][k]        rgW[m][k] += sum_i UR[i][m] X[i][k] + U[i][m] RX[i][k]
//            D[i][k]  = sum_m U[i][m] W[m][k]         DR[i][k]  = sum_m UR[i][m] W[m][k] + U[i][m] RW[m][k]
// A block owns 16 columns (two 8 x 8 tiles per 8-row band) and one row strip; its four warps take a quarter of the strip
// each and walk it band by band.
//   weight gradient: the gW tile IS the C fragment (lane (g, j): row m0 + g, columns k + 2j, k + 2j + 1, one 16-byte
//     load and store), A[g][j] = U[4 ks + j][m0 + g], B[j][g] = X[4 ks + j][k + g]: the samples are the contraction, in
//     NKS = ceil(n / 4) k-steps; the X fragments of the block's columns stay in registers for the whole strip;
//   input gradient: C = D[8 mt + g][k + 2j ..] stays in registers across the strip, A[g][j] = U[8 mt + g][m0 + 4 ks + j],
//     B[j][g] = W[m0 + 4 ks + j][k + g] -- the warp's two 8-byte loads per tile are 4 rows x 64 contiguous bytes each.
// The U-side fragments are 8-byte loads straight from global memory (U / UR are a few tens of KB: L1 / L2), prefetched one
// band ahead like the weight-sized operands.  The four warps' input-gradient fragments are summed through shared memory in
// warp order; only the per-block sums -- a few strips per column group -- go through the scratch and the ordered fold.
// (First version: four warps side by side on 64 columns sharing staged U fragments, 13 strips per column group -- the last
// block's fold of 13 x 2 partial tiles was 14 us of the kernel's 31.)
// ------------------------------------------------------------------------------------------------------------------
struct FsBwdArgs {
  const double* U; const double* UR; const double* W; const double* RW; const double* X; const double* RX;
  double* gW; double* rgW; double* D; double* DR;
  const double* S; const double* RS;   // EPI_SIGMOID_BWD on the input gradient (activations of the layer below), or NULL
  double* part; unsigned* tickets;
  int rows, cols, n, rs, strips, accumulate_d, dual;   // rs: rows per block (a multiple of 32)
};

template <int NKS>
__global__ void __launch_bounds__(kFsThreads, 3) few_samples_bwd_kernel(const FsBwdArgs a) {
  constexpr int MT = (NKS + 1) / 2;
  constexpr int NQ = 2 * MT * 2 * 2;               // input-gradient accumulators per lane: {D, DR} x mt x tile x pair
  __shared__ double sred[4 * NQ * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
  const bool do_w = a.gW != nullptr || a.rgW != nullptr, do_i = a.D != nullptr || a.DR != nullptr;
  const bool has_ur = a.dual && a.UR != nullptr, has_rw = a.dual && a.RW != nullptr, has_rx = a.dual && a.RX != nullptr;
  const bool want_rg = a.dual && a.rgW != nullptr, want_dr = a.dual && a.DR != nullptr;
  const int k0 = blockIdx.x * 16;
  const int wrows = a.rs >> 2;                      // rows per warp (a multiple of 8)
  const int row0 = blockIdx.y * a.rs + warp * wrows;
  int live = (min(a.rows, row0 + wrows) - row0 + 7) >> 3;   // bands of this warp that hold rows
  if (live < 0) live = 0;

  // X-side fragments of the weight gradient: B[j][g] = X[4 ks + j][k0 + 8u + g]
  double xb[2][NKS], rxb[2][NKS];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      const int i = 4 * ks + j, k = k0 + 8 * u + g;
      const bool ok = do_w && i < a.n && k < a.cols;
      xb[u][ks] = ok ? a.X[(long long)i * a.cols + k] : 0.0;
      rxb[u][ks] = (ok && has_rx) ? a.RX[(long long)i * a.cols + k] : 0.0;
    }
  double d[MT][2][2], dr[MT][2][2];
#pragma unroll
  for (int t = 0; t < MT; ++t)
#pragma unroll
    for (int u = 0; u < 2; ++u) d[t][u][0] = d[t][u][1] = dr[t][u][0] = dr[t][u][1] = 0.0;

  // one band's loads: W / RW in B-fragment order, the gW / rgW tiles as C fragments, the U / UR fragments of both products
  struct Band {
    double wv[2][2], rwv[2][2];
    double2 gv[2], rgv[2];
    double uw[NKS], urw[NKS];         // weight gradient A: U[4 ks + j][m0 + g]
    double ui[MT][2], uri[MT][2];     // input gradient A:  U[8 mt + g][m0 + 4 ks2 + j]
  };
  auto load_band = [&](int b, Band& t) {
    const int m0 = row0 + 8 * b;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int kb = k0 + 8 * u + g, kc = k0 + 8 * u + 2 * j;
#pragma unroll
      for (int ks2 = 0; ks2 < 2; ++ks2) {
        const int m = m0 + 4 * ks2 + j;
        const bool ok = do_i && m < a.rows && kb < a.cols;
        t.wv[u][ks2] = ok ? a.W[(long long)m * a.cols + kb] : 0.0;
        t.rwv[u][ks2] = (ok && has_rw) ? a.RW[(long long)m * a.cols + kb] : 0.0;
      }
      const bool okc = m0 + g < a.rows && kc < a.cols;
      const long long oc = (long long)(m0 + g) * a.cols + kc;
      t.gv[u] = (okc && a.gW) ? ld2g(a.gW + oc) : make_double2(0.0, 0.0);
      t.rgv[u] = (okc && want_rg) ? ld2g(a.rgW + oc) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      const int i = 4 * ks + j, m = m0 + g;
      const bool ok = do_w && i < a.n && m < a.rows;
      t.uw[ks] = ok ? a.U[(long long)i * a.rows + m] : 0.0;
      t.urw[ks] = (ok && has_ur) ? a.UR[(long long)i * a.rows + m] : 0.0;
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int ks2 = 0; ks2 < 2; ++ks2) {
        const int i = 8 * mt + g, m = m0 + 4 * ks2 + j;
        const bool ok = do_i && i < a.n && m < a.rows;
        t.ui[mt][ks2] = ok ? a.U[(long long)i * a.rows + m] : 0.0;
        t.uri[mt][ks2] = (ok && has_ur) ? a.UR[(long long)i * a.rows + m] : 0.0;
      }
  };
  auto compute_band = [&](int b, Band& t) {
    const int m0 = row0 + 8 * b;
    // ---- weight gradient: the samples are the contraction ----
    if (do_w) {
#pragma unroll
      for (int ks = 0; ks < NKS; ++ks) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          dmma_8x8x4(t.gv[u].x, t.gv[u].y, t.uw[ks], xb[u][ks]);
          if (has_rx) dmma_8x8x4(t.rgv[u].x, t.rgv[u].y, t.uw[ks], rxb[u][ks]);
        }
        if (has_ur) {
#pragma unroll
          for (int u = 0; u < 2; ++u) dmma_8x8x4(t.rgv[u].x, t.rgv[u].y, t.urw[ks], xb[u][ks]);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int kc = k0 + 8 * u + 2 * j;
        if (m0 + g < a.rows && kc < a.cols) {
          const long long oc = (long long)(m0 + g) * a.cols + kc;
          if (a.gW) st2g(a.gW + oc, t.gv[u]);
          if (want_rg) st2g(a.rgW + oc, t.rgv[u]);
        }
      }
    }
    // ---- input gradient: the band's 8 rows are two k-steps ----
    if (do_i) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks2 = 0; ks2 < 2; ++ks2) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            dmma_8x8x4(d[mt][u][0], d[mt][u][1], t.ui[mt][ks2], t.wv[u][ks2]);
            if (has_rw) dmma_8x8x4(dr[mt][u][0], dr[mt][u][1], t.ui[mt][ks2], t.rwv[u][ks2]);
          }
          if (has_ur) {
#pragma unroll
            for (int u = 0; u < 2; ++u) dmma_8x8x4(dr[mt][u][0], dr[mt][u][1], t.uri[mt][ks2], t.wv[u][ks2]);
          }
        }
    }
  };
  // One band at a time, no software prefetch: measured on C1 (profiles/r02_few_samples.txt), a second band in flight
  // changed nothing (37.5 vs 38.1 us) while its registers halve the resident warps -- and with a few-KB stream per warp
  // it is the NUMBER of warps with loads outstanding that sets the rate (micro-benchmark tools/micro/bw_pattern.cu: the
  // same 33.5 MB read takes 13 us with 4096 warps, 20 us with 1024, whatever the access pattern).
  {
    Band t;
    for (int b = 0; b < live; ++b) {
      load_band(b, t);
      compute_band(b, t);
    }
  }
  if (!do_i) return;

  // the four warps' fragments, summed in warp order: lane-major so that the sums keep the fragment layout
#pragma unroll
  for (int t = 0; t < MT; ++t)
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sred[(warp * NQ + ((0 * MT + t) * 2 + u) * 2 + e) * 32 + lane] = d[t][u][e];
        sred[(warp * NQ + ((1 * MT + t) * 2 + u) * 2 + e) * 32 + lane] = dr[t][u][e];
      }
  __syncthreads();
  const int colgroups = gridDim.x;
  double* mine = a.part + ((long long)blockIdx.y * colgroups + blockIdx.x) * (NQ * 32);
  const long long strip_stride = (long long)colgroups * (NQ * 32);
  const bool single = a.strips == 1;
  double total[NQ * 32 / kFsThreads];
#pragma unroll
  for (int r = 0; r < NQ * 32 / kFsThreads; ++r) {
    const int e = r * kFsThreads + threadIdx.x;
    total[r] = ((sred[e] + sred[NQ * 32 + e]) + sred[2 * NQ * 32 + e]) + sred[3 * NQ * 32 + e];
    if (!single) mine[e] = total[r];
  }
  if (!single) {
    if (!fold_ticket(a.tickets + blockIdx.x, (unsigned)a.strips)) return;
    const double* p = a.part + (long long)blockIdx.x * (NQ * 32);
#pragma unroll
    for (int r = 0; r < NQ * 32 / kFsThreads; ++r) total[r] = 0.0;
    for (int s0 = 0; s0 < a.strips; s0 += 4) {   // strip order, four strips' loads in flight at a time
      double x[NQ * 32 / kFsThreads][4];
#pragma unroll
      for (int r = 0; r < NQ * 32 / kFsThreads; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int s = min(s0 + q, a.strips - 1);
          x[r][q] = __ldcg(p + (long long)s * strip_stride + r * kFsThreads + threadIdx.x);
        }
#pragma unroll
      for (int r = 0; r < NQ * 32 / kFsThreads; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) total[r] += (s0 + q < a.strips) ? x[r][q] : 0.0;
    }
  }
  // fragment element e = (q, lane): q = ((which * MT + mt) * 2 + u) * 2 + pair  ->  sample 8 mt + g, column k0 + 8u + 2j + pair.
  // D and DR of one element sit NQ / 2 * 32 apart: exchange them through shared memory so that one thread finishes both.
  __syncthreads();
#pragma unroll
  for (int r = 0; r < NQ * 32 / kFsThreads; ++r) sred[r * kFsThreads + threadIdx.x] = total[r];
  __syncthreads();
  for (int e = threadIdx.x; e < (NQ / 2) * 32; e += kFsThreads) {
    const int l = e & 31, q = e >> 5;
    const int pair = q & 1, u = (q >> 1) & 1, mt = q >> 2;
    const int i = 8 * mt + (l >> 2), k = k0 + 8 * u + 2 * (l & 3) + pair;
    if (i >= a.n || k >= a.cols) continue;
    double v = sred[e], rv = sred[(NQ / 2) * 32 + e];
    const long long o = (long long)i * a.cols + k;
    if (a.S) {  // math_funcs.go:365-371, operand order kept: u' = x(1-x)u ; uR' = xR(1-x)u - x xR u + x(1-x)uR
      const double x = a.S[o];
      const double xr = (want_dr && a.RS) ? a.RS[o] : 0.0;
      const double uu = v;
      v = x * (1 - x) * uu;
      rv = xr * (1 - x) * uu - x * xr * uu + x * (1 - x) * rv;
    }
    if (a.D) a.D[o] = a.accumulate_d ? a.D[o] + v : v;
    if (want_dr) a.DR[o] = a.accumulate_d ? a.DR[o] + rv : rv;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward, TMA form (the default where its alignment rules hold: even rows and columns).  Same decomposition and the
// same arithmetic as few_samples_bwd_kernel above, but every operand of a band -- the W / RW / gW / rgW tiles (8 rows x
// 16 columns, 128-byte swizzle) and the U / UR fragments' source (16 samples x 8 rows) -- is brought into a WARP-PRIVATE
// ring of shared-memory stages by bulk tensor copies that lane 0 issues kStages bands ahead.  Data in flight costs no
// registers, so twelve warps per SM each keep 18 KB outstanding, where the register version had one band (6 KB) per warp
// and measured latency-bound (resident warps set its rate: 63 / 50 / 43 us for 4 / 8 / 12 warps per SM on C1).
// No block-level synchronisation until the final reduction: a warp's ring and its mbarriers are its own.
// ------------------------------------------------------------------------------------------------------------------
template <int NKS>
__global__ void __launch_bounds__(kFsThreads, 3)
few_samples_bwd_tma_kernel(const __grid_constant__ CUtensorMap mW, const __grid_constant__ CUtensorMap mRW,
                           const __grid_constant__ CUtensorMap mG, const __grid_constant__ CUtensorMap mRG,
                           const __grid_constant__ CUtensorMap mU, const __grid_constant__ CUtensorMap mUR, const FsBwdArgs a) {
  constexpr int MT = (NKS + 1) / 2;
  constexpr int NQ = 2 * MT * 2 * 2;
  extern __shared__ unsigned char fs_ring_raw[];
  __shared__ __align__(8) unsigned long long fs_bars[4 * kRingStages];
  // swizzle atoms must sit on 1024-byte boundaries of the shared address space: the launch reserves 1 KB of slack
  unsigned char* fs_ring = fs_ring_raw + ((1024u - (smem_u32(fs_ring_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
  const bool do_w = a.gW != nullptr || a.rgW != nullptr, do_i = a.D != nullptr || a.DR != nullptr;
  const bool has_ur = a.dual && a.UR != nullptr, has_rw = a.dual && a.RW != nullptr, has_rx = a.dual && a.RX != nullptr;
  const bool want_rg = a.dual && a.rgW != nullptr, want_dr = a.dual && a.DR != nullptr;
  const int k0 = blockIdx.x * 16;
  const int wrows = a.rs >> 2;
  const int row0 = blockIdx.y * a.rs + warp * wrows;
  int live = (min(a.rows, row0 + wrows) - row0 + 7) >> 3;
  if (live < 0) live = 0;
  const uint32_t ring = smem_u32(fs_ring) + warp * kWarpRingBytes;
  const uint32_t bars = smem_u32(fs_bars) + warp * kRingStages * 8;
  const uint32_t tx = kSlotBytes * ((do_i ? 1 : 0) + ((do_i && has_rw) ? 1 : 0) + (a.gW ? 1 : 0) + (want_rg ? 1 : 0) + 1 + (has_ur ? 1 : 0));

  auto issue = [&](int b, int st) {   // lane 0 only
    const uint32_t dst = ring + st * kStageBytes, bar = bars + 8 * st;
    const int m0 = row0 + 8 * b;
    mbar_expect_tx(bar, tx);
    if (do_i) tma_load_2d(dst, &mW, bar, k0, m0);
    if (do_i && has_rw) tma_load_2d(dst + kSlotBytes, &mRW, bar, k0, m0);
    if (a.gW) tma_load_2d(dst + 2 * kSlotBytes, &mG, bar, k0, m0);
    if (want_rg) tma_load_2d(dst + 3 * kSlotBytes, &mRG, bar, k0, m0);
    tma_load_2d(dst + 4 * kSlotBytes, &mU, bar, m0, 0);
    if (has_ur) tma_load_2d(dst + 5 * kSlotBytes, &mUR, bar, m0, 0);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kRingStages; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
    for (int s = 0; s < kRingStages; ++s)
      if (s < live) issue(s, s);
  }
  __syncwarp();

  // X-side fragments of the weight gradient: B[j][g] = X[4 ks + j][k0 + 8u + g]
  double xb[2][NKS], rxb[2][NKS];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      const int i = 4 * ks + j, k = k0 + 8 * u + g;
      const bool ok = do_w && i < a.n && k < a.cols;
      xb[u][ks] = ok ? a.X[(long long)i * a.cols + k] : 0.0;
      rxb[u][ks] = (ok && has_rx) ? a.RX[(long long)i * a.cols + k] : 0.0;
    }
  double d[MT][2][2], dr[MT][2][2];
#pragma unroll
  for (int t = 0; t < MT; ++t)
#pragma unroll
    for (int u = 0; u < 2; ++u) d[t][u][0] = d[t][u][1] = dr[t][u][0] = dr[t][u][1] = 0.0;

  // per-lane fragment offsets inside a stage (loop-invariant)
  uint32_t off_b[2][2], off_c[2], off_uw[NKS], off_ui[MT][2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
#pragma unroll
    for (int ks2 = 0; ks2 < 2; ++ks2) off_b[u][ks2] = tile_off(4 * ks2 + j, 8 * u + g);   // B[j][g] = W[m0 + 4 ks2 + j][k0 + 8u + g]
    off_c[u] = tile_off(g, 8 * u + 2 * j);                                               // C[g][2j..] = gW[m0 + g][k0 + 8u + 2j ..]
  }
#pragma unroll
  for (int ks = 0; ks < NKS; ++ks) off_uw[ks] = (uint32_t)((4 * ks + j) * 64 + g * 8);    // A[g][j] = U[4 ks + j][m0 + g]
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int ks2 = 0; ks2 < 2; ++ks2) off_ui[mt][ks2] = (uint32_t)((8 * mt + g) * 64 + (4 * ks2 + j) * 8);  // A[g][j] = U[8 mt + g][m0 + 4 ks2 + j]

  for (int b = 0; b < live; ++b) {
    const int st = b % kRingStages;
    const uint32_t base = ring + st * kStageBytes;
    mbar_wait(bars + 8 * st, (uint32_t)((b / kRingStages) & 1));
    const int m0 = row0 + 8 * b;
    // ---- weight gradient: the samples are the contraction ----
    if (do_w) {
      double2 gv[2], rgv[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        gv[u] = rgv[u] = make_double2(0.0, 0.0);
        if (a.gW) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(gv[u].x), "=d"(gv[u].y) : "r"(base + 2 * kSlotBytes + off_c[u]));
        if (want_rg) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rgv[u].x), "=d"(rgv[u].y) : "r"(base + 3 * kSlotBytes + off_c[u]));
      }
#pragma unroll
      for (int ks = 0; ks < NKS; ++ks) {
        const double au = lds_f64(base + 4 * kSlotBytes + off_uw[ks]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          dmma_8x8x4(gv[u].x, gv[u].y, au, xb[u][ks]);
          if (has_rx) dmma_8x8x4(rgv[u].x, rgv[u].y, au, rxb[u][ks]);
        }
        if (has_ur) {
          const double aur = lds_f64(base + 5 * kSlotBytes + off_uw[ks]);
#pragma unroll
          for (int u = 0; u < 2; ++u) dmma_8x8x4(rgv[u].x, rgv[u].y, aur, xb[u][ks]);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int kc = k0 + 8 * u + 2 * j;
        if (m0 + g < a.rows && kc < a.cols) {
          const long long oc = (long long)(m0 + g) * a.cols + kc;
          if (a.gW) st2g(a.gW + oc, gv[u]);
          if (want_rg) st2g(a.rgW + oc, rgv[u]);
        }
      }
    }
    // ---- input gradient: the band's 8 rows are two k-steps ----
    if (do_i) {
#pragma unroll
      for (int ks2 = 0; ks2 < 2; ++ks2) {
        double wv[2], rwv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          wv[u] = lds_f64(base + off_b[u][ks2]);
          rwv[u] = has_rw ? lds_f64(base + kSlotBytes + off_b[u][ks2]) : 0.0;
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const double au = lds_f64(base + 4 * kSlotBytes + off_ui[mt][ks2]);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            dmma_8x8x4(d[mt][u][0], d[mt][u][1], au, wv[u]);
            if (has_rw) dmma_8x8x4(dr[mt][u][0], dr[mt][u][1], au, rwv[u]);
          }
          if (has_ur) {
            const double aur = lds_f64(base + 5 * kSlotBytes + off_ui[mt][ks2]);
#pragma unroll
            for (int u = 0; u < 2; ++u) dmma_8x8x4(dr[mt][u][0], dr[mt][u][1], aur, wv[u]);
          }
        }
      }
    }
    // every ld.shared of this stage has been consumed by an issued DMMA (or store): hand the slot back to the copy engine
    __syncwarp();
    if (lane == 0 && b + kRingStages < live) issue(b + kRingStages, st);
  }
  if (!do_i) return;

  // the four warps' fragments, summed in warp order; each warp parks its own in its (now idle) ring
  double* mysum = reinterpret_cast<double*>(fs_ring + warp * kWarpRingBytes);
#pragma unroll
  for (int t = 0; t < MT; ++t)
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        mysum[(((0 * MT + t) * 2 + u) * 2 + e) * 32 + lane] = d[t][u][e];
        mysum[(((1 * MT + t) * 2 + u) * 2 + e) * 32 + lane] = dr[t][u][e];
      }
  __syncthreads();
  const int colgroups = gridDim.x;
  double* mine = a.part + ((long long)blockIdx.y * colgroups + blockIdx.x) * (NQ * 32);
  const long long strip_stride = (long long)colgroups * (NQ * 32);
  const bool single = a.strips == 1;
  constexpr int PER = NQ * 32 / kFsThreads;
  double total[PER];
#pragma unroll
  for (int r = 0; r < PER; ++r) {
    const int e = r * kFsThreads + threadIdx.x;
    const double* w0 = reinterpret_cast<const double*>(fs_ring);
    constexpr int WS = kWarpRingBytes / 8;
    total[r] = ((w0[e] + w0[WS + e]) + w0[2 * WS + e]) + w0[3 * WS + e];
    if (!single) mine[e] = total[r];
  }
  if (!single) {
    if (!fold_ticket(a.tickets + blockIdx.x, (unsigned)a.strips)) return;
    const double* p = a.part + (long long)blockIdx.x * (NQ * 32);
#pragma unroll
    for (int r = 0; r < PER; ++r) total[r] = 0.0;
    for (int s0 = 0; s0 < a.strips; s0 += 4) {   // strip order, four strips' loads in flight at a time
      double x[PER][4];
#pragma unroll
      for (int r = 0; r < PER; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) x[r][q] = __ldcg(p + (long long)min(s0 + q, a.strips - 1) * strip_stride + r * kFsThreads + threadIdx.x);
#pragma unroll
      for (int r = 0; r < PER; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) total[r] += (s0 + q < a.strips) ? x[r][q] : 0.0;
    }
  }
  // D and DR of one element sit NQ / 2 * 32 apart: exchange through shared memory so that one thread finishes both
  __syncthreads();
  double* sred = reinterpret_cast<double*>(fs_ring);
#pragma unroll
  for (int r = 0; r < PER; ++r) sred[r * kFsThreads + threadIdx.x] = total[r];
  __syncthreads();
  for (int e = threadIdx.x; e < (NQ / 2) * 32; e += kFsThreads) {
    const int l = e & 31, q = e >> 5;
    const int pair = q & 1, u = (q >> 1) & 1, mt = q >> 2;
    const int i = 8 * mt + (l >> 2), k = k0 + 8 * u + 2 * (l & 3) + pair;
    if (i >= a.n || k >= a.cols) continue;
    double v = sred[e], rv = sred[(NQ / 2) * 32 + e];
    const long long o = (long long)i * a.cols + k;
    if (a.S) {  // math_funcs.go:365-371, operand order kept
      const double x = a.S[o];
      const double xr = (want_dr && a.RS) ? a.RS[o] : 0.0;
      const double uu = v;
      v = x * (1 - x) * uu;
      rv = xr * (1 - x) * uu - x * xr * uu + x * (1 - x) * rv;
    }
    if (a.D) a.D[o] = a.accumulate_d ? a.D[o] + v : v;
    if (want_dr) a.DR[o] = a.accumulate_d ? a.DR[o] + rv : rv;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient of a layer with FEW OUTPUT ROWS and a large batch (the bench network's 10-class head: gW[10][512] +=
// U^T X over 4096 samples; lin_tran.go:179-270).  The contraction runs over the batch, the output is tiny: as
// tensor-core tiles it needs split-K (atomics) or, in deterministic mode, eight whole-K tiles on eight CTAs (207 us on
// the bench step).  Here a block owns 64 columns and a chunk of samples; its eight warps take a slice of the chunk each
// (lane = column pair, rows x 2 x {g, rg} accumulators in registers, the X / RX stream coalesced), add their sums in warp
// order through shared memory, and the chunks meet in the ordered last-block fold: deterministic, no atomics, no zeroing.
// ------------------------------------------------------------------------------------------------------------------
struct FrWgradArgs {
  const double* U; const double* UR; const double* X; const double* RX;
  double* gW; double* rgW;
  double* part; unsigned* tickets;
  int rows, cols, chunks, dual;
  long long n, per_chunk;   // samples per chunk (a multiple of the 64-sample tile)
};

constexpr int kFrwTile = 64;  // samples per staged tile: eight per warp

template <int MMAX>
__global__ void __launch_bounds__(256, 1) few_rows_wgrad_kernel(const FrWgradArgs a) {
  constexpr int NQ = 2 * MMAX * 2;                 // {g, rg} x row x pair
  __shared__ double buf[NQ * 32];
  // the tile's upstream rows U / UR (64 x rows values each) go through shared memory: read straight from global -- one
  // broadcast load per (sample, row), each consumed at once -- the kernel was a chain of L1 / L2 round trips (ncu: 20
  // cycles per issued instruction, long_scoreboard; 63.6 us on the bench head)
  __shared__ double su[kFrwTile * MMAX], sur[kFrwTile * MMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 64 + 2 * lane;
  const bool col_ok = c < a.cols;
  const bool has_ur = a.dual && a.UR != nullptr, has_rx = a.dual && a.RX != nullptr;
  const long long c0 = (long long)blockIdx.y * a.per_chunk;            // per_chunk is a multiple of kFrwTile
  const long long c1 = min(c0 + a.per_chunk, a.n);
  double g[MMAX][2], rg[MMAX][2];
#pragma unroll
  for (int m = 0; m < MMAX; ++m) g[m][0] = g[m][1] = rg[m][0] = rg[m][1] = 0.0;
  constexpr int SPW = kFrwTile / 8;           // samples per warp and tile, all of whose X / RX loads are in flight together
  for (long long t0 = c0; t0 < c1; t0 += kFrwTile) {
    const long long s0 = t0 + warp * SPW;
    double2 x[SPW], rx[SPW];
#pragma unroll
    for (int q = 0; q < SPW; ++q) {
      const bool ok = col_ok && s0 + q < c1;
      x[q] = ok ? ld2g(a.X + (s0 + q) * a.cols + c) : make_double2(0.0, 0.0);
      rx[q] = (ok && has_rx) ? ld2g(a.RX + (s0 + q) * a.cols + c) : make_double2(0.0, 0.0);
    }
    __syncthreads();  // the previous tile's upstream rows have been consumed
    const int live = (int)min((long long)kFrwTile, c1 - t0) * a.rows;   // U[t0 .. t0 + 64) is one contiguous run
    for (int e = threadIdx.x; e < kFrwTile * a.rows; e += 256) {
      su[e] = e < live ? a.U[t0 * a.rows + e] : 0.0;
      sur[e] = (e < live && has_ur) ? a.UR[t0 * a.rows + e] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < SPW; ++q) {
      const double* u = su + (warp * SPW + q) * a.rows;   // (samples past the end were staged as zeros)
      const double* ur = sur + (warp * SPW + q) * a.rows;
#pragma unroll
      for (int m = 0; m < MMAX; ++m) {
        if (m < a.rows) {
          const double um = u[m];                      // the same address in every lane: a broadcast
          g[m][0] = fma(um, x[q].x, g[m][0]); g[m][1] = fma(um, x[q].y, g[m][1]);
          if (a.dual) {
            const double urm = ur[m];
            rg[m][0] = fma(urm, x[q].x, fma(um, rx[q].x, rg[m][0]));
            rg[m][1] = fma(urm, x[q].y, fma(um, rx[q].y, rg[m][1]));
          }
        }
      }
    }
  }
  // the eight warps' sums, added in warp order
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int m = 0; m < MMAX; ++m)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int q0 = ((0 * MMAX + m) * 2 + e) * 32 + lane, q1 = ((1 * MMAX + m) * 2 + e) * 32 + lane;
          buf[q0] = (w == 0 ? 0.0 : buf[q0]) + g[m][e];
          buf[q1] = (w == 0 ? 0.0 : buf[q1]) + rg[m][e];
        }
    }
    __syncthreads();
  }
  constexpr int PER = NQ * 32 / 256;
  const int colblocks = gridDim.x;
  double total[PER];
#pragma unroll
  for (int r = 0; r < PER; ++r) total[r] = buf[r * 256 + threadIdx.x];
  if (a.chunks > 1) {
    double* mine = a.part + ((long long)blockIdx.y * colblocks + blockIdx.x) * (NQ * 32);
#pragma unroll
    for (int r = 0; r < PER; ++r) mine[r * 256 + threadIdx.x] = total[r];
    if (!fold_ticket(a.tickets + blockIdx.x, (unsigned)a.chunks)) return;
    const double* p = a.part + (long long)blockIdx.x * (NQ * 32);
    const long long stride = (long long)colblocks * (NQ * 32);
#pragma unroll
    for (int r = 0; r < PER; ++r) total[r] = 0.0;
    for (int c0 = 0; c0 < a.chunks; c0 += 4) {   // chunk order, four chunks' loads in flight per element
      double x[PER][4];
#pragma unroll
      for (int r = 0; r < PER; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) x[r][q] = __ldcg(p + (long long)min(c0 + q, a.chunks - 1) * stride + r * 256 + threadIdx.x);
#pragma unroll
      for (int r = 0; r < PER; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) total[r] += (c0 + q < a.chunks) ? x[r][q] : 0.0;
    }
  }
  // element (q, lane): q = ((which * MMAX + m) * 2 + pair)  ->  gW / rgW [m][blockIdx.x * 64 + 2 lane + pair]
#pragma unroll
  for (int r = 0; r < PER; ++r) {
    const int e = r * 256 + threadIdx.x;
    const int l = e & 31, q = e >> 5;
    const int pair = q & 1, m = (q >> 1) % MMAX, which = q / (2 * MMAX);
    const int k = blockIdx.x * 64 + 2 * l + pair;
    if (m >= a.rows || k >= a.cols) continue;
    double* out = which ? a.rgW : a.gW;
    if (out) out[(long long)m * a.cols + k] += total[r];
  }
}

bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// resident blocks per SM for a kernel at a given dynamic shared-memory size (queried once per size class)
template <typename Kern>
int blocks_per_sm(Kern kern, int smem) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kFsThreads, (size_t)smem) != cudaSuccess || nb < 1) nb = 1;
  return nb;
}

template <typename Kern>
int allow_smem(af_ctx* ctx, Kern kern, std::atomic<unsigned long long>& configured) {
  if (!(configured.load(std::memory_order_acquire) >> (ctx->device & 63) & 1ull)) {
    AF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured.fetch_or(1ull << (ctx->device & 63), std::memory_order_release);
  }
  return AF_OK;
}

template <int MT>
int launch_fwd(af_ctx* ctx, FsFwdArgs& a) {
  auto kern = few_samples_fwd_kernel<MT>;
  static std::atomic<unsigned long long> configured{0};
  AF_TRY(allow_smem(ctx, kern, configured));
  const long long rowblocks = (a.rows + 31) / 32;
  const long long slab2 = 2ll * 8 * MT * a.ppitch;          // doubles per K split in the scratch
  // about one wave of blocks: K is split so that rowblocks x ksplits fills the resident slots
  // (three blocks = twelve warps per SM: more K splits only shorten every warp's stream and lengthen the fold)
  const int per_sm = std::min(ctx->few_samples_blocks, blocks_per_sm(kern, 40 * 1024));
  long long ksplits = ((long long)per_sm * ctx->sm_count) / rowblocks;
  const long long max_by_k = (a.cols + 63) / 64, max_by_scratch = ctx->scratch_doubles / slab2;
  if (ksplits > max_by_k) ksplits = max_by_k;
  if (ksplits > max_by_scratch) ksplits = max_by_scratch;
  if (ksplits < 1) ksplits = 1;
  long long kc = ((a.cols + ksplits - 1) / ksplits + 7) / 8 * 8;
  const long long kc_max = (96 * 1024) / (2 * MT * 32 * 16) * 8;  // shared-memory limit of the staged X / RX fragments
  if (kc > kc_max) kc = kc_max;
  ksplits = (a.cols + kc - 1) / kc;
  if (ksplits > max_by_scratch) return AF_ENOTSUP;
  a.kc = (int)kc; a.ksplits = (int)ksplits;
  const size_t smem = (size_t)(kc / 8) * MT * 32 * 16 * 2;
  kern<<<dim3((unsigned)rowblocks, (unsigned)ksplits), kFsThreads, smem, ctx->stream>>>(a);
  return after_launch(ctx, "few_samples_fwd_kernel", KIND_SMALL_N, 8.0 * a.rows * a.cols * ((a.dual && a.RW) ? 2 : 1));
}

template <int NKS>
int launch_bwd(af_ctx* ctx, FsBwdArgs& a) {
  constexpr int MT = (NKS + 1) / 2;
  constexpr int NQ = 2 * MT * 2 * 2;
  auto kern = few_samples_bwd_kernel<NKS>;
  const long long colgroups = (a.cols + 15) / 16;
  const bool do_i = a.D || a.DR;
  // about one wave of blocks: the rows are split so that colgroups x strips fills the resident slots
  const int per_sm = std::min(ctx->few_samples_blocks, blocks_per_sm(kern, 0));
  long long strips = ((long long)per_sm * ctx->sm_count) / colgroups;
  const long long max_by_rows = (a.rows + 63) / 64, max_by_scratch = ctx->scratch_doubles / (colgroups * NQ * 32);
  if (strips > max_by_rows) strips = max_by_rows;
  if (strips < 1) strips = 1;
  if (do_i && strips > 1 && strips > max_by_scratch) strips = max_by_scratch < 1 ? 1 : max_by_scratch;
  const long long rs = ((a.rows + strips - 1) / strips + 31) / 32 * 32;
  strips = (a.rows + rs - 1) / rs;
  a.rs = (int)rs; a.strips = (int)strips;
  kern<<<dim3((unsigned)colgroups, (unsigned)strips), kFsThreads, 0, ctx->stream>>>(a);
  const double bytes = 8.0 * a.rows * a.cols * ((do_i ? ((a.dual && a.RW) ? 2 : 1) : 0) + 2 * ((a.gW ? 1 : 0) + (a.rgW ? 1 : 0)));
  return after_launch(ctx, "few_samples_bwd_kernel", KIND_SMALL_N, bytes);
}

template <int NKS>
int launch_bwd_tma(af_ctx* ctx, FsBwdArgs& a) {
  constexpr int MT = (NKS + 1) / 2;
  constexpr int NQ = 2 * MT * 2 * 2;
  auto kern = few_samples_bwd_tma_kernel<NKS>;
  const int smem = (int)(4 * kWarpRingBytes) + 1024;
  static std::atomic<unsigned long long> configured{0};
  AF_TRY(allow_smem(ctx, kern, configured));
  const long long colgroups = (a.cols + 15) / 16;
  const bool do_i = a.D || a.DR;
  const int per_sm = std::min(ctx->few_samples_blocks, blocks_per_sm(kern, smem));
  long long strips = ((long long)per_sm * ctx->sm_count) / colgroups;
  const long long max_by_rows = (a.rows + 63) / 64, max_by_scratch = ctx->scratch_doubles / (colgroups * NQ * 32);
  if (strips > max_by_rows) strips = max_by_rows;
  if (strips < 1) strips = 1;
  if (do_i && strips > 1 && strips > max_by_scratch) strips = max_by_scratch < 1 ? 1 : max_by_scratch;
  const long long rs = ((a.rows + strips - 1) / strips + 31) / 32 * 32;
  strips = (a.rows + rs - 1) / rs;
  a.rs = (int)rs; a.strips = (int)strips;
  CUtensorMap maps[6];
  const double* w_like[4] = {a.W, a.RW, a.gW, a.rgW};
  const double* any = a.W ? a.W : (a.gW ? a.gW : a.rgW);
  for (int q = 0; q < 4; ++q) AF_TRY(gemm_make_map(ctx, &maps[q], w_like[q] ? w_like[q] : any, LAY_KC, a.rows, a.cols, a.cols, 8));
  AF_TRY(gemm_make_plain_map(ctx, &maps[4], a.U, a.n, a.rows, a.rows, 16, 8));
  AF_TRY(gemm_make_plain_map(ctx, &maps[5], a.UR ? a.UR : a.U, a.n, a.rows, a.rows, 16, 8));
  kern<<<dim3((unsigned)colgroups, (unsigned)strips), kFsThreads, (size_t)smem, ctx->stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4],
                                                                                             maps[5], a);
  const double bytes = 8.0 * a.rows * a.cols * ((do_i ? ((a.dual && a.RW) ? 2 : 1) : 0) + 2 * ((a.gW ? 1 : 0) + (a.rgW ? 1 : 0)));
  return after_launch(ctx, "few_samples_bwd_tma_kernel", KIND_SMALL_N, bytes);
}

template <int MMAX>
int launch_few_rows_wgrad(af_ctx* ctx, FrWgradArgs& a) {
  constexpr int NQ = 2 * MMAX * 2;
  auto kern = few_rows_wgrad_kernel<MMAX>;
  const long long colblocks = (a.cols + 63) / 64;
  long long chunks = ctx->sm_count / colblocks;   // one resident block per SM (accumulators + a tile's X / RX pairs in registers)
  const long long max_by_n = (a.n + 63) / 64, max_by_scratch = ctx->scratch_doubles / (colblocks * NQ * 32);
  if (chunks > max_by_n) chunks = max_by_n;
  if (chunks > max_by_scratch) chunks = max_by_scratch;
  if (chunks < 1) chunks = 1;
  a.per_chunk = ((a.n + chunks - 1) / chunks + kFrwTile - 1) / kFrwTile * kFrwTile;
  a.chunks = (int)((a.n + a.per_chunk - 1) / a.per_chunk);
  kern<<<dim3((unsigned)colblocks, (unsigned)a.chunks), 256, 0, ctx->stream>>>(a);
  return after_launch(ctx, "few_rows_wgrad_kernel", KIND_SMALL_N, 8.0 * a.n * a.cols * ((a.dual && a.RX) ? 2 : 1));
}

}  // namespace

// 16-byte pairs everywhere: even pitches, aligned bases; the fold groups' ticket counters bound the extents
bool few_samples_applicable(af_ctx* ctx, long long rows, long long cols, long long n, std::initializer_list<const void*> ptrs) {
  if (ctx->gemm_path == 1 || ctx->few_samples_off || !ctx->fold_tickets) return false;
  if (n < ctx->few_samples_min || n > kFsMaxSamples || rows < 1 || cols < 2 || (cols & 1)) return false;
  if ((rows + 31) / 32 > kFoldTickets || (cols + 15) / 16 > kFoldTickets || rows > (1 << 28) || cols > (1 << 28)) return false;
  for (const void* p : ptrs)
    if (p && !aligned16p(p)) return false;
  return true;
}

int few_samples_fwd_launch(af_ctx* ctx, const double* W, const double* RW, const double* X, const double* RX, long long rows,
                           long long cols, long long n, const GemmEpilogue& e, bool dual) {
  FsFwdArgs a{};
  a.W = W; a.RW = dual ? RW : nullptr; a.X = X; a.RX = dual ? RX : nullptr;
  a.Y = e.out0; a.RY = dual ? e.out1 : nullptr; a.b = e.bias0; a.Rb = dual ? e.bias1 : nullptr;
  a.part = ctx->scratch; a.tickets = ctx->fold_tickets;
  a.rows = (int)rows; a.cols = (int)cols; a.n = (int)n; a.epi = e.epi; a.dual = dual ? 1 : 0;
  a.ppitch = (int)((rows + 1) & ~1ll);
  return n <= 8 ? launch_fwd<1>(ctx, a) : launch_fwd<2>(ctx, a);
}

int few_samples_bwd_launch(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW, const double* X,
                           const double* RX, double* gW, double* rgW, double* D, double* DR, const double* S, const double* RS,
                           long long rows, long long cols, long long n, bool accumulate_d) {
  FsBwdArgs a{};
  a.dual = (rgW || DR) ? 1 : 0;
  a.U = U; a.UR = a.dual ? UR : nullptr; a.W = W; a.RW = a.dual ? RW : nullptr; a.X = X; a.RX = a.dual ? RX : nullptr;
  a.gW = gW; a.rgW = rgW; a.D = D; a.DR = DR; a.S = S; a.RS = RS;
  a.part = ctx->scratch; a.tickets = ctx->fold_tickets;
  a.rows = (int)rows; a.cols = (int)cols; a.n = (int)n; a.accumulate_d = accumulate_d ? 1 : 0;
  // the TMA form needs 16-byte global strides (even rows for U / UR, even columns for the weight-sized operands) and
  // 16-byte bases
  bool tma_ok = ctx->few_samples_tma && ctx->encode_tiled && (rows % 2 == 0) && (cols % 2 == 0);
  for (const void* p : {(const void*)U, (const void*)UR, (const void*)W, (const void*)RW, (const void*)gW, (const void*)rgW})
    if (p && !aligned16p(p)) tma_ok = false;
  if (tma_ok) {
    if (n <= 4) return launch_bwd_tma<1>(ctx, a);
    if (n <= 8) return launch_bwd_tma<2>(ctx, a);
    if (n <= 12) return launch_bwd_tma<3>(ctx, a);
    return launch_bwd_tma<4>(ctx, a);
  }
  if (n <= 4) return launch_bwd<1>(ctx, a);
  if (n <= 8) return launch_bwd<2>(ctx, a);
  if (n <= 12) return launch_bwd<3>(ctx, a);
  return launch_bwd<4>(ctx, a);
}

// gW (rows x cols, rows <= 16) += U^T X over a large batch, rgW += UR^T X + U^T RX; deterministic
bool few_rows_wgrad_applicable(af_ctx* ctx, long long rows, long long cols, long long n, std::initializer_list<const void*> ptrs) {
  if (ctx->gemm_path == 1 || ctx->small_n_off || !ctx->fold_tickets) return false;
  if (rows < 1 || rows > 16 || n < 64 || cols < 2 || (cols & 1) || (cols + 63) / 64 > kFoldTickets || cols > (1 << 28)) return false;
  for (const void* p : ptrs)
    if (p && !aligned16p(p)) return false;
  return true;
}

int few_rows_wgrad_launch(af_ctx* ctx, const double* U, const double* UR, const double* X, const double* RX, double* gW, double* rgW,
                          long long rows, long long cols, long long n) {
  FrWgradArgs a{};
  a.dual = rgW ? 1 : 0;
  a.U = U; a.UR = a.dual ? UR : nullptr; a.X = X; a.RX = a.dual ? RX : nullptr; a.gW = gW; a.rgW = rgW;
  a.part = ctx->scratch; a.tickets = ctx->fold_tickets;
  a.rows = (int)rows; a.cols = (int)cols; a.n = n;
  if (rows <= 8) return launch_few_rows_wgrad<8>(ctx, a);
  return rows <= 12 ? launch_few_rows_wgrad<12>(ctx, a) : launch_few_rows_wgrad<16>(ctx, a);
}

}  // namespace af
