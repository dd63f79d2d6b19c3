// Whole-step persistent kernel for the reference-exact workload: bench/mlp.go with a handful of samples (n <= 4; the
// reference benchmark as written is n = 1).  At this size the step is a sequence of HBM-bound streaming passes over the
// weight matrices (W, RW read once forward and once for the input gradient, gW, rgW read-modify-written
// once) separated by tiny dependencies, and a launch per pass costs as much as the pass: the 15-launch CUDA graph spent
// ~80 us on 162 MB (24.7 us at the measured HBM rate).  Here ONE cooperative kernel walks all phases, separated by
// grid-wide barriers:
//
//   F_l  (l = 1..L)  S_l = sigmoid(X_l W_l^T + b_l), RS_l      lin_tran.go:90-99,138-152 (the Gemv branch) + lin_add.go
//                                                              + math_funcs.go:288-315, finished inside the block
//   B_l  (l = L..1)  ONE pass over the rows of layer l does      lin_tran.go:187-196,223-241 (Ger), :285-288,329-334
//                    gW += d x^T, rgW += d Rx^T + dR x^T,        (Gemv T), lin_add.go:94-104 (bias), with the sigmoid
//                    gb += d, rgb += dR and D = W^T d, DR = ...  backward of math_funcs.go:357-374 applied to the raw
//                                                                upstream as it is loaded (d = S(1-S) U, ...)
//
// so W/RW/gW/rgW of a layer are touched exactly once per direction (the 8 algorithmic passes; 6 with fresh
// accumulators, which are stored instead of read-modify-written), the launch count is 1, and the L + (L-1)
// dependencies cost a grid barrier each (measured ~3.3 us: two L2 round trips) instead of a kernel boundary.
// One CTA per SM (148 barrier arrivals) made of 2-4 independent 256-thread groups that pick work items on their own.
#include "common.h"

namespace af {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxLayers = 8;
constexpr int kMChunkMax = 32;  // rows of W per backward work item (delta staged in shared memory)

struct SmallMlpArgs {
  int L, n;
  int sizes[kMaxLayers + 1];
  const double* W[kMaxLayers];
  const double* RW[kMaxLayers];  // NULL: variable absent from the RVector (variable.go:85-91) or non-R step
  const double* b[kMaxLayers];
  const double* Rb[kMaxLayers];
  double* gW[kMaxLayers];
  double* rgW[kMaxLayers];
  double* gb[kMaxLayers];
  double* rgb[kMaxLayers];
  double* act[kMaxLayers + 1];     // act[0] = input batch
  double* ract[kMaxLayers + 1];    // ract[0] unused (the input is not in the RVector)
  double* delta[kMaxLayers + 1];   // delta[L] = caller's upstream; delta[l] (0 < l < L) = raw input gradient of layer l+1
  double* rdelta[kMaxLayers + 1];
  int fwd_wpr[kMaxLayers];         // warps per weight row in F_l
  int bwd_mchunk[kMaxLayers];      // rows per work item in B_l
  int backprop;
  int fresh;                       // accumulators start this step at zero: store instead of read-modify-write
  int unit_upstream;               // the top upstream is outGrad = 1, outRGrad = 0, generated in registers
  int head_local;                  // the top layer (<= 16 outputs) runs forward + backward redundantly in every group: two grid barriers fewer
  unsigned* bar;                   // grid barrier counter: monotonic across launches (never reset)
  unsigned bar_base;               // its value when this launch starts
};

// A CTA is G independent groups of kThreads threads ("virtual blocks"): one CTA per SM keeps the grid barrier at 148
// arrivals, while G x 8 warps per SM keep enough loads in flight.  Groups synchronise on their own named barrier.
__device__ __forceinline__ void group_sync(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kThreads) : "memory");
}

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

// All blocks are co-resident (cooperative launch), so a counter in global memory is a barrier: arrive with a release
// add, spin with acquire loads until every block of this generation has arrived.  The counter is never reset: the
// host knows how many arrivals a launch makes and hands the next launch its starting value.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& generation) {
  __syncthreads();  // every thread's writes of this phase happen-before thread 0's release below (CTA-scope barrier)
  if (threadIdx.x == 0) {
    generation += gridDim.x;
    // release-arrive / acquire-poll: two L2 round trips, no membar.gl (measured ~1 us less per barrier than
    // __threadfence + atomicAdd)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned seen;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if ((int)(seen - generation) >= 0) break;  // wrap-safe
      __nanosleep(20);
    }
  }
  __syncthreads();
}

// ---- F_l ------------------------------------------------------------------------------------------------
template <int NMAX, bool R>
__device__ __forceinline__ void forward_phase(const SmallMlpArgs& a, int l, double (*part)[2 * NMAX], int group, int tid,
                                              int vblock, int nvblocks) {
  const int K = a.sizes[l], M = a.sizes[l + 1], n = a.n, wpr = a.fwd_wpr[l];
  const double* __restrict__ W = a.W[l];
  const double* __restrict__ RW = R ? a.RW[l] : nullptr;
  const double* __restrict__ X = a.act[l];
  const double* __restrict__ RX = (R && l > 0) ? a.ract[l] : nullptr;
  const double* __restrict__ bias = a.b[l];
  const double* __restrict__ Rb = R ? a.Rb[l] : nullptr;
  double* __restrict__ S = a.act[l + 1];
  double* __restrict__ RS = a.ract[l + 1];
  const int warp = tid >> 5, lane = tid & 31;
  const int rpb = 8 / wpr, ws = warp % wpr;
  const int items = (M + rpb - 1) / rpb;
  for (int item = vblock; item < items; item += nvblocks) {
    const int m = item * rpb + warp / wpr;
    double y[NMAX], ry[NMAX];
#pragma unroll
    for (int i = 0; i < NMAX; ++i) y[i] = ry[i] = 0.0;
    if (m < M) {
      const double* w = W + (long long)m * K;
      const double* rw = RW ? RW + (long long)m * K : nullptr;
#pragma unroll 4
      for (int k = 2 * lane + 64 * ws; k < K; k += 64 * wpr) {
        const double2 wv = ld2(w + k);
        double2 rwv = make_double2(0.0, 0.0);
        if (R && rw) rwv = ld2(rw + k);
#pragma unroll
        for (int i = 0; i < NMAX; ++i) {
          if (i < n) {
            const double2 xv = ld2(X + (long long)i * K + k);
            y[i] = fma(xv.x, wv.x, fma(xv.y, wv.y, y[i]));
            if (R && rw) ry[i] = fma(xv.x, rwv.x, fma(xv.y, rwv.y, ry[i]));
            if (R && RX) {
              const double2 rxv = ld2(RX + (long long)i * K + k);
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
        if (R) ry[i] += __shfl_xor_sync(0xffffffffu, ry[i], o);
      }
      if (lane == 0) { part[warp][2 * i] = y[i]; part[warp][2 * i + 1] = ry[i]; }
    }
    group_sync(group);
    // the first warp of each row group folds the group's partial sums; lane i finishes sample i:
    // + bias (lin_add.go:37-42), sigmoid and its R form (math_funcs.go:305-309)
    if (m < M && ws == 0 && lane < n) {
      double z = 0.0, zr = 0.0;
      for (int q = 0; q < wpr; ++q) { z += part[warp + q][2 * lane]; zr += part[warp + q][2 * lane + 1]; }
      z += bias[m];
      const double s = 1.0 / (1.0 + exp(-z));
      S[(long long)lane * M + m] = s;
      if (R) {
        zr += Rb ? Rb[m] : 0.0;
        RS[(long long)lane * M + m] = s * (1 - s) * zr;
      }
    }
    group_sync(group);  // part[] is reused by the next item
  }
}

// ---- B_l ------------------------------------------------------------------------------------------------
// `hd` != NULL: layer l lies directly below a head that ran locally (head_phase): its raw upstream U = W_head^T d_head
// (and the R form) is formed here from the head's deltas hd / hdr ([NMAX][kHeadMax], in shared memory) instead of being
// read from delta[l + 1], which nobody wrote.
constexpr int kHeadMax = 16;
template <int NMAX, bool R>
__device__ __forceinline__ void backward_phase(const SmallMlpArgs& a, int l, double* sd /* [2][NMAX][kMChunkMax] */,
                                               int group, int tid, int vblock, int nvblocks, const double* hd = nullptr,
                                               const double* hdr = nullptr) {
  const int K = a.sizes[l], M = a.sizes[l + 1], n = a.n, mchunk = a.bwd_mchunk[l];
  const double* __restrict__ W = a.W[l];
  const double* __restrict__ RW = R ? a.RW[l] : nullptr;
  const double* __restrict__ X = a.act[l];
  const double* __restrict__ RX = (R && l > 0) ? a.ract[l] : nullptr;
  const double* __restrict__ U = a.delta[l + 1];
  const double* __restrict__ UR = R ? a.rdelta[l + 1] : nullptr;
  const double* __restrict__ S = a.act[l + 1];
  const double* __restrict__ RS = R ? a.ract[l + 1] : nullptr;
  double* __restrict__ gW = a.gW[l];
  double* __restrict__ rgW = R ? a.rgW[l] : nullptr;
  double* __restrict__ D = l > 0 ? a.delta[l] : nullptr;  // layer 1's input is not in the gradient maps (lin_tran.go:420)
  double* __restrict__ DR = (R && l > 0) ? a.rdelta[l] : nullptr;
  const int xblocks = (K / 2 + kThreads - 1) / kThreads;
  const int ychunks = (M + mchunk - 1) / mchunk;
  double* sdr = sd + NMAX * kMChunkMax;
  const bool top_unit = a.unit_upstream && l + 1 == a.L;  // outGrad = 1, outRGrad = 0 (bench/mlp.go:118-126)
  const bool fresh = a.fresh != 0;                        // fresh accumulators: store instead of += (NewGradient per step)
  for (int item = vblock; item < xblocks * ychunks; item += nvblocks) {
    const int bx = item % xblocks, by = item / xblocks;
    const int m0 = by * mchunk, mc = min(mchunk, M - m0);
    // stage the layer's upstream for this row chunk with the sigmoid backward applied as it is loaded
    // (math_funcs.go:365-371, operand order kept): d = S(1-S)U ; dR = RS(1-S)U - S RS U + S(1-S)UR
    for (int e = tid; e < NMAX * mchunk; e += kThreads) {
      const int i = e / mchunk, mm = e - i * mchunk;
      double d = 0.0, dr = 0.0;
      if (i < n && mm < mc) {
        const long long o = (long long)i * M + m0 + mm;
        double p = top_unit ? 1.0 : 0.0, pr = 0.0;
        if (hd) {
          // input gradient of the head toward this layer's output m0 + mm (lin_tran.go:285-288,329-334, Gemv T):
          // column m0 + mm of W_head / RW_head against the head's deltas of sample i
          const int Mh = a.sizes[a.L], col = m0 + mm;
          const double* __restrict__ Wh = a.W[a.L - 1];
          const double* __restrict__ RWh = R ? a.RW[a.L - 1] : nullptr;
          for (int r = 0; r < Mh; ++r) {
            const double w = Wh[(long long)r * M + col], dh = hd[i * kHeadMax + r];
            p = fma(dh, w, p);
            if (R) {
              pr = fma(hdr[i * kHeadMax + r], w, pr);
              if (RWh) pr = fma(dh, RWh[(long long)r * M + col], pr);
            }
          }
        } else if (!top_unit) {
          p = U[o];
          if (R) pr = UR[o];
        }
        const double x = S[o];
        d = x * (1 - x) * p;
        if (R) {
          const double xr = RS[o];
          dr = xr * (1 - x) * p - x * xr * p + x * (1 - x) * pr;
        }
      }
      sd[i * kMChunkMax + mm] = d;
      if (R) sdr[i * kMChunkMax + mm] = dr;
    }
    group_sync(group);
    // bias gradients of these rows (lin_add.go:94-104, batched: column sums over the samples)
    if (bx == 0 && tid < mc) {
      double g = 0.0, rg = 0.0;
      for (int i = 0; i < n; ++i) { g += sd[i * kMChunkMax + tid]; if (R) rg += sdr[i * kMChunkMax + tid]; }
      a.gb[l][m0 + tid] = fresh ? g : a.gb[l][m0 + tid] + g;
      if (R) a.rgb[l][m0 + tid] = fresh ? rg : a.rgb[l][m0 + tid] + rg;
    }
    const int k = 2 * (bx * kThreads + tid);
    if (k < K) {
      double2 xv[NMAX], rxv[NMAX], dv[NMAX], drv[NMAX];
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
        xv[i] = rxv[i] = dv[i] = drv[i] = make_double2(0.0, 0.0);
        if (i < n) {
          xv[i] = ld2(X + (long long)i * K + k);
          if (R && RX) rxv[i] = ld2(RX + (long long)i * K + k);
        }
      }
#pragma unroll((NMAX <= 2 && !R) ? 4 : 2)
      for (int mm = 0; mm < mc; ++mm) {
        const long long o = (long long)(m0 + mm) * K + k;
        double2 g = make_double2(0.0, 0.0), rg = make_double2(0.0, 0.0);
        double2 wv = make_double2(0.0, 0.0), rwv = make_double2(0.0, 0.0);
        if (!fresh) g = ld2(gW + o);
        if (R && !fresh) rg = ld2(rgW + o);
        if (D) wv = ld2(W + o);
        if (D && R && RW) rwv = ld2(RW + o);
#pragma unroll
        for (int i = 0; i < NMAX; ++i) {
          if (i < n) {
            const double d = sd[i * kMChunkMax + mm];
            g.x = fma(d, xv[i].x, g.x); g.y = fma(d, xv[i].y, g.y);
            if (D) { dv[i].x = fma(d, wv.x, dv[i].x); dv[i].y = fma(d, wv.y, dv[i].y); }
            if (R) {
              const double dr = sdr[i * kMChunkMax + mm];
              rg.x = fma(d, rxv[i].x, fma(dr, xv[i].x, rg.x));
              rg.y = fma(d, rxv[i].y, fma(dr, xv[i].y, rg.y));
              if (D) {
                drv[i].x = fma(dr, wv.x, fma(d, rwv.x, drv[i].x));
                drv[i].y = fma(dr, wv.y, fma(d, rwv.y, drv[i].y));
              }
            }
          }
        }
        st2(gW + o, g);
        if (R) st2(rgW + o, rg);
      }
      if (D) {
#pragma unroll
        for (int i = 0; i < NMAX; ++i) {
          if (i < n) {
            const long long o = (long long)i * K + k;
            atomicAdd(D + o, dv[i].x); atomicAdd(D + o + 1, dv[i].y);
            if (R) { atomicAdd(DR + o, drv[i].x); atomicAdd(DR + o + 1, drv[i].y); }
          }
        }
      }
    }
    group_sync(group);  // sd[] is reused by the next item
  }
}

// ---- the head, locally -----------------------------------------------------------------------------------
// The top layer of the bench network has 10 outputs: as F_L and B_L it is two phases with almost no work in them -- a
// handful of groups compute ten dot products, then ONE group runs the rank-n update of a 10 x 512 matrix -- and each
// costs every CTA a grid barrier.  Its forward needs nothing but the complete activations of the layer below, and
// everything its backward produces is either tiny (gW_L, gb_L) or needed only row by row by the layer below (the input
// gradient).  So every group computes the head's forward REDUNDANTLY right after the barrier that completes the layer
// below (W_L and RW_L are a few tens of KB from L2; a CTA's groups share the rows), forms the head's deltas in shared memory
// (hd / hdr), group 0 of CTA 0
// stores S_L / RS_L and the head's weight and bias gradients, and B_{L-1} starts at once, building its raw upstream from
// W_L's columns and the deltas (backward_phase, `hd`).  Two grid barriers and two serial phases leave the step.
// rows of the head whose loads are in flight together (a row at a time would be ten dependent L2 round trips)
template <int NMAX> struct HeadBatch { static constexpr int MB = NMAX == 1 ? 5 : (NMAX == 2 ? 3 : 2); };

template <int NMAX, bool R>
__device__ __forceinline__ void head_phase(const SmallMlpArgs& a, double* hpart /* [8][MB][2 NMAX] */, double* hd, double* hdr,
                                           int group, int ngroups, int tid) {
  constexpr int MB = HeadBatch<NMAX>::MB;
  const int l = a.L - 1, K = a.sizes[l], M = a.sizes[l + 1], n = a.n;
  const double* __restrict__ W = a.W[l];
  const double* __restrict__ RW = R ? a.RW[l] : nullptr;
  const double* __restrict__ X = a.act[l];
  const double* __restrict__ RX = (R && l > 0) ? a.ract[l] : nullptr;
  const int warp = tid >> 5, lane = tid & 31;
  const bool fresh = a.fresh != 0, top_unit = a.unit_upstream != 0;
  // forward: the CTA's groups share the head's rows (hd / hdr belong to the CTA); a group's threads split the contraction in
  // column pairs, MB rows of W at a time, all samples
  const int rpg = (M + ngroups - 1) / ngroups, m_end = min(M, (group + 1) * rpg);
  for (int m0 = group * rpg; m0 < m_end; m0 += MB) {
    double y[MB][NMAX], ry[MB][NMAX];
#pragma unroll
    for (int j = 0; j < MB; ++j)
#pragma unroll
      for (int i = 0; i < NMAX; ++i) y[j][i] = ry[j][i] = 0.0;
    for (int k = 2 * tid; k < K; k += 2 * kThreads) {
      double2 wv[MB], rwv[MB], xv[NMAX], rxv[NMAX];
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        const bool ok = m0 + j < m_end;
        wv[j] = ok ? ld2(W + (long long)(m0 + j) * K + k) : make_double2(0.0, 0.0);
        rwv[j] = (ok && R && RW) ? ld2(RW + (long long)(m0 + j) * K + k) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
        xv[i] = i < n ? ld2(X + (long long)i * K + k) : make_double2(0.0, 0.0);
        rxv[i] = (i < n && R && RX) ? ld2(RX + (long long)i * K + k) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int j = 0; j < MB; ++j)
#pragma unroll
        for (int i = 0; i < NMAX; ++i) {
          y[j][i] = fma(xv[i].x, wv[j].x, fma(xv[i].y, wv[j].y, y[j][i]));
          if (R) ry[j][i] = fma(rxv[i].x, wv[j].x, fma(rxv[i].y, wv[j].y, fma(xv[i].x, rwv[j].x, fma(xv[i].y, rwv[j].y, ry[j][i]))));
        }
    }
#pragma unroll
    for (int j = 0; j < MB; ++j)
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          y[j][i] += __shfl_xor_sync(0xffffffffu, y[j][i], o);
          if (R) ry[j][i] += __shfl_xor_sync(0xffffffffu, ry[j][i], o);
        }
        if (lane == 0) {
          hpart[(warp * MB + j) * 2 * NMAX + 2 * i] = y[j][i];
          hpart[(warp * MB + j) * 2 * NMAX + 2 * i + 1] = ry[j][i];
        }
      }
    group_sync(group);
    if (tid < MB * NMAX) {  // thread (j, i) finishes sample i of output m0 + j: bias, sigmoid, R form, top of the backward pass
      const int j = tid / NMAX, i = tid - j * NMAX, m = m0 + j;
      if (m < m_end && i < n) {
        double z = 0.0, zr = 0.0;
        for (int q = 0; q < 8; ++q) {
          z += hpart[(q * MB + j) * 2 * NMAX + 2 * i];
          zr += hpart[(q * MB + j) * 2 * NMAX + 2 * i + 1];
        }
        z += a.b[l][m];
        const double sg = 1.0 / (1.0 + exp(-z));
        const long long o = (long long)i * M + m;
        double rs = 0.0;
        if (R) {
          zr += a.Rb[l] ? a.Rb[l][m] : 0.0;
          rs = sg * (1 - sg) * zr;
        }
        // math_funcs.go:365-371 on the caller's upstream (or outGrad = 1 / outRGrad = 0), operand order kept
        const double u = top_unit ? 1.0 : a.delta[a.L][o];
        hd[i * kHeadMax + m] = sg * (1 - sg) * u;
        if (R) {
          const double ur = top_unit ? 0.0 : a.rdelta[a.L][o];
          hdr[i * kHeadMax + m] = rs * (1 - sg) * u - sg * rs * u + sg * (1 - sg) * ur;
        }
        if (blockIdx.x == 0) {
          a.act[l + 1][o] = sg;
          if (R) a.ract[l + 1][o] = rs;
        }
      }
    }
    group_sync(group);  // hpart[] is reused by the next rows
  }
  __syncthreads();  // every group's rows are in hd / hdr (all groups of a CTA reach this point together: they left the
                    // grid barrier together and run the same head)
  if (blockIdx.x != 0 || group != 0) return;
  // the head's own gradients, once: gb_L += sum_i d_i (lin_add.go:94-104), gW_L += d x^T, rgW_L += d Rx^T + dR x^T (Ger)
  if (tid < M) {
    double g = 0.0, rg = 0.0;
    for (int i = 0; i < n; ++i) { g += hd[i * kHeadMax + tid]; if (R) rg += hdr[i * kHeadMax + tid]; }
    a.gb[l][tid] = fresh ? g : a.gb[l][tid] + g;
    if (R) a.rgb[l][tid] = fresh ? rg : a.rgb[l][tid] + rg;
  }
  for (int k = 2 * tid; k < K; k += 2 * kThreads) {
    double2 xv[NMAX], rxv[NMAX];
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
      xv[i] = rxv[i] = make_double2(0.0, 0.0);
      if (i < n) {
        xv[i] = ld2(X + (long long)i * K + k);
        if (R && RX) rxv[i] = ld2(RX + (long long)i * K + k);
      }
    }
    for (int m = 0; m < M; ++m) {
      const long long o = (long long)m * K + k;
      double2 g = make_double2(0.0, 0.0), rg = make_double2(0.0, 0.0);
      if (!fresh) g = ld2(a.gW[l] + o);
      if (R && !fresh) rg = ld2(a.rgW[l] + o);
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
        if (i < n) {
          const double d = hd[i * kHeadMax + m];
          g.x = fma(d, xv[i].x, g.x); g.y = fma(d, xv[i].y, g.y);
          if (R) {
            const double dr = hdr[i * kHeadMax + m];
            rg.x = fma(d, rxv[i].x, fma(dr, xv[i].x, rg.x));
            rg.y = fma(d, rxv[i].y, fma(dr, xv[i].y, rg.y));
          }
        }
      }
      st2(a.gW[l] + o, g);
      if (R) st2(a.rgW[l] + o, rg);
    }
  }
}

template <int NMAX, bool R, int G>
__global__ void __launch_bounds__(kThreads * G, 1) small_mlp_step_kernel(const SmallMlpArgs a) {
  __shared__ double part[G][8][2 * NMAX];
  __shared__ double sd[G][2 * NMAX * kMChunkMax];
  const int group = threadIdx.x / kThreads, tid = threadIdx.x % kThreads;
  const int vblock = blockIdx.x * G + group, nvblocks = gridDim.x * G;
  unsigned generation = a.bar_base;
  // the raw input gradients are accumulated with red.add: zero them while the forward pass runs
  if (a.backprop) {
    for (int l = 1; l < a.L; ++l) {
      const long long cnt = (long long)a.n * a.sizes[l];
      for (long long e = (long long)vblock * kThreads + tid; e < cnt; e += (long long)nvblocks * kThreads) {
        a.delta[l][e] = 0.0;
        if (R) a.rdelta[l][e] = 0.0;
      }
    }
  }
  if (a.head_local) {  // (host: backprop, L >= 2, head of <= kHeadMax outputs)
    __shared__ double hd[NMAX * kHeadMax], hdr[NMAX * kHeadMax];
    __shared__ double hpart[G][8 * HeadBatch<NMAX>::MB * 2 * NMAX];
    for (int l = 0; l + 1 < a.L; ++l) {
      forward_phase<NMAX, R>(a, l, part[group], group, tid, vblock, nvblocks);
      grid_barrier(a.bar, generation);
    }
    head_phase<NMAX, R>(a, hpart[group], hd, hdr, group, G, tid);
    for (int l = a.L - 2; l >= 0; --l) {
      backward_phase<NMAX, R>(a, l, sd[group], group, tid, vblock, nvblocks, l == a.L - 2 ? hd : nullptr,
                              l == a.L - 2 ? hdr : nullptr);
      if (l > 0) grid_barrier(a.bar, generation);
    }
    return;
  }
  for (int l = 0; l < a.L; ++l) {
    forward_phase<NMAX, R>(a, l, part[group], group, tid, vblock, nvblocks);
    if (l + 1 < a.L || a.backprop) grid_barrier(a.bar, generation);
  }
  if (!a.backprop) return;
  for (int l = a.L - 1; l >= 0; --l) {
    backward_phase<NMAX, R>(a, l, sd[group], group, tid, vblock, nvblocks);
    if (l > 0) grid_barrier(a.bar, generation);
  }
}

template <int NMAX, bool R, int G>
int launch_variant(af_ctx* ctx, SmallMlpArgs& a, double bytes) {
  auto kernel = small_mlp_step_kernel<NMAX, R, G>;
  const int grid = ctx->sm_count;            // one CTA per SM, all co-resident (cooperative launch)
  const int vgrid = grid * G;                // 256-thread virtual blocks
  // Work decomposition per phase.  These passes are latency-bound, not throughput-bound: what matters is bytes in
  // flight (threads x unrolled 16-byte loads), and that the fixed cost of a work item (staging, reduction, block
  // barriers) is amortised over enough rows -- not that every resident block gets an item.
  for (int l = 0; l < a.L; ++l) {
    const long long M = a.sizes[l + 1], K = a.sizes[l];
    // forward: ~256 contraction entries (4 unrolled double2 loads per lane and matrix) per warp
    int wpr = K <= 384 ? 1 : K <= 768 ? 2 : K <= 1536 ? 4 : 8;
    while (wpr < 8 && (M * wpr + 7) / 8 < vgrid / 2 && K / (2 * wpr) >= 64) wpr *= 2;
    a.fwd_wpr[l] = wpr;
    // backward: 8-16 rows per item (2-4 unrolled batches), more when that still leaves >= ~256 items
    const long long xblocks = (K / 2 + kThreads - 1) / kThreads;
    long long mchunk = M * xblocks / 256;
    if (mchunk < 8) mchunk = 8;
    if (mchunk > 16) mchunk = 16;
    if (mchunk > M) mchunk = M;
    a.bwd_mchunk[l] = (int)mchunk;
  }
  // every block arrives at every barrier, so the counter ends this launch at base + barriers * grid
  const unsigned barriers = a.head_local ? (unsigned)((a.L - 1) + (a.L - 2))
                                         : (unsigned)((a.L - 1) + (a.backprop ? 1 + (a.L - 1) : 0));
  a.bar_base = ctx->barrier_base;
  ctx->barrier_base += barriers * (unsigned)grid;
  void* params[] = {(void*)&a};
  const cudaError_t le =
      cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(kThreads * G), params, 0, ctx->stream);
  if (le != cudaSuccess) {
    // the grid cannot be made co-resident here (SMs held by another client, stream being captured, ...): nothing was
    // launched; tell the caller to take the per-kernel path instead
    ctx->barrier_base = a.bar_base;
    cudaGetLastError();
    return fail(AF_ENOTSUP, std::string("cooperative launch of small_mlp_step_kernel refused: ") + cudaGetErrorString(le));
  }
  return after_launch(ctx, "small_mlp_step_kernel", KIND_SMALL_N, bytes);
}

}  // namespace

bool small_mlp_applicable(af_ctx* ctx, int L, const int64_t* sizes, int64_t n) {
  if (ctx->gemm_path == 1 || ctx->small_n_off || ctx->small_mlp_off || ctx->deterministic) return false;  // (input-gradient atomics)
  // n = 5..8 stay on the per-kernel graph: their R variant needs > 200 registers (one group of 8 warps per SM) and
  // measured slower (183 vs 147 us at n = 8)
  if (n < 1 || n > 4 || L < 1 || L > kMaxLayers) return false;
  for (int l = 0; l <= L; ++l)
    if (sizes[l] < 1 || sizes[l] > (1 << 28)) return false;
  for (int l = 0; l < L; ++l)
    if (sizes[l] & 1) return false;  // 16-byte row accesses: every contraction length must be even
  return true;
}

// One R-step (or plain step) of an L-layer sigmoid MLP.  Pointer arrays are indexed like af_mlp's:
// param/rvec/grad/rgrad[2l] = weights of layer l, [2l+1] = biases; rvec entries may be NULL (absent).
int small_mlp_launch(af_ctx* ctx, int L, const int64_t* sizes, int64_t n, double* const* param, const double* const* rvec,
                     double* const* grad, double* const* rgrad, double* const* act, double* const* ract,
                     double* const* delta, double* const* rdelta, bool with_r, bool backprop, bool fresh,
                     bool unit_upstream) {
  if (!ctx->barrier_word) {
    AF_CUDA(cudaMalloc(&ctx->barrier_word, 256));
    AF_CUDA(cudaMemsetAsync(ctx->barrier_word, 0, 256, ctx->stream));
    ctx->barrier_base = 0;
  }
  SmallMlpArgs a{};
  a.L = L; a.n = (int)n; a.backprop = backprop ? 1 : 0; a.bar = ctx->barrier_word;
  a.fresh = fresh ? 1 : 0; a.unit_upstream = unit_upstream ? 1 : 0;
  a.head_local = (backprop && L >= 2 && sizes[L] <= kHeadMax && !ctx->small_mlp_head_off) ? 1 : 0;
  double bytes = 0;
  for (int l = 0; l <= L; ++l) {
    a.sizes[l] = (int)sizes[l];
    a.act[l] = act[l]; a.ract[l] = ract[l]; a.delta[l] = delta[l]; a.rdelta[l] = rdelta[l];
  }
  for (int l = 0; l < L; ++l) {
    a.W[l] = param[2 * l]; a.b[l] = param[2 * l + 1];
    a.RW[l] = with_r ? rvec[2 * l] : nullptr; a.Rb[l] = with_r ? rvec[2 * l + 1] : nullptr;
    a.gW[l] = grad[2 * l]; a.gb[l] = grad[2 * l + 1];
    a.rgW[l] = rgrad[2 * l]; a.rgb[l] = rgrad[2 * l + 1];
    const double mk = 8.0 * sizes[l] * sizes[l + 1];
    const int rmat = with_r ? 2 : 1;
    bytes += mk * (with_r && a.RW[l] ? 2 : 1);                                   // forward: W (, RW)
    if (backprop) bytes += (fresh ? 1 : 2) * mk * rmat + (l > 0 ? mk * (with_r && a.RW[l] ? 2 : 1) : 0);  // gW, rgW (RMW); W, RW again
  }
  // registers per thread decide how many 256-thread groups share an SM (64 K registers): <= 64 -> 4, <= 85 -> 3,
  // <= 128 -> 2, else 1 (checked with -Xptxas -v: no variant spills more than a few words)
#define AF_SMALL_MLP(NM, GR, GN)                                                 \
  (with_r ? launch_variant<NM, true, GR>(ctx, a, bytes) : launch_variant<NM, false, GN>(ctx, a, bytes))
  if (n <= 1) return AF_SMALL_MLP(1, 2, 4);  // measured for the n = 1 R-step: G = 2 54.5 us, 3 61.7, 4 (spills) 63.6
  if (n <= 2) return AF_SMALL_MLP(2, 2, 3);
  return AF_SMALL_MLP(4, 2, 2);
#undef AF_SMALL_MLP
}

}  // namespace af
