// FP64 tensor-core contraction for the LinTran kernels (lin_tran.go:79-359), sm_100a.
//
// One kernel template serves all six BLAS call sites of the reference.
// In logical indices it computes, for an I x J output tile set and contraction length Kd,
//
//     acc0[i][j] = sum_k A0[i,k] * B0[j,k]
//     acc1[i][j] = sum_k A0[i,k] * B1[j,k] + A1[i,k] * B0[j,k]        (NACC == 2 only)
//
//   forward  (G1+G2, lin_tran.go:113,172-173): A=X,RX  B=W,RW        out n x rows
//   igrad    (G5+G6, lin_tran.go:302,354-355): A=U,UR  B=W,RW (k = row of W)   out n x cols
//   wgrad    (G3+G4, lin_tran.go:210,267-268): A=U,UR (k = sample)  B=X,RX     out rows x cols, +=
//
// so the R-operator's dual products share one pass over the operand tiles instead of the
// reference's 2-3 separate Gemm calls.
//
// Data path: TMA (cp.async.bulk.tensor.2d, 128B swizzle) -> 4-stage shared-memory ring guarded by
// full/empty mbarriers, filled by one elected producer lane -> ld.shared.f64 fragments ->
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no f64 kind).  CTA tile 128 x 64 x 16 with
// 8 warps of 32 x 32 (or 64 x 16 for outputs at most 16 wide).
//
// Operand layouts in global memory:
//   LAY_KC  the contraction index is contiguous:  element (r,k) at ptr[r*ld + k]
//   LAY_RC  the row index is contiguous:          element (r,k) at ptr[k*ld + r]
// Shared-memory images (both 128B-swizzled by TMA, 16 doubles = 128 B per line):
//   KC tile of R rows : line r holds k = 0..15           -> one box {16, R}
//   RC tile of R rows : R/16 boxes of {16 rows, 16 k}; line k of box c holds rows 16c..16c+15
// Bank conflicts: a half-warp reads 4 lines x 4 k for one fragment.  The k index an MMA lane j
// supplies is arbitrary as long as A and B agree, so lane j of MMA step s takes
// k = KJ[j] ^ MS[s] with KJ = {0,3,12,15}, MS = {0,1,4,5}: for KC tiles the four k's then cover
// both 64-B halves and both 8-B parities, for RC tiles they differ in bits 1-2, and every
// half-warp load touches 16 distinct 8-byte bank pairs in either layout.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "streamk_walk.h"

namespace af {

constexpr int kBK = 16;
constexpr int kBM = 128;
constexpr int kBN = 64;
// stage count, warp count and CTAs per SM are per tile configuration (TileCfg below)

enum : int { LAY_KC = 0, LAY_RC = 1 };
enum : int {
  EPI_STORE = 0,         // out0 = acc0 ; out1 = acc1
  EPI_BIAS = 1,          // out0 = acc0 + b0[j] ; out1 = acc1 + b1[j]           (lin_add.go:15-17,37-42)
  EPI_BIAS_SIGMOID = 2,  // s = sigmoid(acc0+b0[j]) ; out1 = s(1-s)(acc1+b1[j]) (math_funcs.go:305-309)
  EPI_SIGMOID_BWD = 3,   // sigmoid R-backward with (S,RS) read at (i,j)         (math_funcs.go:365-371)
  EPI_ACCUM = 4,         // out += acc   (beta = 1, lin_tran.go:210,267-268)
  EPI_ACCUM_ATOMIC = 5,  // out += acc with red.global.add.f64 (split-K)
  // EPI_SIGMOID_BWD, and the column sums of what was stored are added to colsum0 / colsum1 = the bias gradient of the
  // layer below (lin_add.go:94-104 batched), which otherwise costs one more pass over the upstream it was computed
  // from.  The two targets travel in GemmArgs::bias0 / bias1 (unused by this epilogue; the parameter block keeps its
  // size).  Only whole-tile launches of the <KC, RC> instantiation and few_rows_igrad_kernel take it.
  EPI_SIGMOID_BWD_COLSUM = 6,
};

// Operands of the fused LSTM step epilogues (SPECIAL = 1 / 2 below; lstm_step.cu).  Everything is one time step's
// slice; `ld` is the row pitch of the [state | x] buffers IN / AUG / DAUG (= H + I), gate buffers have pitch 4H.
struct LstmStepArgs {
  int H, ld;
  // forward (bench/lstm.go:139-151,183-191): G / RG hold x W_x^T + b (and the R form) on entry, the activated gates
  // on exit; INp = state entering the step, INn = state leaving it, AUG = masked state (input of the output layer)
  double* G; double* RG;
  const double* INp; const double* RINp;
  double* INn; double* RINn; double* AUG; double* RAUG;
  // backward (bench/lstm.go:215-234): the contraction adds delta4_t W_h to SGF (= sg f of step t), and the epilogue
  // runs the cell backward of step t - 1 on the completed state gradient: Gm / RGm = gates of step t - 1,
  // INm / INt = state entering / leaving step t - 1, DAUGm = gradient toward its masked state; it writes
  // delta4_{t-1} to D4m / RD4m and the new sg f back to SGF / RSGF (pitch H)
  const double* Gm; const double* RGm;
  const double* INm; const double* RINm; const double* INt; const double* RINt;
  const double* DAUGm; const double* RDAUGm;
  double* SGF; double* RSGF; double* D4m; double* RD4m;
};

struct GemmArgs {
  int I, J, Kd;
  int has_a1, has_b1;  // R operands present (NULL R-vector == zero, variable.go:85-91)
  int epi;
  int kt_total;            // k-tiles per output tile
  int tiles_j;             // output tiles along J (tile index = ti * tiles_j + tj)
  int units_per_cta;       // contiguous (tile, k-tile) units per CTA; == kt_total for one tile per CTA
  int units_per_tile;      // == kt_total (stream-K, one tile per CTA) or splits * units_per_cta (aligned split-K)
  long long total_units;
  int wave_tiles;     // > 0 (persistent whole tiles, wave order): CTA c works on tiles c, c + grid, c + 2 grid, ... so that
                      // the CTAs running at one moment are on neighbouring tiles (L2 reuse); position q = c * m + s of
                      // the tile sequence is tile s * grid + c with m = wave_tiles
  int wave_rem;       // the first wave_rem CTAs have m tiles, the others m - 1
  int transpose_out;  // write element (i,j) at out[j*ldo + i] (operands were swapped by the host)
  int rotate;         // stream-K: every tile segment walks its k-tiles in rotated order so that CTAs sweep k in lockstep
  double* out0;
  double* out1;
  long long ldo;
  const double* bias0;
  const double* bias1;
  const double* s;
  const double* rs;
  long long lds;
};
// extra kernel parameter of the SPECIAL variants (empty for the plain contraction: its parameter block stays as small as
// it was -- appending the LSTM operands to GemmArgs itself cost the bench's 2048-CTA forward launch 7 %)
template <int SPECIAL> struct SpecialArgs {};
template <> struct SpecialArgs<1> { LstmStepArgs lstm; };
template <> struct SpecialArgs<2> { LstmStepArgs lstm; };
// SPECIAL = 3: the plain contraction as stream-K with an ORDERED fix-up instead of atomics (see the kernel): a workspace
// of one partial tile per CTA ([CTA][value][thread] doubles) and one flag per CTA and warp
struct FixupArgs { double* part; unsigned* flags; };
template <> struct SpecialArgs<3> { FixupArgs fix; };

// ---------------------------------------------------------------------------------------------
// epilogue element functions (shared by the DMMA kernel and the generic kernel)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double sigmoid_f64(double z) { return 1.0 / (1.0 + exp(-z)); }

// ---- LSTM cell algebra, one (lane, hidden unit) element; shared by the fused step epilogues and the stand-alone
//      cell kernels of lstm.cu so that both paths evaluate the same expressions in the same order -------------------
// forward: zi..zo = the four gates' pre-activations (contraction + x projection + bias), rz* their R forms
//   s = f sp + m i            Add(Mul(forget, state), Mul(inMask, inState))   bench/lstm.go:147-150
//   masked = s o              Mul(outStateVar, outGate)                       bench/lstm.go:191
// R forms follow MulR (arithmetic.go:325-329: x yR + xR y), AddR (:58-65) and the sigmoid's (math_funcs.go:305-309).
struct LstmCellFwd {
  double i, m, f, o, s, masked;
  double ri, rm, rf, ro, rs, rmasked;
};
template <bool R>
__device__ __forceinline__ LstmCellFwd lstm_cell_fwd(double zi, double zm, double zf, double zo, double sp, double rzi,
                                                     double rzm, double rzf, double rzo, double rsp) {
  LstmCellFwd c;
  c.i = sigmoid_f64(zi); c.m = sigmoid_f64(zm); c.f = sigmoid_f64(zf); c.o = sigmoid_f64(zo);
  c.s = c.f * sp + c.m * c.i;
  c.masked = c.s * c.o;
  if (R) {
    c.ri = c.i * (1 - c.i) * rzi; c.rm = c.m * (1 - c.m) * rzm;
    c.rf = c.f * (1 - c.f) * rzf; c.ro = c.o * (1 - c.o) * rzo;
    c.rs = (c.f * rsp + c.rf * sp) + (c.m * c.ri + c.rm * c.i);
    c.rmasked = c.s * c.ro + c.rs * c.o;
  }
  return c;
}
// backward (bench/lstm.go:215-234 with the node backwards of arithmetic.go:34-49,288-306 / 79-96,349-376 and
// math_funcs.go:326-333 / 357-374 inlined):  carry = state gradient arriving from step t + 1
//   sg  = carry + dmask o                 (grad[outStateVar] added to stateGrad, lstm.go:227-228)
//   do  = dmask s                         -> delta_o = sigmoid'(o) do
//   dm  = sg i ; di = sg m ; df = sg sp   -> delta_* = sigmoid'(gate) d*
//   sgf = sg f                            (Mul(forget, state) toward the state)
struct LstmCellBwd {
  double di, dm, df, d_o, sgf;
  double rdi, rdm, rdf, rd_o, rsgf;
};
template <bool R>
__device__ __forceinline__ LstmCellBwd lstm_cell_bwd(double i, double m, double f, double o, double sp, double s,
                                                     double dmask, double carry, double ri, double rm, double rf, double ro,
                                                     double rsp, double rs, double rdmask, double rcarry) {
  LstmCellBwd c;
  const double sg = carry + dmask * o;
  const double d_o = dmask * s, d_m = sg * i, d_i = sg * m, d_f = sg * sp;
  c.di = i * (1 - i) * d_i;
  c.dm = m * (1 - m) * d_m;
  c.df = f * (1 - f) * d_f;
  c.d_o = o * (1 - o) * d_o;
  c.sgf = sg * f;
  if (R) {
    // downstreamR = u otherR + uR other (arithmetic.go:358-362)
    const double rsg = rcarry + (dmask * ro + rdmask * o);
    const double rd_o = dmask * rs + rdmask * s;
    const double rd_m = sg * ri + rsg * i, rd_i = sg * rm + rsg * m, rd_f = sg * rsp + rsg * sp;
    // sigmoid R-backward, operand order of math_funcs.go:369-370
    c.rdi = ri * (1 - i) * d_i - i * ri * d_i + i * (1 - i) * rd_i;
    c.rdm = rm * (1 - m) * d_m - m * rm * d_m + m * (1 - m) * rd_m;
    c.rdf = rf * (1 - f) * d_f - f * rf * d_f + f * (1 - f) * rd_f;
    c.rd_o = ro * (1 - o) * d_o - o * ro * d_o + o * (1 - o) * rd_o;
    c.rsgf = sg * rf + rsg * f;
  }
  return c;
}

// FAM: the epilogue families an instantiation can meet (the others are not compiled into it -- the tensor-core kernels'
// main loops lose speed with every block of epilogue code around them): 1 = bias / bias + sigmoid (forward),
// 2 = sigmoid backward (input gradient), 4 = plain accumulation (weight gradient, accumulating input gradients);
// plain and atomic stores are always there.
enum : int { FAM_BIAS = 1, FAM_SIGMOID_BWD = 2, FAM_ACCUM = 4, FAM_ALL = 7 };

template <bool DUAL, int FAM = FAM_ALL>
__device__ __forceinline__ void epi_transform(const GemmArgs& p, long long i, int j, double& v0, double& v1) {
  switch (p.epi) {
    case EPI_BIAS:
      if (!(FAM & FAM_BIAS)) break;
      if (p.bias0) v0 += p.bias0[j];
      if (DUAL && p.bias1) v1 += p.bias1[j];
      break;
    case EPI_BIAS_SIGMOID: {
      if (!(FAM & FAM_BIAS)) break;
      double z = v0 + (p.bias0 ? p.bias0[j] : 0.0);
      double s = sigmoid_f64(z);
      v0 = s;
      if (DUAL) {
        double rz = v1 + (p.bias1 ? p.bias1[j] : 0.0);
        v1 = s * (1 - s) * rz;
      }
      break;
    }
    case EPI_SIGMOID_BWD:
    case EPI_SIGMOID_BWD_COLSUM: {
      if (!(FAM & FAM_SIGMOID_BWD)) break;
      // math_funcs.go:365-371, operand order kept: u' = x(1-x)u ; uR' = xR(1-x)u - x xR u + x(1-x)uR
      double x = p.s[i * p.lds + j];
      double u = v0;
      v0 = x * (1 - x) * u;
      if (DUAL) {
        double xr = p.rs ? p.rs[i * p.lds + j] : 0.0;
        v1 = xr * (1 - x) * u - x * xr * u + x * (1 - x) * v1;
      }
      break;
    }
    default:
      break;
  }
}

template <bool DUAL, int FAM = FAM_ALL>
__device__ __forceinline__ void epi_elem(const GemmArgs& p, long long i, int j, double v0, double v1) {
  long long o = p.transpose_out ? (long long)j * p.ldo + i : i * p.ldo + j;
  if ((FAM & FAM_ACCUM) && p.epi == EPI_ACCUM) {
    if (p.out0) p.out0[o] += v0;
    if (DUAL && p.out1) p.out1[o] += v1;
  } else if (p.epi == EPI_ACCUM_ATOMIC) {
    if (p.out0) atomicAdd(p.out0 + o, v0);
    if (DUAL && p.out1) atomicAdd(p.out1 + o, v1);
  } else {
    epi_transform<DUAL, FAM>(p, i, j, v0, v1);
    if (p.out0) p.out0[o] = v0;
    if (DUAL && p.out1) p.out1[o] = v1;
  }
}

// two adjacent columns (j, j+1), j even; 16-byte accesses when the row pitch allows
template <bool DUAL, int FAM = FAM_ALL>
__device__ __forceinline__ void epi_pair(const GemmArgs& p, bool vec_ok, long long i, int j, double a0, double a1,
                                         double r0, double r1) {
  if (j + 1 < p.J && vec_ok) {
    long long o = i * p.ldo + j;
    if ((FAM & FAM_ACCUM) && p.epi == EPI_ACCUM) {
      if (p.out0) {
        double2 c = *reinterpret_cast<double2*>(p.out0 + o);
        c.x += a0; c.y += a1;
        *reinterpret_cast<double2*>(p.out0 + o) = c;
      }
      if (DUAL && p.out1) {
        double2 c = *reinterpret_cast<double2*>(p.out1 + o);
        c.x += r0; c.y += r1;
        *reinterpret_cast<double2*>(p.out1 + o) = c;
      }
    } else if (p.epi == EPI_ACCUM_ATOMIC) {
      if (p.out0) { atomicAdd(p.out0 + o, a0); atomicAdd(p.out0 + o + 1, a1); }
      if (DUAL && p.out1) { atomicAdd(p.out1 + o, r0); atomicAdd(p.out1 + o + 1, r1); }
    } else {
      epi_transform<DUAL, FAM>(p, i, j, a0, r0);
      epi_transform<DUAL, FAM>(p, i, j + 1, a1, r1);
      if (p.out0) *reinterpret_cast<double2*>(p.out0 + o) = make_double2(a0, a1);
      if (DUAL && p.out1) *reinterpret_cast<double2*>(p.out1 + o) = make_double2(r0, r1);
    }
  } else {
    if (j < p.J) epi_elem<DUAL, FAM>(p, i, j, a0, r0);
    if (j + 1 < p.J) epi_elem<DUAL, FAM>(p, i, j + 1, a1, r1);
  }
}

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: try_wait suspends the thread for a hardware-chosen slice per attempt (the back-off); after kMbarSpinLimit
// failed attempts (seconds of wall time -- a healthy wait on a TMA tile lasts microseconds) the thread traps: a TMA that
// faulted or a descriptor that never completes surfaces as a launch failure (AF_ECUDA at the next call) instead of hanging
// the GPU.  The shape of the loop matters: a first version with nanosleep and two exits made ptxas wrap every wait of the
// main loop in BSSY / BSYNC convergence barriers and cost 2.4 % of the bench step (863 k vs 884 k samples/s, A/B in
// profiles/r02_mbar_wait_ab.txt).
constexpr uint32_t kMbarSpinLimit = 1u << 26;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// the retry loop lives in its own function so that the hot path is one TRYWAIT and a not-taken branch
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  for (uint32_t left = kMbarSpinLimit; left != 0; --left)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef AF_UNBOUNDED_MBAR_WAIT  // A/B builds only: the round-1 wait loop
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "AF_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra AF_DONE;\n"
      "bra AF_WAIT;\n"
      "AF_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
  return;
#endif
  if (__builtin_expect(!mbar_try_wait(bar, parity), 0)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// request the line holding `p` into L2 (no register, no dependency): epilogue operands that are known when the CTA starts
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// thread-block cluster helpers (SPECIAL = 2: split-K over a cluster, partial tiles exchanged through DSMEM)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local` (a shared::cta address of this CTA's window) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(local), "r"(rank));
  return a;
}
__device__ __forceinline__ void dsmem_st_f64(uint32_t addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
// D(8x8) += A(8x4, row) * B(4x8, col); lane (g = lane/4, j = lane%4) supplies A[g][j], B[j][g],
// holds C[g][2j], C[g][2j+1].  Lowers to DMMA.8x8x4 on sm_100a.
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Tile configurations.  A CTA is 8 warps laid out WI x WJ; each warp owns TI x TJ DMMA tiles (8x8).
//   CFG 0 "big"    : 4x2 warps of 32x32 -> 128 x 64   (everything with J > 16)
//   CFG 1 "narrow" : 8x1 warps of  8x16 ->  64 x 16   (J <= 16: the 10-wide output layer, and, with
//                    swapped operands + transposed store, its 10-row weight gradient)
template <int CFG> struct TileCfg;
template <> struct TileCfg<0> { static constexpr int WI = 4, WJ = 2, TI = 4, TJ = 4, STAGES = 4, MIN_CTAS = 1, PREFETCH = 1; };
template <> struct TileCfg<1> { static constexpr int WI = 8, WJ = 1, TI = 1, TJ = 2, STAGES = 4, MIN_CTAS = 2, PREFETCH = 1; };
//   CFG 2 "paired" : 2x2 warps of 32x32 ->  64 x 64, 3 stages, TWO CTAs per SM: the two CTAs drift out of
//                    phase, so one CTA's epilogue / pipeline fill overlaps the other's DMMA stream
template <> struct TileCfg<2> { static constexpr int WI = 2, WJ = 2, TI = 4, TJ = 4, STAGES = 3, MIN_CTAS = 2, PREFETCH = 1; };
//   CFG 3 "wide"   : 4x4 warps of 32x16 -> 128 x 64 with FOUR warps per sub-partition (<= 128 registers):
//                    latency is hidden by warp count instead of register double-buffering
template <> struct TileCfg<3> { static constexpr int WI = 4, WJ = 4, TI = 4, TJ = 2, STAGES = 4, MIN_CTAS = 1, PREFETCH = 0; };
//   CFG 4 "step"   : 4x2 warps of 16x32 ->  64 x 64 with EIGHT warps: the LSTM time steps (lstm_step.cu) are 128-tile
//                    launches, one CTA per SM -- two warps per sub-partition cover each other's barrier waits and fragment
//                    loads (with CFG 2's four warps the same launch ran at 60 % of the pipe: 82 vs 50 us per step)
//                    (MIN_CTAS = 2 -- <= 128 registers, so that a programmatically launched successor step can sit beside
//                    the running one and prefetch its weight tiles -- was measured: the register cap costs 3 us per step
//                    (72.5 vs 69.5 us), the forward chain gains 1 us per step from the early launch and the cluster
//                    launches of the backward chain get much slower under programmatic dependent launch (23.5 vs
//                    16.2 ms): one CTA per SM and plain launches stay the default, af_set_option "pdl" 1 is the A/B switch)
template <> struct TileCfg<4> { static constexpr int WI = 4, WJ = 2, TI = 2, TJ = 4, STAGES = 3, MIN_CTAS = 1, PREFETCH = 1; };

template <int NACC, int CFG>
struct GemmSmem {
  using T = TileCfg<CFG>;
  static constexpr int kBMc = T::WI * T::TI * 8;
  static constexpr int kBNc = T::WJ * T::TJ * 8;
  static constexpr uint32_t kABytes = kBMc * kBK * 8;
  static constexpr uint32_t kBBytes = kBNc * kBK * 8;
  static constexpr uint32_t kStageBytes = NACC * (kABytes + kBBytes);
  static constexpr uint32_t kA0 = 0;
  static constexpr uint32_t kA1 = kABytes;                    // NACC == 2 only
  static constexpr uint32_t kB0 = NACC * kABytes;
  static constexpr uint32_t kB1 = NACC * kABytes + kBBytes;   // NACC == 2 only
  static constexpr int kStages = T::STAGES;
  static constexpr int kWarps = T::WI * T::WJ;
  static constexpr int kThreads = kWarps * 32;
  static constexpr uint32_t kBarOffset = kStages * kStageBytes;
  static constexpr uint32_t kTotal = kBarOffset + 2 * kStages * 8 + 1024;  // + alignment slack
  // SPECIAL = 2: exchange buffer of the cluster split-K, [source CTA][warp of the quarter][value][lane] doubles.  It
  // ALIASES the stage ring: it is first written after a cluster barrier that every thread of every CTA of the cluster
  // reaches only when its main loop is over (all TMA tiles landed and consumed), so the kernel needs no extra shared memory
  // and two CTAs fit on an SM.
  static constexpr int kSplit = 4;
  static constexpr uint32_t kXchgOffset = 0;
  static constexpr uint32_t kXchgBytes = kSplit * (kWarps / kSplit) * NACC * T::TI * T::TJ * 2 * 32 * 8;
  static constexpr uint32_t kTotalCluster = kTotal;
  static_assert(kXchgBytes <= kBarOffset, "the exchange buffer must fit inside the stage ring");
  static_assert(kABytes % 1024 == 0 && kBBytes % 1024 == 0, "swizzle atoms must stay 1024-byte aligned");
  static_assert((T::TI == 1 || (T::TI * 8) % 16 == 0) && (T::TJ == 1 || (T::TJ * 8) % 16 == 0), "frag_addr assumes 16-row alignment");
};

template <int LAY, int ROWS>
__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* map, uint32_t bar, int row0, int k0) {
  if (LAY == LAY_KC) {
    tma_load_2d(dst, map, bar, k0, row0);
  } else {
#pragma unroll
    for (int c = 0; c < ROWS / 16; ++c) tma_load_2d(dst + c * 2048, map, bar, row0 + 16 * c, k0);
  }
}

// byte offset of fragment element (row = w0 + g, k) inside a swizzled tile image
template <int LAY>
__device__ __forceinline__ uint32_t frag_base(int w0, int g, int k) {
  const int row = w0 + g;
  if (LAY == LAY_KC) return (uint32_t)(row * 128 + ((((k >> 1) ^ row) & 7) << 4) + ((k & 1) << 3));
  return (uint32_t)((row >> 4) * 2048 + k * 128 + (((((row & 15) >> 1) ^ k) & 7) << 4) + ((row & 1) << 3));
}
// ... of row + 8t (w0 is a multiple of 16 whenever t > 0 is used)
template <int LAY>
__device__ __forceinline__ uint32_t frag_addr(uint32_t base, int t) {
  if (LAY == LAY_KC) return base + t * 1024;
  return (base + (t >> 1) * 2048) ^ ((t & 1) * 64);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// Work decomposition ("stream-K"): the output tiles x k-tiles form one linear sequence of UNITS,
// unit u = (tile u / kt_total, k-tile u % kt_total), tiles ordered j-fastest.  CTA c owns the contiguous
// range [c*U, (c+1)*U).  U == kt_total is the classic one-tile-per-CTA launch (fused epilogue, plain
// stores); any other U lets a CTA end in the middle of a tile or span several, which balances the chip
// for ANY shape (e.g. 512 tiles on 296 CTA slots = 1.73 waves -> every CTA gets 1.73 tiles' worth);
// partial tiles are then combined with red.global.add.f64 (the host zeroes store-type outputs first and
// applies the elementwise step afterwards with gemm_epilogue_kernel).  The TMA ring and the fragment
// prefetch run across tile boundaries, so a new tile's pipeline fill overlaps the previous tile's math.
// MULTI = false: the CTA's range lies inside one tile (one tile per CTA, or aligned split-K): one segment, one
// epilogue after the loop -- the leanest code.  MULTI = true: stream-K, the range may span tiles.
// SPECIAL (lstm_step.cu; one tile per CTA / aligned split only, CFG 2):
//   1  LSTM forward step.  B is the stacked gate matrix read through a 3-D tensor map {k, unit, gate} so that the 64 rows
//      of a tile are {4 gates} x {16 hidden units}, and warp column wj's accumulator block u covers gate u of units
//      8 wj .. 8 wj + 7: every thread's accumulators then hold all four gates of its (lane, unit) elements and the epilogue is the whole cell (bias + x projection read from G, sigmoids, Mul / Add,
//      new state, masked state) -- the step is ONE kernel, nothing is accumulated atomically, no gate buffer is zeroed.
//   2  LSTM backward step.  The contraction over the 4H gate rows is split over a CLUSTER of kSplit CTAs (aligned
//      split-K: CTA r of the cluster takes k-tiles [r kt/4, (r+1) kt/4)); every CTA sends warp w's partial accumulators
//      to CTA w through distributed shared memory, and CTA w adds the four partials in rank order (deterministic, no
//      atomics, no zeroing) and runs the cell backward of the previous time step on its 32 x 32 quarter of the tile.
// OPS: which R operands exist, as a compile-time fact: -1 = read GemmArgs::has_a1 / has_b1 at run time (every k-step then
// branches around the products of an absent operand: eight branches per k-tile that cut the DMMA stream into basic
// blocks), 1 = A1 only, 2 = B1 only, 3 = both.  The one-tile launches of the paired 64 x 64 configuration are instantiated
// per case (gemm.cu), everything else takes -1.
template <int ALAY, int BLAY, int NACC, int CFG, bool MULTI, int SPECIAL = 0, int OPS = -1>
__global__ void __launch_bounds__(TileCfg<CFG>::WI * TileCfg<CFG>::WJ * 32, TileCfg<CFG>::MIN_CTAS)
gemm_dmma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, const GemmArgs p,
                 const SpecialArgs<SPECIAL> sp) {
  using L = GemmSmem<NACC, CFG>;
  using T = TileCfg<CFG>;
  constexpr bool DUAL = (NACC == 2);
  constexpr bool FIXUP = (SPECIAL == 3);                 // stream-K with the ordered fix-up (MULTI only)
  constexpr bool PLAIN = (SPECIAL == 0 || SPECIAL == 3); // not an LSTM step kernel
  static_assert(!FIXUP || MULTI, "the fix-up is a stream-K mode");
  // epilogue families by call site: <KC,KC> forward, <KC,RC> input gradient (incl. accumulating ones), <RC,RC> weight gradient
  constexpr int FAM = (ALAY == LAY_KC && BLAY == LAY_KC) ? FAM_BIAS
                      : (ALAY == LAY_KC && BLAY == LAY_RC) ? (FAM_SIGMOID_BWD | FAM_ACCUM)
                      : (ALAY == LAY_RC && BLAY == LAY_RC) ? FAM_ACCUM : FAM_ALL;
  constexpr int TI = T::TI, TJ = T::TJ, BM = L::kBMc, BN = L::kBNc, kStages = L::kStages, kConsumerWarps = L::kWarps;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle atoms are 1024-B aligned
  const uint32_t bars = base + L::kBarOffset;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt_total = p.kt_total;
  const long long u0 = (long long)blockIdx.x * p.units_per_cta;
  long long u1 = u0 + p.units_per_cta;
  if (u1 > p.total_units) u1 = p.total_units;
  int nun = (int)(u1 - u0);
  const int tile0 = (int)(u0 / p.units_per_tile);
  const int kt0 = (int)(u0 - (long long)tile0 * p.units_per_tile);
  // aligned split-K pads every tile to a multiple of units_per_cta: clip the padding
  if (p.units_per_tile != kt_total && nun > kt_total - kt0) nun = kt_total - kt0;
  if (MULTI && p.wave_tiles > 0) nun = (p.wave_tiles - ((int)blockIdx.x < p.wave_rem ? 0 : 1)) * kt_total;
  if (nun <= 0) return;
  const bool has_a1 = OPS < 0 ? (DUAL && p.has_a1) : (DUAL && (OPS & 1) != 0);
  const bool has_b1 = OPS < 0 ? (DUAL && p.has_b1) : (DUAL && (OPS & 2) != 0);
  // sequence position -> output tile (identity unless the launch is wave-ordered)
  auto tile_of = [&](int q) { return (MULTI && p.wave_tiles > 0) ? (q % p.wave_tiles) * (int)gridDim.x + q / p.wave_tiles : q; };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bars + 8 * s, 1);                           // full: one expect_tx arrival + tx bytes
      mbar_init(bars + 8 * (kStages + s), kConsumerWarps);  // empty: one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (PLAIN) pdl_wait();  // everything above overlapped the predecessor's tail; its output is visible from here on

  // Producer role: lane 0 of warp 0 keeps kStages-1 units in flight.  It is folded into a consumer
  // warp (not an extra warp) because registers are allocated per SM sub-partition: an extra warp would put
  // three warps on one sub-partition and cap every thread at 168 registers, spilling accumulators.
  const uint32_t tx_bytes = L::kABytes + L::kBBytes + (has_a1 ? L::kABytes : 0) + (has_b1 ? L::kBBytes : 0);
  // the producer walks the unit sequence incrementally (no divisions in the loop)
  // Stream-K ranges start at different k-tiles, so CTAs that run side by side would read every operand panel at a
  // different k and nothing would be shared in L2 (ncu: 2.4 GB of DRAM reads for 230 MB of operands).  The ORDER in
  // which a segment's k-tiles are accumulated is free, so with p.rotate a segment [b, b + len) that starts at local
  // time tau (units since the CTA's start; all CTAs of a stream-K launch start together) begins at k-tile b + r and
  // wraps, with r chosen so that k-tile == tau (mod kt_total) for as long as possible: whole tiles are then swept
  // in lockstep by every CTA and a panel's k-slice is fetched from DRAM once for all the tiles that use it
  // (streamk_walk.h: sk_seg_start).
  auto seg_start = [&](int b, int len, int tau) -> int { return sk_seg_start(b, len, tau, kt_total, MULTI && p.rotate); };
#ifndef AF_ROTATE_PRODUCER
#define AF_ROTATE_PRODUCER 0
#endif
  const bool prod_lane = AF_ROTATE_PRODUCER ? (lane == 0) : (threadIdx.x == 0);
  int p_q = tile0, p_tj = tile_of(tile0) % p.tiles_j, p_ti = tile_of(tile0) / p.tiles_j;
  int p_b = kt0, p_len = (kt_total - kt0 < nun) ? kt_total - kt0 : nun, p_done = 0, p_tau = 0;  // current segment (MULTI)
  int p_kt = seg_start(p_b, p_len, 0);
  // the B operand of unit (k-tile kt) into stage `st`; SPECIAL = 1: rows of the tile = {4 gates} x {16 hidden units},
  // box {16 k, 16 units, 4 gates} of the 3-D gate map
  auto load_b = [&](uint32_t st, uint32_t full, int kt) {
    const int k0 = kt * kBK, j0 = p_tj * BN;
    if (SPECIAL == 1) {
      tma_load_3d(st + L::kB0, &tmB0, full, k0, p_tj * 16, 0);
      if (has_b1) tma_load_3d(st + L::kB1, &tmB1, full, k0, p_tj * 16, 0);
    } else {
      tma_load_tile<BLAY, BN>(st + L::kB0, &tmB0, full, j0, k0);
      if (has_b1) tma_load_tile<BLAY, BN>(st + L::kB1, &tmB1, full, j0, k0);
    }
  };
  auto load_a = [&](uint32_t st, uint32_t full, int kt) {
    const int k0 = kt * kBK, i0 = p_ti * BM;
    tma_load_tile<ALAY, BM>(st + L::kA0, &tmA0, full, i0, k0);
    if (has_a1) tma_load_tile<ALAY, BM>(st + L::kA1, &tmA1, full, i0, k0);
  };
  // `issue` = false advances the walk without requesting tiles.  It exists for the ROTATING producer experiment
  // (-DAF_ROTATE_PRODUCER=1: lane 0 of every warp tracks the unit walk and unit t's tiles are requested by lane 0 of warp
  // t mod (warps), so that the empty-barrier wait and the TMA issue do not always stall warp 0's sub-partition).  Measured
  // on one box, alternating (tools/r02_ab_prod.sh): bench 878.0 k vs 892.7 k samples/s with the fixed producer, C3 31.95
  // vs 31.42 ms, C4 26.9 vs 25.2 ms -- every warp now pays the walk's bookkeeping each unit, which costs more than
  // spreading the stall returns.  The fixed producer (lane 0 of warp 0) stays.
  auto produce = [&](int t, bool b_loaded = false, bool issue = true) {
    const int s = t % kStages;
    const uint32_t ph = (t / kStages) & 1;
    const uint32_t full = bars + 8 * s;
    const uint32_t st = base + s * L::kStageBytes;
    if (issue) {
      if (!b_loaded) {
        mbar_wait(bars + 8 * (kStages + s), ph ^ 1);  // all warps released the stage's previous unit
        mbar_expect_tx(full, tx_bytes);
        load_b(st, full, p_kt);
      }
      load_a(st, full, p_kt);
    }
    ++p_kt;
    if (MULTI) {
      if (p_kt == p_b + p_len) p_kt = p_b;  // wrap inside the segment (a rotated walk; otherwise the segment ends here)
      if (++p_done == p_len) {              // next tile of this CTA's range
        p_tau += p_len;
        p_b = 0; p_done = 0;
        p_len = (kt_total < nun - p_tau) ? kt_total : nun - p_tau;
        p_kt = seg_start(0, p_len, p_tau);
        ++p_q;
        if (p.wave_tiles > 0) {
          const int tl = tile_of(p_q);
          p_tj = tl % p.tiles_j; p_ti = tl / p.tiles_j;
        } else if (++p_tj == p.tiles_j) { p_tj = 0; ++p_ti; }
      }
    }
  };
  if (!PLAIN) {
    // LSTM time steps are chains of dependent launches (programmatic dependent launch, lstm_step.cu): this CTA may have
    // become resident while the previous step is still computing.  The B operand -- the weights -- does not depend on
    // it, so the first stages' weight tiles are requested BEFORE griddepcontrol.wait and only the A tiles (the state the
    // previous step is producing) after it: launch latency, barrier set-up and half of the pipeline fill leave the chain.
    if (threadIdx.x == 0) {
      for (int t = 0; t < kStages - 1 && t < nun; ++t) {
        mbar_expect_tx(bars + 8 * t, tx_bytes);
        load_b(base + t * L::kStageBytes, bars + 8 * t, p_kt + t);  // (!MULTI: units are consecutive k-tiles of one tile)
      }
    }
    // The cell epilogue reads, at this thread's output positions, vectors that were written long before this step -- the x
    // projection / the forward pass's gates and states / the output layer's gradient: hundreds of MB ago, so they come
    // from DRAM -- and the step cannot end before they arrive.  They do not depend on the previous step either: request
    // them into L2 now, the main loop covers the latency.
    if constexpr (SPECIAL == 1) {
      const LstmStepArgs& c = sp.lstm;
      const int tl = tile_of(tile0);
      const int unit = (tl % p.tiles_j) * 16 + (warp % T::WJ) * 8 + 2 * (lane & 3);
#pragma unroll
      for (int t = 0; t < TI; ++t) {
        const long long i = (long long)(tl / p.tiles_j) * BM + (warp / T::WJ) * (TI * 8) + 8 * t + (lane >> 2);
        if (i < p.I) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            prefetch_l2(c.G + i * 4 * c.H + unit + u * c.H);
            if (DUAL) prefetch_l2(c.RG + i * 4 * c.H + unit + u * c.H);
          }
        }
      }
    }
    if constexpr (SPECIAL == 2) {
      constexpr int NQ = kConsumerWarps / L::kSplit, NV1 = TI * TJ * 2, PER = NQ * NV1 / kConsumerWarps;
      const LstmStepArgs& c = sp.lstm;
      const int tl = tile_of(tile0);
      const int sw = (int)cluster_ctarank() * NQ + (warp * PER) / NV1;
#pragma unroll
      for (int x = 0; x < PER; x += 2) {
        const int v = (warp * PER) % NV1 + x;
        const long long i = (long long)(tl / p.tiles_j) * BM + (sw / T::WJ) * (TI * 8) + 8 * (v / (TJ * 2)) + (lane >> 2);
        const int j = (tl % p.tiles_j) * BN + (sw % T::WJ) * (TJ * 8) + 8 * ((v / 2) % TJ) + 2 * (lane & 3);
        if (i < p.I && j < p.J) {
          const long long gi = i * 4 * c.H + j, si = i * c.ld + j;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            prefetch_l2(c.Gm + gi + u * c.H);
            if (DUAL) prefetch_l2(c.RGm + gi + u * c.H);
          }
          prefetch_l2(c.INm + si); prefetch_l2(c.INt + si); prefetch_l2(c.DAUGm + si);
          if (DUAL) { prefetch_l2(c.RINm + si); prefetch_l2(c.RINt + si); prefetch_l2(c.RDAUGm + si); }
        }
      }
    }
    pdl_wait();
    if (prod_lane) {
      for (int t = 0; t < kStages - 1 && t < nun; ++t) produce(t, true, warp == 0);
    }
  } else if (prod_lane) {
    for (int t = 0; t < kStages - 1 && t < nun; ++t) produce(t, false, warp == 0);
  }

  // ---------------- consumers: WI (i) x WJ (j) warps ----------------
  const int g = lane >> 2, j4 = lane & 3;
  const int wi0 = (warp / T::WJ) * (TI * 8), wj0 = (warp % T::WJ) * (TJ * 8);
  const int kj = ((j4 & 1) ? 3 : 0) | ((j4 & 2) ? 12 : 0);
  uint32_t fa[4], fb[4];
  {
    const int ms[4] = {0, 1, 4, 5};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      fa[s] = frag_base<ALAY>(wi0, g, kj ^ ms[s]);
      fb[s] = frag_base<BLAY>(SPECIAL == 1 ? (warp % T::WJ) * 8 : wj0, g, kj ^ ms[s]);
    }
  }

  double acc0[TI][TJ][2], acc1[DUAL ? TI : 1][TJ][2];
  auto zero_acc = [&]() {
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
      for (int u = 0; u < TJ; ++u) {
        acc0[t][u][0] = acc0[t][u][1] = 0.0;
        if (DUAL) acc1[t][u][0] = acc1[t][u][1] = 0.0;
      }
  };
  zero_acc();

  // Fragment registers are double-buffered by hand ACROSS unit boundaries: the loads of MMA step q+1
  // (possibly the first step of the next stage) are issued before the DMMAs of step q.
  constexpr int NB = T::PREFETCH ? 2 : 1;
  double a0[NB][TI], b0[NB][TJ], a1[NB][DUAL ? TI : 1], b1[NB][DUAL ? TJ : 1];
  auto load_frags = [&](int buf, uint32_t st, int ks) {
#pragma unroll
    for (int t = 0; t < TI; ++t) a0[buf][t] = lds_f64(st + L::kA0 + frag_addr<ALAY>(fa[ks], t));
#pragma unroll
    for (int u = 0; u < TJ; ++u) b0[buf][u] = lds_f64(st + L::kB0 + (SPECIAL == 1 ? fb[ks] + u * 2048 : frag_addr<BLAY>(fb[ks], u)));
    if (has_a1) {
#pragma unroll
      for (int t = 0; t < TI; ++t) a1[buf][DUAL ? t : 0] = lds_f64(st + L::kA1 + frag_addr<ALAY>(fa[ks], t));
    }
    if (has_b1) {
#pragma unroll
      for (int u = 0; u < TJ; ++u)
        b1[buf][DUAL ? u : 0] = lds_f64(st + L::kB1 + (SPECIAL == 1 ? fb[ks] + u * 2048 : frag_addr<BLAY>(fb[ks], u)));
    }
  };
  auto mma_step = [&](int buf) {
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
      for (int u = 0; u < TJ; ++u) dmma_8x8x4(acc0[t][u][0], acc0[t][u][1], a0[buf][t], b0[buf][u]);
    if (has_b1) {
#pragma unroll
      for (int t = 0; t < TI; ++t)
#pragma unroll
        for (int u = 0; u < TJ; ++u)
          dmma_8x8x4(acc1[DUAL ? t : 0][u][0], acc1[DUAL ? t : 0][u][1], a0[buf][t], b1[buf][DUAL ? u : 0]);
    }
    if (has_a1) {
#pragma unroll
      for (int t = 0; t < TI; ++t)
#pragma unroll
        for (int u = 0; u < TJ; ++u)
          dmma_8x8x4(acc1[DUAL ? t : 0][u][0], acc1[DUAL ? t : 0][u][1], a1[buf][DUAL ? t : 0], b0[buf][u]);
    }
  };

  // registers -> global for the tile that just ended (or was interrupted by the end of this CTA's range)
  const bool vec_ok = !p.transpose_out && ((p.ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.out0) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.out1) & 15) == 0);
  // kind (FIXUP only): 0 = the segment is a whole tile, 1 = it is the TAIL of a tile whose first k-tiles belong to the
  // previous CTA (computed first in time: park the partial sums), 2 = it is the HEAD of a tile that the next CTA finishes
  // (computed last: fetch that CTA's partial, add, run the real epilogue)
  auto flush = [&](int q, int kind = 0, int head_len = 0) {
    const int tile = tile_of(q);
    const int i0 = (tile / p.tiles_j) * BM, j0 = (tile % p.tiles_j) * BN;
    if constexpr (FIXUP) {
      // Stream-K without atomics.  A tile that is split has one HEAD -- the CTA whose range ends inside it, holding its first
      // k-tiles -- and one or more followers: the next CTA(s), whose ranges begin (or lie entirely) inside it.  A follower
      // meets the tile at the very start of its life: it parks its partial accumulators in the workspace -- thread-private
      // slots, [CTA][value][thread], no index arithmetic on the other side; a CTA parks at most one partial, its first
      // segment -- and raises its warp's flag.  The head meets the tile at the very end of its life: it takes the parked
      // values in CTA order, adds them to its own (one fixed order, the same bits on every run) and runs the ordinary fused
      // epilogue with plain stores.
      // No zeroing pass, no red.global.add.f64, no deferred elementwise kernel.  All CTAs of the launch are co-resident or
      // dispatched in index order, so the wait cannot deadlock; it is bounded like the mbarrier waits.
      constexpr int NV1 = TI * TJ * 2, NV = NACC * NV1, NT = kConsumerWarps * 32;
      if (kind == 1) {
        double* mine = sp.fix.part + (size_t)blockIdx.x * (NV * NT) + threadIdx.x;
#pragma unroll
        for (int t = 0; t < TI; ++t)
#pragma unroll
          for (int u = 0; u < TJ; ++u)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int v = (t * TJ + u) * 2 + e;
              __stcg(mine + v * NT, acc0[t][u][e]);
              if (DUAL) __stcg(mine + (NV1 + v) * NT, acc1[DUAL ? t : 0][u][e]);
            }
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(sp.fix.flags + blockIdx.x * kConsumerWarps + warp), "r"(1u) : "memory");
        return;
      }
      if (kind == 2) {
        // the followers: CTA c + 1 holds the next units_per_cta k-tiles of this tile (all of its range, if the tile is
        // longer than a range), c + 2 the ones after that, ... until the tile's last k-tile; their partials are added in
        // that order
        const int followers = (kt_total - head_len + p.units_per_cta - 1) / p.units_per_cta;
        for (int fo = 1; fo <= followers; ++fo) {
          unsigned* f = sp.fix.flags + (blockIdx.x + fo) * kConsumerWarps + warp;
          unsigned seen = 0;
          for (uint32_t left = kMbarSpinLimit; left != 0; --left) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
            if (seen) break;
            __nanosleep(64);
          }
          if (!seen) __trap();
          const double* theirs = sp.fix.part + (size_t)(blockIdx.x + fo) * (NV * NT) + threadIdx.x;
#pragma unroll
          for (int t = 0; t < TI; ++t)
#pragma unroll
            for (int u = 0; u < TJ; ++u)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int v = (t * TJ + u) * 2 + e;
                acc0[t][u][e] += __ldcg(theirs + v * NT);
                if (DUAL) acc1[DUAL ? t : 0][u][e] += __ldcg(theirs + (NV1 + v) * NT);
              }
          __syncwarp();
          if (lane == 0) *f = 0;  // (one writer and one reader per launch: ready for the next launch or graph replay)
        }
      }
    }
    if constexpr (SPECIAL == 1) {
      // ---- LSTM forward cell (see LstmStepArgs): accumulator block u of this thread IS gate u of its two units ----
      static_assert(SPECIAL != 1 || (TJ == 4 && T::WJ == 2 && BLAY == LAY_KC), "the gate layout assumes 4 blocks of 8 units per warp");
      const LstmStepArgs& c = sp.lstm;
      const int H = c.H;
      const int unit = (tile % p.tiles_j) * 16 + (warp % T::WJ) * 8 + 2 * j4;  // and unit + 1
      // all rows' operands first (G / RG are read and written in place, so the compiler cannot move row t + 1's loads above
      // row t's stores on its own: one round trip instead of TI)
      double2 zx[TI][4], rzx[TI][4], spv[TI], rspv[TI];
#pragma unroll
      for (int t = 0; t < TI; ++t) {
        const long long i = i0 + wi0 + 8 * t + g;
        spv[t] = rspv[t] = make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 4; ++u) zx[t][u] = rzx[t][u] = make_double2(0.0, 0.0);
        if (i < p.I) {
          const double* gp = c.G + i * 4 * H + unit;
          const long long si = i * c.ld + unit;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            zx[t][u] = *reinterpret_cast<const double2*>(gp + u * H);  // x W_x^T + b of gate u
            if (DUAL) rzx[t][u] = *reinterpret_cast<const double2*>(c.RG + i * 4 * H + unit + u * H);
          }
          spv[t] = *reinterpret_cast<const double2*>(c.INp + si);
          if (DUAL) rspv[t] = *reinterpret_cast<const double2*>(c.RINp + si);
        }
      }
#pragma unroll
      for (int t = 0; t < TI; ++t) {
        const long long i = i0 + wi0 + 8 * t + g;
        if (i < p.I) {
          double* gp = c.G + i * 4 * H + unit;
          const long long si = i * c.ld + unit;
          double2 z[4], rz[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            z[u] = make_double2(acc0[t][u][0] + zx[t][u].x, acc0[t][u][1] + zx[t][u].y);
            if (DUAL) rz[u] = make_double2(acc1[DUAL ? t : 0][u][0] + rzx[t][u].x, acc1[DUAL ? t : 0][u][1] + rzx[t][u].y);
            else rz[u] = make_double2(0.0, 0.0);
          }
          const double2 sp = spv[t], rsp = rspv[t];
          const LstmCellFwd a = lstm_cell_fwd<DUAL>(z[0].x, z[1].x, z[2].x, z[3].x, sp.x, rz[0].x, rz[1].x, rz[2].x, rz[3].x, rsp.x);
          const LstmCellFwd b = lstm_cell_fwd<DUAL>(z[0].y, z[1].y, z[2].y, z[3].y, sp.y, rz[0].y, rz[1].y, rz[2].y, rz[3].y, rsp.y);
          *reinterpret_cast<double2*>(gp) = make_double2(a.i, b.i);
          *reinterpret_cast<double2*>(gp + H) = make_double2(a.m, b.m);
          *reinterpret_cast<double2*>(gp + 2 * H) = make_double2(a.f, b.f);
          *reinterpret_cast<double2*>(gp + 3 * H) = make_double2(a.o, b.o);
          *reinterpret_cast<double2*>(c.INn + si) = make_double2(a.s, b.s);
          *reinterpret_cast<double2*>(c.AUG + si) = make_double2(a.masked, b.masked);
          if (DUAL) {
            double* rgp = c.RG + i * 4 * H + unit;
            *reinterpret_cast<double2*>(rgp) = make_double2(a.ri, b.ri);
            *reinterpret_cast<double2*>(rgp + H) = make_double2(a.rm, b.rm);
            *reinterpret_cast<double2*>(rgp + 2 * H) = make_double2(a.rf, b.rf);
            *reinterpret_cast<double2*>(rgp + 3 * H) = make_double2(a.ro, b.ro);
            *reinterpret_cast<double2*>(c.RINn + si) = make_double2(a.rs, b.rs);
            *reinterpret_cast<double2*>(c.RAUG + si) = make_double2(a.rmasked, b.rmasked);
          }
        }
      }
      return;
    }
    if constexpr (SPECIAL == 2) {
      // ---- cluster split-K: exchange the partial tiles through DSMEM, then the LSTM cell backward ----
      // The tile is cut into kSplit "quarters" of NQ consecutive warps each; CTA q of the cluster finishes quarter q.
      constexpr int NQ = kConsumerWarps / L::kSplit;     // warps per quarter
      constexpr int NV1 = TI * TJ * 2, NV = NACC * NV1;  // accumulator values per thread (one / both accumulators)
      constexpr int PER = NQ * NV1 / kConsumerWarps;     // values of one accumulator each warp finishes
      static_assert(SPECIAL != 2 || (kConsumerWarps % L::kSplit == 0 && PER >= 2 && PER % 2 == 0 && NV1 % PER == 0), "quarter layout");
      const LstmStepArgs& c = sp.lstm;
      const uint32_t xchg = base + L::kXchgOffset;
      const uint32_t rank = cluster_ctarank();
      const int H = c.H;
      const int first = warp * PER;                 // this warp's chunk of the quarter's NQ x NV1 values
      const int sub = first / NV1;                  // which warp of the quarter produced them
      const int sw = (int)rank * NQ + sub;          // ... and its position in the tile
      const int sw_i0 = (sw / T::WJ) * (TI * 8), sw_j0 = (sw % T::WJ) * (TJ * 8);
      // The cell's operands at this thread's output positions do not depend on the exchange: request them BEFORE the two
      // cluster barriers (and all pairs before the first store -- SGF is read and written in place, so the compiler cannot
      // hoist the second pair's loads above the first pair's stores): their latency, requested into L2 when the CTA started,
      // hides behind the exchange instead of following it pair by pair.
      constexpr int NX = PER / 2;
      struct CellOperands { double2 carry, gI, gM, gF, gO, sp, sn, dm, rcarry, rI, rM, rF, rO, rsp, rsn, rdm; };
      CellOperands op[NX];
#pragma unroll
      for (int xi = 0; xi < NX; ++xi) {
        const int v = first % NV1 + 2 * xi;         // even: the pair (e = 0, 1) of one 8 x 8 block
        const int t = v / (TJ * 2), u = (v / 2) % TJ;
        const long long i = i0 + sw_i0 + 8 * t + g;
        const int j = j0 + sw_j0 + 8 * u + 2 * j4;  // and j + 1
        CellOperands& o = op[xi];
        o.carry = o.gI = o.gM = o.gF = o.gO = o.sp = o.sn = o.dm = make_double2(0.0, 0.0);
        o.rcarry = o.rI = o.rM = o.rF = o.rO = o.rsp = o.rsn = o.rdm = make_double2(0.0, 0.0);
        if (i >= p.I || j >= p.J) continue;
        const long long ei = i * H + j, gi = i * 4 * H + j, si = i * c.ld + j;
        o.carry = *reinterpret_cast<const double2*>(c.SGF + ei);
        o.gI = *reinterpret_cast<const double2*>(c.Gm + gi); o.gM = *reinterpret_cast<const double2*>(c.Gm + gi + H);
        o.gF = *reinterpret_cast<const double2*>(c.Gm + gi + 2 * H); o.gO = *reinterpret_cast<const double2*>(c.Gm + gi + 3 * H);
        o.sp = *reinterpret_cast<const double2*>(c.INm + si); o.sn = *reinterpret_cast<const double2*>(c.INt + si);
        o.dm = *reinterpret_cast<const double2*>(c.DAUGm + si);
        if (DUAL) {
          o.rcarry = *reinterpret_cast<const double2*>(c.RSGF + ei);
          o.rI = *reinterpret_cast<const double2*>(c.RGm + gi); o.rM = *reinterpret_cast<const double2*>(c.RGm + gi + H);
          o.rF = *reinterpret_cast<const double2*>(c.RGm + gi + 2 * H); o.rO = *reinterpret_cast<const double2*>(c.RGm + gi + 3 * H);
          o.rsp = *reinterpret_cast<const double2*>(c.RINm + si); o.rsn = *reinterpret_cast<const double2*>(c.RINt + si);
          o.rdm = *reinterpret_cast<const double2*>(c.RDAUGm + si);
        }
      }
      cluster_sync();  // every CTA of the cluster is running and past its main loop: exchange buffers may be written
      {
        // this warp's partial sums go to the CTA that owns its quarter: slot [source rank][warp within quarter][value][lane]
        const uint32_t dst = dsmem_addr(xchg, (uint32_t)(warp / NQ)) + (((rank * NQ + (warp % NQ)) * NV) * 32 + lane) * 8;
#pragma unroll
        for (int t = 0; t < TI; ++t)
#pragma unroll
          for (int u = 0; u < TJ; ++u)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int v = (t * TJ + u) * 2 + e;
              dsmem_st_f64(dst + v * 256, acc0[t][u][e]);
              if (DUAL) dsmem_st_f64(dst + (NV1 + v) * 256, acc1[DUAL ? t : 0][u][e]);
            }
      }
      cluster_sync();  // all partials have landed (release / acquire at cluster scope)
#pragma unroll
      for (int xi = 0; xi < NX; ++xi) {
        const int v = first % NV1 + 2 * xi;
        const int t = v / (TJ * 2), u = (v / 2) % TJ;
        const long long i = i0 + sw_i0 + 8 * t + g;
        const int j = j0 + sw_j0 + 8 * u + 2 * j4;
        if (i >= p.I || j >= p.J) continue;
        double s0[2], s1[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double a = 0.0, r = 0.0;
#pragma unroll
          for (int src = 0; src < L::kSplit; ++src) {  // rank order: the sum is the same on every run
            const uint32_t slot = xchg + (((src * NQ + sub) * NV + v + e) * 32 + lane) * 8;
            a += lds_f64(slot);
            if (DUAL) r += lds_f64(slot + NV1 * 256);
          }
          s0[e] = a; s1[e] = r;
        }
        const long long ei = i * H + j, gi = i * 4 * H + j;
        const CellOperands& o = op[xi];
        const LstmCellBwd a = lstm_cell_bwd<DUAL>(o.gI.x, o.gM.x, o.gF.x, o.gO.x, o.sp.x, o.sn.x, o.dm.x, o.carry.x + s0[0], o.rI.x, o.rM.x,
                                                  o.rF.x, o.rO.x, o.rsp.x, o.rsn.x, o.rdm.x, o.rcarry.x + s1[0]);
        const LstmCellBwd b = lstm_cell_bwd<DUAL>(o.gI.y, o.gM.y, o.gF.y, o.gO.y, o.sp.y, o.sn.y, o.dm.y, o.carry.y + s0[1], o.rI.y, o.rM.y,
                                                  o.rF.y, o.rO.y, o.rsp.y, o.rsn.y, o.rdm.y, o.rcarry.y + s1[1]);
        *reinterpret_cast<double2*>(c.D4m + gi) = make_double2(a.di, b.di);
        *reinterpret_cast<double2*>(c.D4m + gi + H) = make_double2(a.dm, b.dm);
        *reinterpret_cast<double2*>(c.D4m + gi + 2 * H) = make_double2(a.df, b.df);
        *reinterpret_cast<double2*>(c.D4m + gi + 3 * H) = make_double2(a.d_o, b.d_o);
        *reinterpret_cast<double2*>(c.SGF + ei) = make_double2(a.sgf, b.sgf);
        if (DUAL) {
          *reinterpret_cast<double2*>(c.RD4m + gi) = make_double2(a.rdi, b.rdi);
          *reinterpret_cast<double2*>(c.RD4m + gi + H) = make_double2(a.rdm, b.rdm);
          *reinterpret_cast<double2*>(c.RD4m + gi + 2 * H) = make_double2(a.rdf, b.rdf);
          *reinterpret_cast<double2*>(c.RD4m + gi + 3 * H) = make_double2(a.rd_o, b.rd_o);
          *reinterpret_cast<double2*>(c.RSGF + ei) = make_double2(a.rsgf, b.rsgf);
        }
      }
      return;
    }
#ifndef AF_PERSIST_TILES
#define AF_PERSIST_TILES 0  // 1: compile the persistent whole-tile mode (af_set_option "persist_tiles"; measured slower) into the stream-K kernels
#endif
    if constexpr (PLAIN && MULTI && !AF_PERSIST_TILES && !FIXUP) {
      // stream-K: tile segments always meet in red.global.add.f64 (the elementwise step, if any, follows in
      // gemm_epilogue_kernel).  Nothing else is compiled into these instantiations: their main loops lose 2-3 % whenever
      // the epilogue code around them grows (register allocation is per kernel).
#pragma unroll
      for (int t = 0; t < TI; ++t) {
        const long long i = i0 + wi0 + 8 * t + g;
        if (i < p.I) {
#pragma unroll
          for (int u = 0; u < TJ; ++u) {
            const int j = j0 + wj0 + 8 * u + 2 * j4;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (j + e < p.J) {
                const long long o = p.transpose_out ? (long long)(j + e) * p.ldo + i : i * p.ldo + j + e;
                if (p.out0) atomicAdd(p.out0 + o, acc0[t][u][e]);
                if (DUAL && p.out1) atomicAdd(p.out1 + o, acc1[DUAL ? t : 0][u][e]);
              }
            }
          }
        }
      }
      return;
    }
    if constexpr (PLAIN && (FIXUP || !(MULTI && !AF_PERSIST_TILES))) {
      if ((!MULTI || FIXUP) && ALAY == LAY_KC && BLAY == LAY_KC && p.epi == EPI_BIAS_SIGMOID && vec_ok && i0 + BM <= p.I && j0 + BN <= p.J) {
        // dense forward on an interior tile (lin_add.go:15-17,37-42 + math_funcs.go:292,305-309): one straight-line
        // batch per fragment row -- TJ x 2 independent sigmoids in flight -- instead of the generic path below, which
        // decides the epilogue kind and the bounds anew for every column pair and so runs one exp / reciprocal chain at
        // a time (a 64 x 64 tile is 32 such chains per thread, with the CTA issuing no DMMA meanwhile).  Compiled into the
        // one-tile forward instantiation only: the other instantiations' main loops are sensitive to any growth of their
        // epilogues (the stream-K launches lost 20-25 us each with this block present and never taken).
        double2 bz[TJ], rbz[TJ];
#pragma unroll
        for (int u = 0; u < TJ; ++u) {
          const int j = j0 + wj0 + 8 * u + 2 * j4;
          bz[u] = p.bias0 ? make_double2(__ldg(p.bias0 + j), __ldg(p.bias0 + j + 1)) : make_double2(0.0, 0.0);
          rbz[u] = (DUAL && p.bias1) ? make_double2(__ldg(p.bias1 + j), __ldg(p.bias1 + j + 1)) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int t = 0; t < TI; ++t) {
          const long long row = (i0 + wi0 + 8 * t + g) * p.ldo + j0 + wj0 + 2 * j4;
          double2 sg[TJ];
#pragma unroll
          for (int u = 0; u < TJ; ++u)
            sg[u] = make_double2(sigmoid_f64(acc0[t][u][0] + bz[u].x), sigmoid_f64(acc0[t][u][1] + bz[u].y));
#pragma unroll
          for (int u = 0; u < TJ; ++u) {
            if (p.out0) *reinterpret_cast<double2*>(p.out0 + row + 8 * u) = sg[u];
            if (DUAL && p.out1)
              *reinterpret_cast<double2*>(p.out1 + row + 8 * u) =
                  make_double2(sg[u].x * (1 - sg[u].x) * (acc1[DUAL ? t : 0][u][0] + rbz[u].x),
                               sg[u].y * (1 - sg[u].y) * (acc1[DUAL ? t : 0][u][1] + rbz[u].y));
          }
        }
        return;
      }
      if ((FAM & FAM_ACCUM) && p.epi == EPI_ACCUM && vec_ok) {
        // beta = 1 accumulation (lin_tran.go:210,267-268) with plain stores: fetch a whole fragment row of C before adding
        // to it.  Written pair by pair (load, add, store, load, ...) the stores fence the next load -- the outputs may alias
        // as far as the compiler knows -- and a thread paid 32 serialised L2 round trips per tile.
#pragma unroll
        for (int t = 0; t < TI; ++t) {
          const long long i = i0 + wi0 + 8 * t + g;
          if (i >= p.I) continue;
          double2 c0[TJ], c1[TJ];
#pragma unroll
          for (int u = 0; u < TJ; ++u) {
            const int j = j0 + wj0 + 8 * u + 2 * j4;
            c0[u] = c1[u] = make_double2(0.0, 0.0);
            if (j + 1 < p.J) {
              if (p.out0) c0[u] = *reinterpret_cast<const double2*>(p.out0 + i * p.ldo + j);
              if (DUAL && p.out1) c1[u] = *reinterpret_cast<const double2*>(p.out1 + i * p.ldo + j);
            }
          }
#pragma unroll
          for (int u = 0; u < TJ; ++u) {
            const int j = j0 + wj0 + 8 * u + 2 * j4;
            if (j + 1 < p.J) {
              if (p.out0)
                *reinterpret_cast<double2*>(p.out0 + i * p.ldo + j) = make_double2(c0[u].x + acc0[t][u][0], c0[u].y + acc0[t][u][1]);
              if (DUAL && p.out1)
                *reinterpret_cast<double2*>(p.out1 + i * p.ldo + j) =
                    make_double2(c1[u].x + acc1[DUAL ? t : 0][u][0], c1[u].y + acc1[DUAL ? t : 0][u][1]);
            } else if (j < p.J) {
              epi_pair<DUAL, FAM>(p, vec_ok, i, j, acc0[t][u][0], acc0[t][u][1], DUAL ? acc1[DUAL ? t : 0][u][0] : 0.0,
                             DUAL ? acc1[DUAL ? t : 0][u][1] : 0.0);
            }
          }
        }
        return;
      }
      if constexpr (ALAY == LAY_KC && BLAY == LAY_RC && (!MULTI || FIXUP)) {  // (whole tiles per CTA, or the fix-up's finished tiles)
        // ---- sigmoid backward of the layer below on the finished tile (input-gradient launches; math_funcs.go:365-371,
        // operand order kept), optionally with the tile's column sums (EPI_SIGMOID_BWD_COLSUM) ----
        // S / RS at the output positions are usually cold (written by the forward pass, hundreds of MB ago).  Fetched pair by
        // pair between the stores -- the outputs may alias them as far as the compiler knows -- a thread paid TI x TJ
        // serialised DRAM round trips per tile while its CTA issued no DMMA.  Here a whole fragment row (TJ pairs of both
        // vectors) is requested through the read-only path BEFORE the previous row is computed and stored: TI round
        // trips, each hidden behind a row's arithmetic.  (Requesting the tile into L2 when the CTA starts was measured on
        // top of this and lost 6 us on the bench's layer-2 launch, 793.6 vs 787.5 us: the prefetches compete with the
        // first operand tiles.)
        const bool sig = p.epi == EPI_SIGMOID_BWD || p.epi == EPI_SIGMOID_BWD_COLSUM;
        if (sig && vec_ok && (p.lds & 1) == 0 && ((reinterpret_cast<uintptr_t>(p.s) | reinterpret_cast<uintptr_t>(p.rs)) & 15) == 0) {
          const bool with_colsum = p.epi == EPI_SIGMOID_BWD_COLSUM;
          const bool has_rs = DUAL && p.rs != nullptr;
          double cs0[TJ][2], cs1[TJ][2];
#pragma unroll
          for (int u = 0; u < TJ; ++u) cs0[u][0] = cs0[u][1] = cs1[u][0] = cs1[u][1] = 0.0;
          double2 sx[2][TJ], rx[2][TJ];
          auto fetch = [&](int buf, int t) {
            const long long i = i0 + wi0 + 8 * t + g;
#pragma unroll
            for (int u = 0; u < TJ; ++u) {
              const int j = j0 + wj0 + 8 * u + 2 * j4;
              sx[buf][u] = rx[buf][u] = make_double2(0.0, 0.0);
              if (i < p.I && j < p.J) {
                const long long o = i * p.lds + j;
                if (j + 1 < p.J) {
                  sx[buf][u] = __ldg(reinterpret_cast<const double2*>(p.s + o));
                  if (has_rs) rx[buf][u] = __ldg(reinterpret_cast<const double2*>(p.rs + o));
                } else {
                  sx[buf][u].x = __ldg(p.s + o);
                  if (has_rs) rx[buf][u].x = __ldg(p.rs + o);
                }
              }
            }
          };
          fetch(0, 0);
#pragma unroll
          for (int t = 0; t < TI; ++t) {
            if (t + 1 < TI) fetch((t + 1) & 1, t + 1);
            const long long i = i0 + wi0 + 8 * t + g;
            if (i >= p.I) continue;
#pragma unroll
            for (int u = 0; u < TJ; ++u) {
              const int j = j0 + wj0 + 8 * u + 2 * j4;
              if (j >= p.J) continue;
              const bool two = j + 1 < p.J;
              const double2 x = sx[t & 1][u], xr = rx[t & 1][u];
              const double u0 = acc0[t][u][0], u1 = acc0[t][u][1];
              // u' = x(1-x)u ; uR' = xR(1-x)u - x xR u + x(1-x)uR
              const double a0 = x.x * (1 - x.x) * u0, a1 = x.y * (1 - x.y) * u1;
              double r0 = 0.0, r1 = 0.0;
              if (DUAL) {
                r0 = xr.x * (1 - x.x) * u0 - x.x * xr.x * u0 + x.x * (1 - x.x) * acc1[DUAL ? t : 0][u][0];
                r1 = xr.y * (1 - x.y) * u1 - x.y * xr.y * u1 + x.y * (1 - x.y) * acc1[DUAL ? t : 0][u][1];
              }
              const long long o = i * p.ldo + j;
              if (two) {
                if (p.out0) *reinterpret_cast<double2*>(p.out0 + o) = make_double2(a0, a1);
                if (DUAL && p.out1) *reinterpret_cast<double2*>(p.out1 + o) = make_double2(r0, r1);
              } else {
                if (p.out0) p.out0[o] = a0;
                if (DUAL && p.out1) p.out1[o] = r0;
              }
              cs0[u][0] += a0; cs1[u][0] += r0;
              if (two) { cs0[u][1] += a1; cs1[u][1] += r1; }
            }
          }
          if (with_colsum) {
            // a thread has added its TI rows; three butterfly steps add the eight row lanes, lanes 0..3 send one
            // red.global.add.f64 per column (the targets travel in bias0 / bias1, see EPI_SIGMOID_BWD_COLSUM)
            double* colsum0 = const_cast<double*>(p.bias0);
            double* colsum1 = const_cast<double*>(p.bias1);
#pragma unroll
            for (int u = 0; u < TJ; ++u)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                double a = cs0[u][e], r = cs1[u][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                  a += __shfl_xor_sync(0xffffffffu, a, o);
                  if (DUAL) r += __shfl_xor_sync(0xffffffffu, r, o);
                }
                const int j = j0 + wj0 + 8 * u + 2 * j4 + e;
                if (g == 0 && j < p.J) {
                  if (colsum0) atomicAdd(colsum0 + j, a);
                  if (DUAL && colsum1) atomicAdd(colsum1 + j, r);
                }
              }
          }
          return;
        }
      }
#pragma unroll
      for (int t = 0; t < TI; ++t) {
        const long long i = i0 + wi0 + 8 * t + g;
        if (i < p.I) {
#pragma unroll
          for (int u = 0; u < TJ; ++u) {
            const int j = j0 + wj0 + 8 * u + 2 * j4;
            if (j < p.J)
              epi_pair<DUAL, FAM>(p, vec_ok, i, j, acc0[t][u][0], acc0[t][u][1], DUAL ? acc1[DUAL ? t : 0][u][0] : 0.0,
                             DUAL ? acc1[DUAL ? t : 0][u][1] : 0.0);
          }
        }
      }
    }
  };

  // Outer loop: one iteration per tile segment of this CTA's unit range; the epilogue stays outside the
  // hot loop.  The ring and the fragment prefetch simply continue into the next segment.
  int kt_cur = kt0, tile_cur = tile0;
  int seg_kt0 = kt0, seg_un0 = 0;  // FIXUP: first k-tile and first unit of the segment being accumulated
  if (T::PREFETCH) {
    mbar_wait(bars + 0, 0);
    load_frags(0, base, 0);
  }
  int un = 0;
  while (un < nun) {
    int seg_end = nun;
    if (MULTI) {
      seg_end = un + (kt_total - kt_cur);
      if (seg_end > nun) seg_end = nun;
    }
    for (; un < seg_end; ++un) {
      const int s = un % kStages;
      const uint32_t st = base + s * L::kStageBytes;
      if (prod_lane && un + kStages - 1 < nun)
        produce(un + kStages - 1, false, !AF_ROTATE_PRODUCER || warp == (un + kStages - 1) % kConsumerWarps);
      __syncwarp();
      if (T::PREFETCH) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (ks < 3) {
            load_frags((ks + 1) & 1, st, ks + 1);
          } else if (un + 1 < nun) {
            const int s1 = (un + 1) % kStages;
            mbar_wait(bars + 8 * s1, ((un + 1) / kStages) & 1);
            load_frags(0, base + s1 * L::kStageBytes, 0);
          }
          mma_step(ks & 1);
        }
      } else {
        mbar_wait(bars + 8 * s, (un / kStages) & 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          load_frags(0, st, ks);
          mma_step(0);
        }
      }
      // every ld.shared of this stage has been consumed by an issued DMMA: hand the slot back
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (kStages + s));
    }
    if (!MULTI || un >= nun) break;  // the last (or only) segment is flushed after the loop
    if constexpr (FIXUP) {
      flush(tile_cur, seg_kt0 > 0 ? 1 : 0);  // (a segment flushed here ends at its tile's last k-tile)
      seg_kt0 = 0; seg_un0 = un;
    } else {
      flush(tile_cur);
    }
    zero_acc();
    kt_cur = 0;
    ++tile_cur;
  }
  if constexpr (FIXUP) flush(tile_cur, seg_kt0 > 0 ? 1 : (nun - seg_un0 == kt_total ? 0 : 2), nun - seg_un0);
  else flush(tile_cur);
}

// ---------------------------------------------------------------------------------------------
// generic kernel: any shape / alignment (odd leading dimensions, sliced views).  One output
// element per thread, same index algebra and epilogue; used when TMA's 16-byte rules fail.
// ---------------------------------------------------------------------------------------------
template <int ALAY, int BLAY, bool DUAL>
__global__ void __launch_bounds__(256) gemm_generic_kernel(const double* __restrict__ A0, const double* __restrict__ A1,
                                                           const double* __restrict__ B0, const double* __restrict__ B1,
                                                           long long lda, long long ldb, const GemmArgs p) {
  const int j = blockIdx.x * 16 + (threadIdx.x & 15);
  const long long i = (long long)blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i >= p.I || j >= p.J) return;
  double v0 = 0.0, v1 = 0.0;
  for (int k = 0; k < p.Kd; ++k) {
    const long long ia = (ALAY == LAY_KC) ? i * lda + k : (long long)k * lda + i;
    const long long ib = (BLAY == LAY_KC) ? (long long)j * ldb + k : (long long)k * ldb + j;
    const double a = A0[ia], b = B0[ib];
    v0 = fma(a, b, v0);
    if (DUAL) {
      if (B1) v1 = fma(a, B1[ib], v1);
      if (A1) v1 = fma(A1[ia], b, v1);
    }
  }
  epi_elem<DUAL>(p, i, j, v0, v1);
}

}  // namespace af
