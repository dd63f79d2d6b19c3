// The LSTM of bench/lstm.go (lstmBlock :129-151, lstmNet.StepTime :169-198, PropagateGradient
// :215-234), batched over B lanes and with the R-operator carried through, as one resident plan.
//
// Structure.  The four gate matrices share one input, concat(state, x)
// (bench/lstm.go:140,188), so they are stored stacked as one (4H) x (H+I) matrix W4 and evaluated by
// ONE contraction per time step; bias, sigmoid and their R forms are applied by the cell kernel; the cell
// algebra (Mul/Add, arithmetic.go:17-96,261-376) is one pointwise kernel per step.  Everything that
// does not depend on the recurrence is batched over all T*B rows: the output layer, its backward, and
// every weight / bias gradient (contraction over T*B samples instead of T launches).
//
//   forward  t = 0..T-1 :  Z_t = IN_t W4^T (+ R form)               [dual contraction, 4H x (H+I), n = B]
//                          G_t = sigmoid(Z_t + b4) ; s_t = f.s_{t-1} + m.i ; masked = s_t.o   [lstm_cell_fwd_kernel]
//            after loop  :  RES = sigmoid(AUG Wout^T + bout)         [af_dense_fwd, n = T*B]
//   backward before loop:  dOut = sigmoid'(RES) up ; gWout, gbout ; dAUG = dOut Wout
//            t = T-1..0  :  delta4_t, sg' = sg.f                      [lstm_cell_bwd_kernel]
//                          sg' += delta4_t W4[:, :H]                 [dual contraction, += , split-K]
//            after loop  :  gW4 += delta4^T IN ; gb4 += colsum(delta4)   (n = T*B)
#include <vector>

#include "common.h"
#include "gemm_dmma.cuh"

namespace af {
bool lstm_fused_applicable(af_ctx* ctx, long long H, long long C, long long B);
int lstm_fwd_step_launch(af_ctx* ctx, const double* W4, const double* RW4, const LstmStepArgs& st, bool with_r, bool r_state,
                         long long nb);
int lstm_bwd_step_launch(af_ctx* ctx, const double* D4, const double* RD4, const double* W4, const double* RW4,
                         const LstmStepArgs& st, bool with_r, long long nb);
int colsum_launch(af_ctx* ctx, const double* U, const double* UR, double* gb, double* rgb, long long len, long long n);

namespace {

// dst[r][c] = src[r][c] for a rows x cols block with independent row pitches
__global__ void __launch_bounds__(256) copy2d_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src,
                                                     long long lds, long long rows, long long cols) {
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const long long r = i / cols, c = i - r * cols;
    dst[r * ldd + c] = src[r * lds + c];
  }
}

// Cell forward for one time step, B x H elements.  Gates G = [i | m | f | o] (B x 4H).
//   s = f*sp + m*i            Add(Mul(forget, state), Mul(inMask, inState))   bench/lstm.go:147-150
//   masked = s*o              Mul(outStateVar, outGate)                       bench/lstm.go:191
// R forms follow MulR (arithmetic.go:325-329: x*yR + xR*y) and AddR (:58-65).
// On entry G / RG hold the four gates' raw contraction sums (IN W4^T and its R form); the kernel adds
// the biases (lin_add.go:15-17,37-42), applies the sigmoid and its R form (math_funcs.go:305-309) and
// stores the activated gates back for the backward pass -- the gate contraction needs no epilogue pass.
__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(double* __restrict__ G, double* __restrict__ RG,
                                                            const double* __restrict__ b4, const double* __restrict__ Rb4,
                                                            const double* __restrict__ INp, const double* __restrict__ RINp,
                                                            double* __restrict__ INn, double* __restrict__ RINn,
                                                            double* __restrict__ AUG, double* __restrict__ RAUG, int B,
                                                            int H, int ld) {
  pdl_wait();
  const int n = B * H;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < n; e += gridDim.x * 256) {
    const int b = e / H, j = e - b * H;
    double* g = G + (long long)b * 4 * H;
    const long long si = (long long)b * ld + j;
    double rz[4] = {0.0, 0.0, 0.0, 0.0};
    double* rg = RG ? RG + (long long)b * 4 * H : nullptr;
    if (RG) {
#pragma unroll
      for (int u = 0; u < 4; ++u) rz[u] = rg[u * H + j] + Rb4[u * H + j];
    }
    const double z0 = g[j] + b4[j], z1 = g[H + j] + b4[H + j], z2 = g[2 * H + j] + b4[2 * H + j], z3 = g[3 * H + j] + b4[3 * H + j];
    if (RG) {
      const LstmCellFwd c = lstm_cell_fwd<true>(z0, z1, z2, z3, INp[si], rz[0], rz[1], rz[2], rz[3], RINp[si]);
      g[j] = c.i; g[H + j] = c.m; g[2 * H + j] = c.f; g[3 * H + j] = c.o;
      INn[si] = c.s; AUG[si] = c.masked;
      rg[j] = c.ri; rg[H + j] = c.rm; rg[2 * H + j] = c.rf; rg[3 * H + j] = c.ro;
      RINn[si] = c.rs; RAUG[si] = c.rmasked;
    } else {
      const LstmCellFwd c = lstm_cell_fwd<false>(z0, z1, z2, z3, INp[si], 0.0, 0.0, 0.0, 0.0, 0.0);
      g[j] = c.i; g[H + j] = c.m; g[2 * H + j] = c.f; g[3 * H + j] = c.o;
      INn[si] = c.s; AUG[si] = c.masked;
    }
  }
}

// Cell backward for one time step (bench/lstm.go:215-234 with the node backwards of
// arithmetic.go:34-49,288-306 / 79-96,349-376 and math_funcs.go:326-333 / 357-374 inlined).
//   sg  = sg_carry + dmasked*o          (grad[outStateVar] added to stateGrad, lstm.go:227-228)
//   do  = dmasked*s                     -> delta_o = sigmoid'(o) do
//   dm  = sg*i ; di = sg*m ; df = sg*sp -> delta_* = sigmoid'(gate) d*
//   sg' = sg*f                          (Mul(forget,state) toward the state)
// D4 = [delta_i | delta_m | delta_f | delta_o]; SGn receives sg' and is then accumulated into by the
// input-gradient contraction (delta4 W4[:, :H]).
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(
    const double* __restrict__ G, const double* __restrict__ RG, const double* __restrict__ INp,
    const double* __restrict__ RINp, const double* __restrict__ INn, const double* __restrict__ RINn,
    const double* __restrict__ DAUG, const double* __restrict__ RDAUG, const double* __restrict__ SG,
    const double* __restrict__ RSG, double* __restrict__ SGn, double* __restrict__ RSGn, double* __restrict__ D4,
    double* __restrict__ RD4, int B, int H, int ld) {
  pdl_wait();
  const int n = B * H;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < n; e += gridDim.x * 256) {
    const int b = e / H, j = e - b * H;
    const long long gi = (long long)b * 4 * H, si = (long long)b * ld + j;
    const double i = G[gi + j], m = G[gi + H + j], f = G[gi + 2 * H + j], o = G[gi + 3 * H + j];
    if (RG) {
      const LstmCellBwd c = lstm_cell_bwd<true>(i, m, f, o, INp[si], INn[si], DAUG[si], SG[e], RG[gi + j], RG[gi + H + j],
                                                RG[gi + 2 * H + j], RG[gi + 3 * H + j], RINp[si], RINn[si], RDAUG[si], RSG[e]);
      D4[gi + j] = c.di; D4[gi + H + j] = c.dm; D4[gi + 2 * H + j] = c.df; D4[gi + 3 * H + j] = c.d_o;
      SGn[e] = c.sgf;
      RD4[gi + j] = c.rdi; RD4[gi + H + j] = c.rdm; RD4[gi + 2 * H + j] = c.rdf; RD4[gi + 3 * H + j] = c.rd_o;
      RSGn[e] = c.rsgf;
    } else {
      const LstmCellBwd c = lstm_cell_bwd<false>(i, m, f, o, INp[si], INn[si], DAUG[si], SG[e], 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0);
      D4[gi + j] = c.di; D4[gi + H + j] = c.dm; D4[gi + 2 * H + j] = c.df; D4[gi + 3 * H + j] = c.d_o;
      SGn[e] = c.sgf;
    }
  }
}

int ew_grid(af_ctx* ctx, long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace
}  // namespace af

using namespace af;

struct af_lstm {
  af_ctx* ctx = nullptr;
  int64_t I = 0, H = 0, O = 0, T = 0, B = 0, C = 0;  // C = H + I
  double* slab = nullptr;
  // parameters: W4 (4H x C), b4 (4H), Wout (O x C), bout (O) and their R-vectors / accumulators
  double *W4, *b4, *Wout, *bout, *RW4, *Rb4, *RWout, *Rbout;
  double *gW4, *gb4, *gWout, *gbout, *rgW4, *rgb4, *rgWout, *rgbout;
  bool has_r[10] = {false};
  double *IN, *RIN;      // (T+1) x B x C : [:, :H] = state entering step t, [:, H:] = x_t
  double *G, *RG;        // T x B x 4H gates
  double *AUG, *RAUG;    // T x B x C : [masked | x_t]
  double *RES, *RRES;    // T x B x O
  double *UP, *RUP;      // T x B x O upstream (consumed)
  double *DAUG, *RDAUG;  // T x B x C
  double *D4, *RD4;      // T x B x 4H
  double *SG[2], *RSG[2];
  int64_t grad_doubles = 0;
  // The lanes are independent through the whole recurrence, so the time loops run the two halves of the batch as two
  // concurrent kernel chains (main stream + side stream, forked and joined with events -- also inside the captured
  // graphs): while one half is in a kernel's tail or launch gap the other half's contraction has the SMs.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // CUDA-graph replay of the two time loops (hundreds of short dependent launches): [with_r][fwd/bwd]
  struct LoopGraph { cudaGraphExec_t exec = nullptr; int seen = 0; int64_t launches = 0; int64_t epoch = 0; };
  LoopGraph graphs[2][2];
  int64_t param_len(int idx) const {
    const int k = idx / 2;
    const int64_t rows = k < 4 ? H : O;
    return (idx & 1) ? rows : rows * C;
  }
  // reference order (bench/lstm.go:75-99): In, InGate, Forget, OutputGate, Output; weights then biases
  double* block(int idx, double* w4, double* bb4, double* wout, double* bbout) const {
    const int k = idx / 2;
    if (k < 4) return (idx & 1) ? bb4 + k * H : w4 + k * H * C;
    return (idx & 1) ? bbout : wout;
  }
};

extern "C" {

int af_copy_2d(af_ctx* ctx, double* d_dst, int64_t ld_dst, const double* d_src, int64_t ld_src, int64_t rows,
               int64_t cols) {
  if (!ctx) return fail(AF_EINVAL, "null context");
  DeviceGuard guard(ctx->device);
  if (rows <= 0 || cols <= 0) return AF_OK;
  if (!d_dst || !d_src) return fail(AF_EINVAL, "null pointer argument");
  if (ld_dst < cols || ld_src < cols) return fail(AF_ESHAPE, "row pitch smaller than the row length");
  copy2d_kernel<<<ew_grid(ctx, rows * cols), 256, 0, ctx->stream>>>(d_dst, ld_dst, d_src, ld_src, rows, cols);
  return after_launch(ctx, "copy2d_kernel", KIND_ELEMENTWISE, 16.0 * rows * cols);
}

int af_lstm_create(af_ctx* ctx, int64_t I, int64_t H, int64_t O, int64_t T, int64_t B, af_lstm** out) {
  if (!ctx || !out) return fail(AF_EINVAL, "null argument");
  if (I < 1 || H < 1 || O < 1 || T < 1 || B < 1) return fail(AF_ESHAPE, "LSTM dimensions must be positive");
  if (ctx->recording) return fail(AF_EINVAL, "af_lstm_create: not allowed between af_graph_begin and af_graph_end");
  DeviceGuard guard(ctx->device);
  af_lstm* m = new af_lstm();
  m->ctx = ctx; m->I = I; m->H = H; m->O = O; m->T = T; m->B = B; m->C = H + I;
  const int64_t C = m->C;
  auto pad = [](int64_t x) { return (x + 31) / 32 * 32; };
  int64_t total = 0;
  std::vector<int64_t> off;
  auto take = [&](int64_t len) { off.push_back(total); total += pad(len); };
  const int64_t plens[4] = {4 * H * C, 4 * H, O * C, O};
  for (int rep = 0; rep < 4; ++rep)  // param, rvec, grad, rgrad
    for (int k = 0; k < 4; ++k) take(plens[k]);
  const int64_t TB = T * B;
  take((T + 1) * B * C); take((T + 1) * B * C);  // IN, RIN
  take(TB * 4 * H); take(TB * 4 * H);            // G, RG
  take(TB * C); take(TB * C);                    // AUG, RAUG
  take(TB * O); take(TB * O);                    // RES, RRES
  take(TB * O); take(TB * O);                    // UP, RUP
  take(TB * C); take(TB * C);                    // DAUG, RDAUG
  take(TB * 4 * H); take(TB * 4 * H);            // D4, RD4
  for (int k = 0; k < 4; ++k) take(B * H);       // SG[2], RSG[2]
  cudaError_t e = cudaMalloc(&m->slab, (size_t)total * 8);
  if (e != cudaSuccess) { delete m; return cuda_fail(e, "cudaMalloc(lstm slab)"); }
  e = cudaMemsetAsync(m->slab, 0, (size_t)total * 8, ctx->stream);
  if (e != cudaSuccess) { cudaFree(m->slab); delete m; return cuda_fail(e, "cudaMemsetAsync(lstm slab)"); }
  size_t k = 0;
  auto next = [&]() { return m->slab + off[k++]; };
  m->W4 = next(); m->b4 = next(); m->Wout = next(); m->bout = next();
  m->RW4 = next(); m->Rb4 = next(); m->RWout = next(); m->Rbout = next();
  m->gW4 = next(); m->gb4 = next(); m->gWout = next(); m->gbout = next();
  m->rgW4 = next(); m->rgb4 = next(); m->rgWout = next(); m->rgbout = next();
  m->IN = next(); m->RIN = next(); m->G = next(); m->RG = next(); m->AUG = next(); m->RAUG = next();
  m->RES = next(); m->RRES = next(); m->UP = next(); m->RUP = next(); m->DAUG = next(); m->RDAUG = next();
  m->D4 = next(); m->RD4 = next();
  m->SG[0] = next(); m->SG[1] = next(); m->RSG[0] = next(); m->RSG[1] = next();
  m->grad_doubles = m->IN - m->gW4;  // grad | rgrad blocks are adjacent
  *out = m;
  return AF_OK;
}

int af_lstm_destroy(af_lstm* m) {
  if (!m) return AF_OK;
  DeviceGuard guard(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  for (auto& a : m->graphs)
    for (auto& g : a)
      if (g.exec) cudaGraphExecDestroy(g.exec);
  if (m->side) cudaStreamDestroy(m->side);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  cudaFree(m->slab);
  delete m;
  return AF_OK;
}

static int lstm_idx_ok(af_lstm* m, int idx, int64_t len) {
  if (!m) return fail(AF_EINVAL, "null lstm");
  if (idx < 0 || idx >= 10) return fail(AF_EINVAL, "parameter index out of range (0..9)");
  if (len != m->param_len(idx))
    return fail(AF_ESHAPE, "parameter length should be " + std::to_string(m->param_len(idx)) + " but got " + std::to_string(len));
  return AF_OK;
}

int af_lstm_set_param(af_lstm* m, int idx, const double* h, int64_t len) {
  AF_TRY(lstm_idx_ok(m, idx, len));
  if (!h) return fail(AF_EINVAL, "null host pointer");
  return af_upload(m->ctx, m->block(idx, m->W4, m->b4, m->Wout, m->bout), h, len);
}

// NULL h_value marks the variable as absent from the RVector (zero R, variable.go:85-91).  The four
// gate matrices share one stacked R buffer, so an absent gate block is zero-filled rather than skipped.
int af_lstm_set_rvector(af_lstm* m, int idx, const double* h, int64_t len) {
  if (!m) return fail(AF_EINVAL, "null lstm");
  if (idx < 0 || idx >= 10) return fail(AF_EINVAL, "parameter index out of range (0..9)");
  double* dst = m->block(idx, m->RW4, m->Rb4, m->RWout, m->Rbout);
  if (!h) {
    m->has_r[idx] = false;
    return af_zero(m->ctx, dst, m->param_len(idx));
  }
  AF_TRY(lstm_idx_ok(m, idx, len));
  m->has_r[idx] = true;
  return af_upload(m->ctx, dst, h, len);
}

double* af_lstm_param_ptr(af_lstm* m, int idx) { return (m && idx >= 0 && idx < 10) ? m->block(idx, m->W4, m->b4, m->Wout, m->bout) : nullptr; }
double* af_lstm_grad_ptr(af_lstm* m, int idx) { return (m && idx >= 0 && idx < 10) ? m->block(idx, m->gW4, m->gb4, m->gWout, m->gbout) : nullptr; }
double* af_lstm_rgrad_ptr(af_lstm* m, int idx) { return (m && idx >= 0 && idx < 10) ? m->block(idx, m->rgW4, m->rgb4, m->rgWout, m->rgbout) : nullptr; }
double* af_lstm_output_ptr(af_lstm* m) { return m ? m->RES : nullptr; }
double* af_lstm_routput_ptr(af_lstm* m) { return m ? m->RRES : nullptr; }
double* af_lstm_upstream_ptr(af_lstm* m) { return m ? m->UP : nullptr; }
double* af_lstm_upstream_r_ptr(af_lstm* m) { return m ? m->RUP : nullptr; }
int64_t af_lstm_grad_slab(af_lstm* m, double** d_first) {
  if (!m) return 0;
  if (d_first) *d_first = m->gW4;
  return m->grad_doubles;
}

int af_lstm_zero_grads(af_lstm* m) {
  if (!m) return fail(AF_EINVAL, "null lstm");
  DeviceGuard guard(m->ctx->device);
  AF_CUDA(cudaMemsetAsync(m->gW4, 0, (size_t)m->grad_doubles * 8, m->ctx->stream));
  return AF_OK;
}

// Gradient.AddToVars (gradient.go:40-52) for the plan's own parameters, in HBM
int af_lstm_add_to_vars(af_lstm* m, double scale) {
  if (!m) return fail(AF_EINVAL, "null lstm");
  AF_TRY(af_axpy(m->ctx, scale, m->gW4, m->W4, 4 * m->H * m->C));
  AF_TRY(af_axpy(m->ctx, scale, m->gb4, m->b4, 4 * m->H));
  AF_TRY(af_axpy(m->ctx, scale, m->gWout, m->Wout, m->O * m->C));
  return af_axpy(m->ctx, scale, m->gbout, m->bout, m->O);
}

// inputs: T x B x I (time-major, lane-major within a step: the packing seqfunc.MapBatcher produces,
// seqfunc/map.go:262-270).  Device pointer.
int af_lstm_set_inputs(af_lstm* m, const double* d_x) {
  if (!m || !d_x) return fail(AF_EINVAL, "null argument");
  af_ctx* ctx = m->ctx;
  const int64_t TB = m->T * m->B;
  // x_t is the tail of both concatenations: concat(state, x) and concat(masked, x) (lstm.go:140,193)
  AF_TRY(af_copy_2d(ctx, m->IN + m->H, m->C, d_x, m->I, TB, m->I));
  return af_copy_2d(ctx, m->AUG + m->H, m->C, d_x, m->I, TB, m->I);
}

}  // extern "C"

// Runs `body` eagerly the first time, captures it into a CUDA graph the second time and replays it afterwards.
template <typename F>
static int lstm_graphed(af_lstm* m, af_lstm::LoopGraph& g, F body) {
  af_ctx* ctx = m->ctx;
  if (ctx->tracing || ctx->stream == nullptr || ctx->stream == cudaStreamLegacy || g.seen < 0) return body();
  if (g.exec && g.epoch != ctx->option_epoch) {  // an option changed since the capture: the decomposition may differ
    cudaGraphExecDestroy(g.exec);
    g = af_lstm::LoopGraph();
    g.seen = 1;
  }
  if (g.exec) {
    AF_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
    ctx->launches += g.launches;
    return AF_OK;
  }
  if (g.seen++ == 0) return body();
  const int64_t before = ctx->launches;
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    g.seen = -1;  // this stream cannot be captured: stay eager
    return body();
  }
  const int rc = body();
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
  if (rc != AF_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
  g.launches = ctx->launches - before;
  g.epoch = ctx->option_epoch;
  ctx->launches = before;
  const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) { g.exec = nullptr; g.seen = -1; cudaGetLastError(); return body(); }
  AF_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
  ctx->launches += g.launches;
  return AF_OK;
}

static int lstm_forward_launches(af_lstm* m, int with_r);
static int lstm_backward_launches(af_lstm* m, int with_r);

extern "C" {

int af_lstm_forward(af_lstm* m, int with_r) {
  if (!m) return fail(AF_EINVAL, "null lstm");
  if (m->ctx->recording) return fail(AF_EINVAL, "af_lstm_forward: plans replay graphs of their own and cannot be recorded");
  DeviceGuard guard(m->ctx->device);
  return lstm_graphed(m, m->graphs[with_r ? 1 : 0][0], [&]() { return lstm_forward_launches(m, with_r); });
}

int af_lstm_backward(af_lstm* m, int with_r) {
  if (!m) return fail(AF_EINVAL, "null lstm");
  if (m->ctx->recording) return fail(AF_EINVAL, "af_lstm_backward: plans replay graphs of their own and cannot be recorded");
  DeviceGuard guard(m->ctx->device);
  return lstm_graphed(m, m->graphs[with_r ? 1 : 0][1], [&]() { return lstm_backward_launches(m, with_r); });
}

}  // extern "C"

// number of independent lane groups the time loops run concurrently (1 = single chain)
static int lstm_lane_groups(af_lstm* m) {
  af_ctx* ctx = m->ctx;
  if (ctx->lstm_split_off || ctx->tracing || m->B < 128 || (m->B & 1)) return 1;
  if (ctx->stream == nullptr || ctx->stream == cudaStreamLegacy) return 1;
  if (!m->side) {
    if (cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      return 1;
    }
  }
  return 2;
}
// run body(group, first lane, lanes) for every lane group, group 1 on the side stream
template <typename F>
static int lstm_for_groups(af_lstm* m, int groups, F body) {
  af_ctx* ctx = m->ctx;
  if (groups == 1) return body(0, (int64_t)0, m->B);
  cudaStream_t main_stream = ctx->stream;
  AF_CUDA(cudaEventRecord(m->ev_fork, main_stream));
  AF_CUDA(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
  const int64_t half = m->B / 2;
  int rc = AF_OK;
  for (int g = 0; g < 2 && rc == AF_OK; ++g) {
    ctx->stream = g == 0 ? main_stream : m->side;
    rc = body(g, g * half, half);
  }
  ctx->stream = main_stream;
  AF_CUDA(cudaEventRecord(m->ev_join, m->side));
  AF_CUDA(cudaStreamWaitEvent(main_stream, m->ev_join, 0));
  return rc;
}

static int lstm_forward_launches(af_lstm* m, int with_r) {
  af_ctx* ctx = m->ctx;
  const int64_t H = m->H, C = m->C, B = m->B, T = m->T, O = m->O;
  const bool R = with_r != 0;
  const int64_t BC = B * C;
  const bool fused = lstm_fused_applicable(ctx, H, C, B);
  if (fused) {
    // The x part of concat(state, x) (lstm.go:140) does not depend on the recurrence: project it for all T x B rows at
    // once, bias included, straight into the gate buffers -- G = X W4[:, H:]^T + b4 (and the R form; x has no R-vector) --
    // so that the recurrent contraction is K = H (aligned) and the step's epilogue finds its addend where it will store
    // the activated gates.
    if (few_cols_fwd_applicable(ctx, 4 * H, m->I, T * B)) {  // I <= 32 (the reference's I = 30): a streaming pass
      AF_TRY(few_cols_fwd_launch(ctx, m->IN + H, C, m->W4 + H, R ? m->RW4 + H : nullptr, C, m->b4, R ? m->Rb4 : nullptr, m->G,
                                 R ? m->RG : nullptr, 4 * H, 4 * H, m->I, T * B));
    } else {
      GemmOperand A{m->IN + H, nullptr, LAY_KC, C};
      GemmOperand Bop{m->W4 + H, R ? m->RW4 + H : nullptr, LAY_KC, C};
      GemmEpilogue e;
      e.epi = EPI_BIAS; e.out0 = m->G; e.out1 = R ? m->RG : nullptr; e.ldo = 4 * H; e.bias0 = m->b4; e.bias1 = R ? m->Rb4 : nullptr;
      AF_TRY(gemm_launch(ctx, A, Bop, R, T * B, 4 * H, m->I, e));
    }
  } else {
    // the gate sums of every step are accumulated into zeroed buffers when the contraction is split: clear all
    // T steps at once instead of two memsets per step
    AF_CUDA(cudaMemsetAsync(m->G, 0, (size_t)(T * B * 4 * H) * 8, ctx->stream));
    if (R) AF_CUDA(cudaMemsetAsync(m->RG, 0, (size_t)(T * B * 4 * H) * 8, ctx->stream));
  }
  // initial state is zero (lstm.go:175) with zero R
  AF_CUDA(cudaMemset2DAsync(m->IN, C * 8, 0, H * 8, B, ctx->stream));
  if (R) AF_CUDA(cudaMemset2DAsync(m->RIN, C * 8, 0, H * 8, B, ctx->stream));
  // the time loop is a chain of dependent kernels (contraction -> cell -> contraction ...), one chain per lane group
  struct PdlScope { af_ctx* c; explicit PdlScope(af_ctx* c_) : c(c_) { c->pdl = true; } ~PdlScope() { c->pdl = false; } };
  AF_TRY(lstm_for_groups(m, lstm_lane_groups(m), [&](int, int64_t b0, int64_t nb) -> int {
    for (int64_t t = 0; t < T; ++t) {
      PdlScope pdl_scope(ctx);
      double* INt = m->IN + t * BC + b0 * C;
      double* RINt = m->RIN + t * BC + b0 * C;
      double* Gt = m->G + (t * B + b0) * 4 * H;
      double* RGt = m->RG + (t * B + b0) * 4 * H;
      if (fused) {
        // ONE kernel per step: gate contraction over the state + bias + sigmoids + cell algebra (lstm_step.cu)
        LstmStepArgs st{};
        st.H = (int)H; st.ld = (int)C;
        st.G = Gt; st.RG = R ? RGt : nullptr;
        st.INp = INt; st.RINp = RINt; st.INn = INt + BC; st.RINn = RINt + BC;
        st.AUG = m->AUG + t * BC + b0 * C; st.RAUG = m->RAUG + t * BC + b0 * C;
        AF_TRY(lstm_fwd_step_launch(ctx, m->W4, R ? m->RW4 : nullptr, st, R, t > 0, nb));
        continue;
      }
      // all four gates in one contraction (R-state of step 0 is identically zero: skip its product); bias and
      // sigmoid are applied by the cell kernel, so a split / stream-K launch needs no separate epilogue pass
      {
        GemmOperand A{INt, (R && t > 0) ? RINt : nullptr, LAY_KC, C};
        GemmOperand Bop{m->W4, R ? m->RW4 : nullptr, LAY_KC, C};
        GemmEpilogue e;
        e.epi = EPI_STORE; e.out0 = Gt; e.out1 = R ? RGt : nullptr; e.ldo = 4 * H; e.prezeroed = true;
        AF_TRY(gemm_launch(ctx, A, Bop, R, nb, 4 * H, C, e));
      }
      AF_CUDA(launch_kernel(lstm_cell_fwd_kernel, dim3(ew_grid(ctx, nb * H)), dim3(256), 0, ctx->stream,
                            ctx->pdl && !ctx->pdl_off, Gt, R ? RGt : nullptr, m->b4, m->Rb4, INt, RINt, INt + BC, RINt + BC,
                            m->AUG + t * BC + b0 * C, m->RAUG + t * BC + b0 * C, (int)nb, (int)H, (int)C));
      AF_TRY(after_launch(ctx, "lstm_cell_fwd_kernel", KIND_ELEMENTWISE, (R ? 2.0 : 1.0) * 7 * 8 * nb * H));
    }
    return AF_OK;
  }));
  // output layer for every step at once (lstm.go:193-194)
  return af_dense_fwd(ctx, m->Wout, R ? m->RWout : nullptr, m->bout, R ? m->Rbout : nullptr, m->AUG, R ? m->RAUG : nullptr,
                      m->RES, R ? m->RRES : nullptr, O, C, T * B, 1);
}

static int lstm_backward_launches(af_lstm* m, int with_r) {
  af_ctx* ctx = m->ctx;
  const int64_t H = m->H, C = m->C, B = m->B, T = m->T, O = m->O;
  const bool R = with_r != 0;
  const int64_t TB = T * B, BC = B * C;
  // output layer backward for all steps: sigmoid', bias and weight gradients, gradient toward [masked | x]
  if (R) AF_TRY(af_sigmoid_bwd_r(ctx, m->RES, m->RRES, m->UP, m->RUP, TB * O));
  else AF_TRY(af_sigmoid_bwd(ctx, m->RES, m->UP, TB * O));
  AF_TRY(colsum_launch(ctx, m->UP, R ? m->RUP : nullptr, m->gbout, R ? m->rgbout : nullptr, O, TB));
  AF_TRY(af_lintran_wgrad_r(ctx, m->UP, R ? m->RUP : nullptr, m->AUG, R ? m->RAUG : nullptr, m->gWout, R ? m->rgWout : nullptr, O, C, TB));
  AF_TRY(af_dense_igrad(ctx, m->UP, R ? m->RUP : nullptr, m->Wout, R ? m->RWout : nullptr, nullptr, nullptr, m->DAUG,
                        R ? m->RDAUG : nullptr, O, C, TB));
  AF_CUDA(cudaMemsetAsync(m->SG[0], 0, (size_t)(B * H) * 8, ctx->stream));
  if (R) AF_CUDA(cudaMemsetAsync(m->RSG[0], 0, (size_t)(B * H) * 8, ctx->stream));
  struct PdlScope { af_ctx* c; explicit PdlScope(af_ctx* c_) : c(c_) { c->pdl = true; } ~PdlScope() { c->pdl = false; } };
  const bool fused = lstm_fused_applicable(ctx, H, C, B);
  AF_TRY(lstm_for_groups(m, lstm_lane_groups(m), [&](int, int64_t b0, int64_t nb) -> int {
    int cur = 0;
    for (int64_t t = T - 1; t >= 0; --t) {
      PdlScope pdl_scope(ctx);
      double* D4t = m->D4 + (t * B + b0) * 4 * H;
      double* RD4t = m->RD4 + (t * B + b0) * 4 * H;
      const int64_t go = (t * B + b0) * 4 * H, io = t * BC + b0 * C, so = b0 * H;
      if (!fused || t == T - 1) {
        // (fused path: only the last time step, whose incoming state gradient is zero, runs the stand-alone cell kernel;
        // it leaves sg f in SG[1], which the fused steps then update in place)
        AF_CUDA(launch_kernel(lstm_cell_bwd_kernel, dim3(ew_grid(ctx, nb * H)), dim3(256), 0, ctx->stream,
                              ctx->pdl && !ctx->pdl_off, m->G + go, R ? m->RG + go : nullptr, m->IN + io, m->RIN + io,
                              m->IN + io + BC, m->RIN + io + BC, m->DAUG + io, m->RDAUG + io, m->SG[cur] + so, m->RSG[cur] + so,
                              m->SG[cur ^ 1] + so, m->RSG[cur ^ 1] + so, D4t, RD4t, (int)nb, (int)H, (int)C));
        AF_TRY(after_launch(ctx, "lstm_cell_bwd_kernel", KIND_ELEMENTWISE, (R ? 2.0 : 1.0) * 13 * 8 * nb * H));
      }
      if (fused) {
        if (t > 0) {
          // ONE kernel: stateGrad += delta4_t W4[:, :H] over a cluster of 4 CTAs (split-K through distributed shared
          // memory), then the cell backward of step t - 1 in the epilogue (lstm_step.cu)
          LstmStepArgs st{};
          st.H = (int)H; st.ld = (int)C;
          st.Gm = m->G + go - B * 4 * H; st.RGm = R ? m->RG + go - B * 4 * H : nullptr;
          st.INm = m->IN + io - BC; st.RINm = m->RIN + io - BC; st.INt = m->IN + io; st.RINt = m->RIN + io;
          st.DAUGm = m->DAUG + io - BC; st.RDAUGm = m->RDAUG + io - BC;
          st.SGF = m->SG[1] + so; st.RSGF = m->RSG[1] + so;
          st.D4m = D4t - B * 4 * H; st.RD4m = RD4t - B * 4 * H;
          AF_TRY(lstm_bwd_step_launch(ctx, D4t, R ? RD4t : nullptr, m->W4, R ? m->RW4 : nullptr, st, R, nb));
        }
        continue;
      }
      if (t > 0) {
        // stateGrad += delta4 W4[:, :H]  (the four gates' inputGradient, lin_tran.go:272-359, restricted to the
        // state columns: the x columns go to a Variable that is not in the gradient maps)
        GemmOperand A{D4t, R ? RD4t : nullptr, LAY_KC, 4 * H};
        GemmOperand Bop{m->W4, R ? m->RW4 : nullptr, LAY_RC, C};
        GemmEpilogue e;
        e.epi = EPI_ACCUM; e.out0 = m->SG[cur ^ 1] + so; e.out1 = R ? m->RSG[cur ^ 1] + so : nullptr; e.ldo = H;
        AF_TRY(gemm_launch(ctx, A, Bop, R, nb, H, 4 * H, e));
      }
      cur ^= 1;
    }
    return AF_OK;
  }));
  // gate weight / bias gradients over all T*B rows at once (lstm.go's per-step dataGradient calls add up)
  AF_TRY(colsum_launch(ctx, m->D4, R ? m->RD4 : nullptr, m->gb4, R ? m->rgb4 : nullptr, 4 * H, TB));
  return af_lintran_wgrad_r(ctx, m->D4, R ? m->RD4 : nullptr, m->IN, R ? m->RIN : nullptr, m->gW4, R ? m->rgW4 : nullptr, 4 * H, C, TB);
}

extern "C" {

// Host-buffer form of one bench iteration (bench/lstm.go:41-48): upload inputs (T x B x I) and the
// per-step output gradients (T x B x O; h_upstreams_r may be NULL = zeros), zero the accumulators,
// forward, BPTT, download outputs and every gradient (10 host pointers each, reference order).
// One data-parallel iteration on resident inputs / upstreams: fresh accumulators, forward, BPTT, then the sum of the
// ranks' {grad | rgrad} blocks (Gradient.Add across ranks, gradient.go:28-30) in order on the ctx stream.
int af_lstm_step_dp(af_lstm* m, int with_r) {
  if (!m) return fail(AF_EINVAL, "null lstm");
  AF_TRY(af_lstm_zero_grads(m));
  AF_TRY(af_lstm_forward(m, with_r));
  AF_TRY(af_lstm_backward(m, with_r));
  DeviceGuard guard(m->ctx->device);
  return dp_allreduce(m->ctx, m->gW4, m->grad_doubles, m->ctx->stream, true);
}

int af_lstm_run_host(af_lstm* m, int with_r, const double* h_x, const double* h_up, const double* h_upr, double* h_out,
                     double* h_rout, double* const* h_grads, double* const* h_rgrads) {
  if (!m || !h_x || !h_up) return fail(AF_EINVAL, "null argument");
  af_ctx* ctx = m->ctx;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = ctx->stream;
  const int64_t TB = m->T * m->B;
  // x_t is the tail of both concatenations, concat(state, x) and concat(masked, x) (lstm.go:140,193): a strided copy
  // puts it straight into IN's x columns (no staging buffer -- any input width fits), AUG's come from there on the device
  AF_CUDA(cudaMemcpy2DAsync(m->IN + m->H, (size_t)m->C * 8, h_x, (size_t)m->I * 8, (size_t)m->I * 8, (size_t)TB,
                            cudaMemcpyHostToDevice, st));
  AF_TRY(af_copy_2d(ctx, m->AUG + m->H, m->C, m->IN + m->H, m->C, TB, m->I));
  AF_CUDA(cudaMemcpyAsync(m->UP, h_up, (size_t)(TB * m->O) * 8, cudaMemcpyHostToDevice, st));
  if (with_r) {
    if (h_upr) AF_CUDA(cudaMemcpyAsync(m->RUP, h_upr, (size_t)(TB * m->O) * 8, cudaMemcpyHostToDevice, st));
    else AF_CUDA(cudaMemsetAsync(m->RUP, 0, (size_t)(TB * m->O) * 8, st));
  }
  AF_TRY(af_lstm_zero_grads(m));
  AF_TRY(af_lstm_forward(m, with_r));
  if (h_out) AF_CUDA(cudaMemcpyAsync(h_out, m->RES, (size_t)(TB * m->O) * 8, cudaMemcpyDeviceToHost, st));
  if (h_rout && with_r) AF_CUDA(cudaMemcpyAsync(h_rout, m->RRES, (size_t)(TB * m->O) * 8, cudaMemcpyDeviceToHost, st));
  AF_TRY(af_lstm_backward(m, with_r));
  for (int i = 0; i < 10; ++i) {
    if (h_grads && h_grads[i])
      AF_CUDA(cudaMemcpyAsync(h_grads[i], af_lstm_grad_ptr(m, i), (size_t)m->param_len(i) * 8, cudaMemcpyDeviceToHost, st));
    if (with_r && h_rgrads && h_rgrads[i])
      AF_CUDA(cudaMemcpyAsync(h_rgrads[i], af_lstm_rgrad_ptr(m, i), (size_t)m->param_len(i) * 8, cudaMemcpyDeviceToHost, st));
  }
  AF_CUDA(cudaStreamSynchronize(st));
  return AF_OK;
}

}  // extern "C"
