/*
 * autofunc_b200 -- C ABI of the B200-native dense differentiation path of unixpickle/autofunc.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: the path sits
 * behind the Go interfaces Func/RFunc (func.go:7-27), Batcher/RBatcher (batcher.go:5-29) and
 * Result/RResult (result.go:10-77).  A Go (cgo) shim -- the package go/b200, shown in INTEGRATION.md --
 * implements those interfaces by calling the entry points below; each entry point cites the
 * reference routine it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a negative AF_E* code; af_last_error() gives text.
 *   - all numeric data is IEEE float64; matrices are row-major; a packed batch of n samples is
 *     sample-major (batcher.go:6-17): X is n x cols, Y is n x rows.
 *   - `const double*` / `double*` arguments named d_* are DEVICE pointers (cudaMalloc'ed by
 *     af_malloc, or any other allocator sharing the device's primary context, e.g. torch);
 *     arguments named h_* are HOST pointers.
 *   - a NULL R-operand means "identically zero" (variable.go:85-91: a variable absent from the
 *     RVector has a zero R-vector) and the corresponding product is skipped.
 *   - gradient outputs ACCUMULATE (+=), as in the reference (variable.go:45; beta=1 at
 *     lin_tran.go:210,267-268).
 *   - one af_ctx = one device + one stream; calls on a ctx are serialised by the caller
 *     (the reference is single-goroutine).  Work is asynchronous on the ctx stream until
 *     af_sync() or a host-pointer call.
 */
#ifndef AUTOFUNC_B200_H_
#define AUTOFUNC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AF_OK 0
#define AF_ESHAPE (-1)   /* the reference panics: lin_tran.go:26,35,52,66; arithmetic.go:265 */
#define AF_ECUDA (-2)
#define AF_ENOMEM (-3)
#define AF_EINVAL (-4)
#define AF_ENOTSUP (-5)
#define AF_ENCCL (-6)    /* the collective library failed or is missing (data-parallel entry points only) */

typedef struct af_ctx af_ctx;
typedef struct af_mlp af_mlp;
typedef struct af_lstm af_lstm;

/* ---- context / memory ------------------------------------------------------------------- */
int af_ctx_create(int device, af_ctx** out);
/* Borrow an existing stream (e.g. torch's current stream); stream == NULL means legacy default. */
int af_ctx_create_on_stream(int device, void* cuda_stream, af_ctx** out);
int af_ctx_destroy(af_ctx* ctx);
int af_sync(af_ctx* ctx);
const char* af_last_error(void);
int af_version(void);
/* Number of kernels this library has launched on ctx since creation (bench.py's gpu_launches). */
int64_t af_launch_count(af_ctx* ctx);
/* Launch trace for measurement (bench.py's roofline): between begin and end, one CUDA event is
 * recorded on the ctx stream after every kernel launch.  af_trace_end synchronises and returns, per
 * launch, its kind (AF_KIND_*), its algorithmic work (flops for contractions, bytes for streaming
 * kernels) and the device time since the previous event in milliseconds.                        */
#define AF_KIND_GEMM_DUAL 0     /* TMA+DMMA contraction, value and R accumulators */
#define AF_KIND_GEMM_SINGLE 1   /* TMA+DMMA contraction, one accumulator */
#define AF_KIND_GEMM_GENERIC 2  /* generic contraction (operands not TMA-aligned) */
#define AF_KIND_ELEMENTWISE 3
#define AF_KIND_REDUCE 4
#define AF_KIND_SMALL_N 5       /* n <= 16 streaming kernels (the reference's Gemv/Ger branch); work in bytes */
int af_trace_begin(af_ctx* ctx, int capacity);
int af_trace_end(af_ctx* ctx, int capacity, int* h_kind, double* h_work, float* h_ms, int* h_count);
/* Path selector for the LinTran contractions: 0 = auto (TMA + DMMA tensor-core kernel when the
 * operands satisfy TMA alignment, else the generic kernel), 1 = force generic, 2 = require DMMA
 * (returns AF_ENOTSUP instead of falling back). */
int af_set_gemm_path(af_ctx* ctx, int mode);
/* Tuning knobs (measurement / A-B runs).  "big_cfg": tile configuration of the DMMA kernel for outputs
 * wider than 16 columns: 0 = 128x64 tile, one CTA per SM; 2 = 64x64 tile, two CTAs per SM.
 * "small_n": 1 (default) = batches of <= 16 samples use the streaming Gemv/Ger kernels, 0 = they go
 * through the tensor-core kernel (transposed narrow tiles + split-K).
 * "lstm_split": 1 = the LSTM time loops run the two halves of the lane batch as two concurrent kernel chains on
 * two streams (lanes are independent through the recurrence); 0 (default) = one chain -- measured 24.6 vs 24.2 ms on
 * the C4 LSTM: the half-batch contractions lose more efficiency than the overlap of tails and launch gaps returns.
 * Takes effect for graphs captured afterwards.
 * "pdl": 1 = the LSTM time loops launch their dependent kernel chain (contraction -> cell -> ...) with
 * programmatic dependent launch (each kernel's launch and prologue overlap its predecessor's tail); 0 (default) =
 * plain launches -- measured 24.6 ms vs 24.25 ms on the C4 LSTM: the stream-K contraction already keeps every CTA
 * slot busy until its end, so the early-launched successor only competes for the slots.  Takes effect for graphs
 * captured afterwards.
 * "sk_rotate": 1 (default) = stream-K contraction launches whose CTAs own >= 64 k-tiles accumulate each tile segment's
 * k-tiles in rotated order, so that all CTAs sweep k in lockstep and share operand panels in L2 (DRAM reads of the bench
 * weight gradient 2.44 GB -> 0.46 GB, same time); 0 = every segment ascending from its first k-tile.
 * "sk_bias": stream-K's cost in the work-decomposition model, percent (A/B of the model's threshold).
 * "persist_tiles": 1 | 2 = multi-wave one-tile-per-CTA launches as persistent CTAs over whole tiles (measured slower; only in
 * builds with -DAF_PERSIST_TILES=1, AF_ENOTSUP otherwise: keeping the mode compiled in cost the stream-K launches 2-3 %).
 * "small_mlp": 1 (default) = af_mlp_step with <= 4 samples runs as ONE persistent cooperative kernel
 * (all layers, both directions, grid barriers between phases); 0 = the per-kernel CUDA graph.
 * "small_mlp_head": 1 (default) = that kernel runs a top layer of <= 16 outputs (forward, top sigmoid backward, its
 * gradients) locally in every CTA instead of as two grid-wide phases: two grid barriers fewer; 0 = A/B.
 * "dp_max_ctas": CTA cap of the context's communicator (ncclConfig_t.maxCTAs; default 8, 0 = NCCL's own choice); read by
 * af_dp_init.  "dp_reserve_sms": SMs the plans leave to an all-reduce that runs beside their contractions (default =
 * dp_max_ctas): stream-K launches issued while a reduce is in flight are sized for sm_count - dp_reserve_sms SMs.
 * "few_samples": 1 (default) = LinTran passes with few_samples_min..16 samples ("few_samples_min", default 5) run as the
 * streaming tensor-core kernels of csrc/few_samples.cu (forward; weight + input gradient in one pass over W); 0 = the
 * Gemv / Ger kernels up to 8 samples and the transposed narrow tiles / rank-n tiles of the contraction kernel above.
 * "few_samples_blocks": resident blocks per SM those launches are sized for (default 3); "few_samples_tma": 1 (default) =
 * the backward pass brings its operands through warp-private TMA rings, 0 = register-staged loads.
 * "deterministic": 1 = every floating-point sum is taken in an order fixed by the shapes alone, so results are
 * bit-identical from run to run like the reference's (a sequential program): contractions compute whole output tiles per
 * CTA or stream-K ranges with the ordered fix-up ("sk_fixup"; no partial tiles combined with red.global.add.f64), column sums fold their per-block partial
 * sums in block order, and the atomics-based small-batch kernels and the fused head's bias-gradient atomics are not
 * used.  0 (default) = the faster schedule, whose last bits (well inside 1e-12 + 1e-10 |ref|) depend on the arrival
 * order of partial sums.  Cost measured on the bench step: DESIGN.md 3.1.  (The LSTM's fused time steps are
 * deterministic in both modes; the sum over ranks is NCCL's, which is run-to-run deterministic for a fixed communicator.)
 * "few_rows_wgrad": the weight gradient of a layer with <= 16 output rows over a batch of >= 64 samples (the bench
 * network's 10-class head) as a streaming kernel with ordered sums: 0 never, 1 (default) in deterministic mode, 2 always.
 * "lstm_fused": 1 (default) = the LSTM plans' time steps are the fused kernels (gate contraction + cell, cluster split-K +
 * cell backward); 0 = one contraction and one cell kernel per step (A/B).  "small_wgrad_max": batches up to this many
 * samples (default 8, at most 16) take the streaming rank-n weight-gradient update instead of tensor-core tiles.
 * "sk_fixup": stream-K launches whose partial tiles meet in an ORDERED fix-up instead of red.global.add.f64 (the CTA holding
 * a split tile's first k-tiles adds the following CTAs' parked partial sums in CTA order and runs the fused epilogue; no
 * zeroing, no deferred elementwise pass, bit-reproducible): 0 never, 1 (default) = everywhere in deterministic mode and for
 * store-type launches in the default mode, 2 = for every stream-K launch.  "sk_fix_bias": its cost in the decomposition
 * model for store-type launches, percent (default 103; A/B).
 * "fused_colsum": 1 (default) = the MLP plans' bias gradients (column sums of a layer's upstream, lin_add.go:94-104) are
 * added by the epilogue of the input-gradient launch that produces that upstream (float64 atomics; never in deterministic
 * mode); 0 = always a colsum_kernel pass.  "ops_runtime": 1 = the one-tile launches of the big contractions read the R
 * operands' presence at run time instead of taking the instantiation compiled for it (A/B; default 0).
 * "short_rows": 1 (default) = softmax over rows of <= 32 entries: per-entry work in the coalesced phases, one thread per row for the sums;
 * 0 = the staged group kernels that serve longer rows.
 * Plan-owned CUDA graphs captured before a change of any option are re-captured.                        */
int af_set_option(af_ctx* ctx, const char* name, int value);

int af_malloc(af_ctx* ctx, int64_t n_doubles, double** d_out);
int af_free(af_ctx* ctx, double* d_ptr);
int af_upload(af_ctx* ctx, double* d_dst, const double* h_src, int64_t n);    /* sync */
int af_download(af_ctx* ctx, double* h_dst, const double* d_src, int64_t n);  /* sync */
int af_zero(af_ctx* ctx, double* d_ptr, int64_t n);
int af_copy(af_ctx* ctx, double* d_dst, const double* d_src, int64_t n);

/* ---- LinTran: lin_tran.go --------------------------------------------------------------- */
/* W is rows x cols.  X, RX: n x cols.  Y, RY: n x rows.  U, UR: n x rows.  D, DR: n x cols.   */

/* G1  LinTran.multiply (lin_tran.go:79-117):            Y  = X W^T                             */
int af_lintran_fwd(af_ctx*, const double* d_W, const double* d_X, double* d_Y,
                   int64_t rows, int64_t cols, int64_t n);
/* G1+G2  LinTran.BatchR (lin_tran.go:64-77, 119-177):   Y  = X W^T ; RY = X RW^T + RX W^T
 * d_RW and/or d_RX may be NULL.  d_Y may be NULL (R-output only).                              */
int af_lintran_fwd_r(af_ctx*, const double* d_W, const double* d_RW, const double* d_X,
                     const double* d_RX, double* d_Y, double* d_RY,
                     int64_t rows, int64_t cols, int64_t n);
/* G3  LinTran.dataGradient (lin_tran.go:179-212):       gW += U^T X                            */
int af_lintran_wgrad(af_ctx*, const double* d_U, const double* d_X, double* d_gW,
                     int64_t rows, int64_t cols, int64_t n);
/* G3+G4  dataGradient + dataGradientR (lin_tran.go:214-270, called at :410-418):
 *        gW += U^T X   (skipped when d_gW == NULL, i.e. grad == nil or W not in grad)
 *        rgW += U^T RX + UR^T X  (skipped when d_rgW == NULL)                                  */
int af_lintran_wgrad_r(af_ctx*, const double* d_U, const double* d_UR, const double* d_X,
                       const double* d_RX, double* d_gW, double* d_rgW,
                       int64_t rows, int64_t cols, int64_t n);
/* G5  LinTran.inputGradient (lin_tran.go:272-306):      D  = U W                               */
int af_lintran_igrad(af_ctx*, const double* d_U, const double* d_W, double* d_D,
                     int64_t rows, int64_t cols, int64_t n);
/* G5+G6  inputGradient + inputGradientR (lin_tran.go:308-359, called at :421-422):
 *        D = U W ; DR = UR W + U RW                                                            */
int af_lintran_igrad_r(af_ctx*, const double* d_U, const double* d_UR, const double* d_W,
                       const double* d_RW, double* d_D, double* d_DR,
                       int64_t rows, int64_t cols, int64_t n);

/* linTranResult.PropagateGradient / linTranRResult.PropagateRGradient as ONE call (lin_tran.go:361-391 / 393-425): the
 * weight gradient (G3+G4) and the input gradient (G5+G6) from the same upstream.
 *   d_gW / d_rgW NULL: W is not in that map (skipped);  d_D and d_DR both NULL: the input is constant (lin_tran.go:419).
 *   accumulate_input != 0: D += U W and DR += UR W + U RW -- what the input's own PropagateRGradient does next when it is
 *   a Variable (variable.go:43-47,113-123) -- instead of D = ..., DR = ...
 * For 9..16 samples (the reference's LinTranBenchmark runs 10) one streaming pass over W serves both gradients; other
 * batch sizes run the two passes of af_lintran_wgrad_r and af_lintran_igrad_r.                     */
int af_lintran_backward(af_ctx*, const double* d_U, const double* d_W, const double* d_X, double* d_gW, double* d_D,
                        int64_t rows, int64_t cols, int64_t n, int accumulate_input);
int af_lintran_backward_r(af_ctx*, const double* d_U, const double* d_UR, const double* d_W, const double* d_RW,
                          const double* d_X, const double* d_RX, double* d_gW, double* d_rgW, double* d_D,
                          double* d_DR, int64_t rows, int64_t cols, int64_t n, int accumulate_input);

/* ---- LinAdd (lin_add.go) in its batched form Add(x, Repeat(b,n)) ------------------------- */
/* Y = X + b (per sample), RY = RX + Rb.  lin_add.go:13-50; arithmetic.go:17; slices.go:216.    */
int af_bias_fwd(af_ctx*, const double* d_X, const double* d_b, double* d_Y, int64_t len, int64_t n);
int af_bias_fwd_r(af_ctx*, const double* d_X, const double* d_RX, const double* d_b,
                  const double* d_Rb, double* d_Y, double* d_RY, int64_t len, int64_t n);
/* gb += sum over samples of U (lin_add.go:66-69,94-104; Repeat backward slices.go:244-250).
 * Either accumulator may be NULL.                                                              */
int af_bias_grad(af_ctx*, const double* d_U, double* d_gb, int64_t len, int64_t n);
int af_bias_grad_r(af_ctx*, const double* d_U, const double* d_UR, double* d_gb, double* d_rgb,
                   int64_t len, int64_t n);

/* ---- elementwise RFuncs: math_funcs.go --------------------------------------------------- */
/* Sigmoid (math_funcs.go:286-374).  fwd_r: S = 1/(1+exp(-X)), RS = S(1-S) RX.
 * bwd: U <- U S(1-S) in place (:328-330).  bwd_r (:365-371), in place, reference operand order:
 *   U' = S(1-S)U ;  UR' = RS(1-S)U - S RS U + S(1-S)UR                                         */
int af_sigmoid_fwd(af_ctx*, const double* d_X, double* d_S, int64_t len);
int af_sigmoid_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_S, double* d_RS, int64_t len);
int af_sigmoid_bwd(af_ctx*, const double* d_S, double* d_U, int64_t len);
int af_sigmoid_bwd_r(af_ctx*, const double* d_S, const double* d_RS, double* d_U, double* d_UR, int64_t len);
/* Exp (math_funcs.go:12-97): E = exp(X), RE = E RX; bwd U <- U E; bwd_r UR <- UR E + U RE.     */
int af_exp_fwd(af_ctx*, const double* d_X, double* d_E, int64_t len);
int af_exp_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_E, double* d_RE, int64_t len);
int af_exp_bwd(af_ctx*, const double* d_E, double* d_U, int64_t len);
int af_exp_bwd_r(af_ctx*, const double* d_E, const double* d_RE, double* d_U, double* d_UR, int64_t len);

/* ---- arithmetic.go hot subset ------------------------------------------------------------ */
/* out = a + b (arithmetic.go:17,58) ; out = a - b (:105,144) ; out = a * b (:261)             */
int af_add(af_ctx*, const double* d_a, const double* d_b, double* d_out, int64_t len);
int af_sub(af_ctx*, const double* d_a, const double* d_b, double* d_out, int64_t len);
int af_mul(af_ctx*, const double* d_a, const double* d_b, double* d_out, int64_t len);
/* MulR R-output (arithmetic.go:325-329): out = a*bR + aR*b                                     */
int af_mul_r(af_ctx*, const double* d_a, const double* d_aR, const double* d_b, const double* d_bR,
             double* d_out, int64_t len);
/* Mul backward toward one operand (arithmetic.go:291-293,358-362): D = U*o ; DR = U*oR + UR*o.
 * d_UR/d_oR/d_DR may be NULL for the non-R form.                                               */
int af_mul_bwd(af_ctx*, const double* d_U, const double* d_UR, const double* d_o, const double* d_oR,
               double* d_D, double* d_DR, int64_t len);
/* x <- f*x in place (Scale, arithmetic.go:436-493) ; y += f*x (gradient.go:40-52 Axpy)         */
int af_scale(af_ctx*, double* d_x, double f, int64_t len);
int af_axpy(af_ctx*, double f, const double* d_x, double* d_y, int64_t len);
/* Square (arithmetic.go:696-770): fwd_r O = X^2, RO = 2 X RX; bwd_r UR <- 2(UR X + U RX), U <- 2 X U */
int af_square_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_O, double* d_RO, int64_t len);
int af_square_bwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_U, double* d_UR, int64_t len);
/* SumAll (arithmetic.go:972-1047): *d_out = sum(x); (d_outR = sum(xR) when both non-NULL)      */
int af_sum_all(af_ctx*, const double* d_x, const double* d_xR, double* d_out, double* d_outR, int64_t len);
/* fill (SumAll backward broadcast, arithmetic.go:990-993)                                      */
int af_fill(af_ctx*, double* d_x, double value, int64_t len);

/* ---- the functions and cost heads either side of the dense path -------- */
/* Conventions: every *_fwd_r computes the value form, and the R form too when the R output pointer is
 * non-NULL (a NULL R input means identically zero, variable.go:85-91).  Every *_bwd_r overwrites the
 * upstreams in place as the reference does (result.go:24-29); d_UR may be NULL for the non-R form.
 * Scalars the reference reads on the host (scaler.Output()[0], MaxAbs(), SumAll's upstream) are device
 * scalars passed by pointer, so chains of these calls never synchronise the stream.                  */
/* Log (math_funcs.go:99-189): O = log X, RO = RX / X; bwd U <- U/X, UR <- UR/X - U RX / X^2          */
int af_log_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_O, double* d_RO, int64_t len);
int af_log_bwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_U, double* d_UR, int64_t len);
/* LogSigmoid (math_funcs.go:376-477): branch on the sign of x as :399-410; RO = RX / (1 + e^x)        */
int af_logsigmoid_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_O, double* d_RO, int64_t len);
int af_logsigmoid_bwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_U, double* d_UR, int64_t len);
/* Sin (math_funcs.go:511-597); Cos is Sin(AddScaler(x, pi/2)) (:599-607)                              */
int af_sin_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_O, double* d_RO, int64_t len);
int af_sin_bwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_U, double* d_UR, int64_t len);
/* Tanh: EXTENSION named by BASELINE.json's north_star (the reference has none); Sigmoid's contract.   */
int af_tanh_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_T, double* d_RT, int64_t len);
int af_tanh_bwd_r(af_ctx*, const double* d_T, const double* d_RT, double* d_U, double* d_UR, int64_t len);
/* Inverse (arithmetic.go:772-867): O = 1/X, RO = -O^2 RX; bwd takes the forward OUTPUT d_O.           */
int af_inverse_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_O, double* d_RO, int64_t len);
int af_inverse_bwd_r(af_ctx*, const double* d_O, const double* d_RX, double* d_U, double* d_UR, int64_t len);
/* Pow (arithmetic.go:869-966)                                                                         */
int af_pow_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double p, double* d_O, double* d_RO, int64_t len);
int af_pow_bwd_r(af_ctx*, const double* d_X, const double* d_RX, double p, double* d_U, double* d_UR, int64_t len);
/* Div (arithmetic.go:378-423): O = A/B; bwd DA = U/B, DB = U(-O/B) (either may be NULL).  DivR is
 * MulR(a, InverseR(b)) in the reference (:425-427) and is composed the same way by the caller.        */
int af_div(af_ctx*, const double* d_A, const double* d_B, double* d_O, int64_t len);
int af_div_bwd(af_ctx*, const double* d_U, const double* d_O, const double* d_B, double* d_DA, double* d_DB, int64_t len);
/* AddScaler (arithmetic.go:184-252): O = X + f                                                        */
int af_add_scaler(af_ctx*, const double* d_X, double f, double* d_O, int64_t len);
/* O = X + sign * d_f[0]: AddFirst (arithmetic.go:595-694) and the "- maxVal" of the log-domain sums   */
int af_add_dev(af_ctx*, const double* d_X, const double* d_f, double sign, double* d_O, int64_t len);
/* O = X d_f[0]; RO = X d_fR[0] + RX d_f[0]: ScaleFirst forward and (on the upstreams) backward
 * (arithmetic.go:495-593).  O may alias X and RO may alias RX.                                         */
int af_scale_dev(af_ctx*, const double* d_X, const double* d_RX, const double* d_f, const double* d_fR,
                 double* d_O, double* d_RO, int64_t len);
/* x[i] = d_value[0]: SumAll's backward broadcast (arithmetic.go:987-995,1035-1047)                    */
int af_fill_dev(af_ctx*, double* d_x, const double* d_value, int64_t len);
/* *d_out = sum a.b (+ sum c.d when c, d non-NULL): DotFast (arithmetic.go:523,571-573)                */
int af_dot(af_ctx*, const double* d_a, const double* d_b, const double* d_c, const double* d_d, double* d_out, int64_t len);
/* *d_out = max |x| (linalg MaxAbs, arithmetic.go:1057,1078); combine != 0: max with the value already there */
int af_max_abs(af_ctx*, const double* d_x, double* d_out, int combine, int64_t len);
/* Norm (math_funcs.go:201-284): out = ||X||, outR = <X,RX>/out; bwd D = X u/o, DR = X(-u ro/o^2 + uR/o) + RX u/o */
int af_norm_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double* d_out, double* d_outR, int64_t len);
int af_norm_bwd_r(af_ctx*, const double* d_X, const double* d_RX, const double* d_o, const double* d_ro,
                  const double* d_u, const double* d_uR, double* d_D, double* d_DR, int64_t len);
/* Softmax of every row of an n x len packed batch = RFuncBatcher{&Softmax{T}} (math_funcs.go:479-509,
 * batcher.go:57-84) as ONE kernel per direction; temperature 0 means 1 (:481-485).  No max-subtraction,
 * like the reference.  bwd_r is in place on U, UR and needs only the forward outputs Y, RY.            */
int af_softmax_fwd_r(af_ctx*, const double* d_X, const double* d_RX, double temperature, double* d_Y, double* d_RY,
                     int64_t len, int64_t n);
int af_softmax_bwd_r(af_ctx*, const double* d_Y, const double* d_RY, double temperature, double* d_U, double* d_UR,
                     int64_t len, int64_t n);
/* Squared-error cost head (north_star's "SquaredError"): c = .5 ||A - Y||^2, the reference's
 * Scale(SquaredNorm(Sub(a, y)), .5) (math_funcs.go:191-199, arithmetic.go:105,436,696,972) as one
 * reduction; bwd: upstream (d_u, d_uR device scalars) -> upstream on A (sign = +1) or on Y (sign = -1). */
int af_sqerr_fwd_r(af_ctx*, const double* d_A, const double* d_RA, const double* d_Y, const double* d_RY,
                   double* d_out, double* d_outR, int64_t len);
int af_sqerr_bwd_r(af_ctx*, const double* d_A, const double* d_RA, const double* d_Y, const double* d_RY,
                   const double* d_u, const double* d_uR, double sign, double* d_G, double* d_GR, int64_t len);

/* ---- recorded launch sequences (CUDA graphs) --------------------------------------------------
 * The operator-level calls above are one kernel launch each; with a handful of samples (the
 * reference's own benchmarks run n = 1 and n = 10, bench/mlp.go:36, bench/lintran.go:13-17) a chain
 * of them is bound by launch gaps, not by HBM.  Between af_graph_begin and af_graph_end every
 * launch issued on the context is RECORDED instead of executed; af_graph_launch replays the whole
 * sequence -- same buffers, same sizes -- as one submission.  Run the sequence once eagerly first
 * (first calls size internal scratch and set kernel attributes).  Calls that synchronise, allocate
 * or manage graphs of their own (af_sync, af_malloc, af_free, af_upload, af_download, af_mlp_*
 * steps, af_lstm_*) fail with AF_EINVAL while recording; af_graph_end must still be called
 * (out == NULL abandons the recording).  af_launch_count advances by the recorded number of
 * launches on every replay.                                                                     */
typedef struct af_graph af_graph;
int af_graph_begin(af_ctx*);
int af_graph_end(af_ctx*, af_graph** out);
int af_graph_launch(af_graph*);
int64_t af_graph_launches(af_graph*);
int af_graph_destroy(af_graph*);

/* ---- fused dense layer: LinTran -> LinAdd -> Sigmoid in one kernel per direction ---------- */
/* Forward with R (G1+G2+E1+E2):  S = sigmoid(X W^T + b) ; RS = S(1-S)(X RW^T + RX W^T + Rb).
 * activation: 0 = none (S = pre-activation), 1 = sigmoid.  d_b/d_Rb/d_RW/d_RX/d_RS may be NULL
 * (d_RS == NULL selects the non-R forward).                                                    */
int af_dense_fwd(af_ctx*, const double* d_W, const double* d_RW, const double* d_b, const double* d_Rb,
                 const double* d_X, const double* d_RX, double* d_S, double* d_RS,
                 int64_t rows, int64_t cols, int64_t n, int activation);
/* Input-gradient with the PREVIOUS layer's sigmoid backward fused in (G5+G6+E4):
 *   D = U W ; DR = UR W + U RW ; then, when d_Sprev != NULL, the in-place sigmoid R-backward of
 *   math_funcs.go:365-371 with (S,RS) = (Sprev, RSprev), written to d_D / d_DR.  d_D / d_DR must not overlap
 *   d_Sprev / d_RSprev (the kernels read them through the read-only path).                      */
int af_dense_igrad(af_ctx*, const double* d_U, const double* d_UR, const double* d_W, const double* d_RW,
                   const double* d_Sprev, const double* d_RSprev, double* d_D, double* d_DR,
                   int64_t rows, int64_t cols, int64_t n);

/* ---- whole-network fast path: bench/mlp.go:24-126 with n >= 1 samples ---------------------- */
/* sizes[0..n_layers] = layer widths (bench/mlp.go:17-20: 1000,2000,512,10); batch = n.
 * Parameters, R-vectors, activations and the four accumulators per layer stay resident in HBM. */
int af_mlp_create(af_ctx*, const int64_t* sizes, int n_layers, int64_t batch, af_mlp** out);
int af_mlp_destroy(af_mlp*);
/* Parameter l in [0, 2*n_layers): even = W of layer l/2, odd = bias (order of bench/mlp.go:84-85). */
int af_mlp_set_param(af_mlp*, int index, const double* h_value, int64_t len);
int af_mlp_set_rvector(af_mlp*, int index, const double* h_value, int64_t len); /* NULL h_value => absent */
double* af_mlp_param_ptr(af_mlp*, int index);   /* device pointers for zero-copy callers */
double* af_mlp_rvector_ptr(af_mlp*, int index);
double* af_mlp_grad_ptr(af_mlp*, int index);
double* af_mlp_rgrad_ptr(af_mlp*, int index);
double* af_mlp_input_ptr(af_mlp*);              /* n x sizes[0] */
double* af_mlp_output_ptr(af_mlp*);             /* n x sizes[L] */
double* af_mlp_routput_ptr(af_mlp*);
double* af_mlp_upstream_ptr(af_mlp*);           /* n x sizes[L]: outGrad  (bench/mlp.go:118-126) */
double* af_mlp_upstream_r_ptr(af_mlp*);         /* n x sizes[L]: outRGrad */
int af_mlp_zero_grads(af_mlp*);
/* Gradient.AddToVars (gradient.go:40-52) on the plan's own parameters: param_i += scale * grad_i, in HBM
 * (a training loop that never downloads the gradient).                                  */
int af_mlp_add_to_vars(af_mlp*, double scale);
int af_mlp_get_param(af_mlp*, int index, double* h_value, int64_t len);      /* sync */
/* One iteration of bench/mlp.go:52-54 on resident data: ApplyR, then PropagateRGradient with
 * non-nil grad, accumulating into grad/rgrad.  with_r == 0 runs bench/mlp.go:36-38 instead
 * (Apply + PropagateGradient).  The upstream buffers are consumed (clobbered), as the reference
 * allows (result.go:24-29).  Asynchronous.                                                     */
int af_mlp_step(af_mlp*, int with_r, int backprop);
/* Replay the step from a CUDA graph (captured on the second call of a configuration).  On by default for
 * batches <= 256, where the ~25 short kernels of a step are launch-bound.                          */
/* One iteration of the reference benchmark as ONE call (bench/mlp.go:43-57 with the fresh Gradient maps of
 * :33-35 and outGrad = 1 / outRGrad = 0 of :118-126): equivalent to af_mlp_zero_grads + fill the upstreams +
 * af_mlp_step(with_r, 1).  For batches of <= 4 samples the persistent kernel stores the accumulators instead of
 * read-modify-writing them and generates the upstream in registers, so the zeroing passes disappear.  After a
 * non-R fresh step the R accumulators are unspecified (the reference has no RGradient map in Run()).       */
int af_mlp_step_fresh(af_mlp*, int with_r);
int af_mlp_set_graph(af_mlp*, int enable);
/* Host-buffer form (the reference-facing call): uploads h_X (n x sizes[0]), h_upstream /
 * h_upstream_r (n x sizes[L]; NULL => ones / zeros as in bench/mlp.go:118-126), zeroes the
 * accumulators, runs the step, and downloads h_out / h_rout (n x sizes[L]) and every gradient and
 * R-gradient into h_grads[i] / h_rgrads[i] (arrays of 2*n_layers host pointers; entries or the
 * arrays themselves may be NULL).  Synchronous.                                                */
int af_mlp_step_host(af_mlp*, int with_r, const double* h_X, const double* h_upstream,
                     const double* h_upstream_r, double* h_out, double* h_rout,
                     double* const* h_grads, double* const* h_rgrads);

/* Pipelined form of af_mlp_step_host: returns as soon as the step is queued.  Two steps may be in
 * flight: while step k computes, step k-1 downloads its results on a copy stream and step k+1 uploads
 * its inputs on another.  Host buffers should be pinned and must stay valid until the matching
 * af_mlp_wait(), which waits for the OLDEST queued step (its host outputs are then complete).
 * A third submit first waits for the oldest step.                                                */
int af_mlp_submit_host(af_mlp*, int with_r, const double* h_X, const double* h_upstream,
                       const double* h_upstream_r, double* h_out, double* h_rout,
                       double* const* h_grads, double* const* h_rgrads);
int af_mlp_wait(af_mlp*);
/* The same step in two phases, for data-parallel callers: _compute_ uploads and runs the step into the next
 * slot; the caller may then queue work on the ctx stream that the download must wait for (the NCCL
 * all-reduce of that slot's accumulators); _download_ queues the device-to-host copies behind it.   */
int af_mlp_submit_compute_host(af_mlp*, int with_r, const double* h_X, const double* h_upstream,
                               const double* h_upstream_r);
int af_mlp_submit_download_host(af_mlp*, int with_r, double* h_out, double* h_rout,
                                double* const* h_grads, double* const* h_rgrads);
/* Make slot 0 or 1 (input batch, accumulators, outputs) the one af_mlp_step / af_mlp_*_ptr refer to.   */
int af_mlp_bind_slot(af_mlp*, int slot);
/* Data-parallel hook.  During af_mlp_step, `fn(user, d_ptr, n_doubles)` is called on the host as soon as
 * the launches that complete the accumulator range [d_ptr, d_ptr + n_doubles) are on the ctx stream, so the
 * caller can start that range's all-reduce (gradient.go:28-30 across ranks) while the rest of the backward
 * pass runs.  Accumulators are laid out one block per layer (gW, gb, rgW, rgb); layer 1's weight gradient --
 * the largest, and the last to be produced -- is computed in `layer1_chunks` row chunks, each reported
 * separately.  fn == NULL removes the hook.                                                          */
typedef void (*af_grad_ready_fn)(void* user, double* d_ptr, int64_t n_doubles);
int af_mlp_set_grad_ready(af_mlp*, af_grad_ready_fn fn, void* user, int layer1_chunks);

/* ---- data parallelism: Gradient.Add across ranks (gradient.go:28-30,108-120) ---------------------
 * The path shards over the batch: rank r holds its rows of X; parameters and R-vectors are
 * replicated; the parameter gradients and R-gradients are SUMS over samples, so the sum of the ranks'
 * accumulators is the full-batch gradient.  One af_ctx = one device = one rank; the transport is NCCL
 * (ncclAllReduce, ncclDouble, ncclSum) over NVLink / NVSwitch, resolved at run time (the library loads on
 * hosts without NCCL; the entry points below then return AF_ENCCL).
 *   several processes, one device each : rank 0 calls af_dp_unique_id and ships the 128 bytes to the other
 *                                        ranks by any means; every rank calls af_dp_init.
 *   one process, several devices       : af_dp_init_local over the contexts (one per device).  Calls that
 *                                        issue collectives must then come from one host thread per context,
 *                                        or be bracketed by af_dp_group_begin / af_dp_group_end.          */
#define AF_DP_ID_BYTES 128
int af_dp_unique_id(void* id128);
int af_dp_init(af_ctx*, int nranks, int rank, const void* id128);
int af_dp_init_local(af_ctx* const* ctxs, int n);
int af_dp_destroy(af_ctx*);
int af_dp_info(af_ctx*, int* nranks, int* rank, int* nccl_version);   /* any pointer may be NULL */
int af_dp_group_begin(void);
int af_dp_group_end(void);
int64_t af_dp_calls(af_ctx*);   /* collectives issued through ctx so far */
/* d_ptr[0..n) <- sum over ranks, in place, in order on the ctx stream (asynchronous).                   */
int af_allreduce_sum(af_ctx*, double* d_ptr, int64_t n);
/* Plan level.  A data-parallel step = a step from FRESH (zeroed) accumulators followed by the sum of
 * every rank's {grad | rgrad} block; how the reduce is scheduled is the plan's dp mode:
 *   AF_DP_BLOCKING  one all-reduce of the whole block on the ctx stream after the backward pass
 *   AF_DP_OVERLAP   (default) the backward pass runs its critical chain (input gradients) first, then the weight
 *                   gradients largest layer first; each layer block's all-reduce starts on the communicator's own
 *                   stream as soon as the block is complete and runs beside the remaining contractions on
 *                   "dp_max_ctas" SMs; the ctx stream joins at the end of the step
 *   AF_DP_DEFERRED  as AF_DP_OVERLAP, but the join is deferred to the next use of the accumulators by the library
 *                   (af_mlp_zero_grads, the next fresh step -- which runs its forward pass first --, af_mlp_add_to_vars,
 *                   any host download) or af_mlp_dp_join: the tail of the reduce overlaps the next forward pass.
 * af_mlp_step_dp: af_mlp_step(with_r, backprop = 1) + the reduce; the caller zeroed the accumulators.
 * af_mlp_step_fresh_dp: af_mlp_step_fresh + the reduce.  The host-buffer forms (af_mlp_step_host,
 * af_mlp_submit_*_host) reduce whenever the plan's mode is not AF_DP_OFF.  Without a communicator (or with one
 * rank) every form degenerates to its local step.                                                       */
#define AF_DP_OFF 0
#define AF_DP_BLOCKING 1
#define AF_DP_OVERLAP 2
#define AF_DP_DEFERRED 3
int af_mlp_set_dp(af_mlp*, int mode);
int af_mlp_step_dp(af_mlp*, int with_r);
int af_mlp_step_fresh_dp(af_mlp*, int with_r);
int af_mlp_dp_join(af_mlp*);    /* the ctx stream waits for the plan's outstanding reduces */
/* The contiguous {grad | rgrad} block of the bound slot: returns its length in doubles.                 */
int64_t af_mlp_grad_slab(af_mlp*, double** d_first);
/* Host downloads of a data-parallel plan.  enable != 0: the host-buffer forms copy back only the part of the
 * all-reduced {grad | rgrad} block this rank owns -- elements [rank * S / nranks, (rank + 1) * S / nranks) of the
 * block in parameter order (gW0, gb0, rgW0, rgb0, gW1, ...), written at their usual offsets inside h_grads[i] /
 * h_rgrads[i] -- so the reduced gradient crosses to the host once per step instead of once per rank.      */
int af_mlp_set_host_shard(af_mlp*, int enable);
/* One data-parallel LSTM iteration on resident data: zero the accumulators, forward, BPTT, one all-reduce of
 * the {grad | rgrad} block on the ctx stream (the step is ~100x longer than its reduce).                  */
int af_lstm_step_dp(af_lstm*, int with_r);

/* dst[r*ld_dst + c] = src[r*ld_src + c]: the copy behind Concat / Slice of packed batches
 * (slices.go:13-119,130-206) when the parts are interleaved per sample.                          */
int af_copy_2d(af_ctx*, double* d_dst, int64_t ld_dst, const double* d_src, int64_t ld_src, int64_t rows, int64_t cols);

/* ---- LSTM fast path: bench/lstm.go:25-241, batched over B lanes, R-operator carried through ---- */
/* Parameter index 0..9 in reference order (bench/lstm.go:75-99, 200-213): InWeights, InBiases,
 * InGateWeights, InGateBiases, ForgetGateWeights, ForgetGateBiases, OutputGate, OutputGateBiases,
 * OutputWeights, OutputBiases.  Gate matrices are H x (H+I) over concat(state, x) (lstm.go:140);
 * OutputWeights is O x (H+I) over concat(masked, x) (lstm.go:193).  Sequences are time-major and
 * lane-major within a step: T x B x I inputs, T x B x O outputs / upstreams (the packing
 * seqfunc.MapBatcher produces, seqfunc/map.go:262-284).                                           */
int af_lstm_create(af_ctx*, int64_t input_size, int64_t hidden_size, int64_t output_size, int64_t time_steps,
                   int64_t lanes, af_lstm** out);
int af_lstm_destroy(af_lstm*);
int af_lstm_set_param(af_lstm*, int index, const double* h_value, int64_t len);
int af_lstm_set_rvector(af_lstm*, int index, const double* h_value, int64_t len); /* NULL => absent (zero R) */
double* af_lstm_param_ptr(af_lstm*, int index);
double* af_lstm_grad_ptr(af_lstm*, int index);
double* af_lstm_rgrad_ptr(af_lstm*, int index);
double* af_lstm_output_ptr(af_lstm*);
double* af_lstm_routput_ptr(af_lstm*);
double* af_lstm_upstream_ptr(af_lstm*);
double* af_lstm_upstream_r_ptr(af_lstm*);
/* The contiguous {grad | rgrad} block (for one data-parallel all-reduce): returns its length in doubles. */
int64_t af_lstm_grad_slab(af_lstm*, double** d_first);
int af_lstm_zero_grads(af_lstm*);
int af_lstm_add_to_vars(af_lstm*, double scale);                /* gradient.go:40-52, in HBM */
int af_lstm_set_inputs(af_lstm*, const double* d_x);          /* device T x B x I */
int af_lstm_forward(af_lstm*, int with_r);                   /* Reset + T x StepTime (lstm.go:42-45) */
int af_lstm_backward(af_lstm*, int with_r);                  /* PropagateGradient (lstm.go:215-234); consumes upstreams */
int af_lstm_run_host(af_lstm*, int with_r, const double* h_inputs, const double* h_upstreams,
                     const double* h_upstreams_r, double* h_outputs, double* h_routputs,
                     double* const* h_grads, double* const* h_rgrads);

#ifdef __cplusplus
}
#endif
#endif /* AUTOFUNC_B200_H_ */
