// Shared host-side plumbing for libautofunc_b200: context, error reporting, launch accounting.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <initializer_list>
#include <string>
#include <vector>

#include "../../include/autofunc_b200.h"

struct af_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int sm_count = 148;
  int gemm_path = 0;  // 0 auto, 1 generic, 2 require DMMA
  int big_cfg = 2;    // tile configuration for wide outputs: 0 = 128x64 x 8 warps, 2 = 64x64 x 2 CTAs/SM (default, fastest measured), 3 = 128x64 x 16 warps
  int persist_tiles = 0;  // multi-wave one-tile-per-CTA launches as persistent CTAs over whole tiles: 0 off, 1 contiguous tile ranges, 2 wave-ordered (tile = slot + step * grid)
  // stream-K's cost in the work-decomposition model, percent, for short (< 64 k-tiles per CTA) and long unit ranges
  // (af_set_option "sk_bias" sets both: A/B of the model's threshold)
  int sk_bias = 103, sk_bias_long = 96;
  // stream-K launches with long unit ranges walk each tile segment's k-tiles in rotated order (gemm_dmma.cuh: seg_start)
  // so that all CTAs sweep k in lockstep and share operand panels in L2 (af_set_option "sk_rotate" 0/1: A/B)
  int sk_rotate = 1;
  // batches up to this size take the streaming rank-n weight-gradient update (af_set_option "small_wgrad_max").  Round 1: 16
  // (C1, n = 10: 29.6 us streaming vs 41.0 us as tensor-core tiles); since the accumulate epilogue fetches a fragment row of
  // C before adding to it the tiles take 22.5 us there, so 9..16 samples go to the tensor cores (C1 graph 55.4 -> 49.5 us)
  int small_wgrad_max = 8;
  // af_set_option "deterministic" 1: every sum is taken in an order fixed by the shapes alone -- contractions run whole
  // tiles per CTA (no split-K / stream-K partial tiles through red.global.add.f64), column sums fold their partial sums in
  // block order, the atomics-based small-batch kernels (small_n.cu input gradient, small_mlp.cu) and the fused head's
  // per-block bias-gradient atomics are not used.  Results are then bit-identical from run to run, like the reference's
  // (a sequential program); the default (0) is the faster schedule whose last bits depend on the arrival order of partials.
  bool deterministic = false;
  // LinTran passes for few_samples_min <= n <= 16 samples run as the streaming tensor-core kernels of few_samples.cu
  // (af_set_option "few_samples" 0/1, "few_samples_min"): forward, and the weight + input gradient in ONE pass over W
  bool few_samples_off = false;
  int few_samples_min = 5;
  // the <= 16-row weight gradient over a large batch (few_rows_wgrad_kernel): 0 never, 1 in deterministic mode only, 2 always
  int few_rows_wgrad = 1;
  int few_samples_blocks = 3;  // resident blocks per SM the few-samples launches are sized for ("few_samples_blocks")
  bool few_samples_tma = true;  // backward pass through warp-private TMA rings ("few_samples_tma" 0: the register-staged kernel)
  unsigned* fold_tickets = nullptr;  // ticket counters of the last-block-folds reductions (kFoldTickets, zero between launches)
  bool short_rows_off = false;  // A/B switch ("short_rows" 0): rows of <= 32 entries through the staged group kernels (mathops.cu)
  bool ops_runtime = false;  // A/B switch ("ops_runtime" 1): the big contractions read the R operands' presence at run time (gemm_dmma_kernel OPS = -1)
  int sk_fixup = 1;               // stream-K with the ordered fix-up ("sk_fixup"): 0 never, 1 deterministic mode + store-type launches, 2 wherever stream-K is chosen
  int sk_fix_bias = 103;          // its cost in the decomposition model for store-type launches, percent ("sk_fix_bias")
  double* sk_part = nullptr;      // its workspace: one partial tile per CTA ...
  unsigned* sk_flags = nullptr;   // ... and one flag per CTA and warp (gemm.cu: fixup_workspace)
  bool fused_colsum_off = false;  // A/B switch ("fused_colsum" 0): the MLP plan's bias gradients always as colsum_kernel passes
  bool small_n_off = false;  // A/B switch: route n <= 16 through the tensor-core kernel instead of small_n.cu
  bool small_mlp_head_off = false;  // A/B switch ("small_mlp_head" 0): the persistent step runs the top layer as two grid-wide phases again
  bool small_mlp_off = false;  // A/B switch: af_mlp_step with n <= 8 replays the per-kernel graph instead of small_mlp.cu
  bool lstm_split_off = true;   // opt-in (af_set_option "lstm_split" 1): two concurrent half-batch kernel chains; measured 24.6 vs 24.2 ms
  bool lstm_fused_off = false;  // A/B switch (af_set_option "lstm_fused" 0): the per-kernel LSTM steps of round 1
  bool pdl = false;      // inside a dependent-kernel chain (LSTM time loops): launch with programmatic stream serialization
  bool pdl_off = true;   // per-kernel LSTM chains: opt-in (af_set_option "pdl" 1): measured 24.6 vs 24.25 ms on the C4 LSTM
  bool pdl_fused_off = true;   // fused LSTM steps (lstm_step.cu): opt-in as well -- forward 9.10 vs 9.23 ms, but the backward
                               // chain's cluster launches ran 23.5 vs 16.2 ms under programmatic dependent launch
  unsigned* barrier_word = nullptr;  // grid-barrier counter of the persistent small-batch kernel
  unsigned barrier_base = 0;         // its value once every launch queued so far has finished
  int64_t launches = 0;
  // af_graph_begin / af_graph_end: launches on `stream` are being recorded into a CUDA graph, not executed
  bool recording = false;
  int64_t record_start = 0;
  // driver entry point, resolved at run time so the library loads on hosts without libcuda
  void* encode_tiled = nullptr;
  double* scratch = nullptr;  // small device scratch (reductions)
  int64_t scratch_doubles = 0;
  // launch trace (af_trace_begin/af_trace_end): one CUDA event after every kernel launch
  bool tracing = false;
  std::vector<cudaEvent_t> trace_events;  // [0] = start marker, [i+1] = after launch i
  std::vector<int> trace_kind;
  std::vector<double> trace_work;
  size_t trace_used = 0;
  // bumped by af_set_option / af_set_gemm_path: plan-owned CUDA graphs captured under an older epoch are re-captured
  int64_t option_epoch = 0;
  // ---- data parallelism (dp.cu): one NCCL communicator per context, its own high-priority stream ----
  void* comm = nullptr;           // ncclComm_t, capped at dp_max_ctas CTAs: reduces that overlap the contractions
  void* comm_wide = nullptr;      // its uncapped twin (ncclCommSplit) for reduces that run alone; NULL: use `comm`
  cudaEvent_t dp_order_ev = nullptr;
  int dp_rank = 0, dp_nranks = 1;
  cudaStream_t dp_stream = nullptr;
  int dp_max_ctas = 8;            // CTA cap of the communicator (af_set_option "dp_max_ctas", before af_dp_init)
  // SMs left to a collective that runs beside the contractions: stream-K launches are sized for
  // (sm_count - reserved_sms) SMs while > 0 (set by the plans around their overlapped all-reduces)
  int reserved_sms = 0;
  int dp_reserve_sms = -1;        // af_set_option "dp_reserve_sms"; -1 = dp_max_ctas
  int64_t dp_calls = 0;           // collectives issued through this context (bench.py reports it)
};

namespace af {

// device scratch of a context (16 MB): partial sums of SumAll, of the ordered column sums (deterministic mode) and of the
// few-samples kernels' split extents.  Fixed size: captured graphs hold its address.
constexpr int64_t kScratchDoubles = 1ll << 21;
constexpr int kFoldTickets = 4096;

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define AF_CUDA(expr)                                          \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::af::cuda_fail(_e, #expr);  \
  } while (0)

#define AF_TRY(expr)          \
  do {                        \
    int _rc = (expr);         \
    if (_rc != AF_OK) return _rc; \
  } while (0)

// kernel kinds reported by the launch trace (include/autofunc_b200.h: AF_KIND_*)
enum : int { KIND_GEMM_DUAL = 0, KIND_GEMM_SINGLE = 1, KIND_GEMM_GENERIC = 2, KIND_ELEMENTWISE = 3, KIND_REDUCE = 4, KIND_SMALL_N = 5 };

// after a kernel launch: count it, surface launch errors, and (when tracing) record an event.
// `work` is the launch's ALGORITHMIC work: flops for contractions, bytes for streaming kernels.
int after_launch(af_ctx* ctx, const char* what, int kind, double work);

// Kernel launch that may carry the programmatic-dependent-launch attribute: the kernel may then start while its
// predecessor in the stream drains, and MUST execute griddepcontrol.wait before touching the predecessor's output
// (pdl_wait() below).  Only kernels written that way are launched with `pdl` = true.
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#ifdef __CUDACC__
// wait for the predecessor grid (no-op without the launch attribute), then let the successor start launching: it
// waits on us the same way, so triggering early is always safe
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ---- data parallelism (dp.cu) ----------------------------------------------------------------
// sum all-reduce of d_ptr[0..n) across the context's communicator, enqueued on `stream` (Gradient.Add across ranks,
// gradient.go:28-30,108-120).  No-op when the context has no communicator or a single rank.
int dp_allreduce(af_ctx* ctx, double* d_ptr, long long n, cudaStream_t stream, bool wide = false);
inline bool dp_active(const af_ctx* ctx) { return ctx->comm != nullptr && ctx->dp_nranks > 1; }

// ---- contraction front end (gemm.cu) --------------------------------------------------------
struct GemmOperand {
  const double* p0 = nullptr;  // primary (X, U, W)
  const double* p1 = nullptr;  // R companion or NULL
  int layout = 0;              // LAY_KC / LAY_RC
  long long ld = 0;
};
struct GemmEpilogue {
  int epi = 0;
  double* out0 = nullptr;
  double* out1 = nullptr;
  long long ldo = 0;
  const double* bias0 = nullptr;
  const double* bias1 = nullptr;
  const double* s = nullptr;
  const double* rs = nullptr;
  long long lds = 0;
  bool prezeroed = false;  // store-type outputs are already zero: a split / stream-K launch may skip its memsets
  // EPI_SIGMOID_BWD only: if the launch can, it also adds the column sums of out0 / out1 to colsum0 / colsum1 (the bias
  // gradient of the layer below, lin_add.go:94-104) and sets *colsum_done; otherwise the caller sums in a pass of its own
  double* colsum0 = nullptr;
  double* colsum1 = nullptr;
  bool* colsum_done = nullptr;
};
// acc0 = A0 B0^T ; acc1 = A0 B1^T + A1 B0^T (dual when `dual`), I x J, contraction Kd
int gemm_launch(af_ctx* ctx, const GemmOperand& A, const GemmOperand& B, bool dual, long long I, long long J,
                long long Kd, const GemmEpilogue& e);

// TMA descriptors (gemm.cu).  layout LAY_KC: element (r,k) at p[r*ld + k] -> box {16, tile_rows}; LAY_RC: element (r,k) at
// p[k*ld + r] -> box {16, 16}.  gemm_make_gate_map: the stacked LSTM gate matrix as the 4-D tensor of a forward-step tile.
int gemm_make_map(af_ctx* ctx, CUtensorMap* map, const double* p, int layout, long long rows, long long kd, long long ld,
                  int tile_rows);
int gemm_make_gate_map(af_ctx* ctx, CUtensorMap* map, const double* w4, long long H, long long C);
int gemm_make_plain_map(af_ctx* ctx, CUtensorMap* map, const double* p, long long rows, long long cols, long long ld, int box_rows,
                        int box_cols);

// small-batch (n <= 16) streaming kernels, small_n.cu: the reference's Gemv / Ger branch
bool small_n_applicable(af_ctx* ctx, long long rows, long long cols, long long n, std::initializer_list<const void*> ptrs,
                        int nmax = 8);
int small_fwd_launch(af_ctx* ctx, const double* W, const double* RW, const double* X, const double* RX, long long rows,
                     long long cols, long long n, const GemmEpilogue& e, bool dual);
int small_wgrad_launch(af_ctx* ctx, const double* U, const double* UR, const double* X, const double* RX, double* gW,
                       double* rgW, long long rows, long long cols, long long n);
int small_igrad_launch(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW, long long rows,
                       long long cols, long long n, const GemmEpilogue& e, bool dual);

// LinTran with 9..16 samples (few_samples.cu): streaming tensor-core passes over W.  The launchers return AF_ENOTSUP --
// nothing launched, no error text -- when a shape does not fit the scratch; callers then take their other path.
bool few_samples_applicable(af_ctx* ctx, long long rows, long long cols, long long n, std::initializer_list<const void*> ptrs);
int few_samples_fwd_launch(af_ctx* ctx, const double* W, const double* RW, const double* X, const double* RX, long long rows,
                           long long cols, long long n, const GemmEpilogue& e, bool dual);
// weight gradient (gW / rgW, either may be NULL) and input gradient (D / DR; both NULL = none) in one pass; S / RS: sigmoid
// backward of the layer below applied to the input gradient; accumulate_d: D += instead of D =
int few_samples_bwd_launch(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW, const double* X,
                           const double* RX, double* gW, double* rgW, double* D, double* DR, const double* S, const double* RS,
                           long long rows, long long cols, long long n, bool accumulate_d);

// weight gradient of a layer with at most 16 output rows and a large batch (few_samples.cu): deterministic two-level sum
bool few_rows_wgrad_applicable(af_ctx* ctx, long long rows, long long cols, long long n, std::initializer_list<const void*> ptrs);
int few_rows_wgrad_launch(af_ctx* ctx, const double* U, const double* UR, const double* X, const double* RX, double* gW, double* rgW,
                          long long rows, long long cols, long long n);

// input gradient through a layer with at most 16 output rows and a large batch (streaming, small_n.cu)
bool few_rows_igrad_applicable(af_ctx* ctx, long long rows, long long cols, long long n, std::initializer_list<const void*> ptrs);
int few_rows_igrad_launch(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW, long long rows,
                          long long cols, long long n, const GemmEpilogue& e, bool dual);

// forward through a layer with at most 16 output rows and a large batch (streaming, small_n.cu); with `head` the kernel
// also runs the top of the backward pass on its outputs: sigmoid backward on the upstream U / UR (n x rows, in place; or
// on outGrad = 1 / outRGrad = 0 when `unit`) and the bias gradient gb / rgb (either may be NULL)
struct FewRowsHead {
  bool unit = false;
  double* U = nullptr;
  double* UR = nullptr;
  double* gb = nullptr;
  double* rgb = nullptr;
};
bool few_rows_fwd_applicable(af_ctx* ctx, long long rows, long long cols, long long n, std::initializer_list<const void*> ptrs);
int few_rows_fwd_launch(af_ctx* ctx, const double* W, const double* RW, const double* b, const double* Rb, const double* X,
                        const double* RX, double* S, double* RS, int act, const FewRowsHead* head, long long rows, long long cols,
                        long long n, bool dual);

// forward through a layer with at most 32 input columns and a large batch (streaming, small_n.cu): Y = X W^T + b
bool few_cols_fwd_applicable(af_ctx* ctx, long long rows, long long cols, long long n);
int few_cols_fwd_launch(af_ctx* ctx, const double* X, long long ldx, const double* W, const double* RW, long long ldw,
                        const double* b, const double* Rb, double* Y, double* RY, long long ldy, long long rows, long long cols,
                        long long n);

// whole-step persistent kernel for an MLP plan with n <= 8 samples (small_mlp.cu)
bool small_mlp_applicable(af_ctx* ctx, int L, const int64_t* sizes, int64_t n);
int small_mlp_launch(af_ctx* ctx, int L, const int64_t* sizes, int64_t n, double* const* param, const double* const* rvec,
                     double* const* grad, double* const* rgrad, double* const* act, double* const* ract,
                     double* const* delta, double* const* rdelta, bool with_r, bool backprop, bool fresh,
                     bool unit_upstream);

}  // namespace af
