// C ABI of libautofunc_b200 (include/autofunc_b200.h): context, memory, LinTran entry points,
// fused dense layer, whole-MLP plan.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <vector>

#include "common.h"
#include "gemm_dmma.cuh"

namespace af {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return e == cudaErrorMemoryAllocation ? AF_ENOMEM : AF_ECUDA;
}
int after_launch(af_ctx* ctx, const char* what, int kind, double work) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  ctx->launches += 1;
  if (ctx->tracing && ctx->trace_used + 1 < ctx->trace_events.size()) {
    cudaEventRecord(ctx->trace_events[ctx->trace_used + 1], ctx->stream);
    ctx->trace_kind.push_back(kind);
    ctx->trace_work.push_back(work);
    ctx->trace_used += 1;
  }
  return AF_OK;
}

int colsum_launch(af_ctx* ctx, const double* U, const double* UR, double* gb, double* rgb, long long len, long long n);

}  // namespace af

using namespace af;

#define AF_CHECK_CTX(ctx) \
  if (!(ctx)) return af::fail(AF_EINVAL, "null context"); \
  af::DeviceGuard _guard((ctx)->device)
#define AF_NEED(p) \
  if (!(p)) return af::fail(AF_EINVAL, "null pointer argument: " #p)
// entry points that synchronise, allocate or manage graphs of their own cannot be part of a recorded sequence
#define AF_NOT_RECORDING(ctx, what) \
  if ((ctx)->recording) return af::fail(AF_EINVAL, what ": not allowed between af_graph_begin and af_graph_end")

static int check_dims(int64_t rows, int64_t cols, int64_t n) {
  if (rows < 0 || cols < 0 || n < 0) return fail(AF_ESHAPE, "negative dimension");
  return AF_OK;
}

extern "C" {

// ------------------------------------------------------------------------------------------
// context / memory
// ------------------------------------------------------------------------------------------
static int ctx_create_impl(int device, void* stream, bool own, af_ctx** out) {
  if (!out) return fail(AF_EINVAL, "null out pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (device < 0 || device >= count) return fail(AF_EINVAL, "no such CUDA device " + std::to_string(device));
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  AF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(AF_ENOTSUP, std::string("autofunc_b200 is built for sm_100a only; device is ") + prop.name +
                                " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
  af_ctx* c = new af_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (own) {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreateWithFlags"); }
    c->owns_stream = true;
  } else {
    c->stream = static_cast<cudaStream_t>(stream);
  }
  cudaDriverEntryPointQueryResult q;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &c->encode_tiled, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !c->encode_tiled) {
    if (own) cudaStreamDestroy(c->stream);
    delete c;
    return fail(AF_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  }
  c->scratch_doubles = kScratchDoubles;
  e = cudaMalloc(&c->scratch, c->scratch_doubles * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&c->fold_tickets, kFoldTickets * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(c->fold_tickets, 0, kFoldTickets * sizeof(unsigned));
  // workspace of the ordered stream-K fix-up (gemm.cu: fixup_workspace has the same sizes): allocated here, not at first use,
  // because a plan may re-capture its step graph right after an option change -- and cudaMalloc is not allowed while a
  // thread-local capture is open
  const size_t fix_ctas = 2 * (size_t)c->sm_count + 1;
  if (e == cudaSuccess) e = cudaMalloc(&c->sk_part, fix_ctas * 2 * 4 * 4 * 2 * GemmSmem<2, 2>::kThreads * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&c->sk_flags, fix_ctas * GemmSmem<2, 2>::kWarps * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(c->sk_flags, 0, fix_ctas * GemmSmem<2, 2>::kWarps * sizeof(unsigned));
  if (e != cudaSuccess) {
    if (c->scratch) cudaFree(c->scratch);
    if (c->fold_tickets) cudaFree(c->fold_tickets);
    if (c->sk_part) cudaFree(c->sk_part);
    if (c->sk_flags) cudaFree(c->sk_flags);
    if (own) cudaStreamDestroy(c->stream);
    delete c;
    return cuda_fail(e, "cudaMalloc(scratch)");
  }
  *out = c;
  return AF_OK;
}

int af_ctx_create(int device, af_ctx** out) { return ctx_create_impl(device, nullptr, true, out); }
int af_ctx_create_on_stream(int device, void* cuda_stream, af_ctx** out) {
  return ctx_create_impl(device, cuda_stream, false, out);
}

int af_ctx_destroy(af_ctx* ctx) {
  if (!ctx) return AF_OK;
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  af_dp_destroy(ctx);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->fold_tickets) cudaFree(ctx->fold_tickets);
  if (ctx->sk_part) cudaFree(ctx->sk_part);
  if (ctx->sk_flags) cudaFree(ctx->sk_flags);
  if (ctx->barrier_word) cudaFree(ctx->barrier_word);
  for (cudaEvent_t ev : ctx->trace_events) cudaEventDestroy(ev);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return AF_OK;
}

int af_sync(af_ctx* ctx) {
  AF_CHECK_CTX(ctx);
  AF_NOT_RECORDING(ctx, "af_sync");
  AF_CUDA(cudaStreamSynchronize(ctx->stream));
  return AF_OK;
}

// ---- recorded launch sequences (CUDA graphs) ---------------------------------------------------------
struct af_graph {
  af_ctx* ctx = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
};

int af_graph_begin(af_ctx* ctx) {
  AF_CHECK_CTX(ctx);
  if (ctx->recording) return fail(AF_EINVAL, "af_graph_begin: already recording");
  if (ctx->tracing) return fail(AF_EINVAL, "af_graph_begin: a launch trace is active");
  if (ctx->stream == nullptr || ctx->stream == cudaStreamLegacy)
    return fail(AF_ENOTSUP, "af_graph_begin: the legacy default stream cannot be captured");
  AF_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
  ctx->recording = true;
  ctx->record_start = ctx->launches;
  return AF_OK;
}

int af_graph_end(af_ctx* ctx, af_graph** out) {
  AF_CHECK_CTX(ctx);
  if (!ctx->recording) return fail(AF_EINVAL, "af_graph_end: not recording");
  ctx->recording = false;
  const int64_t launches = ctx->launches - ctx->record_start;
  ctx->launches = ctx->record_start;  // nothing ran
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
  if (ce != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    return cuda_fail(ce, "cudaStreamEndCapture");
  }
  if (!out) {  // the caller abandons the recording
    if (graph) cudaGraphDestroy(graph);
    return AF_OK;
  }
  af_graph* g = new af_graph();
  g->ctx = ctx;
  g->launches = launches;
  const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    delete g;
    return cuda_fail(ie, "cudaGraphInstantiate");
  }
  *out = g;
  return AF_OK;
}

int af_graph_launch(af_graph* g) {
  if (!g) return fail(AF_EINVAL, "null graph");
  af_ctx* ctx = g->ctx;
  DeviceGuard guard(ctx->device);
  AF_NOT_RECORDING(ctx, "af_graph_launch");
  if (ctx->tracing) return fail(AF_EINVAL, "af_graph_launch: a launch trace is active (replays carry no per-launch events)");
  AF_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
  ctx->launches += g->launches;
  return AF_OK;
}

int64_t af_graph_launches(af_graph* g) { return g ? g->launches : 0; }

int af_graph_destroy(af_graph* g) {
  if (!g) return AF_OK;
  DeviceGuard guard(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  if (g->exec) cudaGraphExecDestroy(g->exec);
  delete g;
  return AF_OK;
}

const char* af_last_error(void) { return g_last_error.c_str(); }
int af_version(void) { return 100; }
int64_t af_launch_count(af_ctx* ctx) { return ctx ? ctx->launches : 0; }

int af_trace_begin(af_ctx* ctx, int capacity) {
  AF_CHECK_CTX(ctx);
  AF_NOT_RECORDING(ctx, "af_trace_begin");
  if (capacity < 1) return fail(AF_EINVAL, "trace capacity must be positive");
  while ((int)ctx->trace_events.size() < capacity + 1) {
    cudaEvent_t ev;
    AF_CUDA(cudaEventCreate(&ev));
    ctx->trace_events.push_back(ev);
  }
  ctx->trace_kind.clear();
  ctx->trace_work.clear();
  ctx->trace_used = 0;
  AF_CUDA(cudaEventRecord(ctx->trace_events[0], ctx->stream));
  ctx->tracing = true;
  return AF_OK;
}

int af_trace_end(af_ctx* ctx, int capacity, int* h_kind, double* h_work, float* h_ms, int* h_count) {
  AF_CHECK_CTX(ctx); AF_NEED(h_count);
  ctx->tracing = false;
  AF_CUDA(cudaStreamSynchronize(ctx->stream));
  int n = (int)ctx->trace_used;
  if (n > capacity) n = capacity;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    AF_CUDA(cudaEventElapsedTime(&ms, ctx->trace_events[i], ctx->trace_events[i + 1]));
    if (h_kind) h_kind[i] = ctx->trace_kind[i];
    if (h_work) h_work[i] = ctx->trace_work[i];
    if (h_ms) h_ms[i] = ms;
  }
  *h_count = n;
  return AF_OK;
}

int af_set_option(af_ctx* ctx, const char* name, int value) {
  if (!ctx || !name) return fail(AF_EINVAL, "null argument");
  ctx->option_epoch += 1;  // plan-owned graphs captured under the old settings are re-captured
  if (!strcmp(name, "dp_max_ctas")) {
    if (value < 0 || value > 64) return fail(AF_EINVAL, "dp_max_ctas must lie in [0, 64]");
    if (ctx->comm) return fail(AF_EINVAL, "dp_max_ctas must be set before af_dp_init");
    ctx->dp_max_ctas = value;
    return AF_OK;
  }
  if (!strcmp(name, "dp_reserve_sms")) {
    if (value < -1 || value >= ctx->sm_count) return fail(AF_EINVAL, "dp_reserve_sms must lie in [-1, sm_count)");
    ctx->dp_reserve_sms = value;
    return AF_OK;
  }
  if (!strcmp(name, "big_cfg")) {
    if (value != 0 && value != 2) return fail(AF_EINVAL, "big_cfg must be 0 (128x64, 8 warps, 1 CTA/SM) or 2 (64x64, 4 warps, 2 CTAs/SM)");
    ctx->big_cfg = value;
    return AF_OK;
  }
  if (!strcmp(name, "persist_tiles")) {
    if (value < 0 || value > 2) return fail(AF_EINVAL, "persist_tiles must be 0, 1 or 2");
    if (value != 0 && !AF_PERSIST_TILES) return fail(AF_ENOTSUP, "persist_tiles: the library was built without -DAF_PERSIST_TILES=1 (the mode measured slower)");
    ctx->persist_tiles = value;
    return AF_OK;
  }
  if (!strcmp(name, "sk_bias")) {
    if (value < 50 || value > 200) return fail(AF_EINVAL, "sk_bias is a percentage between 50 and 200");
    ctx->sk_bias = ctx->sk_bias_long = value;
    return AF_OK;
  }
  if (!strcmp(name, "sk_rotate")) {
    if (value != 0 && value != 1) return fail(AF_EINVAL, "sk_rotate must be 0 or 1");
    ctx->sk_rotate = value;
    return AF_OK;
  }
  if (!strcmp(name, "small_wgrad_max")) {
    if (value < 0 || value > 16) return fail(AF_EINVAL, "small_wgrad_max must lie in [0, 16]");
    ctx->small_wgrad_max = value;
    return AF_OK;
  }
  if (!strcmp(name, "small_n")) {
    ctx->small_n_off = value == 0;
    return AF_OK;
  }
  if (!strcmp(name, "short_rows")) {
    ctx->short_rows_off = value == 0;
    return AF_OK;
  }
  if (!strcmp(name, "ops_runtime")) {
    ctx->ops_runtime = value != 0;
    return AF_OK;
  }
  if (!strcmp(name, "sk_fix_bias")) {
    if (value < 50 || value > 200) return fail(AF_EINVAL, "sk_fix_bias is a percentage between 50 and 200");
    ctx->sk_fix_bias = value;
    return AF_OK;
  }
  if (!strcmp(name, "sk_fixup")) {
    if (value < 0 || value > 2) return fail(AF_EINVAL, "sk_fixup must be 0, 1 or 2");
    ctx->sk_fixup = value;
    return AF_OK;
  }
  if (!strcmp(name, "fused_colsum")) {
    ctx->fused_colsum_off = value == 0;
    return AF_OK;
  }
  if (!strcmp(name, "lstm_split")) {
    ctx->lstm_split_off = value == 0;
    return AF_OK;
  }
  if (!strcmp(name, "lstm_fused")) {
    ctx->lstm_fused_off = value == 0;
    return AF_OK;
  }
  if (!strcmp(name, "pdl")) {  // 1: programmatic dependent launch in the LSTM time loops' dependent chains; 0 (default): plain launches
    ctx->pdl_off = value == 0;
    ctx->pdl_fused_off = value == 0;
    return AF_OK;
  }
  if (!strcmp(name, "small_mlp_head")) {
    ctx->small_mlp_head_off = value == 0;
    return AF_OK;
  }
  if (!strcmp(name, "small_mlp")) {
    ctx->small_mlp_off = value == 0;
    return AF_OK;
  }
  if (!strcmp(name, "few_samples")) {
    ctx->few_samples_off = value == 0;
    return AF_OK;
  }
  if (!strcmp(name, "few_rows_wgrad")) {
    if (value < 0 || value > 2) return fail(AF_EINVAL, "few_rows_wgrad must be 0, 1 or 2");
    ctx->few_rows_wgrad = value;
    return AF_OK;
  }
  if (!strcmp(name, "few_samples_blocks")) {
    if (value < 1 || value > 16) return fail(AF_EINVAL, "few_samples_blocks must lie in [1, 16]");
    ctx->few_samples_blocks = value;
    return AF_OK;
  }
  if (!strcmp(name, "few_samples_tma")) {
    ctx->few_samples_tma = value != 0;
    return AF_OK;
  }
  if (!strcmp(name, "few_samples_min")) {
    if (value < 1 || value > 17) return fail(AF_EINVAL, "few_samples_min must lie in [1, 17]");
    ctx->few_samples_min = value;
    return AF_OK;
  }
  if (!strcmp(name, "deterministic")) {
    if (value != 0 && value != 1) return fail(AF_EINVAL, "deterministic must be 0 or 1");
    if (ctx->recording) return fail(AF_EINVAL, "deterministic cannot change while a graph is being recorded");
    ctx->deterministic = value == 1;
    return AF_OK;
  }
  return fail(AF_EINVAL, std::string("unknown option ") + name);
}

int af_set_gemm_path(af_ctx* ctx, int mode) {
  if (!ctx) return fail(AF_EINVAL, "null context");
  if (mode < 0 || mode > 2) return fail(AF_EINVAL, "gemm path must be 0, 1 or 2");
  ctx->option_epoch += 1;
  ctx->gemm_path = mode;
  return AF_OK;
}

int af_malloc(af_ctx* ctx, int64_t n, double** out) {
  AF_CHECK_CTX(ctx); AF_NEED(out);
  if (n < 0) return fail(AF_ESHAPE, "negative allocation size");
  *out = nullptr;
  AF_NOT_RECORDING(ctx, "af_malloc");
  AF_CUDA(cudaMalloc(out, (size_t)(n > 0 ? n : 1) * sizeof(double)));
  return AF_OK;
}

int af_free(af_ctx* ctx, double* p) {
  AF_CHECK_CTX(ctx);
  if (!p) return AF_OK;
  AF_NOT_RECORDING(ctx, "af_free");
  AF_CUDA(cudaStreamSynchronize(ctx->stream));
  AF_CUDA(cudaFree(p));
  return AF_OK;
}

int af_upload(af_ctx* ctx, double* dst, const double* src, int64_t n) {
  AF_CHECK_CTX(ctx);
  if (n <= 0) return AF_OK;
  AF_NEED(dst); AF_NEED(src);
  AF_NOT_RECORDING(ctx, "af_upload");
  AF_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  AF_CUDA(cudaStreamSynchronize(ctx->stream));
  return AF_OK;
}

int af_download(af_ctx* ctx, double* dst, const double* src, int64_t n) {
  AF_CHECK_CTX(ctx);
  if (n <= 0) return AF_OK;
  AF_NEED(dst); AF_NEED(src);
  AF_NOT_RECORDING(ctx, "af_download");
  AF_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  AF_CUDA(cudaStreamSynchronize(ctx->stream));
  return AF_OK;
}

int af_zero(af_ctx* ctx, double* p, int64_t n) {
  AF_CHECK_CTX(ctx);
  if (n <= 0) return AF_OK;
  AF_NEED(p);
  AF_CUDA(cudaMemsetAsync(p, 0, (size_t)n * 8, ctx->stream));
  return AF_OK;
}

int af_copy(af_ctx* ctx, double* dst, const double* src, int64_t n) {
  AF_CHECK_CTX(ctx);
  if (n <= 0) return AF_OK;
  AF_NEED(dst); AF_NEED(src);
  AF_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  return AF_OK;
}

// ------------------------------------------------------------------------------------------
// LinTran
// ------------------------------------------------------------------------------------------
static int dense_fwd_impl(af_ctx* ctx, const double* W, const double* RW, const double* b, const double* Rb,
                          const double* X, const double* RX, double* Y, double* RY, int64_t rows, int64_t cols,
                          int64_t n, int epi, bool dual) {
  GemmOperand A{X, dual ? RX : nullptr, LAY_KC, cols};
  GemmOperand B{W, dual ? RW : nullptr, LAY_KC, cols};
  GemmEpilogue e;
  e.epi = epi; e.out0 = Y; e.out1 = RY; e.ldo = rows; e.bias0 = b; e.bias1 = Rb;
  // 9..16 samples (C1 is n = 10): one streaming tensor-core pass over W, K split over blocks and folded in order
  if ((epi == EPI_STORE || epi == EPI_BIAS || epi == EPI_BIAS_SIGMOID) && few_samples_applicable(ctx, rows, cols, n, {W, RW, X, RX})) {
    const int rc = few_samples_fwd_launch(ctx, W, RW, X, RX, rows, cols, n, e, dual);
    if (rc != AF_ENOTSUP) return rc;
  }
  if (small_n_applicable(ctx, rows, cols, n, {W, RW, X, RX}))  // the reference's Gemv branch, lin_tran.go:90-99,138-152
    return small_fwd_launch(ctx, W, RW, X, RX, rows, cols, n, e, dual);
  // at most 16 output rows and a large batch (a 10-class head): one streaming pass, not 64 x 16 tensor-core tiles
  if ((epi == EPI_STORE || epi == EPI_BIAS || epi == EPI_BIAS_SIGMOID) &&
      few_rows_fwd_applicable(ctx, rows, cols, n, {W, RW, X, RX, Y, RY}))
    return few_rows_fwd_launch(ctx, W, dual ? RW : nullptr, b, dual ? Rb : nullptr, X, dual ? RX : nullptr, Y, dual ? RY : nullptr,
                               epi == EPI_BIAS_SIGMOID, nullptr, rows, cols, n, dual);
  // at most 32 input columns, a large batch and no R-input (a first layer on narrow features; the LSTM's x projection):
  // a register-tiled streaming pass bound by the output stream, not two-k-tile tensor-core tiles
  if ((epi == EPI_STORE || epi == EPI_BIAS) && !(dual && RX) && Y && (!dual || RY) && few_cols_fwd_applicable(ctx, rows, cols, n))
    return few_cols_fwd_launch(ctx, X, cols, W, dual ? RW : nullptr, cols, b, dual ? Rb : nullptr, Y, dual ? RY : nullptr, rows, rows,
                               cols, n);
  if (cols == 0) {
    // empty contraction: fall through the generic kernel (Kd == 0 gives acc = 0)
    int saved = ctx->gemm_path;
    ctx->gemm_path = 1;
    int rc = gemm_launch(ctx, A, B, dual, n, rows, cols, e);
    ctx->gemm_path = saved;
    return rc;
  }
  return gemm_launch(ctx, A, B, dual, n, rows, cols, e);
}

int af_lintran_fwd(af_ctx* ctx, const double* W, const double* X, double* Y, int64_t rows, int64_t cols, int64_t n) {
  AF_CHECK_CTX(ctx); AF_TRY(check_dims(rows, cols, n));
  if (rows * n == 0) return AF_OK;
  AF_NEED(W); AF_NEED(X); AF_NEED(Y);
  return dense_fwd_impl(ctx, W, nullptr, nullptr, nullptr, X, nullptr, Y, nullptr, rows, cols, n, EPI_STORE, false);
}

int af_lintran_fwd_r(af_ctx* ctx, const double* W, const double* RW, const double* X, const double* RX, double* Y,
                     double* RY, int64_t rows, int64_t cols, int64_t n) {
  AF_CHECK_CTX(ctx); AF_TRY(check_dims(rows, cols, n));
  if (rows * n == 0) return AF_OK;
  AF_NEED(W); AF_NEED(X); AF_NEED(RY);
  return dense_fwd_impl(ctx, W, RW, nullptr, nullptr, X, RX, Y, RY, rows, cols, n, EPI_STORE, true);
}

static int wgrad_impl(af_ctx* ctx, const double* U, const double* UR, const double* X, const double* RX, double* gW,
                      double* rgW, int64_t rows, int64_t cols, int64_t n) {
  if (rows * cols == 0 || n == 0) return AF_OK;
  const bool dual = rgW != nullptr;
  if (!gW && !rgW) return AF_OK;
  if (few_samples_applicable(ctx, rows, cols, n, {gW, rgW})) {  // 9..16 samples: the rank-n update as one streaming pass
    const int rc = few_samples_bwd_launch(ctx, U, UR, nullptr, nullptr, X, RX, gW, rgW, nullptr, nullptr, nullptr, nullptr, rows, cols, n, false);
    if (rc != AF_ENOTSUP) return rc;
  }
  // <= 16 output rows and a large batch (the 10-class head): the streaming kernel with ordered sums
  if ((ctx->few_rows_wgrad == 2 || (ctx->few_rows_wgrad == 1 && ctx->deterministic)) &&
      few_rows_wgrad_applicable(ctx, rows, cols, n, {X, RX, gW, rgW}))
    return few_rows_wgrad_launch(ctx, U, UR, X, RX, gW, rgW, rows, cols, n);
  if (small_n_applicable(ctx, rows, cols, n, {X, RX, gW, rgW}, ctx->small_wgrad_max))  // Ger branch, lin_tran.go:187-196,223-241
    return small_wgrad_launch(ctx, U, dual ? UR : nullptr, X, dual ? RX : nullptr, gW, rgW, rows, cols, n);
  GemmOperand A{U, dual ? UR : nullptr, LAY_RC, rows};
  GemmOperand B{X, dual ? RX : nullptr, LAY_RC, cols};
  GemmEpilogue e;
  e.epi = EPI_ACCUM; e.out0 = gW; e.out1 = rgW; e.ldo = cols;
  return gemm_launch(ctx, A, B, dual, rows, cols, n, e);
}

int af_lintran_wgrad(af_ctx* ctx, const double* U, const double* X, double* gW, int64_t rows, int64_t cols, int64_t n) {
  AF_CHECK_CTX(ctx); AF_TRY(check_dims(rows, cols, n));
  if (rows * cols == 0 || n == 0) return AF_OK;
  AF_NEED(U); AF_NEED(X); AF_NEED(gW);
  return wgrad_impl(ctx, U, nullptr, X, nullptr, gW, nullptr, rows, cols, n);
}

int af_lintran_wgrad_r(af_ctx* ctx, const double* U, const double* UR, const double* X, const double* RX, double* gW,
                       double* rgW, int64_t rows, int64_t cols, int64_t n) {
  AF_CHECK_CTX(ctx); AF_TRY(check_dims(rows, cols, n));
  if (rows * cols == 0 || n == 0) return AF_OK;
  AF_NEED(U); AF_NEED(X);
  return wgrad_impl(ctx, U, UR, X, RX, gW, rgW, rows, cols, n);
}

// colsum0 / colsum1 / colsum_done (optional, with S): the kernels that can also add the column sums of D / DR -- the bias
// gradient of the layer below, lin_add.go:94-104 -- to colsum0 / colsum1 and then set *colsum_done
static int igrad_impl(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW,
                      const double* S, const double* RS, double* D, double* DR, int64_t rows, int64_t cols, int64_t n,
                      double* colsum0 = nullptr, double* colsum1 = nullptr, bool* colsum_done = nullptr) {
  if (cols * n == 0) return AF_OK;
  const bool dual = DR != nullptr;
  GemmOperand A{U, dual ? UR : nullptr, LAY_KC, rows};
  GemmOperand B{W, dual ? RW : nullptr, LAY_RC, cols};
  GemmEpilogue e;
  e.epi = S ? EPI_SIGMOID_BWD : EPI_STORE;
  e.out0 = D; e.out1 = DR; e.ldo = cols; e.s = S; e.rs = RS; e.lds = cols;
  if (S && colsum_done && !ctx->fused_colsum_off && !ctx->deterministic) {
    e.colsum0 = colsum0; e.colsum1 = dual ? colsum1 : nullptr; e.colsum_done = colsum_done;
  }
  if (rows > 0 && few_samples_applicable(ctx, rows, cols, n, {D, DR, S, RS})) {  // 9..16 samples: one streaming pass over W
    const int rc = few_samples_bwd_launch(ctx, U, UR, W, RW, nullptr, nullptr, nullptr, nullptr, D, DR, S, RS, rows, cols, n, false);
    if (rc != AF_ENOTSUP) return rc;
  }
  // Gemv T branch, lin_tran.go:285-288,329-334 (its row chunks meet in red.global.add.f64: not in deterministic mode)
  if (rows > 0 && !ctx->deterministic && small_n_applicable(ctx, rows, cols, n, {W, RW, D, DR}))
    return small_igrad_launch(ctx, U, UR, W, RW, rows, cols, n, e, dual);
  if (few_rows_igrad_applicable(ctx, rows, cols, n, {W, RW, D, DR, S, RS}))  // <= 16 output rows: a streaming pass, not a GEMM
    return few_rows_igrad_launch(ctx, U, UR, W, RW, rows, cols, n, e, dual);
  if (rows == 0) {
    int saved = ctx->gemm_path;
    ctx->gemm_path = 1;
    int rc = gemm_launch(ctx, A, B, dual, n, cols, rows, e);
    ctx->gemm_path = saved;
    return rc;
  }
  return gemm_launch(ctx, A, B, dual, n, cols, rows, e);
}

int af_lintran_igrad(af_ctx* ctx, const double* U, const double* W, double* D, int64_t rows, int64_t cols, int64_t n) {
  AF_CHECK_CTX(ctx); AF_TRY(check_dims(rows, cols, n));
  if (cols * n == 0) return AF_OK;
  AF_NEED(U); AF_NEED(W); AF_NEED(D);
  return igrad_impl(ctx, U, nullptr, W, nullptr, nullptr, nullptr, D, nullptr, rows, cols, n);
}

int af_lintran_igrad_r(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW, double* D,
                       double* DR, int64_t rows, int64_t cols, int64_t n) {
  AF_CHECK_CTX(ctx); AF_TRY(check_dims(rows, cols, n));
  if (cols * n == 0) return AF_OK;
  AF_NEED(U); AF_NEED(W); AF_NEED(DR);
  return igrad_impl(ctx, U, UR, W, RW, nullptr, nullptr, D, DR, rows, cols, n);
}

// linTranRResult.PropagateRGradient (lin_tran.go:393-425) as one call: dataGradient(R) and inputGradient(R)
static int backward_impl(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW, const double* X,
                         const double* RX, double* gW, double* rgW, double* D, double* DR, int64_t rows, int64_t cols, int64_t n,
                         bool accumulate_input) {
  const bool do_w = (gW || rgW) && rows * cols != 0 && n != 0, do_i = (D || DR) && cols * n != 0;
  if (do_w && do_i && rows > 0 && few_samples_applicable(ctx, rows, cols, n, {gW, rgW, D, DR})) {  // one pass over W for both
    const int rc = few_samples_bwd_launch(ctx, U, UR, W, RW, X, RX, gW, rgW, D, DR, nullptr, nullptr, rows, cols, n, accumulate_input);
    if (rc != AF_ENOTSUP) return rc;
  }
  if (do_w) AF_TRY(wgrad_impl(ctx, U, UR, X, RX, gW, rgW, rows, cols, n));
  if (!do_i) return AF_OK;
  if (!accumulate_input) return igrad_impl(ctx, U, UR, W, RW, nullptr, nullptr, D, DR, rows, cols, n);
  if (rows == 0) return AF_OK;  // D += 0
  if (few_samples_applicable(ctx, rows, cols, n, {D, DR})) {
    const int rc = few_samples_bwd_launch(ctx, U, UR, W, RW, nullptr, nullptr, nullptr, nullptr, D, DR, nullptr, nullptr, rows, cols, n, true);
    if (rc != AF_ENOTSUP) return rc;
  }
  const bool dual = DR != nullptr;
  GemmOperand A{U, dual ? UR : nullptr, LAY_KC, rows};
  GemmOperand B{W, dual ? RW : nullptr, LAY_RC, cols};
  GemmEpilogue e;
  e.epi = EPI_ACCUM; e.out0 = D; e.out1 = DR; e.ldo = cols;
  return gemm_launch(ctx, A, B, dual, n, cols, rows, e);
}

int af_lintran_backward(af_ctx* ctx, const double* U, const double* W, const double* X, double* gW, double* D, int64_t rows,
                        int64_t cols, int64_t n, int accumulate_input) {
  AF_CHECK_CTX(ctx); AF_TRY(check_dims(rows, cols, n));
  AF_NEED(U);
  if (gW) AF_NEED(X);
  if (D) AF_NEED(W);
  return backward_impl(ctx, U, nullptr, W, nullptr, X, nullptr, gW, nullptr, D, nullptr, rows, cols, n, accumulate_input != 0);
}

int af_lintran_backward_r(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW, const double* X,
                          const double* RX, double* gW, double* rgW, double* D, double* DR, int64_t rows, int64_t cols, int64_t n,
                          int accumulate_input) {
  AF_CHECK_CTX(ctx); AF_TRY(check_dims(rows, cols, n));
  AF_NEED(U);
  if (gW || rgW) AF_NEED(X);
  if (D || DR) AF_NEED(W);
  if (D && !DR) return fail(AF_EINVAL, "af_lintran_backward_r: d_DR is required with d_D (use af_lintran_backward for the plain form)");
  return backward_impl(ctx, U, UR, W, RW, X, RX, gW, rgW, D, DR, rows, cols, n, accumulate_input != 0);
}

// ------------------------------------------------------------------------------------------
// fused dense layer
// ------------------------------------------------------------------------------------------
int af_dense_fwd(af_ctx* ctx, const double* W, const double* RW, const double* b, const double* Rb, const double* X,
                 const double* RX, double* S, double* RS, int64_t rows, int64_t cols, int64_t n, int activation) {
  AF_CHECK_CTX(ctx); AF_TRY(check_dims(rows, cols, n));
  if (activation != 0 && activation != 1) return fail(AF_EINVAL, "activation must be 0 (none) or 1 (sigmoid)");
  if (rows * n == 0) return AF_OK;
  AF_NEED(W); AF_NEED(X); AF_NEED(S);
  return dense_fwd_impl(ctx, W, RW, b, Rb, X, RX, S, RS, rows, cols, n, activation ? EPI_BIAS_SIGMOID : EPI_BIAS,
                        RS != nullptr);
}

int af_dense_igrad(af_ctx* ctx, const double* U, const double* UR, const double* W, const double* RW,
                   const double* Sprev, const double* RSprev, double* D, double* DR, int64_t rows, int64_t cols,
                   int64_t n) {
  AF_CHECK_CTX(ctx); AF_TRY(check_dims(rows, cols, n));
  if (cols * n == 0) return AF_OK;
  AF_NEED(U); AF_NEED(W); AF_NEED(D);
  return igrad_impl(ctx, U, UR, W, RW, Sprev, RSprev, D, DR, rows, cols, n);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// whole-MLP plan (bench/mlp.go:24-126 batched)
// ------------------------------------------------------------------------------------------
struct af_mlp {
  af_ctx* ctx = nullptr;
  int L = 0;
  int64_t n = 0;
  std::vector<int64_t> sizes;
  double* slab = nullptr;  // one allocation; everything below points into it
  std::vector<double*> param, rvec, grad, rgrad;  // 2L each: W0,b0,W1,b1,...
  std::vector<bool> has_rvec;
  std::vector<double*> act, ract;      // L+1: act[0] = input
  std::vector<double*> delta, rdelta;  // L+1: delta[L] = upstream buffer
  int64_t total_params = 0;
  int64_t gr_doubles = 0;  // length of the contiguous {grad | rgrad} block (padded sub-buffers, layer blocks adjacent)
  int64_t param_len(int idx) const {
    const int l = idx / 2;
    return (idx & 1) ? sizes[l + 1] : sizes[l] * sizes[l + 1];
  }
  // host-buffer pipeline (af_mlp_submit_host / af_mlp_wait): slot 1 duplicates of the buffers a step
  // reads from or writes for the host, so step k+1 can compute while step k downloads and k+2 uploads
  struct Slot {
    double* x = nullptr;                 // input batch
    std::vector<double*> grad, rgrad;    // accumulators
    double *out = nullptr, *rout = nullptr;
    cudaEvent_t in_ready = nullptr, compute_done = nullptr, out_done = nullptr;
    bool busy = false;
  };
  Slot slot[2];
  int bound_slot = 0;
  // CUDA-graph replay of the step (small batches are launch-bound: ~25 short kernels per step).  One graph
  // per (with_r, backprop, slot); captured on the second call of a configuration, dropped whenever the set
  // of present R-vectors changes or an option of the context changes (epoch).
  struct StepGraph { cudaGraphExec_t exec = nullptr; int seen = 0; int64_t launches = 0; int64_t epoch = 0; };
  StepGraph graphs[2][2][2];
  bool use_graph = false;
  // data-parallel hook: called (on the host, at enqueue time) as soon as the launches that complete a range of
  // accumulators are on the stream, so the caller can start that range's all-reduce while the rest of the
  // backward pass runs.  Layer 1's weight gradient -- the largest and the last -- is produced in row chunks.
  af_grad_ready_fn grad_ready = nullptr;
  void* grad_ready_user = nullptr;
  int grad_ready_chunks = 1;
  double* alt = nullptr;  // backing store of slot 1
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  int64_t submitted = 0, waited = 0;
  // data parallelism inside the library (dp.cu): mode AF_DP_*, per-layer "block complete" events on the ctx stream,
  // per-slot "reduced" events on the communicator's stream
  int dp_mode = AF_DP_OFF;
  bool host_shard = false;
  std::vector<cudaEvent_t> dp_layer_ev, dp_fence_ev;
  cudaEvent_t dp_done[2] = {nullptr, nullptr};
  bool dp_pending[2] = {false, false};
};
static void mlp_drop_graphs(af_mlp* m) {
  for (auto& a : m->graphs)
    for (auto& b : a)
      for (auto& g : b) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = af_mlp::StepGraph();
      }
}


extern "C" {

int af_mlp_create(af_ctx* ctx, const int64_t* sizes, int n_layers, int64_t batch, af_mlp** out) {
  AF_CHECK_CTX(ctx); AF_NEED(sizes); AF_NEED(out);
  AF_NOT_RECORDING(ctx, "af_mlp_create");
  if (n_layers < 1 || batch < 1) return fail(AF_ESHAPE, "need at least one layer and one sample");
  for (int i = 0; i <= n_layers; ++i)
    if (sizes[i] < 1) return fail(AF_ESHAPE, "layer widths must be positive");
  af_mlp* m = new af_mlp();
  m->ctx = ctx; m->L = n_layers; m->n = batch;
  m->sizes.assign(sizes, sizes + n_layers + 1);
  const int P = 2 * n_layers;
  // every sub-buffer starts on a 256-byte boundary so TMA and vector accesses apply
  auto pad = [](int64_t x) { return (x + 31) / 32 * 32; };
  int64_t total = 0;
  std::vector<int64_t> off;
  auto take = [&](int64_t len) { off.push_back(total); total += pad(len); };
  for (int rep = 0; rep < 2; ++rep)
    for (int i = 0; i < P; ++i) take(m->param_len(i));  // param, rvec
  // accumulators, one contiguous block per layer: gW_l, gb_l, rgW_l, rgb_l (a layer's all-reduce is one range)
  for (int l = 0; l < n_layers; ++l)
    for (int rep = 0; rep < 2; ++rep)
      for (int j = 0; j < 2; ++j) take(m->param_len(2 * l + j));
  for (int rep = 0; rep < 2; ++rep)
    for (int l = 0; l <= n_layers; ++l) take(batch * sizes[l]);  // act, ract
  for (int rep = 0; rep < 2; ++rep)
    for (int l = 0; l <= n_layers; ++l) take(l == 0 ? 0 : batch * sizes[l]);  // delta, rdelta
  cudaError_t e = cudaMalloc(&m->slab, (size_t)total * 8);
  if (e != cudaSuccess) { delete m; return cuda_fail(e, "cudaMalloc(mlp slab)"); }
  e = cudaMemsetAsync(m->slab, 0, (size_t)total * 8, ctx->stream);
  if (e != cudaSuccess) { cudaFree(m->slab); delete m; return cuda_fail(e, "cudaMemsetAsync(mlp slab)"); }
  size_t k = 0;
  auto next = [&]() { return m->slab + off[k++]; };
  for (int i = 0; i < P; ++i) m->param.push_back(next());
  for (int i = 0; i < P; ++i) m->rvec.push_back(next());
  m->grad.resize(P); m->rgrad.resize(P);
  for (int l = 0; l < n_layers; ++l) {
    m->grad[2 * l] = next(); m->grad[2 * l + 1] = next();
    m->rgrad[2 * l] = next(); m->rgrad[2 * l + 1] = next();
  }
  for (int l = 0; l <= n_layers; ++l) m->act.push_back(next());
  for (int l = 0; l <= n_layers; ++l) m->ract.push_back(next());
  for (int l = 0; l <= n_layers; ++l) m->delta.push_back(next());
  for (int l = 0; l <= n_layers; ++l) m->rdelta.push_back(next());
  m->gr_doubles = m->act[0] - m->grad[0];
  m->has_rvec.assign(P, false);
  m->use_graph = batch <= 256;  // launch-bound regime
  for (int i = 0; i < P; ++i) m->total_params += m->param_len(i);
  *out = m;
  return AF_OK;
}

int af_mlp_destroy(af_mlp* m) {
  if (!m) return AF_OK;
  DeviceGuard guard(m->ctx->device);
  if (m->ctx->dp_stream) cudaStreamSynchronize(m->ctx->dp_stream);
  cudaStreamSynchronize(m->ctx->stream);
  if (m->copy_in) { cudaStreamSynchronize(m->copy_in); cudaStreamDestroy(m->copy_in); }
  if (m->copy_out) { cudaStreamSynchronize(m->copy_out); cudaStreamDestroy(m->copy_out); }
  for (auto& sl : m->slot)
    for (cudaEvent_t ev : {sl.in_ready, sl.compute_done, sl.out_done})
      if (ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : m->dp_layer_ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : m->dp_fence_ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : m->dp_done)
    if (ev) cudaEventDestroy(ev);
  mlp_drop_graphs(m);
  if (m->alt) cudaFree(m->alt);
  cudaFree(m->slab);
  delete m;
  return AF_OK;
}

static int mlp_index_ok(af_mlp* m, int index, int64_t len) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  if (index < 0 || index >= 2 * m->L) return fail(AF_EINVAL, "parameter index out of range");
  if (len != m->param_len(index))
    return fail(AF_ESHAPE, "parameter length should be " + std::to_string(m->param_len(index)) + " but got " +
                               std::to_string(len));
  return AF_OK;
}

int af_mlp_set_param(af_mlp* m, int index, const double* h, int64_t len) {
  AF_TRY(mlp_index_ok(m, index, len)); AF_NEED(h);
  return af_upload(m->ctx, m->param[index], h, len);
}

int af_mlp_set_rvector(af_mlp* m, int index, const double* h, int64_t len) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  if (index < 0 || index >= 2 * m->L) return fail(AF_EINVAL, "parameter index out of range");
  if (!h) {
    if (m->has_rvec[index]) mlp_drop_graphs(m);
    m->has_rvec[index] = false;
    return AF_OK;
  }
  AF_TRY(mlp_index_ok(m, index, len));
  if (!m->has_rvec[index]) mlp_drop_graphs(m);
  m->has_rvec[index] = true;
  return af_upload(m->ctx, m->rvec[index], h, len);
}

#define MLP_PTR(name, expr)                      \
  double* name(af_mlp* m, int index) {           \
    if (!m || index < 0 || index >= 2 * m->L) return nullptr; \
    return expr;                                 \
  }
MLP_PTR(af_mlp_param_ptr, m->param[index])
MLP_PTR(af_mlp_rvector_ptr, ((m->has_rvec[index] ? (void)0 : mlp_drop_graphs(m)), m->has_rvec[index] = true, m->rvec[index]))
MLP_PTR(af_mlp_grad_ptr, m->grad[index])
MLP_PTR(af_mlp_rgrad_ptr, m->rgrad[index])
#undef MLP_PTR
double* af_mlp_input_ptr(af_mlp* m) { return m ? m->act[0] : nullptr; }
double* af_mlp_output_ptr(af_mlp* m) { return m ? m->act[m->L] : nullptr; }
double* af_mlp_routput_ptr(af_mlp* m) { return m ? m->ract[m->L] : nullptr; }
double* af_mlp_upstream_ptr(af_mlp* m) { return m ? m->delta[m->L] : nullptr; }
double* af_mlp_upstream_r_ptr(af_mlp* m) { return m ? m->rdelta[m->L] : nullptr; }
int64_t af_mlp_grad_slab(af_mlp* m, double** d_first) {
  if (!m) return 0;
  if (d_first) *d_first = m->grad[0];
  return m->gr_doubles;
}

// the ctx stream waits for the slot's outstanding reduces (AF_DP_DEFERRED leaves them running past the step)
static int mlp_dp_join_slot(af_mlp* m, int k) {
  if (!m->dp_pending[k]) return AF_OK;
  AF_CUDA(cudaStreamWaitEvent(m->ctx->stream, m->dp_done[k], 0));
  m->dp_pending[k] = false;
  return AF_OK;
}

int af_mlp_dp_join(af_mlp* m) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  DeviceGuard guard(m->ctx->device);
  AF_TRY(mlp_dp_join_slot(m, 0));
  return mlp_dp_join_slot(m, 1);
}

int af_mlp_zero_grads(af_mlp* m) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  DeviceGuard guard(m->ctx->device);
  AF_TRY(mlp_dp_join_slot(m, m->bound_slot));
  // grad and rgrad blocks are adjacent in the slab
  AF_CUDA(cudaMemsetAsync(m->grad[0], 0, (size_t)m->gr_doubles * 8, m->ctx->stream));
  return AF_OK;
}

// Gradient.AddToVars (gradient.go:40-52) for the plan's own parameters: param_i += scale * grad_i, in HBM.
int af_mlp_add_to_vars(af_mlp* m, double scale) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  AF_TRY(mlp_dp_join_slot(m, m->bound_slot));
  for (int i = 0; i < 2 * m->L; ++i) AF_TRY(af_axpy(m->ctx, scale, m->grad[i], m->param[i], m->param_len(i)));
  return AF_OK;
}

int af_mlp_get_param(af_mlp* m, int index, double* h, int64_t len) {
  AF_TRY(mlp_index_ok(m, index, len)); AF_NEED(h);
  return af_download(m->ctx, h, m->param[index], len);
}

// Head fusion: when the last layer has at most 16 outputs (the bench network's 10-class head) and the step continues into
// the backward pass, its forward kernel also applies the top sigmoid backward to the upstream (or generates the unit
// upstream of bench/mlp.go:118-126) and, unless `head_bias` is false, accumulates the head's bias gradient
// (small_n.cu: few_rows_fwd_kernel).  The backward pass is then told to skip those launches.
struct MlpHead {
  bool backprop = false;   // in: the caller will run the backward pass right after this forward pass
  bool unit = false;       // in: upstream is outGrad = 1 / outRGrad = 0, not yet materialised
  bool bias = true;        // in: accumulators are ready to receive the head's bias gradient during the forward pass
  bool done = false;       // out: the head kernel ran the top sigmoid backward
  bool bias_done = false;  // out: ... and the head's bias gradient
};
static bool mlp_head_fusable(af_mlp* m, int with_r) {
  const int L = m->L;
  const bool R = with_r != 0;
  // (the head kernel adds per-block bias-gradient partial sums with atomics: not in deterministic mode)
  return !m->grad_ready && !m->ctx->deterministic &&
         few_rows_fwd_applicable(m->ctx, m->sizes[L], m->sizes[L - 1], m->n,
                                 {m->param[2 * L - 2], R ? m->rvec[2 * L - 2] : nullptr, m->act[L - 1], R ? m->ract[L - 1] : nullptr,
                                  m->act[L], R ? m->ract[L] : nullptr});
}
static int mlp_forward_launches(af_mlp* m, int with_r, MlpHead* head = nullptr);
static int mlp_backward_launches(af_mlp* m, int with_r, bool dp_overlap, const MlpHead* head = nullptr);
static int mlp_step_launches(af_mlp* m, int with_r, int backprop) {
  MlpHead head;
  head.backprop = backprop != 0;
  AF_TRY(mlp_forward_launches(m, with_r, &head));
  return backprop ? mlp_backward_launches(m, with_r, false, &head) : AF_OK;
}

static int mlp_pipeline_init(af_mlp* m);
static void mlp_bind_slot(af_mlp* m, int k);

int af_mlp_set_grad_ready(af_mlp* m, af_grad_ready_fn fn, void* user, int layer1_chunks) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  m->grad_ready = fn;
  m->grad_ready_user = user;
  m->grad_ready_chunks = layer1_chunks < 1 ? 1 : layer1_chunks;
  return AF_OK;
}

int af_mlp_set_dp(af_mlp* m, int mode) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  if (mode < AF_DP_OFF || mode > AF_DP_DEFERRED) return fail(AF_EINVAL, "dp mode must be AF_DP_OFF .. AF_DP_DEFERRED");
  DeviceGuard guard(m->ctx->device);
  AF_TRY(af_mlp_dp_join(m));
  m->dp_mode = mode;
  return AF_OK;
}

int af_mlp_set_host_shard(af_mlp* m, int enable) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  m->host_shard = enable != 0;
  return AF_OK;
}

int af_mlp_bind_slot(af_mlp* m, int slot) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  if (slot != 0 && slot != 1) return fail(AF_EINVAL, "slot must be 0 or 1");
  DeviceGuard guard(m->ctx->device);
  AF_TRY(mlp_pipeline_init(m));
  mlp_bind_slot(m, slot);
  return AF_OK;
}

int af_mlp_set_graph(af_mlp* m, int enable) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  m->use_graph = enable != 0;
  if (!enable) mlp_drop_graphs(m);
  return AF_OK;
}

static bool mlp_small_ok(af_mlp* m) {
  return !m->grad_ready && small_mlp_applicable(m->ctx, m->L, m->sizes.data(), m->n);
}
// fresh: the accumulators start this step at zero (stored, not read-modify-written);
// unit_upstream: outGrad = 1, outRGrad = 0 (bench/mlp.go:118-126) generated in registers
static int mlp_small_step(af_mlp* m, int with_r, int backprop, bool fresh, bool unit_upstream) {
  std::vector<const double*> rv(2 * m->L);
  for (int i = 0; i < 2 * m->L; ++i) rv[i] = m->has_rvec[i] ? m->rvec[i] : nullptr;
  return small_mlp_launch(m->ctx, m->L, m->sizes.data(), m->n, m->param.data(), rv.data(), m->grad.data(),
                          m->rgrad.data(), m->act.data(), m->ract.data(), m->delta.data(), m->rdelta.data(), with_r != 0,
                          backprop != 0, fresh, unit_upstream);
}

// upstream of a fresh step: from the host, or outGrad = 1 / outRGrad = 0 (bench/mlp.go:118-126)
static int mlp_load_upstream(af_mlp* m, int with_r, const double* hU, const double* hUR) {
  af_ctx* ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int L = m->L;
  const int64_t top = m->n * m->sizes[L];
  if (hU) AF_CUDA(cudaMemcpyAsync(m->delta[L], hU, (size_t)top * 8, cudaMemcpyHostToDevice, st));
  else AF_TRY(af_fill(ctx, m->delta[L], 1.0, top));  // bench/mlp.go:120-123
  if (with_r) {
    if (hUR) AF_CUDA(cudaMemcpyAsync(m->rdelta[L], hUR, (size_t)top * 8, cudaMemcpyHostToDevice, st));
    else AF_CUDA(cudaMemsetAsync(m->rdelta[L], 0, (size_t)top * 8, st));  // bench/mlp.go:124
  }
  return AF_OK;
}

// One iteration of the reference benchmark (bench/mlp.go:43-57 with :118-126): fresh accumulators, upstream from the
// host (or outGrad = 1 / outRGrad = 0 when NULL), forward + backward.  The inputs are already in act[0].
// dp: sum the accumulators across the context's communicator afterwards (Gradient.Add across ranks), scheduled per
// the plan's dp mode.
static int mlp_fresh_step(af_mlp* m, int with_r, const double* hU, const double* hUR, bool dp) {
  af_ctx* ctx = m->ctx;
  dp = dp && dp_active(ctx);
  const int mode = dp ? (m->dp_mode != AF_DP_OFF ? m->dp_mode : AF_DP_OVERLAP) : AF_DP_OFF;
  const bool unit = !hU && !hUR;
  const bool captured = m->use_graph && !m->grad_ready && !ctx->tracing && ctx->stream != nullptr && ctx->stream != cudaStreamLegacy;
  // the head's forward kernel can generate the unit upstream itself (MlpHead): no fill / zero launches then
  const bool unit_in_head = unit && !captured && !mlp_small_ok(m) && mlp_head_fusable(m, with_r);
  if (mode >= AF_DP_OVERLAP && !captured && !mlp_small_ok(m) && !m->grad_ready) {
    MlpHead head;
    head.backprop = true;
    head.unit = unit_in_head;
    if (!unit_in_head) AF_TRY(mlp_load_upstream(m, with_r, hU, hUR));  // (the upstream buffers are no one else's)
    if (mode == AF_DP_DEFERRED) {
      // the previous step's reduce may still be running (deferred join): only the accumulators depend on it, so the
      // forward pass goes first -- without the head's bias gradient -- and the join sits in front of the zeroing
      head.bias = false;
      AF_TRY(mlp_forward_launches(m, with_r, &head));
    }
    AF_TRY(mlp_dp_join_slot(m, m->bound_slot));
    AF_CUDA(cudaMemsetAsync(m->grad[0], 0, (size_t)m->gr_doubles * 8, ctx->stream));
    if (mode != AF_DP_DEFERRED) AF_TRY(mlp_forward_launches(m, with_r, &head));
    AF_TRY(mlp_backward_launches(m, with_r, true, &head));
    if (mode == AF_DP_OVERLAP) AF_TRY(mlp_dp_join_slot(m, m->bound_slot));
    return AF_OK;
  }
  AF_TRY(mlp_dp_join_slot(m, m->bound_slot));
  int rc = AF_ENOTSUP;
  if (unit && mlp_small_ok(m)) {
    rc = mlp_small_step(m, with_r, 1, true, true);
    if (rc == AF_ENOTSUP) ctx->small_mlp_off = true;  // cooperative launch refused: per-kernel path from now on
  }
  if (rc == AF_ENOTSUP) {
    if (!unit_in_head || mlp_small_ok(m)) AF_TRY(mlp_load_upstream(m, with_r, hU, hUR));
    if (mlp_small_ok(m)) {
      rc = mlp_small_step(m, with_r, 1, true, false);
      if (rc == AF_ENOTSUP) ctx->small_mlp_off = true;
    }
  }
  if (rc == AF_ENOTSUP) {
    AF_TRY(af_mlp_zero_grads(m));
    if (unit_in_head && !mlp_small_ok(m)) {  // eager launches (no step graph at this batch size), unit upstream from the head kernel
      MlpHead head;
      head.backprop = true;
      head.unit = true;
      AF_TRY(mlp_forward_launches(m, with_r, &head));
      rc = mlp_backward_launches(m, with_r, false, &head);
    } else {
      if (unit_in_head) AF_TRY(mlp_load_upstream(m, with_r, hU, hUR));  // (the persistent kernel was refused just now)
      rc = af_mlp_step(m, with_r, 1);
    }
  }
  AF_TRY(rc);
  if (dp) AF_TRY(dp_allreduce(ctx, m->grad[0], m->gr_doubles, ctx->stream, true));
  return AF_OK;
}

int af_mlp_step_fresh(af_mlp* m, int with_r) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  DeviceGuard guard(m->ctx->device);
  AF_NOT_RECORDING(m->ctx, "af_mlp_step_fresh (plans replay graphs of their own)");
  return mlp_fresh_step(m, with_r, nullptr, nullptr, false);
}

int af_mlp_step_fresh_dp(af_mlp* m, int with_r) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  DeviceGuard guard(m->ctx->device);
  AF_NOT_RECORDING(m->ctx, "af_mlp_step_fresh_dp (plans replay graphs of their own)");
  return mlp_fresh_step(m, with_r, nullptr, nullptr, true);
}

int af_mlp_step(af_mlp* m, int with_r, int backprop) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  af_ctx* ctx = m->ctx;
  DeviceGuard guard(ctx->device);
  AF_NOT_RECORDING(ctx, "af_mlp_step (plans replay graphs of their own)");
  if (backprop) AF_TRY(mlp_dp_join_slot(m, m->bound_slot));  // the step accumulates into the slot's gradients
  // a handful of samples: the whole step is one persistent kernel (small_mlp.cu), no graph needed
  if (mlp_small_ok(m)) {
    const int rc = mlp_small_step(m, with_r, backprop, false, false);
    if (rc != AF_ENOTSUP) return rc;
    ctx->small_mlp_off = true;  // cooperative launch refused: per-kernel path from now on
  }
  // the legacy default stream cannot be captured
  if (!m->use_graph || m->grad_ready || ctx->tracing || ctx->stream == nullptr || ctx->stream == cudaStreamLegacy)
    return mlp_step_launches(m, with_r, backprop);
  af_mlp::StepGraph& g = m->graphs[with_r ? 1 : 0][backprop ? 1 : 0][m->bound_slot];
  if (g.exec && g.epoch != ctx->option_epoch) {  // an option changed since the capture: the decomposition may differ
    cudaGraphExecDestroy(g.exec);
    g = af_mlp::StepGraph();
    g.seen = 1;  // kernel attributes are set; capture again right away
  }
  if (g.exec) {
    AF_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
    ctx->launches += g.launches;
    return AF_OK;
  }
  if (g.seen++ == 0) return mlp_step_launches(m, with_r, backprop);  // first call also sets kernel attributes
  const int64_t before = ctx->launches;
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();  // capture not possible on this stream: run eagerly from now on
    m->use_graph = false;
    return mlp_step_launches(m, with_r, backprop);
  }
  const int rc = mlp_step_launches(m, with_r, backprop);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
  if (rc != AF_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
  g.launches = ctx->launches - before;
  g.epoch = ctx->option_epoch;
  ctx->launches = before;
  const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) { g.exec = nullptr; return cuda_fail(ie, "cudaGraphInstantiate"); }
  AF_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
  ctx->launches += g.launches;
  return AF_OK;
}

// af_mlp_step(with_r, 1) from accumulators the caller zeroed, then the sum over ranks (per the plan's dp mode)
int af_mlp_step_dp(af_mlp* m, int with_r) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  af_ctx* ctx = m->ctx;
  DeviceGuard guard(ctx->device);
  AF_NOT_RECORDING(ctx, "af_mlp_step_dp (plans replay graphs of their own)");
  const bool dp = dp_active(ctx);
  const int mode = dp ? (m->dp_mode != AF_DP_OFF ? m->dp_mode : AF_DP_OVERLAP) : AF_DP_OFF;
  const bool captured = m->use_graph && !m->grad_ready && !ctx->tracing && ctx->stream != nullptr && ctx->stream != cudaStreamLegacy;
  if (mode >= AF_DP_OVERLAP && !captured && !mlp_small_ok(m) && !m->grad_ready) {
    AF_TRY(mlp_dp_join_slot(m, m->bound_slot));
    MlpHead head;
    head.backprop = true;
    AF_TRY(mlp_forward_launches(m, with_r, &head));
    AF_TRY(mlp_backward_launches(m, with_r, true, &head));
    if (mode == AF_DP_OVERLAP) AF_TRY(mlp_dp_join_slot(m, m->bound_slot));
    return AF_OK;
  }
  AF_TRY(af_mlp_step(m, with_r, 1));
  if (dp) AF_TRY(dp_allreduce(ctx, m->grad[0], m->gr_doubles, ctx->stream, true));
  return AF_OK;
}

}  // extern "C"

static int mlp_forward_launches(af_mlp* m, int with_r, MlpHead* head) {
  af_ctx* ctx = m->ctx;
  const int L = m->L;
  const int64_t n = m->n;
  const bool R = with_r != 0;
  auto rv = [&](int idx) -> const double* { return (R && m->has_rvec[idx]) ? m->rvec[idx] : nullptr; };
  // forward: func.go:40-45 over {LinTran, LinAdd, Sigmoid} triples, one fused kernel per layer.
  // The input is not in the RVector (bench/mlp.go:52; variable.go:85-91) so RX of layer 1 is zero.
  for (int l = 0; l < L; ++l) {
    const double* RX = (R && l > 0) ? m->ract[l] : nullptr;
    if (l == L - 1 && head && head->backprop && mlp_head_fusable(m, with_r)) {
      FewRowsHead h;
      h.unit = head->unit;
      h.U = m->delta[L]; h.UR = R ? m->rdelta[L] : nullptr;
      if (head->bias) { h.gb = m->grad[2 * l + 1]; h.rgb = R ? m->rgrad[2 * l + 1] : nullptr; }
      AF_TRY(few_rows_fwd_launch(ctx, m->param[2 * l], rv(2 * l), m->param[2 * l + 1], rv(2 * l + 1), m->act[l], RX, m->act[l + 1],
                                 R ? m->ract[l + 1] : nullptr, 1, &h, m->sizes[l + 1], m->sizes[l], n, R));
      head->done = true;
      head->bias_done = head->bias;
      continue;
    }
    AF_TRY(af_dense_fwd(ctx, m->param[2 * l], rv(2 * l), m->param[2 * l + 1], rv(2 * l + 1), m->act[l], RX,
                        m->act[l + 1], R ? m->ract[l + 1] : nullptr, m->sizes[l + 1], m->sizes[l], n, 1));
  }
  return AF_OK;
}

// Backward pass.  dp_overlap = false: the reference's order, layer by layer from the top (bias gradient, weight
// gradient, input gradient).  dp_overlap = true (data parallel, AF_DP_OVERLAP / AF_DP_DEFERRED): the input-gradient
// chain -- the only true dependency chain -- runs first; the weight / bias gradients follow, blocks under 1 MB first
// (their all-reduce rides with a neighbour's), then the large layers largest first, and every completed run of adjacent
// layer blocks is all-reduced on the communicator's stream while the remaining contractions run (stream-K launches are
// sized for sm_count - reserved SMs meanwhile).  The results are identical: every product reads the same operands.
static int mlp_backward_launches(af_mlp* m, int with_r, bool dp_overlap, const MlpHead* head) {
  af_ctx* ctx = m->ctx;
  const int L = m->L;
  const int64_t n = m->n;
  const bool R = with_r != 0;
  auto rv = [&](int idx) -> const double* { return (R && m->has_rvec[idx]) ? m->rvec[idx] : nullptr; };
  // top sigmoid (math_funcs.go:326-333 / 357-374) on the caller's upstream, in place -- unless the head's forward
  // kernel already did it (MlpHead)
  const int64_t top = n * m->sizes[L];
  const bool head_done = head && head->done, head_bias_done = head && head->bias_done;
  if (!head_done) {
    if (R) AF_TRY(af_sigmoid_bwd_r(ctx, m->act[L], m->ract[L], m->delta[L], m->rdelta[L], top));
    else AF_TRY(af_sigmoid_bwd(ctx, m->act[L], m->delta[L], top));
  }
  // bias_done[l]: gb_l / rgb_l (lin_add.go:94-104) were already added -- by the head's forward kernel (top layer) or by the
  // epilogue of the input-gradient launch that produced layer l's upstream
  std::unique_ptr<bool[]> bias_done(new bool[L]());
  bias_done[L - 1] = head_bias_done;
  auto input_gradient = [&](int l) -> int {  // lin_tran.go:420-423 with the sigmoid backward of the layer below fused in
    if (m->sizes[l] * n == 0) return AF_OK;
    return igrad_impl(ctx, m->delta[l + 1], R ? m->rdelta[l + 1] : nullptr, m->param[2 * l], rv(2 * l), m->act[l],
                      R ? m->ract[l] : nullptr, m->delta[l], R ? m->rdelta[l] : nullptr, m->sizes[l + 1], m->sizes[l], n,
                      m->grad[2 * l - 1], R ? m->rgrad[2 * l - 1] : nullptr, &bias_done[l - 1]);
  };
  auto block_end = [&](int l) -> double* { return l + 1 < L ? m->grad[2 * l + 2] : m->grad[0] + m->gr_doubles; };
  if (!dp_overlap) {
    for (int l = L - 1; l >= 0; --l) {
      const int64_t rows = m->sizes[l + 1], cols = m->sizes[l];
      const double* U = m->delta[l + 1];
      const double* UR = R ? m->rdelta[l + 1] : nullptr;
      // lin_add.go:94-104 batched: gb += colsum(U), rgb += colsum(UR)
      if (!bias_done[l])
        AF_TRY(colsum_launch(ctx, U, UR, m->grad[2 * l + 1], R ? m->rgrad[2 * l + 1] : nullptr, rows, n));
      // lin_tran.go:410-418
      const double* RX = (R && l > 0) ? m->ract[l] : nullptr;
      double* gW = m->grad[2 * l];
      double* rgW = R ? m->rgrad[2 * l] : nullptr;
      const int chunks = (m->grad_ready && l == 0 && m->grad_ready_chunks > 1) ? m->grad_ready_chunks : 1;
      // 9..16 samples: one pass over W_l for its weight gradient AND the input gradient toward layer l - 1, the sigmoid
      // backward of that layer applied by the block that folds the row strips (csrc/few_samples.cu)
      if (chunks == 1 && l > 0 && few_samples_applicable(ctx, rows, cols, n, {gW, rgW, m->delta[l], m->rdelta[l], m->act[l], m->ract[l]})) {
        const int rc = few_samples_bwd_launch(ctx, U, UR, m->param[2 * l], rv(2 * l), m->act[l], RX, gW, rgW, m->delta[l],
                                              R ? m->rdelta[l] : nullptr, m->act[l], R ? m->ract[l] : nullptr, rows, cols, n, false);
        if (rc != AF_ENOTSUP) {
          AF_TRY(rc);
          if (m->grad_ready) m->grad_ready(m->grad_ready_user, gW, block_end(l) - gW);
          continue;
        }
      }
      if (chunks == 1) {
        AF_TRY(af_lintran_wgrad_r(ctx, U, UR, m->act[l], RX, gW, rgW, rows, cols, n));
        if (m->grad_ready)  // the whole layer block gW_l .. rgb_l is complete
          m->grad_ready(m->grad_ready_user, gW, block_end(l) - gW);
      } else {
        // row chunks of the weight gradient (rows of W = columns of U): each chunk's all-reduce can start while
        // the next chunk is being computed
        m->grad_ready(m->grad_ready_user, m->grad[2 * l + 1], m->param_len(2 * l + 1));
        if (R) m->grad_ready(m->grad_ready_user, m->rgrad[2 * l + 1], m->param_len(2 * l + 1));
        const int64_t step_rows = ((rows + chunks - 1) / chunks + 63) / 64 * 64;
        for (int64_t r0 = 0; r0 < rows; r0 += step_rows) {
          const int64_t rc = rows - r0 < step_rows ? rows - r0 : step_rows;
          GemmOperand A{U + r0, UR ? UR + r0 : nullptr, LAY_RC, rows};
          GemmOperand Bop{m->act[l], RX, LAY_RC, cols};
          GemmEpilogue e;
          e.epi = EPI_ACCUM; e.out0 = gW + r0 * cols; e.out1 = rgW ? rgW + r0 * cols : nullptr; e.ldo = cols;
          AF_TRY(gemm_launch(ctx, A, Bop, rgW != nullptr, rc, cols, n, e));
          m->grad_ready(m->grad_ready_user, gW + r0 * cols, rc * cols);
          if (rgW) m->grad_ready(m->grad_ready_user, rgW + r0 * cols, rc * cols);
        }
      }
      // lin_tran.go:420-423: the input Variable is not in the gradient maps, so layer 1 stops here
      if (l > 0) AF_TRY(input_gradient(l));
    }
    return AF_OK;
  }

  // ---- data-parallel order ----
  const int slot = m->bound_slot;
  while ((int)m->dp_layer_ev.size() < L) {
    cudaEvent_t ev, fe;
    AF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    m->dp_layer_ev.push_back(ev);
    AF_CUDA(cudaEventCreateWithFlags(&fe, cudaEventDisableTiming));
    m->dp_fence_ev.push_back(fe);
  }
  if (!m->dp_done[slot]) AF_CUDA(cudaEventCreateWithFlags(&m->dp_done[slot], cudaEventDisableTiming));
  for (int l = L - 1; l >= 1; --l) AF_TRY(input_gradient(l));
  std::vector<int> order;
  const int64_t small_block = 1 << 17;  // doubles (1 MB): a collective this short is all latency
  for (int l = L - 1; l >= 0; --l)
    if (block_end(l) - m->grad[2 * l] < small_block) order.push_back(l);
  std::vector<int> large;
  for (int l = 0; l < L; ++l)
    if (block_end(l) - m->grad[2 * l] >= small_block) large.push_back(l);
  std::stable_sort(large.begin(), large.end(), [&](int a, int b) {
    return (block_end(a) - m->grad[2 * a]) > (block_end(b) - m->grad[2 * b]);
  });
  const size_t n_small = order.size();
  order.insert(order.end(), large.begin(), large.end());
  std::vector<char> done(L, 0), reduced(L, 0);
  int issued = 0;
  auto reduce_run = [&](int l) -> int {  // the maximal run of complete, not yet reduced, adjacent blocks around l
    int lo = l, hi = l;
    while (lo > 0 && done[lo - 1] && !reduced[lo - 1]) --lo;
    while (hi + 1 < L && done[hi + 1] && !reduced[hi + 1]) ++hi;
    for (int k = lo; k <= hi; ++k) reduced[k] = 1;
    cudaEvent_t ev = m->dp_layer_ev[issued % L], fence = m->dp_fence_ev[issued % L];
    AF_CUDA(cudaEventRecord(ev, ctx->stream));
    AF_CUDA(cudaStreamWaitEvent(ctx->dp_stream, ev, 0));
    // Launch fence.  The collective's CTAs need whole SMs (640 threads x 96 registers) and the next contraction on the
    // ctx stream would fill every SM before a cross-stream launch gets through -- the collective would then wait for
    // that kernel's end (measured in round 1: "overlap" no faster than blocking).  So the ctx stream itself waits
    // for an event the communicator's stream records right before the collective: the collective is launched
    // in-stream behind it, the next contraction only after the event has crossed streams, and takes the other SMs.
    AF_CUDA(cudaEventRecord(fence, ctx->dp_stream));
    AF_TRY(dp_allreduce(ctx, m->grad[2 * lo], block_end(hi) - m->grad[2 * lo], ctx->dp_stream));
    AF_CUDA(cudaStreamWaitEvent(ctx->stream, fence, 0));
    ++issued;
    ctx->reserved_sms = ctx->dp_reserve_sms >= 0 ? ctx->dp_reserve_sms : ctx->dp_max_ctas;
    return AF_OK;
  };
  struct ReserveScope { af_ctx* c; ~ReserveScope() { c->reserved_sms = 0; } } reserve_scope{ctx};
  for (size_t q = 0; q < order.size(); ++q) {
    const int l = order[q];
    const int64_t rows = m->sizes[l + 1], cols = m->sizes[l];
    const double* U = m->delta[l + 1];
    const double* UR = R ? m->rdelta[l + 1] : nullptr;
    if (!bias_done[l])
      AF_TRY(colsum_launch(ctx, U, UR, m->grad[2 * l + 1], R ? m->rgrad[2 * l + 1] : nullptr, rows, n));
    AF_TRY(af_lintran_wgrad_r(ctx, U, UR, m->act[l], (R && l > 0) ? m->ract[l] : nullptr, m->grad[2 * l],
                              R ? m->rgrad[2 * l] : nullptr, rows, cols, n));
    done[l] = 1;
    if (q >= n_small) AF_TRY(reduce_run(l));
  }
  for (int l = 0; l < L; ++l)  // runs made of small blocks only
    if (done[l] && !reduced[l]) AF_TRY(reduce_run(l));
  AF_CUDA(cudaEventRecord(m->dp_done[slot], ctx->dp_stream));
  m->dp_pending[slot] = true;
  return AF_OK;
}

// ---- pipelined host-buffer steps ------------------------------------------------------------------
static int mlp_pipeline_init(af_mlp* m) {
  if (m->copy_in) return AF_OK;
  const int L = m->L, P = 2 * L;
  auto pad = [](int64_t x) { return (x + 31) / 32 * 32; };
  const int64_t gr_doubles = m->gr_doubles;  // grad | rgrad, padded, adjacent
  const int64_t x_doubles = pad(m->n * m->sizes[0]), o_doubles = pad(m->n * m->sizes[L]);
  AF_CUDA(cudaMalloc(&m->alt, (size_t)(gr_doubles + x_doubles + 2 * o_doubles) * 8));
  AF_CUDA(cudaStreamCreateWithFlags(&m->copy_in, cudaStreamNonBlocking));
  AF_CUDA(cudaStreamCreateWithFlags(&m->copy_out, cudaStreamNonBlocking));
  m->slot[0].x = m->act[0]; m->slot[0].grad = m->grad; m->slot[0].rgrad = m->rgrad;
  m->slot[0].out = m->act[L]; m->slot[0].rout = m->ract[L];
  m->slot[1].grad.resize(P); m->slot[1].rgrad.resize(P);
  for (int i = 0; i < P; ++i) {
    m->slot[1].grad[i] = m->alt + (m->grad[i] - m->grad[0]);
    m->slot[1].rgrad[i] = m->alt + (m->rgrad[i] - m->grad[0]);
  }
  m->slot[1].x = m->alt + gr_doubles;
  m->slot[1].out = m->slot[1].x + x_doubles;
  m->slot[1].rout = m->slot[1].out + o_doubles;
  for (auto& sl : m->slot) {
    AF_CUDA(cudaEventCreateWithFlags(&sl.in_ready, cudaEventDisableTiming));
    AF_CUDA(cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming));
    AF_CUDA(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
  }
  return AF_OK;
}

static void mlp_bind_slot(af_mlp* m, int k) {
  m->bound_slot = k;
  const af_mlp::Slot& sl = m->slot[k];
  m->act[0] = sl.x; m->grad = sl.grad; m->rgrad = sl.rgrad;
  m->act[m->L] = sl.out; m->ract[m->L] = sl.rout;
}

// device-to-host copies of a step's results on `st`.  With af_mlp_set_host_shard on a data-parallel plan only this
// rank's 1/nranks of the reduced {grad | rgrad} block crosses to the host (the ranks' host buffers together hold the
// reduced gradient exactly once).
static int mlp_download(af_mlp* m, int with_r, double* hOut, double* hROut, double* const* hGrads, double* const* hRGrads,
                        cudaStream_t st) {
  const int L = m->L;
  const int64_t top = m->n * m->sizes[L];
  if (hOut) AF_CUDA(cudaMemcpyAsync(hOut, m->act[L], (size_t)top * 8, cudaMemcpyDeviceToHost, st));
  if (hROut && with_r) AF_CUDA(cudaMemcpyAsync(hROut, m->ract[L], (size_t)top * 8, cudaMemcpyDeviceToHost, st));
  int64_t lo = 0, hi = m->gr_doubles;
  if (m->host_shard && dp_active(m->ctx)) {
    lo = m->gr_doubles * m->ctx->dp_rank / m->ctx->dp_nranks;
    hi = m->gr_doubles * (m->ctx->dp_rank + 1) / m->ctx->dp_nranks;
  }
  auto part = [&](double* h, const double* d, int64_t len) -> int {
    if (!h) return AF_OK;
    const int64_t off = d - m->grad[0];
    const int64_t a = off > lo ? off : lo, b = off + len < hi ? off + len : hi;
    if (a < b) AF_CUDA(cudaMemcpyAsync(h + (a - off), d + (a - off), (size_t)(b - a) * 8, cudaMemcpyDeviceToHost, st));
    return AF_OK;
  };
  for (int i = 0; i < 2 * L; ++i) {
    if (hGrads) AF_TRY(part(hGrads[i], m->grad[i], m->param_len(i)));
    if (with_r && hRGrads) AF_TRY(part(hRGrads[i], m->rgrad[i], m->param_len(i)));
  }
  return AF_OK;
}

extern "C" {

int af_mlp_wait(af_mlp* m) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  if (m->waited >= m->submitted) return AF_OK;
  DeviceGuard guard(m->ctx->device);
  af_mlp::Slot& sl = m->slot[m->waited & 1];
  AF_CUDA(cudaEventSynchronize(sl.out_done));
  sl.busy = false;
  m->waited += 1;
  return AF_OK;
}

// Phase 1 of a pipelined step: upload on the copy-in stream, compute on the ctx stream into the next slot.
int af_mlp_submit_compute_host(af_mlp* m, int with_r, const double* hX, const double* hU, const double* hUR) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  AF_NEED(hX);
  af_ctx* ctx = m->ctx;
  DeviceGuard guard(ctx->device);
  AF_NOT_RECORDING(ctx, "af_mlp_submit_compute_host");
  AF_TRY(mlp_pipeline_init(m));
  while (m->submitted - m->waited >= 2) AF_TRY(af_mlp_wait(m));  // at most two steps in flight
  const int k = (int)(m->submitted & 1);
  af_mlp::Slot& sl = m->slot[k];
  // the slot's previous user has fully drained (out_done waited above)
  AF_CUDA(cudaMemcpyAsync(sl.x, hX, (size_t)(m->n * m->sizes[0]) * 8, cudaMemcpyHostToDevice, m->copy_in));
  AF_CUDA(cudaEventRecord(sl.in_ready, m->copy_in));
  cudaStream_t st = ctx->stream;
  AF_CUDA(cudaStreamWaitEvent(st, sl.in_ready, 0));
  mlp_bind_slot(m, k);
  AF_TRY(mlp_fresh_step(m, with_r, hU, hUR, m->dp_mode != AF_DP_OFF));
  sl.busy = true;
  return AF_OK;
}

// Phase 2: everything queued on the ctx stream so far (the step, and e.g. the caller's all-reduce of the slot's
// accumulators) gates the download, which runs on the copy-out stream while the next step computes.
int af_mlp_submit_download_host(af_mlp* m, int with_r, double* hOut, double* hROut, double* const* hGrads,
                                double* const* hRGrads) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  af_ctx* ctx = m->ctx;
  DeviceGuard guard(ctx->device);
  const int k = (int)(m->submitted & 1);
  af_mlp::Slot& sl = m->slot[k];
  if (!sl.busy || !m->copy_out) return fail(AF_EINVAL, "af_mlp_submit_download_host without a pending af_mlp_submit_compute_host");
  AF_CUDA(cudaEventRecord(sl.compute_done, ctx->stream));
  cudaStream_t co = m->copy_out;
  AF_CUDA(cudaStreamWaitEvent(co, sl.compute_done, 0));
  // a deferred reduce of this slot gates the download, not the ctx stream (which joins when it next touches the slot)
  if (m->dp_pending[k]) AF_CUDA(cudaStreamWaitEvent(co, m->dp_done[k], 0));
  if (m->bound_slot != k) mlp_bind_slot(m, k);
  AF_TRY(mlp_download(m, with_r, hOut, hROut, hGrads, hRGrads, co));
  AF_CUDA(cudaEventRecord(sl.out_done, co));
  m->submitted += 1;
  return AF_OK;
}

int af_mlp_submit_host(af_mlp* m, int with_r, const double* hX, const double* hU, const double* hUR, double* hOut,
                       double* hROut, double* const* hGrads, double* const* hRGrads) {
  AF_TRY(af_mlp_submit_compute_host(m, with_r, hX, hU, hUR));
  return af_mlp_submit_download_host(m, with_r, hOut, hROut, hGrads, hRGrads);
}

int af_mlp_step_host(af_mlp* m, int with_r, const double* hX, const double* hU, const double* hUR, double* hOut,
                     double* hROut, double* const* hGrads, double* const* hRGrads) {
  if (!m) return fail(AF_EINVAL, "null mlp");
  AF_NEED(hX);
  af_ctx* ctx = m->ctx;
  DeviceGuard guard(ctx->device);
  while (m->waited < m->submitted) AF_TRY(af_mlp_wait(m));
  if (m->copy_in) mlp_bind_slot(m, 0);
  cudaStream_t st = ctx->stream;
  AF_CUDA(cudaMemcpyAsync(m->act[0], hX, (size_t)(m->n * m->sizes[0]) * 8, cudaMemcpyHostToDevice, st));
  AF_TRY(mlp_fresh_step(m, with_r, hU, hUR, m->dp_mode != AF_DP_OFF));
  AF_TRY(mlp_dp_join_slot(m, m->bound_slot));
  AF_TRY(mlp_download(m, with_r, hOut, hROut, hGrads, hRGrads, st));
  AF_CUDA(cudaStreamSynchronize(st));
  return AF_OK;
}

}  // extern "C"
