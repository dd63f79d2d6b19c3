// Data parallelism behind the C ABI: the multi-device form of Gradient.Add
// (gradient.go:28-30,108-120 -- the reference's gradient is a SUM over samples, so the sum of the
// ranks' partial gradients IS the full-batch gradient).
//
// One af_ctx = one device = one rank.  The communicator is NCCL (ncclAllReduce, ncclDouble, ncclSum over
// NVLink / NVSwitch).  libnccl is resolved with dlopen at af_dp_* time, not at link time: the library must
// load (and export every symbol) on a host without NCCL or a GPU, and a process that already carries a
// libnccl.so.2 (torch's bundled copy) must share it instead of loading a second one.  The communicator runs
// with a CTA cap (ncclConfig_t.maxCTAs, af_set_option "dp_max_ctas") on its own high-priority stream, so a
// plan can keep an all-reduce in flight beside the contractions at the cost of a few SMs (ctx->reserved_sms).
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>

#include "common.h"

namespace af {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  const char* names[] = {getenv("AUTOFUNC_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    const char* e = dlerror();
    g_nccl.error = std::string("libnccl.so.2 could not be loaded: ") + (e ? e : "unknown dlopen error");
    return;
  }
#define AF_SYM(field, name)                                                        \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.handle, name)); \
  if (!g_nccl.field && g_nccl.error.empty()) g_nccl.error = std::string("libnccl lacks ") + name;
  AF_SYM(GetVersion, "ncclGetVersion")
  AF_SYM(GetUniqueId, "ncclGetUniqueId")
  AF_SYM(CommInitRank, "ncclCommInitRank")
  AF_SYM(CommInitAll, "ncclCommInitAll")
  AF_SYM(CommDestroy, "ncclCommDestroy")
  AF_SYM(AllReduce, "ncclAllReduce")
  AF_SYM(GroupStart, "ncclGroupStart")
  AF_SYM(GroupEnd, "ncclGroupEnd")
  AF_SYM(GetErrorString, "ncclGetErrorString")
#undef AF_SYM
  // optional (NCCL >= 2.14 / 2.18): without them the communicator simply has no CTA cap / no uncapped twin
  g_nccl.CommInitRankConfig =
      reinterpret_cast<decltype(g_nccl.CommInitRankConfig)>(dlsym(g_nccl.handle, "ncclCommInitRankConfig"));
  g_nccl.CommSplit = reinterpret_cast<decltype(g_nccl.CommSplit)>(dlsym(g_nccl.handle, "ncclCommSplit"));
}

int nccl_ready() {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.error.empty()) return fail(AF_ENCCL, g_nccl.error);
  return AF_OK;
}

int nccl_fail(ncclResult_t r, const char* what) {
  return fail(AF_ENCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error") + " (" +
                            std::to_string((int)r) + ")");
}

#define AF_NCCL(expr)                                       \
  do {                                                      \
    ncclResult_t _r = (expr);                               \
    if (_r != ncclSuccess) return nccl_fail(_r, #expr);     \
  } while (0)

int dp_make_stream(af_ctx* ctx) {
  if (ctx->dp_stream) return AF_OK;
  int lo = 0, hi = 0;
  AF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically lowest = highest priority
  AF_CUDA(cudaStreamCreateWithPriority(&ctx->dp_stream, cudaStreamNonBlocking, hi));
  AF_CUDA(cudaEventCreateWithFlags(&ctx->dp_order_ev, cudaEventDisableTiming));
  return AF_OK;
}

}  // namespace

// `wide`: a reduce that nothing overlaps (whole accumulator block after the step, small batches, the LSTM, af_allreduce_sum)
// goes through the UNCAPPED twin of the communicator when there is one: with the 8-CTA cap that keeps an overlapped
// reduce out of the contractions' way, 48.5 MB take 0.36 ms at 8 GPUs against ~0.2 ms with NCCL's own CTA count.
int dp_allreduce(af_ctx* ctx, double* d_ptr, long long n, cudaStream_t stream, bool wide) {
  if (!dp_active(ctx) || n <= 0) return AF_OK;
  void* comm = (wide && ctx->comm_wide) ? ctx->comm_wide : ctx->comm;
  AF_NCCL(g_nccl.AllReduce(d_ptr, d_ptr, (size_t)n, ncclDouble, ncclSum, static_cast<ncclComm_t>(comm), stream));
  ctx->dp_calls += 1;
  return AF_OK;
}

}  // namespace af

using namespace af;

extern "C" {

int af_dp_unique_id(void* id128) {
  if (!id128) return fail(AF_EINVAL, "null id buffer");
  AF_TRY(nccl_ready());
  static_assert(sizeof(ncclUniqueId) == AF_DP_ID_BYTES, "AF_DP_ID_BYTES must equal sizeof(ncclUniqueId)");
  ncclUniqueId id;
  AF_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return AF_OK;
}

int af_dp_init(af_ctx* ctx, int nranks, int rank, const void* id128) {
  if (!ctx || !id128) return fail(AF_EINVAL, "null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(AF_EINVAL, "rank must lie in [0, nranks)");
  if (ctx->comm) return fail(AF_EINVAL, "af_dp_init: the context already has a communicator");
  if (ctx->recording) return fail(AF_EINVAL, "af_dp_init: not allowed between af_graph_begin and af_graph_end");
  AF_TRY(nccl_ready());
  DeviceGuard guard(ctx->device);
  AF_TRY(dp_make_stream(ctx));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  if (g_nccl.CommInitRankConfig && ctx->dp_max_ctas > 0) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.maxCTAs = ctx->dp_max_ctas;
    cfg.minCTAs = 1;
    AF_NCCL(g_nccl.CommInitRankConfig(&comm, nranks, id, rank, &cfg));
  } else {
    AF_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  }
  ctx->comm = comm;
  ctx->dp_rank = rank;
  ctx->dp_nranks = nranks;
  // the uncapped twin for reduces that run alone (same ranks, NCCL's own CTA count); a collective call: every rank makes it
  if (g_nccl.CommInitRankConfig && g_nccl.CommSplit && ctx->dp_max_ctas > 0 && nranks > 1) {
    ncclConfig_t wide_cfg = NCCL_CONFIG_INITIALIZER;
    wide_cfg.maxCTAs = 32;  // explicit: an undefined field of a split's config may inherit the parent's cap
    ncclComm_t wide = nullptr;
    // (not fatal: without the twin every reduce simply uses the capped communicator)
    if (g_nccl.CommSplit(comm, 0, rank, &wide, &wide_cfg) == ncclSuccess) ctx->comm_wide = wide;
  }
  return AF_OK;
}

int af_dp_init_local(af_ctx* const* ctxs, int n) {
  if (!ctxs || n < 1) return fail(AF_EINVAL, "af_dp_init_local: need at least one context");
  if (n > 64) return fail(AF_EINVAL, "af_dp_init_local: at most 64 devices");
  AF_TRY(nccl_ready());
  int devs[64];
  ncclComm_t comms[64];
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return fail(AF_EINVAL, "null context");
    if (ctxs[i]->comm) return fail(AF_EINVAL, "af_dp_init_local: a context already has a communicator");
    for (int j = 0; j < i; ++j)
      if (ctxs[j]->device == ctxs[i]->device) return fail(AF_EINVAL, "af_dp_init_local: two contexts share a device");
    devs[i] = ctxs[i]->device;
  }
  // (ncclCommInitAll takes no config: the CTA cap of a local communicator comes from NCCL_MAX_CTAS, if set)
  AF_NCCL(g_nccl.CommInitAll(comms, n, devs));
  for (int i = 0; i < n; ++i) {
    DeviceGuard guard(ctxs[i]->device);
    ctxs[i]->comm = comms[i];
    ctxs[i]->dp_rank = i;
    ctxs[i]->dp_nranks = n;
    AF_TRY(dp_make_stream(ctxs[i]));
  }
  return AF_OK;
}

int af_dp_destroy(af_ctx* ctx) {
  if (!ctx) return AF_OK;
  DeviceGuard guard(ctx->device);
  if (ctx->dp_stream) cudaStreamSynchronize(ctx->dp_stream);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm_wide && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->comm_wide));
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->comm));
  ctx->comm = nullptr;
  ctx->comm_wide = nullptr;
  ctx->dp_rank = 0;
  ctx->dp_nranks = 1;
  if (ctx->dp_order_ev) { cudaEventDestroy(ctx->dp_order_ev); ctx->dp_order_ev = nullptr; }
  if (ctx->dp_stream) { cudaStreamDestroy(ctx->dp_stream); ctx->dp_stream = nullptr; }
  return AF_OK;
}

int af_dp_info(af_ctx* ctx, int* nranks, int* rank, int* nccl_version) {
  if (!ctx) return fail(AF_EINVAL, "null context");
  if (nranks) *nranks = ctx->dp_nranks;
  if (rank) *rank = ctx->dp_rank;
  if (nccl_version) {
    *nccl_version = 0;
    if (ctx->comm && g_nccl.GetVersion) g_nccl.GetVersion(nccl_version);
  }
  return AF_OK;
}

int af_dp_group_begin(void) {
  AF_TRY(nccl_ready());
  AF_NCCL(g_nccl.GroupStart());
  return AF_OK;
}

int af_dp_group_end(void) {
  AF_TRY(nccl_ready());
  AF_NCCL(g_nccl.GroupEnd());
  return AF_OK;
}

int64_t af_dp_calls(af_ctx* ctx) { return ctx ? ctx->dp_calls : 0; }

int af_allreduce_sum(af_ctx* ctx, double* d_ptr, int64_t n) {
  if (!ctx) return fail(AF_EINVAL, "null context");
  if (n < 0) return fail(AF_ESHAPE, "negative length");
  if (n == 0) return AF_OK;
  if (!d_ptr) return fail(AF_EINVAL, "null pointer argument: d_ptr");
  if (!ctx->comm) return fail(AF_EINVAL, "af_allreduce_sum: the context has no communicator (af_dp_init)");
  DeviceGuard guard(ctx->device);
  // (order this reduce behind whatever a plan still has in flight on the communicator's own stream: kernels of the two
  // communicators are never launched in different orders on different ranks)
  if (ctx->dp_stream && ctx->dp_order_ev) {
    AF_CUDA(cudaEventRecord(ctx->dp_order_ev, ctx->dp_stream));
    AF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->dp_order_ev, 0));
  }
  return dp_allreduce(ctx, d_ptr, n, ctx->stream, true);
}

}  // extern "C"
