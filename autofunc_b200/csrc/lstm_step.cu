// Fused LSTM time steps (north_star: "the seqfunc LSTM step becomes a fused gate-GEMM-plus-pointwise kernel").
//
//   forward step  t : ONE kernel = the four gates' contraction over the recurrent state (K = H; the x part of
//                     concat(state, x), bench/lstm.go:140, does not depend on the recurrence and is projected for all
//                     T x B rows beforehand) + bias + sigmoids + the cell algebra (bench/lstm.go:143-150,187-191) in the
//                     epilogue.  gemm_dmma_kernel SPECIAL = 1: the stacked gate matrix is read through a 3-D TMA tensor
//                     map {k, unit, gate}, so a 64-row tile is {4 gates} x {16 units} and every thread's accumulators
//                     hold all four gates of its (lane, unit) elements.  No atomics, no zeroed gate buffers, no separate
//                     cell kernel.
//   backward step t : ONE kernel = stateGrad += delta4_t W4[:, :H] (the four gates' inputGradient, lin_tran.go:272-359;
//                     K = 4H) split over a thread-block CLUSTER of 4 CTAs, partial tiles exchanged through distributed
//                     shared memory and summed in rank order, + the cell backward of step t - 1 (bench/lstm.go:215-234)
//                     in the epilogue.  gemm_dmma_kernel SPECIAL = 2.  Deterministic; no atomics, no memsets.
//
// Host side only: tensor maps, launch configuration (cluster attribute), applicability.
#include <atomic>

#include "common.h"
#include "gemm_dmma.cuh"

namespace af {

namespace {

constexpr int kCfg = 4;  // 64 x 64 tiles, 8 warps of 16 x 32 (TileCfg<4>: two warps per SM sub-partition with one CTA per SM)

template <typename Kern>
int set_smem_once(af_ctx* ctx, Kern kern, int smem, std::atomic<unsigned long long>& configured) {
  if (!(configured.load(std::memory_order_acquire) >> (ctx->device & 63) & 1ull)) {
    AF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.fetch_or(1ull << (ctx->device & 63), std::memory_order_release);
  }
  return AF_OK;
}

template <int NACC>
int launch_fwd(af_ctx* ctx, const CUtensorMap* maps, const GemmArgs& args, const LstmStepArgs& st, unsigned grid) {
  auto kern = gemm_dmma_kernel<LAY_KC, LAY_KC, NACC, kCfg, false, 1>;
  const int smem = (int)GemmSmem<NACC, kCfg>::kTotal;
  static std::atomic<unsigned long long> configured{0};
  AF_TRY(set_smem_once(ctx, kern, smem, configured));
  const cudaError_t le = launch_kernel(kern, dim3(grid), dim3(GemmSmem<NACC, kCfg>::kThreads), (size_t)smem, ctx->stream,
                                       ctx->pdl && !ctx->pdl_fused_off, maps[0], maps[1], maps[2], maps[3], args, SpecialArgs<1>{st});
  if (le != cudaSuccess) return cuda_fail(le, "lstm forward step launch");
  const int products = 1 + (NACC == 2 ? (args.has_a1 ? 1 : 0) + (args.has_b1 ? 1 : 0) : 0);
  return after_launch(ctx, "gemm_dmma_kernel<lstm fwd step>", NACC == 2 ? KIND_GEMM_DUAL : KIND_GEMM_SINGLE,
                      2.0 * args.I * args.J * args.Kd * products);
}

template <int NACC>
int launch_bwd(af_ctx* ctx, const CUtensorMap* maps, const GemmArgs& args, const LstmStepArgs& st, unsigned grid) {
  auto kern = gemm_dmma_kernel<LAY_KC, LAY_RC, NACC, kCfg, false, 2>;
  const int smem = (int)GemmSmem<NACC, kCfg>::kTotalCluster;
  static std::atomic<unsigned long long> configured{0};
  AF_TRY(set_smem_once(ctx, kern, smem, configured));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GemmSmem<NACC, kCfg>::kThreads);
  cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = GemmSmem<NACC, kCfg>::kSplit; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (ctx->pdl && !ctx->pdl_fused_off) ? 2 : 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], args, SpecialArgs<2>{st});
  if (le != cudaSuccess) return cuda_fail(le, "lstm backward step launch (cluster of 4)");
  const int products = 1 + (NACC == 2 ? (args.has_a1 ? 1 : 0) + (args.has_b1 ? 1 : 0) : 0);
  return after_launch(ctx, "gemm_dmma_kernel<lstm bwd step>", NACC == 2 ? KIND_GEMM_DUAL : KIND_GEMM_SINGLE,
                      2.0 * args.I * args.J * args.Kd * products);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// The fused steps need whole 16-unit groups per tile, a contraction that splits four ways, 16-byte vector accesses
// (even pitches) and TMA-compatible operands; anything else takes the per-kernel path of lstm.cu.
bool lstm_fused_applicable(af_ctx* ctx, long long H, long long C, long long B) {
  return ctx->encode_tiled != nullptr && ctx->gemm_path != 1 && !ctx->lstm_fused_off && H >= 16 && H % 16 == 0 && C % 2 == 0 &&
         B >= 1 && H < (1ll << 28) && B < (1ll << 28);
}

// One forward step on `nb` lanes.  W4 / RW4: stacked gate matrices (4H x C).  The rest as in LstmStepArgs.
int lstm_fwd_step_launch(af_ctx* ctx, const double* W4, const double* RW4, const LstmStepArgs& st, bool with_r, bool r_state,
                         long long nb) {
  const long long H = st.H, C = st.ld;
  for (const void* p : {(const void*)W4, (const void*)RW4, (const void*)st.G, (const void*)st.RG, (const void*)st.INp,
                        (const void*)st.RINp, (const void*)st.INn, (const void*)st.RINn, (const void*)st.AUG, (const void*)st.RAUG})
    if (!aligned16(p)) return fail(AF_EINVAL, "lstm forward step: operand not 16-byte aligned");
  GemmArgs args{};
  args.I = (int)nb; args.J = (int)(4 * H); args.Kd = (int)H;
  args.has_a1 = with_r && r_state;  // the R-state entering step 0 is identically zero: its product is skipped
  args.has_b1 = with_r;
  args.epi = EPI_STORE;
  args.kt_total = (int)(H / kBK);
  args.tiles_j = (int)(H / 16);
  args.units_per_cta = args.kt_total; args.units_per_tile = args.kt_total;
  const long long tiles = ((nb + 63) / 64) * args.tiles_j;
  args.total_units = tiles * args.kt_total;
  CUtensorMap maps[4];
  AF_TRY(gemm_make_map(ctx, &maps[0], st.INp, LAY_KC, nb, H, C, 64));
  if (args.has_a1) AF_TRY(gemm_make_map(ctx, &maps[1], st.RINp, LAY_KC, nb, H, C, 64)); else maps[1] = maps[0];
  AF_TRY(gemm_make_gate_map(ctx, &maps[2], W4, H, C));
  if (args.has_b1) AF_TRY(gemm_make_gate_map(ctx, &maps[3], RW4, H, C)); else maps[3] = maps[2];
  return with_r ? launch_fwd<2>(ctx, maps, args, st, (unsigned)tiles) : launch_fwd<1>(ctx, maps, args, st, (unsigned)tiles);
}

// One backward step on `nb` lanes: SGF += D4 W4[:, :H], then the cell backward of the previous time step (LstmStepArgs).
int lstm_bwd_step_launch(af_ctx* ctx, const double* D4, const double* RD4, const double* W4, const double* RW4,
                         const LstmStepArgs& st, bool with_r, long long nb) {
  const long long H = st.H, C = st.ld;
  for (const void* p : {(const void*)D4, (const void*)RD4, (const void*)W4, (const void*)RW4, (const void*)st.Gm, (const void*)st.RGm,
                        (const void*)st.INm, (const void*)st.RINm, (const void*)st.INt, (const void*)st.RINt, (const void*)st.DAUGm,
                        (const void*)st.RDAUGm, (const void*)st.SGF, (const void*)st.RSGF, (const void*)st.D4m, (const void*)st.RD4m})
    if (!aligned16(p)) return fail(AF_EINVAL, "lstm backward step: operand not 16-byte aligned");
  constexpr int split = GemmSmem<2, kCfg>::kSplit;
  GemmArgs args{};
  args.I = (int)nb; args.J = (int)H; args.Kd = (int)(4 * H);
  args.has_a1 = with_r; args.has_b1 = with_r;
  args.epi = EPI_STORE;
  args.kt_total = (int)(4 * H / kBK);
  args.tiles_j = (int)((H + 63) / 64);
  args.units_per_cta = args.kt_total / split;   // aligned split-K: CTA r of a cluster takes k-tiles [r kt/4, (r+1) kt/4)
  args.units_per_tile = args.kt_total;
  const long long tiles = ((nb + 63) / 64) * args.tiles_j;
  args.total_units = tiles * args.kt_total;
  CUtensorMap maps[4];
  AF_TRY(gemm_make_map(ctx, &maps[0], D4, LAY_KC, nb, 4 * H, 4 * H, 64));
  if (with_r) AF_TRY(gemm_make_map(ctx, &maps[1], RD4, LAY_KC, nb, 4 * H, 4 * H, 64)); else maps[1] = maps[0];
  AF_TRY(gemm_make_map(ctx, &maps[2], W4, LAY_RC, H, 4 * H, C, 64));
  if (with_r) AF_TRY(gemm_make_map(ctx, &maps[3], RW4, LAY_RC, H, 4 * H, C, 64)); else maps[3] = maps[2];
  const unsigned grid = (unsigned)(tiles * split);
  return with_r ? launch_bwd<2>(ctx, maps, args, st, grid) : launch_bwd<1>(ctx, maps, args, st, grid);
}

}  // namespace af
