// Host front end of the FP64 contraction kernels: TMA descriptor construction, path selection
// (TMA + DMMA vs. generic), split-K policy, launches.
#include "common.h"
#include "gemm_dmma.cuh"

#include <atomic>
#include <utility>

namespace af {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool tma_compatible(const double* p, long long ld, long long rows, long long kd) {
  if (p == nullptr) return true;
  if (reinterpret_cast<uintptr_t>(p) & 15) return false;  // TMA global address: 16-byte aligned
  if (ld & 1) return false;                               // row pitch: multiple of 16 bytes
  if (ld * 8 >= (1ll << 40)) return false;
  if (rows >= (1ll << 31) || kd >= (1ll << 31)) return false;
  return true;
}

// KC: element (r,k) at p[r*ld + k] -> dims {Kd, R}, box {16, tile_rows}
// RC: element (r,k) at p[k*ld + r] -> dims {R, Kd}, box {16, 16}
int make_map(af_ctx* ctx, CUtensorMap* map, const double* p, int layout, long long rows, long long kd, long long ld,
             int tile_rows) {
  return gemm_make_map(ctx, map, p, layout, rows, kd, ld, tile_rows);
}

}  // namespace

int gemm_make_map(af_ctx* ctx, CUtensorMap* map, const double* p, int layout, long long rows, long long kd, long long ld,
                  int tile_rows) {
  cuuint64_t dims[2];
  cuuint32_t box[2];
  if (layout == LAY_KC) {
    dims[0] = (cuuint64_t)kd; dims[1] = (cuuint64_t)rows;
    box[0] = kBK; box[1] = (cuuint32_t)tile_rows;
  } else {
    dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)kd;
    box[0] = 16; box[1] = kBK;
  }
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(p), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AF_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return AF_OK;
}

// Unswizzled 2-D map of a row-major matrix (element (r, c) at p[r*ld + c]) with box {box_cols, box_rows}: the box lands
// in shared memory as box_rows lines of box_cols doubles (few_samples.cu: the sample-side operands U / UR).
int gemm_make_plain_map(af_ctx* ctx, CUtensorMap* map, const double* p, long long rows, long long cols, long long ld, int box_rows,
                        int box_cols) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(p), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AF_ECUDA, "cuTensorMapEncodeTiled (plain map) failed with CUresult " + std::to_string((int)r));
  return AF_OK;
}

// The stacked LSTM gate matrix W4 (4H x C, gate-major rows: bench/lstm.go's four H x C matrices one after another) as a
// 3-D tensor {k, unit, gate} with box {16, 16, 4}: one box = the 64 rows {4 gates} x {16 units} of a forward-step tile
// (gemm_dmma_kernel SPECIAL = 1), landing in shared memory as 64 consecutive 128-byte lines exactly like a KC tile.
// Only the first H columns (the state part of concat(state, x)) are mapped.
int gemm_make_gate_map(af_ctx* ctx, CUtensorMap* map, const double* w4, long long H, long long C) {
  cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)H, 4};
  cuuint64_t strides[2] = {(cuuint64_t)C * 8, (cuuint64_t)H * C * 8};
  cuuint32_t box[3] = {kBK, 16, 4};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(w4), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AF_ECUDA, "cuTensorMapEncodeTiled (gate map) failed with CUresult " + std::to_string((int)r));
  return AF_OK;
}

namespace {

// algorithmic flops of one launch: 2 I J Kd per product that the reference would also compute
double gemm_flops(const GemmArgs& a, bool dual) {
  int products = 0;
  if (a.out0) products += 1;
  if (dual && a.out1) products += (a.has_a1 ? 1 : 0) + (a.has_b1 ? 1 : 0);
  return 2.0 * a.I * a.J * a.Kd * products;
}

template <int ALAY, int BLAY, int NACC, int CFG, bool MULTI, int OPS = -1>
int launch_dmma_ops(af_ctx* ctx, const CUtensorMap* maps, const GemmArgs& args, dim3 grid) {
  auto kern = gemm_dmma_kernel<ALAY, BLAY, NACC, CFG, MULTI, 0, OPS>;
  const int smem = (int)GemmSmem<NACC, CFG>::kTotal;
  // per template instantiation; the attribute is per device AND function.  Contexts on different devices may be driven
  // by different host threads: a device's bit is published only after its attribute call has returned (two threads may
  // both make the call, which is harmless).
  static std::atomic<unsigned long long> configured{0};
  if (!(configured.load(std::memory_order_acquire) >> (ctx->device & 63) & 1ull)) {
    AF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.fetch_or(1ull << (ctx->device & 63), std::memory_order_release);
  }
  const cudaError_t le = launch_kernel(kern, grid, dim3(GemmSmem<NACC, CFG>::kThreads), (size_t)smem, ctx->stream,
                                       ctx->pdl && !ctx->pdl_off, maps[0], maps[1], maps[2], maps[3], args, SpecialArgs<0>{});
  if (le != cudaSuccess) return cuda_fail(le, "gemm_dmma_kernel launch");
  return after_launch(ctx, "gemm_dmma_kernel", NACC == 2 ? KIND_GEMM_DUAL : KIND_GEMM_SINGLE, gemm_flops(args, NACC == 2));
}

// stream-K with the ordered fix-up (gemm_dmma_kernel SPECIAL = 3): paired configuration only
template <int ALAY, int BLAY, int NACC>
int launch_dmma_fixup(af_ctx* ctx, const CUtensorMap* maps, const GemmArgs& args, dim3 grid) {
  auto kern = gemm_dmma_kernel<ALAY, BLAY, NACC, 2, true, 3>;
  const int smem = (int)GemmSmem<NACC, 2>::kTotal;
  static std::atomic<unsigned long long> configured{0};
  if (!(configured.load(std::memory_order_acquire) >> (ctx->device & 63) & 1ull)) {
    AF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.fetch_or(1ull << (ctx->device & 63), std::memory_order_release);
  }
  SpecialArgs<3> sp{};
  sp.fix.part = ctx->sk_part;
  sp.fix.flags = ctx->sk_flags;
  const cudaError_t le = launch_kernel(kern, grid, dim3(GemmSmem<NACC, 2>::kThreads), (size_t)smem, ctx->stream,
                                       ctx->pdl && !ctx->pdl_off, maps[0], maps[1], maps[2], maps[3], args, sp);
  if (le != cudaSuccess) return cuda_fail(le, "gemm_dmma_kernel (stream-K fix-up) launch");
  return after_launch(ctx, "gemm_dmma_kernel", NACC == 2 ? KIND_GEMM_DUAL : KIND_GEMM_SINGLE, gemm_flops(args, NACC == 2));
}

// workspace of the fix-up: one partial tile per CTA of a stream-K launch (two per SM) and one flag per CTA and warp
int fixup_workspace(af_ctx* ctx) {
  if (ctx->sk_part) return AF_OK;
  const size_t ctas = 2 * (size_t)ctx->sm_count + 1;
  const size_t part_bytes = ctas * 2 * 4 * 4 * 2 * GemmSmem<2, 2>::kThreads * sizeof(double);  // NACC x TI x TJ x 2 values per thread
  AF_CUDA(cudaMalloc(&ctx->sk_part, part_bytes));
  AF_CUDA(cudaMalloc(&ctx->sk_flags, ctas * GemmSmem<2, 2>::kWarps * sizeof(unsigned)));
  AF_CUDA(cudaMemsetAsync(ctx->sk_flags, 0, ctas * GemmSmem<2, 2>::kWarps * sizeof(unsigned), ctx->stream));
  return AF_OK;
}

// the paired configuration of the big dual contractions is compiled per set of present R operands (gemm_dmma_kernel OPS)
template <int ALAY, int BLAY, int NACC, int CFG, bool MULTI>
int launch_dmma_impl(af_ctx* ctx, const CUtensorMap* maps, const GemmArgs& args, dim3 grid) {
  // (one-tile launches only: the stream-K instantiations measured SLOWER with the presence compiled in -- bench weight
  // gradients 762 -> 795 and 1057 -> 1103 us -- while the one-tile forward / input-gradient launches gain 6-8 us each)
  if constexpr (NACC == 2 && CFG == 2 && !MULTI) {
    if (!ctx->ops_runtime) {
      if (args.has_a1 && args.has_b1) return launch_dmma_ops<ALAY, BLAY, NACC, CFG, MULTI, 3>(ctx, maps, args, grid);
      if (args.has_b1) return launch_dmma_ops<ALAY, BLAY, NACC, CFG, MULTI, 2>(ctx, maps, args, grid);
      if (args.has_a1) return launch_dmma_ops<ALAY, BLAY, NACC, CFG, MULTI, 1>(ctx, maps, args, grid);
    }
  }
  return launch_dmma_ops<ALAY, BLAY, NACC, CFG, MULTI, -1>(ctx, maps, args, grid);
}

// stream-K (ranges spanning tiles) is instantiated for the default and the narrow configuration only
template <int ALAY, int BLAY, int NACC, int CFG>
int launch_dmma(af_ctx* ctx, const CUtensorMap* maps, const GemmArgs& args, dim3 grid) {
  if (CFG != 0 && args.units_per_tile == args.kt_total && args.units_per_cta != args.kt_total)
    return launch_dmma_impl<ALAY, BLAY, NACC, CFG, (CFG != 0)>(ctx, maps, args, grid);
  return launch_dmma_impl<ALAY, BLAY, NACC, CFG, false>(ctx, maps, args, grid);
}

template <int ALAY, int BLAY>
int launch_generic(af_ctx* ctx, const GemmOperand& A, const GemmOperand& B, bool dual, const GemmArgs& args) {
  dim3 grid((unsigned)((args.J + 15) / 16), (unsigned)((args.I + 15) / 16));
  if (dual)
    gemm_generic_kernel<ALAY, BLAY, true><<<grid, 256, 0, ctx->stream>>>(A.p0, A.p1, B.p0, B.p1, A.ld, B.ld, args);
  else
    gemm_generic_kernel<ALAY, BLAY, false><<<grid, 256, 0, ctx->stream>>>(A.p0, nullptr, B.p0, nullptr, A.ld, B.ld, args);
  return after_launch(ctx, "gemm_generic_kernel", KIND_GEMM_GENERIC, gemm_flops(args, dual));
}

}  // namespace

// Deferred epilogue: applies the fused elementwise step (bias / sigmoid / sigmoid backward) in place on
// outputs that were accumulated with split-K atomics.  Only small problems (few output tiles) take it.
template <bool DUAL>
__global__ void __launch_bounds__(256) gemm_epilogue_kernel(const GemmArgs p) {
  const long long n = (long long)p.I * p.J;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) {
    const long long i = e / p.J;
    const int j = (int)(e - i * p.J);
    const long long o = i * p.ldo + j;
    double v0 = p.out0 ? p.out0[o] : 0.0, v1 = (DUAL && p.out1) ? p.out1[o] : 0.0;
    epi_transform<DUAL>(p, i, j, v0, v1);
    if (p.out0) p.out0[o] = v0;
    if (DUAL && p.out1) p.out1[o] = v1;
  }
}

int gemm_deferred_epilogue(af_ctx* ctx, const GemmArgs& logical, bool dual) {
  long long blocks = ((long long)logical.I * logical.J + 255) / 256;
  if (blocks > 8ll * ctx->sm_count) blocks = 8ll * ctx->sm_count;
  if (blocks < 1) blocks = 1;
  if (dual) gemm_epilogue_kernel<true><<<(unsigned)blocks, 256, 0, ctx->stream>>>(logical);
  else gemm_epilogue_kernel<false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(logical);
  return after_launch(ctx, "gemm_epilogue_kernel", KIND_ELEMENTWISE, 16.0 * (dual ? 2 : 1) * logical.I * logical.J);
}

int gemm_launch(af_ctx* ctx, const GemmOperand& A_in, const GemmOperand& B_in, bool dual, long long I, long long J,
                long long Kd, const GemmEpilogue& e) {
  if (I <= 0 || J <= 0) return AF_OK;
  if (I >= (1ll << 31) || J >= (1ll << 31) || Kd >= (1ll << 31)) return fail(AF_ESHAPE, "contraction extent exceeds 2^31");
  GemmOperand A = A_in, B = B_in;
  GemmArgs args{};
  args.epi = e.epi;
  args.I = (int)I; args.J = (int)J; args.Kd = (int)Kd;
  args.has_a1 = dual && A.p1 != nullptr;
  args.has_b1 = dual && B.p1 != nullptr;
  args.out0 = e.out0; args.out1 = dual ? e.out1 : nullptr; args.ldo = e.ldo;
  args.bias0 = e.bias0; args.bias1 = e.bias1;
  args.s = e.s; args.rs = e.rs; args.lds = e.lds;
  const GemmArgs logical = args;  // un-swapped view, used by the generic kernel and the deferred epilogue
  const bool accumulate = e.epi == EPI_ACCUM;
  const int kt_total = (int)((Kd + kBK - 1) / kBK);

  const bool tma_ok = ctx->encode_tiled != nullptr && tma_compatible(A.p0, A.ld, I, Kd) &&
                      tma_compatible(A.p1, A.ld, I, Kd) && tma_compatible(B.p0, B.ld, J, Kd) &&
                      tma_compatible(B.p1, B.ld, J, Kd) && Kd > 0;
  const bool use_dmma = ctx->gemm_path != 1 && tma_ok;
  if (!use_dmma && ctx->gemm_path == 2)
    return fail(AF_ENOTSUP, "DMMA path required but operands violate TMA alignment (16-byte base, even pitch)");

  if (!use_dmma) {
    if (A.layout == LAY_KC && B.layout == LAY_KC) return launch_generic<LAY_KC, LAY_KC>(ctx, A, B, dual, args);
    if (A.layout == LAY_KC && B.layout == LAY_RC) return launch_generic<LAY_KC, LAY_RC>(ctx, A, B, dual, args);
    if (A.layout == LAY_RC && B.layout == LAY_RC) return launch_generic<LAY_RC, LAY_RC>(ctx, A, B, dual, args);
    return fail(AF_EINVAL, "unsupported operand layout combination");
  }

  // Very few rows (n <= 16 samples; the 10-row weight gradient of the output layer): compute the
  // transposed product so the short extent lands on the 16-wide J tile of the narrow configuration and
  // the long operand (the weight matrix) is streamed exactly once.  acc1 is symmetric under
  // (A0,A1) <-> (B0,B1).
  const bool contiguous_out = (e.ldo == J);
  bool swapped = false;
  // (deterministic mode: only where the transposed result can be added in place -- transposed STORES go through the
  // atomic partial-tile path and its deferred epilogue)
  if (I <= 16 && J > 16 && (accumulate || (contiguous_out && !ctx->deterministic))) {
    std::swap(A, B);
    std::swap(I, J);
    std::swap(args.has_a1, args.has_b1);
    args.I = (int)I; args.J = (int)J;
    args.transpose_out = 1;
    swapped = true;
  }
  const int cfg = J <= 16 ? 1 : ctx->big_cfg;
  const int bm = cfg == 1 ? GemmSmem<2, 1>::kBMc : cfg == 2 ? GemmSmem<2, 2>::kBMc : GemmSmem<2, 0>::kBMc;
  const int bn = cfg == 1 ? GemmSmem<2, 1>::kBNc : cfg == 2 ? GemmSmem<2, 2>::kBNc : GemmSmem<2, 0>::kBNc;
  const long long tiles_j = (J + bn - 1) / bn, tiles_i = (I + bm - 1) / bm;
  const long long tiles = tiles_i * tiles_j;
  // the narrow configuration streams the long operand and is latency-bound (ncu: 12 % warps active, tensor pipe 34 %,
  // stalls = long_scoreboard / wait with one CTA per SM): two CTAs per SM (80 KB of stages each) double the bytes in flight
  // (while a plan keeps an all-reduce in flight beside its contractions the collective's CTAs own `reserved_sms` SMs)
  const int slots = (ctx->sm_count - ctx->reserved_sms) * (cfg == 0 ? 1 : 2);
  args.kt_total = kt_total;
  args.tiles_j = (int)tiles_j;
  args.total_units = tiles * kt_total;

  // Work decomposition (see gemm_dmma_kernel), chosen by a small cost model in k-tile units per CTA slot:
  //   one tile per CTA : waves(tiles) * (kt + 0.5)                       fused epilogue, plain stores
  //   aligned split-K s: waves(tiles*s) * (ceil(kt/s) + cf)              partial tiles via red.global.add.f64
  //   stream-K         : U + cf * (segments per CTA), U = total/slots    equal contiguous unit ranges
  // Store-type outputs pay cd more on the two atomic paths (zeroing + gemm_epilogue_kernel afterwards).
  const double cf = 3.5, cd = accumulate ? 0.0 : 3.0;
  auto waves_of = [&](long long ctas) { return (double)((ctas + slots - 1) / slots); };
  // deterministic mode: whole tiles per CTA only -- split-K and stream-K add partial tiles with red.global.add.f64 in
  // arrival order
  const bool may_split = kt_total >= 8 && (accumulate || contiguous_out) && !ctx->deterministic;
  const bool must_defer = swapped && !accumulate;  // transposed stores always go through the deferred epilogue
  int mode = 0, best_s = 1;                         // 0 = one tile per CTA, 1 = aligned split, 2 = stream-K
  double best = must_defer ? 1e300 : waves_of(tiles) * (kt_total + 0.5);
  long long sk_units = (args.total_units + slots - 1) / slots;
  if (sk_units < 4) sk_units = 4;
  if (may_split || must_defer) {
    for (int sp = 2; sp <= 64 && sp <= kt_total / 4; ++sp) {
      const double c = waves_of(tiles * sp) * ((kt_total + sp - 1) / sp + cf) + cd;
      if (c < best * 0.98) { best = c; mode = 1; best_s = sp; }
    }
    const double segs = 1.0 + (double)(sk_units - 1) / kt_total;
    // measured (profiles/r01s4_sk_bias_ab.txt): ranges of >= 64 k-tiles amortise a segment's fixed cost better than the
    // model's cf says (bench weight gradient 1082 -> 1065 us, MLP n = 1024 step 1.371 -> 1.326 ms as stream-K), short
    // ranges do not (the LSTM's per-step contractions, 15 k-tiles per CTA: 24.3 -> 26.2 ms)
    // (store-type launches also pay the zeroing and the deferred pass: they lean only while the last-wave loss is
    // large -- under four waves of tiles, e.g. 512 tiles = 1.73 waves at n = 1024; C3's 6.9-wave forward / input-
    // gradient launches lose 0.3 % when they flip)
    const int bias = (sk_units >= 64 && (accumulate || tiles < 4 * slots)) ? ctx->sk_bias_long : ctx->sk_bias;
    const double c = bias * 0.01 * (sk_units + cf * segs) + cd;
    if (cfg != 0 && (c < best * 0.98 || (must_defer && mode == 0))) { best = c; mode = 2; }
  }
  // Stream-K with the ORDERED fix-up (SPECIAL = 3 of the kernel: no atomics, no zeroing, no deferred pass, the same bits on
  // every run): the paired configuration, ranges of at least 16 k-tiles, one wave.  af_set_option "sk_fixup": 0 never;
  // 1 (default) = in deterministic mode wherever stream-K balances better than whole tiles (the mode has nothing else),
  // in the default mode for STORE-type launches -- they save the zeroing and the deferred elementwise pass; accumulating
  // launches keep red.global.add.f64, measured 8-18 us faster per bench weight gradient --; 2 = wherever stream-K is chosen.
  // Store-type launches are costed without the deferred pass (cd) and with their own bias ("sk_fix_bias", percent).
  const bool fixup_ok = cfg == 2 && !swapped && kt_total >= 8 && sk_units >= 16 &&
                        (long long)((args.total_units + sk_units - 1) / sk_units) <= slots &&
                        (ctx->sk_fixup == 2 || (ctx->sk_fixup == 1 && (ctx->deterministic || !accumulate)));
  if (fixup_ok) {
    const double segs = 1.0 + (double)(sk_units - 1) / kt_total;
    if (mode == 2) {
      mode = 4;
    } else if (ctx->deterministic) {
      const int bias = (sk_units >= 64 && (accumulate || tiles < 4 * slots)) ? ctx->sk_bias_long : ctx->sk_bias;
      if (mode == 0 && bias * 0.01 * (sk_units + cf * segs) < best * 0.98) mode = 4;
    } else if (!accumulate) {
      if (ctx->sk_fix_bias * 0.01 * (sk_units + cf * segs) < best * 0.98) mode = 4;
    }
  }
  // Persistent whole tiles: a launch of several waves of one-tile CTAs pays every CTA's prologue (barrier setup,
  // pipeline fill) and epilogue in the open (ncu: tensor pipe 85 % of elapsed vs 92 % for the 296-CTA stream-K
  // launches).  Give each of <= `slots` CTAs m whole tiles instead: the TMA ring runs on into the next tile while the
  // current one is stored, the epilogue stays fused (no atomics, no zeroing, no deferred pass).
  long long tiles_per_cta = 0;
  if (AF_PERSIST_TILES && mode == 0 && ctx->persist_tiles && cfg != 0 && tiles > slots) {
    tiles_per_cta = (tiles + slots - 1) / slots;
    mode = 3;
  }
  bool deferred = false;
  long long units_per_cta = kt_total, units_per_tile = kt_total;
  if (mode == 3) {
    units_per_cta = tiles_per_cta * kt_total;
  } else if (mode == 1) {
    units_per_cta = (kt_total + best_s - 1) / best_s;
    units_per_tile = units_per_cta * best_s;
  } else if (mode == 2 || mode == 4) {
    units_per_cta = sk_units;
    args.rotate = (ctx->sk_rotate && sk_units >= 64) ? 1 : 0;
  }
  if (mode == 4) AF_TRY(fixup_workspace(ctx));
  // column sums of a sigmoid-backward output ride on the epilogue of whole-tile launches (float64 atomics: not in
  // deterministic mode, whose column sums are folded in block order by colsum_kernel)
  if (mode == 0 && !swapped && cfg == 2 && e.epi == EPI_SIGMOID_BWD && (e.colsum0 || e.colsum1) && e.colsum_done &&
      !ctx->deterministic && !ctx->fused_colsum_off && A.layout == LAY_KC && B.layout == LAY_RC) {
    args.epi = EPI_SIGMOID_BWD_COLSUM;
    args.bias0 = e.colsum0;
    args.bias1 = dual ? e.colsum1 : nullptr;
    *e.colsum_done = true;
  }
  if (mode == 1 || mode == 2) {
    if (!accumulate) {
      deferred = true;
      const size_t bytes = (size_t)logical.I * logical.J * 8;
      if (args.out0 && !e.prezeroed) AF_CUDA(cudaMemsetAsync(args.out0, 0, bytes, ctx->stream));
      if (args.out1 && !e.prezeroed) AF_CUDA(cudaMemsetAsync(args.out1, 0, bytes, ctx->stream));
    }
    args.epi = EPI_ACCUM_ATOMIC;
  }
  args.units_per_cta = (int)units_per_cta;
  args.units_per_tile = (int)units_per_tile;
  args.total_units = tiles * units_per_tile;
  dim3 grid((unsigned)((args.total_units + units_per_cta - 1) / units_per_cta), 1, 1);
  if (mode == 3 && ctx->persist_tiles == 2 && tiles_per_cta * (tiles_per_cta - 1) <= tiles) {
    // wave order: CTA c takes tiles c, c + grid, ...; the first `wave_rem` CTAs have m tiles, the rest m - 1
    // (m (m - 1) <= tiles guarantees that the first m - 1 rounds are complete)
    args.wave_tiles = (int)tiles_per_cta;
    args.wave_rem = (int)(tiles - (tiles_per_cta - 1) * (long long)grid.x);
  }

  CUtensorMap maps[4];
  AF_TRY(make_map(ctx, &maps[0], A.p0, A.layout, I, Kd, A.ld, bm));
  AF_TRY(make_map(ctx, &maps[2], B.p0, B.layout, J, Kd, B.ld, bn));
  if (args.has_a1) AF_TRY(make_map(ctx, &maps[1], A.p1, A.layout, I, Kd, A.ld, bm)); else maps[1] = maps[0];
  if (args.has_b1) AF_TRY(make_map(ctx, &maps[3], B.p1, B.layout, J, Kd, B.ld, bn)); else maps[3] = maps[2];

  const int nacc = dual ? 2 : 1;
  int rc = AF_EINVAL;
  bool dispatched = false;
  if (mode == 4) {
#define AF_FIXUP(AL, BL)                                                                                                   \
  if (!dispatched && A.layout == AL && B.layout == BL) {                                                                   \
    dispatched = true;                                                                                                     \
    rc = nacc == 2 ? launch_dmma_fixup<AL, BL, 2>(ctx, maps, args, grid) : launch_dmma_fixup<AL, BL, 1>(ctx, maps, args, grid); \
  }
    AF_FIXUP(LAY_KC, LAY_KC)
    AF_FIXUP(LAY_KC, LAY_RC)
    AF_FIXUP(LAY_RC, LAY_RC)
#undef AF_FIXUP
    if (!dispatched) return fail(AF_EINVAL, "unsupported operand layout combination");
    return rc;
  }
#define AF_DISPATCH(AL, BL)                                                                        \
  if (!dispatched && A.layout == AL && B.layout == BL) {                                           \
    dispatched = true;                                                                             \
    if (cfg == 0)                                                                                  \
      rc = nacc == 2 ? launch_dmma<AL, BL, 2, 0>(ctx, maps, args, grid) : launch_dmma<AL, BL, 1, 0>(ctx, maps, args, grid); \
    else if (cfg == 2)                                                                             \
      rc = nacc == 2 ? launch_dmma<AL, BL, 2, 2>(ctx, maps, args, grid) : launch_dmma<AL, BL, 1, 2>(ctx, maps, args, grid); \
    else                                                                                           \
      rc = nacc == 2 ? launch_dmma<AL, BL, 2, 1>(ctx, maps, args, grid) : launch_dmma<AL, BL, 1, 1>(ctx, maps, args, grid); \
  }
  AF_DISPATCH(LAY_KC, LAY_KC)
  AF_DISPATCH(LAY_KC, LAY_RC)
  AF_DISPATCH(LAY_RC, LAY_RC)
#undef AF_DISPATCH
  if (!dispatched && A.layout == LAY_RC && B.layout == LAY_KC && cfg == 1) {  // swapped input-gradient, narrow only
    dispatched = true;
    rc = nacc == 2 ? launch_dmma<LAY_RC, LAY_KC, 2, 1>(ctx, maps, args, grid) : launch_dmma<LAY_RC, LAY_KC, 1, 1>(ctx, maps, args, grid);
  }
  if (!dispatched) return fail(AF_EINVAL, "unsupported operand layout combination");
  AF_TRY(rc);
  if (deferred && logical.epi != EPI_STORE) return gemm_deferred_epilogue(ctx, logical, dual);
  return AF_OK;
}

}  // namespace af
