// Order in which a stream-K CTA accumulates the k-tiles of one tile segment (gemm_dmma.cuh, producer lane).
// Host + device so that tests/test_streamk_walk_cpu.py can compile it with g++ and check the walk exhaustively.
#pragma once

#if defined(__CUDACC__)
#define AF_HD __host__ __device__ __forceinline__
#else
#define AF_HD inline
#endif

namespace af {

// First k-tile of a segment [b, b + len) of a tile with kt_total k-tiles that the CTA starts at local time tau (units
// since the CTA's start).  The segment is walked from there upwards, wrapping from b + len back to b.  Without
// `rotate` that is simply b.  With it, r = (first k-tile - b) is chosen so that k-tile == tau' (mod kt_total) at every
// local time tau' of the longer of the two pieces the wrap cuts the walk into: all CTAs of a one-wave stream-K launch
// start together, so those that are inside whole tiles sweep k in lockstep and share operand panels in L2.
AF_HD int sk_seg_start(int b, int len, int tau, int kt_total, bool rotate) {
  if (!rotate) return b;
  const int r1 = ((tau - b) % kt_total + kt_total) % kt_total;  // aligned before the wrap, for len - r1 units
  const int r2 = (r1 + len) % kt_total;                          // aligned after the wrap, for r2 units
  int r = 0, span = 0;
  if (r1 < len) { r = r1; span = len - r1; }
  if (r2 < len && r2 > span) r = r2;
  return b + r;
}

}  // namespace af
