"""CPU test of the stream-K rotated k-walk (autofunc_b200/csrc/streamk_walk.h, used by the TMA producer lane of
gemm_dmma_kernel): the header is host + device, so it is compiled here with g++ and driven through the producer's
bookkeeping (restated below from gemm_dmma.cuh `produce`) over random unit ranges.  Checked: every segment's k-tiles
are visited exactly once, tiles in range order, the unrotated walk is the plain ascending one, and on the bench's
layer-1 weight gradient (512 tiles x 256 k-tiles on 296 CTA slots) most units are in lockstep (k-tile == local time
mod kt_total) -- the property that lets co-running CTAs share operand panels in L2."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>
#include "streamk_walk.h"

// the producer's bookkeeping, as in gemm_dmma.cuh (MULTI): returns the (tile, k-tile) sequence of one CTA
static std::vector<std::pair<int, int>> walk(int kt_total, int kt0, int tile0, int nun, bool rotate) {
  std::vector<std::pair<int, int>> out;
  int p_q = tile0, p_b = kt0, p_len = std::min(kt_total - kt0, nun), p_done = 0, p_tau = 0;
  int p_kt = af::sk_seg_start(p_b, p_len, 0, kt_total, rotate);
  for (int t = 0; t < nun; ++t) {
    out.push_back({p_q, p_kt});
    ++p_kt;
    if (p_kt == p_b + p_len) p_kt = p_b;
    if (++p_done == p_len) {
      p_tau += p_len;
      p_b = 0; p_done = 0;
      p_len = std::min(kt_total, nun - p_tau);
      p_kt = af::sk_seg_start(0, p_len, p_tau, kt_total, rotate);
      ++p_q;
    }
  }
  return out;
}

int main() {
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&](int lo, int hi) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return lo + (int)(s % (unsigned long long)(hi - lo + 1)); };
  for (int trial = 0; trial < 20000; ++trial) {
    const int kt = rnd(1, 300), kt0 = rnd(0, kt - 1), nun = rnd(1, 900), tile0 = rnd(0, 5);
    std::vector<std::pair<int, int>> expect;
    for (int t = 0, q = tile0, k = kt0; t < nun; ++t) { expect.push_back({q, k}); if (++k == kt) { k = 0; ++q; } }
    for (int rot = 0; rot < 2; ++rot) {
      auto got = walk(kt, kt0, tile0, nun, rot);
      if (got.size() != expect.size()) { std::printf("FAIL size\n"); return 1; }
      for (size_t i = 0; i < got.size(); ++i)
        if (got[i].first != expect[i].first) { std::printf("FAIL tile order kt=%d kt0=%d nun=%d\n", kt, kt0, nun); return 1; }
      if (!rot && got != expect) { std::printf("FAIL unrotated walk differs\n"); return 1; }
      auto a = got, b = expect;
      std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
      if (a != b) { std::printf("FAIL not a permutation kt=%d kt0=%d nun=%d\n", kt, kt0, nun); return 1; }
    }
  }
  // bench layer-1 weight gradient: 2000 x 1000 outputs in 64 x 64 tiles, 4096 samples = 256 k-tiles, 2 x 148 CTA slots
  const int kt = 256;
  const long long total = 512ll * kt, U = (total + 295) / 296;
  long long aligned = 0, units = 0;
  for (long long u0 = 0; u0 < total; u0 += U) {
    const int nun = (int)std::min(U, total - u0);
    auto got = walk(kt, (int)(u0 % kt), (int)(u0 / kt), nun, true);
    for (int t = 0; t < nun; ++t) { ++units; aligned += got[t].second == t % kt; }
  }
  std::printf("OK aligned %.4f\n", (double)aligned / (double)units);
  return 0;
}
"""


def test_rotated_walk_is_a_permutation_and_mostly_in_lockstep(tmp_path):
    src = tmp_path / "walk.cpp"
    src.write_text(SRC)
    exe = tmp_path / "walk"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "autofunc_b200", "csrc"),
                    str(src), "-o", str(exe)], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK aligned ")
    assert float(r.stdout.split()[-1]) > 0.7  # 0.758 measured: whole tiles fully, partial segments on one side of the wrap
