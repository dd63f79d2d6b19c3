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


# ---- the ordered fix-up of stream-K (gemm_dmma_kernel SPECIAL = 3): roles of a launch's CTAs ------------------------------
def _fixup_roles(kt_total, units_per_cta, total_units):
    """The kernel's segment walk restated (gemm_dmma.cuh: the outer loop of the consumers and its FIXUP flush calls): per
    CTA the list of (tile, first k-tile, length, kind, head_len) it flushes; kind 0 = whole tile, 1 = parks its partial,
    2 = head (adds the followers' partials, runs the epilogue)."""
    out = []
    n_ctas = (total_units + units_per_cta - 1) // units_per_cta
    for c in range(n_ctas):
        u0 = c * units_per_cta
        nun = min(u0 + units_per_cta, total_units) - u0
        tile, kt0 = divmod(u0, kt_total)
        kt_cur, un, seg_kt0, seg_un0, segs = kt0, 0, kt0, 0, []
        while un < nun:
            seg_end = min(un + (kt_total - kt_cur), nun)
            un = seg_end
            if un >= nun:
                break
            segs.append((tile, seg_kt0, un - seg_un0, 1 if seg_kt0 > 0 else 0, 0))
            seg_kt0, seg_un0, kt_cur, tile = 0, un, 0, tile + 1
        last = nun - seg_un0
        segs.append((tile, seg_kt0, last, 1 if seg_kt0 > 0 else (0 if last == kt_total else 2), last))
        out.append(segs)
    return out


def test_streamk_fixup_roles_cover_every_tile_once_in_cta_order():
    """Every tile is finished by exactly one CTA (a whole-tile segment, or the head); a head's followers -- the next
    ceil((kt_total - head_len) / units_per_cta) CTAs, as the kernel computes -- each park exactly one partial, their
    FIRST segment, and together with the head they cover the tile's k-tiles exactly once, in ascending order; no CTA
    parks more than one partial (one workspace slot per CTA) and a parked partial is always consumed."""
    import random
    rng = random.Random(7)
    cases = [(125, 216, 512 * 125), (256, 443, 512 * 256), (256, 222, 256 * 256), (63, 436, 2048 * 63)]  # the bench's launches
    for _ in range(400):
        kt = rng.randint(8, 300)
        tiles = rng.randint(1, 700)
        slots = rng.randint(1, 296)
        total = tiles * kt
        cases.append((kt, max(16, (total + slots - 1) // slots), total))
    for kt, U, total in cases:
        roles = _fixup_roles(kt, U, total)
        parked = {}      # cta -> (tile, first k-tile, length)
        finished = {}    # tile -> cta
        heads = []
        for c, segs in enumerate(roles):
            n_parked = 0
            for i, (tile, k0, ln, kind, head_len) in enumerate(segs):
                assert ln > 0
                if kind == 1:
                    assert i == 0 and k0 > 0, "only a CTA's first segment can start inside a tile"
                    n_parked += 1
                    parked[c] = (tile, k0, ln)
                elif kind == 0:
                    assert k0 == 0 and ln == kt
                    assert tile not in finished
                    finished[tile] = c
                else:
                    assert i == len(segs) - 1 and k0 == 0 and ln < kt and head_len == ln
                    assert tile not in finished
                    finished[tile] = c
                    heads.append((c, tile, ln))
            assert n_parked <= 1
        assert sorted(finished) == list(range(total // kt)), (kt, U, total)
        consumed = set()
        for c, tile, head_len in heads:
            followers = (kt - head_len + U - 1) // U        # the kernel's formula
            k = head_len
            for fo in range(1, followers + 1):
                assert c + fo in parked, (kt, U, total, c, fo)
                ptile, pk0, pln = parked[c + fo]
                assert ptile == tile and pk0 == k, "followers continue the tile in CTA order"
                k += pln
                consumed.add(c + fo)
            assert k == kt, "head + followers cover the tile exactly"
        assert consumed == set(parked), "every parked partial has exactly one reader"
