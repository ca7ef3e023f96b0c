// Host emulation of the package-merge of one warp (zip-ada_b200/csrc/b2_pm.cuh): the merge-path version and the
// binary-search version against a plain sequential package-merge (lists merged with "package before a leaf of equal
// weight", the first 2n-2 items of the last list, walk down).  Test infrastructure; usage: emu_pm <cases> <seed>
#include "cuda_emu.h"
#define B2_EMU 1
#include "../../zip-ada_b200/csrc/b2_pm.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>

void b2_set_error(const char *, int, const char *) {}

struct Item { u64 w; bool pkg; };

static std::vector<int> reference_lengths(const std::vector<u32> &leafw, int max_bits) {
  const int ns = (int)leafw.size(), need = 2 * ns - 2;
  std::vector<std::vector<Item>> lists(max_bits);
  for (int i = 0; i < ns; i++) lists[0].push_back(Item{leafw[i], false});
  for (int lev = 1; lev < max_bits; lev++) {
    const auto &prev = lists[lev - 1];
    std::vector<u64> pk;
    for (size_t b = 0; b + 1 < prev.size(); b += 2) pk.push_back((u32)(prev[b].w + prev[b + 1].w));
    size_t a = 0, b = 0;
    auto &cur = lists[lev];
    while ((a < (size_t)ns || b < pk.size()) && (int)cur.size() < need) {
      const bool takep = b < pk.size() && (a >= (size_t)ns || pk[b] <= leafw[a]);
      if (takep) cur.push_back(Item{pk[b++], true}); else cur.push_back(Item{leafw[a++], false});
    }
  }
  std::vector<int> len(ns, 0);
  int k = need;
  for (int lev = max_bits - 1; lev >= 0; lev--) {
    int p = 0;
    for (int i = 0; i < k && i < (int)lists[lev].size(); i++) p += lists[lev][i].pkg;
    const int a = k - p;
    for (int i = 0; i < a && i < ns; i++) len[i]++;
    k = 2 * p;
  }
  return len;
}

int main(int argc, char **argv) {
  const int cases = argc > 1 ? atoi(argv[1]) : 200;
  std::mt19937 rng(argc > 2 ? atoi(argv[2]) : 1);
  int bad = 0;
  for (int cs = 0; cs < cases; cs++) {
    const int sizes[] = {2, 3, 4, 5, 17, 31, 32, 33, 64, 65, 100, 139, 200, 257, 258, 260};
    const int ns = cs < 16 ? sizes[cs] : 2 + (int)(rng() % 259);
    const int max_bits = (cs % 3 == 0) ? 15 : (cs % 3 == 1 ? 17 : 16);
    std::vector<u32> w(ns);
    const int dist = cs % 6;
    for (int i = 0; i < ns; i++) {
      switch (dist) {
        case 0: w[i] = 1; break;                                          // all equal: ties everywhere
        case 1: w[i] = 1 + rng() % 4; break;                              // many ties
        case 2: w[i] = 1u << std::min(22, i / 2); break;                  // geometric: the length limit binds
        case 3: w[i] = 1 + rng() % 900000; break;
        case 4: w[i] = 1 + (u32)(900000.0 / (1 + rng() % (ns * 4))); break;   // Zipf-like
        default: w[i] = (i % 7 == 0) ? 1 + rng() % 50000 : 1; break;
      }
    }
    std::sort(w.begin(), w.end());
    std::vector<u32> sym(ns);
    for (int i = 0; i < ns; i++) sym[i] = i;
    std::shuffle(sym.begin(), sym.end(), rng);
    const u32 alpha = (u32)ns + (rng() % 3);                              // scratch sized by a larger alphabet of the batch
    const std::vector<int> ref = reference_lengths(w, max_bits);
    for (int variant = 0; variant < 2; variant++) {
      std::vector<u32> scratch(ll_scratch_words(alpha) + 8, 0xCDCDCDCDu);
      std::vector<u8> lens(B2_MAX_ALPHA + 8, 0xEE);
      for (int i = 0; i < ns; i++) scratch[i] = (w[i] << 9) | sym[i];
      emu_launch(1, 32, 0, 1, [&]() {
        LLScratch S = ll_scratch_at(scratch.data(), alpha);
        if (variant == 0) ll_package_merge_warp(S, ns, max_bits, lens.data());
        else ll_package_merge_warp_mp(S, ns, max_bits, lens.data());
      });
      if (scratch[ll_scratch_words(alpha)] != 0xCDCDCDCDu) { printf("case %d variant %d: scratch overrun\n", cs, variant); bad++; }
      for (int i = 0; i < ns; i++)
        if (lens[sym[i]] != ref[i]) { printf("case %d (ns %d, max_bits %d, dist %d) variant %d: leaf %d has length %d, expected %d\n", cs, ns, max_bits, dist, variant, i, lens[sym[i]], ref[i]); bad++; break; }
      // Kraft sum of a complete code
      u64 kraft = 0;
      for (int i = 0; i < ns; i++) kraft += 1ull << (max_bits - lens[sym[i]]);
      if (ns >= 2 && kraft != (1ull << max_bits)) { printf("case %d variant %d: Kraft sum %llu\n", cs, variant, (unsigned long long)kraft); bad++; }
    }
  }
  printf(bad ? "FAILED\n" : "OK\n");
  return bad ? 1 : 0;
}
