// Host emulation of k_scatter2 (zip-ada_b200/csrc/b2_scatter2.cuh) against a plain stable counting sort.
// Test infrastructure (tests/test_emu_scatter.py builds and runs it); usage: emu_scatter <mode> <rr_group> <window> <seed> <shift_step>
//   mode: 0..3 = template MODE of k_scatter2 (bit 0 match.any, bit 1 keys loaded early); 40..71 = k_scatter3 (40 + bits: 2 = tiles
//   of 2048 rows, 4 = rotation indices early, 8 = digit from the key registers, 16 = second early look at the predecessor)
// Blocks of several sizes (empty tail tiles, exactly one tile, one row, several tiles), digits of every pass
// position, skewed and uniform digit distributions; every block must come out as the stable sort of its rows by
// the digit, all other arena positions untouched.
#include "cuda_emu.h"
#define B2_EMU 1
#include "../../zip-ada_b200/csrc/b2_scatter2.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>

// what b2_set_error would do in the library
void b2_set_error(const char *, int, const char *) {}

struct Block { u32 n, off; };

// the dispatch order of b2_bwt.cu (build_tiles_rr): blocks in groups, tile k of every block of the group, then tile k + 1 ...
static void tiles_rr(const std::vector<Block> &blocks, size_t group, std::vector<B2SortTileRR> &rr) {
  rr.clear();
  std::vector<u32> last(blocks.size(), 0xFFFFFFFFu);
  for (size_t g0 = 0; g0 < blocks.size(); g0 += group) {
    const size_t g1 = std::min(blocks.size(), g0 + group);
    u32 max_nt = 0;
    for (size_t j = g0; j < g1; j++) max_nt = std::max(max_nt, (blocks[j].n + SC_TILE - 1) / SC_TILE);
    for (u32 k = 0; k < max_nt; k++)
      for (size_t j = g0; j < g1; j++)
        if ((blocks[j].n + SC_TILE - 1) / SC_TILE > k) { const u32 pos = (u32)rr.size(); rr.push_back(B2SortTileRR{(u32)j, k * SC_TILE, last[j], 0}); last[j] = pos; }
  }
}

// the same dispatch order with the self-contained tile records of k_scatter3 (build_tiles_sc of b2_bwt.cu)
static void tiles_sc(const std::vector<Block> &blocks, size_t group, u32 sc_tile, std::vector<B2ScTile> &out) {
  out.clear();
  std::vector<u32> last(blocks.size(), 0xFFFFFFFFu);
  for (size_t g0 = 0; g0 < blocks.size(); g0 += group) {
    const size_t g1 = std::min(blocks.size(), g0 + group);
    u32 max_nt = 0;
    for (size_t j = g0; j < g1; j++) max_nt = std::max(max_nt, (blocks[j].n + sc_tile - 1) / sc_tile);
    for (u32 k = 0; k < max_nt; k++)
      for (size_t j = g0; j < g1; j++)
        if ((blocks[j].n + sc_tile - 1) / sc_tile > k) {
          const u32 pos = (u32)out.size(), start = k * sc_tile, cnt = std::min(sc_tile, blocks[j].n - start);
          out.push_back(B2ScTile{(u32)j | ((cnt - 1) << 16), blocks[j].off + start, last[j], blocks[j].off});
          last[j] = pos;
        }
  }
}

template <int MODE>
static int run(size_t group, unsigned window, unsigned seed, int shift_step) {
  std::mt19937_64 rng(seed);
  const u32 sizes[] = {1, 31, 4095, 4096, 4097, 3 * 4096, 2 * 4096 + 777, 9000, 300, 5 * 4096 + 1, 2047, 2048, 2049};
  std::vector<Block> blocks;
  u32 pos = 64;
  for (u32 n : sizes) { blocks.push_back(Block{n, pos}); pos += (n + 255u) & ~255u; pos += 256; }
  const u32 total = pos + 64;
  std::vector<B2Job> jobs(blocks.size());
  for (size_t j = 0; j < blocks.size(); j++) { std::memset(&jobs[j], 0, sizeof(B2Job)); jobs[j].na = blocks[j].n; jobs[j].n = blocks[j].n + 5; jobs[j].pos_off = blocks[j].off; }
  std::vector<B2SortTileRR> rr;
  std::vector<B2ScTile> sc;
  constexpr int THREADS3 = ((MODE - 40) & 2) ? 256 : 512;           // MODE 40..71: k_scatter3; 40 + bits: 2 = tiles of 2048 rows, 4 / 8 / 16 = SC3_VEARLY / LEAN / MIDLOOK
  if (MODE >= 40) tiles_sc(blocks, group, THREADS3 * SC_ITEMS, sc); else tiles_rr(blocks, group, rr);
  const size_t ntile = MODE >= 40 ? sc.size() : rr.size();
  std::vector<u32> state(ntile * 256, 0);
  int bad = 0;
  u32 pass_no = 0;
  for (int shift = 0; shift < 64; shift += shift_step) {
    for (int dist = 0; dist < 2; dist++) {
      std::vector<u64> kin(total, 0xDEADBEEFDEADBEEFull), kout(total, 0x1111111111111111ull);
      std::vector<u32> vin(total, 0xABABABABu), vout(total, 0x22222222u);
      std::vector<u32> jobhist(blocks.size() * ST_MAXPASS * 256, 0);
      for (size_t j = 0; j < blocks.size(); j++) {
        u32 cnt[256] = {0};
        for (u32 i = 0; i < blocks[j].n; i++) {
          u64 k = rng();
          if (dist == 0) {                       // skewed: few distinct digits, long runs of equal digits
            const u64 dgt = (rng() % 100 < 70) ? (u64)(i / 37 % 3) * 97 : (rng() % 256);
            k = (k & ~(0xFFull << shift)) | (dgt << shift);
          }
          kin[blocks[j].off + i] = k;
          vin[blocks[j].off + i] = i * 7u + (u32)j;
          cnt[(k >> shift) & 255]++;
        }
        u32 run = 0;
        for (int d = 0; d < 256; d++) { jobhist[(j * ST_MAXPASS + (shift >> 3)) * 256 + d] = run; run += cnt[d]; }
      }
      u32 lb_error = 0;
      const u32 tag = (pass_no & 255u) << 22;
      pass_no++;
      if constexpr (MODE >= 40) {
        const u32 pfd = (shift & 8) ? 3u : 0u;                // with and without the prefetch of a later tile (a no-op here but for its addressing)
        emu_launch((unsigned)sc.size(), THREADS3, sizeof(ScatterSmemT<THREADS3>), window, [&]() {
          k_scatter3<THREADS3, 3, ((MODE - 40) >> 2)>(sc.data(), kin.data(), vin.data(), kout.data(), vout.data(), shift, state.data(), jobhist.data(), &lb_error, tag, pfd);
        });
      } else {
        emu_launch((unsigned)rr.size(), SC_THREADS, sizeof(ScatterSmem), window, [&]() {
          k_scatter2<3, MODE>(rr.data(), jobs.data(), kin.data(), vin.data(), kout.data(), vout.data(), shift, state.data(), jobhist.data(), &lb_error, tag);
        });
      }
      if (lb_error) { printf("look-back error flag set (shift %d)\n", shift); bad++; }
      // reference: stable sort by digit per block; untouched elsewhere
      std::vector<u64> kref(total, 0x1111111111111111ull);
      std::vector<u32> vref(total, 0x22222222u);
      for (size_t j = 0; j < blocks.size(); j++) {
        std::vector<u32> idx(blocks[j].n);
        for (u32 i = 0; i < blocks[j].n; i++) idx[i] = i;
        std::stable_sort(idx.begin(), idx.end(), [&](u32 a, u32 b) { return ((kin[blocks[j].off + a] >> shift) & 255) < ((kin[blocks[j].off + b] >> shift) & 255); });
        for (u32 i = 0; i < blocks[j].n; i++) { kref[blocks[j].off + i] = kin[blocks[j].off + idx[i]]; vref[blocks[j].off + i] = vin[blocks[j].off + idx[i]]; }
      }
      size_t diffs = 0;
      for (u32 i = 0; i < total; i++) if (kref[i] != kout[i] || vref[i] != vout[i]) { if (diffs < 5) printf("mode %d shift %d dist %d: position %u differs\n", MODE, shift, dist, i); diffs++; }
      if (diffs) { printf("mode %d shift %d dist %d: %zu positions differ\n", MODE, shift, dist, diffs); bad++; }
    }
  }
  return bad;
}

int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const size_t group = argc > 2 ? (size_t)atol(argv[2]) : 128;
  const unsigned window = argc > 3 ? (unsigned)atoi(argv[3]) : 1;
  const unsigned seed = argc > 4 ? (unsigned)atoi(argv[4]) : 1;
  const int step = argc > 5 ? atoi(argv[5]) : 8;          // 8: every pass position; 24: passes 0, 3, 6
  int bad = 0;
  switch (mode) {
    case 0: bad = run<0>(group, window, seed, step); break;
    case 1: bad = run<SC2_MATCHANY>(group, window, seed, step); break;
    case 2: bad = run<SC2_EARLY>(group, window, seed, step); break;
    case 40: bad = run<40>(group, window, seed, step); break;
    case 42: bad = run<42>(group, window, seed, step); break;
    case 44: bad = run<44>(group, window, seed, step); break;
    case 46: bad = run<46>(group, window, seed, step); break;
    case 52: bad = run<52>(group, window, seed, step); break;
    case 60: bad = run<60>(group, window, seed, step); break;
    case 68: bad = run<68>(group, window, seed, step); break;
    case 70: bad = run<70>(group, window, seed, step); break;
    default: bad = run<SC2_EARLY | SC2_MATCHANY>(group, window, seed, step); break;
  }
  printf(bad ? "FAILED\n" : "OK\n");
  return bad ? 1 : 0;
}
