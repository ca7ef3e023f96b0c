// Length-limited code lengths by package-merge, one warp per coder (stage A9).  Included by b2_entropy.cu; also
// compiled for the host by tests/emu (thread-per-lane emulation, `-m "not gpu"`).
#pragma once
#include "b2_common.cuh"

// ---------------------------------------------------------------------------------------------
// Length-limited code lengths (huffman-encoding-length_limited_coding.adb:46-280).
//
// The reference runs the *boundary* package-merge (lazy, recursive, :131-163).  What it computes is
// the classic package-merge: list 1 = the sorted leaves; list l+1 = merge (leaves, pair sums of
// consecutive items of list l); take the first 2n-2 items of the last list, then walk down: if p of
// the first k items of a list are packages, the first 2p items of the list below are taken; a leaf
// gets one bit per list in which it is taken (Extract_Bit_Lengths, :180-189).  The boundary
// version's tie rule "new leaf iff sum > leaf weight" (:151) means a package goes BEFORE a leaf of
// equal weight.  Each list is built here by one warp as a parallel merge (binary searches); the
// equality of both formulations is checked on the CPU against the oracle's literal restatement
// (tests/test_oracle.py::test_forward_package_merge_equals_boundary).
// ---------------------------------------------------------------------------------------------
#define LL_MAXBITS 17
#define LL_MAXITEMS (2 * B2_MAX_ALPHA)
#define LL_BITWORDS 17
#define PM_WARPS 4

// Scratch of one warp, carved out of dynamic shared memory sized by the largest alphabet of the batch (a text
// batch needs 4.4 KB per warp instead of 7.3 KB: more warps per SM for a latency-bound kernel).
struct LLScratch {
  u32 *leaf;                               // [alpha + 2]     (weight << 9) | symbol, sorted
  u32 *lvl[2];                             // [2 * alpha]     merged weights of two consecutive lists
  u32 *pk;                                 // [alpha + 2]     pair sums of the list below
  u32 (*pkgbits)[LL_BITWORDS];             // [LL_MAXBITS]    bit p set <=> item p of the list is a package
};
__host__ __device__ inline u32 ll_scratch_words(u32 alpha) { return (alpha + 2) * 2 + 4 * alpha + LL_MAXBITS * LL_BITWORDS; }
__device__ __forceinline__ LLScratch ll_scratch_at(u32 *base, u32 alpha) {
  LLScratch S;
  S.leaf = base; base += alpha + 2;
  S.lvl[0] = base; base += 2 * alpha;
  S.lvl[1] = base; base += 2 * alpha;
  S.pk = base; base += alpha + 2;
  S.pkgbits = reinterpret_cast<u32 (*)[LL_BITWORDS]>(base);
  return S;
}

// walk down from the last list (Extract_Bit_Lengths): if p of the first k items of a list are packages, the
// first 2 p items of the list below are taken; a leaf gets one bit per list in which it is taken
__device__ __forceinline__ void ll_walk_down(LLScratch &S, int ns, int max_bits, u8 *lens) {
  const u32 l = lane_id();
  const int need = 2 * ns - 2;
  u32 cnt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int k = need;
  for (int lev = max_bits - 1; lev >= 0; lev--) {
    const u32 wv = (l < LL_BITWORDS) ? S.pkgbits[lev][l] : 0u;
    const int lo = (int)l * 32;
    const u32 msk = (k >= lo + 32) ? 0xFFFFFFFFu : (k <= lo ? 0u : ((1u << (k - lo)) - 1u));
    u32 p = __popc(wv & msk);
#pragma unroll
    for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    const int a = k - (int)p;                               // leaves taken in this list
#pragma unroll
    for (int j = 0; j < 9; j++) cnt[j] += ((int)l + 32 * j < a);
    k = 2 * (int)p;
  }
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const int i = (int)l + 32 * j;
    if (i < ns) lens[S.leaf[i] & 511u] = (u8)cnt[j];
  }
}


__device__ void ll_package_merge_warp(LLScratch &S, int ns, int max_bits, u8 *lens) {
  const u32 l = lane_id();
  const int need = 2 * ns - 2;
  for (int i = l; i < ns; i += 32) S.lvl[0][i] = S.leaf[i] >> 9;
  if (l < LL_BITWORDS) S.pkgbits[0][l] = 0;
  int len_prev = ns;
  __syncwarp();
  for (int lev = 1; lev < max_bits; lev++) {
    const u32 *prev = S.lvl[(lev - 1) & 1];
    u32 *cur = S.lvl[lev & 1];
    const int npk = len_prev >> 1;
    if (l < LL_BITWORDS) S.pkgbits[lev][l] = 0;
    for (int b = l; b < npk; b += 32) S.pk[b] = prev[2 * b] + prev[2 * b + 1];      // the packages of the list below
    __syncwarp();
    bool changed = false;
    for (int a = l; a < ns; a += 32) {
      const u32 w = S.leaf[a] >> 9;
      int lo = 0, hi = npk;                                 // packages with sum <= w go before this leaf
      while (lo < hi) { int mid = (lo + hi) >> 1; if (S.pk[mid] <= w) lo = mid + 1; else hi = mid; }
      const int pos = a + lo;
      if (pos < need) { changed |= (pos >= len_prev) || prev[pos] != w; cur[pos] = w; }
    }
    for (int b = l; b < npk; b += 32) {
      const u32 pk = S.pk[b];
      int lo = 0, hi = ns;                                  // leaves with weight < sum go before this package
      while (lo < hi) { int mid = (lo + hi) >> 1; if ((S.leaf[mid] >> 9) < pk) lo = mid + 1; else hi = mid; }
      const int pos = b + lo;
      if (pos < need) { changed |= (pos >= len_prev) || prev[pos] != pk; cur[pos] = pk; atomicOr(&S.pkgbits[lev][pos >> 5], 1u << (pos & 31)); }
    }
    const int len_cur = min(need, ns + npk);
    changed = __any_sync(0xffffffffu, changed) || len_cur != len_prev;
    len_prev = len_cur;
    __syncwarp();
    if (!changed) {
      // the list repeats the one below: every list above is built from the same packages, hence equal
      // to this one, items and package flags alike
      if (l < LL_BITWORDS) { const u32 v = S.pkgbits[lev][l]; for (int q = lev + 1; q < max_bits; q++) S.pkgbits[q][l] = v; }
      __syncwarp();
      break;
    }
  }
  ll_walk_down(S, ns, max_bits, lens);
}

// The same lists, each built as ONE merge instead of a binary search per item: lane l makes the items
// [l * ipl, (l + 1) * ipl) of the list.  Where its stretch starts in the two inputs (how many of the items before
// it are packages) is one binary search on the diagonal - package b goes to place b + (leaves lighter than it),
// which is >= D exactly when leaf D - b - 1 is lighter than package b - and then it merges sequentially, a
// package before a leaf of equal weight.  About five times fewer instructions per list than the searches.
__device__ void ll_package_merge_warp_mp(LLScratch &S, int ns, int max_bits, u8 *lens) {
  const u32 l = lane_id();
  const int need = 2 * ns - 2;
  u32 *const lvl = S.lvl[0];                                   // both lists through one base: list k at k * stride
  const u32 stride = (u32)(S.lvl[1] - S.lvl[0]);
  const u32 *const leaf = S.leaf;
  u32 *const pk = S.pk;
  for (int i = l; i < ns; i += 32) lvl[i] = leaf[i] >> 9;
  if (l < LL_BITWORDS) S.pkgbits[0][l] = 0;
  int len_prev = ns;
  __syncwarp();
  for (int lev = 1; lev < max_bits; lev++) {
    const u32 *prev = lvl + ((lev - 1) & 1) * stride;
    u32 *cur = lvl + (lev & 1) * stride;
    const int npk = len_prev >> 1;
    if (l < LL_BITWORDS) S.pkgbits[lev][l] = 0;
    for (int b = l; b < npk; b += 32) pk[b] = prev[2 * b] + prev[2 * b + 1];      // the packages of the list below
    __syncwarp();
    const int len_cur = min(need, ns + npk);
    const int ipl = (len_cur + 31) >> 5;                     // at most 17: my stretch lies in at most two words of flags
    const int D = (int)l * ipl;
    bool changed = false;
    if (D < len_cur) {
      int lo = max(0, D - ns), hi = min(D, npk);              // packages among the first D items
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((leaf[D - mid - 1] >> 9) < pk[mid]) hi = mid; else lo = mid + 1;
      }
      int bi = lo, ai = D - lo;
      const int end = min(len_cur, D + ipl);
      // heads of both inputs; one place behind an input is still inside the scratch (never taken: see takep)
      u32 lv = leaf[ai] >> 9, pv = pk[bi];
      u32 tb = 0, m = 1;                                        // flags of my stretch, first item in bit 0
      for (int pos = D; pos < end; pos++) {
        const bool takep = bi < npk && (ai >= ns || pv <= lv);
        const u32 v = takep ? pv : lv;
        changed |= (pos >= len_prev) || prev[pos] != v;
        cur[pos] = v;
        tb |= takep ? m : 0u;
        m <<= 1;
        bi += takep ? 1 : 0;
        ai += takep ? 0 : 1;
        const u32 nxt = takep ? pk[bi] : leaf[ai];
        pv = takep ? nxt : pv;
        lv = takep ? lv : (nxt >> 9);
      }
      const u32 sh = (u32)D & 31u;
      const u32 bits0 = tb << sh, bits1 = sh ? (tb >> (32u - sh)) : 0u;
      if (bits0) atomicOr(&S.pkgbits[lev][D >> 5], bits0);
      if (bits1) atomicOr(&S.pkgbits[lev][(D >> 5) + 1], bits1);
    }
    changed = __any_sync(0xffffffffu, changed) || len_cur != len_prev;
    len_prev = len_cur;
    __syncwarp();
    if (!changed) {
      // the list repeats the one below: every list above is built from the same packages, hence equal
      // to this one, items and package flags alike
      if (l < LL_BITWORDS) { const u32 v = S.pkgbits[lev][l]; for (int q = lev + 1; q < max_bits; q++) S.pkgbits[q][l] = v; }
      __syncwarp();
      break;
    }
  }
  ll_walk_down(S, ns, max_bits, lens);
}
