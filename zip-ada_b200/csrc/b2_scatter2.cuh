// The radix scatter of the BWT rotation sort (stage A6, zip_lib/bzip2-encoding.adb:219-296 is what the sort
// replaces): tile geometry, shared-memory layout, look-back states and the kernel k_scatter2.
// Included by b2_bwt.cu; also compiled for the host by tests/emu (a thread-per-lane emulation of one CTA
// that checks the kernel against a plain stable counting sort - `-m "not gpu"`), hence the few B2_EMU switches.
#pragma once
#include "b2_common.cuh"
#include "b2_kernels.h"

#ifndef ST_MAXPASS
#define ST_MAXPASS 8
#endif
// The scatter has its own tile geometry (env-free compile-time choice): larger tiles give longer output
// runs per digit and fewer look-back states.
#ifndef SC_THREADS
#define SC_THREADS 512
#endif
#define SC_ITEMS 8
#define SC_TILE (SC_THREADS * SC_ITEMS)
#define SC_WARPS (SC_THREADS / 32)
#define SC_WCHUNK (32 * SC_ITEMS)
#define SC_MINCTAS (1536 / SC_THREADS)
// ---- stable scatter of one digit ---------------------------------------------------------------
struct ScatterSmem {
  u64 keys[SC_TILE];
  u32 vals[SC_TILE];
  u32 warp_cnt[SC_WARPS][256];
  u32 tile_start[256];
  u32 g_off[256];
  u32 scan[40];
};

// Tile states of the decoupled look-back: [31:30] 1 = this tile's count, 2 = count of this tile and all
// tiles of the block before it; [29:22] pass tag (states of other passes read as "not there yet");
// [21:0] the count (a block has at most 1 125 000 rows).
#define LB_AGG 0x40000000u
#define LB_INC 0x80000000u
// tile states are single self-describing words: relaxed device-scope accesses are all that is needed
#ifndef B2_EMU
__device__ __forceinline__ void lb_store(u32 *p, u32 v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ u32 lb_load(const u32 *p) { u32 v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
#else
inline void lb_store(u32 *p, u32 v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
inline u32 lb_load(const u32 *p) { emu_yield(); return __atomic_load_n(p, __ATOMIC_RELAXED); }
#endif

// ---- the same pass with half the instructions per row (round 2) ---------------------------------------
// k_scatter spends 158 instructions per row (profiles/r02_scatter_ncu_full_summary.md) for 24 bytes of
// traffic: it is bound by instruction issue, not by HBM.  k_scatter2 is the same algorithm - same tile
// geometry, same look-back protocol and states, same output - written for the instruction count:
//  * a tile with all its 4096 rows (all but the last tile of a block) runs a body without any bounds test;
//  * the match of equal digits inside a warp is three instructions per digit bit (vote, select, one three-input
//    logic operation; the predicates of seven bits come from one R2P) instead of the six the compiler makes
//    of the C form;
//  * digits and ranks stay in their own registers (no packing / unpacking), every global and shared address
//    is one per-thread base plus a compile-time offset.
__device__ __forceinline__ u32 sc2_match_bit(u32 peers, u32 d, const u32 bitmask) {
#ifdef B2_EMU
  const u32 m = __ballot_sync(0xffffffffu, (d & bitmask) != 0u), x = (d & bitmask) ? 0xffffffffu : 0u;
  return peers & ~(m ^ x);
#else
  // peers &= (my bit set ? lanes with the bit set : lanes with the bit clear) = peers & ~(ballot ^ x), x = my bit on all 32 places
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 m, t, x;\n\t"
               "and.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
               "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
               "selp.b32 x, 0xffffffff, 0, p;\n\t"
               "lop3.b32 %0, %0, m, x, 0x90;\n\t}"
               : "+r"(peers) : "r"(d), "r"(bitmask));
  return peers;
#endif
}

// MODE bit 0: lanes with equal digits from match.any instead of the eight ballots; bit 1: the whole keys and the
// rotation indices are loaded at the top and kept in registers (no second read; for two CTAs of 64 registers).
#define SC2_MATCHANY 1
#define SC2_EARLY 2
template <bool FULL, int MODE>
__device__ __forceinline__ void sc2_body(ScatterSmem &S, const B2SortTileRR tl, const B2SortTileRR *__restrict__ tiles,
                                         const u32 n, const u32 off, const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in,
                                         u64 *__restrict__ keys_out, u32 *__restrict__ vals_out, const int shift, u32 *__restrict__ state,
                                         u32 *__restrict__ lb_error, const u32 tag, const u32 v_early) {
  const u32 tid = threadIdx.x, w = tid >> 5, l = tid & 31u;
  const u32 tix = blockIdx.x;
  const u32 lt_mask = (1u << l) - 1u;
  const u32 cnt = FULL ? (u32)SC_TILE : (n - tl.start);          // rows of this tile
  const u32 wrow = w * SC_WCHUNK + l;                            // my first row inside the tile; row k is wrow + 32 k
  const u64 *kin = keys_in + off + tl.start + wrow;
  const u32 *vin = vals_in + off + tl.start + wrow;
  u32 *wc = &S.warp_cnt[w][0];
  u32 d[SC_ITEMS], rk[SC_ITEMS];
  u64 key[SC_ITEMS];
  u32 val[SC_ITEMS];
  if (MODE & SC2_EARLY) {
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
      const bool valid = FULL || (wrow + 32u * k < cnt);
      key[k] = valid ? kin[k * 32] : 0;
      val[k] = valid ? vin[k * 32] : 0;
    }
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) d[k] = (u32)(key[k] >> shift) & 255u;
  } else {
    const u8 *kb = reinterpret_cast<const u8 *>(kin) + ((u32)shift >> 3);        // little endian: byte shift / 8 of the key
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) d[k] = (FULL || wrow + 32u * k < cnt) ? (u32)kb[(size_t)k * 256] : 0u;
  }
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    const bool valid = FULL || (wrow + 32u * k < cnt);
    u32 peers;
    if (MODE & SC2_MATCHANY) {
      peers = __match_any_sync(0xffffffffu, valid ? d[k] : 256u);            // rows beyond the end only match each other
    } else {
      peers = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, valid);
#pragma unroll
      for (int bb = 0; bb < 8; bb++) peers = sc2_match_bit(peers, d[k], 1u << bb);
    }
    const u32 r = __popc(peers & lt_mask);
    const u32 base = wc[d[k]];
    __syncwarp();
    if (valid && r == 0) wc[d[k]] = base + __popc(peers);
    __syncwarp();
    rk[k] = base + r;                                           // < 256: a warp holds 256 rows
  }
  __syncthreads();
  // per digit: exclusive over warps, tile count; then the look-back (protocol and states of k_scatter).  The
  // 256 digits belong to the threads of warps 0-7: the scan over the digits is a warp scan plus the sums of the
  // warps before (one barrier instead of the three of block_excl_add).
  u32 run = 0, inc = 0;
  if (tid < 256) {
#pragma unroll
    for (int ww = 0; ww < SC_WARPS; ww++) { const u32 c = S.warp_cnt[ww][tid]; S.warp_cnt[ww][tid] = run; run += c; }
    inc = warp_incl_add(run);
    if (l == 31) S.scan[w] = inc;
  }
  __syncthreads();
  if (tid < 256) {
    u32 ts = inc - run;
#pragma unroll
    for (u32 ww = 0; ww < 7; ww++) ts += (ww < w) ? S.scan[ww] : 0u;
    S.tile_start[tid] = ts;
    u32 *stt = state + (size_t)tix * 256 + tid;
    u32 excl = 0;
    if (tl.prev == 0xFFFFFFFFu) {
      lb_store(stt, LB_INC | tag | run);
    } else {
      if ((v_early & 0xFFC00000u) == (LB_INC | tag)) {
        excl = v_early & 0x3FFFFFu;                       // the usual case: no wait, no AGG state needed
      } else {
        lb_store(stt, LB_AGG | tag | run);
        u32 p = tl.prev;
        while (p != 0xFFFFFFFFu) {
          const u32 *pp = state + (size_t)p * 256 + tid;
          u32 v, spins = 0;
          do { v = lb_load(pp); } while (((v & 0x3FC00000u) != tag || (v >> 30) == 0u) && ++spins < (1u << 24));
          if (spins >= (1u << 24)) { *lb_error = 1u; break; }
          excl += v & 0x3FFFFFu;
          if (v & LB_INC) break;
          p = tiles[p].prev;
        }
      }
      lb_store(stt, LB_INC | tag | (excl + run));
    }
    // arena index of the first row of my digit in this tile, minus its place in the staged tile (the arenas
    // have fewer than 2^32 positions; the differences are taken modulo 2^32)
    S.g_off[tid] += off + excl - ts;
  }
  __syncthreads();
  // place of every row in the staged tile, then keys and rotation indices (second read of the keys: L2)
  u32 lp[SC_ITEMS];
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) lp[k] = S.tile_start[d[k]] + wc[d[k]] + rk[k];
  if (!(MODE & SC2_EARLY)) {
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
      const bool valid = FULL || (wrow + 32u * k < cnt);
      key[k] = valid ? kin[k * 32] : 0;
      val[k] = valid ? vin[k * 32] : 0;
    }
  }
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    if (FULL || (wrow + 32u * k < cnt)) { S.keys[lp[k]] = key[k]; S.vals[lp[k]] = val[k]; }
  }
  __syncthreads();
  const u8 *sdig = reinterpret_cast<const u8 *>(&S.keys[tid]) + ((u32)shift >> 3);
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    const u32 q = tid + k * SC_THREADS;
    if (FULL || q < cnt) {
      const u64 kk = S.keys[q];
      const u32 dg = sdig[(size_t)k * SC_THREADS * 8];
      const u32 dst = S.g_off[dg] + q;
      keys_out[dst] = kk;
      vals_out[dst] = S.vals[q];
    }
  }
}

template <int MINCTAS, int MODE>
__global__ void __launch_bounds__(SC_THREADS, MINCTAS)
k_scatter2(const B2SortTileRR *__restrict__ tiles, const B2Job *__restrict__ jobs,
           const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in,
           u64 *__restrict__ keys_out, u32 *__restrict__ vals_out, int shift, u32 *__restrict__ state,
           const u32 *__restrict__ jobhist, u32 *__restrict__ lb_error, u32 tag) {
#ifndef B2_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#else
  unsigned char *smem_raw = emu_dynamic_smem();
#endif
  ScatterSmem &S = *reinterpret_cast<ScatterSmem *>(smem_raw);
  const u32 tid = threadIdx.x;
  const B2SortTileRR tl = tiles[blockIdx.x];
  {
    // every warp clears its own 256 counters: no block barrier before the ranking
    uint4 *z = reinterpret_cast<uint4 *>(&S.warp_cnt[tid >> 5][0]);
    z[tid & 31u] = make_uint4(0u, 0u, 0u, 0u);
    z[(tid & 31u) + 32] = make_uint4(0u, 0u, 0u, 0u);
  }
  const u32 n = jobs[tl.job].na, off = jobs[tl.job].pos_off;
  u32 v_early = LB_INC | tag;
  if (tid < 256) {
    // (only thread tid touches g_off[tid] before the barriers of the body)
    S.g_off[tid] = jobhist[((size_t)tl.job * ST_MAXPASS + (u32)(shift >> 3)) * 256 + tid];
    if (tl.prev != 0xFFFFFFFFu) v_early = lb_load(state + (size_t)tl.prev * 256 + tid);   // only trusted when final
  }
  __syncwarp();
  if (n - tl.start >= SC_TILE) sc2_body<true, MODE>(S, tl, tiles, n, off, keys_in, vals_in, keys_out, vals_out, shift, state, lb_error, tag, v_early);
  else sc2_body<false, MODE>(S, tl, tiles, n, off, keys_in, vals_in, keys_out, vals_out, shift, state, lb_error, tag, v_early);
}

// ---- k_scatter3: the same pass, written against latency -------------------------------------------------------
// Halving the instructions (k_scatter2) bought 7 %: the pass is not bound by issue but by the chain of dependent
// memory waits of a tile (tile record -> block record -> digit bytes from DRAM -> ... -> keys and rotation indices)
// with only three tiles resident per SM: about 32 KB of loads in flight per SM, where the HBM needs 45.  Here
//  * the tile record carries everything the tile needs (first row in the arena, rows, block offset): one load
//    instead of two dependent ones before the first data load can be issued;
//  * every tile asks the L2 for the keys and rotation indices of the tile `pf_dist` places behind it in the dispatch
//    order (prefetch.global.L2, one 128-byte line per thread): when that tile starts, its loads are L2 hits, and the
//    DRAM reads of a tile are in flight one tile lifetime before they are needed;
//  * THREADS = 256 makes tiles of 2048 rows, six resident per SM instead of three (finer interleaving of the phases);
//  * flags (SC3_*): rotation indices requested before the scan / look-back phase; digit of the output phase from the key
//    registers; a second early look at the predecessor's look-back state.
struct B2ScTile { u32 jobcnt; u32 row0; u32 prev; u32 off; };   // jobcnt = block | (rows - 1) << 16; row0 = arena index of the first row

template <int THREADS> struct ScatterSmemT {
  u64 keys[THREADS * SC_ITEMS];
  u32 vals[THREADS * SC_ITEMS];
  u32 warp_cnt[THREADS / 32][256];
  u32 tile_start[256];
  u32 g_off[256];
  u32 scan[40];
};

__device__ __forceinline__ void sc_prefetch_l2(const void *p) {
#ifndef B2_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

#define SC3_VEARLY 1      // rotation indices requested before the scan / look-back phase
#define SC3_LEAN 2        // output phase: the digit comes out of the key registers instead of another shared-memory load
#define SC3_MIDLOOK 4     // a second look at the predecessor's state right after the ranking (the first one, at the start, is too early
                          // when the dispatch groups are small)
template <bool FULL, int THREADS, int FLAGS>
__device__ __forceinline__ void sc3_body(ScatterSmemT<THREADS> &S, const u32 cnt, const u32 row0, const u32 prev, const u32 off,
                                         const B2ScTile *__restrict__ tiles, const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in,
                                         u64 *__restrict__ keys_out, u32 *__restrict__ vals_out, const int shift, u32 *__restrict__ state,
                                         u32 *__restrict__ lb_error, const u32 tag, u32 v_early, const bool pf, const uint4 rec2) {
  constexpr int TILE = THREADS * SC_ITEMS, WARPS = THREADS / 32;
  constexpr bool VEARLY = (FLAGS & SC3_VEARLY) != 0, LEAN = (FLAGS & SC3_LEAN) != 0, MIDLOOK = (FLAGS & SC3_MIDLOOK) != 0;
  const u32 tid = threadIdx.x, w = tid >> 5, l = tid & 31u;
  const u32 tix = blockIdx.x;
  const u32 lt_mask = (1u << l) - 1u;
  const u32 wrow = w * SC_WCHUNK + l;                            // my first row inside the tile; row k is wrow + 32 k
  const u64 *kin = keys_in + row0 + wrow;
  const u32 *vin = vals_in + row0 + wrow;
  u32 *wc = &S.warp_cnt[w][0];
  u32 d[SC_ITEMS], rk[SC_ITEMS];
  {
    const u8 *kb = reinterpret_cast<const u8 *>(kin) + ((u32)shift >> 3);        // little endian: byte shift / 8 of the key
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) d[k] = (FULL || wrow + 32u * k < cnt) ? (u32)kb[(size_t)k * 256] : 0u;
  }
  if (pf) {
    // the tile pf_dist places behind me: its keys (16 per line) and rotation indices (32 per line)
    const u32 cnt2 = (rec2.x >> 16) + 1u, row2 = rec2.y;
    if (tid < (u32)TILE / 16) { if (tid * 16u < cnt2) sc_prefetch_l2(keys_in + row2 + tid * 16u); }
    else { const u32 t = tid - (u32)TILE / 16; if (t * 32u < cnt2) sc_prefetch_l2(vals_in + row2 + t * 32u); }
  }
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    const bool valid = FULL || (wrow + 32u * k < cnt);
    u32 peers = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int bb = 0; bb < 8; bb++) peers = sc2_match_bit(peers, d[k], 1u << bb);
    const u32 r = __popc(peers & lt_mask);
    const u32 base = wc[d[k]];
    __syncwarp();
    if (valid && r == 0) wc[d[k]] = base + __popc(peers);
    __syncwarp();
    rk[k] = base + r;                                           // < 256: a warp holds 256 rows
  }
  if (MIDLOOK) {
    // the tile before mine was dispatched one group of blocks earlier: by now it has usually published its inclusive count
    if (tid < 256 && prev != 0xFFFFFFFFu && (v_early & 0xFFC00000u) != (LB_INC | tag)) v_early = lb_load(state + (size_t)prev * 256 + tid);
  }
  // VEARLY: the rotation indices are requested before the scan and the look-back, whose time hides their latency
  u32 val[SC_ITEMS];
  if (VEARLY) {
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) val[k] = (FULL || (wrow + 32u * k < cnt)) ? vin[k * 32] : 0u;
  }
  __syncthreads();
  u32 run = 0, inc = 0;
  if (tid < 256) {
#pragma unroll
    for (int ww = 0; ww < WARPS; ww++) { const u32 c = S.warp_cnt[ww][tid]; S.warp_cnt[ww][tid] = run; run += c; }
    inc = warp_incl_add(run);
    if (l == 31) S.scan[w] = inc;
  }
  __syncthreads();
  if (tid < 256) {
    u32 ts = inc - run;
#pragma unroll
    for (u32 ww = 0; ww < 7; ww++) ts += (ww < w) ? S.scan[ww] : 0u;
    S.tile_start[tid] = ts;
    u32 *stt = state + (size_t)tix * 256 + tid;
    u32 excl = 0;
    if (prev == 0xFFFFFFFFu) {
      lb_store(stt, LB_INC | tag | run);
    } else {
      if ((v_early & 0xFFC00000u) == (LB_INC | tag)) {
        excl = v_early & 0x3FFFFFu;                       // the usual case: no wait, no AGG state needed
      } else {
        lb_store(stt, LB_AGG | tag | run);
        u32 p = prev;
        while (p != 0xFFFFFFFFu) {
          const u32 *pp = state + (size_t)p * 256 + tid;
          u32 v, spins = 0;
          do { v = lb_load(pp); } while (((v & 0x3FC00000u) != tag || (v >> 30) == 0u) && ++spins < (1u << 24));
          if (spins >= (1u << 24)) { *lb_error = 1u; break; }
          excl += v & 0x3FFFFFu;
          if (v & LB_INC) break;
          p = tiles[p].prev;
        }
      }
      lb_store(stt, LB_INC | tag | (excl + run));
    }
    S.g_off[tid] += off + excl - ts;                     // arena index of my digit's first row of this tile, minus its staged place (mod 2^32)
  }
  __syncthreads();
  u32 lp[SC_ITEMS];
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) lp[k] = S.tile_start[d[k]] + wc[d[k]] + rk[k];
  {
    u64 key[SC_ITEMS];
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
      const bool valid = FULL || (wrow + 32u * k < cnt);
      key[k] = valid ? kin[k * 32] : 0;
      if (!VEARLY) val[k] = valid ? vin[k * 32] : 0;
    }
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
      if (FULL || (wrow + 32u * k < cnt)) { S.keys[lp[k]] = key[k]; S.vals[lp[k]] = val[k]; }
    }
  }
  __syncthreads();
  const u8 *sdig = reinterpret_cast<const u8 *>(&S.keys[tid]) + ((u32)shift >> 3);
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    const u32 q = tid + k * THREADS;
    if (FULL || q < cnt) {
      const u64 kk = S.keys[q];
      const u32 dg = LEAN ? ((u32)(kk >> shift) & 255u) : (u32)sdig[(size_t)k * THREADS * 8];
      const u32 dst = S.g_off[dg] + q;
      keys_out[dst] = kk;
      vals_out[dst] = S.vals[q];
    }
  }
}

template <int THREADS, int MINCTAS, int FLAGS>
__global__ void __launch_bounds__(THREADS, MINCTAS)
k_scatter3(const B2ScTile *__restrict__ tiles, const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in,
           u64 *__restrict__ keys_out, u32 *__restrict__ vals_out, int shift, u32 *__restrict__ state,
           const u32 *__restrict__ jobhist, u32 *__restrict__ lb_error, u32 tag, u32 pf_dist) {
#ifndef B2_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#else
  unsigned char *smem_raw = emu_dynamic_smem();
#endif
  ScatterSmemT<THREADS> &S = *reinterpret_cast<ScatterSmemT<THREADS> *>(smem_raw);
  constexpr u32 TILE = THREADS * SC_ITEMS;
  const u32 tid = threadIdx.x;
  const uint4 rec = *reinterpret_cast<const uint4 *>(tiles + blockIdx.x);
  const bool pf = pf_dist != 0u && blockIdx.x + pf_dist < gridDim.x && tid < TILE / 16 + TILE / 32;
  uint4 rec2 = make_uint4(0u, 0u, 0u, 0u);
  if (pf) rec2 = *reinterpret_cast<const uint4 *>(tiles + blockIdx.x + pf_dist);
  {
    // every warp clears its own 256 counters: no block barrier before the ranking
    uint4 *z = reinterpret_cast<uint4 *>(&S.warp_cnt[tid >> 5][0]);
    z[tid & 31u] = make_uint4(0u, 0u, 0u, 0u);
    z[(tid & 31u) + 32] = make_uint4(0u, 0u, 0u, 0u);
  }
  const u32 job = rec.x & 0xFFFFu, cnt = (rec.x >> 16) + 1u, row0 = rec.y, prev = rec.z, off = rec.w;
  u32 v_early = LB_INC | tag;
  if (tid < 256) {
    S.g_off[tid] = jobhist[((size_t)job * ST_MAXPASS + (u32)(shift >> 3)) * 256 + tid];
    if (prev != 0xFFFFFFFFu) v_early = lb_load(state + (size_t)prev * 256 + tid);   // only trusted when final
  }
  __syncwarp();
  if (cnt == TILE) sc3_body<true, THREADS, FLAGS>(S, cnt, row0, prev, off, tiles, keys_in, vals_in, keys_out, vals_out, shift, state, lb_error, tag, v_early, pf, rec2);
  else sc3_body<false, THREADS, FLAGS>(S, cnt, row0, prev, off, tiles, keys_in, vals_in, keys_out, vals_out, shift, state, lb_error, tag, v_early, pf, rec2);
}
