// Stage A6: Burrows-Wheeler rotation sort for a batch of blocks.
//
// Reference: zip_lib/bzip2-encoding.adb:219-296.  The reference heap-sorts rotation offsets with
// an O(N) comparator that is a total order (:229-255), so the sorted matrix, its last column and
// the origin pointer are unique functions of the block; any sort is admissible.
//
// Here: cyclic prefix doubling with filtering.  Round 0 sorts every rotation by its first 8 bytes,
// each byte replaced by its rank among the bytes in use and packed with b = ceil (log2 (largest
// alphabet of the batch)) bits: b LSD radix passes of 8 bits (5 for lower-case text) - or, when the
// alphabet has more than 128 bytes, by seven characters in radix B and the eighth cut down to
// floor (2^56 / B^7) order-preserving buckets: seven passes, and the doubling goes on from 7.  A
// rank is the row index ("slot") of the first row of a class, so equal prefixes share a rank.  After
// every round the rows that are alone in their class are final; only the others ("active" rows) are
// compacted and re-sorted in round r >= 1 by the 40-bit key (rank[i] << 20 | rank[(i+h) mod n]) with 5
// passes, h = 8, 16, ..., and written back to the slots they came from (a class occupies a contiguous
// range of slots, and the key's high part keeps classes in place).  A block is finished when no row is
// active or when the compared prefix covers the whole rotation (periodic blocks: the classes are then
// sets of EQUAL rotations; their last-column bytes coincide and the origin pointer is the first row of
// the class of rotation 0, because the reference orders equal rotations by offset and rotation 0 has
// offset 0, :254, :276-279).  All blocks of a batch are sorted together: every radix pass is segmented
// by block through a tile table, so no block id is needed in the key.
//
// Radix pass = ONE kernel (k_scatter3 in b2_scatter2.cuh; k_scatter below and k_scatter2 are its earlier
// versions, kept switchable by B2GPU_SCATTER): 12 B/row read, 12 B/row written, staged through shared
// memory so that the writes leave as runs.  The digit totals of all passes of a round come from one read
// of the keys (k_hist_all: LSD passes only permute the rows of a block), the per-tile offsets from a
// decoupled look-back inside the scatter (see k_scatter).
#include "b2_common.cuh"
#include "b2_kernels.h"

#define ST_THREADS 256
#define ST_ITEMS 8
#define ST_TILE (ST_THREADS * ST_ITEMS)
#define ST_WARPS (ST_THREADS / 32)
#define ST_WCHUNK (32 * ST_ITEMS)      // elements per warp

// ---- round-0 keys: first 8 bytes of each rotation, cyclic ------------------------------------
// The bytes are replaced by their rank among the bytes in use (order preserving, so the order of the
// rotations is unchanged) and packed with `bits` bits each: a batch whose blocks all use at most 2^bits
// distinct bytes sorts round 0 in `bits` radix passes instead of 8.  bits = 8 keeps the bytes as they are.
__global__ void __launch_bounds__(ST_THREADS)
k_keys0(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs, const u8 *__restrict__ text,
        u64 *__restrict__ keys, u32 *__restrict__ vals, int bits, u32 base, u32 q8) {
  __shared__ u8 code[256];
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.n, off = job.pos_off;
  const u8 *tx = text + off;
  {
    const u32 t = threadIdx.x, wd = t >> 5;
    u32 c = t;
    if (bits < 8 || base) {
      c = __popc(job.in_use[wd] & ((1u << (t & 31u)) - 1u));
      for (u32 q = 0; q < wd; q++) c += __popc(job.in_use[q]);
    }
    code[t] = (u8)c;
  }
  __syncthreads();
#pragma unroll 4
  for (int k = 0; k < ST_ITEMS; k++) {
    u32 i = tl.start + threadIdx.x + k * ST_THREADS;
    if (i < n) {
      u64 key = 0;
      if (base) {
        // seven characters in radix `base` and the eighth cut down to q8 order-preserving buckets: fewer than
        // 2^56 values, seven passes.  Equal keys share (at least) seven characters: the doubling goes on from 7.
        u32 p = i, c8 = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const u32 cj = code[tx[p]];
          if (j < 7) key = key * base + cj; else c8 = cj;
          p++; if (p >= n) p = 0;
        }
        key = key * q8 + (c8 * q8) / base;
      } else if (i + 8 <= n) {
#pragma unroll
        for (int j = 0; j < 8; j++) key = (key << bits) | code[tx[i + j]];
      } else {
        u32 p = i;
#pragma unroll
        for (int j = 0; j < 8; j++) { key = (key << bits) | code[tx[p]]; p++; if (p >= n) p = 0; }
      }
      keys[off + i] = key;
      vals[off + i] = i;
    }
  }
}

// ---- digit histograms of every pass of a round, per block, from one read of the keys -----------------
// (LSD passes only permute the keys of a block, so the totals of all digits are known up front.)
// A CTA walks HG_TILES consecutive tiles and flushes its counts when the block changes.
#define HG_TILES 8
#define ST_MAXPASS 8
__global__ void __launch_bounds__(ST_THREADS)
k_hist_all(const B2SortTile *__restrict__ tiles, u32 n_tiles, const B2Job *__restrict__ jobs, const u64 *__restrict__ keys,
           int npass, u32 *__restrict__ jobhist) {
  __shared__ u32 h[ST_MAXPASS][256];
  const u32 tid = threadIdx.x;
  for (int p = 0; p < npass; p++) h[p][tid] = 0;
  __syncthreads();
  const u32 t0 = blockIdx.x * HG_TILES, t1 = min(n_tiles, t0 + HG_TILES);
  u32 cur_job = tiles[t0].job;
  for (u32 t = t0; t < t1; t++) {
    const B2SortTile tl = tiles[t];
    if (tl.job != cur_job) {
      __syncthreads();
      for (int p = 0; p < npass; p++) { const u32 c = h[p][tid]; if (c) atomicAdd(&jobhist[((size_t)cur_job * ST_MAXPASS + p) * 256 + tid], c); h[p][tid] = 0; }
      __syncthreads();
      cur_job = tl.job;
    }
    const B2Job &job = jobs[tl.job];
    const u32 n = job.na;
    const u64 *kp = keys + job.pos_off;
#pragma unroll 4
    for (int k = 0; k < ST_ITEMS; k++) {
      const u32 i = tl.start + tid + k * ST_THREADS;
      if (i < n) {
        const u64 key = kp[i];
        for (int p = 0; p < npass; p++) atomicAdd(&h[p][(u32)(key >> (8 * p)) & 255u], 1u);
      }
    }
  }
  __syncthreads();
  for (int p = 0; p < npass; p++) { const u32 c = h[p][tid]; if (c) atomicAdd(&jobhist[((size_t)cur_job * ST_MAXPASS + p) * 256 + tid], c); }
}

// ---- per (block, pass): exclusive scan over the digits, in place ---------------------------------------
__global__ void __launch_bounds__(256)
k_scan_jobs(const B2SortJob *__restrict__ sj, u32 *__restrict__ jobhist) {
  __shared__ u32 sm[40];
  const B2SortJob s = sj[blockIdx.x];
  u32 *h = jobhist + ((size_t)s.job * ST_MAXPASS + blockIdx.y) * 256;
  const u32 c = h[threadIdx.x];
  const u32 base = block_excl_add(c, sm, nullptr);
  h[threadIdx.x] = base;
}

#include "b2_scatter2.cuh"   // tile geometry, shared-memory layout and look-back states of the scatter; k_scatter2

__global__ void __launch_bounds__(SC_THREADS, SC_MINCTAS)
k_scatter(const B2SortTileRR *__restrict__ tiles, const B2Job *__restrict__ jobs,
          const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in,
          u64 *__restrict__ keys_out, u32 *__restrict__ vals_out, int shift, u32 *__restrict__ state,
          const u32 *__restrict__ jobhist, u32 *__restrict__ lb_error, u32 tag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem &S = *reinterpret_cast<ScatterSmem *>(smem_raw);
  const u32 tid = threadIdx.x, w = warp_id(), l = lane_id();
  // Tiles are dispatched in blockIdx order, and that order interleaves the blocks of a group (tile k of
  // every block, then tile k + 1 of every block): the tile before mine in my block was dispatched a whole
  // row of blocks ago and has usually published its inclusive counts, so the look-back below is one early
  // load and no wait.  (A same-address ticket counter costs more than the look-back itself; the wait is
  // bounded and reports through *lb_error instead of hanging should a predecessor never show up.)
  for (int i = tid; i < SC_WARPS * 256; i += SC_THREADS) (&S.warp_cnt[0][0])[i] = 0;
  __syncthreads();
  const u32 tix = blockIdx.x;
  const B2SortTileRR tl = tiles[tix];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.na, off = job.pos_off;
  const u32 lt_mask = (1u << l) - 1u;
  // early look at the state of the tile before mine (usually final by now) and at my digit's base
  u32 v_early = LB_INC | tag;
  if (tid < 256) {
    S.g_off[tid] = jobhist[((size_t)tl.job * ST_MAXPASS + (u32)(shift >> 3)) * 256 + tid];
    if (tl.prev != 0xFFFFFFFFu) v_early = lb_load(state + (size_t)tl.prev * 256 + tid);   // only trusted when final
  }
  const u32 wbase = tl.start + w * SC_WCHUNK;
  // The ranking only needs the digit: its byte is fetched on its own and the whole key is fetched later, next to
  // the rotation index, when its place in the staged tile is known - that second read hits the L2.  (Holding
  // eight 64-bit keys across the ranking made the compiler sink every load next to its use: exposed latency
  // and spills under the 40-register budget of three resident CTAs.)  The state of the ranking is packed: the
  // digits of the even rows in DA, of the odd rows in DB - so the first two rows already need all eight loads,
  // which keeps them together at the top - a validity bit per row, and the rank of a row inside its (warp,
  // digit), at most 255, as a byte of RK0 / RK1.
  u32 DA = 0, DB = 0, vm = 0, RK0 = 0, RK1 = 0;
  {
    const u8 *kb = reinterpret_cast<const u8 *>(keys_in + off) + (shift >> 3);      // little endian: byte shift / 8 of the key
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
      const u32 i = wbase + k * 32 + l;
      const bool valid = i < n;
      const u32 dgk = valid ? (u32)kb[(size_t)i * 8] : 0u;
      if (k & 1) DB |= dgk << (8 * (k >> 1)); else DA |= dgk << (8 * (k >> 1));
      vm |= (valid ? 1u : 0u) << k;
    }
  }
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    const bool valid = (vm >> k) & 1u;
    const u32 d = (((k & 1) ? DB : DA) >> (8 * (k >> 1))) & 255u;
    // lanes holding the same digit: eight ballots (one per digit bit) cost the same for every digit
    // distribution, whereas match.any slows down with the number of distinct digits in the warp
    u32 peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int bb = 0; bb < 8; bb++) {
      const bool bit = (d >> bb) & 1u;
      const u32 m = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? m : ~m;
    }
    const u32 r = __popc(peers & lt_mask);
    const u32 base = valid ? S.warp_cnt[w][d] : 0;
    __syncwarp();
    if (valid && r == 0) S.warp_cnt[w][d] = base + __popc(peers);
    __syncwarp();
    const u32 rank8 = valid ? (base + r) : 0u;              // < 256: a warp holds 256 rows
    if (k < 4) RK0 |= rank8 << (8 * k); else RK1 |= rank8 << (8 * (k - 4));
  }
  __syncthreads();
  // per digit: exclusive over warps, tile count
  {
    u32 run = 0;
    if (tid < 256) {
#pragma unroll
      for (int ww = 0; ww < SC_WARPS; ww++) { u32 c = S.warp_cnt[ww][tid]; S.warp_cnt[ww][tid] = run; run += c; }
    }
    u32 ts = block_excl_add(run, S.scan, nullptr);
    // rows of my digit in the tiles of this block before mine: decoupled look-back, one digit per thread
    if (tid < 256) {
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
      // first place of my digit's rows of this tile in the block, minus their place in the staged tile
      S.g_off[tid] += excl - ts;
    }
  }
  __syncthreads();
  // keys and rotation indices are only needed now
  u64 key[SC_ITEMS];
  u32 val[SC_ITEMS];
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    const u32 i = wbase + k * 32 + l;
    key[k] = (i < n) ? keys_in[off + i] : 0;
    val[k] = (i < n) ? vals_in[off + i] : 0;
  }
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    if ((vm >> k) & 1u) {
      const u32 d = (((k & 1) ? DB : DA) >> (8 * (k >> 1))) & 255u;
      const u32 rank8 = ((k < 4 ? RK0 : RK1) >> (8 * (k & 3))) & 255u;
      const u32 lp = S.tile_start[d] + S.warp_cnt[w][d] + rank8;
      S.keys[lp] = key[k];
      S.vals[lp] = val[k];
    }
  }
  __syncthreads();
  const u32 cnt = (n - tl.start) < SC_TILE ? (n - tl.start) : SC_TILE;
#pragma unroll 4
  for (int k = 0; k < SC_ITEMS; k++) {
    u32 q = tid + k * SC_THREADS;
    if (q < cnt) {
      u64 kk = S.keys[q];
      u32 d = (u32)(kk >> shift) & 255u;
      u32 dst = off + S.g_off[d] + q;
      keys_out[dst] = kk;
      vals_out[dst] = S.vals[q];
    }
  }
}

typedef void (*sc_kernel_t)(const B2SortTileRR *, const B2Job *, const u64 *, const u32 *, u64 *, u32 *, int, u32 *, const u32 *, u32 *, u32);
// B2GPU_SCATTER = 1: round 1's k_scatter; 2 (default): k_scatter2, ballots, three CTAs per SM; 22: two CTAs;
// 3 / 32: match.any; 24: two CTAs, keys loaded once and kept in registers; 34: the same with match.any
static sc_kernel_t sc_pick(int v) {
  switch (v) {
    case 1: return k_scatter;
    case 22: return k_scatter2<2, 0>;
    case 3: return k_scatter2<3, SC2_MATCHANY>;
    case 32: return k_scatter2<2, SC2_MATCHANY>;
    case 24: return k_scatter2<2, SC2_EARLY>;
    case 34: return k_scatter2<2, SC2_EARLY | SC2_MATCHANY>;
    default: return k_scatter2<3, 0>;
  }
}

// ---- class heads, singletons -> ranks, compaction of the still active rows --------------------
// Works on the sorted compact list of a block: c in [0, na).  head(c): key differs from key(c-1);
// single(c): head(c) and (c+1 == na or head(c+1)).
__global__ void __launch_bounds__(ST_THREADS)
k_heads(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs, const u64 *__restrict__ keys,
        i32 *__restrict__ tile_head, u32 *__restrict__ tile_cnt) {
  __shared__ i32 last;
  __shared__ u32 cnt;
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.na;
  const u64 *kp = keys + job.pos_off;
  if (threadIdx.x == 0) { last = -1; cnt = 0; }
  __syncthreads();
  i32 mine = -1;
  u32 act = 0;
#pragma unroll 4
  for (int k = 0; k < ST_ITEMS; k++) {
    u32 i = tl.start + threadIdx.x + k * ST_THREADS;
    if (i < n) {
      const u64 kc = kp[i];
      const bool head = (i == 0) || kc != kp[i - 1];
      const bool nexthead = (i + 1 == n) || kp[i + 1] != kc;
      if (head) mine = (i32)i;
      if (!(head && nexthead)) act++;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { mine = max(mine, __shfl_xor_sync(0xffffffffu, mine, o)); act += __shfl_xor_sync(0xffffffffu, act, o); }
  if (lane_id() == 0) { if (mine >= 0) atomicMax(&last, mine); if (act) atomicAdd(&cnt, act); }
  __syncthreads();
  if (threadIdx.x == 0) { tile_head[blockIdx.x] = last; tile_cnt[blockIdx.x] = cnt; }
}

__global__ void k_scan_heads(const B2SortJob *__restrict__ sj, u32 n_sj, B2Job *jobs,
                             const i32 *__restrict__ tile_head, i32 *__restrict__ carry_in, u32 *__restrict__ tile_cnt) {
  u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sj) return;
  const B2SortJob j = sj[s];
  i32 run = -1;
  u32 pre = 0;
  for (u32 t = 0; t < j.ntiles; t++) {
    carry_in[j.tile0 + t] = run;
    i32 h = tile_head[j.tile0 + t];
    if (h >= 0) run = h;
    u32 c = tile_cnt[j.tile0 + t];
    tile_cnt[j.tile0 + t] = pre;                 // exclusive prefix of active rows
    pre += c;
  }
  jobs[j.job].unsorted = pre;                    // active rows of the next round
}

// slot_in == nullptr: round 0, the compact list is the whole block (slot of c is c).
__global__ void __launch_bounds__(ST_THREADS, 6)
k_ranks(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs, const u64 *__restrict__ keys,
        const u32 *__restrict__ vals, const i32 *__restrict__ carry_in, const u32 *__restrict__ tile_cnt,
        const u32 *__restrict__ slot_in, u32 *__restrict__ rank, u32 *__restrict__ sa_full,
        u32 *__restrict__ slot_out, u32 *__restrict__ vals_out, u32 *__restrict__ grp_out) {
  __shared__ i32 wl[ST_WARPS];
  __shared__ u32 wact[ST_WARPS];
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.na, off = job.pos_off;
  const u64 *kp = keys + off;
  const u32 w = warp_id(), l = lane_id();
  const u32 wbase = tl.start + w * ST_WCHUNK;
  u32 hmask[ST_ITEMS], amask[ST_ITEMS];
  i32 wlast = -1;
  u32 na_w = 0;
#pragma unroll
  for (int k = 0; k < ST_ITEMS; k++) {
    const u32 i = wbase + k * 32 + l;
    bool head = false, act = false;
    if (i < n) {
      const u64 kc = kp[i];
      head = (i == 0) || kc != kp[i - 1];
      const bool nexthead = (i + 1 == n) || kp[i + 1] != kc;
      act = !(head && nexthead);
    }
    const u32 m = __ballot_sync(0xffffffffu, head);
    const u32 am = __ballot_sync(0xffffffffu, act);
    hmask[k] = m; amask[k] = am;
    if (m) wlast = (i32)(wbase + k * 32 + (31 - __clz(m)));
    na_w += __popc(am);
  }
  if (l == 0) { wl[w] = wlast; wact[w] = na_w; }
  __syncthreads();
  i32 carry = carry_in[blockIdx.x];
  u32 cbase = tile_cnt[blockIdx.x];
  for (u32 ww = 0; ww < w; ww++) { carry = max(carry, wl[ww]); cbase += wact[ww]; }
  const u32 le_mask = 0xFFFFFFFFu >> (31 - l);
  const u32 lt_mask = (1u << l) - 1u;
#pragma unroll
  for (int k = 0; k < ST_ITEMS; k++) {
    const u32 i = wbase + k * 32 + l;
    const u32 m = hmask[k], am = amask[k];
    const u32 mm = m & le_mask;
    const i32 headc = mm ? (i32)(wbase + k * 32 + (31 - __clz(mm))) : carry;
    if (m) carry = (i32)(wbase + k * 32 + (31 - __clz(m)));
    if (i < n) {
      const u32 slot = slot_in ? slot_in[off + i] : i;
      const u32 headslot = slot_in ? slot_in[off + (u32)headc] : (u32)headc;
      const u32 v = vals[off + i];
      sa_full[off + slot] = v;
      // later rounds: the key carries the class the row came from (k_keys); a row that stays in the first part of
      // its class keeps its rank, and the random store is skipped
      if (!slot_in || headslot != (u32)(kp[i] >> 20)) rank[off + v] = headslot;
      if ((am >> l) & 1u) {
        const u32 o = off + cbase + __popc(am & lt_mask);
        slot_out[o] = slot;
        vals_out[o] = v;
        grp_out[o] = headslot;
      }
    }
    cbase += __popc(am);
  }
}

// after a round: report the active count and make it the list length of the next round
__global__ void k_collect_unsorted(const B2SortJob *__restrict__ sj, u32 n_sj, B2Job *jobs, u32 *__restrict__ out) {
  u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_sj) { B2Job &j = jobs[sj[s].job]; out[s] = j.unsorted; j.na = j.unsorted; }
}

// ---- doubling keys over the compact list -------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS)
k_keys(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs, const u32 *__restrict__ sa,
       const u32 *__restrict__ grp, const u32 *__restrict__ rank, u64 *__restrict__ keys, u32 h) {
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.n, na = job.na, off = job.pos_off;
  const u32 hh = h % n;
#pragma unroll 4
  for (int k = 0; k < ST_ITEMS; k++) {
    u32 i = tl.start + threadIdx.x + k * ST_THREADS;
    if (i < na) {
      u32 s = sa[off + i] + hh;
      if (s >= n) s -= n;
      keys[off + i] = ((u64)grp[off + i] << 20) | (u64)rank[off + s];
    }
  }
}

// ---- last column + origin pointer (bzip2-encoding.adb:273-280) ---------------------------------
__global__ void __launch_bounds__(ST_THREADS)
k_bwt_out(const B2SortTile *__restrict__ tiles, B2Job *jobs, const u32 *__restrict__ sa,
          const u8 *__restrict__ text, const u32 *__restrict__ rank, u8 *__restrict__ bwt) {
  const B2SortTile tl = tiles[blockIdx.x];
  B2Job &job = jobs[tl.job];
  const u32 n = job.n, off = job.pos_off;
#pragma unroll 4
  for (int k = 0; k < ST_ITEMS; k++) {
    u32 i = tl.start + threadIdx.x + k * ST_THREADS;
    if (i < n) {
      u32 s = sa[off + i];
      bwt[off + i] = text[off + (s == 0 ? n - 1 : s - 1)];
      if (i == 0) job.origin = rank[off];   // first row of the class of rotation 0
    }
  }
}

// =============================================================================================
// Host driver
// =============================================================================================
#include <algorithm>
#include <cstdlib>
#include <vector>

static int build_tiles(const std::vector<u32> &job_ids, const std::vector<u32> &job_n,
                       std::vector<B2SortTile> &tiles, std::vector<B2SortJob> &sj) {
  tiles.clear(); sj.clear();
  for (size_t k = 0; k < job_ids.size(); k++) {
    u32 n = job_n[k];
    if (n == 0) continue;
    B2SortJob s; s.job = job_ids[k]; s.tile0 = (u32)tiles.size(); s.ntiles = (n + ST_TILE - 1) / ST_TILE;
    for (u32 t = 0; t < s.ntiles; t++) tiles.push_back(B2SortTile{s.job, t * ST_TILE});
    sj.push_back(s);
  }
  return 0;
}

// The scatter's dispatch order: blocks are taken in groups of `group`; inside a group, tile k of every
// block, then tile k + 1 of every block, ...  The tile before mine in my block is then `group` tiles
// ahead of me (far enough to have published its counts), while the output runs of a block still arrive
// close enough in time for the L2 to merge them into full lines.
static void build_tiles_rr(const std::vector<B2SortJob> &sj, std::vector<B2SortTileRR> &rr, size_t group) {
  rr.clear();
  std::vector<u32> last(sj.size(), 0xFFFFFFFFu);
  for (size_t g0 = 0; g0 < sj.size(); g0 += group) {
    const size_t g1 = std::min(sj.size(), g0 + group);
    const u32 per = SC_TILE / ST_TILE;                       // sort tiles per scatter tile
    u32 max_nt = 0;
    for (size_t j = g0; j < g1; j++) max_nt = std::max(max_nt, (sj[j].ntiles + per - 1) / per);
    for (u32 k = 0; k < max_nt; k++)
      for (size_t j = g0; j < g1; j++)
        if ((sj[j].ntiles + per - 1) / per > k) { const u32 pos = (u32)rr.size(); rr.push_back(B2SortTileRR{sj[j].job, k * SC_TILE, last[j], 0}); last[j] = pos; }
  }
}

// the same order for k_scatter3, whose tile records are self-contained (first row in the arena, rows, block offset)
static void build_tiles_sc(const std::vector<u32> &ids, const std::vector<u32> &nas, const B2Job *h_jobs, std::vector<B2ScTile> &out,
                           size_t group, u32 sc_tile) {
  out.clear();
  std::vector<u32> jid, jn;
  for (size_t k = 0; k < ids.size(); k++) if (nas[k] > 0) { jid.push_back(ids[k]); jn.push_back(nas[k]); }
  std::vector<u32> last(jid.size(), 0xFFFFFFFFu);
  for (size_t g0 = 0; g0 < jid.size(); g0 += group) {
    const size_t g1 = std::min(jid.size(), g0 + group);
    u32 max_nt = 0;
    for (size_t j = g0; j < g1; j++) max_nt = std::max(max_nt, (jn[j] + sc_tile - 1) / sc_tile);
    for (u32 k = 0; k < max_nt; k++)
      for (size_t j = g0; j < g1; j++)
        if ((jn[j] + sc_tile - 1) / sc_tile > k) {
          const u32 pos = (u32)out.size(), start = k * sc_tile, cnt = std::min(sc_tile, jn[j] - start), off = h_jobs[jid[j]].pos_off;
          out.push_back(B2ScTile{jid[j] | ((cnt - 1) << 16), off + start, last[j], off});
          last[j] = pos;
        }
  }
}

struct EvPair { cudaEvent_t a, b; };

// job_n[k] = post-RLE1 size of block job_ids[k]; the device copies have na == n on entry.
int b2k_bwt_batch(B2SortCtx *cx, cudaStream_t st, B2Job *d_jobs, const std::vector<u32> &job_ids,
                  const std::vector<u32> &job_n, const u8 *d_text, u8 *d_bwt) {
  // B2GPU_SCATTER: see sc_pick; 40..71 = k_scatter3, 40 + bits: 1 = L2 prefetch of the tile B2GPU_SC_PFD places ahead, 2 = tiles of
  // 2048 rows (six CTAs per SM) instead of 4096, 4 = rotation indices requested before the scan, 8 = digit of the output phase
  // from the key registers, 16 = second early look at the predecessor's state
  int sc_variant = 53;                                       // measured best (profiles/r02c_variants.jsonl, r02d_variants.jsonl)
  if (const char *e = getenv("B2GPU_SCATTER")) sc_variant = atoi(e);
  const bool sc3 = sc_variant >= 40 && sc_variant <= 71;
  const int sc_bits = sc3 ? sc_variant - 40 : 0;
  const u32 sc_tile = (sc_bits & 2) ? 2048u : (u32)SC_TILE;
  u32 sc_pfd = 0;
  if (sc_bits & 1) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sc_pfd = (u32)sms * (sc_tile == 2048u ? 2u : 1u);        // a third of the tiles resident on the device (148 measured better than 296 and 592)
    if (const char *e = getenv("B2GPU_SC_PFD")) { long v = atol(e); if (v >= 1 && v < (1 << 20)) sc_pfd = (u32)v; }
  }
  const sc_kernel_t sc_kernel = sc_pick(sc_variant);
  typedef void (*sc3_kernel_t)(const B2ScTile *, const u64 *, const u32 *, u64 *, u32 *, int, u32 *, const u32 *, u32 *, u32, u32);
  sc3_kernel_t sc3_kernel = nullptr;
  {
    const int fl = ((sc_bits & 4) ? SC3_VEARLY : 0) | ((sc_bits & 8) ? SC3_LEAN : 0) | ((sc_bits & 16) ? SC3_MIDLOOK : 0);
    static const sc3_kernel_t k512[8] = {k_scatter3<512, 3, 0>, k_scatter3<512, 3, 1>, k_scatter3<512, 3, 2>, k_scatter3<512, 3, 3>,
                                         k_scatter3<512, 3, 4>, k_scatter3<512, 3, 5>, k_scatter3<512, 3, 6>, k_scatter3<512, 3, 7>};
    static const sc3_kernel_t k256[8] = {k_scatter3<256, 6, 0>, k_scatter3<256, 6, 1>, k_scatter3<256, 6, 2>, k_scatter3<256, 6, 3>,
                                         k_scatter3<256, 6, 4>, k_scatter3<256, 6, 5>, k_scatter3<256, 6, 6>, k_scatter3<256, 6, 7>};
    sc3_kernel = sc_tile == 2048u ? k256[fl] : k512[fl];
  }
  const size_t sc3_smem = sc_tile == 2048u ? sizeof(ScatterSmemT<256>) : sizeof(ScatterSmemT<512>);
  // per device and cheap; set on every call so that handles on several devices / threads all have it
  if (!sc3) B2_CUDA_CHECK(cudaFuncSetAttribute(sc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
  else B2_CUDA_CHECK(cudaFuncSetAttribute(sc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc3_smem));
  std::vector<B2ScTile> tiles_sc;
  std::vector<B2SortTile> tiles;
  std::vector<B2SortTileRR> tiles_rr;
  size_t rr_group = 64;                    // 64 measured best for k_scatter3 (128 for k_scatter2)
  if (const char *e = getenv("B2GPU_RR_GROUP")) { long v = atol(e); if (v >= 1) rr_group = (size_t)v; }
  std::vector<B2SortJob> sj;
  std::vector<u32> ids, ns, nas;             // active blocks: id, size, active rows
  for (size_t k = 0; k < job_ids.size(); k++) if (job_n[k] > 0) { ids.push_back(job_ids[k]); ns.push_back(job_n[k]); }
  nas = ns;
  build_tiles(ids, nas, tiles, sj);
  if (tiles.empty()) return 0;
  if (tiles.size() > cx->max_tiles || sj.size() > cx->max_jobs) { b2_set_error(__FILE__, __LINE__, "sort workspace too small"); return 11; }
  u64 *kA = cx->keysA, *kB = cx->keysB;
  u32 *vA = cx->valsA, *vB = cx->valsB;
  u32 *slotA = cx->slotA, *slotB = cx->slotB;
  std::vector<EvPair> evs;
  auto upload = [&]() -> int {
    B2_CUDA_CHECK(cudaMemcpyAsync(cx->d_tiles, tiles.data(), tiles.size() * sizeof(B2SortTile), cudaMemcpyHostToDevice, st));
    B2_CUDA_CHECK(cudaMemcpyAsync(cx->d_sj, sj.data(), sj.size() * sizeof(B2SortJob), cudaMemcpyHostToDevice, st));
    return 0;
  };
  auto upload_rr = [&]() -> int {
    if (sc3) {
      build_tiles_sc(ids, nas, cx->h_jobs, tiles_sc, rr_group, sc_tile);
      if (tiles_sc.size() > cx->max_tiles) { b2_set_error(__FILE__, __LINE__, "sort workspace too small"); return 11; }
      static_assert(sizeof(B2ScTile) == sizeof(B2SortTileRR), "the two tile records share a buffer");
      B2_CUDA_CHECK(cudaMemcpyAsync(cx->d_tiles_rr, tiles_sc.data(), tiles_sc.size() * sizeof(B2ScTile), cudaMemcpyHostToDevice, st));
      return 0;
    }
    build_tiles_rr(sj, tiles_rr, rr_group);
    B2_CUDA_CHECK(cudaMemcpyAsync(cx->d_tiles_rr, tiles_rr.data(), tiles_rr.size() * sizeof(B2SortTileRR), cudaMemcpyHostToDevice, st));
    return 0;
  };
  // digit totals of all passes of a round (one read of the keys), then their exclusive scans
  u32 max_job = 0;
  for (u32 id : ids) max_job = std::max(max_job, id);
  auto round_hist = [&](int npass) -> int {
    u32 nt = (u32)tiles.size();
    B2_CUDA_CHECK(cudaMemsetAsync(cx->d_digit_base, 0, ((size_t)max_job + 1) * ST_MAXPASS * 256 * sizeof(u32), st));
    k_hist_all<<<(nt + HG_TILES - 1) / HG_TILES, ST_THREADS, 0, st>>>(cx->d_tiles, nt, d_jobs, kA, npass, cx->d_digit_base);
    k_scan_jobs<<<dim3((u32)sj.size(), (u32)npass), 256, 0, st>>>(cx->d_sj, cx->d_digit_base);
    B2_CUDA_CHECK(cudaGetLastError());
    cx->stats.launches += 2;
    return 0;
  };
  const size_t nt0 = tiles.size();       // later rounds never have more tiles than the first
  u32 pass_no = 0;                       // tag of the look-back states; the state array is cleared when it wraps
  auto radix_pass = [&](int shift) -> int {
    u32 nt = (u32)tiles.size();
    if ((pass_no & 255u) == 0) {
      B2_CUDA_CHECK(cudaMemsetAsync(cx->d_hist, 0, nt0 * 256 * sizeof(u32), st));
      B2_CUDA_CHECK(cudaMemsetAsync(cx->d_ticket, 0, sizeof(u32), st));
    }
    const u32 tag = (pass_no & 255u) << 22;
    EvPair ev{nullptr, nullptr};
    if (cx->timing) { cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); cudaEventRecord(ev.a, st); }
    if (!sc3)
      sc_kernel<<<(u32)tiles_rr.size(), SC_THREADS, sizeof(ScatterSmem), st>>>(cx->d_tiles_rr, d_jobs, kA, vA, kB, vB, shift, cx->d_hist, cx->d_digit_base,
                                                                                 cx->d_ticket, tag);
    else
      sc3_kernel<<<(u32)tiles_sc.size(), sc_tile / SC_ITEMS, sc3_smem, st>>>(reinterpret_cast<const B2ScTile *>(cx->d_tiles_rr), kA, vA, kB, vB, shift,
                                                                              cx->d_hist, cx->d_digit_base, cx->d_ticket, tag, sc_pfd);
    pass_no++;
    if (cx->timing) { cudaEventRecord(ev.b, st); evs.push_back(ev); }
    B2_CUDA_CHECK(cudaGetLastError());
    std::swap(kA, kB); std::swap(vA, vB);
    u64 el = 0;
    for (u32 x : nas) el += x;
    cx->stats.scatter_launches++;
    cx->stats.scatter_elems += el;
    cx->stats.launches += 1;
    return 0;
  };
  // heads -> ranks -> compaction; afterwards vA holds the compacted rotation indices, slotA their slots
  auto ranks = [&](bool round0) -> int {
    u32 nt = (u32)tiles.size(), nj = (u32)sj.size();
    k_heads<<<nt, ST_THREADS, 0, st>>>(cx->d_tiles, d_jobs, kA, cx->d_tile_head, cx->d_tile_cnt);
    k_scan_heads<<<(nj + 127) / 128, 128, 0, st>>>(cx->d_sj, nj, d_jobs, cx->d_tile_head, cx->d_carry, cx->d_tile_cnt);
    k_ranks<<<nt, ST_THREADS, 0, st>>>(cx->d_tiles, d_jobs, kA, vA, cx->d_carry, cx->d_tile_cnt, round0 ? nullptr : slotA,
                                       cx->rank, cx->sa_full, slotB, vB, cx->grp);
    k_collect_unsorted<<<(nj + 127) / 128, 128, 0, st>>>(cx->d_sj, nj, d_jobs, cx->d_unsorted);
    cx->stats.launches += 4;
    B2_CUDA_CHECK(cudaGetLastError());
    std::swap(vA, vB); std::swap(slotA, slotB);
    B2_CUDA_CHECK(cudaMemcpyAsync(cx->h_unsorted, cx->d_unsorted, nj * sizeof(u32), cudaMemcpyDeviceToHost, st));
    u32 lb_err = 0;
    B2_CUDA_CHECK(cudaMemcpyAsync(&lb_err, cx->d_ticket, sizeof(u32), cudaMemcpyDeviceToHost, st));
    B2_CUDA_CHECK(cudaStreamSynchronize(st));
    if (lb_err) { b2_set_error(__FILE__, __LINE__, "radix scatter: a tile's predecessor never published its counts"); return 11; }
    return 0;
  };

  int rc;
  if ((rc = upload())) return rc;
  {
    u64 el = 0; for (u32 x : ns) el += x;
    cx->stats.sorted_elems_round0 += el;
  }
  const int sym_bits = (cx->sym_bits >= 1 && cx->sym_bits <= 8) ? cx->sym_bits : 8;   // 8 characters of sym_bits bits = sym_bits passes
  // Alphabets of more than 128 bytes (sym_bits = 8): seven characters in radix B = the largest alphabet of the batch plus
  // the eighth cut down to floor (2^56 / B^7) buckets make a key below 2^56: seven passes instead of eight, and the
  // doubling goes on from a prefix of 7 (B2GPU_R0 = 8 keeps the eight full characters).
  u32 r0_base = 0, r0_q8 = 0;
  {
    int r0 = 7;
    if (const char *e = getenv("B2GPU_R0")) r0 = atoi(e);
    if (r0 == 7 && sym_bits == 8 && cx->max_used > 128 && cx->max_used <= 256) {
      u64 b7 = 1;
      for (int j = 0; j < 7; j++) b7 *= cx->max_used;                 // <= 2^56
      r0_base = cx->max_used;
      r0_q8 = (u32)((1ull << 56) / b7);                                // >= 1
    }
  }
  const int r0_passes = r0_base ? 7 : sym_bits;
  k_keys0<<<(u32)tiles.size(), ST_THREADS, 0, st>>>(cx->d_tiles, d_jobs, d_text, kA, vA, sym_bits, r0_base, r0_q8);
  cx->stats.launches += 1;
  if ((rc = upload_rr())) return rc;
  if ((rc = round_hist(r0_passes))) return rc;
  for (int p = 0; p < r0_passes; p++) if ((rc = radix_pass(8 * p))) return rc;
  if ((rc = ranks(true))) return rc;
  cx->stats.rounds++;
  u64 reflect = r0_base ? 7 : 8;          // characters that rows of one class are known to share
  for (;;) {
    // split finished / unfinished
    std::vector<u32> fin_ids, fin_n, go_ids, go_n, go_na;
    for (size_t k = 0; k < ids.size(); k++) {
      bool done = cx->h_unsorted[k] == 0 || reflect >= ns[k];
      if (done) { fin_ids.push_back(ids[k]); fin_n.push_back(ns[k]); }
      else { go_ids.push_back(ids[k]); go_n.push_back(ns[k]); go_na.push_back(cx->h_unsorted[k]); }
    }
    if (!fin_ids.empty()) {
      build_tiles(fin_ids, fin_n, tiles, sj);
      if ((rc = upload())) return rc;
      k_bwt_out<<<(u32)tiles.size(), ST_THREADS, 0, st>>>(cx->d_tiles, d_jobs, cx->sa_full, d_text, cx->rank, d_bwt);
      B2_CUDA_CHECK(cudaGetLastError());
      cx->stats.launches += 1;
    }
    if (go_ids.empty()) break;
    ids.swap(go_ids); ns.swap(go_n); nas.swap(go_na);
    build_tiles(ids, nas, tiles, sj);
    if ((rc = upload())) return rc;
    {
      u64 el = 0; for (u32 x : nas) el += x;
      cx->stats.sorted_elems_later += el;
    }
    k_keys<<<(u32)tiles.size(), ST_THREADS, 0, st>>>(cx->d_tiles, d_jobs, vA, cx->grp, cx->rank, kA, (u32)(reflect));
    cx->stats.launches += 1;
    if ((rc = upload_rr())) return rc;
    if ((rc = round_hist(5))) return rc;
    for (int p = 0; p < 5; p++) if ((rc = radix_pass(8 * p))) return rc;
    if ((rc = ranks(false))) return rc;
    cx->stats.rounds++;
    reflect *= 2;
  }
  if (cx->timing) {
    B2_CUDA_CHECK(cudaStreamSynchronize(st));
    for (auto &e : evs) {
      float ms = 0;
      cudaEventElapsedTime(&ms, e.a, e.b);
      cx->stats.scatter_ms += ms;
      cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
  }
  return 0;
}
