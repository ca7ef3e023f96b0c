// Stage A1: chunk cutting (Data_Acquisition) as device scans plus a short serial chain.
//
// Reference: zip_lib/bzip2-encoding.adb:1160-1208 (Data_Acquisition) and :1413-1429 (chunk loop with
// the last-two-blocks balancing).  The reference reads bytes while `rle_1_block_size + 5 <
// capacity` (and < 10 x capacity raw bytes), where rle_1_block_size counts the RLE1 size of the
// *completed* runs, a run being cut every 259 bytes (:1171-1179, :1195-1204).
//
// A "piece" is a maximal equal-byte run counted from the chunk start, split every 259; when the
// first byte of a new piece is read the previous piece's size min(len,4)+[len>=4] is committed; the
// chunk ends with the first byte whose commit makes committed + 5 >= capacity.
// Only the FIRST run of a chunk depends on where the chunk starts (a chunk may start in the middle
// of a run of the stream); all later runs are runs of the stream itself.  Hence:
//   k_cut_a   per 2048-byte tile: first / last position where the byte changes       (all SMs)
//   k_cut_s1  exclusive max-scan over tiles -> start of the run that enters each tile
//   k_cut_b   per tile: sum of the commits g(p) defined with the runs of the STREAM      (all SMs)
//   k_cut_s2  inclusive add-scan over tiles -> G
//   k_cut_chain  one warp walks the chunks: first run analytically, then a binary search on G and
//                one in-tile scan locate the end of the chunk.
#include "b2_common.cuh"
#include "b2_kernels.h"

#define CT_TILE 2048
#define CT_THREADS 128          // 16 bytes per thread
#define NONE32 0xFFFFFFFFu
#define B2_TRY_K(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

__device__ __forceinline__ u32 enc_size(u32 len) { return (len < 4 ? len : 4) + (len >= 4 ? 1 : 0); }

__device__ __forceinline__ void load16(const u8 *__restrict__ in, u64 base, u64 n_alloc, u8 *b) {
  // 16-byte aligned vector load when the whole vector is inside the allocation
  if (base + 16 <= n_alloc && ((reinterpret_cast<uintptr_t>(in + base)) & 15u) == 0) {
    const uint4 v = *reinterpret_cast<const uint4 *>(in + base);
    const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; k++) b[k] = (u8)(w[k >> 2] >> (8 * (k & 3)));
  } else {
#pragma unroll
    for (int k = 0; k < 16; k++) b[k] = (base + k < n_alloc) ? in[base + k] : 0;
  }
}

__global__ void __launch_bounds__(CT_THREADS)
k_cut_a(const u8 *__restrict__ in, u64 n, u32 *__restrict__ firstchg, u32 *__restrict__ lastchg) {
  __shared__ u32 s_first, s_last;
  const u64 t0 = (u64)blockIdx.x * CT_TILE;
  const u32 tid = threadIdx.x;
  if (tid == 0) { s_first = NONE32; s_last = 0; }
  __syncthreads();
  const u64 base = t0 + tid * 16;
  u8 b[16];
  load16(in, base, n, b);
  u8 pv = (base > 0 && base < n) ? in[base - 1] : 0;
  u32 f = NONE32, l = 0;     // l holds offset + 1, 0 = none
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const u64 p = base + k;
    if (p < n && (p == 0 || b[k] != pv)) { const u32 o = tid * 16 + k; if (f == NONE32) f = o; l = o + 1; }
    pv = b[k];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
    l = max(l, __shfl_xor_sync(0xffffffffu, l, o));
  }
  if (lane_id() == 0) { if (f != NONE32) atomicMin(&s_first, f); if (l) atomicMax(&s_last, l); }
  __syncthreads();
  if (tid == 0) { firstchg[blockIdx.x] = s_first; lastchg[blockIdx.x] = s_last; }
}

// carry_r[t] = (global position of the last change before tile t) + 1, 0 if none.  One CTA.
__global__ void __launch_bounds__(1024)
k_cut_s1(const u32 *__restrict__ lastchg, u64 ntiles, u64 *__restrict__ carry_r) {
  __shared__ u64 sm[1024];
  const u32 tid = threadIdx.x;
  const u64 per = (ntiles + 1023) / 1024;
  const u64 a = tid * per, e = min(ntiles, a + per);
  u64 loc = 0;
  for (u64 t = a; t < e; t++) { u32 l = lastchg[t]; if (l) loc = t * CT_TILE + l; }   // (+1 kept: l = offset + 1)
  sm[tid] = loc;
  __syncthreads();
  if (tid == 0) { u64 run = 0; for (int i = 0; i < 1024; i++) { u64 v = sm[i]; sm[i] = run; if (v) run = v; } }
  __syncthreads();
  u64 run = sm[tid];
  for (u64 t = a; t < e; t++) { carry_r[t] = run; u32 l = lastchg[t]; if (l) run = t * CT_TILE + l; }
}

// commits with the runs of the stream: g(p) = size of the piece that ends at p-1 if a piece starts at p.
// `m` = (p-1 - run start of p-1) mod 259, the index of byte p-1 inside its piece, carried along so that
// no division is needed per byte; returns g(p) and advances m to the index of byte p.
__device__ __forceinline__ u32 gstep(bool first_of_stream, bool chg, u32 &m) {
  u32 g;
  if (first_of_stream) { m = 0; return 0; }
  if (chg) { g = enc_size(m + 1); m = 0; }
  else if (m == 258) { g = 5; m = 0; }                     // forced break at 259 (:1199)
  else { g = 0; m++; }
  return g;
}
__device__ __forceinline__ u32 mod259(u64 d) { return (d >> 32) ? (u32)(d % 259ull) : ((u32)d % 259u); }

__global__ void __launch_bounds__(CT_THREADS)
k_cut_b(const u8 *__restrict__ in, u64 n, const u64 *__restrict__ carry_r, u32 *__restrict__ tsum, u8 *__restrict__ gs, u16 *__restrict__ gm) {
  __shared__ i32 sm_i[40];
  __shared__ u32 sm_u[40];
  const u64 t0 = (u64)blockIdx.x * CT_TILE;
  const u32 tid = threadIdx.x;
  const u64 base = t0 + tid * 16;
  u8 b[16];
  load16(in, base, n, b);
  const u8 prev = (base > 0 && base < n) ? in[base - 1] : 0;
  i32 lc = -1;
  {
    u8 pv = prev;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const u64 p = base + k;
      if (p < n && (p == 0 || b[k] != pv)) lc = (i32)(tid * 16 + k);
      pv = b[k];
    }
  }
  i32 tot;
  const i32 rin = block_excl_max(lc, -1, sm_i, &tot);
  const u64 cr = carry_r[blockIdx.x];
  const u64 r = rin >= 0 ? t0 + (u64)rin : (cr ? cr - 1 : 0);     // run start of byte base-1 (unused when base == 0)
  u32 local = 0;
  {
    u32 m = (base > 0 && base <= n) ? mod259(base - 1 - r) : 0;   // index of byte base-1 inside its piece
    gm[(size_t)blockIdx.x * CT_THREADS + tid] = (u16)m;
    u8 pv = prev;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const u64 p = base + k;
      if (p < n) local += gstep(p == 0, b[k] != pv, m);
      pv = b[k];
    }
  }
  gs[(size_t)blockIdx.x * CT_THREADS + tid] = (u8)local;          // at most 16 x 5
  u32 total;
  block_excl_add(local, sm_u, &total);
  if (tid == 0) tsum[blockIdx.x] = total;
}

// The two in-tile questions of the chain, answered from the granule sums: one 128-byte load and a warp scan
// find the granule, at most 16 bytes are then stepped through.
//  mode 0: inclusive in-tile prefix of g at global position q          -> returns the prefix
//  mode 1: first global position p in the tile with excl + prefix(p) >= target   -> position or ~0
__device__ u64 warp_tile_g(const u8 *__restrict__ in, u64 n, const u8 *__restrict__ gs, const u16 *__restrict__ gm, u64 t, int mode,
                           u64 q, u64 excl, u64 target) {
  const u32 l = lane_id();
  const u64 t0 = t * CT_TILE;
  const u32 w = *reinterpret_cast<const u32 *>(gs + t * CT_THREADS + 4 * l);      // granules 4l .. 4l+3
  const u32 s0 = w & 255u, s1 = (w >> 8) & 255u, s2 = (w >> 16) & 255u, s3 = w >> 24;
  const u32 c0 = s0, c1 = c0 + s1, c2 = c1 + s2, c3 = c2 + s3;                     // inclusive inside my four
  const u32 incl = warp_incl_add(c3);
  const u32 before = incl - c3;                                                    // granules of the lanes before me
  u32 gq;                                                                          // the granule to step through
  u64 run;                                                                         // prefix before that granule (mode 0: in-tile; mode 1: global)
  u64 need = 0;
  if (mode == 0) {
    gq = (u32)((q - t0) >> 4);
    const u32 k = gq & 3u;
    const u32 mine = before + (k == 0 ? 0u : k == 1 ? c0 : k == 2 ? c1 : c2);
    run = __shfl_sync(0xffffffffu, mine, gq >> 2);
  } else {
    if (target <= excl) need = 0; else need = target - excl;
    // (need == 0 cannot happen: the chain asks for a position strictly behind the chunk's first run)
    const u32 tot = __shfl_sync(0xffffffffu, incl, 31);
    if (need > tot) return ~0ull;
    const u32 nd = (u32)need;
    int k = -1;
    if (before + c0 >= nd) k = 0; else if (before + c1 >= nd) k = 1; else if (before + c2 >= nd) k = 2; else if (before + c3 >= nd) k = 3;
    const u32 bm = __ballot_sync(0xffffffffu, k >= 0);
    if (!bm) return ~0ull;
    const int src = __ffs(bm) - 1;
    const int ks = __shfl_sync(0xffffffffu, k, src);
    const u32 pre = __shfl_sync(0xffffffffu, before + (ks == 0 ? 0u : ks == 1 ? c0 : ks == 2 ? c1 : c2), src);
    gq = 4u * (u32)src + (u32)ks;
    run = excl + pre;
  }
  // step through granule gq (every lane does the same: uniform)
  const u64 base = t0 + 16ull * gq;
  u8 b[16];
  load16(in, base, n, b);
  u8 pv = (base > 0 && base < n) ? in[base - 1] : 0;
  u32 m = gm[t * CT_THREADS + gq];
  u64 res = ~0ull;
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const u64 p = base + k;
    if (p < n && res == ~0ull) {
      run += gstep(p == 0, b[k] != pv, m);
      if (mode == 0 && p == q) res = run;
      if (mode == 1 && run >= target) res = p;
    }
    pv = b[k];
  }
  return res;
}

__global__ void __launch_bounds__(1024)
k_cut_s2(const u32 *__restrict__ tsum, u64 ntiles, u64 *__restrict__ tincl) {
  __shared__ u64 sm[1024];
  const u32 tid = threadIdx.x;
  const u64 per = (ntiles + 1023) / 1024;
  const u64 a = tid * per, e = min(ntiles, a + per);
  u64 loc = 0;
  for (u64 t = a; t < e; t++) loc += tsum[t];
  sm[tid] = loc;
  __syncthreads();
  if (tid == 0) { u64 run = 0; for (int i = 0; i < 1024; i++) { u64 v = sm[i]; sm[i] = run; run += v; } }
  __syncthreads();
  u64 run = sm[tid];
  for (u64 t = a; t < e; t++) { run += tsum[t]; tincl[t] = run; }
}

// One warp scans one tile.  Each lane owns 64 consecutive bytes, loaded once as four 16-byte vectors.
//  mode 0: inclusive in-tile prefix of g at global position q          -> returns the prefix
//  mode 1: first global position p in the tile with excl + prefix(p) >= target   -> position or ~0
//  mode 2: first global position p > q in the tile where the byte changes        -> position or ~0
__device__ u64 warp_tile(const u8 *__restrict__ in, u64 n, const u64 *__restrict__ carry_r, u64 t, int mode,
                         u64 q, u64 excl, u64 target) {
  const u32 l = lane_id();
  const u64 t0 = t * CT_TILE;
  const u64 base = t0 + l * 64;
  u32 w[16];
#pragma unroll
  for (int v = 0; v < 4; v++) {
    u8 b16[16];
    load16(in, base + 16 * v, n, b16);
#pragma unroll
    for (int k = 0; k < 4; k++) w[4 * v + k] = (u32)b16[4 * k] | ((u32)b16[4 * k + 1] << 8) | ((u32)b16[4 * k + 2] << 16) | ((u32)b16[4 * k + 3] << 24);
  }
  const u8 prev = (base > 0 && base < n) ? in[base - 1] : 0;
  u64 res = ~0ull;
  if (mode == 2) {
    u8 pv = prev;
#pragma unroll
    for (int k = 0; k < 64; k++) {
      const u64 p = base + k;
      const u8 c = (u8)(w[k >> 2] >> (8 * (k & 3)));
      if (p < n && p > q && (p == 0 || c != pv) && res == ~0ull) res = p;
      pv = c;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { u64 x = __shfl_xor_sync(0xffffffffu, res, o); res = min(res, x); }
    return res;
  }
  // run start entering my 64 bytes
  i32 lc = -1;
  {
    u8 pv = prev;
#pragma unroll
    for (int k = 0; k < 64; k++) {
      const u64 p = base + k;
      const u8 c = (u8)(w[k >> 2] >> (8 * (k & 3)));
      if (p < n && (p == 0 || c != pv)) lc = (i32)(l * 64 + k);
      pv = c;
    }
  }
  i32 inc = warp_incl_max(lc);
  i32 rin = __shfl_up_sync(0xffffffffu, inc, 1);
  if (l == 0) rin = -1;
  const u64 cr = carry_r[t];
  const u64 r0 = rin >= 0 ? t0 + (u64)rin : (cr ? cr - 1 : 0);
  // my sum
  const u32 m0 = (base > 0 && base <= n) ? mod259(base - 1 - r0) : 0;   // index of byte base-1 inside its piece
  u32 local = 0;
  {
    u32 m = m0;
    u8 pv = prev;
#pragma unroll 8
    for (int k = 0; k < 64; k++) {
      const u64 p = base + k;
      const u8 c = (u8)(w[k >> 2] >> (8 * (k & 3)));
      if (p < n) local += gstep(p == 0, c != pv, m);
      pv = c;
    }
  }
  const u32 incl = warp_incl_add(local);
  u64 run = excl + (incl - local);
  {
    u32 m = m0;
    u8 pv = prev;
#pragma unroll 8
    for (int k = 0; k < 64; k++) {
      const u64 p = base + k;
      const u8 c = (u8)(w[k >> 2] >> (8 * (k & 3)));
      if (p < n) {
        run += gstep(p == 0, c != pv, m);
        if (mode == 0 && p == q) res = run - excl;
        if (mode == 1 && res == ~0ull && run >= target) res = p;
      }
      pv = c;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { u64 x = __shfl_xor_sync(0xffffffffu, res, o); res = min(res, x); }
  return res;
}

// `in` holds the bytes [gbase, gbase + n) of the stream (gbase = 0, and the whole stream, unless the stream is
// spread over several handles: b2_shard_*).  The walk starts at local position pos0 and ends with the first
// chunk that starts at or after `stop` (that start is handed to the next shard through n_chunks[1..2]) or
// with the end of the bytes.
__global__ void __launch_bounds__(32)
k_cut_chain(const u8 *__restrict__ in, u64 n, i64 size_hint, int level, i64 win_lo, i64 win_hi,
            const u32 *__restrict__ firstchg, const u64 *__restrict__ carry_r, const u64 *__restrict__ tincl, u64 ntiles,
            const u8 *__restrict__ gs, const u16 *__restrict__ gm,
            B2Chunk *chunks, u32 *n_chunks, u32 max_chunks, u32 *progress, u64 gbase, u64 pos0, u64 stop) {
  // progress[0] = chunks published so far, progress[1] = 1 when the chain is complete: k_segment may be
  // following on another stream and starts on a chunk as soon as it is there
  const u32 l = lane_id();
  u64 pos = pos0;
  u32 nc = 0;
  for (;;) {
    if (pos >= stop && !(pos0 == 0 && gbase == 0 && n == 0)) break;   // (an empty stream is still one empty chunk, :1428)
    // stream_rest: size_hint - bytes read, sticking at -1 (= unknown_size) once it gets there
    // (bzip2-encoding.adb:1192-1194); the balancing test :1416-1424 was done in float32 on the host
    // and arrives as the integer window [win_lo, win_hi].
    const i64 gpos = (i64)(gbase + pos);
    const i64 rest = size_hint < 0 ? -1 : (gpos <= size_hint ? size_hint - gpos : -1);
    i64 cap = (i64)level * 100000;
    if (rest >= win_lo && rest <= win_hi) cap = rest / 2;
    const u64 avail = n - pos;
    const u64 rawmax = (u64)(10 * cap);                      // multiplier = 10 (:1156)
    const u64 limit = avail < rawmax ? avail : rawmax;       // raw bytes this chunk may take
    const u64 need = cap > 5 ? (u64)(cap - 5) : 0;           // stop once committed >= cap - 5
    u64 len = limit;
    if (need == 0) len = 0;
    else if (limit > 0) {
      const u64 s = pos, end_max = pos + limit;
      // e1: first position after s where the byte changes (end of the chunk's first run)
      u64 e1 = ~0ull;
      {
        u64 t = s / CT_TILE;
        {
          // the usual case: the byte changes within the next 32 positions
          const u64 p = s + 1 + l;
          const bool chg = p < n && in[p] != in[p - 1];
          const u32 bm = __ballot_sync(0xffffffffu, chg);
          if (bm) e1 = s + 1 + (u64)(__ffs(bm) - 1);
        }
        if (e1 == ~0ull) e1 = warp_tile(in, n, carry_r, t, 2, s, 0, 0);
        if (e1 == ~0ull) {
          // skip tiles without any change, 32 at a time
          for (u64 tb = t + 1; tb < ntiles && e1 == ~0ull; tb += 32) {
            const u64 tt = tb + l;
            const u32 f = tt < ntiles ? firstchg[tt] : NONE32;
            const u32 m = __ballot_sync(0xffffffffu, f != NONE32);
            if (m) { const int k = __ffs(m) - 1; const u32 fk = __shfl_sync(0xffffffffu, f, k); e1 = (tb + k) * CT_TILE + fk; }
            if (tb * CT_TILE >= end_max) break;
          }
        }
        if (e1 == ~0ull) e1 = n;
      }
      const u64 run_end = e1 < end_max ? e1 : end_max;       // first run inside the window: [s, run_end)
      const u64 jstar = (need + 4) / 5;                      // forced breaks needed to reach `need` (5 each, :1199)
      const u64 pstar = s + 259 * jstar;
      if (pstar < run_end) len = pstar - s + 1;              // (i) ends inside the first run
      else if (e1 >= end_max) len = limit;                   // (ii) window exhausted inside the first run
      else {
        const u64 L1 = e1 - s;
        const u64 A = 5 * ((L1 - 1) / 259) + enc_size((u32)((L1 - 1) % 259) + 1);
        if (A >= need) len = e1 - s + 1;                     // (iii) ends with the first byte of the second run
        else {
          const u64 te = e1 / CT_TILE;
          const u64 excl_e = te ? tincl[te - 1] : 0;
          const u64 Ge1 = excl_e + warp_tile_g(in, n, gs, gm, te, 0, e1, 0, 0);
          const u64 Gt = Ge1 + (need - A);
          // first tile whose inclusive prefix reaches Gt
          // (32 probes per trip, one per lane; tiles at or beyond the end of the window need not be looked at)
          u64 lo = te, hi = ntiles;                          // answer in [lo, hi]; hi = "none below"
          { const u64 wend = end_max / CT_TILE + 1; if (wend < hi) hi = wend; }
          while (lo < hi) {
            const u64 span = hi - lo, step = (span + 31) / 32;
            const u64 upto = (u64)(l + 1) * step;
            const u64 pr = lo + (upto < span ? upto : span) - 1;      // non-decreasing in the lane, the last ones = hi - 1
            const u32 m = __ballot_sync(0xffffffffu, tincl[pr] >= Gt);
            if (m == 0) { lo = hi; break; }
            const int k = __ffs(m) - 1;
            hi = __shfl_sync(0xffffffffu, pr, k);
            lo = lo + (u64)k * step;
          }
          if (lo < ntiles && lo * CT_TILE < end_max) {
            const u64 excl = lo ? tincl[lo - 1] : 0;
            const u64 p = warp_tile_g(in, n, gs, gm, lo, 1, 0, excl, Gt);
            if (p != ~0ull && p < end_max) len = p - s + 1;  // (iv)
          }
        }
      }
    }
    if (l == 0 && nc < max_chunks) { chunks[nc].start = pos; chunks[nc].len = (u32)len; chunks[nc].cap = (u32)cap; chunks[nc].pad = 0; chunks[nc].pad2 = 0; }
    nc++;
    if (l == 0 && progress) { __threadfence(); *(volatile u32 *)progress = nc < max_chunks ? nc : max_chunks; }
    pos += len;
    if (pos >= n) break;                                     // exit when not More_Bytes (:1428)
    if (len == 0) break;
  }
  if (l == 0) {
    n_chunks[0] = nc;
    n_chunks[2] = (u32)pos; n_chunks[3] = (u32)(pos >> 32);     // where the next shard's first chunk starts (local)
    if (progress) { __threadfence(); ((volatile u32 *)progress)[1] = 1u; }
  }
}

int b2k_cut_scans(cudaStream_t st, const u8 *d_in, u64 n, B2CutWork *w) {
  const u64 ntiles = (n + CT_TILE - 1) / CT_TILE;
  if (ntiles) {
    k_cut_a<<<(u32)ntiles, CT_THREADS, 0, st>>>(d_in, n, w->firstchg, w->lastchg);
    k_cut_s1<<<1, 1024, 0, st>>>(w->lastchg, ntiles, w->carry_r);
    k_cut_b<<<(u32)ntiles, CT_THREADS, 0, st>>>(d_in, n, w->carry_r, w->tsum, w->gs, w->gm);
    k_cut_s2<<<1, 1024, 0, st>>>(w->tsum, ntiles, w->tincl);
    B2_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

int b2k_cut_chain(cudaStream_t st, const u8 *d_in, u64 n, i64 size_hint, int level, i64 win_lo, i64 win_hi,
                  B2Chunk *d_chunks, u32 *d_n_chunks, u32 max_chunks, B2CutWork *w, u32 *d_progress, cudaEvent_t ev_chain_starts,
                  u64 gbase, u64 pos0, u64 stop) {
  const u64 ntiles = (n + CT_TILE - 1) / CT_TILE;
  if (d_progress) B2_CUDA_CHECK(cudaMemsetAsync(d_progress, 0, 2 * sizeof(u32), st));
  if (ev_chain_starts) B2_CUDA_CHECK(cudaEventRecord(ev_chain_starts, st));
  k_cut_chain<<<1, 32, 0, st>>>(d_in, n, size_hint, level, win_lo, win_hi, w->firstchg, w->carry_r, w->tincl, ntiles, w->gs, w->gm,
                                d_chunks, d_n_chunks, max_chunks, d_progress, gbase, pos0, stop);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int b2k_cut(cudaStream_t st, const u8 *d_in, u64 n, i64 size_hint, int level, i64 win_lo, i64 win_hi,
            B2Chunk *d_chunks, u32 *d_n_chunks, u32 max_chunks, B2CutWork *w, u32 *d_progress, cudaEvent_t ev_chain_starts) {
  B2_TRY_K(b2k_cut_scans(st, d_in, n, w));
  return b2k_cut_chain(st, d_in, n, size_hint, level, win_lo, win_hi, d_chunks, d_n_chunks, max_chunks, w, d_progress, ev_chain_starts,
                       0, 0, n);
}
