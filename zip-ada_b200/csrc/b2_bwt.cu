// Stage A6: Burrows-Wheeler rotation sort for a batch of blocks.
//
// Reference: zip_lib/bzip2-encoding.adb:219-296.  The reference heap-sorts rotation offsets with
// an O(N) comparator that is a total order (:229-255), so the sorted matrix, its last column and
// the origin pointer are unique functions of the block; any sort is admissible.
//
// Here: cyclic prefix doubling with filtering.  Round 0 sorts every rotation by its first 8 bytes
// (64-bit key, 8 LSD radix passes of 8 bits).  A rank is the row index ("slot") of the first row of
// a class, so equal prefixes share a rank.  After every round the rows that are alone in their class
// are final; only the others ("active" rows) are compacted and re-sorted in round r >= 1 by the
// 40-bit key (rank[i] << 20 | rank[(i+h) mod n]) with 5 passes, h = 8, 16, ..., and written back to
// the slots they came from (a class occupies a contiguous range of slots, and the key's high part
// keeps classes in place).  A block is finished when no row is active or when the compared prefix
// covers the whole rotation (periodic blocks: the classes are then sets of EQUAL rotations; their
// last-column bytes coincide and the origin pointer is the first row of the class of rotation 0,
// because the reference orders equal rotations by offset and rotation 0 has offset 0, :254,
// :276-279).  All blocks of a batch are sorted together: every radix pass is segmented by block
// through a tile table, so no block id is needed in the key.
//
// Radix pass = histogram (8 B/elt read) + per-block scan (tiny) + scatter (12 B/elt read, 12 B/elt
// write, staged through shared memory so the writes leave as runs).
#include "b2_common.cuh"
#include "b2_kernels.h"

#define ST_THREADS 256
#define ST_ITEMS 8
#define ST_TILE (ST_THREADS * ST_ITEMS)
#define ST_WARPS (ST_THREADS / 32)
#define ST_WCHUNK (32 * ST_ITEMS)      // elements per warp

// ---- round-0 keys: first 8 bytes of each rotation, cyclic ------------------------------------
__global__ void __launch_bounds__(ST_THREADS)
k_keys0(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs, const u8 *__restrict__ text,
        u64 *__restrict__ keys, u32 *__restrict__ vals) {
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.n, off = job.pos_off;
  const u8 *tx = text + off;
#pragma unroll 4
  for (int k = 0; k < ST_ITEMS; k++) {
    u32 i = tl.start + threadIdx.x + k * ST_THREADS;
    if (i < n) {
      u64 key = 0;
      if (i + 8 <= n) {
#pragma unroll
        for (int j = 0; j < 8; j++) key = (key << 8) | tx[i + j];
      } else {
        u32 p = i;
#pragma unroll
        for (int j = 0; j < 8; j++) { key = (key << 8) | tx[p]; p++; if (p >= n) p = 0; }
      }
      keys[off + i] = key;
      vals[off + i] = i;
    }
  }
}

// ---- histogram of one digit per tile -----------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS)
k_hist(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs, const u64 *__restrict__ keys,
       int shift, u32 *__restrict__ hist) {
  __shared__ u32 h[256];
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.na;
  const u64 *kp = keys + job.pos_off;
  if (threadIdx.x < 256) h[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll 4
  for (int k = 0; k < ST_ITEMS; k++) {
    u32 i = tl.start + threadIdx.x + k * ST_THREADS;
    if (i < n) atomicAdd(&h[(u32)(kp[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 256) hist[(size_t)blockIdx.x * 256 + threadIdx.x] = h[threadIdx.x];
}

// ---- per-block exclusive scan over (digit major, tile minor) ------------------------------------
__global__ void __launch_bounds__(256)
k_scan(const B2SortJob *__restrict__ sj, u32 *__restrict__ hist, u32 *__restrict__ digit_base) {
  __shared__ u32 sm[40];
  const B2SortJob s = sj[blockIdx.x];
  const u32 d = threadIdx.x;
  u32 *h = hist + (size_t)s.tile0 * 256 + d;
  u32 total = 0;
  u32 t = 0;
  for (; t + 8 <= s.ntiles; t += 8) {             // 8 independent loads in flight per thread
    u32 c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = h[(size_t)(t + k) * 256];
#pragma unroll
    for (int k = 0; k < 8; k++) { h[(size_t)(t + k) * 256] = total; total += c[k]; }
  }
  for (; t < s.ntiles; t++) { u32 c = h[(size_t)t * 256]; h[(size_t)t * 256] = total; total += c; }
  u32 base = block_excl_add(total, sm, nullptr);
  digit_base[(size_t)s.job * 256 + d] = base;     // added by the scatter kernel
}

// ---- stable scatter of one digit ---------------------------------------------------------------
struct ScatterSmem {
  u64 keys[ST_TILE];
  u32 vals[ST_TILE];
  u32 warp_cnt[ST_WARPS][256];
  u32 tile_start[256];
  u32 g_off[256];
  u32 scan[40];
};

__global__ void __launch_bounds__(ST_THREADS, 6)
k_scatter(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs,
          const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in,
          u64 *__restrict__ keys_out, u32 *__restrict__ vals_out, int shift, const u32 *__restrict__ hist,
          const u32 *__restrict__ digit_base) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem &S = *reinterpret_cast<ScatterSmem *>(smem_raw);
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.na, off = job.pos_off;
  const u32 tid = threadIdx.x, w = warp_id(), l = lane_id();
  const u32 lt_mask = (1u << l) - 1u;
  for (int i = tid; i < ST_WARPS * 256; i += ST_THREADS) (&S.warp_cnt[0][0])[i] = 0;
  if (tid < 256) S.g_off[tid] = hist[(size_t)blockIdx.x * 256 + tid] + digit_base[(size_t)tl.job * 256 + tid];
  __syncthreads();
  u64 key[ST_ITEMS];
  u32 rk[ST_ITEMS];   // rank within (warp, digit) | digit << 16 ; 0xFFFFFFFF = invalid
  const u32 wbase = tl.start + w * ST_WCHUNK;
#pragma unroll
  for (int k = 0; k < ST_ITEMS; k++) {
    u32 i = wbase + k * 32 + l;
    key[k] = (i < n) ? keys_in[off + i] : 0;
  }
#pragma unroll
  for (int k = 0; k < ST_ITEMS; k++) {
    u32 i = wbase + k * 32 + l;
    bool valid = i < n;
    u32 d = (u32)(key[k] >> shift) & 255u;
    u32 mk = valid ? d : (256u + l);
    u32 peers = __match_any_sync(0xffffffffu, mk);
    u32 r = __popc(peers & lt_mask);
    u32 base = valid ? S.warp_cnt[w][d] : 0;
    __syncwarp();
    if (valid && r == 0) S.warp_cnt[w][d] = base + __popc(peers);
    __syncwarp();
    rk[k] = valid ? ((base + r) | (d << 16)) : 0xFFFFFFFFu;
  }
  __syncthreads();
  // per digit: exclusive over warps, tile count
  {
    u32 run = 0;
    if (tid < 256) {
#pragma unroll
      for (int ww = 0; ww < ST_WARPS; ww++) { u32 c = S.warp_cnt[ww][tid]; S.warp_cnt[ww][tid] = run; run += c; }
    }
    u32 ts = block_excl_add(run, S.scan, nullptr);
    if (tid < 256) S.tile_start[tid] = ts;
  }
  __syncthreads();
  // the rotation indices are only needed now: fetching them late keeps the register count low
  // enough for six resident CTAs per SM, whose phases overlap
  u32 val[ST_ITEMS];
#pragma unroll
  for (int k = 0; k < ST_ITEMS; k++) {
    u32 i = wbase + k * 32 + l;
    val[k] = (i < n) ? vals_in[off + i] : 0;
  }
#pragma unroll
  for (int k = 0; k < ST_ITEMS; k++) {
    if (rk[k] != 0xFFFFFFFFu) {
      u32 d = rk[k] >> 16;
      u32 lp = S.tile_start[d] + S.warp_cnt[w][d] + (rk[k] & 0xFFFFu);
      S.keys[lp] = key[k];
      S.vals[lp] = val[k];
    }
  }
  __syncthreads();
  const u32 cnt = (n - tl.start) < ST_TILE ? (n - tl.start) : ST_TILE;
#pragma unroll 4
  for (int k = 0; k < ST_ITEMS; k++) {
    u32 q = tid + k * ST_THREADS;
    if (q < cnt) {
      u64 kk = S.keys[q];
      u32 d = (u32)(kk >> shift) & 255u;
      u32 dst = off + S.g_off[d] + (q - S.tile_start[d]);
      keys_out[dst] = kk;
      vals_out[dst] = S.vals[q];
    }
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
__global__ void __launch_bounds__(ST_THREADS)
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
      rank[off + v] = headslot;
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

struct EvPair { cudaEvent_t a, b; };

// job_n[k] = post-RLE1 size of block job_ids[k]; the device copies have na == n on entry.
int b2k_bwt_batch(B2SortCtx *cx, cudaStream_t st, B2Job *d_jobs, const std::vector<u32> &job_ids,
                  const std::vector<u32> &job_n, const u8 *d_text, u8 *d_bwt) {
  // per device and cheap; set on every call so that handles on several devices / threads all have it
  B2_CUDA_CHECK(cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
  std::vector<B2SortTile> tiles;
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
  auto radix_pass = [&](int shift) -> int {
    u32 nt = (u32)tiles.size();
    k_hist<<<nt, ST_THREADS, 0, st>>>(cx->d_tiles, d_jobs, kA, shift, cx->d_hist);
    k_scan<<<(u32)sj.size(), 256, 0, st>>>(cx->d_sj, cx->d_hist, cx->d_digit_base);
    EvPair ev{nullptr, nullptr};
    if (cx->timing) { cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); cudaEventRecord(ev.a, st); }
    k_scatter<<<nt, ST_THREADS, sizeof(ScatterSmem), st>>>(cx->d_tiles, d_jobs, kA, vA, kB, vB, shift, cx->d_hist, cx->d_digit_base);
    if (cx->timing) { cudaEventRecord(ev.b, st); evs.push_back(ev); }
    B2_CUDA_CHECK(cudaGetLastError());
    std::swap(kA, kB); std::swap(vA, vB);
    u64 el = 0;
    for (u32 x : nas) el += x;
    cx->stats.scatter_launches++;
    cx->stats.scatter_elems += el;
    cx->stats.launches += 3;
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
    B2_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
  };

  int rc;
  if ((rc = upload())) return rc;
  {
    u64 el = 0; for (u32 x : ns) el += x;
    cx->stats.sorted_elems_round0 += el;
  }
  k_keys0<<<(u32)tiles.size(), ST_THREADS, 0, st>>>(cx->d_tiles, d_jobs, d_text, kA, vA);
  cx->stats.launches += 1;
  for (int p = 0; p < 8; p++) if ((rc = radix_pass(8 * p))) return rc;
  if ((rc = ranks(true))) return rc;
  cx->stats.rounds++;
  u64 reflect = 8;
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
