// Stage A7: move-to-front + zero-run coding (RUN_A/RUN_B), parallel over positions.
//
// Reference: zip_lib/bzip2-encoding.adb:318-413.  The reference keeps a 256-entry list, finds the
// symbol by linear search and shifts (:384-396).  The MTF index of position i equals the number of
// DISTINCT symbols seen since the previous occurrence of the same symbol; if there is none, it is
// (rank of the symbol among the used bytes) + (number of distinct symbols seen so far that are
// larger), because the list starts in increasing order (:374-376, :320-328).  That is a pure
// function of the data before i, so every position is computed independently (k_mtf_index).
//
// To keep the backward search short, k_mtf_masks first records which byte values occur in every
// 16-position segment (256-bit mask), every 256-position segment and every 4096-position tile; a
// search that leaves its own 16-segment skips whole segments / tiles by OR-ing their masks until it
// meets one that contains the symbol, descends, and only then walks bytes again.  Worst case per
// position: 2*15 byte steps + 4*15 segment masks + (tiles of the block) tile masks.
//
// The zero-run coding (:348-363, :400-410) is a max-scan (last non-zero position) plus an add-scan
// (output offsets) per block (k_rle2).
#include "b2_common.cuh"
#include "b2_kernels.h"

#define MI_THREADS 256
#define MI_ITEMS (B2_MTF_TILE / MI_THREADS)

struct Mask256 { u32 w[8]; };

__device__ __forceinline__ void mask_or(u32 *seen, const u32 *__restrict__ m) {
  const uint4 a = *reinterpret_cast<const uint4 *>(m), b = *reinterpret_cast<const uint4 *>(m + 4);
  seen[0] |= a.x; seen[1] |= a.y; seen[2] |= a.z; seen[3] |= a.w;
  seen[4] |= b.x; seen[5] |= b.y; seen[6] |= b.z; seen[7] |= b.w;
}
__device__ __forceinline__ void seen_set(u32 *seen, u32 x) {
  const u32 bit = 1u << (x & 31), wsel = x >> 5;
#pragma unroll
  for (int w = 0; w < 8; w++) seen[w] |= (w == (int)wsel) ? bit : 0u;
}

// One CTA per 4096-position tile; thread t owns the 16 positions [16t, 16t+16) of the tile:
// 256 masks of 16 positions, 16 masks of 256 positions, 1 tile mask.
__global__ void __launch_bounds__(MI_THREADS)
k_mtf_masks(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs, const u8 *__restrict__ bwt,
            u32 *__restrict__ m16, u32 *__restrict__ m256, u32 *__restrict__ tilemask) {
  __shared__ u32 tm[8];
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.n, off = job.pos_off;
  const u8 *d = bwt + off;
  const u32 tid = threadIdx.x, l = lane_id();
  if (tid < 8) tm[tid] = 0;
  __syncthreads();
  const u32 p0 = tl.start + tid * 16;
  u32 mk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (p0 < n) {
    const uint4 v = *reinterpret_cast<const uint4 *>(d + p0);      // pos_off is a multiple of 64
    const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (p0 + k < n) seen_set(mk, (w[k >> 2] >> (8 * (k & 3))) & 255u);
    }
    u32 *o = m16 + ((size_t)(off >> 4) + (p0 >> 4)) * 8;
    *reinterpret_cast<uint4 *>(o) = make_uint4(mk[0], mk[1], mk[2], mk[3]);
    *reinterpret_cast<uint4 *>(o + 4) = make_uint4(mk[4], mk[5], mk[6], mk[7]);
  }
  // 256-position masks: OR over the 16 threads of a half warp
  const u32 hm = (l < 16) ? 0x0000FFFFu : 0xFFFF0000u;
  u32 r[8];
#pragma unroll
  for (int q = 0; q < 8; q++) r[q] = __reduce_or_sync(hm, mk[q]);
  if ((l & 15) == 0 && p0 < n) {
    u32 *o = m256 + ((size_t)(off >> 8) + (p0 >> 8)) * 8;
#pragma unroll
    for (int q = 0; q < 8; q++) { o[q] = r[q]; if (r[q]) atomicOr(&tm[q], r[q]); }
  }
  __syncthreads();
  if (tid < 8) tilemask[(size_t)blockIdx.x * 8 + tid] = tm[tid];
}

// OR of `mine` (8 words, only meaningful on lanes with `take`) over the warp, added to seen[] on all lanes.
__device__ __forceinline__ void warp_or_into(u32 *seen, const u32 *mine, bool take) {
#pragma unroll
  for (int q = 0; q < 8; q++) seen[q] |= __reduce_or_sync(0xffffffffu, take ? mine[q] : 0u);
}

// Scans up to 32 masks downwards from index `from` (lane l looks at mask from - l, valid while >= lo).
// Returns the index of the nearest mask that contains b (or -1) and ORs all nearer masks into seen[].
__device__ __forceinline__ i32 warp_scan_masks(const u32 *__restrict__ masks, i32 from, i32 lo, u32 bw, u32 bbit, u32 *seen) {
  const u32 l = lane_id();
  const i32 s = from - (i32)l;
  const bool valid = s >= lo;
  u32 m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (valid) {
    const uint4 a = *reinterpret_cast<const uint4 *>(masks + (size_t)s * 8), c = *reinterpret_cast<const uint4 *>(masks + (size_t)s * 8 + 4);
    m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = c.x; m[5] = c.y; m[6] = c.z; m[7] = c.w;
  }
  u32 wsel = m[0];
#pragma unroll
  for (int q = 1; q < 8; q++) wsel = (bw == (u32)q) ? m[q] : wsel;
  const u32 hit = __ballot_sync(0xffffffffu, valid && (wsel & bbit));
  const u32 first = hit ? (u32)(__ffs(hit) - 1) : 32u;     // nearest mask containing b
  warp_or_into(seen, m, valid && l < first);
  return hit ? from - (i32)first : -1;
}

__global__ void __launch_bounds__(MI_THREADS)
k_mtf_index(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs, const u8 *__restrict__ bwt,
            const u32 *__restrict__ m16, const u32 *__restrict__ m256, const u32 *__restrict__ tilemask,
            u8 *__restrict__ idx_out) {
  __shared__ u16 hard[B2_MTF_TILE];
  __shared__ u32 n_hard;
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.n, off = job.pos_off;
  const u8 *d = bwt + off;
  const u32 *s16 = m16 + (size_t)(off >> 4) * 8;             // block-relative masks
  const u32 *s256 = m256 + (size_t)(off >> 8) * 8;
  const u32 *tmk = tilemask + (size_t)job.tile0 * 8;
  if (threadIdx.x == 0) n_hard = 0;
  __syncthreads();
  // ---- phase A: every thread, own 16-segment only ---------------------------------------------
  for (int k = 0; k < MI_ITEMS; k++) {
    const u32 i = tl.start + k * MI_THREADS + threadIdx.x;
    if (i >= n) continue;
    const u32 b = d[i];
    if (i > 0 && d[i - 1] == b) { idx_out[off + i] = 0; continue; }
    u32 seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool found = false;
    for (i32 j = (i32)i - 1; j >= (i32)(i & ~15u); j--) {
      const u32 x = d[j];
      if (x == b) { found = true; break; }
      seen_set(seen, x);
    }
    if (found) {
      u32 idx = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) idx += __popc(seen[w]);
      idx_out[off + i] = (u8)idx;
    } else {
      hard[atomicAdd(&n_hard, 1u)] = (u16)(i - tl.start);
    }
  }
  __syncthreads();
  // ---- phase B: one warp per remaining position, masks examined 32 at a time ------------------
  const u32 nh = n_hard, l = lane_id();
  for (u32 hI = warp_id(); hI < nh; hI += MI_THREADS / 32) {
    const u32 i = tl.start + hard[hI];
    const u32 b = d[i];
    const u32 bw = b >> 5, bbit = 1u << (b & 31);
    u32 seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u32 mine[8];
    // own 16-segment (b is known not to occur in it before i)
    {
      const u32 p = (i & ~15u) + l;
      const bool take = l < 16 && p < i;
#pragma unroll
      for (int q = 0; q < 8; q++) mine[q] = 0;
      if (take) seen_set(mine, d[p]);
      warp_or_into(seen, mine, take);
    }
    const i32 g16 = (i32)(i >> 4), g256 = (i32)(i >> 8), tile = (i32)(i >> 12);
    i32 hit16 = warp_scan_masks(s16, g16 - 1, g256 << 4, bw, bbit, seen);              // own 256-segment
    if (hit16 < 0) {
      i32 hit256 = warp_scan_masks(s256, g256 - 1, tile << 4, bw, bbit, seen);         // own tile
      if (hit256 < 0) {
        for (i32 t = tile - 1; t >= 0 && hit256 < 0; t -= 32) {                        // earlier tiles
          const i32 ht = warp_scan_masks(tmk, t, 0, bw, bbit, seen);
          if (ht >= 0) { hit256 = warp_scan_masks(s256, (ht << 4) + 15, ht << 4, bw, bbit, seen); break; }
        }
      }
      if (hit256 >= 0) hit16 = warp_scan_masks(s16, (hit256 << 4) + 15, hit256 << 4, bw, bbit, seen);
    }
    bool found = false;
    if (hit16 >= 0) {
      // bytes of the hit segment from its end downwards: lane l looks at byte 15 - l
      const u32 p = ((u32)hit16 << 4) + 15 - l;
      const u32 x = l < 16 ? d[p] : 256u;
      const u32 mt = __ballot_sync(0xffffffffu, x == b);
      const u32 first = (u32)(__ffs(mt) - 1);
      const bool take = l < first;
#pragma unroll
      for (int q = 0; q < 8; q++) mine[q] = 0;
      if (take) seen_set(mine, x);
      warp_or_into(seen, mine, take);
      found = true;
    }
    u32 idx = 0;
    if (found) {
#pragma unroll
      for (int w = 0; w < 8; w++) idx += __popc(seen[w]);
    } else {
      // rank of b among used bytes + seen symbols larger than b
      const u32 bb = b & 31;
#pragma unroll
      for (int w = 0; w < 8; w++) {
        u32 below = (w < (int)bw) ? 0xFFFFFFFFu : ((w == (int)bw) ? ((1u << bb) - 1u) : 0u);
        u32 above = (w > (int)bw) ? 0xFFFFFFFFu : ((w == (int)bw) ? (bb == 31 ? 0u : (0xFFFFFFFFu << (bb + 1))) : 0u);
        idx += __popc(job.in_use[w] & below) + __popc(seen[w] & above);
      }
    }
    if (l == 0) idx_out[off + i] = (u8)idx;
  }
}

#define R2_THREADS 1024
#define R2_ITEMS 4
#define R2_TILE (R2_THREADS * R2_ITEMS)

__global__ void __launch_bounds__(R2_THREADS, 1)
k_rle2(B2Job *jobs, const u8 *__restrict__ idx_in, u16 *__restrict__ mtf) {
  __shared__ i32 sm_i[40];
  __shared__ u32 sm_u[40];
  B2Job &job = jobs[blockIdx.x];
  const u32 n = job.n, off = job.pos_off;
  const u8 *ix = idx_in + off;
  u16 *out = mtf + job.mtf_off;
  const u32 tid = threadIdx.x;
  i32 carry_nz = -1;
  u32 out_base = 0;
  for (u32 t0 = 0; t0 < n; t0 += R2_TILE) {
    const u32 a = t0 + tid * R2_ITEMS;
    u32 v[R2_ITEMS];
    i32 lnz = -1;
#pragma unroll
    for (int k = 0; k < R2_ITEMS; k++) {
      v[k] = (a + k < n) ? ix[a + k] : 0;
      if (a + k < n && v[k]) lnz = (i32)(a + k);
    }
    i32 tot_nz;
    i32 prev_nz = block_excl_max(lnz, -1, sm_i, &tot_nz);
    prev_nz = max(prev_nz, carry_nz);
    u32 cnt = 0;
    {
      i32 p = prev_nz;
#pragma unroll
      for (int k = 0; k < R2_ITEMS; k++) {
        if (a + k < n && v[k]) {
          u32 run = (u32)((i32)(a + k) - 1 - p);
          cnt += 1 + (run ? (31 - __clz(run + 1)) : 0);
          p = (i32)(a + k);
        }
      }
    }
    u32 tile_total;
    u32 base = block_excl_add(cnt, sm_u, &tile_total);
    {
      u32 o = out_base + base;
      i32 p = prev_nz;
#pragma unroll
      for (int k = 0; k < R2_ITEMS; k++) {
        if (a + k < n && v[k]) {
          u32 run = (u32)((i32)(a + k) - 1 - p);
          if (run) {
            u32 rc = run + 1;                      // bijective base 2, RUN_A = 0, RUN_B = 1 (:348-363)
            do { out[o++] = (u16)(rc & 1); rc >>= 1; } while (rc >= 2);
          }
          out[o++] = (u16)(1 + v[k]);              // (:404)
          p = (i32)(a + k);
        }
      }
    }
    out_base += tile_total;
    carry_nz = max(carry_nz, tot_nz);
    __syncthreads();
  }
  if (tid == 0) {
    u32 o = out_base;
    u32 run = (u32)((i32)n - 1 - carry_nz);        // trailing zero run (:409)
    if (run) {
      u32 rc = run + 1;
      do { out[o++] = (u16)(rc & 1); rc >>= 1; } while (rc >= 2);
    }
    out[o++] = (u16)(job.n_used + 1);              // EOB (:330-335, :410)
    job.n_mtf = o;
    job.n_groups = 1 + (o - 1) / B2_GROUP_SIZE;    // selector_count (:968)
  }
}

int b2k_mtf(cudaStream_t st, B2Job *d_jobs, u32 n_jobs, const B2SortTile *d_tiles, u32 n_tiles,
            const u8 *d_bwt, u32 *d_m16, u32 *d_m256, u32 *d_tilemask, u8 *d_idx, u16 *d_mtf) {
  if (n_tiles) {
    k_mtf_masks<<<n_tiles, MI_THREADS, 0, st>>>(d_tiles, d_jobs, d_bwt, d_m16, d_m256, d_tilemask);
    k_mtf_index<<<n_tiles, MI_THREADS, 0, st>>>(d_tiles, d_jobs, d_bwt, d_m16, d_m256, d_tilemask, d_idx);
  }
  if (n_jobs) k_rle2<<<n_jobs, R2_THREADS, 0, st>>>(d_jobs, d_idx, d_mtf);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}
