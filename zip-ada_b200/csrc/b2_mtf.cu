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
// 64-position segment (256-bit mask) and in every 4096-position tile; a search that leaves its own
// segment skips whole segments / tiles by OR-ing their masks until it meets one that contains the
// symbol, and only then walks bytes again.  Worst case per position: 2*64 byte steps + 2*63
// segment masks + (tiles of the block) tile masks.
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

// One CTA per 4096-position tile: 64 segment masks + the tile mask.
__global__ void __launch_bounds__(MI_THREADS)
k_mtf_masks(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs, const u8 *__restrict__ bwt,
            u32 *__restrict__ segmask, u32 *__restrict__ tilemask) {
  __shared__ u32 tm[8];
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.n, off = job.pos_off;
  const u8 *d = bwt + off;
  const u32 w = warp_id(), l = lane_id();
  if (threadIdx.x < 8) tm[threadIdx.x] = 0;
  __syncthreads();
  u32 acc_tile = 0;   // lane q < 8 accumulates word q of the tile mask over this warp's segments
  for (int sgi = 0; sgi < 8; sgi++) {
    const u32 s = (tl.start >> 6) + w * 8 + sgi;       // block-relative segment
    const u32 p0 = s << 6;
    if (p0 >= n) break;
    u32 mine = 0;                                      // lane q < 8 holds word q of the segment mask
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const u32 p = p0 + h * 32 + l;
      const bool v = p < n;
      const u32 b = v ? d[p] : 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        u32 r = __reduce_or_sync(0xffffffffu, (v && (b >> 5) == (u32)q) ? (1u << (b & 31)) : 0u);
        if (l == (u32)q) mine |= r;
      }
    }
    if (l < 8) { segmask[((size_t)(off >> 6) + s) * 8 + l] = mine; acc_tile |= mine; }
  }
  if (l < 8 && acc_tile) atomicOr(&tm[l], acc_tile);
  __syncthreads();
  if (threadIdx.x < 8) tilemask[(size_t)blockIdx.x * 8 + threadIdx.x] = tm[threadIdx.x];
}

__global__ void __launch_bounds__(MI_THREADS)
k_mtf_index(const B2SortTile *__restrict__ tiles, const B2Job *__restrict__ jobs, const u8 *__restrict__ bwt,
            const u32 *__restrict__ segmask, const u32 *__restrict__ tilemask, u8 *__restrict__ idx_out) {
  const B2SortTile tl = tiles[blockIdx.x];
  const B2Job &job = jobs[tl.job];
  const u32 n = job.n, off = job.pos_off;
  const u8 *d = bwt + off;
  const u32 *sm = segmask + (size_t)(off >> 6) * 8;          // block-relative segment masks
  const u32 *tmk = tilemask + (size_t)job.tile0 * 8;         // block-relative tile masks
  for (int k = 0; k < MI_ITEMS; k++) {
    const u32 i = tl.start + k * MI_THREADS + threadIdx.x;
    if (i >= n) continue;
    const u32 b = d[i];
    u32 idx;
    if (i > 0 && d[i - 1] == b) idx = 0;
    else {
      u32 seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const u32 bw = b >> 5, bbit = 1u << (b & 31);
      bool found = false;
      // 1. own segment, byte by byte
      const i32 seg = (i32)(i >> 6);
      for (i32 j = (i32)i - 1; j >= (seg << 6); j--) {
        u32 x = d[j];
        if (x == b) { found = true; break; }
        seen_set(seen, x);
      }
      i32 hit_seg = -1;
      if (!found) {
        // 2. earlier segments of the own tile
        const i32 tile = (i32)(i >> 12);
        for (i32 s = seg - 1; s >= (tile << 6); s--) {
          const u32 *m = sm + (size_t)s * 8;
          if (m[bw] & bbit) { hit_seg = s; break; }
          mask_or(seen, m);
        }
        // 3. earlier tiles
        if (hit_seg < 0) {
          for (i32 t = tile - 1; t >= 0; t--) {
            const u32 *m = tmk + (size_t)t * 8;
            if (m[bw] & bbit) {
              for (i32 s = (t << 6) + 63; s >= (t << 6); s--) {
                const u32 *ms = sm + (size_t)s * 8;
                if (ms[bw] & bbit) { hit_seg = s; break; }
                mask_or(seen, ms);
              }
              break;
            }
            mask_or(seen, m);
          }
        }
        if (hit_seg >= 0) {
          for (i32 j = (hit_seg << 6) + 63; j >= (hit_seg << 6); j--) {
            u32 x = d[j];
            if (x == b) { found = true; break; }
            seen_set(seen, x);
          }
        }
      }
      if (found) {
        idx = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) idx += __popc(seen[w]);
      } else {
        // rank of b among used bytes + seen symbols larger than b
        idx = 0;
        const u32 bb = b & 31;
#pragma unroll
        for (int w = 0; w < 8; w++) {
          u32 below = (w < (int)bw) ? 0xFFFFFFFFu : ((w == (int)bw) ? ((1u << bb) - 1u) : 0u);
          u32 above = (w > (int)bw) ? 0xFFFFFFFFu : ((w == (int)bw) ? (bb == 31 ? 0u : (0xFFFFFFFFu << (bb + 1))) : 0u);
          idx += __popc(job.in_use[w] & below) + __popc(seen[w] & above);
        }
      }
    }
    idx_out[off + i] = (u8)idx;
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
            const u8 *d_bwt, u32 *d_segmask, u32 *d_tilemask, u8 *d_idx, u16 *d_mtf) {
  if (n_tiles) {
    k_mtf_masks<<<n_tiles, MI_THREADS, 0, st>>>(d_tiles, d_jobs, d_bwt, d_segmask, d_tilemask);
    k_mtf_index<<<n_tiles, MI_THREADS, 0, st>>>(d_tiles, d_jobs, d_bwt, d_segmask, d_tilemask, d_idx);
  }
  if (n_jobs) k_rle2<<<n_jobs, R2_THREADS, 0, st>>>(d_jobs, d_idx, d_mtf);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}
