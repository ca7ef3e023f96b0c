// Stage A1 (chunk cutting) and A3 (entropy segmentation) on the device.
//
// Reference: zip_lib/bzip2-encoding.adb:1160-1208 (Data_Acquisition), :1413-1429 (chunk loop,
// last-two-blocks balancing) and zip_lib/data_segmentation.adb:39-105 (Segment_by_Entropy).
//
// k_cut: the chunk chain is serial (chunk k+1 starts where chunk k ended) but each link is a
// scan: a "piece" is a maximal equal-byte run counted from the chunk start, split every 259;
// when the first byte of a new piece is read, the previous piece's RLE1 size min(len,4)+[len>=4]
// is committed; reading stops after the first byte whose commit makes committed+5 >= capacity
// (SURVEY.md §9 R4).  One persistent CTA walks the whole input tile by tile.
#include "b2_common.cuh"
#include "b2_kernels.h"

#define CUT_THREADS 1024
#define CUT_BYTES 32                       // bytes per thread per tile
#define CUT_TILE (CUT_THREADS * CUT_BYTES)

__device__ __forceinline__ u32 enc_size(u32 len) { return (len < 4 ? len : 4) + (len >= 4 ? 1 : 0); }

__global__ void __launch_bounds__(CUT_THREADS, 1)
k_cut(const u8 *__restrict__ in, u64 n, i64 size_hint, int level, i64 win_lo, i64 win_hi,
      B2Chunk *chunks, u32 *n_chunks, u32 max_chunks) {
  __shared__ i32 sm_scan[40];
  __shared__ u32 sm_scan_u[40];
  __shared__ u32 sm_cross;
  __shared__ i64 sm_cap;
  __shared__ u64 sm_limit;
  const u32 tid = threadIdx.x;
  u64 pos = 0;          // bytes read so far == start of the current chunk
  u32 nc = 0;
  for (;;) {
    if (tid == 0) {
      // stream_rest: size_hint - bytes read, sticking at -1 (= unknown_size) once it gets there
      // (bzip2-encoding.adb:1192-1194); balancing test :1416-1424 done in float32 on the host
      // and passed as the integer window [win_lo, win_hi].
      i64 rest = size_hint < 0 ? -1 : ((i64)pos <= size_hint ? size_hint - (i64)pos : -1);
      i64 cap = (i64)level * 100000;
      if (rest >= win_lo && rest <= win_hi) cap = rest / 2;
      sm_cap = cap;
      u64 lim = n - pos;
      u64 rawmax = (u64)(10 * cap);                      // multiplier = 10 (:1156)
      sm_limit = lim < rawmax ? lim : rawmax;
      sm_cross = 0xFFFFFFFFu;
    }
    __syncthreads();
    const i64 cap = sm_cap;
    const u32 limit = (u32)sm_limit;                      // <= 9_000_000
    const u32 need = cap > 5 ? (u32)(cap - 5) : 0;        // stop once committed >= cap - 5
    u32 len = limit;
    if (need == 0) len = 0;                               // loop condition false before the first read
    else {
      const u8 *src = in + pos;                           // chunk-relative addressing
      u32 committed = 0;                                  // committed size before this tile
      i32 carry_r = -1;                                   // run start (relative) of the last byte of the previous tile
      for (u32 t0 = 0; t0 < limit; t0 += CUT_TILE) {
        const u32 a = t0 + tid * CUT_BYTES;               // first relative position of this thread
        u8 b[CUT_BYTES];
        u8 prev = 0;
#pragma unroll
        for (int k = 0; k < CUT_BYTES; k++) b[k] = (a + k < limit) ? src[a + k] : 0;
        if (a > 0 && a < limit) prev = src[a - 1];
        // phase 1: last change position in my segment
        i32 lc = -1;
        {
          u8 pv = prev;
#pragma unroll
          for (int k = 0; k < CUT_BYTES; k++) {
            u32 p = a + k;
            if (p < limit && (p == 0 || b[k] != pv)) lc = (i32)p;
            pv = b[k];
          }
        }
        i32 tot_lc;
        i32 r_in = block_excl_max(lc, -1, sm_scan, &tot_lc);
        r_in = max(r_in, carry_r);
        // phase 2: commits
        u32 run[CUT_BYTES];
        u32 local = 0;
        {
          i32 r = r_in;
          u8 pv = prev;
#pragma unroll
          for (int k = 0; k < CUT_BYTES; k++) {
            u32 p = a + k;
            if (p < limit) {
              bool chg = (p == 0 || b[k] != pv);
              u32 c = 0;
              if (p > 0) {
                u32 mprev = (u32)((i32)(p - 1) - r) % 259u;          // index in piece of byte p-1
                if (chg) c = enc_size(mprev + 1);
                else if (mprev == 258) c = 5;                        // forced break at 259 (:1199)
              }
              if (chg) r = (i32)p;
              local += c;
            }
            run[k] = local;
            pv = b[k];
          }
        }
        u32 tile_total;
        u32 base = block_excl_add(local, sm_scan_u, &tile_total);
        // phase 3: first position where committed + 5 >= cap
        if (committed + tile_total >= need) {
          u32 before = committed + base;
          if (before + local >= need) {
#pragma unroll
            for (int k = 0; k < CUT_BYTES; k++) {
              if (a + k < limit && before + run[k] >= need) { atomicMin(&sm_cross, a + k); break; }
            }
          }
        }
        __syncthreads();
        u32 cross = sm_cross;
        if (cross != 0xFFFFFFFFu) { len = cross + 1; break; }
        committed += tile_total;
        carry_r = max(carry_r, tot_lc);
        __syncthreads();
      }
    }
    if (tid == 0) {
      if (nc < max_chunks) { chunks[nc].start = pos; chunks[nc].len = len; chunks[nc].cap = (u32)cap; chunks[nc].pad = 0; }
    }
    nc++;
    pos += len;
    __syncthreads();
    if (pos >= n) break;                                  // exit when not More_Bytes (:1428)
    if (len == 0) break;                                  // cannot make progress (cap <= 5): not reachable for levels 1..9
  }
  if (tid == 0) *n_chunks = nc;
}

// ---------------------------------------------------------------------------------------------
// k_segment: one thread per chunk replays the running FP64 entropy sum.  Both profiles
// (bzip2-encoding.adb:1262-1264) share window_size = 16_000, hence the same entropy series;
// they differ in thresholds and marks only.  T[c] = -(c/16000)*ln(c/16000) is built on the
// host with glibc `log`; the sum itself is order dependent and is replayed
// with round-to-nearest adds, no contraction.
// ---------------------------------------------------------------------------------------------
#define SEG_THREADS 4
#define SEG_WINDOW 16000

__global__ void __launch_bounds__(SEG_THREADS)
k_segment(const u8 *__restrict__ in, const B2Chunk *__restrict__ chunks, u32 n_chunks,
          const double *__restrict__ T, u32 *__restrict__ seg, u32 *__restrict__ nseg) {
  __shared__ u16 freq_s[256 * SEG_THREADS];
  const u32 c = blockIdx.x * SEG_THREADS + threadIdx.x;
  if (c >= n_chunks) return;
  u16 *freq = freq_s + threadIdx.x;                      // stride SEG_THREADS
  for (int b = 0; b < 256; b++) freq[b * SEG_THREADS] = 0;
  const u8 *buf = in + chunks[c].start;
  const i32 len = (i32)chunks[c].len;
  const double thr[2] = {(double)0.6f, (double)0.4f};    // Float generic formal widened (data_segmentation.ads:43)
  const i32 ithr[2] = {4000, 8000};
  bool act[2];
  i32 index_mark[2] = {1, 1};
  double mark[2] = {0.0, 0.0};
  u32 cnt[2] = {0, 0};
  u32 *out[2] = {seg + (size_t)(2 * c) * B2_MAX_SEG, seg + (size_t)(2 * c + 1) * B2_MAX_SEG};
  act[0] = len > SEG_WINDOW + ithr[0];
  act[1] = len > SEG_WINDOW + ithr[1];
  if (act[0] || act[1]) {
    double entropy = 0.0;
    for (i32 i = 1; i <= len; i++) {
      u32 bt = buf[i - 1];
      u32 f = (u32)freq[bt * SEG_THREADS] + 1;
      freq[bt * SEG_THREADS] = (u16)f;
      if (i == SEG_WINDOW) {
        for (int b = 0; b < 256; b++) {
          u32 fb = freq[b * SEG_THREADS];
          if (fb > 0) entropy = __dadd_rn(entropy, T[fb]);
        }
        mark[0] = entropy; mark[1] = entropy;
      } else if (i > SEG_WINDOW) {
        entropy = __dsub_rn(entropy, T[f - 1]);
        entropy = __dadd_rn(entropy, T[f]);
        u32 bo = buf[i - SEG_WINDOW - 1];
        u32 g = freq[bo * SEG_THREADS];
        entropy = __dsub_rn(entropy, T[g]);
        g--;
        freq[bo * SEG_THREADS] = (u16)g;
        if (g > 0) entropy = __dadd_rn(entropy, T[g]);
#pragma unroll
        for (int k = 0; k < 2; k++) {
          if (act[k] && fabs(__dsub_rn(entropy, mark[k])) > thr[k]) {
            i32 seg_point = i - SEG_WINDOW;
            if (seg_point - index_mark[k] > ithr[k]) {
              if (cnt[k] < B2_MAX_SEG - 1) out[k][cnt[k]] = (u32)seg_point;
              cnt[k]++;
              index_mark[k] = seg_point;
              mark[k] = entropy;
            }
          }
        }
      }
    }
  }
  for (int k = 0; k < 2; k++) {
    if (len > 0) { if (cnt[k] < B2_MAX_SEG) out[k][cnt[k]] = (u32)len; cnt[k]++; }
    nseg[2 * c + k] = cnt[k];
  }
}

// ---------------------------------------------------------------------------------------------
int b2k_cut(cudaStream_t st, const u8 *d_in, u64 n, i64 size_hint, int level, i64 win_lo, i64 win_hi,
            B2Chunk *d_chunks, u32 *d_n_chunks, u32 max_chunks) {
  k_cut<<<1, CUT_THREADS, 0, st>>>(d_in, n, size_hint, level, win_lo, win_hi, d_chunks, d_n_chunks, max_chunks);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int b2k_segment(cudaStream_t st, const u8 *d_in, const B2Chunk *d_chunks, u32 n_chunks, const double *d_T,
                u32 *d_seg, u32 *d_nseg) {
  if (n_chunks == 0) return 0;
  k_segment<<<(n_chunks + SEG_THREADS - 1) / SEG_THREADS, SEG_THREADS, 0, st>>>(d_in, d_chunks, n_chunks, d_T, d_seg, d_nseg);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}
