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
// k_segment: one WARP per chunk replays the running FP64 entropy sum.  Both profiles
// (bzip2-encoding.adb:1262-1264) share window_size = 16_000, hence the same entropy series;
// they differ in thresholds and marks only.  T[c] = -(c/16000)*ln(c/16000) is built on the
// host with glibc `log`.  The sum is order dependent, so it is replayed
// sequentially with round-to-nearest adds; what the warp parallelises is everything around it:
// for 32 consecutive steps the lanes derive the window counts of the incoming / outgoing bytes
// (integer, exact) and fetch the four table terms, then all lanes run the same 4-add chain.
// ---------------------------------------------------------------------------------------------
#define SEG_WINDOW 16000

__global__ void __launch_bounds__(32)
k_segment(const u8 *__restrict__ in, const B2Chunk *__restrict__ chunks, u32 n_chunks,
          const double *__restrict__ T, u32 *__restrict__ seg, u32 *__restrict__ nseg) {
  __shared__ u32 F[256];
  __shared__ __align__(16) double sE[32 * 4];
  __shared__ __align__(16) u8 sBn[32];
  __shared__ __align__(16) u8 sBo[32];
  const u32 *sBn32 = reinterpret_cast<const u32 *>(sBn);
  const u32 *sBo32 = reinterpret_cast<const u32 *>(sBo);
  const u32 c = blockIdx.x;
  if (c >= n_chunks) return;
  const u32 l = threadIdx.x;
  const u32 lt = (1u << l) - 1u;
  const u8 *buf = in + chunks[c].start;
  const i32 len = (i32)chunks[c].len;
  const double thr[2] = {(double)0.6f, (double)0.4f};    // Float generic formal widened (data_segmentation.ads:43)
  const i32 ithr[2] = {4000, 8000};
  i32 index_mark[2] = {1, 1};
  double mark[2] = {0.0, 0.0};
  u32 cnt[2] = {0, 0};
  u32 *out[2] = {seg + (size_t)(2 * c) * B2_MAX_SEG, seg + (size_t)(2 * c + 1) * B2_MAX_SEG};
  const bool act[2] = {len > SEG_WINDOW + ithr[0], len > SEG_WINDOW + ithr[1]};
  if (act[0] || act[1]) {
    for (int b = l; b < 256; b += 32) F[b] = 0;
    __syncwarp();
    for (i32 i = l; i < SEG_WINDOW; i += 32) atomicAdd(&F[buf[i]], 1u);
    __syncwarp();
    // initial entropy, b = 0 .. 255 in order (data_segmentation.adb:63-72)
    double entropy = 0.0;
    for (int g = 0; g < 8; g++) {
      u32 f = F[g * 32 + l];
      double tv = T[f];
      for (int k = 0; k < 32; k++) {
        u32 fk = __shfl_sync(0xffffffffu, f, k);
        double tk = __shfl_sync(0xffffffffu, tv, k);
        if (fk > 0) entropy = __dadd_rn(entropy, tk);
      }
    }
    mark[0] = entropy; mark[1] = entropy;
    // byte-prefix masks for "count bytes j < l" / "j <= l" over a 32-byte row held as 8 words
    u32 pm_lt[8], pm_le[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int lo = 4 * q;
      const int a_lt = (int)l - lo, a_le = (int)l + 1 - lo;
      pm_lt[q] = a_lt >= 4 ? 0xFFFFFFFFu : (a_lt <= 0 ? 0u : ((1u << (8 * a_lt)) - 1u));
      pm_le[q] = a_le >= 4 ? 0xFFFFFFFFu : (a_le <= 0 ? 0u : ((1u << (8 * a_le)) - 1u));
    }
    // Software pipeline: while the 4-add chain of batch t runs out of shared memory, the table terms
    // of batch t+1 are already being fetched into registers.
    double n1 = 0, n2 = 0, n3 = 0, n4 = 0;
    u32 g4_next = 0, g4_cur = 0;
    auto fetch = [&](i32 i0) {
      const i32 i = i0 + (i32)l;
      const bool valid = i < len;
      const u32 bn = valid ? buf[i] : 0u;                     // incoming byte
      const u32 bo = valid ? buf[i - SEG_WINDOW] : 0u;        // outgoing byte
      sBn[l] = (u8)bn; sBo[l] = (u8)bo;
      __syncwarp();
      const u32 key_n = valid ? bn : (256u + l), key_o = valid ? bo : (512u + l);
      const u32 c_nn = __popc(__match_any_sync(0xffffffffu, key_n) & lt);   // #{j<k : bn_j = bn_k}
      const u32 c_oo = __popc(__match_any_sync(0xffffffffu, key_o) & lt);   // #{j<k : bo_j = bo_k}
      const u32 sp_n = bn * 0x01010101u, sp_o = bo * 0x01010101u;
      u32 c_on = 0, c_no = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        c_on += __popc(__vcmpeq4(sBo32[q], sp_n) & pm_lt[q]);   // #{j<k  : bo_j = bn_k}
        c_no += __popc(__vcmpeq4(sBn32[q], sp_o) & pm_le[q]);   // #{j<=k : bn_j = bo_k}
      }
      c_on >>= 3; c_no >>= 3;
      u32 fnb = 0, fob = 1;
      if (valid) { fnb = F[bn] + c_nn - c_on; fob = F[bo] + c_no - c_oo; }
      const u32 f4 = fob - 1;
      n1 = T[fnb]; n2 = T[fnb + 1]; n3 = T[fob]; n4 = T[f4];
      g4_next = __ballot_sync(0xffffffffu, valid && f4 > 0);
      __syncwarp();
      if (valid) { atomicAdd(&F[bn], 1u); atomicSub(&F[bo], 1u); }
      __syncwarp();
    };
    fetch(SEG_WINDOW);
    for (i32 i0 = SEG_WINDOW; i0 < len; i0 += 32) {          // 0-based step index; reference i = index + 1
      // publish the terms of this batch, start fetching the next one
      sE[l * 4 + 0] = n1; sE[l * 4 + 1] = n2; sE[l * 4 + 2] = n3; sE[l * 4 + 3] = n4;
      g4_cur = g4_next;
      __syncwarp();
      if (i0 + 32 < len) fetch(i0 + 32);
      const int steps = min(32, len - i0);
      double mine = 0.0;
      for (int k = 0; k < steps; k++) {
        const double2 ab = *reinterpret_cast<const double2 *>(&sE[k * 4]);
        const double2 cd = *reinterpret_cast<const double2 *>(&sE[k * 4 + 2]);
        entropy = __dsub_rn(entropy, ab.x);                   // data_segmentation.adb:75-90
        entropy = __dadd_rn(entropy, ab.y);
        entropy = __dsub_rn(entropy, cd.x);
        const double e4 = __dadd_rn(entropy, cd.y);
        entropy = ((g4_cur >> k) & 1u) ? e4 : entropy;
        if ((int)l == k) mine = entropy;
      }
      __syncwarp();
      // threshold tests for the 32 steps at once (:91-97).  A cut moves index_mark to within 32 of
      // every later step of the batch, so at most one cut per profile can happen in a batch.
      const i32 sp = i0 + (i32)l + 1 - SEG_WINDOW;
#pragma unroll
      for (int p = 0; p < 2; p++) {
        if (act[p] && (i0 + steps - SEG_WINDOW - index_mark[p] > ithr[p])) {
          const bool cond = ((int)l < steps) && fabs(__dsub_rn(mine, mark[p])) > thr[p] && (sp - index_mark[p] > ithr[p]);
          const u32 m = __ballot_sync(0xffffffffu, cond);
          if (m) {
            const int k = __ffs(m) - 1;
            const i32 seg_point = i0 + k + 1 - SEG_WINDOW;
            if (l == 0 && cnt[p] < B2_MAX_SEG - 1) out[p][cnt[p]] = (u32)seg_point;
            cnt[p]++;
            index_mark[p] = seg_point;
            mark[p] = __shfl_sync(0xffffffffu, mine, k);
          }
        }
      }
    }
  }
  if (l == 0) {
    for (int k = 0; k < 2; k++) {
      if (len > 0) { if (cnt[k] < B2_MAX_SEG) out[k][cnt[k]] = (u32)len; cnt[k]++; }   // :102-104
      nseg[2 * c + k] = cnt[k];
    }
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
  k_segment<<<n_chunks, 32, 0, st>>>(d_in, d_chunks, n_chunks, d_T, d_seg, d_nseg);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}
