// Stage A3: entropy segmentation (Segment_by_Entropy) on the device.
//
// Reference: zip_lib/data_segmentation.adb:39-105, instantiated with the two profiles of
// zip_lib/bzip2-encoding.adb:1262-1281.
#include "b2_common.cuh"
#include "b2_kernels.h"

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
    // bytes are loaded one batch ahead of their use (pre_n / pre_o), so that neither the byte loads
    // nor the table loads that depend on them stall the warp
    u32 pre_n = (SEG_WINDOW + (i32)l < len) ? buf[SEG_WINDOW + l] : 0u;
    u32 pre_o = (SEG_WINDOW + (i32)l < len) ? buf[l] : 0u;
    auto fetch = [&](i32 i0) {
      const i32 i = i0 + (i32)l;
      const bool valid = i < len;
      const u32 bn = pre_n;                                   // incoming byte
      const u32 bo = pre_o;                                   // outgoing byte
      {
        const i32 i2 = i + 32;
        const bool v2 = i2 < len;
        pre_n = v2 ? buf[i2] : 0u;
        pre_o = v2 ? buf[i2 - SEG_WINDOW] : 0u;
      }
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
      __syncwarp();
      if (valid) { atomicAdd(&F[bn], 1u); atomicSub(&F[bo], 1u); }
      __syncwarp();
    };
    fetch(SEG_WINDOW);
    for (i32 i0 = SEG_WINDOW; i0 < len; i0 += 32) {          // 0-based step index; reference i = index + 1
      const double c1 = n1, c2 = n2, c3 = n3, c4 = n4;        // terms of this batch; T[0] = 0 covers "p = 0" (:84-89)
      if (i0 + 32 < len) fetch(i0 + 32);                      // start fetching the next batch
      const int steps = min(32, len - i0);
      double mine = 0.0;
      bool done = false;
      // Exact integer replay.  While the running sum stays inside one binade [2^e, 2^(e+1)) it is an
      // integer multiple k of u = 2^(e-52) and fl (k u + d) = (k + rint (d / u)) u unless d / u lies
      // exactly half way between two integers (then the parity of k decides).  The 4 x 32 roundings of
      // a batch thus become an integer prefix sum; any batch that leaves the binade or meets a tie is
      // replayed with the sequential floating-point chain below.
      if (entropy >= 0.00390625 && entropy < 8.0) {
        const int e = ilogb(entropy);
        const double scale = ldexp(1.0, 52 - e), inv = ldexp(1.0, e - 52);
        const long long K0 = (long long)(entropy * scale);
        const bool v = (int)l < steps;
        const double x1 = v ? -c1 * scale : 0.0, x2 = v ? c2 * scale : 0.0, x3 = v ? -c3 * scale : 0.0, x4 = v ? c4 * scale : 0.0;
        const double r1 = rint(x1), r2 = rint(x2), r3 = rint(x3), r4 = rint(x4);
        const bool tie = fabs(x1 - r1) == 0.5 || fabs(x2 - r2) == 0.5 || fabs(x3 - r3) == 0.5 || fabs(x4 - r4) == 0.5;
        const long long p1 = (long long)r1, p2 = p1 + (long long)r2, p3 = p2 + (long long)r3, p4 = p3 + (long long)r4;
        long long incl = p4;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, incl, o); if (l >= (u32)o) incl += t; }
        const long long base = K0 + incl - p4;
        const long long lo = 1ll << 52, hi = 1ll << 53;
        const long long k1 = base + p1, k2 = base + p2, k3 = base + p3, k4 = base + p4;
        const bool bad = tie || k1 < lo || k1 >= hi || k2 < lo || k2 >= hi || k3 < lo || k3 >= hi || k4 < lo || k4 >= hi;
        if (__ballot_sync(0xffffffffu, v && bad) == 0) {
          mine = (double)k4 * inv;
          entropy = __shfl_sync(0xffffffffu, mine, steps - 1);
          done = true;
        }
      }
      if (!done) {
        __syncwarp();
        sE[l * 4 + 0] = c1; sE[l * 4 + 1] = c2; sE[l * 4 + 2] = c3; sE[l * 4 + 3] = c4;
        __syncwarp();
        for (int k = 0; k < steps; k++) {
          const double2 ab = *reinterpret_cast<const double2 *>(&sE[k * 4]);
          const double2 cd = *reinterpret_cast<const double2 *>(&sE[k * 4 + 2]);
          entropy = __dsub_rn(entropy, ab.x);                 // data_segmentation.adb:75-90
          entropy = __dadd_rn(entropy, ab.y);
          entropy = __dsub_rn(entropy, cd.x);
          entropy = __dadd_rn(entropy, cd.y);
          if ((int)l == k) mine = entropy;
        }
        __syncwarp();
      }
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

int b2k_segment(cudaStream_t st, const u8 *d_in, const B2Chunk *d_chunks, u32 n_chunks, const double *d_T,
                u32 *d_seg, u32 *d_nseg) {
  if (n_chunks == 0) return 0;
  k_segment<<<n_chunks, 32, 0, st>>>(d_in, d_chunks, n_chunks, d_T, d_seg, d_nseg);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}
