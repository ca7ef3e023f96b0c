// Stage A3: entropy segmentation (Segment_by_Entropy) on the device.
//
// Reference: zip_lib/data_segmentation.adb:39-105, instantiated with the two profiles of
// zip_lib/bzip2-encoding.adb:1262-1281.
//
// One CTA per chunk replays the running FP64 entropy sum.  Both profiles share window_size =
// 16_000, hence the same entropy series; they differ in thresholds and marks only.
// T[c] = -(c/16000)*ln(c/16000) is built on the host with glibc `log`.
//
// The sum is order dependent (four roundings per input byte, :75-90).  It is replayed 256 steps
// at a time:
//  * the window counts met by the incoming / outgoing byte of every step are derived in parallel
//    (integer, exact): counts at the batch start + occurrences earlier in the batch;
//  * while the running sum stays inside one binade [2^e, 2^(e+1)) it is an integer multiple k of
//    u = 2^(e-52), and fl (k u + d) = (k + rint (d / u)) u; when d / u lies exactly half way between
//    two integers (frequent: d has only a few more bits than u) round-half-even makes the result
//    depend on the parity of k, so every term is a map on k that only looks at k's parity; such
//    maps compose associatively and the 4 x 256 roundings of a batch become a prefix scan over the CTA;
//  * a batch that leaves the binade is replayed by one warp with the literal sequential chain of
//    round-to-nearest adds.
#include "b2_common.cuh"
#include "b2_kernels.h"

#define SEG_WINDOW 16000
#define SG_THREADS 256
#define SG_WARPS (SG_THREADS / 32)

// k -> k + (k even ? ae : ao)
struct PF { long long ae, ao; };
__device__ __forceinline__ PF pf_term(long long m, bool tie) {
  PF f;
  f.ae = m + ((tie && (m & 1)) ? 1 : 0);
  f.ao = m + ((tie && ((m + 1) & 1)) ? 1 : 0);
  return f;
}
__device__ __forceinline__ PF pf_compose(const PF &f, const PF &g) {   // first f, then g
  PF r;
  r.ae = f.ae + ((f.ae & 1) ? g.ao : g.ae);
  r.ao = f.ao + (((1 + f.ao) & 1) ? g.ao : g.ae);
  return r;
}

struct SegSmem {
  u32 F[256];                      // window counts at the batch start
  u32 Hn[SG_WARPS][256];           // occurrences of each byte among the incoming bytes of a warp
  u32 Ho[SG_WARPS][256];           // ... among the outgoing bytes
  u16 Pn[SG_WARPS][256];           // the same, summed over the warps before
  u16 Po[SG_WARPS][256];
  double terms[SG_THREADS * 4];    // fallback: the four table terms of every step
  double series[SG_THREADS];       // fallback: the sum after every step
  long long wtot_e[SG_WARPS], wtot_o[SG_WARPS];
  double cutval[2];
  u32 first[2];
  u8 bn[SG_THREADS], bo[SG_THREADS];
  double entropy;
};

__global__ void __launch_bounds__(SG_THREADS)
k_segment(const u8 *__restrict__ in, const B2Chunk *__restrict__ chunks, u32 n_chunks,
          const double *__restrict__ T, u32 *__restrict__ seg, u32 *__restrict__ nseg) {
  __shared__ SegSmem S;
  const u32 c = blockIdx.x;
  if (c >= n_chunks) return;
  const u32 tid = threadIdx.x, l = lane_id(), w = warp_id();
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
    S.F[tid] = 0;
    for (int i = tid; i < SG_WARPS * 256; i += SG_THREADS) { (&S.Hn[0][0])[i] = 0; (&S.Ho[0][0])[i] = 0; }
    __syncthreads();
    for (i32 i = tid; i < SEG_WINDOW; i += SG_THREADS) atomicAdd(&S.F[buf[i]], 1u);
    __syncthreads();
    // initial entropy, b = 0 .. 255 in order (data_segmentation.adb:63-72), by warp 0
    if (w == 0) {
      double e = 0.0;
      for (int g = 0; g < 8; g++) {
        const u32 f = S.F[g * 32 + l];
        const double tv = T[f];
        for (int k = 0; k < 32; k++) {
          const u32 fk = __shfl_sync(0xffffffffu, f, k);
          const double tk = __shfl_sync(0xffffffffu, tv, k);
          if (fk > 0) e = __dadd_rn(e, tk);
        }
      }
      if (l == 0) S.entropy = e;
    }
    __syncthreads();
    double entropy = S.entropy;
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
    const u32 *rown = reinterpret_cast<const u32 *>(S.bn + 32 * w);
    const u32 *rowo = reinterpret_cast<const u32 *>(S.bo + 32 * w);
    // bytes are loaded one batch ahead of their use
    u32 pre_n = (SEG_WINDOW + (i32)tid < len) ? buf[SEG_WINDOW + tid] : 0u;
    u32 pre_o = (SEG_WINDOW + (i32)tid < len) ? buf[tid] : 0u;
    for (i32 i0 = SEG_WINDOW; i0 < len; i0 += SG_THREADS) {   // 0-based step index; reference i = index + 1
      const i32 i = i0 + (i32)tid;
      const bool valid = i < len;
      const u32 bn = pre_n, bo = pre_o;                       // incoming / outgoing byte of my step
      {
        const i32 i2 = i + SG_THREADS;
        const bool v2 = i2 < len;
        pre_n = v2 ? buf[i2] : 0u;
        pre_o = v2 ? buf[i2 - SEG_WINDOW] : 0u;
      }
      const int steps = min(SG_THREADS, len - i0);
      S.bn[tid] = (u8)bn; S.bo[tid] = (u8)bo;
      if (tid < 2) S.first[tid] = 0xFFFFFFFFu;
      if (valid) { atomicAdd(&S.Hn[w][bn], 1u); atomicAdd(&S.Ho[w][bo], 1u); }
      __syncthreads();
      {                                                       // per byte value: occurrences in the warps before
        u32 rn = 0, ro = 0;
#pragma unroll
        for (int ww = 0; ww < SG_WARPS; ww++) {
          S.Pn[ww][tid] = (u16)rn; S.Po[ww][tid] = (u16)ro;
          rn += S.Hn[ww][tid]; ro += S.Ho[ww][tid];
        }
      }
      __syncthreads();
      // occurrences earlier in my own warp
      const u32 key_n = valid ? bn : (256u + l), key_o = valid ? bo : (512u + l);
      const u32 c_nn = __popc(__match_any_sync(0xffffffffu, key_n) & lt);   // #{j<k : bn_j = bn_k}
      const u32 c_oo = __popc(__match_any_sync(0xffffffffu, key_o) & lt);   // #{j<k : bo_j = bo_k}
      const u32 sp_n = bn * 0x01010101u, sp_o = bo * 0x01010101u;
      u32 c_on = 0, c_no = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        c_on += __popc(__vcmpeq4(rowo[q], sp_n) & pm_lt[q]);   // #{j<k  : bo_j = bn_k}
        c_no += __popc(__vcmpeq4(rown[q], sp_o) & pm_le[q]);   // #{j<=k : bn_j = bo_k}
      }
      c_on >>= 3; c_no >>= 3;
      // invalid lanes hold byte 0 in the rows: they are the LAST lanes of the last warp, so they are
      // never counted by a valid lane (prefix masks), and their own results are discarded
      u32 fnb = 0, fob = 1;
      if (valid) {
        fnb = S.F[bn] + (S.Pn[w][bn] + c_nn) - (S.Po[w][bn] + c_on);        // count met by the incoming byte
        fob = S.F[bo] + (S.Pn[w][bo] + c_no) - (S.Po[w][bo] + c_oo);        // count met by the outgoing byte
      }
      const double c1 = T[fnb], c2 = T[fnb + 1], c3 = T[fob], c4 = T[fob - 1];   // T[0] = 0 covers "p = 0" (:84-89)
      double mine = 0.0;
      bool bad = true;
      const bool try_int = entropy >= 0.00390625 && entropy < 8.0;
      if (try_int) {
        const int e = ilogb(entropy);
        const double scale = ldexp(1.0, 52 - e), inv = ldexp(1.0, e - 52);
        const long long K0 = (long long)(entropy * scale);
        // every term d becomes the map k -> k + m + t * [(k + m) odd]: m = rint (d / u), or, when d / u is
        // exactly half way (t = 1), m = floor (d / u) and round-half-even picks the even neighbour.
        // Such a map only looks at the parity of k: it is the pair (ae, ao) of its increments for even
        // and odd k, and pairs compose associatively -> prefix scan.
        double x[4] = {-c1 * scale, c2 * scale, -c3 * scale, c4 * scale};
        long long m[4]; bool t[4];
        PF mine_f{0, 0};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (!valid) x[j] = 0.0;
          const double r = rint(x[j]);
          t[j] = fabs(x[j] - r) == 0.5;
          m[j] = (long long)(t[j] ? floor(x[j]) : r);
          mine_f = pf_compose(mine_f, pf_term(m[j], t[j]));
        }
        PF inc = mine_f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          PF nb;
          nb.ae = __shfl_up_sync(0xffffffffu, inc.ae, o);
          nb.ao = __shfl_up_sync(0xffffffffu, inc.ao, o);
          if (l >= (u32)o) inc = pf_compose(nb, inc);
        }
        if (l == 31) { S.wtot_e[w] = inc.ae; S.wtot_o[w] = inc.ao; }
        PF excl;                                               // lanes before me in my warp
        excl.ae = __shfl_up_sync(0xffffffffu, inc.ae, 1);
        excl.ao = __shfl_up_sync(0xffffffffu, inc.ao, 1);
        if (l == 0) { excl.ae = 0; excl.ao = 0; }
        __syncthreads();
        PF pre{0, 0};                                          // warps before mine
        for (u32 ww = 0; ww < w; ww++) pre = pf_compose(pre, PF{S.wtot_e[ww], S.wtot_o[ww]});
        pre = pf_compose(pre, excl);
        long long k = K0 + ((K0 & 1) ? pre.ao : pre.ae);
        const long long lo = 1ll << 52, hi = 1ll << 53;
        bool mybad = false;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          k = k + m[j] + ((t[j] && ((k + m[j]) & 1)) ? 1 : 0);
          mybad = mybad || k < lo || k >= hi;
        }
        bad = __syncthreads_or(valid && mybad) != 0;
        if (!bad) {
          mine = (double)k * inv;
          if ((int)tid == steps - 1) S.entropy = mine;
        }
      }
      if (bad) {
        S.terms[tid * 4 + 0] = c1; S.terms[tid * 4 + 1] = c2; S.terms[tid * 4 + 2] = c3; S.terms[tid * 4 + 3] = c4;
        __syncthreads();
        if (w == 0) {
          double e = entropy;
          for (int k = 0; k < steps; k++) {
            const double2 ab = *reinterpret_cast<const double2 *>(&S.terms[k * 4]);
            const double2 cd = *reinterpret_cast<const double2 *>(&S.terms[k * 4 + 2]);
            e = __dsub_rn(e, ab.x);                           // data_segmentation.adb:75-90
            e = __dadd_rn(e, ab.y);
            e = __dsub_rn(e, cd.x);
            e = __dadd_rn(e, cd.y);
            if (l == 0) S.series[k] = e;
          }
          if (l == 0) S.entropy = e;
        }
        __syncthreads();
        mine = S.series[tid];
      }
      // threshold tests for the 256 steps at once (:91-97).  A cut moves index_mark to within 256 of
      // every later step of the batch, so at most one cut per profile can happen in a batch.
      const i32 sp = i + 1 - SEG_WINDOW;
      bool maybe[2];
#pragma unroll
      for (int p = 0; p < 2; p++) {
        maybe[p] = act[p] && (i0 + steps - SEG_WINDOW - index_mark[p] > ithr[p]);
        if (maybe[p] && valid && fabs(__dsub_rn(mine, mark[p])) > thr[p] && (sp - index_mark[p] > ithr[p])) atomicMin(&S.first[p], tid);
      }
      // update the window counts and clear my histogram entries for the next batch
      if (valid) { atomicAdd(&S.F[bn], 1u); atomicSub(&S.F[bo], 1u); atomicSub(&S.Hn[w][bn], 1u); atomicSub(&S.Ho[w][bo], 1u); }
      __syncthreads();
#pragma unroll
      for (int p = 0; p < 2; p++) if (maybe[p] && S.first[p] == tid) S.cutval[p] = mine;
      __syncthreads();
      entropy = S.entropy;
#pragma unroll
      for (int p = 0; p < 2; p++) {
        if (maybe[p] && S.first[p] != 0xFFFFFFFFu) {
          const i32 seg_point = i0 + (i32)S.first[p] + 1 - SEG_WINDOW;
          if (tid == 0 && cnt[p] < B2_MAX_SEG - 1) out[p][cnt[p]] = (u32)seg_point;
          cnt[p]++;
          index_mark[p] = seg_point;
          mark[p] = S.cutval[p];
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    for (int k = 0; k < 2; k++) {
      if (len > 0) { if (cnt[k] < B2_MAX_SEG) out[k][cnt[k]] = (u32)len; cnt[k]++; }   // :102-104
      nseg[2 * c + k] = cnt[k];
    }
  }
}

int b2k_segment(cudaStream_t st, const u8 *d_in, const B2Chunk *d_chunks, u32 n_chunks, const double *d_T,
                u32 *d_seg, u32 *d_nseg) {
  if (n_chunks == 0) return 0;
  k_segment<<<n_chunks, SG_THREADS, 0, st>>>(d_in, d_chunks, n_chunks, d_T, d_seg, d_nseg);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}
