// Stage A8-A10: entropy-coder search — initial clustering by ranking, iterative re-clustering,
// length-limited Huffman code lengths, total cost, brute force over (max_code_len, sample_width,
// coders).
//
// Reference: zip_lib/bzip2-encoding.adb:433-978 (Entropy_Calculations / Multiple_Entropy_Coders),
// zip_lib/huffman-encoding-length_limited_coding.adb:46-280 (boundary package-merge and its
// unstable Quick_sort), and — NOT under /root/reference — GNAT's
// Ada.Containers.Generic_Constrained_Array_Sort (heap sort) whose tie order decides the initial
// clustering (:562-568, :619).
//
// `Construct (sample_width)` for a given (max_code_len, coders) is a pure function of the block's
// MTF symbols: it rewrites every selector and every used descriptor.  All 20
// triples of all blocks of a batch therefore advance together, one reclassification iteration
// (:793-802) per round of kernels:
//   k_ent_hist   cluster histograms, updated only for groups that changed cluster (:643-652), Avoid_Zeros
//   k_ent_qsort  the reference's unstable Quick_sort of the leaves, 32 independent sorts per warp in lockstep
//   k_ent_pm     package-merge code lengths, one warp per (block, triple, coder)
//   k_ent_cost   bits of every group under every coder (:731-737)
//   k_ent_sweep  the serial reclassification through the selector MTF list (:672-718): the sweeps of the
//                20 triples of a block run in the 20 lanes of one warp
// A triple whose sweep finds no defector is finished (:801) and skipped by later rounds.  k_choose then
// replays the reference's loop order, its `low_cluster_usage` gate and its strict-< cost selection
// (:930-952) over the stored results; the stored result of the winner is what the reference's final
// `Construct` (:961) would recompute.
#include "b2_common.cuh"
#include "b2_kernels.h"

#define HSTRIDE 260                    // padded alphabet stride of hist / leaves rows
#define QS_SMALL 100                   // alphabets up to this size are sorted by a launch with less shared memory per warp

// ---------------------------------------------------------------------------------------------
// Ranking keys (:593-614): key(group) = number of symbols in run_a .. min(EOB-1, sample_width-1)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_group_keys(const B2Job *__restrict__ jobs, const u16 *__restrict__ mtf, u32 *__restrict__ rank3, u32 *__restrict__ rank4) {
  const B2Job &job = jobs[blockIdx.x];
  const u32 M = job.n_mtf, G = job.n_groups;
  const u16 *m = mtf + job.mtf_off;
  const u32 eob = job.n_used + 1;
  const u32 lim3 = eob < 3 ? eob : 3, lim4 = eob < 4 ? eob : 4;
  for (u32 g = threadIdx.x; g < G; g += blockDim.x) {
    u32 s0 = g * B2_GROUP_SIZE, s1 = min(s0 + B2_GROUP_SIZE, M);
    u32 k3 = 0, k4 = 0;
    for (u32 s = s0; s < s1; s++) { u32 sym = m[s]; k3 += sym < lim3; k4 += sym < lim4; }
    rank3[job.grp_off + g] = (k3 << 16) | (g + 1);
    rank4[job.grp_off + g] = (k4 << 16) | (g + 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Sparse histogram of every group of 50 symbols, built once per block and shared by the 20 triples
// and all their iterations: entries (count << 9 | symbol) in order of first
// appearance, at the group's own 50 slots; gdist = number of entries.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_group_hist(const B2Job *__restrict__ jobs, const u16 *__restrict__ mtf, u16 *__restrict__ ghist, u8 *__restrict__ gdist) {
  // one warp per group: symbols 0..31 of the group in the lanes, then symbols 32..49; equal symbols
  // are found with match_any, the second part is merged into the entries of the first
  const B2Job &job = jobs[blockIdx.x];
  const u32 M = job.n_mtf, G = job.n_groups;
  const u16 *m = mtf + job.mtf_off;
  u16 *gh = ghist + job.mtf_off;
  const u32 l = lane_id(), lt = (1u << l) - 1u;
  for (u32 g = warp_id(); g < G; g += 8) {
    const u32 s0 = g * B2_GROUP_SIZE, s1 = min(s0 + B2_GROUP_SIZE, M);
    const u32 nA = min(32u, s1 - s0), nB = (s1 - s0) - nA;
    // part A
    const u32 symA = l < nA ? m[s0 + l] : (0x10000u + l);
    const u32 peersA = __match_any_sync(0xffffffffu, symA);
    const bool leadA = l < nA && (peersA & lt) == 0;
    const u32 bmA = __ballot_sync(0xffffffffu, leadA);
    const u32 DA = __popc(bmA);
    const u32 slotA = __popc(bmA & lt);                  // entry index of my symbol (leaders)
    u32 cntA = __popc(peersA);
    // part B
    const u32 symB = l < nB ? m[s0 + 32 + l] : (0x20000u + l);
    const u32 peersB = __match_any_sync(0xffffffffu, symB);
    const bool leadB = l < nB && (peersB & lt) == 0;
    const u32 cntB = __popc(peersB);
    // does my B symbol already have an A entry?  lane j (leader of A) broadcasts its symbol
    u32 foundAt = 0xFFFFFFFFu;                           // A leader lane holding my B symbol
    u32 addA = 0;                                        // count to add to my A entry
    u32 todo = bmA;
    while (todo) {
      const u32 j = (u32)(__ffs(todo) - 1);
      todo &= todo - 1;
      const u32 sj = __shfl_sync(0xffffffffu, symA, j);
      const u32 hit = __ballot_sync(0xffffffffu, leadB && symB == sj);
      if (hit) {
        const u32 src = (u32)(__ffs(hit) - 1);
        const u32 cb = __shfl_sync(0xffffffffu, cntB, src);
        if (l == j) addA = cb;
        if (l == src) foundAt = j;
      }
    }
    cntA += addA;
    const bool newB = leadB && foundAt == 0xFFFFFFFFu;
    const u32 bmB = __ballot_sync(0xffffffffu, newB);
    u16 *e = gh + s0;
    if (leadA) e[slotA] = (u16)((cntA << 9) | symA);
    if (newB) e[DA + __popc(bmB & lt)] = (u16)((cntB << 9) | symB);
    if (l == 0) gdist[job.grp_off + g] = (u8)(DA + __popc(bmB));
  }
}

// ---------------------------------------------------------------------------------------------
// Ranking_Sort (:564-568, :619) = GNAT Generic_Constrained_Array_Sort: in-place heap sort,
// Floyd's variant (sift the hole to a leaf, then climb), compare on key only.  Serial by nature
// (its tie order is the point); one thread per array, array staged in shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_rank_sort(const B2Job *__restrict__ jobs, u32 *__restrict__ rank3, u32 *__restrict__ rank4, u32 g_lo, u32 g_hi) {
  extern __shared__ __align__(16) u32 a[];   // 1-based: a[1..G]
  const B2Job &job = jobs[blockIdx.x >> 1];
  u32 *arr = ((blockIdx.x & 1) ? rank4 : rank3) + job.grp_off;
  const i32 G = (i32)job.n_groups;
  if ((u32)G <= g_lo || (u32)G > g_hi) return;     // launched once per size class (shared memory per sort)
  for (i32 i = threadIdx.x; i < G; i += 32) a[i + 1] = arr[i];
  __syncwarp();
  if (threadIdx.x == 0) {
    i32 mx = G;
    u32 temp;
#define KEY(x) ((x) >> 16)
    auto sift = [&](i32 s) {
      i32 c = s;
      // two levels per trip while both are complete: the children and the grandchildren are fetched
      // together (one 8-byte and one 16-byte load), which halves the chain of dependent loads
      // ... and three levels per trip higher up in the heap
      while (8 * c + 7 <= mx) {
        const uint2 pr = *reinterpret_cast<const uint2 *>(&a[2 * c]);
        const uint4 qd = *reinterpret_cast<const uint4 *>(&a[4 * c]);
        const uint4 o0 = *reinterpret_cast<const uint4 *>(&a[8 * c]);
        const uint4 o1 = *reinterpret_cast<const uint4 *>(&a[8 * c + 4]);
        const bool r1 = KEY(pr.x) < KEY(pr.y);
        const i32 s1 = 2 * c + (r1 ? 1 : 0);
        const u32 g0 = r1 ? qd.z : qd.x, g1 = r1 ? qd.w : qd.y;
        const bool r2 = KEY(g0) < KEY(g1);
        const i32 s2 = 2 * s1 + (r2 ? 1 : 0);
        const u32 ox = r1 ? o1.x : o0.x, oy = r1 ? o1.y : o0.y, oz = r1 ? o1.z : o0.z, ow = r1 ? o1.w : o0.w;
        const u32 h0 = r2 ? oz : ox, h1 = r2 ? ow : oy;
        const bool r3 = KEY(h0) < KEY(h1);
        a[c] = r1 ? pr.y : pr.x;
        a[s1] = r2 ? g1 : g0;
        a[s2] = r3 ? h1 : h0;
        c = 2 * s2 + (r3 ? 1 : 0);
      }
      while (4 * c + 3 <= mx) {
        const uint2 pr = *reinterpret_cast<const uint2 *>(&a[2 * c]);
        const uint4 qd = *reinterpret_cast<const uint4 *>(&a[4 * c]);
        const bool r1 = KEY(pr.x) < KEY(pr.y);
        const i32 s1 = 2 * c + (r1 ? 1 : 0);
        const u32 g0 = r1 ? qd.z : qd.x, g1 = r1 ? qd.w : qd.y;
        const bool r2 = KEY(g0) < KEY(g1);
        a[c] = r1 ? pr.y : pr.x;
        a[s1] = r2 ? g1 : g0;
        c = 2 * s1 + (r2 ? 1 : 0);
      }
      for (;;) {
        i32 son = 2 * c;
        if (son > mx) break;
        if (son < mx && KEY(a[son]) < KEY(a[son + 1])) son++;
        a[c] = a[son];
        c = son;
      }
      while (c != s) {
        i32 father = c >> 1;
        if (KEY(a[father]) < KEY(temp)) { a[c] = a[father]; c = father; } else break;
      }
      a[c] = temp;
    };
    for (i32 j = mx / 2; j >= 1; j--) { temp = a[j]; sift(j); }
    while (mx > 1) {
      temp = a[mx];
      a[mx] = a[1];
      mx--;
      sift(1);
    }
#undef KEY
  }
  __syncwarp();
  for (i32 i = threadIdx.x; i < G; i += 32) arr[i] = a[i + 1];
}

// ---------------------------------------------------------------------------------------------
// Per-problem bookkeeping.  Problem p = block * B2_N_TRIPLES + triple.
// ---------------------------------------------------------------------------------------------
struct EntArgs {
  const B2Job *jobs;
  const u16 *mtf;
  const u16 *ghist;                     // sparse group histograms (same layout as mtf)
  const u8 *gdist;                      // [total_groups] entries per group
  const u32 *rank3, *rank4;
  u8 *sel, *selprev;                    // [triple][total_groups]
  u32 *gpack;                           // [triple][total_groups] six 4-bit fields: bits above the cheapest coder, clipped at 7
  u16 *gselcost;                        // [triple][total_groups] exact bits of the group under its current coder
  u32 *hist;                            // [p][6][HSTRIDE] raw cluster histograms
  u32 *leaves;                          // interleaved: [(slot / 32)][HSTRIDE][slot % 32], slot = place in the work list
  u32 *wl, *wl_count;                   // work list of this round: q = p * 6 + coder of every coder to (re)build
  u32 *wl_off;                          // [p] first slot of problem p in the work list (k_ent_wl_scan)
  u32 *chg;                             // [p] bit c: the histogram of coder c changed in this round (bit 31: first round, all coders)
  u32 *pl, *pl_count;                   // unfinished problems of this round, in (block, triple) order: the sweeps run over this list
  u8 *lens;                             // [p][6][B2_MAX_ALPHA]
  u32 *stat;                            // [p][2]: defectors of the last sweep, finished flag
  u32 *selcost;                         // [p]
  u32 *cost_all, *low_all;              // [p]
  u32 total_groups;
  int level, n_triples;
  u32 n_jobs;
};

// Coder counts that the brute force always tries (coder_choices, bzip2-encoding.adb:907-915); the
// others are tried only when the previous Construct left low_cluster_usage set (:936-939).
__host__ __device__ inline u32 b2_choice_mask(int level, u32 M) {
  if (level == 9) {
    if (M <= 5000) return (1u << 2) | (1u << 3) | (1u << 6);
    if (M <= 10000) return (1u << 3) | (1u << 4) | (1u << 6);
    return (1u << 3) | (1u << 4) | (1u << 5) | (1u << 6);
  }
  return (1u << 4) | (1u << 6);
}

__device__ __forceinline__ size_t leaf_index(u32 q, u32 e) { return ((size_t)(q >> 5) * HSTRIDE + e) * 32 + (q & 31); }

// Initial_Clustering_by_Rank (:572-588, :625-631)
__global__ void __launch_bounds__(256)
k_ent_init(EntArgs a) {
  const int t = blockIdx.x;
  const u32 jb = blockIdx.y;
  const B2Job &job = a.jobs[jb];
  const u32 G = job.n_groups;
  int max_len, sw, ec;
  b2_triple(a.level, t, max_len, sw, ec);
  const u32 p = jb * B2_N_TRIPLES + t;
  u8 *sel = a.sel + (size_t)t * a.total_groups + job.grp_off;
  u8 *selprev = a.selprev + (size_t)t * a.total_groups + job.grp_off;
  const u32 *rk = (sw == 3 ? a.rank3 : a.rank4) + job.grp_off;
  const int attr_tab[5][6] = {{2, 1, 0, 0, 0, 0}, {3, 1, 2, 0, 0, 0}, {4, 2, 1, 3, 0, 0}, {5, 3, 1, 2, 4, 0}, {6, 4, 2, 1, 3, 5}};
  for (u32 i = threadIdx.x; i < G; i += 256) {
    // rank position i+1 belongs to range a (1-based) iff high_(a-1) < i+1 <= high_a, high_a = a*G/ec
    const u32 pos1 = i + 1;
    int r = 1;
    while ((u64)r * G / ec < pos1) r++;
    sel[(rk[i] & 0xFFFFu) - 1] = (u8)attr_tab[ec - 2][r - 1];
    selprev[i] = 0;
  }
  u32 *h = a.hist + (size_t)p * (B2_MAX_CODERS * HSTRIDE);
  for (u32 i = threadIdx.x; i < B2_MAX_CODERS * HSTRIDE; i += 256) h[i] = 0;
  u8 *ln = a.lens + (size_t)p * B2_MAX_CODERS * B2_MAX_ALPHA;
  for (u32 i = threadIdx.x; i < B2_MAX_CODERS * B2_MAX_ALPHA; i += 256) ln[i] = 0;
  if (threadIdx.x == 0) {
    const bool always = (b2_choice_mask(a.level, job.n_mtf) >> ec) & 1u;
    a.stat[2 * p] = 1;
    a.stat[2 * p + 1] = always ? 0u : 2u;        // 0 running, 1 finished, 2 not scheduled (gated)
    a.chg[p] = 0x80000000u;
  }
}

// Places of the unfinished problems in the round's work list, in (block, triple) order: neighbours in
// the list then belong to the same block and share its alphabet size, which the 32 lock-step sorts of a
// warp need (k_ent_qsort), and finished problems cost nothing.
#define WL_TILE 2048
// coders problem p adds to the work list (low 16 bits) and 1 << 16 if the problem is unfinished: both lists
// come out of the same scan (a tile of 2048 problems adds at most 12 288 coders)
__device__ __forceinline__ u32 wl_items(const EntArgs &a, u32 p) {
  const int t = (int)(p % B2_N_TRIPLES);
  if (p >= a.n_jobs * B2_N_TRIPLES || t >= a.n_triples || a.stat[2 * p + 1] != 0) return 0;
  return (u32)__popc(a.chg[p] & 63u) | 0x10000u;
}
__global__ void __launch_bounds__(1024)
k_ent_wl_sum(EntArgs a, u32 *tile_sum) {
  __shared__ u32 sm[40];
  const u32 p = blockIdx.x * WL_TILE + 2 * threadIdx.x;
  u32 total;
  block_excl_add(wl_items(a, p) + wl_items(a, p + 1), sm, &total);
  if (threadIdx.x == 0) { tile_sum[2 * blockIdx.x] = total & 0xFFFFu; tile_sum[2 * blockIdx.x + 1] = total >> 16; }
}
__global__ void __launch_bounds__(1024)
k_ent_wl_scan(EntArgs a, const u32 *tile_sum) {
  __shared__ u32 sm[40];
  u32 before_c = 0, before_p = 0;
  for (u32 b = threadIdx.x; b < blockIdx.x; b += 1024) { before_c += tile_sum[2 * b]; before_p += tile_sum[2 * b + 1]; }
  u32 base_c, base_p;
  block_excl_add(before_c, sm, &base_c);
  __syncthreads();
  block_excl_add(before_p, sm, &base_p);
  __syncthreads();
  const u32 P = a.n_jobs * B2_N_TRIPLES;
  const u32 p = blockIdx.x * WL_TILE + 2 * threadIdx.x;
  const u32 c0 = wl_items(a, p), c1 = wl_items(a, p + 1);
  u32 total;
  const u32 ex = block_excl_add(c0 + c1, sm, &total);
  if (p < P) { a.wl_off[p] = base_c + (ex & 0xFFFFu); if (c0 >> 16) a.pl[base_p + (ex >> 16)] = p; }
  if (p + 1 < P) { a.wl_off[p + 1] = base_c + ((ex + c0) & 0xFFFFu); if (c1 >> 16) a.pl[base_p + ((ex + c0) >> 16)] = p + 1; }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { *a.wl_count = base_c + (total & 0xFFFFu); *a.pl_count = base_p + (total >> 16); }
}

// Define_Descriptors, first half (:643-652) + Avoid_Zeros (:439-462)
__global__ void __launch_bounds__(256)
k_ent_hist(EntArgs a) {
  __shared__ u32 h[B2_MAX_CODERS][HSTRIDE];
  __shared__ u32 changed;
  const int t = blockIdx.x;
  const u32 jb = blockIdx.y;
  const u32 p = jb * B2_N_TRIPLES + t;
  if (a.stat[2 * p + 1]) return;
  const B2Job &job = a.jobs[jb];
  const u32 G = job.n_groups;
  int max_len, sw, ec;
  b2_triple(a.level, t, max_len, sw, ec);
  const u8 *sel = a.sel + (size_t)t * a.total_groups + job.grp_off;
  u8 *selprev = a.selprev + (size_t)t * a.total_groups + job.grp_off;
  u32 *hg = a.hist + (size_t)p * (B2_MAX_CODERS * HSTRIDE);
  const u32 tid = threadIdx.x;
  for (u32 i = tid; i < (u32)ec * HSTRIDE; i += 256) (&h[0][0])[i] = hg[i];
  if (tid == 0) changed = 0;
  __syncthreads();
  const u16 *gh = a.ghist + job.mtf_off;
  const u8 *gd = a.gdist + job.grp_off;
  u32 mych = 0;
  for (u32 g = tid; g < G; g += 256) {
    const u32 c = sel[g], o = selprev[g];
    if (c == o) continue;
    const u16 *e = gh + g * B2_GROUP_SIZE;
    const u32 D = gd[g];
    for (u32 k = 0; k < D; k++) {
      const u32 v = e[k], sym = v & 511u, cn = v >> 9;
      atomicAdd(&h[c - 1][sym], cn);
      if (o) atomicSub(&h[o - 1][sym], cn);
    }
    mych |= (1u << (c - 1)) | (o ? (1u << (o - 1)) : 0u);
    selprev[g] = (u8)c;
  }
  mych = __reduce_or_sync(0xffffffffu, mych);
  if (lane_id() == 0 && mych) atomicOr(&changed, mych);
  __syncthreads();
  for (u32 i = tid; i < (u32)ec * HSTRIDE; i += 256) hg[i] = (&h[0][0])[i];
  // Coders whose histogram did not change keep their code lengths (Define_Descriptors is a function of the
  // histogram, :495-513): only the others join the round's work list.  The first round builds all of them.
  if (tid == 0) a.chg[p] = (a.chg[p] & 0x80000000u) ? ((1u << ec) - 1u) : changed;
}

// Avoid_Zeros (:439-462) and the leaves, in alphabet order (:230-235), of every coder on the work list
__global__ void __launch_bounds__(256)
k_ent_leaves(EntArgs a) {
  const int t = blockIdx.x;
  const u32 jb = blockIdx.y;
  const u32 p = jb * B2_N_TRIPLES + t;
  if (a.stat[2 * p + 1]) return;
  const u32 mask = a.chg[p] & 63u;
  const u32 w = warp_id(), l = lane_id();
  if (!((mask >> w) & 1u)) return;
  const int A = (int)a.jobs[jb].n_used + 2;
  const u32 *h = a.hist + (size_t)p * (B2_MAX_CODERS * HSTRIDE) + (size_t)w * HSTRIDE;
  const u32 slot = a.wl_off[p] + __popc(mask & ((1u << w) - 1u));
  u32 zeroes = 0;
  for (int s = l; s < A; s += 32) zeroes += (h[s] == 0);
#pragma unroll
  for (int o = 16; o; o >>= 1) zeroes += __shfl_xor_sync(0xffffffffu, zeroes, o);
  if (l == 0) a.wl[slot] = p * B2_MAX_CODERS + w;
  for (int s = l; s < A; s += 32) {
    u32 v = h[s];
    if (zeroes > 0 && zeroes <= 100) v = max(1u, v);
    else if (zeroes > 100) v = (v == 0) ? 1u : v * 2u;
    a.leaves[leaf_index(slot, (u32)s)] = (v << 9) | (u32)s;     // every count is > 0 here
  }
}

// ---------------------------------------------------------------------------------------------
// The reference's Quick_sort (huffman-encoding-length_limited_coding.adb:191-223), whose tie order
// decides which of several equal-weight symbols get the longer codes.  32 independent sorts per warp,
// one per lane, advanced in lockstep as a small state machine; arrays interleaved in shared memory
// (element e of lane l at e*32+l: conflict free).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_ent_qsort(EntArgs a, u32 nq, u32 alpha_max, u32 n_lo, u32 n_hi) {
  extern __shared__ u32 qs_smem[];             // alpha_max x 32 elements, then 16 x 32 stack words
  u32 *s = qs_smem;
  u32 *stk = qs_smem + alpha_max * 32;         // per-lane stack of (first << 16 | count)
  const u32 l = threadIdx.x;
  const u32 slot = blockIdx.x * 32 + l;
  const u32 count = min(*a.wl_count, nq);
  if (blockIdx.x * 32 >= count) return;
  u32 n = 0;
  if (slot < count) {
    const u32 q = a.wl[slot];
    const u32 jb = (q / B2_MAX_CODERS) / B2_N_TRIPLES;
    n = a.jobs[jb].n_used + 2;
  }
  const u32 nmax = __reduce_max_sync(0xffffffffu, n);
  if (nmax <= n_lo || nmax > n_hi) return;     // launched once per alphabet class (shared memory per warp)
  u32 *g = a.leaves + (size_t)blockIdx.x * HSTRIDE * 32;
  for (u32 e = 0; e < nmax; e++) s[e * 32 + l] = g[e * 32 + l];
  __syncwarp();
#define EL(x) s[(x) * 32 + l]
  // Every lane walks scan-up -> scan-down -> check in ONE pass of the loop body, so that a lane whose
  // scans stop at once (the common case with many equal weights) completes a swap per pass.
  enum { SI = 1, SJ = 2, CK = 3 };
  int sp = 0, st = SI;
  u32 f = 0, nn = n, pw = 0;
  i32 i = 0, j = (i32)n - 1;
  bool active = n >= 2;
  if (active) pw = EL(nn / 2) >> 9;
  while (__any_sync(0xffffffffu, active)) {
    if (active) {
      if (st == SI) {
#pragma unroll
        for (int u = 0; u < 4; u++) { if (st == SI) { if ((EL(f + i) >> 9) < pw) i++; else st = SJ; } }
      }
      if (st == SJ) {
#pragma unroll
        for (int u = 0; u < 4; u++) { if (st == SJ) { if (pw < (EL(f + j) >> 9)) j--; else st = CK; } }
      }
      if (st == CK) {
        if (i >= j) {
          // Quick_sort (a (first .. first+i-1)); Quick_sort (a (first+i .. last)): the smaller range is
          // sorted next, the larger one waits on the stack (the ranges are disjoint, order is free)
          const u32 n1 = (u32)i, n2 = nn - (u32)i;
          u32 fa = f, na = n1, fb = f + (u32)i, nb = n2;       // a = next, b = later
          if (n1 > n2) { fa = f + (u32)i; na = n2; fb = f; nb = n1; }
          if (na < 2) { fa = fb; na = nb; nb = 0; }
          if (nb >= 2) stk[(sp++) * 32 + l] = (fb << 16) | nb;
          if (na < 2) {
            if (sp == 0) active = false;
            else { const u32 v = stk[(--sp) * 32 + l]; fa = v >> 16; na = v & 0xFFFFu; }
          }
          if (active) {
            f = fa; nn = na;
            pw = EL(f + nn / 2) >> 9;
            i = 0; j = (i32)nn - 1;
          }
        } else {
          const u32 x = EL(f + i), y = EL(f + j);
          EL(f + i) = y; EL(f + j) = x;
          i++; j--;
        }
        st = SI;
      }
    }
  }
#undef EL
  __syncwarp();
  for (u32 e = 0; e < nmax; e++) g[e * 32 + l] = s[e * 32 + l];
}

#include "b2_pm.cuh"    // length-limited code lengths: the package-merge of one warp

__global__ void __launch_bounds__(32 * PM_WARPS)
k_ent_pm(EntArgs a, u32 nq, u32 alpha_max, int variant) {
  extern __shared__ u32 pm_smem[];
  const u32 w = warp_id(), l = lane_id();
  const u32 slot = blockIdx.x * PM_WARPS + w;
  if (slot >= min(*a.wl_count, nq)) return;
  LLScratch S = ll_scratch_at(pm_smem + (size_t)w * ll_scratch_words(alpha_max), alpha_max);
  const u32 q = a.wl[slot];
  const u32 p = q / B2_MAX_CODERS, c = q % B2_MAX_CODERS;
  const u32 jb = p / B2_N_TRIPLES, t = p % B2_N_TRIPLES;
  int max_len, sw, ec;
  b2_triple(a.level, (int)t, max_len, sw, ec);
  const int ns = (int)a.jobs[jb].n_used + 2;
  for (int e = l; e < ns; e += 32) S.leaf[e] = a.leaves[leaf_index(slot, (u32)e)];
  __syncwarp();
  u8 *lens = a.lens + ((size_t)p * B2_MAX_CODERS + c) * B2_MAX_ALPHA;
  if (variant == 0) ll_package_merge_warp(S, ns, max_len, lens);
  else ll_package_merge_warp_mp(S, ns, max_len, lens);
}

// ---------------------------------------------------------------------------------------------
// The selector MTF list (:669-717, :816-835) is kept as places: a 4-bit field per coder holding its
// 1-based place.  Moving the coder at place p to the front increments every place < p and sets its
// own to 1; with 4-bit fields "place < p" for all coders at once is a carry-free add and mask.
// ---------------------------------------------------------------------------------------------
// bits of every group under every coder (:731-737), six 10-bit fields per group
__global__ void __launch_bounds__(256)
k_ent_cost(EntArgs a) {
  __shared__ unsigned long long lenpack[HSTRIDE];
  const int t = blockIdx.x;
  const u32 jb = blockIdx.y;
  const u32 p = jb * B2_N_TRIPLES + t;
  if (a.stat[2 * p + 1]) return;
  const B2Job &job = a.jobs[jb];
  const u32 M = job.n_mtf, G = job.n_groups;
  const int A = (int)job.n_used + 2;
  const u16 *m = a.mtf + job.mtf_off;
  int max_len, sw, ec;
  b2_triple(a.level, t, max_len, sw, ec);
  const u8 *lens = a.lens + (size_t)p * B2_MAX_CODERS * B2_MAX_ALPHA;
  for (int s = threadIdx.x; s < A; s += 256) {
    unsigned long long v = 0;
    for (int c = 0; c < ec; c++) v |= (unsigned long long)lens[c * B2_MAX_ALPHA + s] << (10 * c);
    lenpack[s] = v;
  }
  __syncthreads();
  u32 *gpack = a.gpack + (size_t)t * a.total_groups + job.grp_off;
  u16 *gselcost = a.gselcost + (size_t)t * a.total_groups + job.grp_off;
  const u8 *sel = a.sel + (size_t)t * a.total_groups + job.grp_off;
  const u16 *gh = a.ghist + job.mtf_off;
  const u8 *gd = a.gdist + job.grp_off;
  // The sparse histograms of 32 consecutive groups are one contiguous piece of 3200 bytes: every warp
  // brings its piece to shared memory with coalesced loads, then each lane walks its own row there
  // (rows are 25 words apart: no bank conflicts).
  __shared__ u32 rows[8][32 * B2_GROUP_SIZE / 2];
  const u32 wid = threadIdx.x >> 5, ln = threadIdx.x & 31u;
  const u32 *gh32 = reinterpret_cast<const u32 *>(gh);       // the arena offset of a block is a multiple of 4 symbols
  for (u32 gb = 0; gb < G; gb += 256) {
    const u32 gw = gb + 32 * wid;                             // first group of my warp
    const u32 nrow = gw < G ? min(32u, G - gw) : 0u;
    __syncwarp();
    for (u32 i = ln; i < nrow * (B2_GROUP_SIZE / 2); i += 32) rows[wid][i] = gh32[(size_t)gw * (B2_GROUP_SIZE / 2) + i];
    __syncwarp();
    const u32 g = gw + ln;
    if (g >= G) continue;
    const u32 *e2 = rows[wid] + ln * (B2_GROUP_SIZE / 2);            // two entries per word
    const u32 D = gd[g];
    unsigned long long acc = 0;
    for (u32 k = 0; k + 1 < D; k += 2) {
      const u32 v = e2[k >> 1];
      acc += lenpack[v & 511u] * (unsigned long long)((v >> 9) & 127u);
      acc += lenpack[(v >> 16) & 511u] * (unsigned long long)(v >> 25);
    }
    if (D & 1u) { const u32 v = e2[D >> 1] & 0xFFFFu; acc += lenpack[v & 511u] * (unsigned long long)(v >> 9); }
    // Only cost differences matter to the reclassification, and a coder 7 or more bits above the
    // cheapest can never win (places cost 1..6, :683-695): keep six clipped excesses in 4-bit fields
    // (coders beyond ec get 7 and, in the sweep, place 7).
    u32 c[6], mn = 0xFFFFFFFFu;
#pragma unroll
    for (int cl = 0; cl < 6; cl++) { c[cl] = (u32)((acc >> (10 * cl)) & 1023u); if (cl < ec) mn = min(mn, c[cl]); }
    u32 pk = 0;
#pragma unroll
    for (int cl = 0; cl < 6; cl++) pk |= (cl < ec ? min(c[cl] - mn, 7u) : 7u) << (4 * cl);
    gpack[g] = pk;
    const u32 sc = sel[g] - 1;
    u32 mine = c[0];
#pragma unroll
    for (int cl = 1; cl < 6; cl++) mine = (sc == (u32)cl) ? c[cl] : mine;
    gselcost[g] = (u16)mine;
  }
}

// Simulate_Entropy_Coding_Variants_and_Reclassify (:661-753): serial through the selector MTF list.
// One lane per unfinished (block, triple): 32 neighbours of the round's problem list per warp (they mostly
// belong to the same block, hence the same number of groups).
__global__ void __launch_bounds__(32)
k_ent_sweep(EntArgs a) {
  const u32 slot = blockIdx.x * 32 + threadIdx.x;
  const u32 count = *a.pl_count;
  if (blockIdx.x * 32 >= count) return;
  const bool active = slot < count;
  const u32 p = a.pl[active ? slot : blockIdx.x * 32];
  const u32 jb = p / B2_N_TRIPLES, t = p % B2_N_TRIPLES;
  const B2Job &job = a.jobs[jb];
  const u32 G = active ? job.n_groups : 0u;
  const u32 Gmax = __reduce_max_sync(0xffffffffu, G);
  int max_len, sw, ec;
  b2_triple(a.level, (int)t, max_len, sw, ec);
  const size_t base = (size_t)t * a.total_groups + job.grp_off;
  const u32 *gc = a.gpack + base;
  u8 *sel = a.sel + base;
  u32 posv = 0;                       // place of coder cl in field cl; coders beyond ec sit at place 7
#pragma unroll
  for (int cl = 0; cl < 6; cl++) posv |= (cl < ec ? (u32)cl + 1u : 7u) << (4 * cl);
  u32 def = 0;
  // 16 groups per step, loaded one step ahead (register double buffer) and prefetched into L2 further ahead
  u32 nx[16];
  u32 ns[4] = {0, 0, 0, 0};
  auto load16g = [&](u32 g0) {
#pragma unroll
    for (int k = 0; k < 16; k++) nx[k] = 0;
    ns[0] = ns[1] = ns[2] = ns[3] = 0;
    if (g0 < G) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint4 x = *reinterpret_cast<const uint4 *>(gc + g0 + 4 * k);
        nx[4 * k] = x.x; nx[4 * k + 1] = x.y; nx[4 * k + 2] = x.z; nx[4 * k + 3] = x.w;
      }
      const uint4 sv = *reinterpret_cast<const uint4 *>(sel + g0);
      ns[0] = sv.x; ns[1] = sv.y; ns[2] = sv.z; ns[3] = sv.w;
      if (g0 + 256 < G) asm volatile("prefetch.global.L2 [%0];" ::"l"(gc + g0 + 256));
    }
  };
  load16g(0);
  for (u32 g0 = 0; g0 < Gmax; g0 += 16) {
    u32 c16[16];
#pragma unroll
    for (int k = 0; k < 16; k++) c16[k] = nx[k];
    const u32 s4[4] = {ns[0], ns[1], ns[2], ns[3]};
    load16g(g0 + 16);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (g0 + k < G) {
        const u32 clk = (s4[k >> 2] >> (8 * (k & 3))) & 255u;
        // cost of coder cl = excess + place, both in 4-bit fields (no carries: <= 7 + 7).  key = (cost << 3)
        // | coder0: the minimum is the cheapest coder, lowest coder on ties (strict "<" scanning cl
        // upward, :691-695).
        // The six keys are built three at a time as bytes (even coders, odd coders) and reduced with one
        // per-byte minimum and two scalar ones.
        const u32 sum = c16[k] + posv;
        const u32 ev = ((sum & 0x000F0F0Fu) << 3) | 0x00040200u;            // coders 0, 2, 4
        const u32 od = (((sum >> 4) & 0x000F0F0Fu) << 3) | 0x00050301u;     // coders 1, 3, 5
        const u32 m3 = __vminu4(ev, od);
        const u32 key = min(min(m3 & 255u, (m3 >> 8) & 255u), (m3 >> 16) & 255u);
        const u32 best0 = key & 7u;
        if (best0 + 1 != clk) { def++; sel[g0 + k] = (u8)(best0 + 1); }
        // move to front (:707-717): places below the chosen one move down by one, the chosen one becomes 1
        const u32 pl = (posv >> (4 * best0)) & 15u;
        posv += (~(posv + (8u - pl) * 0x111111u) & 0x888888u) >> 3;
        posv = (posv & ~(15u << (4 * best0))) | (1u << (4 * best0));
      }
    }
  }
  if (active) { a.stat[2 * p] = def; if (def == 0) a.stat[2 * p + 1] = 1; }
}

// Compute_Selectors_Cost (:815-837), lanes = triples
__global__ void __launch_bounds__(32)
k_ent_selcost(EntArgs a) {
  const u32 jb = blockIdx.x;
  const u32 t = threadIdx.x;
  const B2Job &job = a.jobs[jb];
  const u32 G = job.n_groups;
  const bool active = (int)t < a.n_triples;
  const u8 *sel = a.sel + (size_t)(active ? t : 0) * a.total_groups + job.grp_off;
  u32 posv = 0x654321u;               // initial list 1, 2, 3, ... (:820-822)
  u32 bits = 0;
  uint4 nx = *reinterpret_cast<const uint4 *>(sel);
  for (u32 g0 = 0; g0 < G; g0 += 16) {
    const uint4 cur = nx;
    if (g0 + 16 < G) nx = *reinterpret_cast<const uint4 *>(sel + g0 + 16);
    const u32 w4[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (g0 + k < G) {
        const u32 cl0 = ((w4[k >> 2] >> (8 * (k & 3))) & 255u) - 1;
        const u32 pl = (posv >> (4 * cl0)) & 15u;
        bits += pl;
        posv += (~(posv + (8u - pl) * 0x111111u) & 0x888888u) >> 3;
        posv = (posv & ~(15u << (4 * cl0))) | (1u << (4 * cl0));
      }
    }
  }
  if (active) a.selcost[jb * B2_N_TRIPLES + t] = bits;
}

// Cluster_Statistics (:757-778) + Compute_Total_Entropy_Cost (:811-887)
__global__ void __launch_bounds__(256)
k_ent_final(EntArgs a) {
  __shared__ u32 stat[8];
  __shared__ u32 red[8];
  const int t = blockIdx.x;
  const u32 jb = blockIdx.y;
  const u32 p = jb * B2_N_TRIPLES + t;
  if (a.stat[2 * p + 1] == 2u) return;             // not scheduled (yet)
  const B2Job &job = a.jobs[jb];
  const u32 G = job.n_groups;
  const int A = (int)job.n_used + 2;
  int max_len, sw, ec;
  b2_triple(a.level, t, max_len, sw, ec);
  const u8 *lens = a.lens + (size_t)p * B2_MAX_CODERS * B2_MAX_ALPHA;
  const u8 *sel = a.sel + (size_t)t * a.total_groups + job.grp_off;
  const u16 *gselcost = a.gselcost + (size_t)t * a.total_groups + job.grp_off;
  const u32 tid = threadIdx.x;
  if (tid < 8) stat[tid] = 0;
  __syncthreads();
  u32 part = 0;
  for (u32 g = tid; g < G; g += 256) {
    const u32 c = sel[g];
    atomicAdd(&stat[c], 1u);
    part += gselcost[g];                                            // data bits (:873-881)
  }
  for (int i = tid; i < ec * A; i += 256) {                         // code length tables (:839-865)
    const int c = i / A, s = i % A;
    const int cur = s == 0 ? lens[c * B2_MAX_ALPHA] : lens[c * B2_MAX_ALPHA + s - 1];
    const int nw = lens[c * B2_MAX_ALPHA + s];
    const int dlt = nw > cur ? nw - cur : cur - nw;
    part += 2 * dlt + 1 + (s == 0 ? 5 : 0);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane_id() == 0) red[warp_id()] = part;
  __syncthreads();
  if (tid == 0) {
    u32 total = a.selcost[p];
    for (int w = 0; w < 8; w++) total += red[w];
    const u32 uniform_usage = G / (u32)ec;
    u32 low = 0;
    for (int c = 1; c <= ec; c++) if (stat[c] < uniform_usage / 2) low = 1;
    a.cost_all[p] = total;
    a.low_all[p] = low;
    a.stat[2 * p + 1] = 1;                         // finished (converged or all rounds done)
  }
}

// ---------------------------------------------------------------------------------------------
// Schedules the gated triples whose turn has come: a coder count outside coder_choices is tried iff
// the Construct that ran just before it in the reference's loop (ec from 6 down to 2 inside every
// (max_code_len, sample_width) pair, :930-952) left low_cluster_usage set.
// ---------------------------------------------------------------------------------------------
__global__ void k_ent_gate(EntArgs a, u32 *activated) {
  const u32 jb = blockIdx.x * blockDim.x + threadIdx.x;
  if (jb >= a.n_jobs) return;
  const u32 mask = b2_choice_mask(a.level, a.jobs[jb].n_mtf);
  u32 count = 0;
  for (int t0 = 0; t0 < a.n_triples; t0 += 5) {
    bool low = false;            // low_cluster_usage as the loop reaches each coder count
    bool known = true;           // false once a Construct that would run has not finished yet
    for (int c = 0; c < 5 && known; c++) {
      const int ec = 6 - c;
      const u32 p = jb * B2_N_TRIPLES + t0 + c;
      const bool runs = low || ((mask >> ec) & 1u);
      if (!runs) continue;
      const u32 f = a.stat[2 * p + 1];
      if (f == 2u) { a.stat[2 * p + 1] = 0; count++; known = false; }     // schedule it; what follows depends on it
      else low = a.low_all[p] != 0;
    }
  }
  if (count) atomicAdd(activated, count);
}

// ---------------------------------------------------------------------------------------------
// Replay of the brute-force loop (:930-952) over the stored per-triple results.
// ---------------------------------------------------------------------------------------------
__global__ void k_choose(B2Job *jobs, u32 n_jobs, const u32 *__restrict__ cost_all, const u32 *__restrict__ low_all,
                         int level, int n_triples) {
  u32 jb = blockIdx.x * blockDim.x + threadIdx.x;
  if (jb >= n_jobs) return;
  B2Job &job = jobs[jb];
  const u32 M = job.n_mtf;
  const u32 choice_mask = b2_choice_mask(level, M);   // bit ec set if ec in coder_choices (:907-915)
  bool low = false;
  u32 best_cost = 0x7FFFFFFFu, best = 0;
  for (int t = 0; t < n_triples; t++) {
    int max_len, sw, ec;
    b2_triple(level, t, max_len, sw, ec);
    if (low || ((choice_mask >> ec) & 1u)) {
      u32 cost = cost_all[(size_t)jb * B2_N_TRIPLES + t];
      low = low_all[(size_t)jb * B2_N_TRIPLES + t] != 0;
      if (cost < best_cost) { best_cost = cost; best = (u32)t; }
    }
  }
  job.best = best;
  job.best_cost = best_cost;
  // block bits: header 48+32+1+24, map 16 + 16 per used 16-range, 3 + 15, then best_cost
  u32 ranges = 0;
  for (int i = 0; i < 16; i++) {
    u32 wv = job.in_use[i >> 1];
    u32 h = (i & 1) ? (wv >> 16) : (wv & 0xFFFFu);
    if (h) ranges++;
  }
  job.nbits = 105ull + 16ull + 16ull * ranges + 3ull + 15ull + (u64)best_cost;
}

int b2k_entropy(cudaStream_t st, B2Job *d_jobs, u32 n_jobs, u32 max_groups_per_job, u32 total_groups,
                const u16 *d_mtf, u16 *d_ghist, u8 *d_gdist, u32 *d_rank3, u32 *d_rank4, u8 *d_sel, u8 *d_selprev,
                u32 *d_gpack, u16 *d_gselcost,
                u32 *d_hist, u32 *d_leaves, u32 *d_wl, u8 *d_lens, u32 *d_stat, u32 *d_selcost, u32 *d_cost, u32 *d_low,
                int level, u32 max_alpha, u32 *d_activated, u64 *launches) {
  B2_CUDA_CHECK(cudaFuncSetAttribute(k_rank_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 18064 * 4));
  B2_CUDA_CHECK(cudaFuncSetAttribute(k_ent_qsort, cudaFuncAttributeMaxDynamicSharedMemorySize, (HSTRIDE + 16) * 32 * 4));
  if (n_jobs == 0) return 0;
  if (max_alpha < 2) max_alpha = 2;
  if (max_alpha > HSTRIDE) max_alpha = HSTRIDE;
  const size_t qs_smem = ((size_t)max_alpha + 16) * 32 * 4;
  const size_t pm_smem = (size_t)PM_WARPS * ll_scratch_words(max_alpha) * 4;
  int pm_variant = 1;                    // B2GPU_PM = 0: lists merged by binary searches (round 1); 1: by a merge path
  if (const char *e = getenv("B2GPU_PM")) pm_variant = atoi(e) ? 1 : 0;
  const int n_triples = level == 9 ? 20 : 5;
  k_group_keys<<<n_jobs, 256, 0, st>>>(d_jobs, d_mtf, d_rank3, d_rank4);
  k_group_hist<<<n_jobs, 256, 0, st>>>(d_jobs, d_mtf, d_ghist, d_gdist);
  // small blocks (the four parts, segments, archive entries) do not pay for the shared memory of a full block
  {
    u32 lo = 0;
    for (u32 hi = 2304; lo < max_groups_per_job; hi *= 2) {
      const u32 top = hi >= max_groups_per_job ? max_groups_per_job : hi;
      k_rank_sort<<<n_jobs * 2, 32, ((size_t)top + 2) * 4, st>>>(d_jobs, d_rank3, d_rank4, lo, top);
      *launches += 1;
      lo = top;
    }
  }
  EntArgs a;
  a.jobs = d_jobs; a.mtf = d_mtf; a.ghist = d_ghist; a.gdist = d_gdist; a.rank3 = d_rank3; a.rank4 = d_rank4; a.sel = d_sel; a.selprev = d_selprev;
  a.gpack = d_gpack; a.gselcost = d_gselcost; a.hist = d_hist; a.leaves = d_leaves; a.wl = d_wl; a.wl_count = d_activated + 1; a.pl_count = d_activated + 2; a.wl_off = d_wl + (size_t)n_jobs * B2_N_TRIPLES * B2_MAX_CODERS + 32;
  a.chg = a.wl_off + (size_t)n_jobs * B2_N_TRIPLES + 32; a.pl = a.chg + (size_t)n_jobs * B2_N_TRIPLES + 32; a.lens = d_lens; a.stat = d_stat; a.selcost = d_selcost;
  a.cost_all = d_cost; a.low_all = d_low; a.total_groups = total_groups; a.level = level; a.n_triples = n_triples;
  a.n_jobs = n_jobs;
  const dim3 grid(n_triples, n_jobs);
  const u32 nq = n_jobs * B2_N_TRIPLES * B2_MAX_CODERS;
  const u32 wl_tiles = (n_jobs * B2_N_TRIPLES + WL_TILE - 1) / WL_TILE;
  u32 *wl_tile_sum = a.pl + (size_t)n_jobs * B2_N_TRIPLES + 32;
  const u32 np = n_jobs * B2_N_TRIPLES;
  k_ent_init<<<grid, 256, 0, st>>>(a);
  *launches += 4;
  for (int phase = 0; phase < 4; phase++) {
    for (int it = 0; it <= 10; it++) {
      // iterations 1..10 (:793-802); round 10 is the extra Define_Descriptors for triples still moving (:803-807)
      k_ent_hist<<<grid, 256, 0, st>>>(a);
      k_ent_wl_sum<<<wl_tiles, 1024, 0, st>>>(a, wl_tile_sum);
      k_ent_wl_scan<<<wl_tiles, 1024, 0, st>>>(a, wl_tile_sum);
      k_ent_leaves<<<grid, 256, 0, st>>>(a);
      if (max_alpha > QS_SMALL) {             // small alphabets keep their high occupancy next to large ones
        k_ent_qsort<<<(nq + 31) / 32, 32, ((size_t)QS_SMALL + 16) * 32 * 4, st>>>(a, nq, QS_SMALL, 0, QS_SMALL);
        k_ent_qsort<<<(nq + 31) / 32, 32, qs_smem, st>>>(a, nq, max_alpha, QS_SMALL, HSTRIDE);
        *launches += 1;
      } else {
        k_ent_qsort<<<(nq + 31) / 32, 32, qs_smem, st>>>(a, nq, max_alpha, 0, HSTRIDE);
      }
      k_ent_pm<<<(nq + PM_WARPS - 1) / PM_WARPS, 32 * PM_WARPS, pm_smem, st>>>(a, nq, max_alpha, pm_variant);
      k_ent_cost<<<grid, 256, 0, st>>>(a);
      *launches += 7;
      if (it < 10) { k_ent_sweep<<<(np + 31) / 32, 32, 0, st>>>(a); *launches += 1; }
    }
    k_ent_selcost<<<n_jobs, 32, 0, st>>>(a);
    k_ent_final<<<grid, 256, 0, st>>>(a);
    // gated coder counts whose predecessor asked for them (at most three more phases)
    B2_CUDA_CHECK(cudaMemsetAsync(d_activated, 0, sizeof(u32), st));
    k_ent_gate<<<(n_jobs + 127) / 128, 128, 0, st>>>(a, d_activated);
    *launches += 3;
    u32 activated = 0;
    B2_CUDA_CHECK(cudaMemcpyAsync(&activated, d_activated, sizeof(u32), cudaMemcpyDeviceToHost, st));
    B2_CUDA_CHECK(cudaStreamSynchronize(st));
    if (!activated) break;
  }
  k_choose<<<(n_jobs + 127) / 128, 128, 0, st>>>(d_jobs, n_jobs, d_cost, d_low, level, n_triples);
  *launches += 3;
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}
