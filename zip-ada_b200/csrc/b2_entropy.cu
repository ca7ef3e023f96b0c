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
// triples of a block therefore run concurrently, one CTA each (k_construct); k_choose then replays
// the reference's loop order, its `low_cluster_usage` gate and its strict-< cost selection
// (:930-952) over the stored results, and the stored result of the winner is what the reference's
// final `Construct` (:961) would recompute.
#include "b2_common.cuh"
#include "b2_kernels.h"

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
// Ranking_Sort (:564-568, :619) = GNAT Generic_Constrained_Array_Sort: in-place heap sort,
// Floyd's variant (sift the hole to a leaf, then climb), compare on key only.  Serial by nature
// (its tie order is the point); one thread per array, array staged in shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_rank_sort(const B2Job *__restrict__ jobs, u32 *__restrict__ rank3, u32 *__restrict__ rank4) {
  extern __shared__ u32 a[];   // 1-based: a[1..G]
  const B2Job &job = jobs[blockIdx.x >> 1];
  u32 *arr = ((blockIdx.x & 1) ? rank4 : rank3) + job.grp_off;
  const i32 G = (i32)job.n_groups;
  for (i32 i = threadIdx.x; i < G; i += 32) a[i + 1] = arr[i];
  __syncwarp();
  if (threadIdx.x == 0) {
    i32 mx = G;
    u32 temp;
#define KEY(x) ((x) >> 16)
    auto sift = [&](i32 s) {
      i32 c = s;
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
// Length-limited code lengths (huffman-encoding-length_limited_coding.adb:46-280).
//
// The reference runs the *boundary* package-merge (lazy, recursive, :131-163).  What it computes is
// the classic package-merge: list 1 = the sorted leaves; list l+1 = merge (leaves, pair sums of
// consecutive items of list l); take the first 2n-2 items of the last list, then walk down: if p of
// the first k items of a list are packages, the first 2p items of the list below are taken; a leaf
// gets one bit per list in which it is taken (Extract_Bit_Lengths, :180-189).  The boundary
// version's tie rule "new leaf iff sum > leaf weight" (:151) means a package goes BEFORE a leaf of
// equal weight.  Each list is built here by one warp as a parallel merge (binary searches); the
// equality of both formulations was checked on the CPU against the oracle's literal restatement
// (tests/test_oracle.py::test_forward_package_merge_equals_boundary).  The leaf order comes from
// the reference's own unstable Quick_sort (:191-223), replayed literally by lane 0 because its tie
// order decides which of several equal-weight symbols get the longer codes.
// ---------------------------------------------------------------------------------------------
#define LL_MAXBITS 17
#define LL_MAXITEMS (2 * B2_MAX_ALPHA)
#define LL_BITWORDS 17

struct LLScratch {
  u32 leaf[B2_MAX_ALPHA + 2];              // (weight << 9) | symbol
  u32 lvl[2][LL_MAXITEMS];                 // merged weights of two consecutive lists
  u32 pkgbits[LL_MAXBITS][LL_BITWORDS];    // bit p set <=> item p of the list is a package
};

__device__ void ll_quick_sort(u32 *a0, i32 n0) {   // :191-223, compares weights only
  // explicit stack of (first, n) ranges; sub-ranges are disjoint so their order is irrelevant
  i32 stk_f[40], stk_n[40];
  int sp = 0;
  stk_f[0] = 0; stk_n[0] = n0; sp = 1;
  while (sp) {
    sp--;
    i32 f = stk_f[sp], nn = stk_n[sp];
    if (nn < 2) continue;
    u32 *a = a0 + f;
    const u32 pw = a[nn / 2] >> 9;
    i32 i = 0, j = nn - 1;
    for (;;) {
      while ((a[i] >> 9) < pw) i++;
      while (pw < (a[j] >> 9)) j--;
      if (i >= j) break;
      u32 t = a[i]; a[i] = a[j]; a[j] = t;
      i++; j--;
    }
    // Quick_sort (a (first .. first+i-1)); Quick_sort (a (first+i .. last)); larger range pushed first
    i32 n1 = i, n2 = nn - i;
    if (n1 > n2) {
      stk_f[sp] = f; stk_n[sp] = n1; sp++;
      stk_f[sp] = f + i; stk_n[sp] = n2; sp++;
    } else {
      stk_f[sp] = f + i; stk_n[sp] = n2; sp++;
      stk_f[sp] = f; stk_n[sp] = n1; sp++;
    }
  }
}

// One warp.  counts[0..n-1] (already through Avoid_Zeros) -> lens[0..n-1].
__device__ void ll_length_limited_warp(LLScratch &S, const u32 *counts, int n, int max_bits, u8 *lens) {
  const u32 l = lane_id();
  const u32 lt = (1u << l) - 1u;
  int ns = 0;
  for (int base = 0; base < n; base += 32) {                // leaves in alphabet order (:230-235)
    const int a = base + (int)l;
    const u32 c = a < n ? counts[a] : 0;
    const u32 m = __ballot_sync(0xffffffffu, c > 0);
    if (c > 0) S.leaf[ns + __popc(m & lt)] = (c << 9) | (u32)a;
    ns += __popc(m);
    if (a < n) lens[a] = 0;
  }
  __syncwarp();
  if (ns == 0) return;
  if (ns == 1) { if (l == 0) lens[S.leaf[0] & 511u] = 1; __syncwarp(); return; }
  if (l == 0) ll_quick_sort(S.leaf, ns);
  __syncwarp();
  const int need = 2 * ns - 2;
  for (int i = l; i < ns; i += 32) S.lvl[0][i] = S.leaf[i] >> 9;
  if (l < LL_BITWORDS) S.pkgbits[0][l] = 0;
  int len_prev = ns;
  __syncwarp();
  for (int lev = 1; lev < max_bits; lev++) {
    const u32 *prev = S.lvl[(lev - 1) & 1];
    u32 *cur = S.lvl[lev & 1];
    const int npk = len_prev >> 1;
    if (l < LL_BITWORDS) S.pkgbits[lev][l] = 0;
    __syncwarp();
    for (int a = l; a < ns; a += 32) {
      const u32 w = S.leaf[a] >> 9;
      int lo = 0, hi = npk;                                 // packages with sum <= w go before this leaf
      while (lo < hi) { int mid = (lo + hi) >> 1; if (prev[2 * mid] + prev[2 * mid + 1] <= w) lo = mid + 1; else hi = mid; }
      const int pos = a + lo;
      if (pos < need) cur[pos] = w;
    }
    for (int b = l; b < npk; b += 32) {
      const u32 pk = prev[2 * b] + prev[2 * b + 1];
      int lo = 0, hi = ns;                                  // leaves with weight < sum go before this package
      while (lo < hi) { int mid = (lo + hi) >> 1; if ((S.leaf[mid] >> 9) < pk) lo = mid + 1; else hi = mid; }
      const int pos = b + lo;
      if (pos < need) { cur[pos] = pk; atomicOr(&S.pkgbits[lev][pos >> 5], 1u << (pos & 31)); }
    }
    len_prev = min(need, ns + npk);
    __syncwarp();
  }
  // walk down from the last list
  u32 cnt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int k = need;
  for (int lev = max_bits - 1; lev >= 0; lev--) {
    const u32 wv = (l < LL_BITWORDS) ? S.pkgbits[lev][l] : 0u;
    const int lo = (int)l * 32;
    const u32 msk = (k >= lo + 32) ? 0xFFFFFFFFu : (k <= lo ? 0u : ((1u << (k - lo)) - 1u));
    u32 p = __popc(wv & msk);
#pragma unroll
    for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    const int a = k - (int)p;                               // leaves taken in this list
#pragma unroll
    for (int j = 0; j < 9; j++) cnt[j] += ((int)l + 32 * j < a);
    k = 2 * (int)p;
  }
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const int i = (int)l + 32 * j;
    if (i < ns) lens[S.leaf[i] & 511u] = (u8)cnt[j];
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// selector MTF list (:669-717, :816-835) kept as positions: pos[cl] = 1-based place of coder cl+1.
// Moving the coder at place p to the front increments every place < p and sets its own to 1.
// ---------------------------------------------------------------------------------------------
struct SelList {
  u32 pos[6];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int c = 0; c < 6; c++) pos[c] = (u32)c + 1;
  }
  __device__ __forceinline__ u32 place(u32 cl0) const {      // cl0 = coder - 1
    u32 p = pos[0];
#pragma unroll
    for (int c = 1; c < 6; c++) p = (cl0 == (u32)c) ? pos[c] : p;
    return p;
  }
  __device__ __forceinline__ void to_front(u32 cl0, u32 p) {
#pragma unroll
    for (int c = 0; c < 6; c++) pos[c] = (cl0 == (u32)c) ? 1u : pos[c] + (pos[c] < p ? 1u : 0u);
  }
};

#define CT_THREADS 256

struct ConstructSmem {
  u32 hist[B2_MAX_CODERS][B2_MAX_ALPHA + 2];
  unsigned long long lenpack[B2_MAX_ALPHA + 2];
  u8 lens[B2_MAX_CODERS][B2_MAX_ALPHA + 2];
  LLScratch ll[B2_MAX_CODERS];
  u32 stat[8];
  u32 red[40];
  u32 defectors;
  u32 selcost;
};

// Define_Descriptors (:635-657) given selectors: histograms -> Avoid_Zeros -> lengths -> lenpack
__device__ void ct_define_descriptors(ConstructSmem &S, const u16 *__restrict__ m, const u8 *__restrict__ sel,
                                      u32 M, u32 G, int A, int ec, int max_len) {
  const u32 tid = threadIdx.x;
  for (u32 i = tid; i < B2_MAX_CODERS * (B2_MAX_ALPHA + 2); i += CT_THREADS) (&S.hist[0][0])[i] = 0;
  __syncthreads();
  for (u32 g = tid; g < G; g += CT_THREADS) {
    const u32 c = sel[g] - 1;
    const u32 s0 = g * B2_GROUP_SIZE, s1 = min(s0 + B2_GROUP_SIZE, M);
    u32 n0 = 0, n1 = 0;
    for (u32 s = s0; s < s1; s++) {
      u32 sym = m[s];
      if (sym == 0) n0++;
      else if (sym == 1) n1++;
      else atomicAdd(&S.hist[c][sym], 1u);
    }
    if (n0) atomicAdd(&S.hist[c][0], n0);
    if (n1) atomicAdd(&S.hist[c][1], n1);
  }
  __syncthreads();
  // one warp per coder: Avoid_Zeros (:439-462) in parallel, then the serial length limiter on lane 0
  const u32 w = warp_id(), l = lane_id();
  if ((int)w < ec) {
    u32 zeroes = 0;
    for (int s = l; s < A; s += 32) zeroes += (S.hist[w][s] == 0);
#pragma unroll
    for (int o = 16; o; o >>= 1) zeroes += __shfl_xor_sync(0xffffffffu, zeroes, o);
    if (zeroes > 0 && zeroes <= 100) { for (int s = l; s < A; s += 32) S.hist[w][s] = max(1u, S.hist[w][s]); }
    else if (zeroes > 100) { for (int s = l; s < A; s += 32) { u32 v = S.hist[w][s]; S.hist[w][s] = v == 0 ? 1u : v * 2u; } }
    __syncwarp();
    ll_length_limited_warp(S.ll[w], S.hist[w], A, max_len, S.lens[w]);
  }
  __syncthreads();
  for (int s = tid; s < A; s += CT_THREADS) {
    unsigned long long p = 0;
    for (int c = 0; c < ec; c++) p |= (unsigned long long)S.lens[c][s] << (10 * c);
    S.lenpack[s] = p;
  }
  __syncthreads();
}

__device__ void ct_group_costs(ConstructSmem &S, const u16 *__restrict__ m, u32 M, u32 G,
                               unsigned long long *__restrict__ gcost) {
  for (u32 g = threadIdx.x; g < G; g += CT_THREADS) {
    const u32 s0 = g * B2_GROUP_SIZE, s1 = min(s0 + B2_GROUP_SIZE, M);
    unsigned long long acc = 0;
    for (u32 s = s0; s < s1; s++) acc += S.lenpack[m[s]];
    gcost[g] = acc;
  }
  __syncthreads();
}

// grid: (n_triples, n_jobs)
__global__ void __launch_bounds__(CT_THREADS)
k_construct(const B2Job *__restrict__ jobs, const u16 *__restrict__ mtf, const u32 *__restrict__ rank3,
            const u32 *__restrict__ rank4, u8 *__restrict__ sel_all, unsigned long long *__restrict__ gcost_all,
            u8 *__restrict__ lens_all, u32 *__restrict__ cost_all, u32 *__restrict__ low_all,
            int level, int n_triples, u32 total_groups) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ConstructSmem &S = *reinterpret_cast<ConstructSmem *>(smem_raw);
  const int t = blockIdx.x;
  const u32 jb = blockIdx.y;
  const B2Job &job = jobs[jb];
  const u32 M = job.n_mtf, G = job.n_groups;
  const int A = (int)job.n_used + 2;                 // symbols 0 .. EOB
  const u16 *m = mtf + job.mtf_off;
  int max_len, sw, ec;
  b2_triple(level, t, max_len, sw, ec);
  u8 *sel = sel_all + (size_t)t * total_groups + job.grp_off;
  unsigned long long *gcost = gcost_all + (size_t)t * total_groups + job.grp_off;
  const u32 tid = threadIdx.x;

  // Initial_Clustering_by_Rank (:572-588, :625-631)
  {
    const u32 *rk = (sw == 3 ? rank3 : rank4) + job.grp_off;
    const int attr_tab[5][6] = {{2, 1, 0, 0, 0, 0}, {3, 1, 2, 0, 0, 0}, {4, 2, 1, 3, 0, 0}, {5, 3, 1, 2, 4, 0}, {6, 4, 2, 1, 3, 5}};
    for (u32 i = tid; i < G; i += CT_THREADS) {
      // rank position i+1 belongs to range a (1-based) iff low_a <= i+1 <= high_a, high_a = a*G/ec
      u32 pos1 = i + 1;
      int a = 1;
      while ((u64)a * G / ec < pos1) a++;
      sel[(rk[i] & 0xFFFFu) - 1] = (u8)attr_tab[ec - 2][a - 1];
    }
  }
  __syncthreads();

  u32 defectors = 0;
  for (int iteration = 1; iteration <= 10; iteration++) {                 // :793-802
    ct_define_descriptors(S, m, sel, M, G, A, ec, max_len);
    ct_group_costs(S, m, M, G, gcost);
    // Simulate_Entropy_Coding_Variants_and_Reclassify (:661-753): serial through the selector MTF list
    if (warp_id() == 0) {
      const u32 l = lane_id();
      SelList L; L.init();
      u32 def = 0;
      for (u32 g0 = 0; g0 < G; g0 += 32) {
        const u32 g = g0 + l;
        unsigned long long c = g < G ? gcost[g] : 0;
        u32 cur = g < G ? sel[g] : 0;
        u32 mine = cur;
        const u32 cntk = min(32u, G - g0);
        for (u32 k = 0; k < cntk; k++) {
          const unsigned long long ck = __shfl_sync(0xffffffffu, c, k);
          const u32 clk = __shfl_sync(0xffffffffu, cur, k);
          // key = (cost << 6) | (coder0 << 3) | place: the minimum is the cheapest coder, lowest
          // coder on ties (strict "<" scanning cl upward, :691-695), and carries its place along
          u32 key = 0xFFFFFFFFu;
#pragma unroll
          for (int cl = 0; cl < 6; cl++) {
            if (cl < ec) {
              const u32 cost = (u32)((ck >> (10 * cl)) & 1023u) + L.pos[cl];
              key = min(key, (cost << 6) | ((u32)cl << 3) | L.pos[cl]);
            }
          }
          const u32 best0 = (key >> 3) & 7u;
          if (best0 + 1 != clk) { def++; if (l == k) mine = best0 + 1; }
          L.to_front(best0, key & 7u);
        }
        if (g < G && mine != cur) sel[g] = (u8)mine;
      }
      if (l == 0) S.defectors = def;
    }
    __syncthreads();
    defectors = S.defectors;
    if (defectors == 0) break;
  }
  if (defectors > 0) {                                                    // :803-807
    ct_define_descriptors(S, m, sel, M, G, A, ec, max_len);
    ct_group_costs(S, m, M, G, gcost);
  }
  // Cluster_Statistics (:757-778)
  if (tid < 8) S.stat[tid] = 0;
  if (tid == 0) S.selcost = 0;
  __syncthreads();
  u32 data_bits = 0;
  for (u32 g = tid; g < G; g += CT_THREADS) {
    u32 c = sel[g];
    atomicAdd(&S.stat[c], 1u);
    data_bits += (u32)((gcost[g] >> (10 * (c - 1))) & 1023u);
  }
  // Compute_Selectors_Cost (:815-837), serial on warp 0
  if (warp_id() == 0) {
    const u32 l = lane_id();
    SelList L; L.init();
    u32 bits = 0;
    for (u32 g0 = 0; g0 < G; g0 += 32) {
      const u32 g = g0 + l;
      u32 cur = g < G ? sel[g] : 1;
      const u32 cntk = min(32u, G - g0);
      for (u32 k = 0; k < cntk; k++) {
        const u32 cl0 = __shfl_sync(0xffffffffu, cur, k) - 1;
        const u32 p = L.place(cl0);
        bits += p;
        L.to_front(cl0, p);
      }
    }
    if (l == 0) S.selcost = bits;
  }
  // Compute_Huffman_Bit_Lengths_Cost (:839-865): 5 + sum (2*|delta| + 1)
  u32 len_bits = 0;
  for (int i = tid; i < ec * A; i += CT_THREADS) {
    int c = i / A, s = i % A;
    int cur = s == 0 ? S.lens[c][0] : S.lens[c][s - 1];
    int nw = S.lens[c][s];
    int dlt = nw > cur ? nw - cur : cur - nw;
    len_bits += 2 * dlt + 1 + (s == 0 ? 5 : 0);
  }
  u32 part = data_bits + len_bits;
#pragma unroll
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  __syncthreads();
  if (lane_id() == 0) S.red[warp_id()] = part;
  __syncthreads();
  if (tid == 0) {
    u32 total = S.selcost;
    for (int w = 0; w < CT_THREADS / 32; w++) total += S.red[w];
    const u32 uniform_usage = G / (u32)ec;
    u32 low = 0;
    for (int c = 1; c <= ec; c++) if (S.stat[c] < uniform_usage / 2) low = 1;
    cost_all[(size_t)jb * B2_N_TRIPLES + t] = total;
    low_all[(size_t)jb * B2_N_TRIPLES + t] = low;
  }
  u8 *lo = lens_all + ((size_t)jb * B2_N_TRIPLES + t) * (B2_MAX_CODERS * B2_MAX_ALPHA);
  for (int i = tid; i < B2_MAX_CODERS * B2_MAX_ALPHA; i += CT_THREADS) {
    int c = i / B2_MAX_ALPHA, s = i % B2_MAX_ALPHA;
    lo[i] = (c < ec && s < A) ? S.lens[c][s] : 0;
  }
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
  u32 choice_mask;   // bit ec set if ec in coder_choices (:907-915)
  if (level == 9) {
    if (M <= 5000) choice_mask = (1u << 2) | (1u << 3) | (1u << 6);
    else if (M <= 10000) choice_mask = (1u << 3) | (1u << 4) | (1u << 6);
    else choice_mask = (1u << 3) | (1u << 4) | (1u << 5) | (1u << 6);
  } else choice_mask = (1u << 4) | (1u << 6);
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
                const u16 *d_mtf, u32 *d_rank3, u32 *d_rank4, u8 *d_sel, unsigned long long *d_gcost,
                u8 *d_lens, u32 *d_cost, u32 *d_low, int level) {
  static bool attr_set = false;
  if (!attr_set) {
    B2_CUDA_CHECK(cudaFuncSetAttribute(k_construct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConstructSmem)));
    B2_CUDA_CHECK(cudaFuncSetAttribute(k_rank_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 18004 * 4));
    attr_set = true;
  }
  if (n_jobs == 0) return 0;
  const int n_triples = level == 9 ? 20 : 5;
  k_group_keys<<<n_jobs, 256, 0, st>>>(d_jobs, d_mtf, d_rank3, d_rank4);
  size_t sort_smem = ((size_t)max_groups_per_job + 2) * 4;
  k_rank_sort<<<n_jobs * 2, 32, sort_smem, st>>>(d_jobs, d_rank3, d_rank4);
  dim3 grid(n_triples, n_jobs);
  k_construct<<<grid, CT_THREADS, sizeof(ConstructSmem), st>>>(d_jobs, d_mtf, d_rank3, d_rank4, d_sel, d_gcost, d_lens,
                                                             d_cost, d_low, level, n_triples, total_groups);
  k_choose<<<(n_jobs + 127) / 128, 128, 0, st>>>(d_jobs, n_jobs, d_cost, d_low, level, n_triples);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}
