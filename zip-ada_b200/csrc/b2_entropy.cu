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
// Length-limited code lengths, literal restatement of the reference's boundary package-merge
// (huffman-encoding-length_limited_coding.adb).  Serial; runs on one lane, scratch in shared
// memory.  Recursion (:131-163) is unrolled on an explicit LIFO of list indices, which preserves
// the depth-first order of the two recursive calls.
// ---------------------------------------------------------------------------------------------
#define LL_MAXBITS 17
#define LL_POOL (2 * LL_MAXBITS * (LL_MAXBITS + 1))
#define LL_NULL 0xFFFFu

struct LLScratch {
  u32 leaf_w[B2_MAX_ALPHA];
  u32 node_w[LL_POOL];
  u16 leaf_s[B2_MAX_ALPHA];
  u16 node_cnt[LL_POOL];
  u16 node_tail[LL_POOL];
  u16 lists[LL_MAXBITS][2];
  u8 node_use[LL_POOL];
  u8 stack[2 * LL_MAXBITS + 4];
};

__device__ void ll_quick_sort(LLScratch &S, i32 first, i32 n) {   // :191-223
  // explicit stack of (first, n) ranges; sub-ranges are disjoint so their order is irrelevant
  i32 stk_f[40], stk_n[40];
  int sp = 0;
  stk_f[0] = first; stk_n[0] = n; sp = 1;
  while (sp) {
    sp--;
    i32 f = stk_f[sp], nn = stk_n[sp];
    if (nn < 2) continue;
    u32 pw = S.leaf_w[f + nn / 2];
    i32 i = 0, j = nn - 1;
    for (;;) {
      while (S.leaf_w[f + i] < pw) i++;
      while (pw < S.leaf_w[f + j]) j--;
      if (i >= j) break;
      u32 tw = S.leaf_w[f + i]; S.leaf_w[f + i] = S.leaf_w[f + j]; S.leaf_w[f + j] = tw;
      u16 ts = S.leaf_s[f + i]; S.leaf_s[f + i] = S.leaf_s[f + j]; S.leaf_s[f + j] = ts;
      i++; j--;
    }
    // Quick_sort (a (first .. first+i-1)); Quick_sort (a (first+i .. last))
    // push the larger range first so that the stack stays shallow
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

__device__ u32 ll_get_free_node(LLScratch &S, int max_bits, u32 &pool_next, bool use_lists) {   // :98-122
  const u32 pool_size = 2u * max_bits * (max_bits + 1);
  for (;;) {
    if (pool_next >= pool_size) {
      for (u32 i = 0; i < pool_size; i++) S.node_use[i] = 0;
      if (use_lists) {
        for (int i = 0; i < max_bits * 2; i++) {
          u32 node = S.lists[i / 2][i % 2];
          while (node != LL_NULL) { S.node_use[node] = 1; node = S.node_tail[node]; }
        }
      }
      pool_next = 0;
    }
    if (!S.node_use[pool_next]) break;
    pool_next++;
  }
  pool_next++;
  return pool_next - 1;
}

// counts[0..n-1] (already through Avoid_Zeros) -> lens[0..n-1]
__device__ void ll_length_limited(LLScratch &S, const u32 *counts, int n, int max_bits, u8 *lens) {
  const u32 pool_size = 2u * max_bits * (max_bits + 1);
  for (u32 i = 0; i < pool_size; i++) { S.node_use[i] = 0; S.node_tail[i] = LL_NULL; }
  u32 pool_next = 0;
  i32 num_symbols = 0;
  for (int a = 0; a < n; a++) lens[a] = 0;
  for (int a = 0; a < n; a++)
    if (counts[a] > 0) { S.leaf_w[num_symbols] = counts[a]; S.leaf_s[num_symbols] = (u16)a; num_symbols++; }
  if (num_symbols == 0) return;
  if (num_symbols == 1) { lens[S.leaf_s[0]] = 1; return; }
  ll_quick_sort(S, 0, num_symbols);
  auto init_node = [&](u32 weight, u32 count, u32 tail, u32 idx) {
    S.node_w[idx] = weight; S.node_cnt[idx] = (u16)count; S.node_tail[idx] = (u16)tail; S.node_use[idx] = 1;
  };
  {  // Init_Lists :167-174
    u32 node0 = ll_get_free_node(S, max_bits, pool_next, false);
    u32 node1 = ll_get_free_node(S, max_bits, pool_next, false);
    init_node(S.leaf_w[0], 1, LL_NULL, node0);
    init_node(S.leaf_w[1], 2, LL_NULL, node1);
    for (int i = 0; i < max_bits; i++) { S.lists[i][0] = (u16)node0; S.lists[i][1] = (u16)node1; }
  }
  const i32 runs = 2 * num_symbols - 4;
  for (i32 run = 1; run <= runs; run++) {
    const bool final_top = (run == runs);
    int sp = 0;
    S.stack[sp++] = (u8)(max_bits - 1);
    bool top = true;
    while (sp) {
      const int index = S.stack[--sp];
      const bool final = top && final_top;
      top = false;
      const u32 lastcount = S.node_cnt[S.lists[index][1]];
      if (index == 0 && (i32)lastcount >= num_symbols) continue;
      const u32 newchain = ll_get_free_node(S, max_bits, pool_next, true);
      const u32 oldchain = S.lists[index][1];
      S.lists[index][0] = (u16)oldchain; S.lists[index][1] = (u16)newchain;
      if (index == 0) {
        init_node(S.leaf_w[lastcount], lastcount + 1, LL_NULL, newchain);
      } else {
        const u32 sum = S.node_w[S.lists[index - 1][0]] + S.node_w[S.lists[index - 1][1]];
        if ((i32)lastcount < num_symbols && sum > S.leaf_w[lastcount]) {
          init_node(S.leaf_w[lastcount], lastcount + 1, S.node_tail[oldchain], newchain);
        } else {
          init_node(sum, lastcount, S.lists[index - 1][1], newchain);
          if (!final) { S.stack[sp++] = (u8)(index - 1); S.stack[sp++] = (u8)(index - 1); }
        }
      }
    }
  }
  // Extract_Bit_Lengths :180-189
  u32 node = S.lists[max_bits - 1][1];
  while (node != LL_NULL) {
    u32 c = S.node_cnt[node];
    for (u32 i = 0; i < c; i++) lens[S.leaf_s[i]]++;
    node = S.node_tail[node];
  }
}

// ---------------------------------------------------------------------------------------------
// selector MTF list (:669-717, :816-835): positions 1..ec, packed 4 bits per entry
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 selmtf_init(int ec) {
  u32 v = 0;
  for (int w = 0; w < ec; w++) v |= (u32)(w + 1) << (4 * w);
  return v;
}
__device__ __forceinline__ int selmtf_pos(u32 v, int cl) {      // 1-based position of coder cl
  int p = 1;
#pragma unroll
  for (int w = 0; w < 6; w++) { if (((v >> (4 * w)) & 15u) == (u32)cl) p = w + 1; }
  return p;
}
__device__ __forceinline__ u32 selmtf_front(u32 v, int pos, int cl) {   // move entry at pos to front
  u32 lowmask = (pos >= 8) ? 0xFFFFFFFFu : ((1u << (4 * (pos - 1))) - 1u);   // entries before pos
  u32 keep_hi = (pos >= 8) ? 0u : (v & ~((1u << (4 * pos)) - 1u));
  return keep_hi | ((v & lowmask) << 4) | (u32)cl;
}

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
    if (l == 0) ll_length_limited(S.ll[w], S.hist[w], A, max_len, S.lens[w]);
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
      u32 list = selmtf_init(ec);
      u32 def = 0;
      for (u32 g0 = 0; g0 < G; g0 += 32) {
        const u32 g = g0 + l;
        unsigned long long c = g < G ? gcost[g] : 0;
        u32 cur = g < G ? sel[g] : 0;
        u32 mine = cur;
        const u32 cntk = min(32u, G - g0);
        for (u32 k = 0; k < cntk; k++) {
          unsigned long long ck = __shfl_sync(0xffffffffu, c, k);
          u32 clk = __shfl_sync(0xffffffffu, cur, k);
          u32 min_bits = 0x7FFFFFFFu;
          u32 best = clk;
#pragma unroll
          for (int cl = 1; cl <= 6; cl++) {
            if (cl <= ec) {
              u32 cost = (u32)((ck >> (10 * (cl - 1))) & 1023u) + (u32)selmtf_pos(list, cl);
              if (cost < min_bits) { min_bits = cost; best = (u32)cl; }
            }
          }
          if (best != clk) { def++; if (l == k) mine = best; }
          list = selmtf_front(list, selmtf_pos(list, (int)best), (int)best);
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
    u32 list = selmtf_init(ec);
    u32 bits = 0;
    for (u32 g0 = 0; g0 < G; g0 += 32) {
      const u32 g = g0 + l;
      u32 cur = g < G ? sel[g] : 1;
      const u32 cntk = min(32u, G - g0);
      for (u32 k = 0; k < cntk; k++) {
        u32 clk = __shfl_sync(0xffffffffu, cur, k);
        int p = selmtf_pos(list, (int)clk);
        bits += (u32)p;
        list = selmtf_front(list, p, (int)clk);
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
