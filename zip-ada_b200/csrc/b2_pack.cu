// Stages A10-A15: canonical codes, block header, tree descriptors, entropy-coded data, and the
// final shift-concatenation of the winning blocks into the stream.
//
// Reference: zip_lib/huffman-encoding.adb:45-80 (Prepare_Codes), zip_lib/bzip2-encoding.adb:984-994
// (Put_Block_Header), :996-1086 (Put_Block_Trees_Descriptors), :1088-1112 (Entropy_Output),
// zip_lib/bzip2-buffers.adb:18-31 (MSB-first Put_Bits).
//
// Every block is packed at bit offset 0 of its own word-aligned slot (a candidate's bitstream does
// not depend on the incoming bit offset); bit offsets inside the block come from
// prefix sums of code lengths (R11).  The stream is a sequence of big-endian 32-bit words: words
// are assembled MSB-first in registers and byte-swapped on the way to memory.
#include "b2_common.cuh"
#include "b2_kernels.h"

#define PK_THREADS 256

__device__ __forceinline__ u32 bswap32(u32 x) { return __byte_perm(x, 0, 0x0123); }

// OR `n` (1..32) low bits of `value` into the stream at bit position `p` (MSB-first).
__device__ __forceinline__ void put_bits_atomic(u32 *words, u64 p, u32 value, u32 n) {
  u32 sh = (u32)(p & 31);
  u64 v64 = ((u64)value << (64 - n)) >> sh;
  u32 hi = (u32)(v64 >> 32), lo = (u32)v64;
  u64 w = p >> 5;
  if (hi) atomicOr(&words[w], bswap32(hi));
  if (lo) atomicOr(&words[w + 1], bswap32(lo));
}

struct PackSmem {
  u32 code[B2_MAX_CODERS][B2_MAX_ALPHA + 2];
  u8 len[B2_MAX_CODERS][B2_MAX_ALPHA + 2];
  u32 scan[40];
  u32 table_bits[B2_MAX_CODERS];
  u64 sec_sel, sec_tab, sec_data;   // bit offsets of the sections
};

// selector MTF helpers (same packing as in b2_entropy.cu)
__device__ __forceinline__ int pk_pos(u32 v, int cl) {
  int p = 1;
#pragma unroll
  for (int w = 0; w < 6; w++) { if (((v >> (4 * w)) & 15u) == (u32)cl) p = w + 1; }
  return p;
}
__device__ __forceinline__ u32 pk_front(u32 v, int pos, int cl) {
  u32 lowmask = (1u << (4 * (pos - 1))) - 1u;
  u32 keep_hi = v & ~((1u << (4 * pos)) - 1u);
  return keep_hi | ((v & lowmask) << 4) | (u32)cl;
}

__global__ void __launch_bounds__(PK_THREADS)
k_pack(const B2Job *__restrict__ jobs, const u16 *__restrict__ mtf, const u8 *__restrict__ sel_all,
       const u8 *__restrict__ lens_all, u8 *__restrict__ selpos_all, u32 *__restrict__ bits, int level, u32 total_groups) {
  __shared__ PackSmem S;
  const B2Job &job = jobs[blockIdx.x];
  const u32 M = job.n_mtf, G = job.n_groups;
  const int A = (int)job.n_used + 2;
  const u16 *m = mtf + job.mtf_off;
  const int t = (int)job.best;
  int max_len, sw, ec;
  b2_triple(level, t, max_len, sw, ec);
  const u8 *sel = sel_all + (size_t)t * total_groups + job.grp_off;
  u8 *selpos = selpos_all + job.grp_off;
  const u8 *lens = lens_all + ((size_t)blockIdx.x * B2_N_TRIPLES + t) * (B2_MAX_CODERS * B2_MAX_ALPHA);
  u32 *w = bits + job.bits_off;
  const u32 tid = threadIdx.x;

  for (int i = tid; i < ec * A; i += PK_THREADS) { int c = i / A, s = i % A; S.len[c][s] = lens[c * B2_MAX_ALPHA + s]; }
  __syncthreads();
  // Prepare_Codes (huffman-encoding.adb:45-80), one thread per coder; also the size of each table
  if ((int)tid < ec) {
    const int c = tid;
    u32 bl_count[LL_MAXCODE + 1], next_code[LL_MAXCODE + 1];
    for (int b = 0; b <= LL_MAXCODE; b++) { bl_count[b] = 0; next_code[b] = 0; }
    for (int s = 0; s < A; s++) bl_count[S.len[c][s]]++;
    u32 code = 0;
    for (int b = 1; b <= max_len; b++) { code = (code + bl_count[b - 1]) * 2; next_code[b] = code; }
    u32 tb = 5;
    int cur = S.len[c][0];
    for (int s = 0; s < A; s++) {
      int bl = S.len[c][s];
      if (bl > 0) { S.code[c][s] = next_code[bl]; next_code[bl]++; } else S.code[c][s] = 0;
      int d = bl > cur ? bl - cur : cur - bl;
      tb += 2 * d + 1;
      cur = bl;
    }
    S.table_bits[c] = tb;
  }
  // selector MTF positions (:1032-1050), serial on warp 0
  if (warp_id() == 0) {
    const u32 l = lane_id();
    u32 list = 0;
    for (int q = 0; q < ec; q++) list |= (u32)(q + 1) << (4 * q);
    for (u32 g0 = 0; g0 < G; g0 += 32) {
      const u32 g = g0 + l;
      u32 cur = g < G ? sel[g] : 1;
      u32 mypos = 0;
      const u32 cntk = min(32u, G - g0);
      for (u32 k = 0; k < cntk; k++) {
        u32 clk = __shfl_sync(0xffffffffu, cur, k);
        int p = pk_pos(list, (int)clk);
        if (l == k) mypos = (u32)p;
        list = pk_front(list, p, (int)clk);
      }
      if (g < G) selpos[g] = (u8)mypos;
    }
  }
  __syncthreads();

  // ---- header + mapping table + counts (thread 0; < 400 bits) -----------------------------------
  u32 ranges = 0;
  for (int i = 0; i < 16; i++) {
    u32 wv = job.in_use[i >> 1];
    u32 h = (i & 1) ? (wv >> 16) : (wv & 0xFFFFu);
    if (h) ranges++;
  }
  const u64 sec_sel = 105ull + 16ull + 16ull * ranges + 3ull + 15ull;
  if (tid == 0) {
    u64 p = 0;
    put_bits_atomic(w, p, 0x314159u, 24); p += 24;          // block_header_magic "1AY&SY" (bzip2.ads:122)
    put_bits_atomic(w, p, 0x265359u, 24); p += 24;
    put_bits_atomic(w, p, job.crc, 32); p += 32;            // :988
    put_bits_atomic(w, p, 0, 1); p += 1;                    // randomised = False (:992)
    put_bits_atomic(w, p, job.origin, 24); p += 24;         // :993
    u32 map16 = 0;
    for (int i = 0; i < 16; i++) {
      u32 wv = job.in_use[i >> 1];
      u32 h = (i & 1) ? (wv >> 16) : (wv & 0xFFFFu);
      if (h) map16 |= 1u << (15 - i);
    }
    put_bits_atomic(w, p, map16, 16); p += 16;              // :1010-1012
    for (int i = 0; i < 16; i++) {                          // :1014-1020
      u32 wv = job.in_use[i >> 1];
      u32 h = (i & 1) ? (wv >> 16) : (wv & 0xFFFFu);
      if (h) {
        u32 rev = __brev(h) >> 16;                          // in_use (16 i + j) for j = 0..15, MSB first
        put_bits_atomic(w, p, rev, 16); p += 16;
      }
    }
    put_bits_atomic(w, p, (u32)ec, 3); p += 3;              // :1083
    put_bits_atomic(w, p, G, 15); p += 15;                  // :1027
  }
  // ---- selectors, unary (:1045-1049) --------------------------------------------------------
  u32 sel_run = 0;
  for (u32 g0 = 0; g0 < G; g0 += PK_THREADS) {
    const u32 g = g0 + tid;
    u32 p = g < G ? selpos[g] : 0;
    u32 tot;
    u32 ex = block_excl_add(p, S.scan, &tot);
    if (g < G) put_bits_atomic(w, sec_sel + sel_run + ex, ((1u << (p - 1)) - 1u) << 1, p);
    sel_run += tot;
  }
  // ---- code length tables (:1053-1079) ------------------------------------------------------
  const u64 sec_tab = sec_sel + sel_run;
  u64 sec_data = sec_tab;
  for (int c = 0; c < ec; c++) sec_data += S.table_bits[c];
  if ((int)tid < ec) {
    const int c = tid;
    u64 p = sec_tab;
    for (int q = 0; q < c; q++) p += S.table_bits[q];
    int cur = S.len[c][0];
    put_bits_atomic(w, p, (u32)cur, 5); p += 5;
    for (int s = 0; s < A; s++) {
      int nw = S.len[c][s];
      while (cur != nw) {
        if (cur < nw) { put_bits_atomic(w, p, 2u, 2); cur++; }   // '1','0'
        else { put_bits_atomic(w, p, 3u, 2); cur--; }            // '1','1'
        p += 2;
      }
      p += 1;                                                    // '0' (already zero)
    }
  }
  // ---- data (:1094-1109) -------------------------------------------------------------------
  u64 data_run = 0;
  for (u32 g0 = 0; g0 < G; g0 += PK_THREADS) {
    const u32 g = g0 + tid;
    u32 gb = 0;
    u32 c = 0, s0 = 0, s1 = 0;
    if (g < G) {
      c = sel[g] - 1; s0 = g * B2_GROUP_SIZE; s1 = min(s0 + B2_GROUP_SIZE, M);
      for (u32 s = s0; s < s1; s++) gb += S.len[c][m[s]];
    }
    u32 tot;
    u32 ex = block_excl_add(gb, S.scan, &tot);
    if (g < G) {
      u64 p = sec_data + data_run + ex;
      // accumulate MSB-first into a 64-bit window aligned to the word grid
      u64 wi = p >> 5;
      u32 fill = (u32)(p & 31);          // bits already occupied in the current word (by other threads)
      u64 acc = 0;                       // top `fill + pending` bits of a 64-bit big-endian window
      u32 nb = fill;
      for (u32 s = s0; s < s1; s++) {
        u32 sym = m[s];
        u32 ln = S.len[c][sym], cd = S.code[c][sym];
        acc |= ((u64)cd << (64 - ln)) >> nb;
        nb += ln;
        if (nb >= 32) {
          atomicOr(&w[wi], bswap32((u32)(acc >> 32)));
          wi++; acc <<= 32; nb -= 32;
        }
      }
      if (nb) { u32 hi = (u32)(acc >> 32); if (hi) atomicOr(&w[wi], bswap32(hi)); }
    }
    data_run += tot;
  }
}

// ---------------------------------------------------------------------------------------------
// Shift-concatenation: copy `nbits` bits of a packed block (starting at bit 0 of word `src_word`)
// to bit offset `dst_bit` of the stream.  One thread per destination word.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_concat(const B2ConcatItem *__restrict__ items, const u32 *__restrict__ bits, u32 *__restrict__ out) {
  const B2ConcatItem it = items[blockIdx.y];
  const u32 sh = (u32)(it.dst_bit & 31);
  const u64 w0 = it.dst_bit >> 5;
  const u64 nwords_src = (it.nbits + 31) >> 5;
  const u64 nwords_dst = (sh + it.nbits + 31) >> 5;
  const u32 *src = bits + it.src_word;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < nwords_dst; j += (u64)gridDim.x * blockDim.x) {
    u32 a = (j >= 1 && j - 1 < nwords_src) ? bswap32(src[j - 1]) : 0;
    u32 b = (j < nwords_src) ? bswap32(src[j]) : 0;
    u32 v = sh ? ((a << (32 - sh)) | (b >> sh)) : b;
    if (v) {
      if (j == 0 || j + 1 >= nwords_dst) atomicOr(&out[w0 + j], bswap32(v));
      else out[w0 + j] = bswap32(v);
    }
  }
}

// Stream header "BZh<level>" (:1384-1391) and footer (:1395-1407): 48-bit magic 0x177245385090 and
// the combined CRC at the stream's current bit offset; the zero padding to a byte is already there.
__global__ void k_stream_ends(const B2StreamEnd *__restrict__ ends, u32 n, int level, u32 *__restrict__ out) {
  const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const B2StreamEnd E = ends[s];
  u32 *w = out + (E.out_off >> 2);                 // regions are 8-byte aligned
  // pad bit 0: no header, bit 1: no footer (a shard of a stream that is not its first / last, b2_shard_finish)
  if (!(E.pad & 1u)) atomicOr(&w[0], bswap32(0x425A6800u | (u32)('0' + level)));
  if (E.pad & 2u) return;
  u64 p = E.end_bit;
  put_bits_atomic(w, p, 0x177245u, 24); p += 24;
  put_bits_atomic(w, p, 0x385090u, 24); p += 24;
  put_bits_atomic(w, p, E.crc, 32);
}

// Copies every stream from its region to its packed place (both 8-byte aligned), 8 bytes per thread.
__global__ void __launch_bounds__(256)
k_pack_streams(const B2PackItem *__restrict__ items, const u8 *__restrict__ src, u8 *__restrict__ dst) {
  const B2PackItem it = items[blockIdx.x];
  const u64 nw = (it.len + 7) >> 3;
  const u64 *s = reinterpret_cast<const u64 *>(src + it.src_off);
  u64 *d = reinterpret_cast<u64 *>(dst + it.dst_off);
  for (u64 i = threadIdx.x; i < nw; i += blockDim.x) d[i] = s[i];
}

__global__ void k_bits_layout(B2Job *jobs, u32 n_jobs, u64 *total_words) {
  // serial exclusive scan of per-block word counts (<= a few thousand blocks)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    u64 run = 0;
    for (u32 j = 0; j < n_jobs; j++) {
      jobs[j].bits_off = run;
      run += ((jobs[j].nbits + 31) >> 5) + 2;   // +2: put_bits may touch word+1
    }
    *total_words = run;
  }
}

int b2k_stream_ends(cudaStream_t st, const B2StreamEnd *d_ends, u32 n, int level, u32 *d_out) {
  if (n == 0) return 0;
  k_stream_ends<<<(n + 127) / 128, 128, 0, st>>>(d_ends, n, level, d_out);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int b2k_pack_streams(cudaStream_t st, const B2PackItem *d_items, u32 n, const u8 *d_src, u8 *d_dst) {
  if (n == 0) return 0;
  k_pack_streams<<<n, 256, 0, st>>>(d_items, d_src, d_dst);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int b2k_bits_layout(cudaStream_t st, B2Job *d_jobs, u32 n_jobs, u64 *d_total_words) {
  k_bits_layout<<<1, 32, 0, st>>>(d_jobs, n_jobs, d_total_words);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int b2k_pack(cudaStream_t st, const B2Job *d_jobs, u32 n_jobs, const u16 *d_mtf, const u8 *d_sel, const u8 *d_lens,
             u8 *d_selpos, u32 *d_bits, int level, u32 total_groups) {
  if (n_jobs == 0) return 0;
  k_pack<<<n_jobs, PK_THREADS, 0, st>>>(d_jobs, d_mtf, d_sel, d_lens, d_selpos, d_bits, level, total_groups);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int b2k_concat(cudaStream_t st, const B2ConcatItem *d_items, u32 n_items, const u32 *d_bits, u32 *d_out) {
  if (n_items == 0) return 0;
  dim3 grid(64, n_items);
  k_concat<<<grid, 256, 0, st>>>(d_items, d_bits, d_out);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}
