// BZip2 decode / verify on the device, block-parallel.
//
// Reference: zip_lib/bzip2-decoding.adb (Decode_Block :545-609, Receive_Mapping_Table :94-119,
// Receive_Selectors :121-180, Receive_Huffman_Bit_Lengths :233-261, Receive_MTF_Values :293-468,
// BWT_Detransform :470-487, RLE_1 :489-543) — the decoder UnZip uses for method 12
// (unzip-decompress.adb:1898-1915).  It decodes one block after the other because a block's end is only known
// once it is decoded.  Here every block is decoded at the same time:
//   k_v_find    every bit position of the stream is tested for the 48-bit block magic (bzip2.ads:122) and the
//               stream-footer magic (bzip2.ads:118-119): candidates
//   k_v_decode  one warp per candidate (reading the bits of a block is a serial chain: lane 0; the lanes share the
//               move-to-front shifts and the zero-run fills): header, used map, selectors, code lengths, canonical
//               decoding tables, Huffman symbols -> inverse MTF / zero runs -> the counting pass and the pointer
//               walk of the inverse BWT -> RLE1 decoding with the block CRC over the raw bytes.  The RLE1-coded bytes are kept so that the raw bytes can be compared with an
//               expected input once every block knows where its bytes start.
//   host        follows the chain "a block starts where the one before it ended" from bit 32 to the footer —
//               candidates that are not on the chain (a magic inside compressed data) are never looked at —
//               and folds the block CRCs into the stream CRC (bzip2-decoding.adb: computed_combined_crc).
//   k_v_compare one warp per block of the chain: RLE1 expansion against the expected bytes.
// Randomised blocks (obsolete, never written by the reference's encoder) are reported as unsupported.
#include "b2_common.cuh"
#include "b2_kernels.h"

#define V_MAGIC_BLOCK 0x314159265359ull
#define V_MAGIC_END 0x177245385090ull

__device__ __forceinline__ u32 v_bswap(u32 x) { return __byte_perm(x, 0, 0x0123); }

// ---- candidates ------------------------------------------------------------------------------------
// Thread j looks at the 32 bit positions that start inside the 4-byte word j (the stream is padded with zeros).
__global__ void __launch_bounds__(256)
k_v_find(const u32 *__restrict__ words, u64 n_words, u64 n_bits, u64 *__restrict__ cand, u32 *__restrict__ n_cand, u32 cap) {
  const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_words) return;
  const u32 w0 = v_bswap(words[j]), w1 = v_bswap(words[j + 1]), w2 = v_bswap(words[j + 2]);
  const u64 hi = ((u64)w0 << 32) | w1;
#pragma unroll 4
  for (u32 s = 0; s < 32; s++) {
    // 48 bits starting at bit s of the 96-bit window
    const u64 win = s == 0 ? hi : ((hi << s) | ((u64)w2 >> (32 - s)));
    const u64 v = win >> 16;
    if (v == V_MAGIC_BLOCK || v == V_MAGIC_END) {
      const u64 bit = j * 32 + s;
      if (bit + 48 <= n_bits) {
        const u32 k = atomicAdd(n_cand, 1u);
        if (k < cap) cand[k] = (bit << 1) | (v == V_MAGIC_END ? 1ull : 0ull);
      }
    }
  }
}

// ---- one block ------------------------------------------------------------------------------------------
struct VBits {
  const u8 *p; u64 pos, end;       // next bit to read, number of bits in the stream (the buffer is padded with zeros)
  __device__ __forceinline__ u32 get(u32 n) {            // n <= 24; past the end: zeros (caught by `pos > end`)
    if (pos >= end) { pos += n; return 0; }
    const u64 byte = pos >> 3;
    const u32 sh = (u32)(pos & 7);
    const u32 w = ((u32)p[byte] << 24) | ((u32)p[byte + 1] << 16) | ((u32)p[byte + 2] << 8) | (u32)p[byte + 3];
    pos += n;
    return (w << sh) >> (32 - n);
  }
  __device__ __forceinline__ u32 bit() {
    if (pos >= end) { pos++; return 0; }
    const u32 b = (p[pos >> 3] >> (7 - (pos & 7))) & 1u;
    pos++;
    return b;
  }
};

struct VTables {
  i32 limit[6][22];
  i32 base[6][23];
  u16 perm[6][B2_MAX_ALPHA];
  u32 minlen[6];
  u8 len[6][B2_MAX_ALPHA];
  u8 unseq[256];
  u8 yy[256];
  u32 cf[257];
};

// One warp per candidate.  Lane 0 reads the bits (a serial chain); the lanes share the move-to-front shifts and
// the fills of the zero runs.
__global__ void __launch_bounds__(32)
k_v_decode(const u8 *__restrict__ stream, u64 n_bits, B2VBlock *__restrict__ blocks, u32 n_blocks, u32 max_n,
           u32 *__restrict__ link_all, u8 *__restrict__ lcol_all, u8 *__restrict__ rle_all, u8 *__restrict__ sel_all,
           const u32 *__restrict__ crc_tab) {
  __shared__ VTables T;
  const u32 bi = blockIdx.x;
  if (bi >= n_blocks) return;
  const u32 lane = threadIdx.x;
  B2VBlock &B = blocks[bi];
  if (B.status == 0xFFu) return;                 // a stream-footer candidate, not a block
  u32 *link = link_all + (size_t)bi * (max_n + 32);
  u8 *lcol = lcol_all + (size_t)bi * (max_n + 32);
  u8 *rle = rle_all + (size_t)bi * (max_n + 32);
  u8 *sel = sel_all + (size_t)bi * 18016;
  VBits br{stream, B.start_bit + 48, n_bits};
  u32 status = 0;
  u32 n = 0;
  u32 n_inuse = 0, n_groups = 0, n_sel = 0, orig_ptr = 0;
  if (lane == 0) {
    B.stored_crc = (br.get(16) << 16) | br.get(16);                      // (:581)
    if (br.bit()) status = 2;                                             // randomised block: not supported
    orig_ptr = br.get(24);
    // ---- mapping table (:94-119)
    if (!status) {
      const u32 used16 = br.get(16);
      for (u32 i = 0; i < 16; i++) {
        if ((used16 >> (15 - i)) & 1u) {
          const u32 m = br.get(16);
          for (u32 k = 0; k < 16; k++) if ((m >> (15 - k)) & 1u) T.unseq[n_inuse++] = (u8)(16 * i + k);
        }
      }
    }                                                                     // (an empty block has no byte in use: only EOB follows)
    // ---- selectors (:121-180)
    if (!status) {
      n_groups = br.get(3);
      n_sel = br.get(15);
      if (n_groups < 2 || n_groups > 6 || n_sel < 1 || n_sel > 18002) status = 4;
    }
    if (!status) {
      u8 pos[6];
      for (u32 v = 0; v < 6; v++) pos[v] = (u8)v;
      for (u32 i = 0; i < n_sel && !status; i++) {
        u32 j = 0;
        while (br.bit()) { j++; if (j >= n_groups) { status = 5; break; } }
        if (status) break;
        const u8 tmp = pos[j];
        for (; j > 0; j--) pos[j] = pos[j - 1];
        pos[0] = tmp;
        sel[i] = tmp;
      }
    }
    // ---- code lengths (:233-261) and canonical decoding tables (:182-222, :263-291)
    const u32 alpha = n_inuse + 2;
    if (!status) {
      for (u32 t = 0; t < n_groups && !status; t++) {
        i32 curr = (i32)br.get(5);
        for (u32 i = 0; i < alpha; i++) {
          for (;;) {
            if (curr < 1 || curr > 20) { status = 6; break; }
            if (!br.bit()) break;
            curr += br.bit() ? -1 : 1;
          }
          if (status) break;
          T.len[t][i] = (u8)curr;
        }
      }
    }
    if (!status) {
      for (u32 t = 0; t < n_groups; t++) {
        u32 mn = 32, mx = 0;
        for (u32 i = 0; i < alpha; i++) { const u32 l = T.len[t][i]; mn = min(mn, l); mx = max(mx, l); }
        T.minlen[t] = mn;
        u32 pp = 0;
        for (u32 l = mn; l <= mx; l++) for (u32 i = 0; i < alpha; i++) if (T.len[t][i] == l) T.perm[t][pp++] = (u16)i;
        for (u32 l = 0; l < 23; l++) T.base[t][l] = 0;
        for (u32 i = 0; i < alpha; i++) T.base[t][T.len[t][i] + 1]++;
        for (u32 l = 1; l < 23; l++) T.base[t][l] += T.base[t][l - 1];
        for (u32 l = 0; l < 22; l++) T.limit[t][l] = 0;
        i32 vec = 0;
        for (u32 l = mn; l <= mx; l++) { vec += T.base[t][l + 1] - T.base[t][l]; T.limit[t][l] = vec - 1; vec <<= 1; }
        for (u32 l = mn + 1; l <= mx; l++) T.base[t][l] = ((T.limit[t][l - 1] + 1) << 1) - T.base[t][l];
        for (u32 l = mx + 1; l < 22; l++) T.limit[t][l] = 0x7FFFFFFF;     // longer than the longest code: stop (and fail below)
      }
    }
  }
  for (u32 i = lane; i < 256; i += 32) { T.yy[i] = (u8)i; T.cf[i] = 0; }
  status = __shfl_sync(0xffffffffu, status, 0);
  n_inuse = __shfl_sync(0xffffffffu, n_inuse, 0);
  __syncwarp();
  const u32 alpha = n_inuse + 2, eob = n_inuse + 1;
  // ---- Huffman symbols, inverse move-to-front, zero runs (:293-468)
  if (!status) {
    u32 g = 0, left = 0, t = 0;                 // lane 0: selector cursor
    u32 es = 0, N = 1;                          // pending zero run (bijective base 2, low digit first)
    for (;;) {
      u32 sym = eob;
      if (lane == 0) {
        if (left == 0) {
          if (g >= n_sel) status = 7;
          else { t = sel[g++]; left = B2_GROUP_SIZE; }
        }
        if (!status) {
          left--;
          u32 zn = T.minlen[t];
          i32 zvec = (i32)br.get(zn);
          while (zvec > T.limit[t][zn]) {
            zn++;
            if (zn > 20) { status = 8; break; }
            zvec = (zvec << 1) | (i32)br.bit();
          }
          if (!status) {
            const i32 k = zvec - T.base[t][zn];
            if (k < 0 || k >= (i32)alpha) status = 9; else sym = T.perm[t][k];
          }
        }
        if (!status && sym != eob && sym <= 1) {
          es += N << sym;                        // run_a adds N, run_b adds 2 N
          N <<= 1;
          if (N > (1u << 21)) status = 10;
        }
      }
      sym = __shfl_sync(0xffffffffu, sym, 0);
      status = __shfl_sync(0xffffffffu, status, 0);
      if (status) break;
      if (sym != eob && sym <= 1) continue;
      const u32 run = __shfl_sync(0xffffffffu, es, 0);
      if (run) {                                 // the pending run ends here: `run` copies of the front byte
        if (n_inuse == 0) { status = 3; break; }
        const u8 b = T.unseq[T.yy[0]];
        if (n + run > max_n) { status = 11; break; }
        for (u32 k = lane; k < run; k += 32) lcol[n + k] = b;
        if (lane == 0) { T.cf[b] += run; es = 0; N = 1; }
        n += run;
      }
      if (sym == eob) break;
      const u32 nn = sym - 1;
      const u8 v = T.yy[nn];
      u8 keep[8];
#pragma unroll
      for (int q = 0; q < 8; q++) { const u32 j = lane + 1 + 32 * q; keep[q] = j <= nn ? T.yy[j - 1] : 0; }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; q++) { const u32 j = lane + 1 + 32 * q; if (j <= nn) T.yy[j] = keep[q]; }
      if (n >= max_n) { status = 11; break; }
      if (lane == 0) { T.yy[0] = v; const u8 b = T.unseq[v]; T.cf[b]++; lcol[n] = b; }
      n++;
      __syncwarp();
    }
    if (!status && __shfl_sync(0xffffffffu, (u32)(br.pos > n_bits), 0)) status = 12;
  }
  __syncwarp();
  if (lane != 0) return;
  B.end_bit = br.pos;
  B.n_rle = n;
  B.orig_ptr = orig_ptr;
  u64 raw_len = 0;
  u32 crc = 0xFFFFFFFFu;
  if (!status && n > 0) {
    if (orig_ptr >= n) status = 13;
    else {
      // ---- inverse BWT: counting pass (BWT_Detransform :470-487), then the pointer walk inside RLE_1 (:489-543)
      u32 run = 0;
      for (u32 i = 0; i < 256; i++) { const u32 c = T.cf[i]; T.cf[i] = run; run += c; }
      for (u32 p = 0; p < n; p++) { const u32 b = lcol[p]; link[T.cf[b]++] = p; }
      u32 idx = link[orig_ptr];
      u32 rl = 0, prev = 256;
      for (u32 k = 0; k < n; k++) {
        const u32 b = lcol[idx];
        idx = link[idx];
        rle[k] = (u8)b;
        if (rl == 4) {                                       // the byte after four equal ones is a count
          for (u32 q = 0; q < b; q++) crc = crc_tab[(crc >> 24) ^ prev] ^ (crc << 8);
          raw_len += b;
          rl = 0; prev = 256;
        } else {
          if (b == prev) rl++; else { rl = 1; prev = b; }
          crc = crc_tab[(crc >> 24) ^ b] ^ (crc << 8);
          raw_len++;
        }
      }
    }
  }
  B.computed_crc = ~crc;
  B.raw_len = raw_len;
  B.status = status;
}

// ---- the raw bytes against the expected input: one warp per block of the chain --------------------------------
__global__ void __launch_bounds__(32)
k_v_compare(const B2VBlock *__restrict__ blocks, const u32 *__restrict__ chain, u32 n_chain, u32 max_n, const u8 *__restrict__ rle_all,
            const u8 *__restrict__ expect, u64 expect_n, unsigned long long *__restrict__ first_bad) {
  const u32 ci = blockIdx.x;
  if (ci >= n_chain || threadIdx.x != 0) return;
  const u32 bi = chain[ci];
  const B2VBlock &B = blocks[bi];
  const u8 *rle = rle_all + (size_t)bi * (max_n + 32);
  u64 o = B.raw_off;
  u32 rl = 0, prev = 256;
  unsigned long long bad = ~0ull;
  for (u32 k = 0; k < B.n_rle && bad == ~0ull; k++) {
    const u32 b = rle[k];
    if (rl == 4) {
      for (u32 q = 0; q < b; q++, o++) if (o >= expect_n || expect[o] != prev) { bad = o; break; }
      rl = 0; prev = 256;
    } else {
      if (b == prev) rl++; else { rl = 1; prev = b; }
      if (o >= expect_n || expect[o] != b) bad = o;
      o++;
    }
  }
  if (bad != ~0ull) atomicMin(first_bad, bad);
}

int b2k_verify_find(cudaStream_t st, const u8 *d_stream, u64 n_bytes, u64 *d_cand, u32 *d_n_cand, u32 cap) {
  const u64 n_words = (n_bytes + 3) / 4;
  if (n_words == 0) return 0;
  k_v_find<<<(u32)((n_words + 255) / 256), 256, 0, st>>>(reinterpret_cast<const u32 *>(d_stream), n_words, n_bytes * 8, d_cand, d_n_cand, cap);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int b2k_verify_decode(cudaStream_t st, const u8 *d_stream, u64 n_bytes, B2VBlock *d_blocks, u32 n_blocks, u32 max_n, u32 *d_link, u8 *d_lcol,
                      u8 *d_rle, u8 *d_sel, const u32 *d_crc_tab) {
  if (n_blocks == 0) return 0;
  k_v_decode<<<n_blocks, 32, 0, st>>>(d_stream, n_bytes * 8, d_blocks, n_blocks, max_n, d_link, d_lcol, d_rle, d_sel, d_crc_tab);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int b2k_verify_compare(cudaStream_t st, const B2VBlock *d_blocks, const u32 *d_chain, u32 n_chain, u32 max_n, const u8 *d_rle,
                       const u8 *d_expect, u64 expect_n, unsigned long long *d_first_bad) {
  if (n_chain == 0) return 0;
  k_v_compare<<<n_chain, 32, 0, st>>>(d_blocks, d_chain, n_chain, max_n, d_rle, d_expect, expect_n, d_first_bad);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}
