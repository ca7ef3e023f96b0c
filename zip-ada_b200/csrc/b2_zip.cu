// Archive side of the path: the Zip CRC-32 of every entry and the assembly of
// the archive image on the device.
//
// Reference: zip_lib/zip-crc_crypto.adb:31-61 (table CRC-32, reflected polynomial 0xEDB88320, Init =
// 0xFFFFFFFF, Final = not), computed by the reference one byte at a time inside the Read_Byte callback
// (zip-compress-bzip2_e.adb:70-98); zip-create.adb:194-297 (entry = local header, name, payload).
//
// A reflected CRC is the MSB-first CRC of the bit-reversed bytes, bit-reversed: the GF(2) machinery of
// the block CRC (b2_rle1.cu) is reused.  With a zero start value leading zero bytes do not change the
// remainder, so the tiles of an entry are aligned to its END and the first tile is padded on the left:
// every thread then has a fixed power of x, whatever the entry size.
#include "b2_common.cuh"
#include "b2_kernels.h"

#define ZC_THREADS 1024
#define ZC_BYTES 64
#define ZC_TILE (ZC_THREADS * ZC_BYTES)

__global__ void __launch_bounds__(ZC_THREADS)
k_zipcrc_tiles(const u8 *__restrict__ in, const B2ZipTile *__restrict__ tiles, const B2ZipCrcTables *__restrict__ zt,
               u32 *__restrict__ partial) {
  __shared__ u32 tab[256];
  __shared__ u32 red[32];
  const B2ZipTile tl = tiles[blockIdx.x];
  const u32 tid = threadIdx.x;
  if (tid < 256) tab[tid] = zt->byte_tab[tid];
  __syncthreads();
  // my 64 bytes end at tl.end - 64 * (1023 - tid); bytes before tl.begin count as zeros
  const i64 hi = (i64)tl.end - (i64)ZC_BYTES * (ZC_THREADS - 1 - tid);
  const i64 lo = hi - ZC_BYTES;
  u32 c = 0;
  if (hi > (i64)tl.begin) {
    const i64 from = lo > (i64)tl.begin ? lo : (i64)tl.begin;
    if (from == lo && ((lo & 15) == 0)) {
      const uint4 *p = reinterpret_cast<const uint4 *>(in + lo);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint4 v = p[q];
        const u32 wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const u32 r = __brev(wv[k]);           // byte j of the word, reversed, sits at bits 31-8j .. 24-8j
#pragma unroll
          for (int j = 0; j < 4; j++) c = tab[(c >> 24) ^ ((r >> (24 - 8 * j)) & 255u)] ^ (c << 8);
        }
      }
    } else {
      for (i64 i = from; i < hi; i++) c = tab[(c >> 24) ^ (__brev((u32)in[i]) >> 24)] ^ (c << 8);
    }
    c = gf_mulmod(c, zt->xp_thread[tid]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
  if (lane_id() == 0) red[warp_id()] = c;
  __syncthreads();
  if (tid < 32) {
    u32 v = red[tid];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
    if (tid == 0) partial[blockIdx.x] = v;
  }
}

__device__ __forceinline__ u32 zc_xpow_bytes(const B2ZipCrcTables *zt, u64 nbytes) {   // x^(8 * nbytes) mod P
  u32 r = 1;
  for (int k = 0; k < 48 && (nbytes >> k); k++) if ((nbytes >> k) & 1ull) r = r == 1 ? zt->pw2[k] : gf_mulmod(r, zt->pw2[k]);
  return r;
}

// One warp per entry: folds the tile remainders, applies Init and Final.
__global__ void __launch_bounds__(128)
k_zipcrc_fold(const B2ZipEntry *__restrict__ ents, u32 n, const u32 *__restrict__ partial,
              const B2ZipCrcTables *__restrict__ zt, u32 *__restrict__ crc_out) {
  const u32 e = blockIdx.x * 4 + warp_id();
  if (e >= n) return;
  const B2ZipEntry en = ents[e];
  const u32 l = lane_id();
  const u32 nt = en.n_tiles;
  const u32 per = (nt + 31) / 32;
  const u32 k0 = min(nt, l * per), k1 = min(nt, k0 + per);
  u32 acc = 0;
  const u32 xt = zt->pw2[16];                       // x^(8 * 65536) = one tile
  for (u32 k = k0; k < k1; k++) acc = gf_mulmod(acc, xt) ^ partial[en.tile0 + k];
  if (k1 < nt && acc) acc = gf_mulmod(acc, zc_xpow_bytes(zt, (u64)ZC_TILE * (nt - k1)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
  if (l == 0) {
    const u32 reg = acc ^ gf_mulmod(0xFFFFFFFFu, zc_xpow_bytes(zt, en.len));      // Init (zip-crc_crypto.adb:47-50)
    crc_out[e] = ~__brev(reg);                                                     // Final (:57-60)
  }
}

// Byte-granular gather: item = (source buffer, source offset, destination offset, length <= 64 KiB).
__global__ void __launch_bounds__(256)
k_zip_gather(const B2ZipCopy *__restrict__ items, const u8 *__restrict__ src0, const u8 *__restrict__ src1,
             u8 *__restrict__ dst) {
  const B2ZipCopy it = items[blockIdx.x];
  const u8 *s = (it.which ? src1 : src0) + it.src_off;
  u8 *d = dst + it.dst_off;
  const u32 n = it.len;
  const u32 head = min(n, (u32)((4 - ((uintptr_t)d & 3)) & 3));
  if (threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x];
  const u32 nw = (n - head) >> 2;
  const u8 *sb = s + head;
  u32 *dw = reinterpret_cast<u32 *>(d + head);
  const u32 sh = (u32)((uintptr_t)sb & 3);
  const u32 *sw = reinterpret_cast<const u32 *>(sb - sh);
  if (sh == 0) {
    for (u32 i = threadIdx.x; i < nw; i += blockDim.x) dw[i] = sw[i];
  } else {
    for (u32 i = threadIdx.x; i < nw; i += blockDim.x) dw[i] = __funnelshift_r(sw[i], sw[i + 1], 8 * sh);
  }
  const u32 done = head + 4 * nw;
  if (threadIdx.x < n - done) d[done + threadIdx.x] = s[done + threadIdx.x];
}

int b2k_zipcrc(cudaStream_t st, const u8 *d_in, const B2ZipTile *d_tiles, u32 n_tiles, const B2ZipEntry *d_ents, u32 n_entries,
               const B2ZipCrcTables *d_zt, u32 *d_partial, u32 *d_crc) {
  if (n_tiles) k_zipcrc_tiles<<<n_tiles, ZC_THREADS, 0, st>>>(d_in, d_tiles, d_zt, d_partial);
  if (n_entries) k_zipcrc_fold<<<(n_entries + 3) / 4, 128, 0, st>>>(d_ents, n_entries, d_partial, d_zt, d_crc);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int b2k_zip_gather(cudaStream_t st, const B2ZipCopy *d_items, u32 n, const u8 *d_src0, const u8 *d_src1, u8 *d_dst) {
  if (n == 0) return 0;
  k_zip_gather<<<n, 256, 0, st>>>(d_items, d_src0, d_src1, d_dst);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static u32 hz_mulmod(u32 a, u32 b) {
  u32 r = 0;
  for (int i = 31; i >= 0; i--) {
    r = (r << 1) ^ ((r & 0x80000000u) ? CRC_POLY : 0u);
    if ((b >> i) & 1u) r ^= a;
  }
  return r;
}

void b2k_make_zipcrc_tables(B2ZipCrcTables *t) {
  for (u32 i = 0; i < 256; i++) {
    u32 c = i << 24;
    for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ CRC_POLY : (c << 1);
    t->byte_tab[i] = c;
  }
  u32 p = 2;                                         // x
  for (int k = 0; k < 3; k++) p = hz_mulmod(p, p);   // x^8
  for (int k = 0; k < 48; k++) { t->pw2[k] = p; p = hz_mulmod(p, p); }    // x^(8 * 2^k)
  const u32 x64 = t->pw2[6];                         // x^(8 * 64): one thread's bytes
  u32 r = 1;
  for (int i = ZC_THREADS - 1; i >= 0; i--) { t->xp_thread[i] = r; r = hz_mulmod(r, x64); }
}
