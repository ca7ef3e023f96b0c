// Stage A5: RLE_1 + block CRC + byte usage map, one CTA per block, tile loop with carries.
//
// Reference: zip_lib/bzip2-encoding.adb:165-210 (RLE_1), zip_lib/bzip2.adb:34-118 (CRC).
// Runs of 4..259 equal bytes become 4 bytes + a count byte (run-4); count bytes are also
// marked in `in_use` (:174).  Output offsets are an exclusive prefix sum of per-position
// emissions; the CRC (MSB-first, poly 0x04C11DB7) of a slice is combined
// from per-thread partial CRCs by GF(2) polynomial arithmetic (R6).
#include "b2_common.cuh"
#include "b2_kernels.h"

#define RLE_THREADS 1024
#define RLE_BYTES 16
#define RLE_TILE (RLE_THREADS * RLE_BYTES)
__global__ void __launch_bounds__(RLE_THREADS, 1)
k_rle1(const u8 *__restrict__ in, B2Job *jobs, u8 *__restrict__ text, const B2CrcTables *__restrict__ ct) {
  __shared__ u32 crc_tab[256];
  __shared__ i32 sm_scan[40];
  __shared__ u32 sm_scan_u[40];
  __shared__ u8 stage[RLE_TILE + RLE_TILE / 4 + 64];
  __shared__ u32 used[8];
  __shared__ u32 sm_crc[32];
  const u32 tid = threadIdx.x;
  B2Job &job = jobs[blockIdx.x];
  const u8 *src = in + job.raw_off;
  const u32 len = job.raw_len;
  u8 *dst = text + job.pos_off;
  if (tid < 256) crc_tab[tid] = ct->byte_tab[tid];
  if (tid < 8) used[tid] = 0;
  __syncthreads();
  u32 out_base = 0;
  i32 carry_r = -1;
  u32 crc_run = 0;       // crc0 of everything so far (thread 0 only)
  u32 ntiles = 0;
  u32 lu[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // per-thread 256-bit usage map (static indexing only)
  for (u32 t0 = 0; t0 < len; t0 += RLE_TILE, ntiles++) {
    const u32 a = t0 + tid * RLE_BYTES;
    u8 b[RLE_BYTES];
    u8 prev = 0, next = 0;
#pragma unroll
    for (int k = 0; k < RLE_BYTES; k++) b[k] = (a + k < len) ? src[a + k] : 0;
    if (a > 0 && a < len) prev = src[a - 1];
    if (a + RLE_BYTES < len) next = src[a + RLE_BYTES];
    // per-thread crc0 over its bytes (zero padded on the right in the last tile)
    u32 c = 0;
#pragma unroll
    for (int k = 0; k < RLE_BYTES; k++) c = crc_tab[(c >> 24) ^ b[k]] ^ (c << 8);
    c = gf_mulmod(c, ct->xp_thread[tid]);
#pragma unroll
    for (int o = 16; o; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
    // phase 1: last change position
    i32 lc = -1;
    {
      u8 pv = prev;
#pragma unroll
      for (int k = 0; k < RLE_BYTES; k++) {
        u32 p = a + k;
        if (p < len && (p == 0 || b[k] != pv)) lc = (i32)p;
        pv = b[k];
      }
    }
    i32 tot_lc;
    i32 r_in = block_excl_max(lc, -1, sm_scan, &tot_lc);
    r_in = max(r_in, carry_r);
    if (lane_id() == 0) sm_crc[warp_id()] = c;
    // phase 2: emissions.  em[k] bit0 = literal, bit1 = count byte after it
    u32 m_of[RLE_BYTES];
    u32 emits = 0;
    u32 local = 0;
    {
      i32 r = r_in;
      u8 pv = prev;
#pragma unroll
      for (int k = 0; k < RLE_BYTES; k++) {
        u32 p = a + k;
        u32 e = 0, m = 0;
        if (p < len) {
          if (p == 0 || b[k] != pv) r = (i32)p;
          m = (u32)((i32)p - r) % 259u;
          u8 nx = (k + 1 < RLE_BYTES) ? b[k + 1] : next;
          bool last = (p + 1 == len) || (nx != b[k]) || (m == 258);
          if (m < 4) e |= 1;
          if (last && m >= 3) e |= 2;
        }
        m_of[k] = m;
        emits |= e << (2 * k);
        local += (e & 1) + (e >> 1);
        pv = b[k];
      }
    }
    u32 tile_total;
    u32 base = block_excl_add(local, sm_scan_u, &tile_total);
    // stage this tile's output in shared memory, then copy out coalesced
    {
      u32 o = base;
#pragma unroll
      for (int k = 0; k < RLE_BYTES; k++) {
        u32 e = (emits >> (2 * k)) & 3;
        if (e & 1) {
          stage[o++] = b[k];
#pragma unroll
          for (int w = 0; w < 8; w++) lu[w] |= (w == (b[k] >> 5)) ? (1u << (b[k] & 31)) : 0u;
        }
        if (e & 2) {
          u8 cb = (u8)(m_of[k] - 3);
          stage[o++] = cb;
#pragma unroll
          for (int w = 0; w < 8; w++) lu[w] |= (w == (cb >> 5)) ? (1u << (cb & 31)) : 0u;
        }
      }
    }
    __syncthreads();
    for (u32 i = tid; i < tile_total; i += RLE_THREADS) dst[out_base + i] = stage[i];
    if (tid == 0) {
      u32 tc = 0;
      for (int w = 0; w < RLE_THREADS / 32; w++) tc ^= sm_crc[w];
      crc_run = gf_mulmod(crc_run, ct->x_tile) ^ tc;
    }
    out_base += tile_total;
    carry_r = max(carry_r, tot_lc);
    __syncthreads();
  }
#pragma unroll
  for (int w = 0; w < 8; w++) {
    u32 v = __reduce_or_sync(0xffffffffu, lu[w]);
    if (lane_id() == 0 && v) atomicOr(&used[w], v);
  }
  __syncthreads();
  if (tid == 0) {
    // undo the right zero padding of the last tile, then add the init/final terms
    u32 padded = ntiles * RLE_TILE;
    u32 trail = padded - len;
    u32 c0 = crc_run;
    if (trail) { c0 = gf_mulmod(c0, ct->xinv_a[trail >> 4]); c0 = gf_mulmod(c0, ct->xinv_b[trail & 15]); }
    u32 xl = 1;                                     // x^(8*len)
    for (int k = 0; k < 32; k++) if ((len >> k) & 1u) xl = gf_mulmod(xl, ct->pw2[k]);
    u32 crc = c0 ^ gf_mulmod(0xFFFFFFFFu, xl);      // Init = 0xFFFFFFFF (bzip2.adb:110-113)
    job.crc = ~crc;                                 // Final (bzip2.adb:115-118)
    job.n = out_base;
    u32 nu = 0;
    for (int i = 0; i < 8; i++) { job.in_use[i] = used[i]; nu += __popc(used[i]); }
    job.n_used = nu;
  }
}

int b2k_rle1(cudaStream_t st, const u8 *d_in, B2Job *d_jobs, u32 n_jobs, u8 *d_text, const B2CrcTables *d_ct) {
  if (n_jobs == 0) return 0;
  k_rle1<<<n_jobs, RLE_THREADS, 0, st>>>(d_in, d_jobs, d_text, d_ct);
  B2_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// Host-side construction of the GF(2) tables (pure constants of the format).
static u32 h_mulmod(u32 a, u32 b) {
  u32 r = 0;
  for (int i = 31; i >= 0; i--) {
    r = (r << 1) ^ ((r & 0x80000000u) ? CRC_POLY : 0u);
    if ((b >> i) & 1u) r ^= a;
  }
  return r;
}
static u32 h_powx(u64 e_bits) {  // x^e_bits mod P
  u32 r = 1, base = 2;           // x
  while (e_bits) { if (e_bits & 1) r = h_mulmod(r, base); base = h_mulmod(base, base); e_bits >>= 1; }
  return r;
}
void b2k_make_crc_tables(B2CrcTables *t) {
  for (u32 i = 0; i < 256; i++) {
    u32 c = i << 24;
    for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ CRC_POLY : (c << 1);
    t->byte_tab[i] = c;
  }
  for (u32 i = 0; i < RLE_THREADS; i++) t->xp_thread[i] = h_powx((u64)8 * RLE_BYTES * (RLE_THREADS - 1 - i));
  t->x_tile = h_powx((u64)8 * RLE_TILE);
  // x^-1 = (P - 1) / x = x^31 + (0x04C11DB6 >> 1)
  const u32 xinv = 0x80000000u | (0x04C11DB6u >> 1);
  u32 xinv8 = 1;
  for (int i = 0; i < 8; i++) xinv8 = h_mulmod(xinv8, xinv);
  t->xinv_b[0] = 1;
  for (int i = 1; i < 16; i++) t->xinv_b[i] = h_mulmod(t->xinv_b[i - 1], xinv8);
  u32 xinv128 = h_mulmod(t->xinv_b[15], xinv8);
  t->xinv_a[0] = 1;
  for (int i = 1; i < 1024; i++) t->xinv_a[i] = h_mulmod(t->xinv_a[i - 1], xinv128);
  for (int k = 0; k < 32; k++) t->pw2[k] = h_powx((u64)8 << k);
}
