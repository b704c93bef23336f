/* p3_k1.cuh -- device code of the Huffman stage (K1), shared by k_huffman (p3_kernels.cu) and the
 * persistent fused kernel (p3_fused.cu).  Reference: Read_Main_L3 1376-1435, Read_Huffman 2051-2115,
 * Huffman_Decode 1593-1643, Get_Main_Data 1096-1122 of /root/reference/pdmp3.c. */
#pragma once
#include "p3_device.cuh"

/* MSB-first bit reader over the big-endian words in shared memory: two consecutive words in registers
 * plus a bit offset, so a 32-bit look-ahead is ONE funnel shift and advancing is branch-free (the next
 * word is loaded unconditionally, off the critical path).  The reference does a byte access per BIT
 * (pdmp3.c:1489-1527). */
struct k1_bits {
  const uint32_t *sw; uint32_t cur, nxt, off, widx;       /* cur = sw[widx], nxt = sw[widx+1], off in 0..31 */
  __device__ __forceinline__ void init(const uint32_t *s, uint32_t bitpos)
  {
    sw = s; widx = bitpos >> 5; off = bitpos & 31; cur = s[widx]; nxt = s[widx + 1];
  }
  __device__ __forceinline__ uint32_t pos() const { return widx * 32 + off; }
  __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(nxt, cur, off); }   /* next 32 bits */
  __device__ __forceinline__ void skip(uint32_t n)                                              /* n <= 32 */
  {
    const uint32_t w = sw[widx + 2];
    off += n;
    const bool adv = off >= 32;
    cur = adv ? nxt : cur; nxt = adv ? w : nxt;
    widx += adv ? 1u : 0u; off &= 31u;
  }
};

/* one Huffman-coded pair (pdmp3.c:1593-1643); tb = base | pbits<<16 | linbits<<24 */
__device__ __forceinline__ uint32_t k1_pair(k1_bits &bb, const uint16_t *lut, uint32_t tb)
{
  const uint32_t base = tb & 0xffffu, linbits = (tb >> 24) & 31u;
  const uint32_t w0 = bb.peek();
  uint32_t cw = (tb >> 16) & 31u, used = 0;
  uint32_t e = lut[base + (w0 >> (32 - cw))];
  while (e & 0x8000u) {                                 /* next LUT level (codes longer than 8 bits) */
    used += cw; cw = (e >> 10) & 7;
    e = lut[base + (e & 1023u) + ((w0 << used) >> (32 - cw))];
  }
  used += (e >> 8) & 31;
  int x = (e >> 4) & 15, y = e & 15;
  /* bits after the code word: x linbits, x sign, y linbits, y sign (pdmp3.c:1637-1640).  If code + worst-case
   * escapes cannot fit the 32-bit window (only tables with >= 8 linbits), re-read after the code word. */
  uint32_t w = w0 << used, total = used;
  if (used + 2 * linbits + 2 > 32) { bb.skip(used); w = bb.peek(); total = 0; }
  const uint32_t ex = x == 15 ? linbits : 0u;
  x += (int)((w >> 1) >> (31 - ex)); w <<= ex;
  const uint32_t sx = x != 0;
  x = (sx & (w >> 31)) ? -x : x; w <<= sx;
  const uint32_t ey = y == 15 ? linbits : 0u;
  y += (int)((w >> 1) >> (31 - ey)); w <<= ey;
  const uint32_t sy = y != 0;
  y = (sy & (w >> 31)) ? -y : y;
  bb.skip(total + ex + sx + ey + sy);
  return (uint32_t)(x & 0xffff) | ((uint32_t)y << 16);
}

/* Per-thread output staging: 8 words (16 spectral values = half a 32-byte sector) are collected in
 * shared memory, transposed [word][thread] so that neither the word writes nor the flush conflict,
 * and leave as one 16-byte store; the two halves of a sector meet in L2, so every DRAM sector of the
 * spectra is written once. */
struct k1_out {
  uint32_t *ring;                 /* &ring[0][tid] */
  uint32_t stride;                /* threads sharing the staging area */
  uint4 *dst;                     /* this granule-channel's 1152-byte row */
  uint32_t pw;                    /* words emitted so far */
  __device__ __forceinline__ void flush(uint32_t q)          /* q: index of the 16-byte quarter-row chunk */
  {
    uint4 a;
    a.x = ring[0]; a.y = ring[stride]; a.z = ring[2 * stride]; a.w = ring[3 * stride];
    dst[q] = a;
  }
  __device__ __forceinline__ void put(uint32_t word)
  {
    ring[(pw & 3u) * stride] = word;
    if ((pw & 3u) == 3u) flush(pw >> 2);
    pw++;
  }
};


/* Gather the main data of frames [F0,F1) plus the up-to-512 reservoir bytes in front of them into
 * shared memory as big-endian words: smem byte 512 <-> logical main-data byte frames[F0].main_pos.
 * `sw` must be zeroed; ends with __syncthreads().  s_fb: one int64 of shared scratch. */
__device__ __forceinline__ void k1_gather(const uint8_t *__restrict__ raw, const p3_frame *__restrict__ frames,
                                          const uint8_t *__restrict__ tail, int64_t F0, int64_t F1, uint32_t *sw, int64_t *s_fb_p)
{
  const uint64_t base0 = frames[F0].main_pos;
#define s_fb (*s_fb_p)
  if (threadIdx.x == 0) {                                /* earliest frame whose data reaches into the window */
    int64_t fs = F0; uint32_t acc = 0;
    while (fs > 0 && acc < 512) { fs--; acc += frames[fs].main_size; }
    s_fb = fs;
  }
  __syncthreads();
  uint8_t *sb8 = reinterpret_cast<uint8_t *>(sw);
  {
    /* bytes that precede frame 0 of the batch come from the context's tail buffer */
    const int64_t fb = s_fb;
    const int64_t lo = (int64_t)base0 - 512;             /* logical position of smem byte 0 */
    if (fb == 0) {
      int64_t first = (int64_t)frames[0].main_pos;       /* logical start of the batch */
      for (int64_t L = lo + threadIdx.x; L < first; L += blockDim.x) {
        int64_t back = first - L;                        /* 1..512 */
        if (back <= 512) sb8[(uint32_t)(L - lo) ^ 3u] = tail[512 - back];
      }
    }
    /* one warp per source frame: whole destination words via two aligned loads + funnel shift, the
     * ragged ends byte by byte (a word can straddle two frames' data) */
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int64_t fs = fb + warp; fs < F1; fs += nwarp) {
      const uint8_t *src = raw + frames[fs].main_off;
      const int64_t d0 = (int64_t)frames[fs].main_pos - lo;   /* smem byte of the frame's first data byte */
      const int32_t n = frames[fs].main_size;
      const int32_t b0 = d0 < 0 ? (int32_t)(-d0) : 0;    /* first byte that lands inside the window */
      if (b0 >= n) continue;
      const int32_t lead = (int32_t)((4 - ((d0 + b0) & 3)) & 3);          /* bytes up to the next word boundary */
      const int32_t wb = b0 + lead;                      /* first byte of the first whole word */
      const int32_t nw = wb < n ? (n - wb) >> 2 : 0;     /* whole words */
      for (int32_t b = b0 + (int32_t)lane; b < min(wb, n); b += 32) sb8[(uint32_t)(d0 + b) ^ 3u] = src[b];
      for (int32_t b = wb + 4 * nw + (int32_t)lane; b < n; b += 32) sb8[(uint32_t)(d0 + b) ^ 3u] = src[b];
      const uintptr_t sa = (uintptr_t)(src + wb);
      const uint32_t *al = reinterpret_cast<const uint32_t *>(sa & ~(uintptr_t)3);
      const uint32_t sh = (uint32_t)(sa & 3) * 8;
      uint32_t *dw = sw + ((d0 + wb) >> 2);
      for (int32_t k = lane; k < nw; k += 32) {
        const uint32_t v = __funnelshift_r(__ldcs(al + k), __ldcs(al + k + 1), sh);   /* 4 stream bytes, first byte in bits 0-7 */
        dw[k] = __byte_perm(v, 0, 0x0123);               /* first byte to the MSB */
      }
    }
  }
  __syncthreads();

#undef s_fb
}

/* Scalefactors + Huffman of ONE granule-channel by ONE thread, spectra through `ob`, scalefactor bytes
 * to scf[64] (zeroed by the caller).  Returns count1. */
__device__ __forceinline__ uint32_t k1_decode_gc(const uint32_t *sw, const uint16_t *lut, const p3_tables *__restrict__ T,
                                                 const p3_gc *__restrict__ gcs, const p3_frame &fr, const p3_gc &g, int64_t f,
                                                 uint32_t gr, uint32_t ch, uint64_t base0, k1_out &ob, uint8_t *scf)
{
  uint32_t c1 = 0;
  {
    const bool ok = ch < fr.nch && !(fr.flags & (P3_FRAME_NODATA | P3_FRAME_BAD));
    const uint32_t p23l = ok ? P3_GC_P23L(g) : 0u;
    const uint32_t fstart = (uint32_t)((int64_t)fr.main_pos - fr.main_begin - ((int64_t)base0 - 512)) * 8u;
    uint32_t pos = fstart + P3_GC_START(g);
    const uint32_t part2_start = pos;
    const bool is_short = P3_GC_WINSW(g) && P3_GC_BTYPE(g) == 2;
    /* ---- part 2: scalefactors (pdmp3.c:1382-1435); a zero-length part carries none ---- */
    if (ok) {
      const uint32_t slen1 = p23l ? T->slen[P3_GC_SFCOMP(g)][0] : 0u, slen2 = p23l ? T->slen[P3_GC_SFCOMP(g)][1] : 0u;
      if (is_short) {
        uint32_t first = 0;
        if (P3_GC_MIXED(g)) { for (int sfb = 0; sfb < 8; sfb++) scf[sfb] = (uint8_t)p3_getbits(sw, pos, slen1); first = 3; }
        for (uint32_t sfb = first; sfb < 12; sfb++)
          for (int win = 0; win < 3; win++) scf[P3_SCF_S_OFF + 3 * sfb + win] = (uint8_t)p3_getbits(sw, pos, sfb < 6 ? slen1 : slen2);
      } else {
        const uint32_t scfsi = gr == 1 ? (fr.scfsi >> (4 * ch)) & 15u : 0u;
        /* granule 1 may reuse granule 0's scalefactors (scfsi): re-read them from granule 0's part 2 */
        const p3_gc g0 = gcs[4 * f + ch];
        const bool g0_ok = !(P3_GC_WINSW(g0) && P3_GC_BTYPE(g0) == 2) && P3_GC_P23L(g0) != 0;
        const uint32_t a1 = T->slen[P3_GC_SFCOMP(g0)][0], a2 = T->slen[P3_GC_SFCOMP(g0)][1];
        const uint32_t g0pos = fstart + P3_GC_START(g0);
        #pragma unroll
        for (int band = 0; band < 4; band++) {
          const int lo = band == 0 ? 0 : band == 1 ? 6 : band == 2 ? 11 : 16, hi = band == 0 ? 6 : band == 1 ? 11 : band == 2 ? 16 : 21;
          if ((scfsi >> band) & 1) {
            uint32_t p = g0pos + (band == 0 ? 0u : band == 1 ? 6 * a1 : band == 2 ? 11 * a1 : 11 * a1 + 5 * a2), nb = band < 2 ? a1 : a2;
            for (int sfb = lo; sfb < hi; sfb++) scf[sfb] = g0_ok ? (uint8_t)p3_getbits(sw, p, nb) : 0;
          } else {
            for (int sfb = lo; sfb < hi; sfb++) scf[sfb] = (uint8_t)p3_getbits(sw, pos, band < 2 ? slen1 : slen2);
          }
        }
      }
    }
    if (p23l) {
      /* ---- part 3 (pdmp3.c:2063-2113) ---- */
      const uint32_t bit_pos_end = part2_start + p23l - 1;
      uint32_t r1s, r2s;
      if (is_short) { r1s = 36; r2s = 576; }
      else { r1s = T->sfb_l[fr.sfreq][P3_GC_REG0(g) + 1]; r2s = T->sfb_l[fr.sfreq][P3_GC_REG0(g) + P3_GC_REG1(g) + 2]; }
      const uint32_t bv2 = 2 * P3_GC_BIGV(g);
      uint32_t tbs[3];
      #pragma unroll
      for (int r = 0; r < 3; r++) {
        const uint32_t t = P3_GC_TSEL(g, r);
        const int book = T->table_book[t];
        tbs[r] = book < 0 ? 0xffffffffu : ((uint32_t)T->book_base[book] | (uint32_t)T->book_pbits[book] << 16 | (uint32_t)T->table_linbits[t] << 24);
      }
      /* one loop over all big_values pairs; the table changes at the region boundaries, so every lane of
       * the warp stays in the same loop whatever its region split */
      uint32_t tb = tbs[0];
      k1_bits bb; bb.init(sw, pos);
      for (uint32_t i = 0; i < bv2; i += 2) {
        if (i == r1s) tb = tbs[1];
        if (i == r2s) tb = tbs[2];
        ob.put(tb == 0xffffffffu ? 0u : k1_pair(bb, lut, tb));            /* empty tables: zeros, no bits (pdmp3.c:1599-1602) */
      }
      /* count1 quads (pdmp3.c:2091-2103) */
      uint32_t is_pos = bv2;
      const bool tabB = P3_GC_C1TAB(g);                  /* reference quirk Q1: table B = leaf 0011, no code bits */
      const uint32_t qbase = T->book_base[T->table_book[32]], qbits = T->book_pbits[T->table_book[32]];
      pos = bb.pos();
      while (is_pos <= 572 && pos <= bit_pos_end) {
        uint32_t w = bb.peek(), used = 0, leaf = 3;
        if (!tabB) { uint32_t e = lut[qbase + (w >> (32 - qbits))]; used = (e >> 8) & 31; leaf = e & 15; w <<= used; }
        int v = (leaf >> 3) & 1, ww = (leaf >> 2) & 1, x = (leaf >> 1) & 1, y = leaf & 1;
        if (v) { if (w >> 31) v = -1; w <<= 1; used++; }
        if (ww) { if (w >> 31) ww = -1; w <<= 1; used++; }
        if (x) { if (w >> 31) x = -1; w <<= 1; used++; }
        if (y) { if (w >> 31) y = -1; used++; }
        bb.skip(used); pos += used;
        ob.put((uint32_t)(v & 0xffff) | ((uint32_t)ww << 16));
        ob.put((uint32_t)(x & 0xffff) | ((uint32_t)y << 16));
        is_pos += 4;
      }
      if (pos > bit_pos_end + 1) is_pos = is_pos >= 4 ? is_pos - 4 : 0;    /* pdmp3.c:2105-2106 */
      c1 = is_pos;
    }
    /* ---- close the row: words [c1/2, 288) are the rzero region; 72 chunks of 4 words ---- */
    {
      const uint32_t pw_hi = ob.pw, pw_new = c1 >> 1, S = pw_hi & ~3u;
      uint32_t zs;                                        /* first chunk that is entirely zero */
      if (pw_new >= S) {
        if (S < 288) {
          for (uint32_t k = pw_new & 3u; k < 4; k++) ob.ring[k * ob.stride] = 0;
          ob.flush(S >> 2);
        }
        zs = (S >> 2) + 1;
      } else {                                            /* rolled back into a chunk that already left */
        uint32_t *dw = reinterpret_cast<uint32_t *>(ob.dst);
        for (uint32_t k = pw_new; k < S; k++) dw[k] = 0;
        zs = S >> 2;
      }
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (uint32_t q = zs; q < 72; q++) ob.dst[q] = z;
    }
  }
  return c1;
}
