/* p3_k1.cuh -- device code of the Huffman stage (K1, k_huffman in p3_kernels.cu).  Reference: Read_Main_L3 1376-1435, Read_Huffman 2051-2115,
 * Huffman_Decode 1593-1643, Get_Main_Data 1096-1122 of /root/reference/pdmp3.c. */
#pragma once
#include "p3_device.cuh"

/* K1_LUT_GLOBAL (experiment, tools/build_variant.sh lutg -DK1_LUT_GLOBAL): the Huffman LUT is read through L1 from global memory
 * instead of a per-CTA copy in shared memory -- 9 KB less per CTA, six resident CTAs per SM instead of five. */
#ifdef K1_LUT_GLOBAL
#define K1_LD(p) __ldg(p)
#else
#define K1_LD(p) (*(p))
#endif

/* MSB-first bit reader over the big-endian words of the CTA's window of the compact main-data stream in shared
 * memory: two consecutive words in registers plus a bit offset, so a 32-bit look-ahead is ONE funnel shift;
 * advancing adds to the offset and, when it crosses a word, shifts the pair and loads the next word under a
 * predicate -- no branch, one shared-memory load per 32 bits.  The reference does a byte access per BIT
 * (pdmp3.c:1489-1527). */
#ifdef K1_PAIR2
/* (experiment build: the same reader with ONE 32-bit shared-memory address instead of a generic pointer, so that a refill
 *  advances one register) */
struct k1_bits {
  uint32_t wa, base, hi, lo, off;                          /* wa: shared-memory address of the next word; hi:lo = the two words before it */
  __device__ __forceinline__ void init(const uint32_t *s, uint32_t bitpos)
  {
    const uint32_t i = bitpos >> 5;
    base = (uint32_t)__cvta_generic_to_shared(s); wa = base + (i + 2) * 4; off = bitpos & 31; hi = s[i]; lo = s[i + 1];
  }
  __device__ __forceinline__ uint32_t pos() const { return (wa - base - 8) * 8 + off; }
  __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(lo, hi, off); }
  __device__ __forceinline__ void skip(uint32_t n)
  {
    off += n;
    /* crossing a word: hi <- lo, lo <- next word, in place under one predicate (no copies; the in/out operands also keep the load
     * behind the loads that produced hi / lo, i.e. behind the mbarrier wait) */
    asm("{\n .reg .pred p;\n setp.ge.u32 p, %3, 32;\n @p mov.b32 %0, %1;\n @p ld.shared.b32 %1, [%2];\n @p add.u32 %2, %2, 4;\n}"
        : "+r"(hi), "+r"(lo), "+r"(wa) : "r"(off));
    off &= 31u;
  }
};
#else
struct k1_bits {
  const uint32_t *wp; uint32_t hi, lo, off;               /* hi:lo = words at wp[-2], wp[-1]; off in 0..31 */
  const uint32_t *base;
  __device__ __forceinline__ void init(const uint32_t *s, uint32_t bitpos)
  {
    base = s; wp = s + (bitpos >> 5) + 2; off = bitpos & 31; hi = wp[-2]; lo = wp[-1];
  }
  __device__ __forceinline__ uint32_t pos() const { return (uint32_t)(wp - base - 2) * 32 + off; }
  __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(lo, hi, off); }     /* next 32 bits */
  __device__ __forceinline__ void skip(uint32_t n)                                              /* n <= 32 */
  {
    off += n;
    if (off >= 32) { hi = lo; lo = *wp; wp++; }
    off &= 31u;
  }
};
#endif

/* One Huffman-coded pair (pdmp3.c:1593-1643) as a packed int16 pair.  tl: the code book's LUT, sh = 32 - width of
 * its first level.  Fast path (no escape): the LUT leaf gives x, y and the code length; the sign bits follow the
 * code word, x's first; a zero value has no sign bit, and negating zero is harmless, so no test is needed. */
__device__ __forceinline__ uint32_t k1_pair(k1_bits &bb, const uint16_t *tl, uint32_t sh, uint32_t linbits)
{
  const uint32_t w0 = bb.peek();
  uint32_t e = K1_LD(tl + (w0 >> sh)), used = 0;
  if (e & 0x8000u) {                                      /* codes longer than the first level: walk the next levels */
    uint32_t cw = 32 - sh;
    do { used += cw; cw = (e >> 10) & 7; e = K1_LD(tl + (e & 1023u) + ((w0 << used) >> (32 - cw))); } while (e & 0x8000u);
  }
  used += (e >> 8) & 31;
  int x = (e >> 4) & 15, y = e & 15;
  if ((e & 0x4000u) && linbits) {                         /* x or y is 15 and the table has linbits (pdmp3.c:1637-1640): rare */
    bb.skip(used);
    uint32_t w = bb.peek(), n = 0;
    if (x == 15) { x += (int)(w >> (32 - linbits)); w <<= linbits; n += linbits; }
    if (x) { if ((int)w < 0) x = -x; w <<= 1; n++; }
    if (y == 15) { y += (int)(w >> (32 - linbits)); w <<= linbits; n += linbits; }
    if (y) { if ((int)w < 0) y = -y; n++; }
    bb.skip(n);
    return __byte_perm((uint32_t)x, (uint32_t)y, 0x5410);
  }
  const uint32_t nx = min(x, 1), ny = min(y, 1);
  uint32_t w = w0 << used;
  int m = (int)w >> 31; x = (x ^ m) - m;                  /* the bit after the code word is x's sign (if x != 0) */
  w <<= nx;
  m = (int)w >> 31; y = (y ^ m) - m;
  bb.skip(used + nx + ny);
  return __byte_perm((uint32_t)x, (uint32_t)y, 0x5410);    /* (x & 0xffff) | y << 16 */
}

#ifdef K1_PAIR2
/* EXPERIMENT (tools/build_variant.sh pair2 -DK1_PAIR2): the same pair decode with ONE test in front of the fast path --
 * `slow` = 0x8000 (entry is a link to the next LUT level) | 0x4000 if the table has linbits (x or y is 15: escape) -- instead of
 * two reconvergence regions, and the sign / length arithmetic written as min / shift / xor-sub chains. */
__device__ __forceinline__ uint32_t k1_pair2(k1_bits &bb, const uint16_t *tl, uint32_t sh, uint32_t linbits, uint32_t slow)
{
  const uint32_t w0 = bb.peek();
  uint32_t e = tl[w0 >> sh], used = 0;
  if (e & slow) {
    if (e & 0x8000u) {
      uint32_t cw = 32 - sh;
      do { used += cw; cw = (e >> 10) & 7; e = tl[(e & 1023u) + ((w0 << used) >> (32 - cw))]; } while (e & 0x8000u);
    }
    if (e & slow) {                                        /* a leaf now: only the escape bit can still match */
      used += (e >> 8) & 31;
      int x = (e >> 4) & 15, y = e & 15;
      bb.skip(used);
      uint32_t w = bb.peek(), n = 0;
      if (x == 15) { x += (int)(w >> (32 - linbits)); w <<= linbits; n += linbits; }
      if (x) { if ((int)w < 0) x = -x; w <<= 1; n++; }
      if (y == 15) { y += (int)(w >> (32 - linbits)); w <<= linbits; n += linbits; }
      if (y) { if ((int)w < 0) y = -y; n++; }
      bb.skip(n);
      return __byte_perm((uint32_t)x, (uint32_t)y, 0x5410);
    }
  }
  const uint32_t len = used + ((e >> 8) & 31u);
  uint32_t x = (e >> 4) & 15u, y = e & 15u, nx, ny;
  asm("min.u32 %0, %1, 1;" : "=r"(nx) : "r"(x));
  asm("min.u32 %0, %1, 1;" : "=r"(ny) : "r"(y));
  uint32_t w = w0 << len;
  uint32_t m = (uint32_t)((int)w >> 31); x = (x ^ m) - m;
  w <<= nx;
  m = (uint32_t)((int)w >> 31); y = (y ^ m) - m;
  bb.skip(len + nx + ny);
  return __byte_perm(x, y, 0x5410);
}
#endif

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


/* Scalefactors + Huffman of ONE granule-channel by ONE thread, spectra through `ob`, scalefactor bytes
 * to scf[64] (zeroed by the caller).  Returns count1. */
__device__ __forceinline__ uint32_t k1_decode_gc(const uint32_t *sw /* the CTA's window of the compact main-data stream, shared memory */, const uint16_t *lut, const p3_tables *__restrict__ T,
                                                 const p3_gc *__restrict__ gcs, const p3_frame &fr, const p3_gc &g, int64_t f,
                                                 uint32_t gr, uint32_t ch, int64_t win0 /* stream byte of window byte 0 */, k1_out &ob, uint8_t *scf)
{
  uint32_t c1 = 0;
  {
    const bool ok = ch < fr.nch && !(fr.flags & (P3_FRAME_NODATA | P3_FRAME_BAD));
    const uint32_t p23l = ok ? P3_GC_P23L(g) : 0u;
    /* where this frame's bits start: main_data_begin bytes in front of its own data, relative to the window */
    const uint32_t fstart = (uint32_t)((int64_t)fr.main_pos - fr.main_begin - win0) * 8u;
    uint32_t pos = fstart + P3_GC_START(g);
    const uint32_t part2_start = pos;
    const bool is_short = P3_GC_WINSW(g) && P3_GC_BTYPE(g) == 2;
    /* ---- part 2: scalefactors (pdmp3.c:1382-1435).  The reference reads them whatever part2_3_length says: a zero-length
     *      part still takes its slen bits from the stream (and the next part starts behind them: the parsers' part-start
     *      rule).  ISO mode: a zero-length part carries no bits, its scalefactors are 0. ---- */
    if (ok) {
      const bool has2 = p23l != 0 || !(fr.flags & P3_FRAME_ISO);
      const uint32_t slen1 = has2 ? T->slen[P3_GC_SFCOMP(g)][0] : 0u, slen2 = has2 ? T->slen[P3_GC_SFCOMP(g)][1] : 0u;
      if (is_short) {
        uint32_t first = 0;
        if (P3_GC_MIXED(g)) { for (int sfb = 0; sfb < 8; sfb++) scf[sfb] = (uint8_t)p3_getbits(sw, pos, slen1); first = 3; }
        for (uint32_t sfb = first; sfb < 12; sfb++)
          for (int win = 0; win < 3; win++) scf[P3_SCF_S_OFF + 3 * sfb + win] = (uint8_t)p3_getbits(sw, pos, sfb < 6 ? slen1 : slen2);
      } else {
        const uint32_t scfsi = gr == 1 ? (fr.scfsi >> (4 * ch)) & 15u : 0u;
        /* granule 1 may reuse granule 0's scalefactors (scfsi): re-read them from granule 0's part 2 */
        const p3_gc g0 = gcs[4 * f + ch];
        /* (granule 0 a short block: the reference copies whatever an earlier frame left in scalefac_l[0]; ISO forbids scfsi
         *  there, and 0 is what this decoder uses -- documented deviation) */
        const bool g0_ok = !(P3_GC_WINSW(g0) && P3_GC_BTYPE(g0) == 2) && (P3_GC_P23L(g0) != 0 || !(fr.flags & P3_FRAME_ISO)) && has2;
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
      const uint16_t *tls[3]; uint32_t shs[3], lbs[3];
#ifdef K1_PAIR2
      uint32_t sms[3];
#endif
      #pragma unroll
      for (int r = 0; r < 3; r++) {
        const uint32_t t = P3_GC_TSEL(g, r);
        const int book = T->table_book[t];
        /* empty tables 0/4/14: zeros, no bits (pdmp3.c:1599-1602) = a pseudo book whose every leaf is (0, 0, length 0) */
        tls[r] = lut + (book < 0 ? T->hlut_zero : T->book_base[book]);
        shs[r] = 32u - (book < 0 ? 1u : T->book_pbits[book]);
        lbs[r] = book < 0 ? 0u : T->table_linbits[t];
#ifdef K1_PAIR2
        sms[r] = lbs[r] ? 0xc000u : 0x8000u;
#endif
      }
      /* one loop over all big_values pairs; the table changes at the region boundaries, so every lane of the warp
       * stays in the same loop whatever its region split.  Four pairs (half a sector) per iteration leave straight
       * from registers as one 16-byte store; the up to three pairs left over go through the staging ring below. */
      const uint16_t *tl = tls[0]; uint32_t sh = shs[0], lb = lbs[0];
      uint32_t nsw = r1s;                                   /* next value index at which the table changes */
      k1_bits bb; bb.init(sw, pos);
      uint32_t i = 0;
#ifdef K1_PAIR2
      uint32_t sm = sms[0];
      auto pair_at = [&](uint32_t at) -> uint32_t {
        if (at == nsw) {
          if (at == r1s) { tl = tls[1]; sh = shs[1]; lb = lbs[1]; sm = sms[1]; }
          if (at == r2s) { tl = tls[2]; sh = shs[2]; lb = lbs[2]; sm = sms[2]; }
          nsw = r2s > at ? r2s : 0xffffffffu;
        }
        return k1_pair2(bb, tl, sh, lb, sm);
      };
#else
      auto pair_at = [&](uint32_t at) -> uint32_t {
        if (at == nsw) {
          if (at == r1s) { tl = tls[1]; sh = shs[1]; lb = lbs[1]; }
          if (at == r2s) { tl = tls[2]; sh = shs[2]; lb = lbs[2]; }
          nsw = r2s > at ? r2s : 0xffffffffu;
        }
        return k1_pair(bb, tl, sh, lb);
      };
#endif
      for (; i + 8 <= bv2; i += 8) {
        uint4 o;
        o.x = pair_at(i); o.y = pair_at(i + 2); o.z = pair_at(i + 4); o.w = pair_at(i + 6);
        ob.dst[i >> 3] = o;
      }
      ob.pw = i >> 1;
      for (; i < bv2; i += 2) ob.put(pair_at(i));
      /* count1 quads (pdmp3.c:2091-2103) */
      uint32_t is_pos = bv2;
      const bool tabB = P3_GC_C1TAB(g);                  /* reference quirk Q1: table B = leaf 0011, no code bits */
      const bool isoB = tabB && (fr.flags & P3_FRAME_ISO); /* ISO mode: table B is the 4-bit code it is, value = ~code */
      const uint32_t qbase = T->book_base[T->table_book[32]], qbits = T->book_pbits[T->table_book[32]];
      pos = bb.pos();
      while (is_pos <= 572 && pos <= bit_pos_end) {
        uint32_t w = bb.peek(), used = 0, leaf = 3;
        if (!tabB) { uint32_t e = K1_LD(lut + qbase + (w >> (32 - qbits))); used = (e >> 8) & 31; leaf = e & 15; w <<= used; }
        else if (isoB) { leaf = (~w) >> 28; used = 4; w <<= 4; }
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


#ifdef K1_COOP_TRIAL
/* ---------------------------------------------------------------------------------------------------------------------------
 * TRIAL (never built into the product; tools/build_variant.sh coop -DK1_COOP_TRIAL): north_star's "one warp per granule,
 * __ballot/__shfl to walk the code words" -- the 32 lanes of a warp decode ONE part together: every lane decodes a pair
 * speculatively at one of 32 consecutive bit offsets, the true chain is then followed through the lanes with shuffles
 * (lane 0 -> lane 0 + its length -> ...), typically 3-4 symbols per window of 32 bits; the window then moves on.  Same LUTs, same
 * bit reader, same outputs as the thread-per-part decoder above, so the GPU parity tests run on it unchanged.  Measured
 * against it in profiles/ (DESIGN.md 4.6). */
__device__ __forceinline__ uint32_t k1_decode_gc_coop(const uint32_t *sw, const uint16_t *lut, const p3_tables *__restrict__ T,
                                                      const p3_gc *__restrict__ gcs, const p3_frame &fr, const p3_gc &g, int64_t f,
                                                      uint32_t gr, uint32_t ch, int64_t win0, uint32_t *dstw /* the part's 288-word row */, uint8_t *scf)
{
  const uint32_t lane = threadIdx.x & 31;
  uint32_t c1 = 0;
  const bool ok = ch < fr.nch && !(fr.flags & (P3_FRAME_NODATA | P3_FRAME_BAD));
  const uint32_t p23l = ok ? P3_GC_P23L(g) : 0u;
  const uint32_t fstart = (uint32_t)((int64_t)fr.main_pos - fr.main_begin - win0) * 8u;
  uint32_t pos = fstart + P3_GC_START(g);
  const uint32_t part2_start = pos;
  const bool is_short = P3_GC_WINSW(g) && P3_GC_BTYPE(g) == 2;
  if (ok) {                                                  /* part 2: every lane runs the same (short, uniform) code; lane 0 stores */
    uint8_t tmp[64];
    #pragma unroll
    for (int i = 0; i < 64; i++) tmp[i] = 0;
    const bool has2 = p23l != 0 || !(fr.flags & P3_FRAME_ISO);
    const uint32_t slen1 = has2 ? T->slen[P3_GC_SFCOMP(g)][0] : 0u, slen2 = has2 ? T->slen[P3_GC_SFCOMP(g)][1] : 0u;
    if (is_short) {
      uint32_t first = 0;
      if (P3_GC_MIXED(g)) { for (int sfb = 0; sfb < 8; sfb++) tmp[sfb] = (uint8_t)p3_getbits(sw, pos, slen1); first = 3; }
      for (uint32_t sfb = first; sfb < 12; sfb++)
        for (int win = 0; win < 3; win++) tmp[P3_SCF_S_OFF + 3 * sfb + win] = (uint8_t)p3_getbits(sw, pos, sfb < 6 ? slen1 : slen2);
    } else {
      const uint32_t scfsi = gr == 1 ? (fr.scfsi >> (4 * ch)) & 15u : 0u;
      const p3_gc g0 = gcs[4 * f + ch];
      const bool g0_ok = !(P3_GC_WINSW(g0) && P3_GC_BTYPE(g0) == 2) && (P3_GC_P23L(g0) != 0 || !(fr.flags & P3_FRAME_ISO)) && has2;
      const uint32_t a1 = T->slen[P3_GC_SFCOMP(g0)][0], a2 = T->slen[P3_GC_SFCOMP(g0)][1];
      const uint32_t g0pos = fstart + P3_GC_START(g0);
      for (int band = 0; band < 4; band++) {
        const int lo = band == 0 ? 0 : band == 1 ? 6 : band == 2 ? 11 : 16, hi = band == 0 ? 6 : band == 1 ? 11 : band == 2 ? 16 : 21;
        if ((scfsi >> band) & 1) {
          uint32_t p = g0pos + (band == 0 ? 0u : band == 1 ? 6 * a1 : band == 2 ? 11 * a1 : 11 * a1 + 5 * a2), nb = band < 2 ? a1 : a2;
          for (int sfb = lo; sfb < hi; sfb++) tmp[sfb] = g0_ok ? (uint8_t)p3_getbits(sw, p, nb) : 0;
        } else
          for (int sfb = lo; sfb < hi; sfb++) tmp[sfb] = (uint8_t)p3_getbits(sw, pos, band < 2 ? slen1 : slen2);
      }
    }
    if (lane == 0) for (int i = 0; i < 64; i++) scf[i] = tmp[i];
  }
  uint32_t nw = 0;                                           /* words (pairs) of the row written so far */
  if (p23l) {
    const uint32_t bit_pos_end = part2_start + p23l - 1;
    uint32_t r1s, r2s;
    if (is_short) { r1s = 36; r2s = 576; }
    else { r1s = T->sfb_l[fr.sfreq][P3_GC_REG0(g) + 1]; r2s = T->sfb_l[fr.sfreq][P3_GC_REG0(g) + P3_GC_REG1(g) + 2]; }
    const uint32_t bv2 = 2 * P3_GC_BIGV(g);
    uint32_t i = 0;
    for (int r = 0; r < 3; r++) {
      const uint32_t t = P3_GC_TSEL(g, r);
      const int book = T->table_book[t];
      const uint16_t *tl = lut + (book < 0 ? T->hlut_zero : T->book_base[book]);
      const uint32_t sh = 32u - (book < 0 ? 1u : T->book_pbits[book]), lb = book < 0 ? 0u : T->table_linbits[t];
      const uint32_t iend = min(bv2, r == 0 ? r1s : r == 1 ? r2s : 576u);
      while (i < iend) {
        k1_bits bb; bb.init(sw, pos + lane);                 /* speculation: a pair that starts at bit pos + lane */
        const uint32_t v = k1_pair(bb, tl, sh, lb);
        const uint32_t len = bb.pos() - (pos + lane);
        uint32_t cur = 0, cnt = 0, mine = 0;
        while (cur < 32 && cnt < 32 && i < iend) {           /* the true chain through the lanes (cnt: the empty tables' code words have length 0) */
          const uint32_t pv = __shfl_sync(0xffffffffu, v, cur), pl = __shfl_sync(0xffffffffu, len, cur);
          if (lane == cnt) mine = pv;
          cnt++; i += 2; cur += pl;
        }
        if (lane < cnt) dstw[nw + lane] = mine;              /* the window's pairs, one coalesced store */
        nw += cnt; pos += cur;
      }
    }
    /* count1 quads, same scheme (pdmp3.c:2091-2106) */
    uint32_t is_pos = bv2;
    const bool tabB = P3_GC_C1TAB(g), isoB = tabB && (fr.flags & P3_FRAME_ISO);
    const uint32_t qbase = T->book_base[T->table_book[32]], qbits = T->book_pbits[T->table_book[32]];
    while (is_pos <= 572 && pos <= bit_pos_end) {
      const uint32_t p0 = pos + lane;
      uint32_t w = p3_peek32(sw, p0), used = 0, leaf = 3;
      if (!tabB) { const uint32_t e = K1_LD(lut + qbase + (w >> (32 - qbits))); used = (e >> 8) & 31; leaf = e & 15; w <<= used; }
      else if (isoB) { leaf = (~w) >> 28; used = 4; w <<= 4; }
      int v = (leaf >> 3) & 1, ww = (leaf >> 2) & 1, x = (leaf >> 1) & 1, y = leaf & 1;
      if (v) { if (w >> 31) v = -1; w <<= 1; used++; }
      if (ww) { if (w >> 31) ww = -1; w <<= 1; used++; }
      if (x) { if (w >> 31) x = -1; w <<= 1; used++; }
      if (y) { if (w >> 31) y = -1; used++; }
      const uint32_t q0 = (uint32_t)(v & 0xffff) | ((uint32_t)ww << 16), q1 = (uint32_t)(x & 0xffff) | ((uint32_t)y << 16);
      uint32_t cur = 0, cnt = 0, m0 = 0, m1 = 0;
      while (cur < 32 && cnt < 32 && is_pos <= 572 && pos + cur <= bit_pos_end) {
        const uint32_t a = __shfl_sync(0xffffffffu, q0, cur), b = __shfl_sync(0xffffffffu, q1, cur), pl = __shfl_sync(0xffffffffu, used, cur);
        if (lane == cnt) { m0 = a; m1 = b; }
        cnt++; is_pos += 4; cur += pl;
        if (pl == 0) break;                                  /* table B in compat mode with no sign bits: cannot happen (leaf 0011 has two), guard anyway */
      }
      if (lane < cnt) { dstw[nw + 2 * lane] = m0; dstw[nw + 2 * lane + 1] = m1; }
      nw += 2 * cnt; pos += cur;
    }
    if (pos > bit_pos_end + 1) is_pos = is_pos >= 4 ? is_pos - 4 : 0;
    c1 = is_pos;
  }
  __syncwarp();
  for (uint32_t k = (c1 >> 1) + lane; k < 288; k += 32) dstw[k] = 0;        /* rzero region (and whatever a rolled-back quad left) */
  return c1;
}
#endif
