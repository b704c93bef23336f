/* oracle/p3_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into libpdmp3_b200.so; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * A plain-C, scalar restatement of the reference's granule decode path, written against the
 * batch descriptors of include/pdmp3_b200.h (so it also checks the host parser).  Each stage
 * cites the reference lines whose ARITHMETIC (operand order, float/double width, rounding
 * points) it follows, because the parity target is bit-exact agreement with the reference
 * compiled -O2 without FMA.  It is pinned by tests/test_cpu_oracle.py against the compiled
 * reference itself (oracle/_ref) and by the committed fixtures in tests/golden/.
 *
 * Deliberate differences from the reference (all outside the parity envelope):
 *   Q5  pseudo-band scalefactors (long sfb 21, short sfb 12) and pretab[21] are 0 (ISO), the
 *       reference reads out of bounds;
 *   Q10 malformed side info is flagged by the parser and decodes as silence.
 */
#include "../include/pdmp3_b200.h"
#include "p3_tables.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>

typedef struct {
  int16_t *is_huff;  int32_t *count1;  uint8_t *scf_l;  uint8_t *scf_s;
  float *xr_req, *xr_reo, *xr_ste, *xr_ali, *y_hyb;
} p3o_taps;

/* ---- MSB-first bit reader over the header-stripped main-data stream (pdmp3.c:1489-1541) ---- */
typedef struct { const uint8_t *d; uint64_t pos; } rd_t;
static inline unsigned rd_bit(rd_t *r) { unsigned b = (r->d[r->pos >> 3] >> (7 - (r->pos & 7))) & 1; r->pos++; return b; }
static inline unsigned rd_bits(rd_t *r, unsigned n) { unsigned v = 0; while (n--) v = (v << 1) | rd_bit(r); return v; }

/* ---- Huffman word: bit-serial match against the canonical code list (pdmp3.c:1593-1643) ---- */
static void huff_word(rd_t *r, const p3_tables *T, unsigned table, int iso, int *x, int *y, int *v, int *w)
{
  int book = T->table_book[table];
  *x = *y = *v = *w = 0;
  if (book < 0) return;                                   /* empty tables 0/4/14: zeros, no bits (1599-1602) */
  int leaf = -1;
  if (table == 33 && !iso) leaf = 0x03;                   /* Q1: table B is mis-wired to a leaf: no code bits, value 0011 */
  else if (table == 33) leaf = (int)(~rd_bits(r, 4) & 15u);   /* ISO mode: count1 table B is a plain 4-bit code, value = ~code (ISO 11172-3 table B.7) */
  else {
    const p3_hcode *c; int n = p3_book_codes(book, &c);
    unsigned code = 0, len = 0;
    while (leaf < 0 && len < 24) {
      code = (code << 1) | rd_bit(r); len++;
      for (int k = 0; k < n; k++) if (c[k].len == len && c[k].code == code) { leaf = (c[k].x << 4) | c[k].y; break; }
    }
  }
  int lx = (leaf >> 4) & 15, ly = leaf & 15;
  if (table > 31) {                                        /* quads (1627-1635) */
    *v = (ly >> 3) & 1; *w = (ly >> 2) & 1; *x = (ly >> 1) & 1; *y = ly & 1;
    if (*v && rd_bit(r)) *v = -*v;
    if (*w && rd_bit(r)) *w = -*w;
    if (*x && rd_bit(r)) *x = -*x;
    if (*y && rd_bit(r)) *y = -*y;
  } else {                                                 /* pairs (1636-1640) */
    unsigned lb = T->table_linbits[table];
    if (lb && lx == 15) lx += (int)rd_bits(r, lb);
    if (lx && rd_bit(r)) lx = -lx;
    if (lb && ly == 15) ly += (int)rd_bits(r, lb);
    if (ly && rd_bit(r)) ly = -ly;
    *x = lx; *y = ly;
  }
}

typedef struct {
  unsigned scf_l[2][2][22];      /* [gr][ch][sfb]; [21] = pseudo band, always 0 */
  unsigned scf_s[2][2][13][3];   /* [12] = pseudo band, always 0 */
  unsigned count1[2][2];
  float store[2][32][18];        /* IMDCT overlap (pdmp3.c:1755) */
  float vfifo[2][16][64];        /* polyphase history as a ring: slot t lives at [t & 15] (pdmp3.c:1983, 2006-2007) */
  uint64_t slot[2];
} ostate;

/* ---- part 2: scalefactors (pdmp3.c:1376-1435) ---- */
static void read_scalefacs(rd_t *r, const p3_tables *T, ostate *S, const p3_frame *fr, const p3_gc *g, unsigned gr, unsigned ch)
{
  unsigned slen1 = T->slen[P3_GC_SFCOMP(*g)][0], slen2 = T->slen[P3_GC_SFCOMP(*g)][1];
  /* the reference reads the scalefactor bits even of a part with part2_3_length == 0 (1379-1435 run before Read_Huffman's
   * early return, 2057-2061); ISO mode: such a part has no bits, its scalefactors are 0 */
  if (P3_GC_P23L(*g) == 0 && (fr->flags & P3_FRAME_ISO)) {
    memset(S->scf_l[gr][ch], 0, sizeof S->scf_l[gr][ch]); memset(S->scf_s[gr][ch], 0, sizeof S->scf_s[gr][ch]);
    return;
  }
  if (P3_GC_WINSW(*g) && P3_GC_BTYPE(*g) == 2) {
    unsigned first = 0;
    if (P3_GC_MIXED(*g)) { for (unsigned sfb = 0; sfb < 8; sfb++) S->scf_l[gr][ch][sfb] = rd_bits(r, slen1); first = 3; }
    for (unsigned sfb = first; sfb < 12; sfb++)
      for (unsigned win = 0; win < 3; win++) S->scf_s[gr][ch][sfb][win] = rd_bits(r, sfb < 6 ? slen1 : slen2);
  } else {
    static const unsigned lo[5] = {0, 6, 11, 16, 21};
    unsigned scfsi = (fr->scfsi >> (4 * ch)) & 15;
    for (unsigned band = 0; band < 4; band++)
      for (unsigned sfb = lo[band]; sfb < lo[band + 1]; sfb++) {
        if (gr == 1 && ((scfsi >> band) & 1)) S->scf_l[1][ch][sfb] = S->scf_l[0][ch][sfb];
        else S->scf_l[gr][ch][sfb] = rd_bits(r, band < 2 ? slen1 : slen2);
      }
  }
}

/* ---- part 3: Read_Huffman (pdmp3.c:2051-2115) ---- */
static void read_huffman(rd_t *r, const p3_tables *T, ostate *S, const p3_frame *fr, const p3_gc *g,
                         unsigned gr, unsigned ch, uint64_t part2_start, float is[576])
{
  unsigned p23l = P3_GC_P23L(*g);
  const int iso = (fr->flags & P3_FRAME_ISO) != 0;
  if (p23l == 0) { memset(is, 0, 576 * sizeof(float)); if (iso) S->count1[gr][ch] = 0; return; }      /* reference: count1 stays stale (Q6) */
  uint64_t bit_pos_end = part2_start + p23l - 1;
  unsigned r1s, r2s;
  if (P3_GC_WINSW(*g) && P3_GC_BTYPE(*g) == 2) { r1s = 36; r2s = 576; }
  else { r1s = T->sfb_l[fr->sfreq][P3_GC_REG0(*g) + 1]; r2s = T->sfb_l[fr->sfreq][P3_GC_REG0(*g) + P3_GC_REG1(*g) + 2]; }
  unsigned is_pos, bv2 = 2 * P3_GC_BIGV(*g);
  int x, y, v, w;
  for (is_pos = 0; is_pos < bv2; is_pos++) {
    unsigned t = is_pos < r1s ? P3_GC_TSEL(*g, 0) : is_pos < r2s ? P3_GC_TSEL(*g, 1) : P3_GC_TSEL(*g, 2);
    huff_word(r, T, t, iso, &x, &y, &v, &w);
    is[is_pos++] = (float)x; is[is_pos] = (float)y;
  }
  unsigned tq = 32 + P3_GC_C1TAB(*g);
  for (is_pos = bv2; is_pos <= 572 && r->pos <= bit_pos_end; is_pos++) {
    huff_word(r, T, tq, iso, &x, &y, &v, &w);
    is[is_pos++] = (float)v; if (is_pos >= 576) break;
    is[is_pos++] = (float)w; if (is_pos >= 576) break;
    is[is_pos++] = (float)x; if (is_pos >= 576) break;
    is[is_pos] = (float)y;
  }
  if (r->pos > bit_pos_end + 1) is_pos = is_pos >= 4 ? is_pos - 4 : 0;   /* 2105-2106 (the reference would wrap below 4) */
  S->count1[gr][ch] = is_pos;
  for (; is_pos < 576; is_pos++) is[is_pos] = 0.0f;
  r->pos = bit_pos_end + 1;
}

/* ---- L3_Requantize (pdmp3.c:1829-1905, 2121-2152) ---- */
static inline float requant(const p3_tables *T, float isv, unsigned e2, int q)
{
  float t1 = T->t1h[e2], t2 = T->t2[q + P3_T2_BIAS];
  float t3 = isv < 0.0f ? -T->pow43[(unsigned)(-isv)] : T->pow43[(unsigned)isv];
  return t1 * t2 * t3;
}
static void requantize(const p3_tables *T, ostate *S, const p3_frame *fr, const p3_gc *g, unsigned gr, unsigned ch, float is[576])
{
  unsigned sf = fr->sfreq, mult = P3_GC_SCALE(*g) ? 2 : 1, c1 = S->count1[gr][ch], i;
  int gg = (int)P3_GC_GAIN(*g) - 210;
  if (P3_GC_WINSW(*g) && P3_GC_BTYPE(*g) == 2) {
    i = 0;
    if (P3_GC_MIXED(*g)) {
      for (; i < 36; i++) {
        unsigned sfb = T->line_sfb_l[sf][i];
        is[i] = requant(T, is[i], mult * (S->scf_l[gr][ch][sfb] + P3_GC_PREF(*g) * T->pretab[sfb]), gg);
      }
    }
    while (i < c1) {                                  /* whole sfb triplets while i < count1 (1856,1877) */
      unsigned sfb = T->line_sfbw_s[sf][i] & 15, wl = T->sfb_s[sf][sfb + 1] - T->sfb_s[sf][sfb];
      for (unsigned win = 0; win < 3; win++)
        for (unsigned j = 0; j < wl; j++, i++)
          is[i] = requant(T, is[i], mult * (sfb < 12 ? S->scf_s[gr][ch][sfb][win] : 0), gg - 8 * (int)P3_GC_SBG(*g, win));
    }
  } else {
    for (i = 0; i < c1 && i < 576; i++) {
      unsigned sfb = T->line_sfb_l[sf][i];
      unsigned s = sfb < 21 ? S->scf_l[gr][ch][sfb] + P3_GC_PREF(*g) * T->pretab[sfb] : 0;
      is[i] = requant(T, is[i], mult * s, gg);
    }
  }
}

/* ---- L3_Reorder (pdmp3.c:1786-1823): unconditional permutation of the short part ---- */
static void reorder(const p3_tables *T, const p3_frame *fr, const p3_gc *g, float is[576])
{
  if (!(P3_GC_WINSW(*g) && P3_GC_BTYPE(*g) == 2)) return;
  float re[576];
  unsigned first = P3_GC_MIXED(*g) ? 36 : 0;
  for (unsigned d = first; d < 576; d++) re[d] = is[T->reorder_src[fr->sfreq][d]];
  memcpy(is + first, re + first, (576 - first) * sizeof(float));
}

/* ---- L3_Stereo (pdmp3.c:1911-1972, 2158-2220) ---- */
static void stereo(const p3_tables *T, ostate *S, const p3_frame *fr, const p3_gc *g0, unsigned gr, float l[576], float r[576])
{
  if (fr->mode != 1 || fr->mode_ext == 0) return;
  unsigned sf = fr->sfreq;
  if (fr->flags & P3_FRAME_ISO) {
    /* ISO mode (not the reference): a band starting at or above the right channel's count1 whose intensity position
     * (RIGHT channel's scalefactor) is below 7 is intensity coded, short blocks with the ratio multiply as well; every
     * other line below max(count1) takes MS if that is on.  Same band walk as the reference otherwise. */
    unsigned c0 = S->count1[gr][0], c1r = S->count1[gr][1], msn = (fr->mode_ext & 2) ? (c0 > c1r ? c0 : c1r) : 0;
    unsigned char done[576]; memset(done, 0, sizeof done);
    if (fr->mode_ext & 1) {
      int sh = P3_GC_WINSW(*g0) && P3_GC_BTYPE(*g0) == 2;
      unsigned lim = sh ? (P3_GC_MIXED(*g0) ? 8 : 0) : 21, first = sh ? (P3_GC_MIXED(*g0) ? 3 : 0) : 12;
      for (unsigned sfb = 0; sfb < lim; sfb++) if (T->sfb_l[sf][sfb] >= c1r) {
        unsigned p = S->scf_l[gr][1][sfb]; if (p >= 7) continue;
        for (unsigned i = T->sfb_l[sf][sfb]; i < T->sfb_l[sf][sfb + 1]; i++) { float x = l[i]; l[i] = T->is_l[p] * x; r[i] = T->is_r[p] * x; done[i] = 1; }
      }
      for (unsigned sfb = first; sfb < 12; sfb++) if (3u * T->sfb_s[sf][sfb] >= c1r) {
        unsigned wl = T->sfb_s[sf][sfb + 1] - T->sfb_s[sf][sfb];
        for (unsigned win = 0; win < 3; win++) {
          unsigned p = S->scf_s[gr][1][sfb][win]; if (p >= 7) continue;
          for (unsigned i = 3 * T->sfb_s[sf][sfb] + wl * win, e = i + wl; i < e; i++) { float x = l[i]; l[i] = T->is_l[p] * x; r[i] = T->is_r[p] * x; done[i] = 1; }
        }
      }
    }
    for (unsigned i = 0; i < msn && i < 576; i++) if (!done[i]) {
      float a = l[i] + r[i], b = l[i] - r[i];
      l[i] = (float)(a * 0.70710678118654752440);
      r[i] = (float)(b * 0.70710678118654752440);
    }
    return;
  }
  if (fr->mode_ext & 2) {
    unsigned n = S->count1[gr][0] > S->count1[gr][1] ? S->count1[gr][1] : S->count1[gr][0];   /* min, sic (1920) */
    for (unsigned i = 0; i < n && i < 576; i++) {
      float a = l[i] + r[i], b = l[i] - r[i];
      l[i] = (float)(a * 0.70710678118654752440);     /* float sum times a DOUBLE constant (168,1923-1926) */
      r[i] = (float)(b * 0.70710678118654752440);
    }
  }
  if (fr->mode_ext & 1) {
    unsigned c1r = S->count1[gr][1];
    if (P3_GC_WINSW(*g0) && P3_GC_BTYPE(*g0) == 2) {
      unsigned first = 0;
      if (P3_GC_MIXED(*g0)) {
        for (unsigned sfb = 0; sfb < 8; sfb++) if (T->sfb_l[sf][sfb] >= c1r) {
          unsigned p = S->scf_l[gr][0][sfb]; if (p == 7) continue;
          for (unsigned i = T->sfb_l[sf][sfb]; i < T->sfb_l[sf][sfb + 1]; i++) { float x = l[i]; l[i] = T->is_l[p & 7] * x; r[i] = T->is_r[p & 7] * x; }
        }
        first = 3;
      }
      for (unsigned sfb = first; sfb < 12; sfb++) if (3u * T->sfb_s[sf][sfb] >= c1r) {
        unsigned wl = T->sfb_s[sf][sfb + 1] - T->sfb_s[sf][sfb];
        for (unsigned win = 0; win < 3; win++) {
          if (S->scf_s[gr][0][sfb][win] == 7) continue;
          /* Q4: the reference assigns through an `unsigned` (2191,2212-2213): both channels get
           * (float)(unsigned)x; for negative x this is UB, x86-64 gcc gives the low 32 bits of the int64 */
          for (unsigned i = 3 * T->sfb_s[sf][sfb] + wl * win, e = i + wl; i < e; i++) {
            float x = (float)(unsigned)(long long)l[i]; l[i] = x; r[i] = x;
          }
        }
      }
    } else {
      for (unsigned sfb = 0; sfb < 21; sfb++) if (T->sfb_l[sf][sfb] >= c1r) {
        unsigned p = S->scf_l[gr][0][sfb]; if (p == 7) continue;                   /* channel-0 scalefactor, sic (2163) */
        for (unsigned i = T->sfb_l[sf][sfb]; i < T->sfb_l[sf][sfb + 1]; i++) { float x = l[i]; l[i] = T->is_l[p & 7] * x; r[i] = T->is_r[p & 7] * x; }
      }
    }
  }
}

/* ---- L3_Antialias (pdmp3.c:1706-1732) ---- */
static void antialias(const p3_tables *T, const p3_gc *g, float is[576])
{
  int sh = P3_GC_WINSW(*g) && P3_GC_BTYPE(*g) == 2;
  if (sh && !P3_GC_MIXED(*g)) return;
  unsigned sblim = sh ? 2 : 32;
  for (unsigned sb = 1; sb < sblim; sb++) for (unsigned i = 0; i < 8; i++) {
    unsigned li = 18 * sb - 1 - i, ui = 18 * sb + i;
    float lb = is[li] * T->cs[i] - is[ui] * T->ca[i];
    float ub = is[ui] * T->cs[i] + is[li] * T->ca[i];
    is[li] = lb; is[ui] = ub;
  }
}

/* ---- L3_Hybrid_Synthesis + IMDCT_Win + L3_Frequency_Inversion (pdmp3.c:1649-1700,1738-1780) ---- */
static void hybrid(const p3_tables *T, ostate *S, const p3_gc *g, unsigned ch, float is[576])
{
  for (unsigned sb = 0; sb < 32; sb++) {
    unsigned bt = (P3_GC_WINSW(*g) && P3_GC_MIXED(*g) && sb < 2) ? 0 : P3_GC_BTYPE(*g);
    float raw[36], *in = is + 18 * sb;
    for (int i = 0; i < 36; i++) raw[i] = 0.0f;
    if (bt == 2) {
      for (unsigned w = 0; w < 3; w++) for (unsigned p = 0; p < 12; p++) {
        float sum = 0.0f;
        for (unsigned m = 0; m < 6; m++) sum += in[w + 3 * m] * T->cos12[m][p];
        raw[6 * w + p + 6] += sum * T->imdct_win[2][p];
      }
    } else {
      for (unsigned p = 0; p < 36; p++) {
        float sum = 0.0f;
        for (unsigned m = 0; m < 18; m++) sum += in[m] * T->cos36[m][p];
        raw[p] = sum * T->imdct_win[bt][p];
      }
    }
    for (unsigned i = 0; i < 18; i++) {
      float yv = raw[i] + S->store[ch][sb][i];
      S->store[ch][sb][i] = raw[i + 18];
      in[i] = ((sb & 1) && (i & 1)) ? -yv : yv;          /* frequency inversion (1741-1743) */
    }
  }
}

/* ---- L3_Subband_Synthesis (pdmp3.c:1978-2045) for one channel of one granule ---- */
static void polyphase(const p3_tables *T, ostate *S, unsigned ch, const float y[576], int16_t *pcm, unsigned stride)
{
  for (unsigned ss = 0; ss < 18; ss++) {
    uint64_t t = S->slot[ch]++;
    float *V = S->vfifo[ch][t & 15];
    for (unsigned i = 0; i < 64; i++) {
      float sum = 0.0f;
      for (unsigned j = 0; j < 32; j++) sum += T->synth_n[i][j] * y[j * 18 + ss];
      V[i] = sum;
    }
    for (unsigned j = 0; j < 32; j++) {
      float sum = 0.0f;
      for (unsigned k = 0; k < 16; k++) {                /* U[32k+j]: k even -> V(t-k)[j], k odd -> V(t-k)[32+j] (2015-2026) */
        const float *Vk = S->vfifo[ch][(t - k) & 15];
        float u = Vk[(k & 1) ? 32 + j : j] * T->synth_d[32 * k + j];
        sum += u;
      }
      double d = sum * 32767.0;                          /* double multiply, C truncation (2028) */
      int32_t samp = (d > -2147483649.0 && d < 2147483648.0) ? (int32_t)d : INT_MIN;   /* x86 cvttsd2si out-of-range result */
      if (samp > 32767) samp = 32767; else if (samp < -32767) samp = -32767;
      pcm[(32 * ss + j) * stride] = (int16_t)samp;
    }
  }
}

/* Decode n_frames parsed frames of `raw`.  pcm: [n_pcm_frames][1152][nch] int16.  taps may be NULL.
 * Starts from the zero state of pdmp3_open_feed (pdmp3.c:2377-2379) on a zeroed handle (G0). */
int p3o_decode(const uint8_t *raw, const p3_frame *fr, const p3_gc *gc, int64_t n_frames, const p3o_taps *tp, int16_t *pcm)
{
  const p3_tables *T = p3_tables_get();
  ostate *S = calloc(1, sizeof *S);
  /* header-stripped main-data stream == what Get_Main_Data assembles frame by frame (1096-1122) */
  uint64_t total = n_frames ? fr[n_frames - 1].main_pos + fr[n_frames - 1].main_size - fr[0].main_pos : 0, base = n_frames ? fr[0].main_pos : 0;
  uint8_t *ms = calloc(total + 4096, 1);   /* slack: a corrupt part may be read far past its end */
  for (int64_t f = 0; f < n_frames; f++) memcpy(ms + (fr[f].main_pos - base), raw + fr[f].main_off, fr[f].main_size);
  float (*is)[2][576] = malloc(sizeof(float) * 4 * 576);
  for (int64_t f = 0; f < n_frames; f++) {
    const p3_frame *F = &fr[f];
    unsigned nch = F->nch;
    int silent = (F->flags & (P3_FRAME_NODATA | P3_FRAME_BAD)) != 0 || F->main_pos - base < F->main_begin;
    for (unsigned gr = 0; gr < 2; gr++) for (unsigned ch = 0; ch < nch; ch++) {
      const p3_gc *g = &gc[4 * f + 2 * gr + ch];
      size_t o = ((size_t)f * 2 + gr) * 2 + ch;
      if (silent) { memset(is[gr][ch], 0, sizeof is[gr][ch]); S->count1[gr][ch] = 0; }
      else {
        rd_t r = {ms, (F->main_pos - base - F->main_begin) * 8 + P3_GC_START(*g)};
        uint64_t p2 = r.pos;
        read_scalefacs(&r, T, S, F, g, gr, ch);
        read_huffman(&r, T, S, F, g, gr, ch, p2, is[gr][ch]);
      }
      if (tp && tp->is_huff) for (int i = 0; i < 576; i++) tp->is_huff[o * 576 + i] = (int16_t)is[gr][ch][i];
      if (tp && tp->count1) tp->count1[o] = (int32_t)S->count1[gr][ch];
      if (tp && tp->scf_l) for (int i = 0; i < 21; i++) tp->scf_l[o * 21 + i] = (uint8_t)S->scf_l[gr][ch][i];
      if (tp && tp->scf_s) for (int i = 0; i < 36; i++) tp->scf_s[o * 36 + i] = (uint8_t)S->scf_s[gr][ch][i / 3][i % 3];
    }
    for (unsigned gr = 0; gr < 2; gr++) {               /* Decode_L3 stage order (pdmp3.c:1029-1047) */
      for (unsigned ch = 0; ch < nch; ch++) {
        const p3_gc *g = &gc[4 * f + 2 * gr + ch];
        size_t o = (((size_t)f * 2 + gr) * 2 + ch) * 576;
        requantize(T, S, F, g, gr, ch, is[gr][ch]); if (tp && tp->xr_req) memcpy(tp->xr_req + o, is[gr][ch], 2304);
        reorder(T, F, g, is[gr][ch]);               if (tp && tp->xr_reo) memcpy(tp->xr_reo + o, is[gr][ch], 2304);
      }
      if (nch == 2) stereo(T, S, F, &gc[4 * f + 2 * gr], gr, is[gr][0], is[gr][1]);
      for (unsigned ch = 0; ch < nch; ch++) {
        const p3_gc *g = &gc[4 * f + 2 * gr + ch];
        size_t o = (((size_t)f * 2 + gr) * 2 + ch) * 576;
        if (tp && tp->xr_ste) memcpy(tp->xr_ste + o, is[gr][ch], 2304);
        antialias(T, g, is[gr][ch]);                if (tp && tp->xr_ali) memcpy(tp->xr_ali + o, is[gr][ch], 2304);
        hybrid(T, S, g, ch, is[gr][ch]);            if (tp && tp->y_hyb) memcpy(tp->y_hyb + o, is[gr][ch], 2304);
        if (F->flags & P3_FRAME_DECODE)
          polyphase(T, S, ch, is[gr][ch], pcm + ((size_t)F->pcm_index * 1152 + gr * 576) * nch + ch, nch);
        else { int16_t scratch[576]; polyphase(T, S, ch, is[gr][ch], scratch, 1); }
      }
    }
  }
  free(is); free(ms); free(S);
  return 0;
}
