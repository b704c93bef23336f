/* p3_synth.c -- see p3_synth.h.  Plain C; uses the canonical code lists of p3_tables. */
#include "p3_synth.h"
#include "../pdmp3_b200/csrc/p3_tables.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rnd(rng_t *r) { uint64_t z = (r->s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
static inline uint32_t rndn(rng_t *r, uint32_t n) { return (uint32_t)((rnd(r) >> 11) % n); }
static inline double rndu(rng_t *r) { return (double)(rnd(r) >> 11) * (1.0 / 9007199254740992.0); }

/* ---- bit writer over the logical (header-stripped) main-data stream ---- */
typedef struct { uint8_t *d; uint64_t pos; } bitwr;
static inline void put(bitwr *w, uint32_t v, unsigned n)
{
  while (n) {
    unsigned room = 8 - (unsigned)(w->pos & 7), k = n < room ? n : room;
    uint32_t bits = (v >> (n - k)) & ((1u << k) - 1);
    w->d[w->pos >> 3] |= (uint8_t)(bits << (room - k));
    w->pos += k; n -= k;
  }
}

/* encoder tables: per book, (x,y) -> code/len */
typedef struct { uint32_t code[256]; uint8_t len[256]; int maxv; } encbook;
static encbook g_enc[20];
static int g_enc_ready;
static void enc_init(void)
{
  if (g_enc_ready) return;
  for (int b = 0; b < p3_book_count(); b++) {
    const p3_hcode *c; int n = p3_book_codes(b, &c);
    g_enc[b].maxv = 0;
    for (int k = 0; k < n; k++) {
      g_enc[b].code[c[k].x * 16 + c[k].y] = c[k].code; g_enc[b].len[c[k].x * 16 + c[k].y] = c[k].len;
      if (c[k].x > g_enc[b].maxv) g_enc[b].maxv = c[k].x;
    }
  }
  g_enc_ready = 1;
}

static const uint16_t k_bitrate[15] = {0,32,40,48,56,64,80,96,112,128,160,192,224,256,320};
static const uint32_t k_sfreq[3] = {44100, 48000, 32000};

typedef struct {
  unsigned p23l, bigv, gain, sfc, ws, bt, mixed, ts[3], sbg[3], r0, r1, pre, scale, c1t;
} gcinfo;

/* geometric-ish magnitude in [0, maxv] with mean around `mu` */
static inline int draw_mag(rng_t *r, double mu, int maxv)
{
  if (mu <= 0.02) return rndu(r) < mu * 8 ? 1 <= maxv ? 1 : 0 : 0;
  double u = rndu(r); if (u < 1e-12) u = 1e-12;
  int v = (int)(-log(u) * mu);
  return v > maxv ? maxv : v;
}

static unsigned pair_bits(int table, int x, int y)
{
  const p3_tables *T = p3_tables_get();
  int b = T->table_book[table]; if (b < 0) return 0;
  unsigned lb = T->table_linbits[table], n;
  int cx = x > 15 ? 15 : x, cy = y > 15 ? 15 : y;
  n = g_enc[b].len[cx * 16 + cy];
  if (lb && cx == 15) n += lb;
  if (x) n++;
  if (lb && cy == 15) n += lb;
  if (y) n++;
  return n;
}
static void pair_put(bitwr *w, int table, int x, int sx, int y, int sy)
{
  const p3_tables *T = p3_tables_get();
  int b = T->table_book[table]; if (b < 0) return;
  unsigned lb = T->table_linbits[table];
  int cx = x > 15 ? 15 : x, cy = y > 15 ? 15 : y;
  put(w, g_enc[b].code[cx * 16 + cy], g_enc[b].len[cx * 16 + cy]);
  if (lb && cx == 15) put(w, (uint32_t)(x - 15), lb);
  if (x) put(w, (uint32_t)sx, 1);
  if (lb && cy == 15) put(w, (uint32_t)(y - 15), lb);
  if (y) put(w, (uint32_t)sy, 1);
}

/* largest value a big_values table can carry */
static int table_max(int t)
{
  const p3_tables *T = p3_tables_get();
  int b = T->table_book[t]; if (b < 0) return 0;
  int m = g_enc[b].maxv;
  if (T->table_linbits[t]) m = 15 + (1 << T->table_linbits[t]) - 1;
  return m;
}

/* Encode one granule-channel into w.  Returns bits written (= part2_3_length unless overrun). */
static void encode_gc(rng_t *r, const p3_synth_cfg *cfg, bitwr *w, gcinfo *g, unsigned budget,
                      unsigned gr, unsigned ch, unsigned scfsi, int intensity_ch0, int sf,
                      uint8_t scf_l_gr0[21], int16_t *is_out, unsigned *count1_out, int *maxabs_out)
{
  const p3_tables *T = p3_tables_get();
  uint64_t start = w->pos;
  unsigned slen1 = T->slen[g->sfc][0], slen2 = T->slen[g->sfc][1];
  int is_short = g->ws && g->bt == 2;
  /* ---- part 2: scalefactors (layouts of pdmp3.c:1382-1435) ---- */
  if (is_short) {
    unsigned sfb0 = 0;
    if (g->mixed) { for (unsigned sfb = 0; sfb < 8; sfb++) put(w, sfb == 0 ? 0 : rndn(r, 1u << slen1), slen1); sfb0 = 3; }
    for (unsigned sfb = sfb0; sfb < 12; sfb++) {
      unsigned nb = sfb < 6 ? slen1 : slen2;
      for (unsigned win = 0; win < 3; win++) {
        unsigned v = (sfb == 0) ? 0 : rndn(r, 1u << nb);
        if (intensity_ch0 && v > 7) v &= 7;
        put(w, v, nb);
      }
    }
  } else {
    static const unsigned lo[5] = {0, 6, 11, 16, 21};
    for (unsigned band = 0; band < 4; band++) {
      if (gr == 1 && ((scfsi >> band) & 1)) continue;          /* reused from granule 0 */
      unsigned nb = band < 2 ? slen1 : slen2;
      for (unsigned sfb = lo[band]; sfb < lo[band + 1]; sfb++) {
        unsigned v = (sfb == 0) ? 0 : rndn(r, 1u << nb);
        if (intensity_ch0 && v > 7) v &= 7;                     /* G7 */
        put(w, v, nb);
        if (gr == 0 && scf_l_gr0) scf_l_gr0[sfb] = (uint8_t)v;
      }
    }
  }
  /* ---- part 3: Huffman ---- */
  unsigned r1s, r2s;
  if (is_short) { r1s = 36; r2s = 576; }
  else { r1s = T->sfb_l[sf][g->r0 + 1]; r2s = T->sfb_l[sf][g->r0 + g->r1 + 2]; }
  /* G2: short (gr1,ch1) keeps count1 <= 3*s[12]; G3: preflag only if count1 <= l[21] */
  unsigned line_cap = 576;
  if (is_short && gr == 1 && ch == 1) line_cap = 3u * T->sfb_s[sf][12] - 48;
  if (g->pre) line_cap = T->sfb_l[sf][21] - 48;
  int16_t tmp[580]; memset(tmp, 0, sizeof tmp);
  unsigned used = (unsigned)(w->pos - start);
  unsigned bv_budget = budget > used ? (unsigned)((budget - used) * (0.70 + 0.25 * rndu(r))) : 0;
  unsigned pos = 0, bits = 0;
  double tilt = 0.004 + 0.012 * rndu(r);
  double mu_scale[3];
  for (int k = 0; k < 3; k++) {
    int tm = table_max((int)g->ts[k]);
    mu_scale[k] = tm <= 1 ? 0.35 : tm <= 3 ? 0.8 : tm <= 7 ? 1.6 : tm <= 15 ? 3.0 : 3.0 + 0.02 * (tm > 600 ? 600 : tm);
  }
  while (pos + 2 <= line_cap && pos + 2 <= 576) {
    int reg = pos < r1s ? 0 : pos < r2s ? 1 : 2;
    int t = (int)g->ts[reg], tm = table_max(t);
    double mu = mu_scale[reg] * exp(-tilt * pos);
    int x = draw_mag(r, mu, tm), y = draw_mag(r, mu, tm);
    /* rare big escapes so linbits paths are exercised without constant clipping */
    if (tm > 15 && rndn(r, 400) == 0) x = 15 + (int)rndn(r, (uint32_t)(tm - 14));
    unsigned nb = pair_bits(t, x, y);
    if (bits + nb > bv_budget) break;
    int sx = (int)rndn(r, 2), sy = (int)rndn(r, 2);
    pair_put(w, t, x, sx, y, sy);
    tmp[pos] = (int16_t)(sx ? -x : x); tmp[pos + 1] = (int16_t)(sy ? -y : y);
    bits += nb; pos += 2;
  }
  g->bigv = pos / 2;
  /* count1 region */
  int bq = g->c1t ? 16 : 15;
  unsigned last_len = 0;
  double p1 = 0.15 + 0.5 * rndu(r);
  while (pos + 4 <= line_cap && pos <= 572) {
    unsigned q = 0, nb;
    for (int k = 0; k < 4; k++) q = (q << 1) | (rndu(r) < p1 ? 1u : 0u);
    nb = g_enc[bq].len[q] + (unsigned)__builtin_popcount(q);
    if ((unsigned)(w->pos - start) + nb > budget) break;
    put(w, g_enc[bq].code[q], g_enc[bq].len[q]);
    for (int k = 3; k >= 0; k--) if ((q >> k) & 1) { unsigned s = rndn(r, 2); put(w, s, 1); tmp[pos + 3 - k] = (int16_t)(s ? -1 : 1); }
    last_len = nb; pos += 4;
    if (rndn(r, 64) == 0) break;
  }
  unsigned total = (unsigned)(w->pos - start);
  unsigned count1 = pos;
  if (total == 0) { put(w, 1, 1); total = 1; g->bigv = 0; count1 = 4; }   /* G5: part2_3_length > 0; '1' is the count1-A code of 0000 */
  /* deliberate overrun: declare part2_3_length 1..last_len-1 bits short (exercises pdmp3.c:2105-2106) */
  if (last_len > 1 && cfg->overrun_pm && rndn(r, 1000) < (uint32_t)cfg->overrun_pm && !g->c1t) {
    unsigned cut = 1 + rndn(r, last_len - 1);
    /* rewind the writer: clear the cut bits so the next part starts there */
    for (unsigned k = 0; k < cut; k++) { uint64_t p = w->pos - 1 - k; w->d[p >> 3] &= (uint8_t)~(0x80u >> (p & 7)); }
    w->pos -= cut; total -= cut;
    count1 = 0xffff;   /* unknown: depends on the following bits */
  }
  g->p23l = total;
  if (is_out) memcpy(is_out, tmp, 576 * sizeof(int16_t));
  { int m = 0; for (int i = 0; i < 576; i++) { int a = tmp[i] < 0 ? -tmp[i] : tmp[i]; if (a > m) m = a; } *maxabs_out = m; }
  *count1_out = count1;
}

int64_t p3_synth(const p3_synth_cfg *cfg, int64_t n_frames, uint8_t *out, uint64_t cap, int16_t *is_out)
{
  const p3_tables *T = p3_tables_get();
  enc_init();
  rng_t rg = {cfg->seed * 0x2545f4914f6cdd1dull + 12345};
  int nch = cfg->mode == 3 ? 1 : 2, sf = cfg->sfreq, silen = nch == 1 ? 17 : 32;
  unsigned hdr = 4 + (cfg->crc ? 2 : 0);
  /* pass 1: frame sizes */
  uint32_t *fsize = malloc(sizeof(uint32_t) * (size_t)n_frames);
  uint8_t *bri = malloc((size_t)n_frames), *padb = malloc((size_t)n_frames);
  uint64_t *mpos = malloc(sizeof(uint64_t) * (size_t)(n_frames + 1));
  uint64_t rem = 0, total_main = 0;
  for (int64_t f = 0; f < n_frames; f++) {
    unsigned br = cfg->bitrate_index ? (unsigned)cfg->bitrate_index : 1 + rndn(&rg, 14);
    /* padding so that the mean frame length is exact for CBR (ISO 2.4.2.3 rule) */
    unsigned num = 144u * k_bitrate[br] * 1000u, base = num / k_sfreq[sf];
    rem += num % k_sfreq[sf];
    unsigned pad = 0; if (rem >= k_sfreq[sf]) { pad = 1; rem -= k_sfreq[sf]; }
    bri[f] = (uint8_t)br; padb[f] = (uint8_t)pad; fsize[f] = base + pad;
    mpos[f] = total_main; total_main += fsize[f] - hdr - silen;
  }
  mpos[n_frames] = total_main;
  uint8_t *logical = calloc(total_main + 64, 1);
  uint8_t *side = calloc((size_t)n_frames * 32 + 32, 1);
  uint8_t *modes = malloc((size_t)n_frames * 2);
  if (!logical || !side) return -2;
  bitwr w = {logical, 0};
  int bstate[2] = {0, 0};
  uint8_t scf0[2][21];
  for (int64_t f = 0; f < n_frames; f++) {
    unsigned main_size = fsize[f] - hdr - silen;
    uint64_t cur = (w.pos + 7) >> 3;                      /* byte-align the start of the frame's data */
    uint64_t earliest = mpos[f] > 511 ? mpos[f] - 511 : 0;
    if (!cfg->reservoir) cur = mpos[f];                   /* previous data never passes mpos[f] */
    else if (cur < earliest) cur = earliest;
    w.pos = cur * 8;
    unsigned main_begin = (unsigned)(mpos[f] - cur);
    uint64_t avail_bits = (mpos[f] + main_size - cur) * 8;
    double fill = cfg->fill_pm / 1000.0 * (0.55 + 0.9 * rndu(&rg));
    uint64_t target = (uint64_t)(main_size * 8.0 * fill);
    if (target > avail_bits) target = avail_bits;
    unsigned mode_ext = 0;
    if (cfg->mode == 1) mode_ext = cfg->mode_ext >= 0 ? (unsigned)cfg->mode_ext : rndn(&rg, 4);
    /* block types per granule / channel */
    gcinfo gi[2][2]; memset(gi, 0, sizeof gi);
    unsigned scfsi[2] = {0, 0};
    for (unsigned gr = 0; gr < 2; gr++) {
      for (int ch = 0; ch < nch; ch++) {
        gcinfo *g = &gi[gr][ch];
        int st;
        if (ch == 1 && cfg->mode == 1) st = (int)gi[gr][0].bt;                 /* joint stereo: same block type */
        else if (!cfg->blocks) st = 0;
        else {
          int s0 = bstate[ch];
          if (s0 == 0) st = rndn(&rg, 8) == 0 ? 1 : 0;
          else if (s0 == 1) st = 2;
          else if (s0 == 2) st = rndn(&rg, 2) ? 2 : 3;
          else st = 0;
        }
        bstate[ch] = st;
        g->bt = (unsigned)st; g->ws = st != 0;
        g->mixed = (st == 2) ? (ch == 1 && cfg->mode == 1 ? gi[gr][0].mixed : rndn(&rg, 2)) : 0;
      }
    }
    /* G6: no intensity stereo when channel 0 of a granule is short */
    if (!cfg->iso && (mode_ext & 1) && ((gi[0][0].bt == 2) || (gi[1][0].bt == 2))) mode_ext &= 2;
    for (int ch = 0; ch < nch; ch++)
      if (gi[0][ch].bt != 2 && gi[1][ch].bt != 2 && cfg->scalefacs) scfsi[ch] = rndn(&rg, 4) == 0 ? rndn(&rg, 16) : 0;
    uint64_t frame_start_bit = w.pos;
    for (unsigned gr = 0; gr < 2; gr++) for (int ch = 0; ch < nch; ch++) {
      gcinfo *g = &gi[gr][ch];
      unsigned idx = gr * (unsigned)nch + (unsigned)ch, left = 2u * (unsigned)nch - idx;
      uint64_t spent = w.pos - frame_start_bit;
      uint64_t remain = target > spent ? target - spent : 0;
      unsigned budget = (unsigned)(remain / left * (left > 1 ? 0.7 + 0.6 * rndu(&rg) : 1.0));
      if (budget > remain) budget = (unsigned)remain;
      if (budget < 48) budget = 48;
      {   /* never starve the parts still to come (G5: every part2_3_length > 0) */
        uint64_t keep = 48ull * (left - 1), room = avail_bits > spent + keep ? avail_bits - spent - keep : 0;
        if (budget > room) budget = (unsigned)room;
      }
      if (budget > 4000) budget = 4000;
      if (cfg->iso && rndn(&rg, 1000) < 30) budget = 0;       /* an empty part */
      int intensity = (mode_ext & 1) && cfg->mode == 1;
      /* side-info fields */
      g->gain = (unsigned)(cfg->gain + (int)rndn(&rg, 13) - 6);
      if (cfg->scalefacs && budget >= 300) g->sfc = rndn(&rg, 16);
      else g->sfc = 0;
      g->scale = cfg->scalefacs ? rndn(&rg, 2) : 0;
      g->pre = (cfg->scalefacs && !(g->ws && g->bt == 2)) ? (rndn(&rg, 4) == 0) : 0;
      g->c1t = cfg->count1_b_pm && rndn(&rg, 1000) < (uint32_t)cfg->count1_b_pm;
      /* the reference decodes table B quads from 2 sign bits only (Q1), so count1 cannot be bounded
       * by the encoder: keep such granules away from the G2/G3 limits */
      if (g->c1t) { g->pre = 0; if (g->ws && g->bt == 2 && gr == 1 && ch == 1) g->c1t = 0; }
      int mt = cfg->max_table > 0 ? cfg->max_table : 31;
      for (int k = 0; k < 3; k++) {
        int t;
        do { t = (int)rndn(&rg, (uint32_t)mt + 1); } while (t == 4 || t == 14 || (t == 0 && rndn(&rg, 4)));
        g->ts[k] = (unsigned)t;
      }
      if (g->ws) {
        g->ts[2] = 0;
        for (int k = 0; k < 3; k++) g->sbg[k] = g->bt == 2 ? rndn(&rg, 8) >> rndn(&rg, 3) : 0;
        g->r0 = (g->bt == 2 && !g->mixed) ? 8 : 7; g->r1 = 20 - g->r0;
      } else {
        g->r0 = rndn(&rg, 16); g->r1 = rndn(&rg, 8);
        while (g->r0 + g->r1 + 2 > 22) { if (g->r0) g->r0--; if (g->r0 + g->r1 + 2 > 22 && g->r1) g->r1--; }
      }
      unsigned c1 = 0; int mx = 0;
      if (budget == 0) { g->p23l = 0; g->bigv = 0; }        /* cannot happen with sane fill; kept for safety (violates G5) */
      else encode_gc(&rg, cfg, &w, g, budget, gr, (unsigned)ch, scfsi[ch], intensity && ch == (cfg->iso ? 1 : 0), sf,
                     gr == 0 ? scf0[ch] : NULL, is_out ? is_out + (((size_t)f * 2 + gr) * 2 + (size_t)ch) * 576 : NULL, &c1, &mx);
      if (mx > 1) {   /* keep the loudest line below ~0.3 of full scale so that clipping stays rare */
        int lim = 210 + (int)floor(4.0 * log2((cfg->peak_pm > 0 ? cfg->peak_pm * 1e-3 : 0.30) / pow((double)mx, 4.0 / 3.0)));
        if ((int)g->gain > lim) g->gain = (unsigned)(lim < 0 ? 0 : lim);
      }
    }
    /* side info bits */
    uint8_t *s = side + (size_t)f * 32; bitwr sw = {s, 0};
    put(&sw, main_begin, 9); put(&sw, 0, nch == 1 ? 5 : 3);
    for (int ch = 0; ch < nch; ch++) for (int b = 0; b < 4; b++) put(&sw, (scfsi[ch] >> b) & 1, 1);
    for (unsigned gr = 0; gr < 2; gr++) for (int ch = 0; ch < nch; ch++) {
      gcinfo *g = &gi[gr][ch];
      put(&sw, g->p23l, 12); put(&sw, g->bigv, 9); put(&sw, g->gain, 8); put(&sw, g->sfc, 4); put(&sw, g->ws, 1);
      if (g->ws) { put(&sw, g->bt, 2); put(&sw, g->mixed, 1); put(&sw, g->ts[0], 5); put(&sw, g->ts[1], 5);
                   put(&sw, g->sbg[0], 3); put(&sw, g->sbg[1], 3); put(&sw, g->sbg[2], 3); }
      else { put(&sw, g->ts[0], 5); put(&sw, g->ts[1], 5); put(&sw, g->ts[2], 5); put(&sw, g->r0, 4); put(&sw, g->r1, 3); }
      put(&sw, g->pre, 1); put(&sw, g->scale, 1); put(&sw, g->c1t, 1);
    }
    modes[2 * f] = (uint8_t)cfg->mode; modes[2 * f + 1] = (uint8_t)mode_ext;
  }
  /* pass 3: assemble raw frames */
  uint64_t o = 0;
  rng_t rj = {cfg->seed ^ 0xabcdef};
  for (int64_t f = 0; f < n_frames; f++) {
    if (cfg->garbage_pm && rndn(&rj, 1000) < (uint32_t)cfg->garbage_pm) {
      unsigned k = 1 + rndn(&rj, 40);
      if (o + k > cap) { o = (uint64_t)-1; break; }
      for (unsigned i = 0; i < k; i++) { uint8_t b = (uint8_t)rnd(&rj); out[o++] = b == 0xff ? 0xfe : b; }
    }
    unsigned main_size = fsize[f] - hdr - silen;
    if (o + fsize[f] > cap) { o = (uint64_t)-1; break; }
    out[o] = 0xff; out[o + 1] = (uint8_t)(0xfa | (cfg->crc ? 0 : 1));
    out[o + 2] = (uint8_t)(bri[f] << 4 | sf << 2 | padb[f] << 1);
    out[o + 3] = (uint8_t)(modes[2 * f] << 6 | modes[2 * f + 1] << 4 | 0x4);
    if (cfg->crc) { out[o + 4] = (uint8_t)rnd(&rj); out[o + 5] = (uint8_t)rnd(&rj); }
    memcpy(out + o + hdr, side + (size_t)f * 32, (size_t)silen);
    memcpy(out + o + hdr + silen, logical + mpos[f], main_size);
    o += fsize[f];
  }
  free(fsize); free(bri); free(padb); free(mpos); free(logical); free(side); free(modes);
  (void)T;
  return (int64_t)o;
}
