/* p3_tables.c -- builds p3_tables once per process (see p3_tables.h). Plain C. */
#include "p3_tables.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "p3_huffcodes.inc"
#include "p3_synwin.inc"

static p3_tables g_t;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

/* "print with %f, parse back as float" -- how the reference's 6-decimal literals came to be */
static float r6(double v) { char b[64]; snprintf(b, sizeof b, "%f", v); return strtof(b, NULL); }

/* ISO 11172-3 Table B.8 scalefactor band boundaries (44.1, 48, 32 kHz) */
static const uint16_t k_sfb_l[3][23] = {
  {0,4,8,12,16,20,24,30,36,44,52,62,74,90,110,134,162,196,238,288,342,418,576},
  {0,4,8,12,16,20,24,30,36,42,50,60,72,88,106,128,156,190,230,276,330,384,576},
  {0,4,8,12,16,20,24,30,36,44,54,66,82,102,126,156,194,240,296,364,448,550,576}};
static const uint16_t k_sfb_s[3][14] = {
  {0,4,8,12,16,22,30,40,52,66,84,106,136,192},
  {0,4,8,12,16,22,28,38,50,64,80,100,126,192},
  {0,4,8,12,16,22,30,42,58,78,104,138,180,192}};
static const uint8_t k_pretab[21] = {0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,3,2};
static const uint8_t k_slen[16][2] = {{0,0},{0,1},{0,2},{0,3},{3,0},{1,1},{1,2},{1,3},
                                      {2,1},{2,2},{2,3},{3,1},{3,2},{3,3},{4,2},{4,3}};

/* ---- Huffman LUT construction from the canonical code list ------------------------------ */
static uint32_t g_fill;                /* next free hlut entry */
static uint16_t g_base;                /* base of the book under construction */

/* Fill one level: entries for all codes below `prefix` (prefix_len bits) using `width` bits. */
static void build_level(const p3_hcode *c, int n, int prefix_len, uint32_t prefix, int width, uint32_t at)
{
  uint32_t size = 1u << width;
  for (uint32_t i = 0; i < size; i++) g_t.hlut[at + i] = 0;   /* len 0 = invalid (cannot happen: books are complete) */
  for (int k = 0; k < n; k++) {
    int l = c[k].len;
    if (l <= prefix_len || (c[k].code >> (l - prefix_len)) != prefix) continue;
    int rem = l - prefix_len;
    if (rem <= width) {                 /* leaf, replicated over the don't-care bits */
      uint32_t lo = (c[k].code & ((1u << rem) - 1)) << (width - rem);
      for (uint32_t j = 0; j < (1u << (width - rem)); j++)
        g_t.hlut[at + lo + j] = (uint16_t)((rem << 8) | (c[k].x << 4) | c[k].y | ((c[k].x == 15 || c[k].y == 15) ? 0x4000 : 0));
    }
  }
  /* links: group the longer codes by their next `width` bits */
  for (uint32_t p2 = 0; p2 < size; p2++) {
    int maxrem = 0;
    uint32_t full = (prefix << width) | p2;
    for (int k = 0; k < n; k++) {
      int l = c[k].len;
      if (l > prefix_len + width && (c[k].code >> (l - prefix_len - width)) == full)
        if (l - prefix_len - width > maxrem) maxrem = l - prefix_len - width;
    }
    if (!maxrem) continue;
    int w = maxrem < P3_HLUT_SBITS ? maxrem : P3_HLUT_SBITS;
    uint32_t sub = g_fill; g_fill += 1u << w;
    if (g_fill > P3_HLUT_MAX || sub - g_base > 1023) { fprintf(stderr, "p3_tables: hlut overflow\n"); abort(); }
    g_t.hlut[at + p2] = (uint16_t)(0x8000u | (w << 10) | (sub - g_base));
    build_level(c, n, prefix_len + width, full, w, sub);
  }
}

static void build_huffman(void)
{
  g_fill = 0;
  for (int b = 0; b < P3_NBOOKS; b++) {
    const p3_hcode *c = p3_books[b].codes; int n = p3_books[b].n, maxlen = 0;
    for (int k = 0; k < n; k++) if (c[k].len > maxlen) maxlen = c[k].len;
    int pb = maxlen < P3_HLUT_PBITS ? maxlen : P3_HLUT_PBITS;
    g_t.book_base[b] = g_base = (uint16_t)g_fill;
    g_t.book_pbits[b] = (uint8_t)pb;
    g_fill += 1u << pb;
    build_level(c, n, 0, 0, pb, g_base);
  }
  /* pseudo book of the empty tables 0/4/14 (pdmp3.c:1599-1602): whatever bit comes next, a leaf of length 0 with x = y = 0 */
  g_t.hlut_zero = g_fill; g_t.hlut[g_fill] = g_t.hlut[g_fill + 1] = 0; g_fill += 2;
  g_t.hlut_used = g_fill;
  for (int t = 0; t < 34; t++) { g_t.table_book[t] = p3_table_book[t]; g_t.table_linbits[t] = p3_table_linbits[t]; }
}

static void build(void)
{
  const float pif = (float)3.14159265358979323846;
  const double PI = 3.14159265358979323846;
  static const double ci[8] = {-0.6,-0.535,-0.33,-0.185,-0.095,-0.041,-0.0142,-0.0037};
  int i, j, m, p, sf;

  for (i = 0; i < 8; i++) {
    g_t.cs[i] = r6(1.0 / sqrt(1.0 + ci[i] * ci[i]));
    g_t.ca[i] = r6(ci[i] / sqrt(1.0 + ci[i] * ci[i]));
  }
  /* intensity ratios (pdmp3.c:575, 2167-2173): float arithmetic on the 6-decimal tan table */
  for (i = 0; i < 8; i++) {
    if (i < 6) {
      float r = r6(tan(i * PI / 12.0));
      g_t.is_l[i] = r / (1.0f + r);
      g_t.is_r[i] = 1.0f / (1.0f + r);
    } else { g_t.is_l[i] = 1.0f; g_t.is_r[i] = 0.0f; }   /* is_pos 6; 7 never used */
  }
  /* IMDCT windows, layout of pdmp3.c:1657-1667; the 6-decimal literals carry a float-pi argument */
  {
    float a36 = pif / 36, a12 = pif / 12;
    memset(g_t.imdct_win, 0, sizeof g_t.imdct_win);
    for (i = 0; i < 36; i++) g_t.imdct_win[0][i] = r6(sin((double)a36 * (i + 0.5)));
    for (i = 0; i < 18; i++) g_t.imdct_win[1][i] = r6(sin((double)a36 * (i + 0.5)));
    for (i = 18; i < 24; i++) g_t.imdct_win[1][i] = 1.0f;
    for (i = 24; i < 30; i++) g_t.imdct_win[1][i] = r6(sin((double)a12 * (i + 0.5 - 18.0)));
    for (i = 0; i < 12; i++) g_t.imdct_win[2][i] = r6(sin((double)a12 * (i + 0.5)));
    for (i = 6; i < 12; i++) g_t.imdct_win[3][i] = r6(sin((double)a12 * (i + 0.5 - 6.0)));
    for (i = 12; i < 18; i++) g_t.imdct_win[3][i] = 1.0f;
    for (i = 18; i < 36; i++) g_t.imdct_win[3][i] = r6(sin((double)a36 * (i + 0.5)));
  }
  for (m = 0; m < 6; m++) for (p = 0; p < 12; p++)
    g_t.cos12[m][p] = r6(cos((double)(pif / 24) * ((2 * p + 7) * (2 * m + 1))));
  for (m = 0; m < 18; m++) for (p = 0; p < 36; p++)
    g_t.cos36[m][p] = r6(cos((double)(pif / 72) * ((2 * p + 19) * (2 * m + 1))));
  for (i = 0; i < 512; i++) {
    char b[64]; snprintf(b, sizeof b, "%.9f", p3_synwin_k[i] / 65536.0);
    g_t.synth_d[i] = strtof(b, NULL);
  }
  for (i = 0; i < 64; i++) for (j = 0; j < 32; j++)      /* pdmp3.c:1992: float product, double cos */
    g_t.synth_n[i][j] = (float)cos((double)((float)((16 + i) * (2 * j + 1))) * (PI / 64.0));
  for (i = 0; i < 8208; i++) g_t.pow43[i] = (float)pow((double)(float)i, 4.0 / 3.0);
  for (i = 0; i < 40; i++) g_t.t1h[i] = (float)pow(2.0, -0.5 * i);
  for (i = 0; i < 320; i++) g_t.t2[i] = (float)pow(2.0, 0.25 * (i - P3_T2_BIAS));

  memset(g_t.pretab, 0, sizeof g_t.pretab); memcpy(g_t.pretab, k_pretab, 21);
  memcpy(g_t.slen, k_slen, sizeof k_slen);
  for (sf = 0; sf < 3; sf++) {
    memcpy(g_t.sfb_l[sf], k_sfb_l[sf], sizeof k_sfb_l[sf]);
    memcpy(g_t.sfb_s[sf], k_sfb_s[sf], sizeof k_sfb_s[sf]);
    for (int sfb = 0; sfb < 22; sfb++)
      for (i = k_sfb_l[sf][sfb]; i < k_sfb_l[sf][sfb + 1]; i++) g_t.line_sfb_l[sf][i] = (uint8_t)sfb;
    for (int sfb = 0; sfb < 13; sfb++) {
      int start = 3 * k_sfb_s[sf][sfb], wl = k_sfb_s[sf][sfb + 1] - k_sfb_s[sf][sfb];
      for (int w = 0; w < 3; w++) for (j = 0; j < wl; j++) {
        int src = start + w * wl + j, dst = start + 3 * j + w;
        g_t.line_sfbw_s[sf][src] = (uint8_t)(sfb | (w << 4));
        g_t.reorder_src[sf][dst] = (uint16_t)src;
      }
    }
  }
  build_huffman();
}

const p3_tables *p3_tables_get(void) { pthread_once(&g_once, build); return &g_t; }
int p3_book_count(void) { return P3_NBOOKS; }
int p3_book_codes(int book, const p3_hcode **codes) { *codes = p3_books[book].codes; return p3_books[book].n; }
