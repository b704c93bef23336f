/* p3_tables.h -- constant tables of the MPEG-1 Layer III granule decode path.
 *
 * Plain C, host side.  Every floating-point table is REGENERATED here from a formula that
 * reproduces the reference's values bit for bit; nothing is pasted from
 * the reference.  tests/test_cpu_tables.py compares each table with the compiled reference
 * (oracle/_ref/libref_taps.so).  The same struct is uploaded verbatim to the device.
 *
 * Reference data being reproduced (file:line in /root/reference/pdmp3.c):
 *   cs/ca 573-574, is_ratios 575, g_imdct_win 577-603, cos_N12 606-619, cos_N36 620-729,
 *   g_synth_dtbl 740-870, g_synth_n_win 1989-1993, powtab34 973-981, pretab 2123,
 *   mpeg1_scalefac_sizes 530-533, g_sf_band_indices 879-892, Huffman books 235-570.
 */
#ifndef P3_TABLES_H
#define P3_TABLES_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define P3_HLUT_PBITS 8      /* first-level width of the Huffman decode LUT */
#define P3_HLUT_SBITS 6      /* width of every further level */
#define P3_HLUT_MAX   5120   /* u16 entries, all books together (4504 used at 8/6) */
#define P3_T2_BIAS    266    /* t2[q + P3_T2_BIAS] = 2^(q/4), q = global_gain-210-8*subblock_gain in [-266,45] */

typedef struct { uint8_t len; uint32_t code; uint8_t x, y; } p3_hcode;

/* Huffman LUT entry (u16):
 *   leaf: bit15=0, bit14 = x or y is 15 (an escape if the table has linbits), bits12..8 = code bits consumed AT THIS LEVEL,
 *         bits7..4 = x, bits3..0 = y
 *   link: bit15=1, bits12..10 = width w of the next level, bits9..0 = offset of the next level
 *         relative to the book base; the caller consumes this level's full width and indexes
 *         the next level with the following w bits. */
typedef struct {
  float cs[8], ca[8];
  float is_l[8], is_r[8];          /* intensity ratios for is_pos 0..6 (7 unused): pdmp3.c:2167-2173 */
  float imdct_win[4][36];
  float cos12[6][12];
  float cos36[18][36];
  float synth_d[512];
  float synth_n[64][32];
  float pow43[8208];
  float t1h[40];                   /* 2^(-e/2), e = (scale?2:1)*(scalefac+preflag*pretab) */
  float t2[320];                   /* 2^(q/4) */
  uint16_t sfb_l[3][24];           /* [sfreq][0..22] */
  uint16_t sfb_s[3][16];           /* [sfreq][0..13] */
  uint8_t  pretab[24];             /* [0..20]; 21 = pseudo band (ISO: 0) */
  uint8_t  slen[16][2];
  /* per-line helpers, [sfreq][line] */
  uint8_t  line_sfb_l[3][576];     /* long-block sfb of a line (21 = lines >= l[21]) */
  uint8_t  line_sfbw_s[3][576];    /* short-block (sfb | win<<4) of a line in BITSTREAM order */
  uint16_t reorder_src[3][576];    /* short blocks: output line d takes bitstream line reorder_src[d]
                                      (pdmp3.c:1811-1816); mixed blocks use it for d >= 36 only */
  /* Huffman */
  uint16_t hlut[P3_HLUT_MAX];
  uint16_t book_base[20];
  uint8_t  book_pbits[20];
  int8_t   table_book[36];         /* -1: empty table (0, 4, 14) */
  uint8_t  table_linbits[36];
  uint32_t hlut_used;
  uint32_t hlut_zero;              /* two all-zero leaves: the code book of the empty tables 0/4/14 */
} p3_tables;

const p3_tables *p3_tables_get(void);

/* canonical code lists, for encoders / exhaustive tests */
int p3_book_count(void);
int p3_book_codes(int book, const p3_hcode **codes);

#ifdef __cplusplus
}
#endif
#endif
