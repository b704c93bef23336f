/* p3_parse.c -- host side, plain C: frame sync, header, side info, bit-reservoir bookkeeping.
 *
 * Re-implements (does not copy) the frame layer of the reference for a whole buffer at once:
 *   Search_Header/Read_Header  pdmp3.c:1322-1340, 1252-1320   -> hop()
 *   Read_CRC                   pdmp3.c:1206-1210              -> CRC word skipped, not verified
 *   Read_Audio_L3              pdmp3.c:1129-1200              -> parse_side()
 *   Get_Main_Data              pdmp3.c:1096-1122              -> main_pos prefix sums + `top` rule
 * The sequential part is only the header hop (frame length depends on the header); side info
 * is parsed by a pool of threads.  Output: SoA descriptors for the device (pdmp3_b200.h).
 */
#include "../../include/pdmp3_b200.h"
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

static const uint16_t k_bitrate[15] = {0,32,40,48,56,64,80,96,112,128,160,192,224,256,320}; /* pdmp3.c:524-527 */
static const uint32_t k_sfreq[3] = {44100, 48000, 32000};                                  /* pdmp3.c:529 */

/* valid MPEG-1 Layer III header at p?  (checks of pdmp3.c:1264, 1293-1315, 1329) */
static inline int header_ok(const uint8_t *p)
{
  if (p[0] != 0xff || (p[1] & 0xf0) != 0xf0) return 0;       /* 12-bit sync */
  if (!(p[1] & 0x08)) return 0;                              /* id must be 1 (MPEG-1) */
  if (((p[1] >> 1) & 3) != 1) return 0;                      /* layer field 01 = Layer III */
  unsigned br = p[2] >> 4, sf = (p[2] >> 2) & 3;
  if (br == 0 || br == 15 || sf == 3) return 0;
  return 1;
}

/* Search_Header without consuming (pdmp3.c:1322-1340): 1 = header found, 0 = need more data,
 * -1 = no valid header within 1152 bytes of garbage. */
int p3_find_header(const uint8_t *data, uint64_t n, int *nch, int *sfreq)
{
  uint64_t lim = n < 4 ? 0 : (n - 3 < 1153 ? n - 3 : 1153);
  for (uint64_t p = 0; p < lim; p++)
    if (header_ok(data + p)) { *nch = (data[p + 3] >> 6) == 3 ? 1 : 2; *sfreq = (data[p + 2] >> 2) & 3; return 1; }
  return lim == 1153 ? -1 : 0;
}

typedef struct { const uint8_t *d; unsigned pos; } bitrd;
static inline unsigned getbits(bitrd *b, unsigned n)     /* MSB first, n <= 16 (pdmp3.c:1547-1561) */
{
  unsigned byte = b->pos >> 3, sh = b->pos & 7;
  uint32_t w = ((uint32_t)b->d[byte] << 24) | ((uint32_t)b->d[byte + 1] << 16) | ((uint32_t)b->d[byte + 2] << 8);
  b->pos += n;
  return (w << sh) >> (32 - n);
}

static const uint8_t k_slen[16][2] = {{0,0},{0,1},{0,2},{0,3},{3,0},{1,1},{1,2},{1,3},{2,1},{2,2},{2,3},{3,1},{3,2},{3,3},{4,2},{4,3}};   /* pdmp3.c:530-533 */

/* Bits of part 2 (scalefactors) that Read_Main_L3 reads for a granule-channel (pdmp3.c:1379-1435).  Needed for one case
 * only: a part with part2_3_length == 0.  The reference still reads those bits and Read_Huffman returns without
 * Set_Main_Pos (pdmp3.c:2057-2061), so the NEXT part of the frame starts behind them instead of at the prefix sum. */
static unsigned part2_bits(unsigned sfc, unsigned ws, unsigned bt, unsigned mixed, unsigned gr, unsigned scfsi4)
{
  const unsigned s1 = k_slen[sfc][0], s2 = k_slen[sfc][1];
  if (ws && bt == 2) return mixed ? 17 * s1 + 18 * s2 : 18 * s1 + 18 * s2;
  if (gr == 0) scfsi4 = 0;
  return ((scfsi4 & 1) ? 0 : 6 * s1) + ((scfsi4 & 2) ? 0 : 5 * s1) + ((scfsi4 & 4) ? 0 : 5 * s2) + ((scfsi4 & 8) ? 0 : 5 * s2);
}

/* side info of one frame -> 4 gc descriptors + scfsi + validation.  si points at the side info. */
static void parse_side(const uint8_t *si, p3_frame *fr, p3_gc *gc)
{
  uint8_t buf[36];
  unsigned nch = fr->nch, silen = nch == 1 ? 17 : 32;
  memcpy(buf, si, silen); memset(buf + silen, 0, sizeof buf - silen);
  bitrd b = {buf, 0};
  (void)getbits(&b, 9);                              /* main_data_begin, already taken by hop() */
  (void)getbits(&b, nch == 1 ? 5 : 3);               /* private bits */
  unsigned scfsi = 0;
  for (unsigned ch = 0; ch < nch; ch++)
    for (unsigned band = 0; band < 4; band++) scfsi |= getbits(&b, 1) << (4 * ch + band);
  fr->scfsi = (uint8_t)scfsi;
  unsigned start = 0, bad = 0;
  memset(gc, 0, 4 * sizeof *gc);
  for (unsigned gr = 0; gr < 2; gr++) for (unsigned ch = 0; ch < nch; ch++) {
    p3_gc *g = &gc[gr * 2 + ch];
    unsigned p23l = getbits(&b, 12), bigv = getbits(&b, 9), gain = getbits(&b, 8), sfc = getbits(&b, 4);
    unsigned ws = getbits(&b, 1), bt = 0, mixed = 0, ts[3] = {0, 0, 0}, sbg[3] = {0, 0, 0}, r0, r1;
    if (ws) {
      bt = getbits(&b, 2); mixed = getbits(&b, 1);
      ts[0] = getbits(&b, 5); ts[1] = getbits(&b, 5);
      sbg[0] = getbits(&b, 3); sbg[1] = getbits(&b, 3); sbg[2] = getbits(&b, 3);
      r0 = (bt == 2 && !mixed) ? 8 : 7;              /* implicit (pdmp3.c:1181-1185) */
      r1 = 20 - r0;
      if (bt == 0) bad = 1;                          /* forbidden by ISO; the reference would read a stale table_select[2] */
    } else {
      ts[0] = getbits(&b, 5); ts[1] = getbits(&b, 5); ts[2] = getbits(&b, 5);
      r0 = getbits(&b, 4); r1 = getbits(&b, 3);
      if (r0 + r1 + 2 > 22) bad = 1;                 /* sfb index out of the table */
    }
    unsigned pre = getbits(&b, 1), scale = getbits(&b, 1), c1t = getbits(&b, 1);
    if (bigv > 288) bad = 1;
    g->w0 = p23l | bigv << 12 | gain << 21 | pre << 29 | scale << 30 | c1t << 31;
    g->w1 = sfc | ws << 4 | bt << 5 | mixed << 7 | ts[0] << 8 | ts[1] << 13 | ts[2] << 18 | r0 << 23 | r1 << 27;
    g->w2 = sbg[0] | sbg[1] << 3 | sbg[2] << 6 | (start & 0x3fffu) << 9;
    /* where the next part starts: behind this one -- or, for an empty part in reference-compatible mode, behind the
     * scalefactor bits the reference reads regardless (ISO mode: an empty part has no bits at all) */
    start += p23l ? p23l : ((fr->flags & P3_FRAME_ISO) ? 0u : part2_bits(sfc, ws, bt, mixed, gr, (scfsi >> (4 * ch)) & 15u));
  }
  if (start > 8u * ((unsigned)fr->main_begin + fr->main_size)) bad = 1;   /* parts overrun the frame's data */
  if (bad) fr->flags |= P3_FRAME_BAD;
}

typedef struct { const uint8_t *data; p3_frame *fr; p3_gc *gc; int64_t lo, hi; int any_empty; } job_t;
static void *worker(void *arg)
{
  job_t *j = (job_t *)arg;
  for (int64_t f = j->lo; f < j->hi; f++) {
    p3_frame *fr = &j->fr[f];
    parse_side(j->data + fr->main_off - (fr->nch == 1 ? 17 : 32), fr, &j->gc[4 * f]);
    for (unsigned k = 0; k < 4; k++) if ((k & 1) < fr->nch && P3_GC_P23L(j->gc[4 * f + k]) == 0) j->any_empty = 1;
  }
  return NULL;
}

int p3_parse(const uint8_t *data, uint64_t n, const p3_parse_opts *o, p3_parse_state *st, p3_parsed *out)
{
  return p3_parse_into(data, n, o, st, out, NULL, NULL, 0);
}

/* Same, writing into caller-owned descriptor arrays (e.g. page-locked) of `cap` frames; the result is
 * then not freed by p3_parsed_free(). */
int p3_parse_into(const uint8_t *data, uint64_t n, const p3_parse_opts *o, p3_parse_state *st, p3_parsed *out,
                  p3_frame *frames_buf, p3_gc *gcs_buf, int64_t buf_cap)
{
  p3_parse_opts od = {0, 0, 0, 0, 0, 0};
  p3_parse_state sd = {0, 0, 0, -1, -1};
  if (!data || !out) return P3_EINVAL;
  if (!o) o = &od;
  if (!st) st = &sd;
  memset(out, 0, sizeof *out);

  /* ---- phase 1: sequential header hop ---- */
  int64_t cap = o->max_frames > 0 ? o->max_frames : (int64_t)(n / 96 + 16), nf = 0;
  if (cap > (int64_t)(n / 96 + 16)) cap = (int64_t)(n / 96 + 16);
  const int ext = frames_buf && gcs_buf;
  if (ext && cap > buf_cap) cap = buf_cap;
  p3_frame *fr = ext ? frames_buf : (p3_frame *)malloc((size_t)(cap > 0 ? cap : 1) * sizeof *fr);
  if (!fr) return P3_ENOMEM;
  uint64_t pos = 0;
  int stop = 0;
  while (1) {
    if ((o->max_frames > 0 && nf >= o->max_frames) || nf >= cap) { stop = 1; break; }
    if (n - pos < (o->lookahead ? o->lookahead : 4)) break;
    uint64_t p = pos, lim = pos + 1153 < n - 3 ? pos + 1153 : n - 3;   /* resync window (pdmp3.c:1337) */
    while (p < lim && !header_ok(data + p)) p++;
    if (p >= lim) { if (lim == pos + 1153) stop = 2; break; }
    const uint8_t *h = data + p;
    unsigned prot = h[1] & 1, br = h[2] >> 4, sf = (h[2] >> 2) & 3, pad = (h[2] >> 1) & 1;
    unsigned mode = h[3] >> 6, mext = (h[3] >> 4) & 3;
    unsigned nch = mode == 3 ? 1 : 2, silen = nch == 1 ? 17 : 32;
    if (nf > 0 && (nch != fr[0].nch || sf != fr[0].sfreq)) { stop = 3; pos = p; break; }   /* format change: next batch */
    unsigned fsize = 144u * k_bitrate[br] * 1000u / k_sfreq[sf] + pad;            /* pdmp3.c:1135-1138 */
    unsigned hdr = 4 + (prot ? 0 : 2);
    if (fsize < hdr + silen || p + fsize > n) break;                              /* incomplete frame */
    /* the hop is one cache miss per frame; fetch the headers a few frames ahead assuming a similar
     * frame length (exact for CBR, harmless for VBR) */
    if (p + 8 * (uint64_t)fsize + 64 < n) { __builtin_prefetch(data + p + 6 * (uint64_t)fsize); __builtin_prefetch(data + p + 6 * (uint64_t)fsize + 64); __builtin_prefetch(data + p + 8 * (uint64_t)fsize); }
    p3_frame *f = &fr[nf];
    f->main_off = p + hdr + silen;
    f->main_size = (uint16_t)(fsize - hdr - silen);
    f->main_begin = (uint16_t)((h[hdr] << 1) | (h[hdr + 1] >> 7));
    f->main_pos = st->main_pos;
    f->nch = (uint8_t)nch; f->mode = (uint8_t)mode; f->mode_ext = (uint8_t)mext; f->sfreq = (uint8_t)sf;
    f->scfsi = 0; f->bitrate_kbps = k_bitrate[br];
    f->flags = o->iso ? P3_FRAME_ISO : 0;
    /* reservoir rule of Get_Main_Data (pdmp3.c:1101-1120) */
    if (f->main_begin > st->top) { f->flags |= P3_FRAME_NODATA; st->top += f->main_size; }
    else st->top = (uint32_t)f->main_begin + f->main_size;
    if ((uint64_t)nf < o->warmup_frames) { f->flags |= P3_FRAME_WARMUP; f->pcm_index = 0xffffffffu; }
    else { f->flags |= P3_FRAME_DECODE; f->pcm_index = st->pcm_index++; out->n_pcm_frames++; }
    st->main_pos += f->main_size;
    st->nch = (int32_t)nch; st->sfreq = (int32_t)sf;
    pos = p + fsize;
    nf++;
  }
  out->consumed = pos; out->n_frames = nf; out->frames = fr; out->stop = stop;
  out->external = ext;
  out->gcs = ext ? gcs_buf : (p3_gc *)malloc((size_t)(nf > 0 ? nf : 1) * 4 * sizeof(p3_gc));
  if (!out->gcs) { free(fr); return P3_ENOMEM; }

  if (o->hop_only) {                                      /* the side info is parsed on the device (k_sideinfo) */
    out->hop_only = 1;
    memset(out->gcs, 0, (size_t)(nf > 0 ? nf : 1) * 4 * sizeof(p3_gc));
    return P3_OK;
  }
  /* ---- phase 2: side info, in parallel ---- */
  int nt = o->nthreads;
  if (nt <= 0) { long c = sysconf(_SC_NPROCESSORS_ONLN); nt = c > 16 ? 16 : (int)c; }
  int any_empty = 0;
  if (nf < 4096 || nt < 2) { job_t j = {data, fr, out->gcs, 0, nf, 0}; worker(&j); any_empty = j.any_empty; }
  else {
    pthread_t th[64]; job_t jb[64];
    if (nt > 64) nt = 64;
    for (int t = 0; t < nt; t++) {
      jb[t] = (job_t){data, fr, out->gcs, nf * t / nt, nf * (t + 1) / nt, 0};
      if (pthread_create(&th[t], NULL, worker, &jb[t])) { worker(&jb[t]); th[t] = 0; }
    }
    for (int t = 0; t < nt; t++) { if (th[t]) pthread_join(th[t], NULL); any_empty |= jb[t].any_empty; }
  }
  /* Q6 (pdmp3.c:2057-2061): a zero-length part leaves count1 of its [gr][ch] slot stale.  Record how
   * many frames back the slot was last written so the device can fetch that count1 (w3 = 0: own).
   * ISO mode: an empty part simply has count1 = 0 (w3 stays 0). */
  if (any_empty && !o->iso) {
    int64_t last[4] = {-1, -1, -1, -1};
    for (int64_t f = 0; f < nf; f++) for (unsigned k = 0; k < 4; k++) {
      if ((k & 1) >= fr[f].nch) continue;
      p3_gc *g = &out->gcs[4 * f + k];
      if (P3_GC_P23L(*g) != 0 || (fr[f].flags & (P3_FRAME_NODATA | P3_FRAME_BAD))) { last[k] = f; g->w3 = 0; }
      else g->w3 = last[k] >= 0 ? (uint32_t)(f - last[k]) : 0x7fffffffu;
    }
  }
  return P3_OK;
}

void p3_parsed_free(p3_parsed *p) { if (p) { if (!p->external) { free(p->frames); free(p->gcs); } p->frames = NULL; p->gcs = NULL; } }
