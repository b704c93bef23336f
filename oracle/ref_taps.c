/* oracle/ref_taps.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Stage-tap harness around the UNMODIFIED reference decoder.  All stage functions of the
 * reference are `static` (pdmp3.c:198-233), so the only way to call them one by one is to
 * compile the reference translation unit into this one.  The reference source is NOT copied
 * into this repo: it is #included from where it lies (REF_SRC, default
 * /root/reference/pdmp3.c) and the result goes to oracle/_ref/ (git-ignored).
 *
 * The driver loop below mirrors pdmp3_read() (pdmp3.c:2431-2481) and Decode_L3()
 * (pdmp3.c:1024-1060) so that the per-stage state can be captured between the calls.
 */
#define NDEBUG 1
#ifndef REF_SRC
#define REF_SRC "/root/reference/pdmp3.c"
#endif
#include REF_SRC

#include <stdint.h>

typedef struct {
  /* every pointer may be NULL; arrays are [max_frames][...] */
  int32_t  *side;     /* [f][2][2][20] parsed side info, layout documented in tests/p3harness.py (gc_fields) */
  int32_t  *hdr;      /* [f][8]  mode, mode_ext, sfreq, bitrate_index, padding, protection, main_data_begin, nch */
  uint8_t  *scf_l;    /* [f][2][2][21] */
  uint8_t  *scf_s;    /* [f][2][2][12][3] */
  int16_t  *is_huff;  /* [f][2][2][576] after Read_Huffman */
  int32_t  *count1;   /* [f][2][2] */
  float    *xr_req;   /* [f][2][2][576] after L3_Requantize */
  float    *xr_reo;   /* after L3_Reorder */
  float    *xr_ste;   /* after L3_Stereo */
  float    *xr_ali;   /* after L3_Antialias */
  float    *y_hyb;    /* after L3_Hybrid_Synthesis + L3_Frequency_Inversion */
  int16_t  *pcm;      /* [f][1152][2] interleaved L,R (both channels always, pdmp3.c:2032-2041) */
} ref_taps_t;

static void cp576(float *dst, pdmp3_handle *id, size_t f, unsigned gr, unsigned ch) {
  if (dst) memcpy(dst + ((f * 2 + gr) * 2 + ch) * 576, id->g_main_data.is[gr][ch], 576 * sizeof(float));
}

/* Decode a whole in-memory stream frame by frame, capturing taps.  Returns frames decoded. */
long ref_taps_decode(const unsigned char *stream, size_t nbytes, long max_frames, ref_taps_t *t)
{
  pdmp3_handle *id = calloc(1, sizeof(pdmp3_handle));      /* G0: zeroed handle */
  size_t fed = 0;
  long f = 0;
  pdmp3_open_feed(id);
  for (;;) {
    /* keep the 16 KiB ring as full as it will go (never exactly full: SURVEY 8b hazard) */
    while (fed < nbytes) {
      size_t fr = Get_Inbuf_Free(id);
      size_t n = nbytes - fed;
      if (fr <= 1) break;
      if (n > fr - 1) n = fr - 1;
      if (n > 4096) n = 4096;
      if (pdmp3_feed(id, stream + fed, n) != PDMP3_OK) break;
      fed += n;
    }
    if (f >= max_frames) break;
    if (Get_Inbuf_Filled(id) < 2 * 576) break;                /* pdmp3.c:2445 */
    size_t pos = id->processed; unsigned mark = id->istart;
    int res = Read_Frame(id);                                 /* pdmp3.c:2449 */
    if (!(res == PDMP3_OK || res == PDMP3_NEW_FORMAT)) {
      id->processed = pos; id->istart = mark;
      break;
    }
    unsigned nch = (id->g_frame_header.mode == mpeg1_mode_single_channel ? 1 : 2);
    t_mpeg1_side_info *si = &id->g_side_info;
    if (t->hdr) {
      int32_t *h = t->hdr + f * 8;
      h[0] = id->g_frame_header.mode; h[1] = id->g_frame_header.mode_extension;
      h[2] = id->g_frame_header.sampling_frequency; h[3] = id->g_frame_header.bitrate_index;
      h[4] = id->g_frame_header.padding_bit; h[5] = id->g_frame_header.protection_bit;
      h[6] = si->main_data_begin; h[7] = nch;
    }
    for (unsigned gr = 0; gr < 2; gr++) for (unsigned ch = 0; ch < nch; ch++) {
      size_t g = (f * 2 + gr) * 2 + ch;
      if (t->side) {
        int32_t *s = t->side + g * 20;
        s[0] = si->part2_3_length[gr][ch]; s[1] = si->big_values[gr][ch]; s[2] = si->global_gain[gr][ch];
        s[3] = si->scalefac_compress[gr][ch]; s[4] = si->win_switch_flag[gr][ch]; s[5] = si->block_type[gr][ch];
        s[6] = si->mixed_block_flag[gr][ch];
        s[7] = si->table_select[gr][ch][0]; s[8] = si->table_select[gr][ch][1]; s[9] = si->table_select[gr][ch][2];
        s[10] = si->subblock_gain[gr][ch][0]; s[11] = si->subblock_gain[gr][ch][1]; s[12] = si->subblock_gain[gr][ch][2];
        s[13] = si->region0_count[gr][ch]; s[14] = si->region1_count[gr][ch];
        s[15] = si->preflag[gr][ch]; s[16] = si->scalefac_scale[gr][ch]; s[17] = si->count1table_select[gr][ch];
        s[18] = si->scfsi[ch][0] | si->scfsi[ch][1] << 1 | si->scfsi[ch][2] << 2 | si->scfsi[ch][3] << 3;
        s[19] = 0;
      }
      if (t->count1) t->count1[g] = si->count1[gr][ch];
      if (t->scf_l) for (int i = 0; i < 21; i++) t->scf_l[g * 21 + i] = id->g_main_data.scalefac_l[gr][ch][i];
      if (t->scf_s) for (int i = 0; i < 36; i++) t->scf_s[g * 36 + i] = id->g_main_data.scalefac_s[gr][ch][i / 3][i % 3];
      if (t->is_huff) for (int i = 0; i < 576; i++) t->is_huff[g * 576 + i] = (int16_t)id->g_main_data.is[gr][ch][i];
    }
    /* Decode_L3 (pdmp3.c:1024-1060) with taps between the stages */
    for (unsigned gr = 0; gr < 2; gr++) {
      for (unsigned ch = 0; ch < nch; ch++) {
        L3_Requantize(id, gr, ch); cp576(t->xr_req, id, f, gr, ch);
        L3_Reorder(id, gr, ch);    cp576(t->xr_reo, id, f, gr, ch);
      }
      L3_Stereo(id, gr);
      for (unsigned ch = 0; ch < nch; ch++) cp576(t->xr_ste, id, f, gr, ch);
      for (unsigned ch = 0; ch < nch; ch++) {
        L3_Antialias(id, gr, ch);  cp576(t->xr_ali, id, f, gr, ch);
        L3_Hybrid_Synthesis(id, gr, ch);
        L3_Frequency_Inversion(id, gr, ch); cp576(t->y_hyb, id, f, gr, ch);
        L3_Subband_Synthesis(id, gr, ch, id->out[gr]);
      }
    }
    if (t->pcm) {
      int16_t *p = t->pcm + f * 2304;
      for (unsigned gr = 0; gr < 2; gr++) for (int i = 0; i < 576; i++) {
        p[(gr * 576 + i) * 2 + 0] = (int16_t)((id->out[gr][i] >> 16) & 0xffff);
        p[(gr * 576 + i) * 2 + 1] = (int16_t)(id->out[gr][i] & 0xffff);
      }
    }
    f++;
  }
  free(id);
  return f;
}

/* Decode ONE Huffman word from a bit string with the reference tree walk (pdmp3.c:1593-1643).
 * bits: MSB-first bytes (at least 16 valid).  out = {x, y, v, w, bits_consumed}. */
void ref_huff_decode(unsigned table_num, const unsigned char *bits, int nbytes, int32_t out[5])
{
  static pdmp3_handle *id;
  if (!id) id = calloc(1, sizeof(pdmp3_handle));
  for (int i = 0; i < nbytes && i < 64; i++) id->g_main_data_vec[i] = bits[i];
  Set_Main_Pos(id, 0);
  int32_t x = 0, y = 0, v = 0, w = 0;
  (void)Huffman_Decode(id, table_num, &x, &y, &v, &w);
  out[0] = x; out[1] = y; out[2] = v; out[3] = w; out[4] = (int32_t)Get_Main_Pos(id);
}

/* Expose the reference's constant tables for bit-for-bit comparison with p3_tables.c */
const float *ref_tab_cs(void)        { return cs; }
const float *ref_tab_ca(void)        { return ca; }
const float *ref_tab_is_ratios(void) { return is_ratios; }
const float *ref_tab_imdct_win(void) { return &g_imdct_win[0][0]; }
const float *ref_tab_cos_n12(void)   { return &cos_N12[0][0]; }
const float *ref_tab_cos_n36(void)   { return &cos_N36[0][0]; }
const float *ref_tab_synth_dtbl(void){ return g_synth_dtbl; }
float        ref_pow43(unsigned i)   { return Requantize_Pow_43(i); }
const unsigned *ref_tab_sfb_l(int sfreq) { return g_sf_band_indices[sfreq].l; }
const unsigned *ref_tab_sfb_s(int sfreq) { return g_sf_band_indices[sfreq].s; }
unsigned ref_sizeof_handle(void) { return sizeof(pdmp3_handle); }
