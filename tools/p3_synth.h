/* p3_synth.h -- synthetic MPEG-1 Layer III stream generator (TEST / BENCH INFRASTRUCTURE).
 * There is no MP3 encoder or sample file in the build image, so streams are produced by a
 * forward Huffman encoder over random spectra, staying inside the envelope in which the
 * reference decoder is well defined. */
#ifndef P3_SYNTH_H
#define P3_SYNTH_H
#include <stdint.h>
typedef struct {
  uint64_t seed;
  int32_t bitrate_index;   /* 1..14 CBR; 0 = VBR, uniform 1..14 per frame                        */
  int32_t mode;            /* 0 stereo, 1 joint stereo, 2 dual channel, 3 mono                    */
  int32_t mode_ext;        /* 0..3 fixed; -1 = uniform per frame (joint stereo only)              */
  int32_t sfreq;           /* 0: 44.1k 1: 48k 2: 32k                                              */
  int32_t blocks;          /* 0 long only; 1 block-type state machine incl. short and mixed       */
  int32_t reservoir;       /* 0: main_data_begin always 0; 1: bit reservoir in use                */
  int32_t scalefacs;       /* 0: scalefac_compress 0; 1: random scalefactors (sfb 0 kept 0, G1)   */
  int32_t gain;            /* centre of global_gain (e.g. 150)                                    */
  int32_t fill_pm;         /* mean fraction of the frame's main-data bits to spend, per mille      */
  int32_t crc;             /* 1: protection_bit=0 + 2-byte CRC word (not verified by the decoders) */
  int32_t count1_b_pm;     /* per-mille of granules using count1table_select=1 (reference quirk Q1)*/
  int32_t overrun_pm;      /* per-mille of granules whose part2_3_length ends inside the last quad  */
  int32_t max_table;       /* highest big_values table number to use (e.g. 31; 15 = no linbits)    */
  int32_t garbage_pm;      /* per-mille of frames preceded by 1..40 junk bytes (resync test)       */
  int32_t iso;             /* 1: stream for the ISO mode of the decoder (P3_FRAME_ISO), outside the reference's envelope:
                              intensity stereo also with short blocks in channel 0 (G6 lifted), intensity positions <= 7 in
                              the RIGHT channel's scalefactors, 3 % of the parts empty (part2_3_length = 0, G5 lifted)      */
  int32_t peak_pm;         /* loudest spectral line allowed, per mille of full scale (0 = the default 300: clipping stays rare);
                              ~2000-4000 with a high `gain` gives near-full-scale PCM with a clip fraction around 1e-3 */
} p3_synth_cfg;

/* Writes n_frames frames.  Returns bytes written or <0 if `cap` is too small.
 * is_out (optional): the encoded quantised spectra, [n_frames][2][2][576] int16, zero beyond count1. */
int64_t p3_synth(const p3_synth_cfg *cfg, int64_t n_frames, uint8_t *out, uint64_t cap, int16_t *is_out);
#endif
