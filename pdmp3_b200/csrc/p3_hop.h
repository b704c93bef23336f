/* p3_hop.h -- data structures of the device-side frame hop (p3_hop.cu), shared with p3_cabi.cu */
#pragma once
#include <stdint.h>
#include "../../include/pdmp3_b200.h"

#ifndef HOP_SEG
#define HOP_SEG   16384u               /* bytes of the raw stream per segment (one thread each); must exceed the longest frame + the resync window (2594) */
#endif
#define HOP_LCAP  (HOP_SEG / 96 + 6)   /* headers a segment's chain can visit: frames are >= 96 bytes (32 kbps at 48 kHz) */
#define HOP_PCAP  8                    /* headers kept in front of the meeting point before the list is rewritten instead */

typedef struct {
  uint64_t entry;                      /* search position the segment was last resolved from (speculation: its first byte) */
  uint64_t l_exit;                     /* where the chain of list L leaves the segment; ~0: it ends inside */
  uint64_t l_term;                     /* ... and the position at which it ended */
  uint64_t term_pos;                   /* position at which the TRUE chain ended inside this segment (stop != 0) */
  uint16_t nL, j0, nP;                 /* true frames of the segment = P[0..nP) then L[j0..nL)  (offsets from the segment's first byte) */
  uint8_t  stop, l_stop;               /* 0: the chain goes on; 1: out of data (p3_parsed.stop 0); 3: no header within 1153 bytes (stop 2) */
  uint16_t P[HOP_PCAP];
  /* aggregates over the true frames (k_hop_agg) */
  uint32_t sum_ms; int32_t fa; uint32_t fb;      /* main-data bytes; reservoir function top' = min(top, fa) + fb */
  uint16_t cnt, chg, max_ms;                     /* frames; index of the first frame whose format differs from the segment's first (0xffff: none); largest main_size */
  uint8_t  fmt, pad;                             /* nch | sfreq << 2 of the first frame */
  /* exclusive prefix: first within the segment's block of 128 (k_hop_agg), then over the whole stream (k_hop_apply) */
  uint64_t base_idx, base_pos; uint32_t top_in, pad2;
  int32_t  la; uint32_t lb;                      /* reservoir function of the block's segments in front of this one */
} hop_seg;

typedef struct { int64_t cnt, ms, a, b; } hop_part;   /* aggregate of a block of 128 segments; after k_hop_scan: the exclusive prefix over the blocks */

typedef struct {
  int64_t  n_total;                    /* frames on the chain before truncation */
  int64_t  n_frames, n_pcm_frames;
  uint64_t consumed;
  int32_t  stop, changed, nch, sfreq;  /* changed: the last k_hop_resolve round moved an exit (run another round) */
  uint32_t max_main, maxg;             /* largest main_size; largest main-data span of a group of 32 frames (K1's window) */
  uint64_t total_ms;                   /* main-data bytes of the kept frames */
  p3_parse_state st;                   /* parser state after the kept frames */
  /* scratch of the scan kernels */
  unsigned long long mismatch;         /* index of the first frame whose format differs from frame 0's (~0: none) */
  unsigned long long term_pos; int32_t term_stop, pad_;
  int64_t tot_a, tot_b;
} p3_hop_result;

#ifdef __CUDACC__
struct p3_hop_work {                   /* device scratch, grown on demand */
  hop_seg *seg; uint16_t *lists; uint64_t *exit[2]; int64_t cap_seg; hop_part *part;
  p3_hop_result *d_res; p3_hop_result *h_res, *h_res_dev;     /* h_res: page-locked and mapped (h_res_dev: its device address) */
};
int p3_hop_work_ensure(p3_hop_work *w, int64_t nseg);
void p3_hop_work_free(p3_hop_work *w);
/* phase 1: chain + counts -> *h_res (synchronises `st`); phase 2: the records -> d_frames[h_res->n_frames], maxg, the
 * tail for the next batch -> second synchronisation */
int p3_hop_count(p3_hop_work *w, cudaStream_t st, const uint8_t *d_raw, uint64_t n, const p3_parse_opts *o, const p3_parse_state *ps, int64_t frame_cap);
int p3_hop_emit(p3_hop_work *w, cudaStream_t st, const uint8_t *d_raw, uint64_t n, const p3_parse_opts *o, const p3_parse_state *ps,
                p3_frame *d_frames, const uint8_t *d_tail_in, uint8_t *d_tail_out);
#endif
