/* p3_hop.cu -- the frame hop ON THE DEVICE: Search_Header / Read_Header (pdmp3.c:1322-1340, 1252-1320), the frame length
 * of Read_Main_L3 (1353-1368) and the reservoir bookkeeping of Get_Main_Data (1096-1122) for a whole buffer at once,
 * producing the same p3_frame[] array as the sequential host hop of p3_parse.c (tests hold the two to the same bytes).
 *
 * The hop is a chain: from a search position `pos` the decoder takes the first valid header p >= pos (within 1153 bytes)
 * and continues at p + frame_length(p).  next(pos) does not depend on anything before pos, so chains started at
 * different positions merge as soon as they meet the same header.  That makes the chain parallel:
 *
 *   k_hop_spec     one thread per 16 KB segment follows the chain from the segment's first byte (a speculative entry:
 *                  the true chain usually enters a few hundred bytes later) and records the headers it visits (list L)
 *                  and where it leaves the segment;
 *   k_hop_resolve  one thread per segment takes the exit of the previous segment as its TRUE entry and follows the chain
 *                  from there until it meets a header of L (typically the first or second one): true frames = the few
 *                  headers before the meeting point (list P) + the rest of L.  If every chain met its list before the
 *                  end of the segment, all exits are unchanged and the entries were the true ones (induction from
 *                  segment 0, whose entry is exact).  Otherwise the kernel is run again with the new exits; every round
 *                  fixes at least one more segment, and in practice one round is enough (a false sync inside main data
 *                  is followed by a real header one hop later);
 *   k_hop_agg      per segment: number of frames, main-data bytes, the reservoir function, format of the frames;
 *   k_hop_scan     exclusive scan of those over the segments (one CTA), truncation rules (max_frames, format change),
 *                  the result block the host reads;
 *   k_hop_write    per segment: the p3_frame records, in place.
 *
 * Reservoir rule as a scan: Get_Main_Data keeps `top` = bytes available; a frame with main_data_begin > top is not
 * decodable (NODATA) and top += main_size, otherwise top = main_data_begin + main_size.  Both cases are
 * top' = min(top, main_data_begin) + main_size  (NODATA iff main_data_begin > top), and functions min(t, a) + b compose
 * as (a1, b1) then (a2, b2) = (min(a1, a2 - b1), b1 + b2): an associative scan.
 */
#include "p3_device.cuh"
#include "p3_hop.h"

#define HOP_TERM   (~(uint64_t)0)          /* exit / entry value: the chain has ended */

__device__ __forceinline__ bool hop_header_ok(const uint8_t *p)     /* checks of pdmp3.c:1264, 1293-1315, 1329 (p3_parse.c header_ok) */
{
  const uint32_t b1 = p[1], b2 = p[2];
  if (p[0] != 0xff || (b1 & 0xfe) != 0xfa) return false;             /* sync, id = 1 (MPEG-1), layer field 01 (Layer III) */
  const uint32_t br = b2 >> 4, sf = (b2 >> 2) & 3;
  return br != 0 && br != 15 && sf != 3;
}

__constant__ uint16_t c_bitrate[16] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0};   /* pdmp3.c:524-527 */
__constant__ uint32_t c_sfreq[4] = {44100, 48000, 32000, 1};                                                   /* pdmp3.c:529 */

struct hop_hdr { uint32_t fsize, hdr, silen, br_kbps; uint8_t nch, mode, mext, sf; };
__device__ __forceinline__ hop_hdr hop_parse_header(const uint8_t *h)
{
  hop_hdr r;
  const uint32_t prot = h[1] & 1, br = h[2] >> 4, sf = (h[2] >> 2) & 3, pad = (h[2] >> 1) & 1;
  r.mode = (uint8_t)(h[3] >> 6); r.mext = (uint8_t)((h[3] >> 4) & 3);
  r.nch = r.mode == 3 ? 1 : 2; r.sf = (uint8_t)sf;
  r.silen = r.nch == 1 ? 17u : 32u;
  r.br_kbps = c_bitrate[br];
  r.fsize = 144u * r.br_kbps * 1000u / c_sfreq[sf] + pad;             /* pdmp3.c:1135-1138 */
  r.hdr = 4 + (prot ? 0 : 2);
  return r;
}

/* One hop from search position `pos` (the loop body of p3_parse_into, p3_parse.c): 0 = a complete frame at p (length
 * fsize); otherwise the chain ends here with stop code (ret - 1): 1 -> stop 0 (out of data), 3 -> stop 2 (no header within
 * 1153 bytes, pdmp3.c:1337). */
__device__ __forceinline__ int hop_step(const uint8_t *__restrict__ raw, uint64_t n, uint64_t pos, uint32_t lookahead, uint64_t &p_out, uint32_t &fsize_out)
{
  if (n - pos < (lookahead ? lookahead : 4u)) return 1;
  const uint64_t lim = pos + 1153 < n - 3 ? pos + 1153 : n - 3;      /* resync window */
  uint64_t p = pos;
  if (p < lim && !hop_header_ok(raw + p)) {
    /* junk in front of the header (or a speculative entry in the middle of main data): look for the next 0xff byte a word
     * at a time, test the candidates */
    p++;
    while (p < lim && (reinterpret_cast<uintptr_t>(raw + p) & 3)) { if (hop_header_ok(raw + p)) goto found; p++; }
    while (p + 4 <= lim) {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(raw + p));
      if (((~w - 0x01010101u) & w & 0x80808080u) != 0) {            /* some byte of w is 0xff (exact test: haszero(~w)) */
        for (int k = 0; k < 4; k++) if (((w >> (8 * k)) & 0xffu) == 0xffu && hop_header_ok(raw + p + k)) { p += k; goto found; }
      }
      p += 4;
    }
    while (p < lim) { if (hop_header_ok(raw + p)) goto found; p++; }
  }
found:
  if (p >= lim) return lim == pos + 1153 ? 3 : 1;
  const hop_hdr h = hop_parse_header(raw + p);
  if (h.fsize < h.hdr + h.silen || p + h.fsize > n) return 1;         /* incomplete frame */
  p_out = p; fsize_out = h.fsize;
  return 0;
}

/* ---- speculative chains ---- */
extern "C" __global__ void __launch_bounds__(128)
k_hop_spec(const uint8_t *__restrict__ raw, uint64_t n, uint32_t lookahead, int64_t nseg, hop_seg *__restrict__ seg, uint16_t *__restrict__ lists, uint64_t *__restrict__ exit0)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nseg) return;
  const uint64_t b = (uint64_t)k * HOP_SEG, bn = b + HOP_SEG;
  uint16_t *L = lists + (size_t)k * HOP_LCAP;
  uint64_t pos = b, ex; uint32_t nL = 0, stop = 0;
  for (;;) {
    if (pos >= bn) { ex = pos; break; }
    uint64_t p; uint32_t fs;
    const int r = hop_step(raw, n, pos, lookahead, p, fs);
    if (r) { stop = (uint32_t)r; ex = HOP_TERM; break; }
    if (nL < HOP_LCAP) L[nL] = (uint16_t)(p - b);
    nL++;
    pos = p + fs;
  }
  hop_seg s;
  s.entry = b; s.l_exit = ex; s.l_term = pos; s.term_pos = pos;
  s.nL = (uint16_t)nL; s.j0 = 0; s.nP = 0; s.stop = (uint8_t)stop; s.l_stop = (uint8_t)stop;
  for (int i = 0; i < HOP_PCAP; i++) s.P[i] = 0;
  s.sum_ms = 0; s.fa = 0; s.fb = 0; s.cnt = 0; s.chg = 0xffff; s.max_ms = 0; s.fmt = 0; s.pad = 0;
  s.base_idx = 0; s.base_pos = 0; s.top_in = 0; s.pad2 = 0;
  seg[k] = s;
  exit0[k] = ex;
}

/* ---- one round of resolution: entry of segment k := exit of segment k-1 after the previous round ---- */
extern "C" __global__ void __launch_bounds__(128)
k_hop_resolve(const uint8_t *__restrict__ raw, uint64_t n, uint32_t lookahead, int64_t nseg, hop_seg *__restrict__ seg, uint16_t *__restrict__ lists,
              const uint64_t *__restrict__ exit_prev, uint64_t *__restrict__ exit_cur, int *__restrict__ changed)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nseg) return;
  if (k == 0) { exit_cur[0] = exit_prev[0]; return; }                /* segment 0 enters at byte 0: its speculation is the truth */
  const uint64_t c = exit_prev[k - 1];
  hop_seg s = seg[k];
  if (c == s.entry) { exit_cur[k] = exit_prev[k]; return; }          /* same entry as last time: nothing to redo */
  const uint64_t b = (uint64_t)k * HOP_SEG, bn = b + HOP_SEG;
  uint16_t *L = lists + (size_t)k * HOP_LCAP;
  uint64_t ex;
  s.entry = c; s.stop = 0; s.term_pos = 0;
  if (c == HOP_TERM || c >= bn) { s.nP = 0; s.j0 = s.nL; ex = c; }   /* the chain ended before this segment (or, impossible with 16 KB segments, skips it) */
  else {
    uint64_t pos = c; uint32_t nP = 0, j = 0; int how = 0;           /* how: 1 met the list, 2 left the segment / ended without meeting it, 3 rewrite */
    for (;;) {
      if (pos >= bn) { ex = pos; how = 2; break; }
      uint64_t p; uint32_t fs;
      const int r = hop_step(raw, n, pos, lookahead, p, fs);
      if (r) { s.stop = (uint8_t)r; s.term_pos = pos; ex = HOP_TERM; how = 2; break; }
      const uint32_t rel = (uint32_t)(p - b);
      while (j < s.nL && L[j] < rel) j++;
      if (j < s.nL && L[j] == rel) { how = 1; break; }               /* the same header: from here on the two chains are one */
      if (nP == HOP_PCAP) { how = 3; break; }
      s.P[nP++] = (uint16_t)rel;
      pos = p + fs;
    }
    if (how == 1) { s.nP = (uint16_t)nP; s.j0 = (uint16_t)j; s.stop = s.l_stop; s.term_pos = s.l_term; ex = s.l_exit; }
    else if (how == 2) { s.nP = (uint16_t)nP; s.j0 = s.nL; }
    else {
      /* no meeting point within HOP_PCAP headers: the true chain itself becomes the segment's list */
      pos = c; uint32_t nL = 0;
      for (;;) {
        if (pos >= bn) { ex = pos; break; }
        uint64_t p; uint32_t fs;
        const int r = hop_step(raw, n, pos, lookahead, p, fs);
        if (r) { s.stop = (uint8_t)r; s.term_pos = pos; ex = HOP_TERM; break; }
        if (nL < HOP_LCAP) L[nL] = (uint16_t)(p - b);
        nL++;
        pos = p + fs;
      }
      s.nL = (uint16_t)nL; s.j0 = 0; s.nP = 0; s.l_stop = s.stop; s.l_term = s.term_pos; s.l_exit = ex;
    }
  }
  seg[k] = s;
  exit_cur[k] = ex;
  if (ex != exit_prev[k]) *changed = 1;
}

/* the i-th true frame of a segment: its header position */
__device__ __forceinline__ uint64_t hop_frame_pos(const hop_seg &s, const uint16_t *L, uint64_t b, uint32_t i)
{
  return b + (i < s.nP ? s.P[i] : L[s.j0 + i - s.nP]);
}

#define HOP_A_INF 0x3fffffff

/* ---- per-segment aggregates over the true frames, their exclusive prefix inside the block of 128 segments, the block's total ---- */
extern "C" __global__ void __launch_bounds__(128)
k_hop_agg(const uint8_t *__restrict__ raw, int64_t nseg, hop_seg *__restrict__ seg, const uint16_t *__restrict__ lists, hop_part *__restrict__ part)
{
  __shared__ int32_t s_cnt[128], s_ms[128], s_a[128], s_b[128];
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t cnt = 0, sum = 0, mx = 0, fmt0 = 0, chg = 0xffff; int64_t fa = HOP_A_INF, fb = 0;
  if (k < nseg) {
    const hop_seg s = seg[k];
    const uint64_t b = (uint64_t)k * HOP_SEG;
    const uint16_t *L = lists + (size_t)k * HOP_LCAP;
    cnt = (uint32_t)s.nP + (s.nL - s.j0);
    for (uint32_t i = 0; i < cnt; i++) {
      const uint8_t *h = raw + hop_frame_pos(s, L, b, i);
      const hop_hdr hh = hop_parse_header(h);
      const uint32_t ms = hh.fsize - hh.hdr - hh.silen, mb = ((uint32_t)h[hh.hdr] << 1) | (h[hh.hdr + 1] >> 7);
      const uint32_t fmt = hh.nch | (uint32_t)hh.sf << 2;
      if (i == 0) fmt0 = fmt; else if (fmt != fmt0 && chg == 0xffff) chg = i;
      sum += ms; mx = max(mx, ms);
      fa = min(fa, (int64_t)mb - fb); fb += ms;                       /* then top' = min(top, mb) + ms */
    }
  }
  s_cnt[threadIdx.x] = (int32_t)cnt; s_ms[threadIdx.x] = (int32_t)sum; s_a[threadIdx.x] = (int32_t)fa; s_b[threadIdx.x] = (int32_t)fb;
  __syncthreads();
  if (threadIdx.x == 0) {                                            /* exclusive scan over the block's 128 segments (serial: 128 steps) */
    int64_t rc = 0, rm = 0, ra = HOP_A_INF, rb = 0;
    for (int i = 0; i < 128; i++) {
      const int64_t c = s_cnt[i], m = s_ms[i], aa = s_a[i], bb = s_b[i];
      s_cnt[i] = (int32_t)rc; s_ms[i] = (int32_t)rm; s_a[i] = (int32_t)ra; s_b[i] = (int32_t)rb;
      rc += c; rm += m; ra = min(ra, aa - rb); rb += bb;
    }
    hop_part p; p.cnt = rc; p.ms = rm; p.a = ra; p.b = rb;
    part[blockIdx.x] = p;
  }
  __syncthreads();
  if (k < nseg) {
    hop_seg &o = seg[k];
    o.cnt = (uint16_t)cnt; o.sum_ms = sum; o.fa = (int32_t)fa; o.fb = (uint32_t)fb;
    o.chg = (uint16_t)chg; o.max_ms = (uint16_t)mx; o.fmt = (uint8_t)fmt0;
    o.base_idx = (uint64_t)s_cnt[threadIdx.x]; o.base_pos = (uint64_t)s_ms[threadIdx.x]; o.la = s_a[threadIdx.x]; o.lb = (uint32_t)s_b[threadIdx.x];
  }
}

/* ---- exclusive scan over the blocks' totals; ONE CTA of 1024 threads (128 x fewer items than segments) ---- */
extern "C" __global__ void __launch_bounds__(1024)
k_hop_scan(int64_t nblk, hop_part *__restrict__ part, p3_hop_result *__restrict__ res)
{
  __shared__ int64_t s_cnt[1024], s_ms[1024], s_a[1024], s_b[1024];
  const int t = threadIdx.x;
  const int64_t per = (nblk + 1023) / 1024, lo = (int64_t)t * per, hi = min(lo + per, nblk);
  int64_t cnt = 0, ms = 0, a = HOP_A_INF, b = 0;
  for (int64_t k = lo; k < hi; k++) { const hop_part p = part[k]; cnt += p.cnt; ms += p.ms; a = min(a, p.a - b); b += p.b; }
  s_cnt[t] = cnt; s_ms[t] = ms; s_a[t] = a; s_b[t] = b;
  __syncthreads();
  if (t == 0) {
    int64_t rc = 0, rm = 0, ra = HOP_A_INF, rb = 0;
    for (int i = 0; i < 1024; i++) {
      const int64_t c = s_cnt[i], m = s_ms[i], aa = s_a[i], bb = s_b[i];
      s_cnt[i] = rc; s_ms[i] = rm; s_a[i] = ra; s_b[i] = rb;
      rc += c; rm += m; ra = min(ra, aa - rb); rb += bb;
    }
    res->n_total = rc; res->total_ms = (uint64_t)rm; res->tot_a = ra; res->tot_b = rb;
    res->mismatch = ~0ull; res->term_stop = 0; res->term_pos = 0; res->max_main = 0;
  }
  __syncthreads();
  cnt = s_cnt[t]; ms = s_ms[t]; a = s_a[t]; b = s_b[t];
  for (int64_t k = lo; k < hi; k++) {
    const hop_part p = part[k];
    hop_part e; e.cnt = cnt; e.ms = ms; e.a = a; e.b = b;
    part[k] = e;
    cnt += p.cnt; ms += p.ms; a = min(a, p.a - b); b += p.b;
  }
}

/* ---- per segment: the global prefix (block prefix, then the prefix inside the block), first format change, end of the chain ---- */
extern "C" __global__ void __launch_bounds__(128)
k_hop_apply(int64_t nseg, hop_seg *__restrict__ seg, const hop_part *__restrict__ part, const uint64_t *__restrict__ exits, p3_parse_state st0, p3_hop_result *res)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nseg) return;
  hop_seg &s = seg[k];
  const hop_part p = part[blockIdx.x];
  const int64_t idx = p.cnt + (int64_t)s.base_idx, ms = p.ms + (int64_t)s.base_pos;
  const int64_t a = min(p.a, (int64_t)s.la - p.b), b = p.b + (int64_t)s.lb;     /* block prefix, then the block's segments in front of this one */
  s.base_idx = (uint64_t)idx; s.base_pos = st0.main_pos + (uint64_t)ms;
  s.top_in = (uint32_t)(min((int64_t)st0.top, a) + b);
  const uint32_t fmt0 = seg[0].cnt ? seg[0].fmt : 0u;
  if (s.cnt) {
    if (s.fmt != fmt0) atomicMin(&res->mismatch, (unsigned long long)idx);
    else if (s.chg != 0xffff) atomicMin(&res->mismatch, (unsigned long long)(idx + s.chg));
    atomicMax(&res->max_main, (uint32_t)s.max_ms);
  }
  if (s.entry != HOP_TERM && exits[k] == HOP_TERM) { res->term_stop = s.stop; res->term_pos = s.term_pos; }    /* the one segment in which the chain ends */
}

/* ---- the truncation rules (p3_parse.c loop head: frame limit first, then the format change) ---- */
extern "C" __global__ void k_hop_final(const hop_seg *__restrict__ seg, p3_parse_state st0, int64_t max_frames, uint32_t warmup, p3_hop_result *__restrict__ res)
{
  const uint32_t fmt0 = seg[0].cnt ? seg[0].fmt : 0u;
  const int64_t n_total = res->n_total;
  const int64_t lim = max_frames > 0 ? max_frames : INT64_MAX;
  const int64_t mis = res->mismatch == ~0ull ? INT64_MAX : (int64_t)res->mismatch;
  res->nch = fmt0 ? (int)(fmt0 & 3) : 2; res->sfreq = (int)(fmt0 >> 2);
  res->maxg = 0;
  if (mis < n_total && mis < lim) { res->n_frames = mis; res->stop = 3; res->consumed = 0; }          /* consumed: k_hop_write (header of frame n_frames) */
  else if (n_total >= lim) { res->n_frames = lim; res->stop = 1; res->consumed = 0; }                  /* consumed: k_hop_write (end of frame n_frames-1) */
  else { res->n_frames = n_total; res->stop = res->term_stop == 3 ? 2 : 0; res->consumed = res->term_pos; }
  res->n_pcm_frames = res->n_frames > (int64_t)warmup ? res->n_frames - warmup : 0;
  res->st = st0;                                                     /* (n_frames > 0: k_hop_write overwrites, total_ms too) */
}

/* ---- the records ---- */
extern "C" __global__ void __launch_bounds__(128)
k_hop_write(const uint8_t *__restrict__ raw, int64_t nseg, const hop_seg *__restrict__ seg, const uint16_t *__restrict__ lists,
            p3_parse_state st0, uint32_t warmup, uint32_t iso, p3_hop_result *__restrict__ res, p3_frame *__restrict__ frames)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nseg) return;
  const hop_seg s = seg[k];
  if (!s.cnt) return;
  const int64_t nf = res->n_frames; const int stop = res->stop;
  if ((int64_t)s.base_idx > nf) return;
  const uint64_t b = (uint64_t)k * HOP_SEG;
  const uint16_t *L = lists + (size_t)k * HOP_LCAP;
  uint64_t mpos = s.base_pos; uint32_t top = s.top_in;
  for (uint32_t i = 0; i < s.cnt; i++) {
    const int64_t idx = (int64_t)s.base_idx + i;
    const uint64_t p = hop_frame_pos(s, L, b, i);
    if (idx >= nf) { if (idx == nf && stop == 3) res->consumed = p; break; }     /* format change: the next batch starts at this header */
    const uint8_t *h = raw + p;
    const hop_hdr hh = hop_parse_header(h);
    const uint32_t ms = hh.fsize - hh.hdr - hh.silen, mb = ((uint32_t)h[hh.hdr] << 1) | (h[hh.hdr + 1] >> 7);
    p3_frame f;
    f.main_off = p + hh.hdr + hh.silen; f.main_pos = mpos; f.main_size = (uint16_t)ms; f.main_begin = (uint16_t)mb;
    f.nch = hh.nch; f.mode = hh.mode; f.mode_ext = hh.mext; f.sfreq = hh.sf; f.scfsi = 0; f.bitrate_kbps = (uint16_t)hh.br_kbps;
    uint32_t fl = iso ? P3_FRAME_ISO : 0u;
    if (mb > top) fl |= P3_FRAME_NODATA;                             /* reservoir rule of Get_Main_Data (pdmp3.c:1101-1120) */
    top = min(top, mb) + ms;
    if ((uint64_t)idx < warmup) { fl |= P3_FRAME_WARMUP; f.pcm_index = 0xffffffffu; }
    else { fl |= P3_FRAME_DECODE; f.pcm_index = st0.pcm_index + (uint32_t)(idx - warmup); }
    f.flags = (uint8_t)fl;
    mpos += ms;
    uint4 *dst = reinterpret_cast<uint4 *>(frames + idx); const uint4 *src = reinterpret_cast<const uint4 *>(&f);
    dst[0] = src[0]; dst[1] = src[1];
    if (idx == nf - 1) {                                             /* the last kept frame: parser state for the next batch */
      if (stop == 1) res->consumed = p + hh.fsize;
      p3_parse_state so;
      so.main_pos = mpos; so.top = top; so.pcm_index = st0.pcm_index + (uint32_t)(nf > (int64_t)warmup ? nf - warmup : 0);
      so.nch = hh.nch; so.sfreq = hh.sf;
      res->st = so; res->total_ms = mpos - st0.main_pos;
    }
  }
}

/* largest main-data span of a group of 32 consecutive frames (aligned to the batch start): K1's shared-memory window */
extern "C" __global__ void __launch_bounds__(128)
k_hop_groups(const p3_frame *__restrict__ frames, p3_hop_result *res)
{
  const int64_t nf = res->n_frames, g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, f0 = g * 32;
  if (f0 >= nf) return;
  const int64_t f1 = min(f0 + 32, nf);
  const uint64_t span = frames[f1 - 1].main_pos + frames[f1 - 1].main_size - frames[f0].main_pos;
  atomicMax(&res->maxg, (uint32_t)span);
}

/* the last 512 bytes of the batch's header-stripped main data (the reservoir the next batch may reach back into), as
 * compute_tail() of p3_cabi.cu does on the host: one warp, frames from the end backwards */
extern "C" __global__ void __launch_bounds__(32)
k_hop_tail(const uint8_t *__restrict__ raw, const p3_frame *__restrict__ frames, const p3_hop_result *__restrict__ res,
           const uint8_t *__restrict__ tail_in, uint8_t *__restrict__ tail_out)
{
  const int64_t nf = res->n_frames;
  int filled = 0;
  for (int64_t f = nf - 1; f >= 0 && filled < 512; f--) {
    const int n = frames[f].main_size, take = min(n, 512 - filled);
    const uint8_t *src = raw + frames[f].main_off + n - take;
    for (int i = threadIdx.x; i < take; i += 32) tail_out[512 - filled - take + i] = src[i];
    filled += take;
  }
  __syncwarp();
  for (int i = threadIdx.x; i < 512 - filled; i += 32) tail_out[i] = tail_in[filled + i];     /* (tail_in and tail_out are different buffers) */
}

/* The result block goes to the host by STORES into mapped page-locked memory, not by a device-to-host copy: in the streaming path
 * the D2H copy engine is busy with the previous batch's PCM (2.9 ms per 32 768 frames), and a 128-byte cudaMemcpyAsync queued
 * behind it made every hop wait for it -- twice per batch (measured: 99 of 102 ms of pdmp3_read were spent there). */
extern "C" __global__ void k_hop_publish(const p3_hop_result *__restrict__ d, p3_hop_result *__restrict__ h)
{
  const uint32_t *s = reinterpret_cast<const uint32_t *>(d); volatile uint32_t *t = reinterpret_cast<volatile uint32_t *>(h);
  for (uint32_t i = threadIdx.x; i < sizeof(p3_hop_result) / 4; i += blockDim.x) t[i] = s[i];
  __threadfence_system();
}

/* ---- host side ---- */
#include <stdio.h>
#include <string.h>
#define HCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "p3_hop: %s: %s\n", #x, cudaGetErrorString(e_)); return P3_ECUDA; } } while (0)

int p3_hop_work_ensure(p3_hop_work *w, int64_t nseg)
{
  if (!w->d_res) { HCK(cudaMalloc(&w->d_res, sizeof(p3_hop_result))); HCK(cudaHostAlloc((void **)&w->h_res, sizeof(p3_hop_result), cudaHostAllocPortable | cudaHostAllocMapped)); HCK(cudaHostGetDevicePointer((void **)&w->h_res_dev, w->h_res, 0)); }
  if (nseg <= w->cap_seg) return P3_OK;
  cudaFree(w->seg); cudaFree(w->lists); cudaFree(w->exit[0]); cudaFree(w->exit[1]); cudaFree(w->part);
  w->seg = NULL; w->lists = NULL; w->exit[0] = w->exit[1] = NULL; w->part = NULL; w->cap_seg = 0;
  const int64_t cap = nseg + nseg / 8 + 16;
  HCK(cudaMalloc(&w->seg, (size_t)cap * sizeof(hop_seg)));
  HCK(cudaMalloc(&w->lists, (size_t)cap * HOP_LCAP * sizeof(uint16_t)));
  HCK(cudaMalloc(&w->exit[0], (size_t)cap * sizeof(uint64_t)));
  HCK(cudaMalloc(&w->exit[1], (size_t)cap * sizeof(uint64_t)));
  HCK(cudaMalloc(&w->part, (size_t)(cap / 128 + 2) * sizeof(hop_part)));
  w->cap_seg = cap;
  return P3_OK;
}

void p3_hop_work_free(p3_hop_work *w)
{
  cudaFree(w->seg); cudaFree(w->lists); cudaFree(w->exit[0]); cudaFree(w->exit[1]); cudaFree(w->part); cudaFree(w->d_res);
  if (w->h_res) cudaFreeHost(w->h_res);
  memset(w, 0, sizeof *w);
}

static int64_t hop_nseg(uint64_t n) { return n ? (int64_t)((n + HOP_SEG - 1) / HOP_SEG) : 1; }

int p3_hop_count(p3_hop_work *w, cudaStream_t st, const uint8_t *d_raw, uint64_t n, const p3_parse_opts *o, const p3_parse_state *ps, int64_t frame_cap)
{
  const int64_t nseg = hop_nseg(n);
  int rc = p3_hop_work_ensure(w, nseg);
  if (rc) return rc;
  int64_t maxf = o->max_frames > 0 ? o->max_frames : 0;
  if (frame_cap > 0 && (maxf == 0 || frame_cap < maxf)) maxf = frame_cap;
  const unsigned grid = (unsigned)((nseg + 127) / 128);
  k_hop_spec<<<grid, 128, 0, st>>>(d_raw, n, o->lookahead, nseg, w->seg, w->lists, w->exit[0]);
  for (int64_t round = 1; ; round++) {
    const uint64_t *ep = w->exit[(round - 1) & 1]; uint64_t *ec = w->exit[round & 1];
    HCK(cudaMemsetAsync(&w->d_res->changed, 0, sizeof(int), st));
    k_hop_resolve<<<grid, 128, 0, st>>>(d_raw, n, o->lookahead, nseg, w->seg, w->lists, ep, ec, &w->d_res->changed);
    k_hop_agg<<<grid, 128, 0, st>>>(d_raw, nseg, w->seg, w->lists, w->part);
    k_hop_scan<<<1, 1024, 0, st>>>((int64_t)grid, w->part, w->d_res);
    k_hop_apply<<<grid, 128, 0, st>>>(nseg, w->seg, w->part, ec, *ps, w->d_res);
    k_hop_final<<<1, 1, 0, st>>>(w->seg, *ps, maxf, o->warmup_frames, w->d_res);
    k_hop_publish<<<1, 64, 0, st>>>(w->d_res, w->h_res_dev);
    HCK(cudaGetLastError());
    HCK(cudaStreamSynchronize(st));
    if (!w->h_res->changed) { w->h_res->changed = (int32_t)round; break; }    /* reports the number of rounds it took */
    if (round > nseg + 1) return P3_EINVAL;                          /* cannot happen: every round fixes at least one more segment */
  }
  return P3_OK;
}

int p3_hop_emit(p3_hop_work *w, cudaStream_t st, const uint8_t *d_raw, uint64_t n, const p3_parse_opts *o, const p3_parse_state *ps,
                p3_frame *d_frames, const uint8_t *d_tail_in, uint8_t *d_tail_out)
{
  const int64_t nseg = hop_nseg(n), nf = w->h_res->n_frames;
  const int rounds = w->h_res->changed;
  if (nf > 0) {
    k_hop_write<<<(unsigned)((nseg + 127) / 128), 128, 0, st>>>(d_raw, nseg, w->seg, w->lists, *ps, o->warmup_frames, o->iso, w->d_res, d_frames);
    k_hop_groups<<<(unsigned)(((nf + 31) / 32 + 127) / 128), 128, 0, st>>>(d_frames, w->d_res);
    if (d_tail_out) k_hop_tail<<<1, 32, 0, st>>>(d_raw, d_frames, w->d_res, d_tail_in, d_tail_out);
    HCK(cudaGetLastError());
  }
  k_hop_publish<<<1, 64, 0, st>>>(w->d_res, w->h_res_dev);
  HCK(cudaGetLastError());
  HCK(cudaStreamSynchronize(st));
  w->h_res->changed = rounds;
  return P3_OK;
}
