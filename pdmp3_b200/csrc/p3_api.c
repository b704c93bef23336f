/* p3_api.c -- the reference's streaming API (libmpg123 subset) on top of the batch decoder.
 * Plain C.  Same prototypes, return codes and observable behaviour as pdmp3.c:2351-2535; what is
 * different is WHEN work happens: pdmp3_read() parses as many buffered frames as the caller's
 * buffer can hold and sends them through the GPU as one batch instead of decoding frame by frame.
 *
 *   reference                          here
 *   pdmp3_new        2351-2353         malloc + option string (ignored by the reference)
 *   pdmp3_open_feed  2369-2384         reset cursors and the device-side filter state
 *   pdmp3_feed       2391-2423         copy into the input buffer; PDMP3_NO_SPACE if it does not fit
 *   pdmp3_read       2431-2481         flush pending PCM, then batch-decode; 1152-byte look-ahead rule (2445)
 *   pdmp3_decode     2491-2520         feed(min(free,insize)) + read, or header peek
 *   pdmp3_getformat  2526-2535         rate / channels / encoding of the last header
 *   pdmp3            2540-2589         CLI loop, raw writer only
 *
 * Documented differences (supersets, SURVEY 8b "hazards"): the input buffer is linear with
 * compaction, so feeding exactly `ring` bytes into an empty buffer works; a frame is only read when
 * it is completely buffered; a frame whose main_data_begin underflows the reservoir is emitted as
 * silence instead of being re-read (Q8).
 */
#include "../../include/pdmp3.h"
#include "../../include/pdmp3_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <fcntl.h>
#include <unistd.h>
#include <pthread.h>
#include <time.h>

static double now_ms(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }

#define P3_DEFAULT_RING 16384u              /* INBUF_SIZE (pdmp3.c:123) */
#define P3_API_CHUNK    32768               /* frames per GPU batch inside one pdmp3_read() */
#define P3_API_DIRECT   64                  /* reads of at least this many frames are decoded straight into the caller's buffer */
#define P3_API_AHEAD    4096                /* smaller reads: up to this many buffered frames are decoded at once into the handle's PCM queue */

struct pdmp3_handle {
  unsigned char *in; size_t cap, istart, iend;      /* buffered bytes are in[istart,iend) */
  size_t processed;
  int16_t *pcm; size_t pcm_cap;                     /* PCM queue: frames decoded ahead of small reads (pcm_cap frames of room) */
  size_t pend_pos, pend_end;                        /* undelivered PCM bytes of the queue: [pend_pos,pend_end) */
  int in_pinned;
  unsigned char *own;                               /* the handle's own input buffer (`in` points at it unless a fed buffer is borrowed) */
  int borrow, borrowed;                             /* feed=borrow: pdmp3_feed into an empty handle keeps the caller's pointer instead of copying */
  p3_frame *dfr[3]; p3_gc *dgc[3]; int dnext;       /* page-locked descriptor arrays, rotated over the in-flight batches */
  p3_ctx *ctx; int device; int ctx_failed; int mode; int host_sideinfo; int iso;
  int64_t batch;                                    /* frames per GPU batch of a large read (batch=<n>, default P3_API_CHUNK) */
  int host_hop; size_t bpf_est;                     /* hop=host: frame hop on the host (default: on the device); bytes per frame seen so far */
  p3_parse_state ps;
  int new_header;                                   /* 0 none yet, 1 seen, -1 reported (pdmp3.c:1318,2470,2531) */
  int nch, sfreq;
  int opened;
};

static const long k_rates[3] = {44100, 48000, 32000};

pdmp3_handle *pdmp3_new(const char *decoder, int *error)
{
  pdmp3_handle *id = (pdmp3_handle *)calloc(1, sizeof *id);
  if (!id) { if (error) *error = PDMP3_ERR; return NULL; }
  id->cap = P3_DEFAULT_RING; id->device = 0; id->mode = P3_MODE_FAST; id->batch = P3_API_CHUNK;
  if (decoder) {                                    /* "b200:ring=<bytes>,device=<n>" */
    const char *p;
    if ((p = strstr(decoder, "ring="))) { unsigned long long v = strtoull(p + 5, NULL, 10); if (v >= 4096) id->cap = (size_t)v; }
    if ((p = strstr(decoder, "device="))) id->device = atoi(p + 7);
    if (strstr(decoder, "mode=exact")) id->mode = P3_MODE_EXACT;   /* bit-identical PCM; default is FAST (<= 1 LSB) */
    if ((p = strstr(decoder, "batch="))) { long long v = atoll(p + 6); if (v >= 1024 && v <= (1 << 20)) id->batch = v; }
    if (strstr(decoder, "feed=borrow")) id->borrow = 1;             /* no copy in pdmp3_feed: the caller keeps its buffer alive and unchanged until it is consumed */
    if (strstr(decoder, "sideinfo=host")) id->host_sideinfo = 1;   /* parse the side info on the host instead of on the device */
    if (strstr(decoder, "hop=host") || id->host_sideinfo) id->host_hop = 1;   /* frame hop of large reads on the host as well (default: on the device, p3_hop.cu) */
    for (p = decoder; (p = strstr(p, "iso")) != NULL; p += 3)     /* "iso" as an option of its own: ISO 11172-3 semantics instead of the reference's quirks (P3_FRAME_ISO) */
      if ((p == decoder || p[-1] == ':' || p[-1] == ',') && (p[3] == 0 || p[3] == ',')) id->iso = 1;
  }
  if (id->cap > (1u << 20)) { id->in = (unsigned char *)p3_host_alloc_dev(id->device, id->cap); id->in_pinned = id->in != NULL; }   /* page-locked: full-speed H2D */
  if (!id->in) id->in = (unsigned char *)malloc(id->cap);
  if (!id->in) { free(id); if (error) *error = PDMP3_ERR; return NULL; }
  id->own = id->in;
  id->ps.nch = id->ps.sfreq = -1; id->nch = 2; id->sfreq = 0;
  if (error) *error = PDMP3_OK;
  return id;
}

void pdmp3_delete(pdmp3_handle *id)
{
  if (!id) return;
  if (id->ctx) p3_ctx_destroy(id->ctx);
  if (id->in_pinned) p3_host_free(id->own); else free(id->own);
  for (int k = 0; k < 3; k++) { p3_host_free(id->dfr[k]); p3_host_free(id->dgc[k]); }
  free(id->pcm); free(id);
}

int pdmp3_open_feed(pdmp3_handle *id)
{
  if (!id) return PDMP3_ERR;
  id->istart = id->iend = 0; id->processed = 0; id->new_header = 0;
  id->in = id->own; id->borrowed = 0;
  id->pend_pos = id->pend_end = 0; id->bpf_est = 0;
  memset(&id->ps, 0, sizeof id->ps); id->ps.nch = id->ps.sfreq = -1;
  if (id->ctx) p3_ctx_reset(id->ctx);               /* hsynth_init / synth_init / g_main_data_top (pdmp3.c:2377-2379) */
  id->opened = 1;
  return PDMP3_OK;
}

static size_t in_filled(const pdmp3_handle *id) { return id->iend - id->istart; }
static size_t in_free(const pdmp3_handle *id) { return id->cap - in_filled(id); }

/* large feeds are copied by several threads (a single core moves ~10 GB/s, the PCIe link 50+): one per online core, at most 32 */
typedef struct { unsigned char *d; const unsigned char *s; size_t n; } cpjob;
static void *cp_worker(void *a) { cpjob *j = (cpjob *)a; memcpy(j->d, j->s, j->n); return NULL; }
static void big_memcpy(unsigned char *d, const unsigned char *s, size_t n)
{
  enum { NTMAX = 32 };
  if (n < ((size_t)32 << 20)) { memcpy(d, s, n); return; }
  long nc = sysconf(_SC_NPROCESSORS_ONLN);
  const int NT = nc < 2 ? 2 : nc > NTMAX ? NTMAX : (int)nc;
  pthread_t th[NTMAX]; cpjob jb[NTMAX]; int ok[NTMAX];
  for (int t = 0; t < NT; t++) {
    size_t lo = n * t / NT, hi = n * (t + 1) / NT;
    jb[t].d = d + lo; jb[t].s = s + lo; jb[t].n = hi - lo;
    ok[t] = pthread_create(&th[t], NULL, cp_worker, &jb[t]) == 0;
    if (!ok[t]) cp_worker(&jb[t]);
  }
  for (int t = 0; t < NT; t++) if (ok[t]) pthread_join(th[t], NULL);
}

int pdmp3_feed(pdmp3_handle *id, const unsigned char *in, size_t size)
{
  if (!(id && in && size)) return PDMP3_ERR;
  if (id->borrow && in_filled(id) == 0) {             /* feed=borrow: decode straight out of the caller's buffer (page-locked for full-speed uploads) */
    id->in = (unsigned char *)(uintptr_t)in; id->istart = 0; id->iend = size; id->borrowed = 1;
    return PDMP3_OK;
  }
  if (id->borrowed) {                                 /* more data while a borrowed buffer is not used up: its rest moves into the handle's own buffer */
    const size_t rest = in_filled(id);
    if (rest + size > id->cap) return PDMP3_NO_SPACE;
    memmove(id->own, id->in + id->istart, rest);
    id->in = id->own; id->istart = 0; id->iend = rest; id->borrowed = 0;
  }
  if (size > in_free(id)) return PDMP3_NO_SPACE;
  if (id->iend + size > id->cap) {                  /* compact */
    memmove(id->in, id->in + id->istart, in_filled(id));
    id->iend -= id->istart; id->istart = 0;
  }
  big_memcpy(id->in + id->iend, in, size);
  id->iend += size;
  return PDMP3_OK;
}

static int ensure_ctx(pdmp3_handle *id)
{
  if (id->ctx) return PDMP3_OK;
  if (id->ctx_failed) return PDMP3_ERR;
  if (p3_ctx_create(id->device, &id->ctx) != P3_OK) {
    fprintf(stderr, "pdmp3_b200: %s\n", p3_last_error());
    id->ctx_failed = 1; id->ctx = NULL;
    return PDMP3_ERR;
  }
  p3_ctx_set_mode(id->ctx, id->mode);
  return PDMP3_OK;
}

int pdmp3_read(pdmp3_handle *id, unsigned char *outmemory, size_t outsize, size_t *done)
{
  if (!(id && outmemory && outsize && done)) return PDMP3_ERR;
  int res = PDMP3_ERR, inflight = 0;
  const int trace = getenv("P3_TRACE") != NULL; double t_parse = 0, t_decode = 0, t_sync = 0; int nb = 0;
  *done = 0;
  if (id->pend_pos < id->pend_end) {                /* PCM decoded earlier: the rest of a frame (pdmp3.c:2437-2442) and the frames decoded ahead */
    size_t n = id->pend_end - id->pend_pos; if (n > outsize) n = outsize;
    memcpy(outmemory, (unsigned char *)id->pcm + id->pend_pos, n);
    id->pend_pos += n; outmemory += n; outsize -= n; *done += n;
    res = PDMP3_OK;
  }
  while (outsize) {
    if (in_filled(id) < 2 * 576) { res = PDMP3_NEED_MORE; break; }          /* pdmp3.c:2445,2466 */
    size_t fbytes = 1152 * sizeof(int16_t) * (size_t)(id->nch == 1 ? 1 : 2);
    /* Large reads: whole frames go straight into the caller's buffer, P3_API_CHUNK frames per GPU batch, the batches
     * double-buffered (upload / kernels / download overlap, and so does the parsing of the next batch).
     * Small reads (the reference CLI asks for 16 KiB = 3.6 frames at a time, pdmp3.c:2564): every frame that is
     * already buffered and decodable goes through the GPU as ONE batch into the handle's PCM queue and is handed out
     * piecewise (the reference's ostart cursor, pdmp3.c:2317-2344, extended from one frame to a queue).  Which
     * frames are decoded, the bytes delivered and the return codes are the reference's; only the moment differs. */
    int direct = outsize >= P3_API_DIRECT * fbytes;
    int64_t want = direct ? (int64_t)(outsize / fbytes) : P3_API_AHEAD;
    if (direct && !id->host_hop && want > id->batch) want = id->batch;
    else if (want > P3_API_CHUNK && !(direct && !id->host_hop)) want = P3_API_CHUNK;
    if (direct && want >= 1024 && !id->host_hop) {
      /* Large reads: nothing is parsed on the host.  A window of the buffered bytes goes to the device, the frame hop
       * (Search_Header / Read_Header, pdmp3.c:1252-1340) runs there (p3_hop.cu) and tells how many frames it found and
       * where the next window starts; kernels and the PCM download of that batch are queued behind it while this loop
       * already uploads the next window. */
      int nch0 = 2, sf0 = 0;
      const size_t avail = in_filled(id);
      const int fh = p3_find_header(id->in + id->istart, avail, &nch0, &sf0);  /* format of the first frame: sizes the PCM slots */
      if (fh <= 0) { res = fh < 0 ? PDMP3_ERR : PDMP3_NEED_MORE; break; }      /* no header within a frame's length: pdmp3.c:1337 */
      if (nch0 != (id->nch == 1 ? 1 : 2)) { id->nch = nch0; continue; }       /* re-plan with the right frame size */
      if (ensure_ctx(id) != PDMP3_OK) { res = PDMP3_ERR; break; }
      size_t win = (size_t)want * (id->bpf_est ? id->bpf_est + id->bpf_est / 32 + 1 : 1441) + 8192;
      uint32_t look = 2 * 576;                                                 /* the 1152-byte rule counts from the end of what is buffered (pdmp3.c:2445) */
      if (win >= avail || avail - win < 8192) win = avail; else look = 0;
      p3_parse_opts po = {want, look, 0, 0, 1u, (uint32_t)id->iso};
      p3_parse_state ps = id->ps; ps.pcm_index = 0;                           /* slots restart at 0 for every batch */
      p3_parsed pb;
      double t0 = trace ? now_ms() : 0;
      int rc = p3_decode_raw_async(id->ctx, id->in + id->istart, win, &po, &ps, &pb, (int16_t *)outmemory);
      if (trace) { t_decode += now_ms() - t0; nb++; }
      if (rc != P3_OK) { fprintf(stderr, "pdmp3_b200: %s\n", p3_last_error()); res = PDMP3_ERR; break; }
      if (pb.n_frames == 0) { res = pb.stop == 2 ? PDMP3_ERR : PDMP3_NEED_MORE; break; }
      inflight = 1;
      id->ps = ps; id->ps.pcm_index = 0; id->sfreq = ps.sfreq;
      id->bpf_est = (size_t)((pb.consumed + (uint64_t)pb.n_frames - 1) / (uint64_t)pb.n_frames);
      if (!id->new_header) id->new_header = 1;                                 /* pdmp3.c:1318 */
      id->istart += pb.consumed; id->processed += pb.consumed;
      { size_t n = (size_t)pb.n_frames * fbytes; outmemory += n; outsize -= n; *done += n; }
      res = PDMP3_OK;
      continue;
    }
    p3_parse_opts po = {want, 2 * 576, want >= 8192 ? 4 : 1, 0, id->host_sideinfo ? 0u : 1u, (uint32_t)id->iso};   /* side info: parsed on the device */
    p3_parse_state ps = id->ps;
    p3_parsed pb;
    p3_frame *dfr = NULL; p3_gc *dgc = NULL;
    if (direct && want >= 1024) {                                             /* big batches: descriptors in page-locked memory, no allocation */
      const int k = id->dnext;                                                /* rotated only once the batch has been submitted (below) */
      if (!id->dfr[k]) { id->dfr[k] = (p3_frame *)p3_host_alloc(sizeof(p3_frame) * P3_API_CHUNK); id->dgc[k] = (p3_gc *)p3_host_alloc(sizeof(p3_gc) * 4 * P3_API_CHUNK); }
      if (id->dfr[k] && id->dgc[k]) { dfr = id->dfr[k]; dgc = id->dgc[k]; }
    }
    double t0 = trace ? now_ms() : 0;
    if (p3_parse_into(id->in + id->istart, in_filled(id), &po, &ps, &pb, dfr, dgc, P3_API_CHUNK) != P3_OK) { res = PDMP3_ERR; break; }
    if (trace) { t_parse += now_ms() - t0; nb++; }
    if (pb.n_frames == 0) {
      int stop = pb.stop;
      p3_parsed_free(&pb);
      res = stop == 2 ? PDMP3_ERR : PDMP3_NEED_MORE;                         /* no header within a frame's length: pdmp3.c:1337 */
      break;
    }
    if (ensure_ctx(id) != PDMP3_OK) { p3_parsed_free(&pb); res = PDMP3_ERR; break; }
    int nch = pb.frames[0].nch;
    size_t fb2 = 1152 * sizeof(int16_t) * (size_t)nch;
    if (fb2 != fbytes) {                                                      /* channel count changed: re-plan with the right frame size */
      id->nch = nch; p3_parsed_free(&pb); continue;
    }
    int16_t *target = (int16_t *)outmemory;
    if (!direct) {
      if (id->pcm_cap < (size_t)pb.n_frames) {
        free(id->pcm); id->pcm_cap = (size_t)pb.n_frames + 16;
        id->pcm = (int16_t *)malloc(id->pcm_cap * 1152 * 2 * sizeof(int16_t));
        if (!id->pcm) { id->pcm_cap = 0; p3_parsed_free(&pb); res = PDMP3_ERR; break; }
      }
      target = id->pcm;
    }
    for (int64_t f = 0; f < pb.n_frames; f++) pb.frames[f].pcm_index = (uint32_t)f;   /* slots restart at 0 for every batch */
    int64_t nfr = pb.n_frames; uint64_t used = pb.consumed; int sf_last = pb.frames[pb.n_frames - 1].sfreq;
    const uint64_t upl = in_filled(id) < used + 64 ? in_filled(id) : used + 64;   /* bytes this batch can touch */
    t0 = trace ? now_ms() : 0;
    int rc = direct ? p3_decode_batch_async(id->ctx, id->in + id->istart, upl, &pb, target)
                    : p3_decode_batch(id->ctx, id->in + id->istart, upl, &pb, target, NULL);
    if (trace) t_decode += now_ms() - t0;
    p3_parsed_free(&pb);                                                     /* (the async call took the arrays over) */
    if (rc != P3_OK) { fprintf(stderr, "pdmp3_b200: %s\n", p3_last_error()); res = PDMP3_ERR; break; }
    inflight |= direct;
    if (dfr) id->dnext = (id->dnext + 1) % 3;                                /* this descriptor buffer now belongs to an in-flight batch */
    id->ps = ps; id->ps.pcm_index = 0;
    id->sfreq = sf_last;
    if (!id->new_header) id->new_header = 1;                                 /* pdmp3.c:1318 */
    id->istart += used; id->processed += used;
    if (direct) {
      size_t n = (size_t)nfr * fbytes;
      outmemory += n; outsize -= n; *done += n;
    } else {
      const size_t have = (size_t)nfr * fbytes, n = have < outsize ? have : outsize;
      memcpy(outmemory, id->pcm, n);
      id->pend_pos = n; id->pend_end = have;
      *done += n; outmemory += n; outsize -= n;
    }
    res = PDMP3_OK;
  }
  { double t0 = trace ? now_ms() : 0;
    if (inflight && p3_batch_sync(id->ctx) != P3_OK) { fprintf(stderr, "pdmp3_b200: %s\n", p3_last_error()); res = PDMP3_ERR; }
    if (trace) { t_sync = now_ms() - t0; fprintf(stderr, "pdmp3_read: %d batches, parse %.1f ms, enqueue (incl. waiting for a slot) %.1f ms, final sync %.1f ms\n", nb, t_parse, t_decode, t_sync); } }
  if (id->istart == id->iend) { id->istart = id->iend = 0; id->in = id->own; id->borrowed = 0; }
  if (id->new_header == 1 && res == PDMP3_OK) res = PDMP3_NEW_FORMAT;       /* pdmp3.c:2470-2472 */
  return res;
}

int pdmp3_decode(pdmp3_handle *id, const unsigned char *in, size_t insize, unsigned char *out, size_t outsize, size_t *done)
{
  if (!id || !done) return PDMP3_ERR;
  size_t fr = in_free(id);
  int res;
  *done = 0;
  if (fr > insize) fr = insize;                     /* silently drops what does not fit (pdmp3.c:2497-2498) */
  res = pdmp3_feed(id, in, fr);
  if (res == PDMP3_OK) {
    if (out && outsize) { size_t avail = 0; res = pdmp3_read(id, out, outsize, &avail); *done = avail; }
    else if (id->processed == 0) {                  /* header peek (pdmp3.c:2507-2516) */
      int nch, sf;
      res = PDMP3_NEED_MORE;
      if (in_filled(id) > 4) {
        int r = p3_find_header(id->in + id->istart, in_filled(id), &nch, &sf);
        if (r == 1) { id->nch = nch; id->sfreq = sf; if (!id->new_header) id->new_header = 1; res = PDMP3_OK; }
        else if (r < 0) res = PDMP3_ERR;
      }
      if (id->new_header == 1) res = PDMP3_NEW_FORMAT;
    }
  }
  return res;
}

int pdmp3_getformat(pdmp3_handle *id, long *rate, int *channels, int *encoding)
{
  if (!(id && rate && channels && encoding)) return PDMP3_ERR;
  *encoding = PDMP3_ENC_SIGNED_16;
  *rate = k_rates[id->sfreq % 3];
  *channels = id->nch == 1 ? 1 : 2;
  id->new_header = -1;
  return PDMP3_OK;
}

/* CLI loop of the reference (pdmp3.c:2540-2589): read 16 KiB of PCM at a time, feed 4096 bytes on
 * NEED_MORE.  Output: <file>.raw, or stdout for "-" (the OUTPUT_RAW writer, pdmp3.c:2236-2257). */
void pdmp3(char * const *mp3s)
{
  unsigned char out[16384], in[4096];
  pdmp3_handle *id;
  if (!mp3s) return;
  if (*mp3s && !strncmp("/dev/dsp", *mp3s, 8)) mp3s++;        /* OSS device argument: accepted, ignored */
  id = pdmp3_new(NULL, NULL);
  if (!id) { fputs("Cannot open stream API (out of memory)", stderr); exit(0); }
  while (*mp3s) {
    const char *filename = *mp3s++;
    FILE *fp = !strcmp(filename, "-") ? stdin : fopen(filename, "r");
    int fd, res; size_t done;
    if (!fp) { fputs("Cannot open file\n", stderr); exit(0); }
    if (strcmp(filename, "-")) { char fname[1024]; snprintf(fname, sizeof fname - 1, "%s.raw", filename); fd = open(fname, O_WRONLY | O_CREAT | O_TRUNC, 0666); if (fd == -1) { perror(fname); exit(-1); } }
    else fd = 1;
    pdmp3_open_feed(id);
    while ((res = pdmp3_read(id, out, sizeof out, &done)) != PDMP3_ERR) {
      if (done && write(fd, out, done) != (ssize_t)done) { fputs("Unable to write raw data\n", stderr); exit(-1); }
      if (res == PDMP3_NEED_MORE) {
        size_t n = fread(in, 1, sizeof in, fp);
        if (!n) break;
        pdmp3_feed(id, in, n);
      }
    }
    if (fd != 1) close(fd);
    if (fp != stdin) fclose(fp);
  }
  pdmp3_delete(id);
}
