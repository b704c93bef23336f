/* p3_cabi.cu -- the thin C-ABI between the plain-C host side and the sm_100a kernels
 * (include/pdmp3_b200.h).  Owns device memory, the stream, the carried decoder state and the
 * kernel launch sequence K1..K4, which is the device-side Decode_L3 (pdmp3.c:1024-1060). */
#include "p3_device.cuh"
#include "p3_kernels.h"
#include "p3_hop.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>

extern "C" int p3_fused_upload_consts(const p3_tables *T, const float *dct4);
extern "C" size_t p3_synthw_smem_bytes(void);
extern "C" int p3_synthw_warps_per_cta(void);
extern "C" int p3_synthw_warps_per_sm(void);
#define SW_DECL(NAME) extern "C" __global__ void NAME(const p3_frame *frames, const p3_gc *gcs, const p3_tables *T, int64_t f_first, int64_t f_end, int frames_per_warp, \
    const int16_t *is_in, const int32_t *count1, const uint8_t *scf, const p3_state *st_in, p3_state *st_out, int16_t *pcm, const float *pow43s, int classify)
SW_DECL(k_synth_warp); SW_DECL(k_synth_warp_same); SW_DECL(k_synth_warp_lean); SW_DECL(k_synth_warp_iso); SW_DECL(k_synth_warp_iso_same); SW_DECL(k_synth_warp_iso_lean);
extern "C" __global__ void k_synth_fast(const p3_frame *frames, const p3_gc *gcs, const p3_tables *T, int64_t f_first, int64_t f_end, int frames_per_cta,
    const int16_t *is_in, const int32_t *count1, const uint8_t *scf, const p3_state *st_in, p3_state *st_out, int16_t *pcm, float *xr_tap, float *y_tap);

#include <time.h>
static double now_ms(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }
static double g_tr_wait, g_tr_stage, g_tr_launch; static int g_trace = -1;     /* P3_TRACE=1: where the host time of the async path goes */
static thread_local char g_err[512];
static int fail(int code, const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); return code; }
extern "C" const char *p3_last_error(void) { return g_err; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(P3_ECUDA, "%s: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

struct dbuf { void *p; size_t cap; };
static int ensure(dbuf *b, size_t n)
{
  if (n <= b->cap) return P3_OK;
  if (b->p) cudaFree(b->p);
  b->p = NULL; b->cap = 0;
  size_t want = n + n / 8 + 256;
  cudaError_t e = cudaMalloc(&b->p, want);
  if (e != cudaSuccess) return fail(P3_ENOMEM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
  b->cap = want;
  return P3_OK;
}

/* One in-flight batch: its inputs, its PCM and the events that order H2D -> kernels -> D2H.  Two slots
 * let the upload of batch k+1 and the download of batch k-1 overlap the kernels of batch k. */
struct p3_slot {
  dbuf raw, frames, gcs, pcm, ms;         /* ms: compact main-data stream (k_compact -> k_huffman) */
  const uint8_t *raw_dev;                 /* the staged byte stream the kernels read: raw.p, or a caller-owned device buffer (p3_sharded_decode) */
  uint64_t ms_bytes;
  uint8_t *d_tail; uint8_t *h_tail;       /* 512 main-data bytes in front of the batch (pinned host copy) */
  int *d_any_empty; int hop_only;         /* device side-info parser: flag for the Q6 chain; pending for this slot */
  int iso;                                /* the staged batch is flagged P3_FRAME_ISO */
  cudaEvent_t h2d_done, compute_done, d2h_done;
  p3_parsed keep; int have_keep;          /* host descriptors owned until the upload has completed */
  int busy;
};

struct p3_ctx {
  int device, mode;
  cudaStream_t stream, s_h2d, s_d2h;     /* kernels | uploads | downloads */
  cudaEvent_t ev[10];
  p3_tables *d_tables;
  p3_state *d_state[2]; int cur;          /* double-buffered carried state: kernels read [cur], write [cur^1] */
  uint8_t h_tail[512];                    /* last 512 main-data bytes before the next batch */
  p3_slot slot[2]; int cur_slot;
  dbuf is16, count1, scf, xr, y;          /* intermediates, only touched by the kernel streams */
  int n_sm;
  /* current (most recently uploaded) batch */
  int64_t n_frames, n_pcm_frames; uint32_t nch; uint64_t raw_bytes;
  uint32_t k1_smem_words; int64_t chunk_frames;
  int launches; int launches_parse; int taps; int fpc;
  float *d_pow43s; int synth_kernel;      /* signed |is|^(4/3) table (k_synth_warp); 0 = pick, 1 = always k_synth_fast, 2 = k_synth_warp without content classes */
  uint8_t next_tail[512]; int have_next_tail;
  /* device-side frame hop (p3_hop.cu): scratch, and the reservoir carry kept on the device for batches staged from raw bytes */
  p3_hop_work hop;
  uint8_t *d_tail_cur, *d_tail_nxt; int tail_on_device, have_next_tail_dev;
  float hop_ms;
  /* K1 / synthesis overlap (FAST mode, run_all_overlap): the batch goes through in chunks of ov_chunk frames, K0 + K1 of chunk
   * i+1 on their own stream while the synthesis of chunk i runs on c->stream; OV_NBUF sets of intermediates */
  cudaStream_t s_k1[2]; cudaEvent_t ov_start, ov_k1_done[3], ov_synth_done[3];
  int64_t ov_chunk; int ov_prio, ov_synth_pad, ov_k1_pad;
};
#define OV_NBUF 3
#define OV_MAX_PAD (64 * 1024)                              /* largest shared-memory padding p3_ctx_set_overlap accepts (caps the CTAs per SM of a kernel) */

extern "C" void *p3_host_alloc(size_t bytes) { void *p = NULL; return cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess ? p : NULL; }
/* same, after selecting `device`: the allocation then does not create a primary context on device 0 for a decoder that runs on device N */
extern "C" void *p3_host_alloc_dev(int device, size_t bytes) { if (cudaSetDevice(device) != cudaSuccess) return NULL; return p3_host_alloc(bytes); }
extern "C" void p3_host_free(void *p) { if (p) cudaFreeHost(p); }

static int ctx_init(p3_ctx *c, int n_sm);
extern "C" void p3_ctx_destroy(p3_ctx *c);

extern "C" int p3_ctx_create(int device, p3_ctx **out)
{
  int ndev = 0;
  if (!out) return fail(P3_EINVAL, "null out");
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(P3_ENODEV, "no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(P3_ENODEV, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(P3_ENODEV, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
  CK(cudaSetDevice(device));
  p3_ctx *c = (p3_ctx *)calloc(1, sizeof *c);
  if (!c) return fail(P3_ENOMEM, "calloc");
  c->device = device; c->mode = P3_MODE_EXACT;
  const int rc = ctx_init(c, prop.multiProcessorCount);
  if (rc != P3_OK) { char keep[sizeof g_err]; memcpy(keep, g_err, sizeof keep); p3_ctx_destroy(c); memcpy(g_err, keep, sizeof keep); return rc; }   /* nothing leaks; the caller may retry */
  *out = c;
  return P3_OK;
}

/* dynamic shared memory the synthesis kernels may be launched with: their own + the tuning pad of p3_ctx_set_overlap, within the 227 KB a CTA can have */
static int synthw_smem_cap(void) { const size_t v = p3_synthw_smem_bytes() + OV_MAX_PAD; return (int)(v < 227 * 1024 ? v : 227 * 1024); }

static int ctx_init(p3_ctx *c, int n_sm)
{
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));          /* lo: least (what c->stream has), hi: greatest */
    CK(cudaStreamCreateWithPriority(&c->s_k1[0], cudaStreamNonBlocking, lo));     /* ov_prio 0: K1 and the synthesis at the same priority */
    CK(cudaStreamCreateWithPriority(&c->s_k1[1], cudaStreamNonBlocking, hi < lo ? lo - 1 : lo));   /* ov_prio 1: K1's CTAs are placed first */
    CK(cudaEventCreateWithFlags(&c->ov_start, cudaEventDisableTiming));
    for (int i = 0; i < OV_NBUF; i++) { CK(cudaEventCreateWithFlags(&c->ov_k1_done[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->ov_synth_done[i], cudaEventDisableTiming)); }
  }
  CK(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  c->n_sm = n_sm;
  for (int i = 0; i < 2; i++) {
    p3_slot *sl = &c->slot[i];
    CK(cudaEventCreateWithFlags(&sl->h2d_done, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&sl->compute_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&sl->d2h_done, cudaEventDisableTiming));
    CK(cudaMalloc(&sl->d_tail, 512)); CK(cudaMemset(sl->d_tail, 0, 512));
    CK(cudaMalloc(&sl->d_any_empty, sizeof(int)));
    CK(cudaHostAlloc((void **)&sl->h_tail, 512, cudaHostAllocDefault)); memset(sl->h_tail, 0, 512);
  }
  for (int i = 0; i < 10; i++) CK(cudaEventCreate(&c->ev[i]));
  CK(cudaMalloc(&c->d_tail_cur, 512)); CK(cudaMemset(c->d_tail_cur, 0, 512));
  CK(cudaMalloc(&c->d_tail_nxt, 512)); CK(cudaMemset(c->d_tail_nxt, 0, 512));
  CK(cudaMalloc(&c->d_tables, sizeof(p3_tables)));
  CK(cudaMemcpy(c->d_tables, p3_tables_get(), sizeof(p3_tables), cudaMemcpyHostToDevice));
  for (int i = 0; i < 2; i++) { CK(cudaMalloc(&c->d_state[i], sizeof(p3_state))); CK(cudaMemset(c->d_state[i], 0, sizeof(p3_state))); }
  CK(cudaFuncSetAttribute(k_huffman, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_polyphase, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  c->chunk_frames = 1 << 18; c->fpc = 32;
  { const char *e = getenv("P3_FPC"); if (e && atoi(e) >= 1) c->fpc = atoi(e); }   /* tuning: frames per run (warp of k_synth_warp / CTA of k_synth_fast) */
  CK(cudaFuncSetAttribute(k_synth_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, synthw_smem_cap()));
  CK(cudaFuncSetAttribute(k_synth_warp_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, synthw_smem_cap()));
  CK(cudaFuncSetAttribute(k_synth_warp_same, cudaFuncAttributeMaxDynamicSharedMemorySize, synthw_smem_cap()));
  CK(cudaFuncSetAttribute(k_synth_warp_iso_same, cudaFuncAttributeMaxDynamicSharedMemorySize, synthw_smem_cap()));
  CK(cudaFuncSetAttribute(k_synth_warp_iso, cudaFuncAttributeMaxDynamicSharedMemorySize, synthw_smem_cap()));
  CK(cudaFuncSetAttribute(k_synth_warp_iso_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, synthw_smem_cap()));
  {
    /* pow43s[8207 + v] = sign(v) * |v|^(4/3): requantization without abs / sign fix-up (pdmp3.c:2125-2132) */
    static float h[2 * 8207 + 1];
    const p3_tables *T = p3_tables_get();
    for (int v = -8207; v <= 8207; v++) h[8207 + v] = v < 0 ? -T->pow43[-v] : T->pow43[v];
    CK(cudaMalloc(&c->d_pow43s, sizeof h));
    CK(cudaMemcpy(c->d_pow43s, h, sizeof h, cudaMemcpyHostToDevice));
    const char *e = getenv("P3_SYNTH"); c->synth_kernel = (e && !strcmp(e, "cta")) ? 1 : 0;
  }
  {
    static float dct4[18 * 18];
    for (int k = 0; k < 18; k++) for (int m = 0; m < 18; m++) dct4[k * 18 + m] = (float)cos(3.14159265358979323846 / 18.0 * (k + 0.5) * (m + 0.5));
    if (p3_fused_upload_consts(p3_tables_get(), dct4) != 0) return fail(P3_ECUDA, "constant upload failed");
  }
  return P3_OK;
}

static void slot_release(p3_slot *sl)
{
  if (sl->busy) { cudaEventSynchronize(sl->d2h_done); sl->busy = 0; }
  if (sl->have_keep) { p3_parsed_free(&sl->keep); sl->have_keep = 0; }
}

extern "C" void p3_ctx_destroy(p3_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);                                /* (every member may still be NULL: p3_ctx_create cleans up through here) */
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->s_h2d) cudaStreamSynchronize(c->s_h2d);
  if (c->s_d2h) cudaStreamSynchronize(c->s_d2h);
  for (int i = 0; i < 2; i++) {
    p3_slot *sl = &c->slot[i];
    slot_release(sl);
    dbuf *bs[] = {&sl->raw, &sl->frames, &sl->gcs, &sl->pcm, &sl->ms};
    for (dbuf *b : bs) if (b->p) cudaFree(b->p);
    cudaFree(sl->d_tail); cudaFree(sl->d_any_empty); if (sl->h_tail) cudaFreeHost(sl->h_tail);
    if (sl->h2d_done) cudaEventDestroy(sl->h2d_done);
    if (sl->compute_done) cudaEventDestroy(sl->compute_done);
    if (sl->d2h_done) cudaEventDestroy(sl->d2h_done);
  }
  dbuf *bs[] = {&c->is16, &c->count1, &c->scf, &c->xr, &c->y};
  for (dbuf *b : bs) if (b->p) cudaFree(b->p);
  cudaFree(c->d_tables); cudaFree(c->d_state[0]); cudaFree(c->d_state[1]); cudaFree(c->d_pow43s);
  cudaFree(c->d_tail_cur); cudaFree(c->d_tail_nxt); p3_hop_work_free(&c->hop);
  for (int i = 0; i < 10; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  for (int i = 0; i < 2; i++) if (c->s_k1[i]) cudaStreamDestroy(c->s_k1[i]);
  if (c->ov_start) cudaEventDestroy(c->ov_start);
  for (int i = 0; i < OV_NBUF; i++) { if (c->ov_k1_done[i]) cudaEventDestroy(c->ov_k1_done[i]); if (c->ov_synth_done[i]) cudaEventDestroy(c->ov_synth_done[i]); }
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  cudaGetLastError();
  free(c);
}

extern "C" int p3_ctx_reset(p3_ctx *c)
{
  if (!c) return fail(P3_EINVAL, "null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->s_h2d)); CK(cudaStreamSynchronize(c->stream)); CK(cudaStreamSynchronize(c->s_d2h));
  for (int i = 0; i < 2; i++) { slot_release(&c->slot[i]); CK(cudaMemset(c->d_state[i], 0, sizeof(p3_state))); }
  memset(c->h_tail, 0, 512); c->have_next_tail = 0;
  CK(cudaMemset(c->d_tail_cur, 0, 512)); c->tail_on_device = 0; c->have_next_tail_dev = 0;
  return P3_OK;
}

extern "C" int p3_ctx_set_overlap(p3_ctx *c, int64_t chunk_frames, int prio, int synth_pad_bytes, int k1_pad_bytes);
extern "C" int p3_ctx_set_mode(p3_ctx *c, int mode)
{
  if (!c || (mode != P3_MODE_EXACT && mode != P3_MODE_FAST)) return fail(P3_EINVAL, "bad mode");
  c->mode = mode;
  /* frames per kernel-sequence launch: EXACT keeps fp32 intermediates of every stage in HBM (28 KB/frame), FAST only the int16 spectra */
  c->chunk_frames = mode == P3_MODE_FAST ? (1 << 21) : (1 << 18);
  { const char *e = getenv("P3_CHUNK"); if (e && atoi(e) >= 64) c->chunk_frames = atoi(e); }
  /* K1's shared-memory window is sized over groups of K1_FPB frames aligned to the batch start (stage_batch); a launch
   * sequence must start on such a boundary, or a group could span more bytes than the window */
  c->chunk_frames -= c->chunk_frames % K1_FPB;
  { const char *e = getenv("P3_OVERLAP");                  /* "chunk[,prio[,synth_pad[,k1_pad]]]" */
    if (e) { long long ch = 0; int pr = 0, sp = 0, kp = 0; sscanf(e, "%lld,%d,%d,%d", &ch, &pr, &sp, &kp); int rc = p3_ctx_set_overlap(c, ch, pr, sp, kp); if (rc) return rc; } }
  return P3_OK;
}
/* K1 / synthesis overlap of FAST mode (run_all_overlap): chunk_frames = 0 turns it off; prio 1 puts K1's stream above the kernel
 * stream; the pads add unused dynamic shared memory to the synthesis kernels / to K1, which caps their CTAs per SM (tuning). */
extern "C" int p3_ctx_set_overlap(p3_ctx *c, int64_t chunk_frames, int prio, int synth_pad_bytes, int k1_pad_bytes)
{
  if (!c || chunk_frames < 0 || prio < 0 || prio > 1 || synth_pad_bytes < 0 || synth_pad_bytes > OV_MAX_PAD || k1_pad_bytes < 0 || k1_pad_bytes > OV_MAX_PAD)
    return fail(P3_EINVAL, "bad overlap parameters");
  if (chunk_frames && chunk_frames < 4 * K1_FPB) chunk_frames = 4 * K1_FPB;
  c->ov_chunk = chunk_frames - chunk_frames % K1_FPB;      /* a launch sequence starts on a K1 group boundary (see p3_ctx_set_mode) */
  c->ov_prio = prio; c->ov_synth_pad = synth_pad_bytes; c->ov_k1_pad = k1_pad_bytes;
  return P3_OK;
}
extern "C" int p3_ctx_set_taps(p3_ctx *c, int on) { if (!c) return P3_EINVAL; c->taps = on; return P3_OK; }
extern "C" int p3_ctx_set_frames_per_cta(p3_ctx *c, int n) { if (!c || n < 1) return P3_EINVAL; c->fpc = n; return P3_OK; }
extern "C" int p3_ctx_set_synth_kernel(p3_ctx *c, int which) { if (!c || which < 0 || which > 2) return P3_EINVAL; c->synth_kernel = which; return P3_OK; }
extern "C" void *p3_ctx_stream(p3_ctx *c) { return c ? (void *)c->stream : NULL; }
extern "C" int p3_kernel_launch_count(p3_ctx *c) { return c ? c->launches : 0; }

/* last 512 bytes of the batch's header-stripped main-data stream (host side, tiny) */
static void compute_tail(const uint8_t *raw, const p3_parsed *b, const uint8_t prev_tail[512], uint8_t out[512])
{
  int filled = 0;
  for (int64_t f = b->n_frames - 1; f >= 0 && filled < 512; f--) {
    int n = b->frames[f].main_size, take = n < 512 - filled ? n : 512 - filled;
    memcpy(out + 512 - filled - take, raw + b->frames[f].main_off + n - take, (size_t)take);
    filled += take;
  }
  if (filled < 512) memcpy(out, prev_tail + filled, (size_t)(512 - filled));
}

/* the batch just decoded becomes history: its last 512 main-data bytes are the reservoir of the next batch */
static void commit_tail(p3_ctx *c)
{
  if (c->have_next_tail) { memcpy(c->h_tail, c->next_tail, 512); c->have_next_tail = 0; c->tail_on_device = 0; }
  if (c->have_next_tail_dev) { uint8_t *t = c->d_tail_cur; c->d_tail_cur = c->d_tail_nxt; c->d_tail_nxt = t; c->have_next_tail_dev = 0; c->tail_on_device = 1; }
}

/* device buffers of a batch of nf frames (npcm of them with a PCM slot), K1's shared-memory window for groups spanning at
 * most maxg bytes of main data, the compact main-data stream of total_ms bytes */
static int size_batch(p3_ctx *c, p3_slot *sl, int64_t nf, int64_t npcm, uint64_t maxg, uint64_t total_ms, cudaStream_t st)
{
  int rc;
  if ((rc = ensure(&sl->frames, (size_t)(nf ? nf : 1) * sizeof(p3_frame)))) return rc;
  if ((rc = ensure(&sl->gcs, (size_t)(nf ? nf : 1) * 4 * sizeof(p3_gc)))) return rc;
  if ((rc = ensure(&sl->pcm, (size_t)(npcm ? npcm : 1) * 1152 * c->nch * sizeof(int16_t)))) return rc;
  int64_t cf = nf < c->chunk_frames ? nf : c->chunk_frames;
  if (c->ov_chunk > 0 && OV_NBUF * c->ov_chunk > cf) cf = OV_NBUF * c->ov_chunk;   /* the overlapped pipeline keeps OV_NBUF chunks of intermediates */
  if (cf) {
    if ((rc = ensure(&c->is16, (size_t)cf * 4 * 576 * 2))) return rc;
    if ((rc = ensure(&c->count1, (size_t)cf * 4 * 4))) return rc;
    if ((rc = ensure(&c->scf, (size_t)cf * 4 * P3_SCF_STRIDE))) return rc;
    if (c->mode == P3_MODE_EXACT || c->taps) {
      if ((rc = ensure(&c->xr, (size_t)cf * 4 * 576 * 4))) return rc;
      if ((rc = ensure(&c->y, (size_t)cf * 4 * 576 * 4))) return rc;
    }
  }
  /* K1 shared-memory window: 512 reservoir bytes + the largest group of K1_FPB frames (+ alignment and read-ahead slack) */
  maxg *= (K1_FPB + 31) / 32;                              /* (maxg is measured over groups of 32 frames; K1_FPB is 32 unless an experiment build says otherwise) */
  c->k1_smem_words = (uint32_t)(((512 + maxg + 64 + 15) & ~(uint64_t)15) / 4);
  if ((size_t)c->k1_smem_words * 4 + sizeof(((p3_tables *)0)->hlut) + 4 * K1_THREADS * 4 > 200 * 1024) return fail(P3_EINVAL, "frame group too large for shared memory");
  /* compact main-data stream: 512 reservoir bytes + all main data of the batch, zero padded (the bit readers run a few words ahead) */
  sl->ms_bytes = 512 + total_ms;
  if ((rc = ensure(&sl->ms, sl->ms_bytes + 512))) return rc;
  CK(cudaMemsetAsync((uint8_t *)sl->ms.p + (sl->ms_bytes & ~(uint64_t)3), 0, 128, st));
  return P3_OK;
}

/* Stage a parsed batch in slot `sl`: allocate, size K1's window, enqueue the uploads on `st`. */
static int stage_batch(p3_ctx *c, p3_slot *sl, const uint8_t *raw, uint64_t raw_bytes, const p3_parsed *b, cudaStream_t st)
{
  int64_t nf = b->n_frames;
  c->n_frames = nf; c->n_pcm_frames = b->n_pcm_frames; c->raw_bytes = raw_bytes;
  c->nch = nf ? b->frames[0].nch : 2;
  if (nf == 0) return P3_OK;
  for (int64_t f = 1; f < nf; f++)
    if (b->frames[f].nch != c->nch) return fail(P3_EINVAL, "channel count changes inside a batch (frame %lld)", (long long)f);
  int rc;
  if (c->tail_on_device) {                                 /* the previous batch was staged from raw bytes: its reservoir carry lives on the device */
    CK(cudaMemcpyAsync(sl->h_tail, c->d_tail_cur, 512, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    memcpy(c->h_tail, sl->h_tail, 512); c->tail_on_device = 0;
  }
  if ((rc = ensure(&sl->raw, raw_bytes + 64))) return rc;
  uint64_t maxg = 0;
  for (int64_t f0 = 0; f0 < nf; f0 += 32) {                /* groups of 32 frames, like k_hop_groups */
    int64_t f1 = f0 + 32 < nf ? f0 + 32 : nf;
    uint64_t span = b->frames[f1 - 1].main_pos + b->frames[f1 - 1].main_size - b->frames[f0].main_pos;
    if (span > maxg) maxg = span;
  }
  if ((rc = size_batch(c, sl, nf, b->n_pcm_frames, maxg, b->frames[nf - 1].main_pos + b->frames[nf - 1].main_size - b->frames[0].main_pos, st))) return rc;
  memcpy(sl->h_tail, c->h_tail, 512);
  CK(cudaMemcpyAsync(sl->raw.p, raw, raw_bytes, cudaMemcpyHostToDevice, st));
  sl->raw_dev = (const uint8_t *)sl->raw.p;
  CK(cudaMemcpyAsync(sl->frames.p, b->frames, (size_t)nf * sizeof(p3_frame), cudaMemcpyHostToDevice, st));
  sl->hop_only = b->hop_only;
  sl->iso = b->n_frames > 0 && (b->frames[0].flags & P3_FRAME_ISO) != 0;
  if (!b->hop_only) CK(cudaMemcpyAsync(sl->gcs.p, b->gcs, (size_t)nf * 4 * sizeof(p3_gc), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(sl->d_tail, sl->h_tail, 512, cudaMemcpyHostToDevice, st));
  compute_tail(raw, b, c->h_tail, c->next_tail); c->have_next_tail = 1;
  return P3_OK;
}

extern "C" int p3_batch_upload(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, const p3_parsed *b)
{
  if (!c || !raw || !b) return fail(P3_EINVAL, "null argument");
  CK(cudaSetDevice(c->device));
  p3_slot *sl = &c->slot[c->cur_slot];
  slot_release(sl);
  return stage_batch(c, sl, raw, raw_bytes, b, c->stream);
}

/* FAST mode synthesis (K2+K3+K4 fused, p3_fused.cu) of frames [f0,f1): stereo batches go to the warp-autonomous packed
 * kernel, mono batches and tapped runs to the one-channel-per-thread kernel. */
static void launch_synth(p3_ctx *c, p3_slot *sl, int64_t f0, int64_t f1, const int16_t *is16, const int32_t *c1, const uint8_t *scf,
                         const p3_state *si, p3_state *so)
{
  const p3_frame *fr = (const p3_frame *)sl->frames.p; const p3_gc *gc = (const p3_gc *)sl->gcs.p;
  const int64_t nf = f1 - f0;
  if (c->nch == 2 && !c->taps && c->synth_kernel != 1) {
    const int classify = c->synth_kernel == 0;
    const int wpb = p3_synthw_warps_per_cta();
    const int64_t warps = (nf + c->fpc - 1) / c->fpc;
    /* the same grid three times: every CTA classifies its frames and only the kernel of that class decodes them (p3_synthw.cuh) */
    const unsigned grid = (unsigned)((warps + wpb - 1) / wpb);
    (sl->iso ? k_synth_warp_iso_lean : k_synth_warp_lean)<<<grid, wpb * 32, p3_synthw_smem_bytes() + (size_t)c->ov_synth_pad, c->stream>>>(fr, gc, c->d_tables, f0, f1, c->fpc,
        is16, c1, scf, si, so, (int16_t *)sl->pcm.p, c->d_pow43s + 8207, classify);
    (sl->iso ? k_synth_warp_iso_same : k_synth_warp_same)<<<grid, wpb * 32, p3_synthw_smem_bytes() + (size_t)c->ov_synth_pad, c->stream>>>(fr, gc, c->d_tables, f0, f1, c->fpc,
        is16, c1, scf, si, so, (int16_t *)sl->pcm.p, c->d_pow43s + 8207, classify);
    (sl->iso ? k_synth_warp_iso : k_synth_warp)<<<grid, wpb * 32, p3_synthw_smem_bytes() + (size_t)c->ov_synth_pad, c->stream>>>(fr, gc, c->d_tables, f0, f1, c->fpc,
        is16, c1, scf, si, so, (int16_t *)sl->pcm.p, c->d_pow43s + 8207, classify);
    c->launches += 2;
  } else
    k_synth_fast<<<(unsigned)((nf + c->fpc - 1) / c->fpc), 128, 0, c->stream>>>(fr, gc, c->d_tables, f0, f1, c->fpc, is16, c1, scf, si, so,
        (int16_t *)sl->pcm.p, c->taps ? (float *)c->xr.p : NULL, c->taps ? (float *)c->y.p : NULL);
}

/* Launch K1..K4 over frames [f0,f1) of the batch staged in `sl`.  ev != NULL: record stage events. */
static int run_chunk(p3_ctx *c, p3_slot *sl, int64_t f0, int64_t f1, cudaEvent_t *ev)
{
  const p3_frame *fr = (const p3_frame *)sl->frames.p; const p3_gc *gc = (const p3_gc *)sl->gcs.p;
  const int64_t nf = f1 - f0;
  p3_state *si = c->d_state[c->cur], *so = c->d_state[c->cur ^ 1];
  CK(cudaMemcpyAsync(so, si, sizeof(p3_state), cudaMemcpyDeviceToDevice, c->stream));   /* fields a launch does not rewrite carry over */
  if (ev) CK(cudaEventRecord(ev[0], c->stream));
  k_compact<<<(unsigned)((nf + 3) / 4), 128, 0, c->stream>>>(sl->raw_dev, fr, sl->d_tail, f0, f1, (uint32_t *)sl->ms.p);
  if (ev) CK(cudaEventRecord(ev[1], c->stream));
  size_t smem1 = (size_t)c->k1_smem_words * 4 + 4 * K1_THREADS * 4 + K1_LUT_SMEM(p3_tables_get()->hlut_used) + 16;
  k_huffman<<<(unsigned)((nf + K1_FPB - 1) / K1_FPB), K1_THREADS, smem1, c->stream>>>((const uint32_t *)sl->ms.p, fr, gc, c->d_tables, f0, f1,
      c->k1_smem_words, (int16_t *)c->is16.p, (int32_t *)c->count1.p, (uint8_t *)c->scf.p);
  if (ev) CK(cudaEventRecord(ev[2], c->stream));
  if (c->mode == P3_MODE_FAST) {
    launch_synth(c, sl, f0, f1, (const int16_t *)c->is16.p, (const int32_t *)c->count1.p, (const uint8_t *)c->scf.p, si, so);
    if (ev) { CK(cudaEventRecord(ev[3], c->stream)); CK(cudaEventRecord(ev[4], c->stream)); CK(cudaEventRecord(ev[5], c->stream)); }
    CK(cudaGetLastError());
    c->cur ^= 1; c->launches += 3;
    return P3_OK;
  }
  k_requant<<<(unsigned)(2 * nf), K2_THREADS, 0, c->stream>>>(fr, gc, c->d_tables, f0, f1, (const int16_t *)c->is16.p,
      (const int32_t *)c->count1.p, (const uint8_t *)c->scf.p, si, so, (float *)c->xr.p);
  if (ev) CK(cudaEventRecord(ev[3], c->stream));
  k_imdct<<<(unsigned)(4 * nf), K3_THREADS, 0, c->stream>>>(fr, gc, c->d_tables, f0, f1, (const float *)c->xr.p, si, so, (float *)c->y.p);
  if (ev) CK(cudaEventRecord(ev[4], c->stream));
  size_t smem4 = (size_t)(2048 + 512 + 2 * (15 + K4_SLOTS) * 96) * 4;
  k_polyphase<<<(unsigned)((2 * nf + K4_GRAN - 1) / K4_GRAN), K4_THREADS, smem4, c->stream>>>(fr, c->d_tables, f0, f1,
      (const float *)c->y.p, si, so, (int16_t *)sl->pcm.p);
  if (ev) CK(cudaEventRecord(ev[5], c->stream));
  CK(cudaGetLastError());
  c->cur ^= 1;
  c->launches += 5;
  return P3_OK;
}

/* Read_Audio_L3 on the device (k_sideinfo + the Q6 chain): once per staged batch, in front of the first decode */
static int run_sideinfo_range(p3_ctx *c, p3_slot *sl, int64_t f0, int64_t f1)
{
  if (f1 <= f0) return P3_OK;
  CK(cudaMemsetAsync(sl->d_any_empty, 0, sizeof(int), c->stream));
  k_sideinfo<<<(unsigned)((f1 - f0 + 127) / 128), 128, 0, c->stream>>>(sl->raw_dev, (p3_frame *)sl->frames.p + f0, (p3_gc *)sl->gcs.p + 4 * f0, f1 - f0, sl->d_any_empty);
  /* (a range that does not start at frame 0: an empty part whose slot was last written before f0 gets w3 = "not in this
   *  batch", which the kernels resolve through the carried state -- what they do at a launch boundary anyway) */
  k_q6_chain<<<1, 1024, 0, c->stream>>>((const p3_frame *)sl->frames.p + f0, (p3_gc *)sl->gcs.p + 4 * f0, f1 - f0, sl->d_any_empty);
  CK(cudaGetLastError());
  c->launches_parse = 2;
  return P3_OK;
}

static int run_sideinfo(p3_ctx *c, p3_slot *sl)
{
  if (!sl->hop_only || c->n_frames == 0) return P3_OK;
  int rc = run_sideinfo_range(c, sl, 0, c->n_frames);
  sl->hop_only = 0;                                        /* the descriptors are complete now; later runs of the same batch reuse them */
  return rc;
}

/* FAST mode, K1 and the synthesis overlapped.  K1 is integer / shared-memory work that leaves the FMA pipe idle (16 % busy), the
 * synthesis is packed-fp32 work bound by that pipe and by issue slots at 57 % each: run one after the other, each leaves half of
 * the SM's issue slots empty.  Here the batch goes through in chunks: K0 + K1 of chunk i+1 on their own stream while the
 * synthesis kernels of chunk i run on c->stream, chained by events, OV_NBUF sets of intermediates; a chunk is small enough that
 * neither kernel fills the machine on its own, so CTAs of both are resident on an SM together.  Same kernels, same launch
 * arithmetic per chunk as run_chunk: the PCM is bit-identical to the sequential path (tests/test_gpu_fast.py). */
static int overlap_on(const p3_ctx *c)
{
  return c->mode == P3_MODE_FAST && c->ov_chunk > 0 && !c->taps && c->n_frames > c->ov_chunk;
}

static int run_all_overlap(p3_ctx *c, p3_slot *sl)
{
  const p3_frame *fr = (const p3_frame *)sl->frames.p; const p3_gc *gc = (const p3_gc *)sl->gcs.p;
  const int64_t C = c->ov_chunk;
  if (c->is16.cap < (size_t)OV_NBUF * C * 4 * 576 * 2) return fail(P3_EINVAL, "batch was staged without room for the overlapped pipeline");
  cudaStream_t sk = c->s_k1[c->ov_prio];
  const size_t smem1 = (size_t)c->k1_smem_words * 4 + 4 * K1_THREADS * 4 + K1_LUT_SMEM(p3_tables_get()->hlut_used) + 16 + (size_t)c->ov_k1_pad;
  CK(cudaEventRecord(c->ov_start, c->stream)); CK(cudaStreamWaitEvent(sk, c->ov_start, 0));   /* K1 starts behind whatever the kernel stream holds (side info, the previous batch) */
  int i = 0;
  for (int64_t f0 = 0; f0 < c->n_frames; f0 += C, i++) {
    const int64_t f1 = f0 + C < c->n_frames ? f0 + C : c->n_frames, nf = f1 - f0;
    const int b = i % OV_NBUF;
    int16_t *is16 = (int16_t *)c->is16.p + (size_t)b * C * 4 * 576; int32_t *c1 = (int32_t *)c->count1.p + (size_t)b * C * 4;
    uint8_t *scf = (uint8_t *)c->scf.p + (size_t)b * C * 4 * P3_SCF_STRIDE;
    if (i >= OV_NBUF) CK(cudaStreamWaitEvent(sk, c->ov_synth_done[b], 0));          /* the synthesis of chunk i - OV_NBUF has read this set */
    k_compact<<<(unsigned)((nf + 3) / 4), 128, 0, sk>>>(sl->raw_dev, fr, sl->d_tail, f0, f1, (uint32_t *)sl->ms.p);
    k_huffman<<<(unsigned)((nf + K1_FPB - 1) / K1_FPB), K1_THREADS, smem1, sk>>>((const uint32_t *)sl->ms.p, fr, gc, c->d_tables, f0, f1, c->k1_smem_words, is16, c1, scf);
    CK(cudaEventRecord(c->ov_k1_done[b], sk));
    CK(cudaStreamWaitEvent(c->stream, c->ov_k1_done[b], 0));
    p3_state *si = c->d_state[c->cur], *so = c->d_state[c->cur ^ 1];
    CK(cudaMemcpyAsync(so, si, sizeof(p3_state), cudaMemcpyDeviceToDevice, c->stream));
    launch_synth(c, sl, f0, f1, is16, c1, scf, si, so);
    CK(cudaEventRecord(c->ov_synth_done[b], c->stream));
    CK(cudaGetLastError());
    c->cur ^= 1; c->launches += 3;
  }
  return P3_OK;
}

static int run_all(p3_ctx *c, p3_slot *sl)
{
  c->launches = 0;
  { int rc = run_sideinfo(c, sl); if (rc) return rc; }
  if (overlap_on(c)) return run_all_overlap(c, sl);
  for (int64_t f0 = 0; f0 < c->n_frames; f0 += c->chunk_frames) {
    int64_t f1 = f0 + c->chunk_frames < c->n_frames ? f0 + c->chunk_frames : c->n_frames;
    int rc = run_chunk(c, sl, f0, f1, NULL);
    if (rc) return rc;
  }
  return P3_OK;
}

extern "C" int p3_batch_run(p3_ctx *c)
{
  if (!c) return fail(P3_EINVAL, "null ctx");
  CK(cudaSetDevice(c->device));
  return run_all(c, &c->slot[c->cur_slot]);
}

extern "C" int p3_batch_sync(p3_ctx *c)
{
  if (!c) return fail(P3_EINVAL, "null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->s_h2d)); CK(cudaStreamSynchronize(c->stream)); CK(cudaStreamSynchronize(c->s_d2h));
  for (int i = 0; i < 2; i++) slot_release(&c->slot[i]);
  if (g_trace > 0 && g_tr_stage > 0) { fprintf(stderr, "p3 async: slot wait %.1f ms, stage %.1f ms, launch %.1f ms\n", g_tr_wait, g_tr_stage, g_tr_launch); g_tr_wait = g_tr_stage = g_tr_launch = 0; }
  return P3_OK;
}

extern "C" void *p3_batch_pcm_device(p3_ctx *c, uint64_t *bytes)
{
  if (bytes) *bytes = (uint64_t)c->n_pcm_frames * 1152 * c->nch * sizeof(int16_t);
  return c->slot[c->cur_slot].pcm.p;
}

extern "C" int p3_batch_download(p3_ctx *c, int16_t *pcm, const p3_taps *t)
{
  if (!c) return fail(P3_EINVAL, "null ctx");
  CK(cudaSetDevice(c->device));
  p3_slot *sl = &c->slot[c->cur_slot];
  if (pcm && c->n_pcm_frames)
    CK(cudaMemcpyAsync(pcm, sl->pcm.p, (size_t)c->n_pcm_frames * 1152 * c->nch * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream));
  if (t) {
    if (c->n_frames > c->chunk_frames) return fail(P3_EINVAL, "taps need the batch to fit one chunk (%lld frames)", (long long)c->chunk_frames);
    size_t ngc = (size_t)c->n_frames * 4;
    if (t->is_huff) CK(cudaMemcpyAsync(t->is_huff, c->is16.p, ngc * 576 * 2, cudaMemcpyDeviceToHost, c->stream));
    if (t->count1)  CK(cudaMemcpyAsync(t->count1, c->count1.p, ngc * 4, cudaMemcpyDeviceToHost, c->stream));
    if (t->scf)     CK(cudaMemcpyAsync(t->scf, c->scf.p, ngc * P3_SCF_STRIDE, cudaMemcpyDeviceToHost, c->stream));
    if (t->xr)      CK(cudaMemcpyAsync(t->xr, c->xr.p, ngc * 576 * 4, cudaMemcpyDeviceToHost, c->stream));
    if (t->y)       CK(cudaMemcpyAsync(t->y, c->y.p, ngc * 576 * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  return P3_OK;
}

extern "C" int p3_batch_download_desc(p3_ctx *c, p3_frame *frames, p3_gc *gcs)
{
  if (!c) return fail(P3_EINVAL, "null ctx");
  CK(cudaSetDevice(c->device));
  p3_slot *sl = &c->slot[c->cur_slot];
  CK(cudaStreamSynchronize(c->stream));
  if (frames && c->n_frames) CK(cudaMemcpy(frames, sl->frames.p, (size_t)c->n_frames * sizeof(p3_frame), cudaMemcpyDeviceToHost));
  if (gcs && c->n_frames) CK(cudaMemcpy(gcs, sl->gcs.p, (size_t)c->n_frames * 4 * sizeof(p3_gc), cudaMemcpyDeviceToHost));
  return P3_OK;
}

extern "C" int p3_decode_batch(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, const p3_parsed *b, int16_t *pcm, const p3_taps *t)
{
  int rc;
  if (!c) return fail(P3_EINVAL, "null ctx");
  c->taps = t != NULL;
  if ((rc = p3_batch_sync(c))) return rc;                 /* drain anything still in flight from the async path */
  if ((rc = p3_batch_upload(c, raw, raw_bytes, b))) return rc;
  if ((rc = p3_batch_run(c))) return rc;
  if ((rc = p3_batch_download(c, pcm, t))) return rc;
  commit_tail(c);                                          /* reservoir for the next batch */
  return P3_OK;
}

/* K3 + K4 only, from host spectra (BASELINE configs[1]: Huffman and requantization still on the host) */
extern "C" int p3_synth_from_xr(p3_ctx *c, const float *xr, const p3_parsed *b, int16_t *pcm)
{
  if (!c || !xr || !b || !pcm) return fail(P3_EINVAL, "null argument");
  if (b->hop_only) return fail(P3_EINVAL, "p3_synth_from_xr needs host-parsed descriptors");
  int rc;
  if ((rc = p3_batch_sync(c))) return rc;
  CK(cudaSetDevice(c->device));
  const int64_t nf = b->n_frames;
  if (nf == 0) return P3_OK;
  p3_slot *sl = &c->slot[c->cur_slot];
  c->n_frames = nf; c->n_pcm_frames = b->n_pcm_frames; c->nch = b->frames[0].nch;
  if ((rc = ensure(&sl->frames, (size_t)nf * sizeof(p3_frame)))) return rc;
  if ((rc = ensure(&sl->gcs, (size_t)nf * 4 * sizeof(p3_gc)))) return rc;
  if ((rc = ensure(&sl->pcm, (size_t)(b->n_pcm_frames ? b->n_pcm_frames : 1) * 1152 * c->nch * sizeof(int16_t)))) return rc;
  const int64_t cf = nf < (1 << 17) ? nf : (1 << 17);               /* frames per launch pair: 2 x 1.2 GB of fp32 intermediates */
  if ((rc = ensure(&c->xr, (size_t)cf * 4 * 576 * 4))) return rc;
  if ((rc = ensure(&c->y, (size_t)cf * 4 * 576 * 4))) return rc;
  CK(cudaMemcpyAsync(sl->frames.p, b->frames, (size_t)nf * sizeof(p3_frame), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(sl->gcs.p, b->gcs, (size_t)nf * 4 * sizeof(p3_gc), cudaMemcpyHostToDevice, c->stream));
  const p3_frame *fr = (const p3_frame *)sl->frames.p; const p3_gc *gc = (const p3_gc *)sl->gcs.p;
  c->launches = 0;
  for (int64_t f0 = 0; f0 < nf; f0 += cf) {
    const int64_t f1 = f0 + cf < nf ? f0 + cf : nf, n = f1 - f0;
    p3_state *si = c->d_state[c->cur], *so = c->d_state[c->cur ^ 1];
    CK(cudaMemcpyAsync(so, si, sizeof(p3_state), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->xr.p, xr + (size_t)f0 * 4 * 576, (size_t)n * 4 * 576 * 4, cudaMemcpyHostToDevice, c->stream));
    k_imdct<<<(unsigned)(4 * n), K3_THREADS, 0, c->stream>>>(fr, gc, c->d_tables, f0, f1, (const float *)c->xr.p, si, so, (float *)c->y.p);
    size_t smem4 = (size_t)(2048 + 512 + 2 * (15 + K4_SLOTS) * 96) * 4;
    k_polyphase<<<(unsigned)((2 * n + K4_GRAN - 1) / K4_GRAN), K4_THREADS, smem4, c->stream>>>(fr, c->d_tables, f0, f1, (const float *)c->y.p, si, so, (int16_t *)sl->pcm.p);
    CK(cudaGetLastError());
    c->cur ^= 1; c->launches += 2;
  }
  if (b->n_pcm_frames)
    CK(cudaMemcpyAsync(pcm, sl->pcm.p, (size_t)b->n_pcm_frames * 1152 * c->nch * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return P3_OK;
}

/* Asynchronous batch: uploads on the H2D stream, kernels on the kernel stream, the PCM download on
 * the D2H stream, chained by events; returns as soon as everything is enqueued.  Alternating between
 * two slots overlaps the transfer of one batch with the kernels of the other.  Takes ownership of
 * `b` (freed once its upload has completed).  `raw` and `pcm` must stay valid until p3_batch_sync(). */
extern "C" int p3_decode_batch_async(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, p3_parsed *b, int16_t *pcm)
{
  if (!c || !raw || !b) return fail(P3_EINVAL, "null argument");
  CK(cudaSetDevice(c->device));
  c->taps = 0;
  c->cur_slot ^= 1;
  p3_slot *sl = &c->slot[c->cur_slot];
  if (g_trace < 0) g_trace = getenv("P3_TRACE") != NULL;
  double t0 = g_trace ? now_ms() : 0;
  slot_release(sl);                                        /* wait for the batch that used this slot two calls ago */
  double t1 = g_trace ? now_ms() : 0;
  int rc = stage_batch(c, sl, raw, raw_bytes, b, c->s_h2d);
  if (rc) return rc;
  double t2 = g_trace ? now_ms() : 0;
  if (!b->external) { sl->keep = *b; sl->have_keep = 1; b->frames = NULL; b->gcs = NULL; }   /* caller-owned arrays: the caller rotates them */
  CK(cudaEventRecord(sl->h2d_done, c->s_h2d));
  CK(cudaStreamWaitEvent(c->stream, sl->h2d_done, 0));
  if ((rc = run_all(c, sl))) return rc;
  CK(cudaEventRecord(sl->compute_done, c->stream));
  CK(cudaStreamWaitEvent(c->s_d2h, sl->compute_done, 0));
  if (pcm && c->n_pcm_frames)
    CK(cudaMemcpyAsync(pcm, sl->pcm.p, (size_t)c->n_pcm_frames * 1152 * c->nch * sizeof(int16_t), cudaMemcpyDeviceToHost, c->s_d2h));
  CK(cudaEventRecord(sl->d2h_done, c->s_d2h));
  /* the other slot's kernels must not start before this download has its data: ordering on c->stream is implicit;
   * its next stage_batch() waits for d2h_done through slot_release() */
  sl->busy = 1;
  commit_tail(c);
  if (g_trace) { g_tr_wait += t1 - t0; g_tr_stage += t2 - t1; g_tr_launch += now_ms() - t2; }
  return P3_OK;
}

/* CUDA-event timing of the kernel sequence with everything resident in HBM.  The carried state is
 * restored before every iteration so that each one decodes the same thing.  ms_stage: K0 (compact), K1..K4. */
extern "C" int p3_batch_time(p3_ctx *c, int iters, float *ms_total, float *ms_stage)
{
  if (!c || iters <= 0) return fail(P3_EINVAL, "bad argument");
  CK(cudaSetDevice(c->device));
  { int rc = run_sideinfo(c, &c->slot[c->cur_slot]); if (rc) return rc; }
  p3_state *save; CK(cudaMalloc(&save, sizeof(p3_state)));
  CK(cudaMemcpyAsync(save, c->d_state[c->cur], sizeof(p3_state), cudaMemcpyDeviceToDevice, c->stream));
  float tot = 0, st[5] = {0, 0, 0, 0, 0};
  for (int it = 0; it < iters; it++) {
    CK(cudaMemcpyAsync(c->d_state[c->cur], save, sizeof(p3_state), cudaMemcpyDeviceToDevice, c->stream));
    c->launches = 0;
    CK(cudaEventRecord(c->ev[8], c->stream));
    const bool one = c->n_frames <= c->chunk_frames && !overlap_on(c);       /* one launch sequence: per-kernel times exist */
    if (one) { int rc = run_chunk(c, &c->slot[c->cur_slot], 0, c->n_frames, c->ev); if (rc) { cudaFree(save); return rc; } }
    else { int rc = run_all(c, &c->slot[c->cur_slot]); if (rc) { cudaFree(save); return rc; } }
    CK(cudaEventRecord(c->ev[9], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms; CK(cudaEventElapsedTime(&ms, c->ev[8], c->ev[9])); tot += ms;
    if (one)
      for (int k = 0; k < 5; k++) { CK(cudaEventElapsedTime(&ms, c->ev[k], c->ev[k + 1])); st[k] += ms; }
  }
  cudaFree(save);
  if (ms_total) *ms_total = tot / iters;
  if (ms_stage) for (int k = 0; k < 5; k++) ms_stage[k] = st[k] / iters;
  return P3_OK;
}

/* BASELINE configs[1] with everything resident in HBM: the batch must have been run once in P3_MODE_EXACT as ONE launch
 * sequence (n_frames <= chunk), so that c->xr holds the spectra after requantize / reorder / stereo / antialias; times
 * k_imdct + k_polyphase reading them, `iters` times, from the same carried state.  ms_stage: [0] k_imdct, [1] k_polyphase. */
extern "C" int p3_batch_time_xr(p3_ctx *c, int iters, float *ms_total, float *ms_stage)
{
  if (!c || iters <= 0) return fail(P3_EINVAL, "bad argument");
  if (c->mode != P3_MODE_EXACT || c->n_frames == 0 || c->n_frames > c->chunk_frames || !c->xr.p || !c->y.p)
    return fail(P3_EINVAL, "p3_batch_time_xr needs a batch run in P3_MODE_EXACT as one launch sequence");
  CK(cudaSetDevice(c->device));
  p3_slot *sl = &c->slot[c->cur_slot];
  const p3_frame *fr = (const p3_frame *)sl->frames.p; const p3_gc *gc = (const p3_gc *)sl->gcs.p;
  const int64_t nf = c->n_frames;
  p3_state *save; CK(cudaMalloc(&save, sizeof(p3_state)));
  CK(cudaMemsetAsync(save, 0, sizeof(p3_state), c->stream));
  float tot = 0, st[2] = {0, 0};
  const size_t smem4 = (size_t)(2048 + 512 + 2 * (15 + K4_SLOTS) * 96) * 4;
  for (int it = 0; it < iters; it++) {
    p3_state *si = c->d_state[c->cur], *so = c->d_state[c->cur ^ 1];
    CK(cudaMemcpyAsync(si, save, sizeof(p3_state), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(so, save, sizeof(p3_state), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaEventRecord(c->ev[0], c->stream));
    k_imdct<<<(unsigned)(4 * nf), K3_THREADS, 0, c->stream>>>(fr, gc, c->d_tables, 0, nf, (const float *)c->xr.p, si, so, (float *)c->y.p);
    CK(cudaEventRecord(c->ev[1], c->stream));
    k_polyphase<<<(unsigned)((2 * nf + K4_GRAN - 1) / K4_GRAN), K4_THREADS, smem4, c->stream>>>(fr, c->d_tables, 0, nf, (const float *)c->y.p, si, so, (int16_t *)sl->pcm.p);
    CK(cudaEventRecord(c->ev[2], c->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    float ms;
    CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); st[0] += ms;
    CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); st[1] += ms;
    CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[2])); tot += ms;
  }
  cudaFree(save);
  c->launches = 2;
  if (ms_total) *ms_total = tot / iters;
  if (ms_stage) { ms_stage[0] = st[0] / iters; ms_stage[1] = st[1] / iters; }
  return P3_OK;
}

/* ---- batches staged from RAW BYTES: the frame hop runs on the device (p3_hop.cu) ---------------------------------- */

/* Stage the frames found in raw[0, raw_bytes) (host memory, or device memory when raw_on_device) in the current slot: copy
 * the bytes, hop on the device, size the buffers.  *st (may be NULL) is the parser state in and, when the batch is
 * committed by p3_decode_raw(), out; *info receives n_frames / n_pcm_frames / consumed / stop as p3_parse() would report
 * them (its frames / gcs pointers stay NULL: the descriptors exist on the device only). */
/* the hop of a staged byte stream (sl->raw_dev, raw_bytes): frames -> sl->frames, the result block -> c->hop.h_res */
static int hop_staged(p3_ctx *c, p3_slot *sl, uint64_t raw_bytes, const p3_parse_opts *o, const p3_parse_state *st_in, cudaStream_t st)
{
  int rc;
  if (c->tail_on_device) CK(cudaMemcpyAsync(sl->d_tail, c->d_tail_cur, 512, cudaMemcpyDeviceToDevice, st));
  else { memcpy(sl->h_tail, c->h_tail, 512); CK(cudaMemcpyAsync(sl->d_tail, sl->h_tail, 512, cudaMemcpyHostToDevice, st)); }
  CK(cudaEventRecord(c->ev[6], st));
  if ((rc = p3_hop_count(&c->hop, st, sl->raw_dev, raw_bytes, o, st_in, 0))) return fail(rc, "device frame hop failed");
  const p3_hop_result *r = c->hop.h_res;
  c->n_frames = r->n_frames; c->n_pcm_frames = r->n_pcm_frames; c->raw_bytes = raw_bytes; c->nch = (uint32_t)r->nch;
  if ((rc = ensure(&sl->frames, (size_t)(r->n_frames ? r->n_frames : 1) * sizeof(p3_frame)))) return rc;
  if ((rc = p3_hop_emit(&c->hop, st, sl->raw_dev, raw_bytes, o, st_in, (p3_frame *)sl->frames.p, sl->d_tail, c->d_tail_nxt))) return fail(rc, "device frame hop failed");
  CK(cudaEventRecord(c->ev[7], st)); CK(cudaEventSynchronize(c->ev[7]));
  CK(cudaEventElapsedTime(&c->hop_ms, c->ev[6], c->ev[7]));            /* device time of the hop: kernels + the two result read-backs */
  c->have_next_tail_dev = 1; c->have_next_tail = 0;
  sl->hop_only = 1; sl->iso = o->iso != 0;
  return P3_OK;
}

static int stage_raw(p3_ctx *c, p3_slot *sl, const uint8_t *raw, uint64_t raw_bytes, int raw_on_device, const p3_parse_opts *o, const p3_parse_state *st_in, p3_parsed *info, cudaStream_t st)
{
  p3_parse_opts od; memset(&od, 0, sizeof od);
  p3_parse_state sd = {0, 0, 0, -1, -1};
  if (!o) o = &od;
  if (!st_in) st_in = &sd;
  int rc;
  if ((rc = ensure(&sl->raw, raw_bytes + 64))) return rc;
  sl->raw_dev = (const uint8_t *)sl->raw.p;
  if (raw_bytes) CK(cudaMemcpyAsync(sl->raw.p, raw, raw_bytes, raw_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync((uint8_t *)sl->raw.p + raw_bytes, 0, 64, st));
  if ((rc = hop_staged(c, sl, raw_bytes, o, st_in, st))) return rc;
  const p3_hop_result *r = c->hop.h_res;
  if ((rc = size_batch(c, sl, r->n_frames, r->n_pcm_frames, r->maxg, r->total_ms, st))) return rc;
  if (info) {
    memset(info, 0, sizeof *info);
    info->n_frames = r->n_frames; info->n_pcm_frames = r->n_pcm_frames; info->consumed = r->consumed; info->stop = r->stop; info->hop_only = 1; info->external = 1;
  }
  return P3_OK;
}

extern "C" int p3_batch_upload_raw(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, int raw_on_device, const p3_parse_opts *o, const p3_parse_state *st, p3_parsed *info)
{
  if (!c || (!raw && raw_bytes)) return fail(P3_EINVAL, "null argument");
  CK(cudaSetDevice(c->device));
  p3_slot *sl = &c->slot[c->cur_slot];
  slot_release(sl);
  return stage_raw(c, sl, raw, raw_bytes, raw_on_device, o, st, info, c->stream);
}

extern "C" int p3_batch_channels(p3_ctx *c) { return c ? (int)c->nch : 0; }
extern "C" float p3_hop_ms(p3_ctx *c) { return c ? c->hop_ms : 0.0f; }
extern "C" int p3_hop_rounds(p3_ctx *c) { return c && c->hop.h_res ? c->hop.h_res->changed : 0; }

/* Asynchronous, double-buffered decode from raw bytes (behind pdmp3_read() for large reads): upload of the byte window and the
 * device hop on the H2D stream -- the call returns once the hop result is known (frames found, bytes consumed), which is
 * while the PREVIOUS batch is still in its kernels / download --, then kernels and the PCM download of exactly the frames
 * found are enqueued.  raw and pcm must stay valid until p3_batch_sync().  *st is advanced. */
extern "C" int p3_decode_raw_async(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, const p3_parse_opts *o, p3_parse_state *st, p3_parsed *info, int16_t *pcm)
{
  if (!c || !raw || !info) return fail(P3_EINVAL, "null argument");
  CK(cudaSetDevice(c->device));
  c->taps = 0;
  c->cur_slot ^= 1;
  p3_slot *sl = &c->slot[c->cur_slot];
  if (g_trace < 0) g_trace = getenv("P3_TRACE") != NULL;
  double t0 = g_trace ? now_ms() : 0;
  slot_release(sl);                                        /* wait for the batch that used this slot two calls ago */
  double t1 = g_trace ? now_ms() : 0;
  int rc = stage_raw(c, sl, raw, raw_bytes, 0, o, st, info, c->s_h2d);
  if (rc) return rc;
  double t2 = g_trace ? now_ms() : 0;
  if (c->n_frames > 0) {
    CK(cudaEventRecord(sl->h2d_done, c->s_h2d));
    CK(cudaStreamWaitEvent(c->stream, sl->h2d_done, 0));
    if ((rc = run_all(c, sl))) return rc;
    CK(cudaEventRecord(sl->compute_done, c->stream));
    CK(cudaStreamWaitEvent(c->s_d2h, sl->compute_done, 0));
    if (pcm && c->n_pcm_frames)
      CK(cudaMemcpyAsync(pcm, sl->pcm.p, (size_t)c->n_pcm_frames * 1152 * c->nch * sizeof(int16_t), cudaMemcpyDeviceToHost, c->s_d2h));
    CK(cudaEventRecord(sl->d2h_done, c->s_d2h));
    sl->busy = 1;
    commit_tail(c);
    if (st) *st = c->hop.h_res->st;
  } else { c->have_next_tail_dev = 0; }
  if (g_trace) { g_tr_wait += t1 - t0; g_tr_stage += t2 - t1; g_tr_launch += now_ms() - t2; }
  return P3_OK;
}

/* Decode raw[0, raw_bytes) in one call, hop included: the device-side counterpart of p3_parse() + p3_decode_batch().
 * *st is updated to the parser state after the decoded frames; info->consumed tells where the next call continues. */
extern "C" int p3_decode_raw(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, const p3_parse_opts *o, p3_parse_state *st, p3_parsed *info,
                             int16_t *pcm, int64_t pcm_cap_frames, const p3_taps *t)
{
  int rc;
  if (!c) return fail(P3_EINVAL, "null ctx");
  c->taps = t != NULL;
  if ((rc = p3_batch_sync(c))) return rc;
  p3_parse_opts oo; if (o) oo = *o; else memset(&oo, 0, sizeof oo);
  if (pcm_cap_frames > 0 && (oo.max_frames <= 0 || oo.max_frames > pcm_cap_frames + oo.warmup_frames)) oo.max_frames = pcm_cap_frames + oo.warmup_frames;
  if ((rc = p3_batch_upload_raw(c, raw, raw_bytes, 0, &oo, st, info))) return rc;
  if ((rc = p3_batch_run(c))) return rc;
  if ((rc = p3_batch_download(c, pcm, t))) return rc;
  commit_tail(c);
  if (st && c->hop.h_res->n_frames > 0) *st = c->hop.h_res->st;
  return P3_OK;
}

#include "p3_dist.cuh"
