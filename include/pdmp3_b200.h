/* pdmp3_b200.h -- inner C-ABI: host C (parser, batcher, streaming API) -> sm_100a CUDA kernels.
 *
 * Plain C types only.  Every entry point returns 0 on success or a negative P3_E* code; the
 * library has NO CPU fallback: if no CUDA device / kernel image is available the calls fail.
 *
 * The batch path replaces, for a whole batch of frames at once, what the reference does per
 * frame in  Read_Main_L3 -> Read_Huffman (pdmp3.c:1346-1442, 2051-2115)  and
 * Decode_L3 (pdmp3.c:1024-1060: L3_Requantize 1829, L3_Reorder 1786, L3_Stereo 1911,
 * L3_Antialias 1706, L3_Hybrid_Synthesis 1752, L3_Frequency_Inversion 1738,
 * L3_Subband_Synthesis 1978) plus Convert_Frame_S16 (2307).
 */
#ifndef PDMP3_B200_H
#define PDMP3_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define P3_OK         0
#define P3_EINVAL    -1
#define P3_ENOMEM    -2
#define P3_ECUDA     -3      /* CUDA runtime / launch failure (p3_last_error() has the text) */
#define P3_ENODEV    -4      /* no usable sm_100 device: there is no CPU fallback */

/* ---- host-side descriptors produced by the parser (replaces Read_Header 1252-1320,
 *      Read_Audio_L3 1129-1200 and the bookkeeping of Get_Main_Data 1096-1122) ------------- */

#define P3_FRAME_DECODE 1u   /* decode and emit PCM for this frame                             */
#define P3_FRAME_NODATA 2u   /* main_data_begin exceeds the reservoir (pdmp3.c:1101): silence   */
#define P3_FRAME_BAD    4u   /* side info fails validation: silence               */
#define P3_FRAME_WARMUP 8u   /* decoded for filter state only (shard halo), no PCM emitted     */
#define P3_FRAME_ISO    16u  /* ISO 11172-3 semantics where the reference deviates: count1 table B
                                decoded as the 4-bit code it is, MS stereo up to max(count1), intensity positions from the
                                RIGHT channel's scalefactors with the ratio multiply in short blocks too, an empty part
                                has count1 = 0.  Off (default) = bit-compatible with pdmp3.c.                          */

typedef struct {             /* 32 bytes, one per MP3 frame */
  uint64_t main_off;         /* byte offset of this frame's main data inside the raw stream     */
  uint64_t main_pos;         /* position of that data in the header-stripped main-data stream   */
  uint16_t main_size;        /* bytes of main data carried by this frame (pdmp3.c:1146-1148)    */
  uint16_t main_begin;       /* main_data_begin, bytes back into the reservoir (pdmp3.c:1157)   */
  uint8_t  nch;              /* 1 or 2                                                          */
  uint8_t  mode;             /* header mode (pdmp3.c:1287)                                      */
  uint8_t  mode_ext;         /* header mode_extension (pdmp3.c:1288)                            */
  uint8_t  sfreq;            /* 0: 44.1 kHz, 1: 48 kHz, 2: 32 kHz                                */
  uint8_t  scfsi;            /* ch0 bands in bits 0-3, ch1 bands in bits 4-7 (pdmp3.c:1163)      */
  uint8_t  flags;            /* P3_FRAME_*                                                       */
  uint16_t bitrate_kbps;
  uint32_t pcm_index;        /* output slot (units of 1152 sample-frames); unused for warm-up    */
} p3_frame;

typedef struct {             /* 16 bytes, one per granule-channel, order [frame][gr][ch(2)]      */
  uint32_t w0;               /* part2_3_length:12 big_values:9 global_gain:8 preflag:1 scalefac_scale:1 count1table_select:1 */
  uint32_t w1;               /* scalefac_compress:4 win_switch:1 block_type:2 mixed:1 table_select0:5 1:5 2:5 region0:4 region1:4 (implicit values reach 13) */
  uint32_t w2;               /* subblock_gain0:3 1:3 2:3 | part2 start bit rel. to (main_pos-main_begin)*8 :14 */
  uint32_t w3;               /* frames back to the last non-empty part of this [gr][ch] slot (0 = this one; Q6) */
} p3_gc;

#define P3_GC_P23L(g)   ((g).w0 & 0xfffu)
#define P3_GC_BIGV(g)   (((g).w0 >> 12) & 0x1ffu)
#define P3_GC_GAIN(g)   (((g).w0 >> 21) & 0xffu)
#define P3_GC_PREF(g)   (((g).w0 >> 29) & 1u)
#define P3_GC_SCALE(g)  (((g).w0 >> 30) & 1u)
#define P3_GC_C1TAB(g)  (((g).w0 >> 31) & 1u)
#define P3_GC_SFCOMP(g) ((g).w1 & 0xfu)
#define P3_GC_WINSW(g)  (((g).w1 >> 4) & 1u)
#define P3_GC_BTYPE(g)  (((g).w1 >> 5) & 3u)
#define P3_GC_MIXED(g)  (((g).w1 >> 7) & 1u)
#define P3_GC_TSEL(g,r) (((g).w1 >> (8 + 5 * (r))) & 0x1fu)
#define P3_GC_REG0(g)   (((g).w1 >> 23) & 0xfu)
#define P3_GC_REG1(g)   (((g).w1 >> 27) & 0xfu)
#define P3_GC_SBG(g,w)  (((g).w2 >> (3 * (w))) & 7u)
#define P3_GC_START(g)  (((g).w2 >> 9) & 0x3fffu)

typedef struct {             /* parser state carried from one batch to the next (streaming)     */
  uint64_t main_pos;         /* logical main-data bytes seen so far                              */
  uint32_t top;              /* g_main_data_top of the reference (pdmp3.c:1109,1120)             */
  uint32_t pcm_index;        /* next output slot                                                 */
  int32_t  nch, sfreq;       /* format of the last parsed frame (-1 = none yet)                  */
} p3_parse_state;

typedef struct {
  int64_t  max_frames;       /* <=0: no limit                                                    */
  uint32_t lookahead;        /* a frame is read only if >= this many bytes are buffered at its
                                sync position (reference: 1152, pdmp3.c:2445); 0 = every complete frame */
  int32_t  nthreads;         /* side-info parse threads (<=0: pick)                              */
  uint32_t warmup_frames;    /* first N frames flagged WARMUP (no PCM slot)                      */
  uint32_t hop_only;         /* 1: only the sequential header hop runs on the host; the side info (Read_Audio_L3,
                                pdmp3.c:1129-1200) is parsed ON THE DEVICE by k_sideinfo when the batch is staged */
  uint32_t iso;              /* 1: flag every frame P3_FRAME_ISO (ISO-correct decoding instead of the reference's quirks) */
} p3_parse_opts;

typedef struct {
  int64_t  n_frames;
  p3_frame *frames;          /* malloc'd [n_frames]   */
  p3_gc    *gcs;             /* malloc'd [n_frames*4] */
  uint64_t consumed;         /* bytes of `data` used up to the end of the last parsed frame      */
  int64_t  n_pcm_frames;     /* frames holding a PCM slot                                        */
  int32_t  external;         /* arrays belong to the caller (p3_parse_into) */
  int32_t  stop;             /* 0: ran out of data, 1: max_frames, 2: no sync within 1152 bytes (pdmp3.c:1337), 3: channel count / sample rate changes at `consumed` */
  int32_t  hop_only;         /* gcs[] (and frames[].scfsi / the BAD flag) are NOT filled: the device parses the side info */
  int32_t  pad_;
} p3_parsed;

int  p3_parse(const uint8_t *data, uint64_t n, const p3_parse_opts *opts, p3_parse_state *state, p3_parsed *out);
int  p3_parse_into(const uint8_t *data, uint64_t n, const p3_parse_opts *opts, p3_parse_state *state, p3_parsed *out,
                   p3_frame *frames_buf, p3_gc *gcs_buf, int64_t buf_cap);   /* caller-owned (e.g. page-locked) descriptor arrays */
void p3_parsed_free(p3_parsed *p);
int  p3_find_header(const uint8_t *data, uint64_t n, int *nch, int *sfreq);   /* 1 found, 0 need more, -1 none within 1152 B */

/* page-locked host memory for stream / PCM buffers (NULL if no device is usable) */
void *p3_host_alloc(size_t bytes);
void *p3_host_alloc_dev(int device, size_t bytes);   /* same, after selecting `device` (no stray context on device 0) */
void  p3_host_free(void *p);

/* ---- device context ------------------------------------------------------------------------ */
typedef struct p3_ctx p3_ctx;

#define P3_MODE_EXACT 0      /* direct-form transforms, reference summation order (bit-exact PCM) */
#define P3_MODE_FAST  1      /* fast transforms (<= 1 LSB of int16 from the reference)            */

int  p3_ctx_create(int device, p3_ctx **out);
void p3_ctx_destroy(p3_ctx *c);
int  p3_ctx_reset(p3_ctx *c);                 /* zero overlap / FIFO / reservoir state (pdmp3_open_feed, pdmp3.c:2377-2379) */
int  p3_ctx_set_mode(p3_ctx *c, int mode);
int  p3_ctx_set_taps(p3_ctx *c, int on);              /* keep stage taps of the next batches on the device */
int  p3_ctx_set_overlap(p3_ctx *c, int64_t chunk_frames, int prio, int synth_pad_bytes, int k1_pad_bytes);
                                                  /* FAST mode: K0 + K1 of chunk i+1 run on their own stream under the synthesis of chunk i (0 = off);
                                                     prio 1: K1's stream above the kernel stream; pads: unused dynamic shared memory that caps the
                                                     CTAs per SM of the synthesis kernels / of K1 (tuning).  Same PCM bits as the sequential path. */
int  p3_ctx_set_frames_per_cta(p3_ctx *c, int n);     /* FAST mode: frames each CTA (k_synth_fast) / warp (k_synth_warp) walks (default 32) */
int  p3_ctx_set_synth_kernel(p3_ctx *c, int which);   /* FAST mode: 0 = k_synth_warp(_lean) for stereo batches (default), 1 = always k_synth_fast,
                                                         2 = k_synth_warp only, no content classes (check: same bits as 0) */
const char *p3_last_error(void);

/* Tap buffers (device side, optional; for stage-level parity tests). NULL = not captured. */
typedef struct {
  int16_t *is_huff;   /* [n_frames][2][2][576] after Huffman                                    */
  int32_t *count1;    /* [n_frames][2][2]                                                       */
  uint8_t *scf;       /* [n_frames][2][2][64]: scalefac_l[21] | pad3 | scalefac_s[12][3] | pad4  */
  float   *xr;        /* [n_frames][2][2][576] after requantize+reorder+stereo+antialias        */
  float   *y;         /* [n_frames][2][2][576] after hybrid synthesis + frequency inversion, [sb][18] */
} p3_taps;

/* Decode a parsed batch.  `raw` is the raw MP3 byte stream the descriptors index into
 * (host memory; copied to the device inside the call), `pcm` receives
 * n_pcm_frames*1152*nch int16 (host memory).  State (IMDCT overlap, polyphase FIFO,
 * reservoir tail) is carried inside the context from call to call.
 * host_taps: optional host tap buffers (NULL entries skipped). */
int p3_decode_batch(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, const p3_parsed *batch,
                    int16_t *pcm, const p3_taps *host_taps);

/* BASELINE configs[1]: the two transform stages only.  `xr` = host spectra after requantize / reorder / stereo /
 * antialias (what Decode_L3 hands to L3_Hybrid_Synthesis, pdmp3.c:1042), [n_frames][2][2][576] fp32; the device runs
 * L3_Hybrid_Synthesis + L3_Frequency_Inversion (1752-1780, 1738-1746) and L3_Subband_Synthesis (1978-2045) with the
 * reference's summation order (bit-exact PCM).  Only block types / flags of the descriptors are used. */
int p3_synth_from_xr(p3_ctx *c, const float *xr, const p3_parsed *batch, int16_t *pcm);

/* Asynchronous, double-buffered variant behind pdmp3_read(): enqueue upload, kernels and PCM download of
 * one batch and return; takes ownership of *batch.  raw/pcm must stay valid until p3_batch_sync(). */
int p3_decode_batch_async(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, p3_parsed *batch, int16_t *pcm);

/* Device-resident variant used by the benchmark and the multi-GPU driver: upload once, run the
 * kernels any number of times, download once.  Pointers returned are DEVICE pointers. */
int p3_batch_upload(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, const p3_parsed *batch);
/* Same, but from RAW BYTES: the frame hop (Search_Header / Read_Header, pdmp3.c:1322-1340, 1252-1320; frame length 1353-1368;
 * reservoir rule of Get_Main_Data 1101-1120) runs ON THE DEVICE (p3_hop.cu) and produces the p3_frame[] array that p3_parse()
 * computes on the host, byte for byte; the side info follows on the device as with hop_only.  raw: host memory, or device
 * memory when raw_on_device.  opts (lookahead, max_frames, warmup_frames, iso) and *state as for p3_parse(); *info receives
 * n_frames / n_pcm_frames / consumed / stop (frames and gcs stay NULL: the descriptors exist on the device only,
 * p3_batch_download_desc() fetches them).  The state is not advanced: p3_decode_raw() does that. */
int p3_batch_upload_raw(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, int raw_on_device, const p3_parse_opts *opts,
                        const p3_parse_state *state, p3_parsed *info);
int p3_batch_channels(p3_ctx *c);                     /* channels of the staged batch (1 or 2) */
float p3_hop_ms(p3_ctx *c);                           /* device time of the last device hop (CUDA events: kernels + result read-backs) */
int p3_hop_rounds(p3_ctx *c);                         /* resolution rounds the last device hop needed (1 = every speculative chain met the true one) */
/* device hop + decode + download in one call: the counterpart of p3_parse() + p3_decode_batch(); *state is advanced,
 * info->consumed says where the next call continues; at most pcm_cap_frames frames are decoded (<= 0: no limit) */
int p3_decode_raw(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, const p3_parse_opts *opts, p3_parse_state *state, p3_parsed *info,
                  int16_t *pcm, int64_t pcm_cap_frames, const p3_taps *host_taps);
/* asynchronous, double-buffered variant behind pdmp3_read(): returns as soon as the hop result is known; raw / pcm stay in
 * use until p3_batch_sync() */
int p3_decode_raw_async(p3_ctx *c, const uint8_t *raw, uint64_t raw_bytes, const p3_parse_opts *opts, p3_parse_state *state, p3_parsed *info, int16_t *pcm);
int p3_batch_run(p3_ctx *c);                          /* launches the kernel sequence on the ctx stream */
int p3_batch_sync(p3_ctx *c);
int p3_batch_download(p3_ctx *c, int16_t *pcm, const p3_taps *host_taps);
int p3_batch_download_desc(p3_ctx *c, p3_frame *frames, p3_gc *gcs);   /* the descriptors as the kernels see them (tests of the device parser) */
void *p3_batch_pcm_device(p3_ctx *c, uint64_t *bytes);
void *p3_ctx_stream(p3_ctx *c);                       /* cudaStream_t of the context */
int  p3_batch_time(p3_ctx *c, int iters, float *ms_total, float *ms_stage /*[8]: k_compact, K1, K2 or the fused synthesis, K3, K4*/);  /* CUDA-event timing of p3_batch_run */
int  p3_batch_time_xr(p3_ctx *c, int iters, float *ms_total, float *ms_stage /*[2]: k_imdct, k_polyphase*/);   /* BASELINE configs[1] resident in HBM: the two transform kernels over the spectra a P3_MODE_EXACT run of the batch left on the device */
int  p3_kernel_launch_count(p3_ctx *c);               /* kernels launched by the last p3_batch_run */

/* ---- BASELINE configs[4]: one stream on rank 0, frame-sharded over the GPUs of the box, PCM gathered to rank 0 ------------
 * One process per GPU, each with its own p3_ctx.  NCCL (loaded at run time with dlopen, called from C) carries exactly what
 * north_star names: the scatter of the compressed byte ranges from rank 0 and the gather of the PCM to rank 0 -- chunked, so
 * that blocks already decoded travel while later ones decode.  Shards are independent up to a warm-up (the frame in front of
 * the shard for the filter state + earlier frames for reservoir bytes), so the decode itself has no
 * exchange step and the result is bit-identical to the single-GPU decode of the same stream. */
typedef struct p3_dist p3_dist;
typedef struct {
  int64_t  n_frames_total, n_frames_mine, warmup_mine, chunks;
  int32_t  nch, stop, launches, pad_;        /* pad_: transport flags -- 1: PCM by copy-engine peer copies (else ncclSend/ncclRecv), 2: byte ranges pushed by rank 0's copy engine (else ncclSend/ncclRecv) */
  uint64_t consumed, bytes_in, bytes_out;    /* bytes_in: compressed bytes this rank received; bytes_out: PCM bytes it sent */
  float    ms, ms_scatter;                   /* device time (CUDA events) of this rank's part of the call / until its scatter traffic was done */
  float    ms_staged, ms_decoded;            /* ... until its bytes were staged and hopped / until its own frames were decoded */
} p3_shard_result;
int  p3_dist_unique_id(uint8_t *out256);     /* rank 0: two NCCL unique ids (scatter and gather communicators), 2 x 128 bytes, to be handed to every rank */
int  p3_dist_init(p3_ctx *c, const uint8_t *ids256, int rank, int world, p3_dist **out);
void p3_dist_destroy(p3_dist *d);
int  p3_dist_nccl_version(void);
int  p3_dist_gather_transport(p3_dist *d);   /* 1: PCM blocks go to rank 0 as copy-engine peer copies through a CUDA IPC mapping of its output buffer
                                                (no SM on either end); 0: ncclSend / ncclRecv (P3_GATHER=nccl, or the mapping is not possible) */
/* collective over all ranks; rank 0 passes the stream (host memory, or device memory when raw_on_device: used in place, 64
 * readable bytes must follow), the others NULL.  Afterwards rank 0 holds the PCM of the whole stream in the context's PCM
 * buffer: p3_batch_pcm_device() / p3_batch_download(). */
int  p3_sharded_decode(p3_dist *d, const uint8_t *raw, uint64_t raw_bytes, int raw_on_device, const p3_parse_opts *opts,
                       int64_t chunk_frames, p3_shard_result *res);
/* the floor of the gather on this box: every rank > 0 sends bytes_per_rank to rank 0, nothing else running */
int64_t p3_dist_chunk_start(int64_t j, int64_t chunk_frames, int64_t n_frames);   /* first frame of chunk j of a shard's schedule (== n_frames past the end) */
int  p3_dist_measure_ingest(p3_dist *d, uint64_t bytes_per_rank, int iters, float *ms_per_iter);

#ifdef __cplusplus
}
#endif
#endif
