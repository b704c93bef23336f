/* p3_dist.cuh -- BASELINE configs[4]: one stream held on rank 0, frame-sharded over the GPUs of the box, PCM gathered to rank 0
 * (SURVEY 8e).  Included at the end of p3_cabi.cu.  NCCL is called from here (C side), loaded at run time with dlopen so
 * that the library neither links against nor needs NCCL for single-GPU use.
 *
 *   rank 0                                               rank r > 0
 *   hop of the whole stream on the device (p3_hop.cu)
 *   k_shard_plan: frame range, warm-up, byte range        |
 *   of every rank  ---- ncclBroadcast (plan) ---------->  plan
 *   ncclSend bytes of rank 1, 2, ...  (comm S, in rank    ncclRecv its byte range (previous frame for the filter state + the
 *   order: rank 1 can start while rank 7 still waits)     frames before it that hold reservoir bytes: the warm-up)
 *   decodes its own shard in chunks, straight into the    device hop of the range, side info, then chunk by chunk:
 *   output buffer                                         K0 -> K1 -> synthesis -> ncclSend of the chunk's PCM (comm G)
 *   ncclRecv of every (rank, chunk) PCM block straight      while the next chunk decodes
 *   into its place in the output buffer, posted in the
 *   order in which the blocks become ready
 *
 * Two communicators so that the scatter (rank 0 egress) and the gather (rank 0 ingress) run side by side: NVLink is full
 * duplex and rank 0's ingress -- 4608 bytes per frame from W-1 peers -- is the floor of this configuration.
 *
 * Gather transport.  NCCL's point-to-point kernels occupy SMs on both ends while the decode kernels want all of them (measured
 * at 2 GPUs: the decode of 10^6 frames takes 11.7-13.5 ms instead of 9.6 ms next to them).  Where CUDA IPC works between the
 * ranks (one box, one container: it does) rank 0 therefore exports its output buffer, every rank maps it, and a finished
 * chunk's PCM goes straight to its place in rank 0's HBM with a copy-engine peer copy over NVLink (no SM involved, no
 * matching receive to schedule); NCCL then only carries the plan, the byte ranges and a completion token per rank.
 * P3_GATHER=nccl forces the ncclSend/ncclRecv gather (the fallback when a mapping fails).
 * No collective touches the decode itself: shards are independent (warm-up rule of SURVEY 3.5, as pdmp3_b200/shard.py).
 */
#include <dlfcn.h>
#include <nccl.h>

struct nccl_api {
  void *h;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  const char *(*GetErrorString)(ncclResult_t);
  ncclResult_t (*GetVersion)(int *);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
};
static nccl_api g_nccl;

static int nccl_load(void)
{
  if (g_nccl.h) return P3_OK;
  /* RTLD_NOLOAD first: a host process that already carries NCCL (e.g. PyTorch's bundled copy) must not get a second one */
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(P3_ENODEV, "libnccl.so.2 not found: the multi-GPU path needs NCCL (%s)", dlerror());
#define NCCL_SYM(field, name) do { *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return fail(P3_ENODEV, "NCCL symbol %s missing", name); } while (0)
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); NCCL_SYM(CommInitRank, "ncclCommInitRank"); NCCL_SYM(CommDestroy, "ncclCommDestroy");
  NCCL_SYM(Send, "ncclSend"); NCCL_SYM(Recv, "ncclRecv"); NCCL_SYM(Broadcast, "ncclBroadcast");
  NCCL_SYM(GroupStart, "ncclGroupStart"); NCCL_SYM(GroupEnd, "ncclGroupEnd"); NCCL_SYM(GetErrorString, "ncclGetErrorString");
  NCCL_SYM(GetVersion, "ncclGetVersion"); NCCL_SYM(AllReduce, "ncclAllReduce");
#undef NCCL_SYM
  g_nccl.h = h;
  return P3_OK;
}
#define NK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return fail(P3_ECUDA, "%s: %s (%s:%d)", #x, g_nccl.GetErrorString(r_), __FILE__, __LINE__); } while (0)

typedef struct {                          /* one per rank, computed on rank 0 (k_shard_plan), broadcast */
  int64_t  first, last, warmup;           /* frames [first, last) of the stream are this rank's; `warmup` frames in front of them are decoded for state only */
  uint64_t byte_lo, byte_hi;              /* byte range of the stream the rank needs */
  uint64_t ms_bytes;                      /* main-data bytes of frames [first - warmup, last) */
} p3_shard_plan;

typedef struct {                          /* broadcast in front of the plans */
  int64_t n_frames; int32_t nch, iso; uint32_t maxg, stop; uint64_t consumed;
} p3_shard_head;

typedef struct {                          /* rank 0's buffers, for the peers to map (travels with the plan) */
  cudaIpcMemHandle_t h; uint64_t bytes; uint64_t serial;          /* the output buffer (PCM of the whole stream) */
} p3_ipc_msg;

typedef struct { cudaIpcMemHandle_t h; uint64_t serial; uint64_t bytes; } p3_ipc_peer;   /* a peer's staging buffer for its byte range, for rank 0 to map */

#define P3_DIST_MAXW 16

struct p3_dist {
  p3_ctx *c; int rank, world;
  ncclComm_t comm_s, comm_g;             /* scatter (+ plan broadcast) | gather */
  cudaStream_t s_scatter, s_gather;
  cudaEvent_t ev_a, ev_b, ev_t0, ev_t1, ev_s1, ev_h, ev_d;
  cudaEvent_t *ev_chunk; int n_ev;
  uint8_t *d_plan, *h_plan;              /* p3_shard_head + world x p3_shard_plan */
  /* copy-engine gather through a CUDA IPC mapping of rank 0's output buffer */
  int use_ipc;                           /* 1: agreed by all ranks at init */
  p3_ipc_msg *d_ipc, *h_ipc;             /* message buffer (device for the broadcast, page-locked host copy) */
  void *exp_ptr; size_t exp_cap; uint64_t exp_serial; cudaIpcMemHandle_t exp_handle;   /* rank 0: the allocation the current handle stands for */
  void *map_ptr; uint64_t map_serial;    /* peers: the mapping currently open */
  /* copy-engine scatter: rank 0 maps every peer's staging buffer and pushes the byte ranges into them */
  p3_ipc_peer *d_peer, *h_peer;          /* [world] message buffers */
  void *peer_map[P3_DIST_MAXW]; uint64_t peer_serial[P3_DIST_MAXW];
  void *exp_in; size_t exp_in_cap; uint64_t exp_in_serial; cudaIpcMemHandle_t exp_in_handle;    /* peers: the staging allocation the current handle stands for */
  int *d_tok;                            /* completion tokens */
};

/* plan of every rank from the frame records of the whole stream (same rule as pdmp3_b200/shard.py::plan_shards): one thread per rank */
extern "C" __global__ void k_shard_plan(const p3_frame *__restrict__ fr, int64_t n, int world, p3_shard_plan *__restrict__ plan)
{
  const int r = threadIdx.x;
  if (r >= world) return;
  const int64_t a = n * r / world, b = n * (r + 1) / world;
  int64_t w = a;
  if (a > 0 && b > a) {
    w = a - 1;                                              /* the frame that must decode correctly: it primes the IMDCT overlap and the polyphase history */
    const int64_t need = (int64_t)fr[w].main_pos - fr[w].main_begin;
    while (w > 0 && (int64_t)fr[w].main_pos > need) w--;    /* earlier frames: reservoir bytes only */
  }
  p3_shard_plan p;
  p.first = a; p.last = b; p.warmup = a - w;
  p.byte_lo = w == 0 ? 0 : fr[w - 1].main_off + fr[w - 1].main_size;
  p.byte_hi = b <= a ? p.byte_lo : fr[b - 1].main_off + fr[b - 1].main_size;
  p.ms_bytes = b <= a ? 0 : fr[b - 1].main_pos + fr[b - 1].main_size - fr[w].main_pos;
  plan[r] = p;
}

extern "C" int p3_dist_unique_id(uint8_t *out /* 2 x 128 bytes */)
{
  int rc = nccl_load(); if (rc) return rc;
  if (!out) return fail(P3_EINVAL, "null argument");
  ncclUniqueId a, b;
  NK(g_nccl.GetUniqueId(&a)); NK(g_nccl.GetUniqueId(&b));
  memcpy(out, &a, 128); memcpy(out + 128, &b, 128);
  return P3_OK;
}

extern "C" void p3_dist_destroy(p3_dist *d)
{
  if (!d) return;
  cudaSetDevice(d->c->device);
  if (d->comm_s) g_nccl.CommDestroy(d->comm_s);
  if (d->comm_g) g_nccl.CommDestroy(d->comm_g);
  if (d->s_scatter) cudaStreamDestroy(d->s_scatter);
  if (d->s_gather) cudaStreamDestroy(d->s_gather);
  cudaEvent_t evs[] = {d->ev_a, d->ev_b, d->ev_t0, d->ev_t1, d->ev_s1, d->ev_h, d->ev_d};
  for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
  for (int i = 0; i < d->n_ev; i++) cudaEventDestroy(d->ev_chunk[i]);
  free(d->ev_chunk);
  cudaFree(d->d_plan); if (d->h_plan) cudaFreeHost(d->h_plan);
  if (d->map_ptr) cudaIpcCloseMemHandle(d->map_ptr);
  for (int r = 0; r < P3_DIST_MAXW; r++) if (d->peer_map[r]) cudaIpcCloseMemHandle(d->peer_map[r]);
  cudaFree(d->d_peer); if (d->h_peer) cudaFreeHost(d->h_peer);
  cudaFree(d->d_ipc); if (d->h_ipc) cudaFreeHost(d->h_ipc); cudaFree(d->d_tok);
  free(d);
}

/* every rank calls this with the 256 bytes rank 0 got from p3_dist_unique_id() (how they travel is the caller's business:
 * MPI, a file, torch.distributed ...).  The context's device must be the rank's GPU. */
extern "C" int p3_dist_init(p3_ctx *c, const uint8_t *ids, int rank, int world, p3_dist **out)
{
  if (!c || !ids || !out || rank < 0 || rank >= world || world > P3_DIST_MAXW) return fail(P3_EINVAL, "bad argument");
  int rc = nccl_load(); if (rc) return rc;
  CK(cudaSetDevice(c->device));
  p3_dist *d = (p3_dist *)calloc(1, sizeof *d);
  if (!d) return fail(P3_ENOMEM, "calloc");
  d->c = c; d->rank = rank; d->world = world;
  ncclUniqueId a, b; memcpy(&a, ids, 128); memcpy(&b, ids + 128, 128);
#define DK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { p3_dist_destroy(d); return fail(P3_ECUDA, "%s: %s", #x, cudaGetErrorString(e_)); } } while (0)
#define DN(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { p3_dist_destroy(d); return fail(P3_ECUDA, "%s: %s", #x, g_nccl.GetErrorString(r_)); } } while (0)
  DN(g_nccl.CommInitRank(&d->comm_s, world, a, rank));
  DN(g_nccl.CommInitRank(&d->comm_g, world, b, rank));
  DK(cudaStreamCreateWithFlags(&d->s_scatter, cudaStreamNonBlocking));
  DK(cudaStreamCreateWithFlags(&d->s_gather, cudaStreamNonBlocking));
  DK(cudaEventCreateWithFlags(&d->ev_a, cudaEventDisableTiming)); DK(cudaEventCreateWithFlags(&d->ev_b, cudaEventDisableTiming));
  DK(cudaEventCreate(&d->ev_t0)); DK(cudaEventCreate(&d->ev_t1)); DK(cudaEventCreate(&d->ev_s1)); DK(cudaEventCreate(&d->ev_h)); DK(cudaEventCreate(&d->ev_d));
  const size_t pb = sizeof(p3_shard_head) + (size_t)world * sizeof(p3_shard_plan) + sizeof(p3_ipc_msg);
  DK(cudaMalloc(&d->d_plan, pb)); DK(cudaHostAlloc((void **)&d->h_plan, pb, cudaHostAllocPortable));
  DK(cudaMalloc(&d->d_peer, (size_t)world * sizeof(p3_ipc_peer))); DK(cudaHostAlloc((void **)&d->h_peer, (size_t)world * sizeof(p3_ipc_peer), cudaHostAllocPortable));
  DK(cudaMalloc(&d->d_ipc, sizeof(p3_ipc_msg))); DK(cudaHostAlloc((void **)&d->h_ipc, sizeof(p3_ipc_msg), cudaHostAllocPortable));
  DK(cudaMalloc(&d->d_tok, (size_t)(world + 1) * sizeof(int))); DK(cudaMemset(d->d_tok, 0, (size_t)(world + 1) * sizeof(int)));
#undef DK
#undef DN
  /* can every rank map rank 0's memory?  (one trial allocation; the answer is agreed with an all-reduce) */
  {
    const char *e = getenv("P3_GATHER");
    int want = !(e && !strcmp(e, "nccl")), ok = want;
    void *probe = NULL, *mapped = NULL;
    if (rank == 0) { if (cudaMalloc(&probe, 1 << 20) != cudaSuccess || cudaIpcGetMemHandle(&d->h_ipc->h, probe) != cudaSuccess) ok = 0; }
    if (rank == 0) cudaMemcpy(d->d_ipc, d->h_ipc, sizeof(p3_ipc_msg), cudaMemcpyHostToDevice);
    if (g_nccl.Broadcast(d->d_ipc, d->d_ipc, sizeof(p3_ipc_msg), ncclUint8, 0, d->comm_s, d->s_scatter) != ncclSuccess) ok = 0;
    cudaMemcpyAsync(d->h_ipc, d->d_ipc, sizeof(p3_ipc_msg), cudaMemcpyDeviceToHost, d->s_scatter); cudaStreamSynchronize(d->s_scatter);
    if (rank != 0 && ok) { if (cudaIpcOpenMemHandle(&mapped, d->h_ipc->h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); } }
    int *flag = d->d_tok + world;
    cudaMemcpy(flag, &ok, sizeof(int), cudaMemcpyHostToDevice);
    if (g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMin, d->comm_s, d->s_scatter) != ncclSuccess) ok = 0;
    int agreed = 0; cudaMemcpyAsync(&agreed, flag, sizeof(int), cudaMemcpyDeviceToHost, d->s_scatter); cudaStreamSynchronize(d->s_scatter);
    if (mapped) cudaIpcCloseMemHandle(mapped);
    /* (rank 0 frees the probe only after every peer has closed it: one more collective as the barrier) */
    g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMin, d->comm_s, d->s_scatter); cudaStreamSynchronize(d->s_scatter);
    if (probe) cudaFree(probe);
    d->use_ipc = ok && agreed;
    cudaGetLastError();
  }
  *out = d;
  return P3_OK;
}

extern "C" int p3_dist_gather_transport(p3_dist *d) { return d ? d->use_ipc : -1; }     /* 1: copy-engine peer copies through CUDA IPC, 0: ncclSend / ncclRecv */

extern "C" int p3_dist_nccl_version(void) { int v = 0; if (nccl_load() == P3_OK) g_nccl.GetVersion(&v); return v; }

static int dist_events(p3_dist *d, int n)
{
  if (n <= d->n_ev) return P3_OK;
  cudaEvent_t *e = (cudaEvent_t *)realloc(d->ev_chunk, (size_t)n * sizeof *e);
  if (!e) return fail(P3_ENOMEM, "realloc");
  d->ev_chunk = e;
  for (int i = d->n_ev; i < n; i++) CK(cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming));
  d->n_ev = n;
  return P3_OK;
}

/* Chunk schedule of a shard of nf frames (warm-up included): small chunks first so that PCM is on the wire early (C/4, C/4,
 * C/2), C frames in the middle -- few enough waves of CTAs per launch that the tail of each launch stays small --, small
 * chunks again at the end (C/2, C/4, C/4) so that little is left to send once the last kernel has finished.  Every rank
 * derives the same schedule from the plan.  Returns the first frame of chunk j (== nf once past the end). */
static inline int64_t chunk_start(int64_t j, int64_t C, int64_t nf)
{
  const int64_t q = (C / 4) - (C / 4) % K1_FPB;
  if (q < K1_FPB || nf < 3 * C) { const int64_t f = j * (q >= K1_FPB && nf < 3 * C && nf >= 8 * q ? q : C); return f < nf ? f : nf; }   /* short shards: uniform chunks */
  const int64_t body0 = 4 * q, body1 = nf - 4 * q - (nf - 4 * q) % K1_FPB;                 /* head [0, 4q) | body | tail [body1, nf) */
  const int64_t nbody = (body1 - body0 + C - 1) / C;
  int64_t f;
  if (j <= 0) f = 0; else if (j == 1) f = q; else if (j == 2) f = 2 * q;
  else if (j < 3 + nbody) f = body0 + (j - 3) * C;
  else if (j == 3 + nbody) f = body1; else if (j == 4 + nbody) f = body1 + 2 * q; else if (j == 5 + nbody) f = body1 + 3 * q;
  else f = nf;
  return f < nf ? f : nf;
}
extern "C" int64_t p3_dist_chunk_start(int64_t j, int64_t chunk_frames, int64_t n_frames) { return chunk_start(j, chunk_frames, n_frames); }   /* (the schedule, for tests) */
static inline int64_t chunk_count(int64_t C, int64_t nf) { int64_t j = 0; while (chunk_start(j, C, nf) < nf) j++; return j; }
/* PCM slots [lo, hi) (in frames of the rank's shard) that chunk j of a shard with `wu` warm-up frames and `cnt` own frames produces */
static inline void chunk_slots(int64_t j, int64_t C, int64_t wu, int64_t cnt, int64_t *lo, int64_t *hi)
{
  const int64_t f0 = chunk_start(j, C, wu + cnt), f1 = chunk_start(j + 1, C, wu + cnt);
  *lo = (f0 > wu ? f0 : wu) - wu; *hi = f1 > wu ? f1 - wu : 0;
  if (*hi < *lo) *hi = *lo;
}

/* The sharded decode.  Rank 0: `raw` is the whole stream (host memory, or device memory when raw_on_device -- then it is used
 * in place and must be followed by 64 readable bytes); other ranks pass NULL / 0.  chunk_frames: frames per launch sequence and
 * per PCM block on the wire (<= 0: four waves of the synthesis kernel, 227 328 frames on a B200; rounded to a multiple of 32).  On return rank 0's PCM of the WHOLE stream is in the
 * context's PCM buffer (p3_batch_pcm_device(), p3_batch_download()), [n_frames][1152][nch] int16.  res (optional):
 * frames decoded by this rank, frames of the stream, device time of this rank's part in ms (CUDA events, from the first
 * byte staged to the last PCM block sent / received), and of the scatter alone. */
extern "C" int p3_sharded_decode(p3_dist *d, const uint8_t *raw, uint64_t raw_bytes, int raw_on_device, const p3_parse_opts *o_in,
                                 int64_t chunk_frames, p3_shard_result *res)
{
  if (!d) return fail(P3_EINVAL, "null argument");
  p3_ctx *c = d->c; const int W = d->world, R = d->rank;
  if (R == 0 && !raw) return fail(P3_EINVAL, "rank 0 holds the stream");
  CK(cudaSetDevice(c->device));
  int rc;
  if ((rc = p3_batch_sync(c))) return rc;
  if ((rc = p3_ctx_reset(c))) return rc;                   /* every shard starts from zero state + its warm-up (rank 0: like pdmp3_open_feed) */
  p3_parse_opts o; if (o_in) o = *o_in; else memset(&o, 0, sizeof o);
  o.max_frames = 0; o.warmup_frames = 0; o.hop_only = 1;
  /* default chunk: four waves of the synthesis kernel (n_sm x resident warps per SM x 32 frames = 56 832 frames on a B200), so that the
   * quarter chunks at both ends of the schedule are exactly one wave and no launch ends in a mostly empty wave */
  int64_t C = chunk_frames > 0 ? chunk_frames : 4 * (int64_t)c->n_sm * p3_synthw_warps_per_sm() * c->fpc;
  C -= C % K1_FPB; if (C < K1_FPB) C = K1_FPB;
  c->chunk_frames = C; c->taps = 0;
  p3_slot *sl = &c->slot[c->cur_slot];
  slot_release(sl);
  p3_shard_head *hd = (p3_shard_head *)d->h_plan; p3_shard_plan *pl = (p3_shard_plan *)(d->h_plan + sizeof(p3_shard_head));
  const size_t pb = sizeof(p3_shard_head) + (size_t)W * sizeof(p3_shard_plan) + sizeof(p3_ipc_msg);
  p3_ipc_msg *ipc = (p3_ipc_msg *)(d->h_plan + sizeof(p3_shard_head) + (size_t)W * sizeof(p3_shard_plan));
  const p3_parse_state st0 = {0, 0, 0, -1, -1};
  CK(cudaEventRecord(d->ev_t0, c->stream));

  /* ---- rank 0: stage, hop, plan ---- */
  if (R == 0) {
    if (raw_on_device) sl->raw_dev = raw;
    else {
      if ((rc = ensure(&sl->raw, raw_bytes + 64))) return rc;
      sl->raw_dev = (const uint8_t *)sl->raw.p;
      CK(cudaMemcpyAsync(sl->raw.p, raw, raw_bytes, cudaMemcpyHostToDevice, c->stream));
      CK(cudaMemsetAsync((uint8_t *)sl->raw.p + raw_bytes, 0, 64, c->stream));
    }
    if ((rc = hop_staged(c, sl, raw_bytes, &o, &st0, c->stream))) return rc;
    const p3_hop_result *r = c->hop.h_res;
    memset(hd, 0, sizeof *hd);
    hd->n_frames = r->n_frames; hd->nch = r->nch; hd->iso = (int32_t)o.iso; hd->maxg = r->maxg; hd->stop = (uint32_t)r->stop; hd->consumed = r->consumed;
    CK(cudaMemcpyAsync(d->d_plan, hd, sizeof *hd, cudaMemcpyHostToDevice, c->stream));
    k_shard_plan<<<1, 32, 0, c->stream>>>((const p3_frame *)sl->frames.p, r->n_frames, W, (p3_shard_plan *)(d->d_plan + sizeof(p3_shard_head)));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(pl, d->d_plan + sizeof(p3_shard_head), (size_t)W * sizeof(p3_shard_plan), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    /* the output buffer (PCM of the whole stream) and, for the copy-engine gather, its IPC handle: it travels with the plan */
    c->n_frames = r->n_frames; c->n_pcm_frames = r->n_frames; c->nch = (uint32_t)r->nch;
    if ((rc = size_batch(c, sl, r->n_frames, r->n_frames, r->maxg, pl[0].ms_bytes, c->stream))) return rc;
    memset(ipc, 0, sizeof *ipc);
    if (d->use_ipc && W > 1) {
      if (d->exp_ptr != sl->pcm.p || d->exp_cap != sl->pcm.cap) {      /* (a new allocation may come back at the old address: the size tells) */
        CK(cudaIpcGetMemHandle(&d->exp_handle, sl->pcm.p)); d->exp_ptr = sl->pcm.p; d->exp_cap = sl->pcm.cap; d->exp_serial++;
      }
      ipc->h = d->exp_handle; ipc->bytes = sl->pcm.cap; ipc->serial = d->exp_serial;
    }
    CK(cudaMemcpyAsync(d->d_plan + pb - sizeof(p3_ipc_msg), ipc, sizeof *ipc, cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(d->ev_a, c->stream));
    CK(cudaStreamWaitEvent(d->s_scatter, d->ev_a, 0));
  }
  NK(g_nccl.Broadcast(d->d_plan, d->d_plan, pb, ncclUint8, 0, d->comm_s, d->s_scatter));
  if (R != 0) {
    CK(cudaMemcpyAsync(d->h_plan, d->d_plan, pb, cudaMemcpyDeviceToHost, d->s_scatter));
    CK(cudaStreamSynchronize(d->s_scatter));
    if (d->use_ipc && (!d->map_ptr || d->map_serial != ipc->serial)) {      /* map rank 0's output buffer (kept until it moves) */
      if (d->map_ptr) { cudaIpcCloseMemHandle(d->map_ptr); d->map_ptr = NULL; }
      CK(cudaIpcOpenMemHandle(&d->map_ptr, ipc->h, cudaIpcMemLazyEnablePeerAccess));
      d->map_serial = ipc->serial;
    }
  }
  const int push = d->use_ipc && W > 1;                     /* scatter by copy-engine pushes out of rank 0 (agreed at init) */
  const p3_shard_plan me = pl[R];
  const int64_t n_total = hd->n_frames; const int nch = hd->nch;
  o.iso = (uint32_t)hd->iso;
  const size_t fbytes = (size_t)1152 * sizeof(int16_t) * (size_t)nch;

  /* ---- scatter: the byte ranges leave rank 0 in rank order (rank 1 can start decoding while rank 7 still waits).
   *      Copy-engine variant: every peer exports its staging buffer, rank 0 maps them and pushes each range with a peer copy
   *      (writes travel on rank 0's egress, which nothing else uses; reads pulled by the peers would have to get their requests
   *      through rank 0's ingress, where the PCM arrives: measured, the last ranks got their bytes after 24-40 ms), followed
   *      by a token; NCCL variant: rank 0 sends the ranges one after the other. ---- */
  if (R == 0) {
    if (push) {
      NK(g_nccl.GroupStart());
      for (int r = 1; r < W; r++) NK(g_nccl.Recv(d->d_peer + r, sizeof(p3_ipc_peer), ncclUint8, r, d->comm_s, d->s_scatter));
      NK(g_nccl.GroupEnd());
      CK(cudaMemcpyAsync(d->h_peer, d->d_peer, (size_t)W * sizeof(p3_ipc_peer), cudaMemcpyDeviceToHost, d->s_scatter));
      CK(cudaStreamSynchronize(d->s_scatter));
      for (int r = 1; r < W; r++) {
        if (!d->peer_map[r] || d->peer_serial[r] != d->h_peer[r].serial) {
          if (d->peer_map[r]) { cudaIpcCloseMemHandle(d->peer_map[r]); d->peer_map[r] = NULL; }
          CK(cudaIpcOpenMemHandle(&d->peer_map[r], d->h_peer[r].h, cudaIpcMemLazyEnablePeerAccess));
          d->peer_serial[r] = d->h_peer[r].serial;
        }
        const uint64_t len = pl[r].byte_hi - pl[r].byte_lo;
        if (len) CK(cudaMemcpyAsync(d->peer_map[r], sl->raw_dev + pl[r].byte_lo, len, cudaMemcpyDeviceToDevice, d->s_scatter));
        NK(g_nccl.Send(d->d_tok, sizeof(int), ncclUint8, r, d->comm_s, d->s_scatter));
      }
    } else
      for (int r = 1; r < W; r++)
        if (pl[r].byte_hi > pl[r].byte_lo) NK(g_nccl.Send(sl->raw_dev + pl[r].byte_lo, pl[r].byte_hi - pl[r].byte_lo, ncclUint8, r, d->comm_s, d->s_scatter));
    CK(cudaEventRecord(d->ev_s1, d->s_scatter));
  } else {
    const uint64_t len = me.byte_hi - me.byte_lo;
    if ((rc = ensure(&sl->raw, len + 64))) return rc;
    sl->raw_dev = (const uint8_t *)sl->raw.p;
    if (push) {
      if (d->exp_in != sl->raw.p || d->exp_in_cap != sl->raw.cap) {
        CK(cudaIpcGetMemHandle(&d->exp_in_handle, sl->raw.p)); d->exp_in = sl->raw.p; d->exp_in_cap = sl->raw.cap; d->exp_in_serial++;
      }
      d->h_peer[R].h = d->exp_in_handle; d->h_peer[R].serial = d->exp_in_serial; d->h_peer[R].bytes = sl->raw.cap;
      CK(cudaMemcpyAsync(d->d_peer + R, d->h_peer + R, sizeof(p3_ipc_peer), cudaMemcpyHostToDevice, d->s_scatter));
      NK(g_nccl.Send(d->d_peer + R, sizeof(p3_ipc_peer), ncclUint8, 0, d->comm_s, d->s_scatter));
      NK(g_nccl.Recv(d->d_tok, sizeof(int), ncclUint8, 0, d->comm_s, d->s_scatter));      /* rank 0's push has landed */
    } else if (len) NK(g_nccl.Recv(sl->raw.p, len, ncclUint8, 0, d->comm_s, d->s_scatter));
    CK(cudaMemsetAsync((uint8_t *)sl->raw.p + len, 0, 64, d->s_scatter));
    CK(cudaEventRecord(d->ev_s1, d->s_scatter));
    CK(cudaStreamWaitEvent(c->stream, d->ev_s1, 0));
    p3_parse_opts oo = o; oo.warmup_frames = (uint32_t)me.warmup; oo.max_frames = me.warmup + (me.last - me.first);
    if (me.last > me.first) {
      if ((rc = hop_staged(c, sl, len, &oo, &st0, c->stream))) return rc;
      const p3_hop_result *r = c->hop.h_res;
      if (r->n_frames != me.warmup + (me.last - me.first)) return fail(P3_EINVAL, "shard of rank %d holds %lld frames, the plan says %lld", R, (long long)r->n_frames, (long long)(me.warmup + me.last - me.first));
      if ((rc = size_batch(c, sl, r->n_frames, r->n_pcm_frames, r->maxg, r->total_ms, c->stream))) return rc;
    } else { c->n_frames = 0; c->n_pcm_frames = 0; c->nch = (uint32_t)nch; }
  }

  /* ---- decode the rank's frames chunk by chunk; every finished chunk's PCM goes on the wire while the next one decodes ---- */
  const int64_t my_f0 = R == 0 ? me.first : 0, my_f1 = R == 0 ? me.last : c->n_frames;     /* frames of the staged batch this rank decodes */
  const int64_t nchunk = chunk_count(C, my_f1 - my_f0);
  if ((rc = dist_events(d, (int)(nchunk > 0 ? nchunk : 1)))) return rc;
  c->launches = 0;
  CK(cudaEventRecord(d->ev_h, c->stream));                 /* staged, hopped, planned (rank 0) / bytes received and hopped (others) */
  if (my_f1 > my_f0) {
    if ((rc = run_sideinfo_range(c, sl, my_f0, my_f1))) return rc;
    sl->hop_only = 0;
  }
  for (int64_t j = 0; j < nchunk; j++) {
    const int64_t f0 = my_f0 + chunk_start(j, C, my_f1 - my_f0), f1 = my_f0 + chunk_start(j + 1, C, my_f1 - my_f0);
    if ((rc = run_chunk(c, sl, f0, f1, NULL))) return rc;
    if (R != 0) {
      int64_t lo, hi; chunk_slots(j, C, me.warmup, me.last - me.first, &lo, &hi);
      CK(cudaEventRecord(d->ev_chunk[j], c->stream));
      CK(cudaStreamWaitEvent(d->s_gather, d->ev_chunk[j], 0));
      if (hi > lo) {
        if (d->use_ipc)                                      /* copy engine, straight into rank 0's output buffer over NVLink */
          CK(cudaMemcpyAsync((uint8_t *)d->map_ptr + (size_t)(me.first + lo) * fbytes, (const uint8_t *)sl->pcm.p + (size_t)lo * fbytes, (size_t)(hi - lo) * fbytes, cudaMemcpyDeviceToDevice, d->s_gather));
        else NK(g_nccl.Send((const uint8_t *)sl->pcm.p + (size_t)lo * fbytes, (size_t)(hi - lo) * fbytes, ncclUint8, 0, d->comm_g, d->s_gather));
      }
    }
  }
  if (d->use_ipc && W > 1) {                                 /* completion: a token per rank once its copies have landed */
    if (R != 0) NK(g_nccl.Send(d->d_tok + R, sizeof(int), ncclUint8, 0, d->comm_g, d->s_gather));
    else {
      NK(g_nccl.GroupStart());
      for (int r = 1; r < W; r++) NK(g_nccl.Recv(d->d_tok + r, sizeof(int), ncclUint8, r, d->comm_g, d->s_gather));
      NK(g_nccl.GroupEnd());
    }
  }
  CK(cudaEventRecord(d->ev_d, c->stream));                 /* this rank's own frames are decoded */
  /* ---- gather on rank 0: every (rank, chunk) block straight into its place, posted in the order the blocks become ready:
   *      rank r's bytes leave rank 0 after those of ranks 1..r-1, then its chunks follow at the decode rate ---- */
  if (R == 0 && W > 1 && !d->use_ipc) {
    struct blk { double t; int r; int64_t lo, hi; } *bl; int nb = 0;
    int64_t maxblk = 0; for (int r = 1; r < W; r++) maxblk += chunk_count(C, pl[r].warmup + pl[r].last - pl[r].first) + 1;
    bl = (blk *)malloc((size_t)(maxblk > 0 ? maxblk : 1) * sizeof *bl);
    if (!bl) return fail(P3_ENOMEM, "malloc");
    double t_sc = 0;
    for (int r = 1; r < W; r++) {
      t_sc += (double)(pl[r].byte_hi - pl[r].byte_lo) / 650e3;                      /* ~650 GB/s on the wire, in microseconds */
      const int64_t wu = pl[r].warmup, cnt = pl[r].last - pl[r].first, nc = chunk_count(C, wu + cnt);
      for (int64_t j = 0; j < nc; j++) {
        int64_t lo, hi; chunk_slots(j, C, wu, cnt, &lo, &hi);
        if (hi > lo) { bl[nb].t = t_sc + (double)chunk_start(j + 1, C, wu + cnt) * 0.0105; bl[nb].r = r; bl[nb].lo = lo; bl[nb].hi = hi; nb++; }   /* ~10.5 ns of kernels per frame */
      }
    }
    for (int i = 1; i < nb; i++) { blk x = bl[i]; int k = i - 1; while (k >= 0 && bl[k].t > x.t) { bl[k + 1] = bl[k]; k--; } bl[k + 1] = x; }   /* stable insertion sort: the order per rank is kept */
    int G = 4; { const char *e = getenv("P3_GATHER_GROUP"); if (e && atoi(e) >= 1) G = atoi(e); }
    for (int i = 0; i < nb; i += G) {
      NK(g_nccl.GroupStart());
      for (int k = i; k < i + G && k < nb; k++)
        NK(g_nccl.Recv((uint8_t *)sl->pcm.p + (size_t)(pl[bl[k].r].first + bl[k].lo) * fbytes, (size_t)(bl[k].hi - bl[k].lo) * fbytes, ncclUint8, bl[k].r, d->comm_g, d->s_gather));
      NK(g_nccl.GroupEnd());
    }
    free(bl);
  }
  /* ---- join ---- */
  CK(cudaEventRecord(d->ev_b, d->s_gather));
  CK(cudaStreamWaitEvent(c->stream, d->ev_b, 0));
  if (R == 0) CK(cudaStreamWaitEvent(c->stream, d->ev_s1, 0));
  CK(cudaEventRecord(d->ev_t1, c->stream));
  CK(cudaStreamSynchronize(c->stream)); CK(cudaStreamSynchronize(d->s_scatter)); CK(cudaStreamSynchronize(d->s_gather));
  c->have_next_tail_dev = 0; c->have_next_tail = 0;        /* a sharded decode is one whole stream: nothing is carried into a next call */
  if (R == 0) { c->n_frames = n_total; c->n_pcm_frames = n_total; }
  if (res) {
    memset(res, 0, sizeof *res);
    res->n_frames_total = n_total; res->n_frames_mine = me.last - me.first; res->warmup_mine = me.warmup; res->nch = nch;
    res->stop = (int32_t)hd->stop; res->consumed = hd->consumed; res->chunks = nchunk; res->launches = c->launches;
    res->pad_ = (d->use_ipc ? 1 : 0) | (push ? 2 : 0);
    res->bytes_in = R == 0 ? 0 : me.byte_hi - me.byte_lo; res->bytes_out = R == 0 ? 0 : (uint64_t)(me.last - me.first) * fbytes;
    CK(cudaEventElapsedTime(&res->ms, d->ev_t0, d->ev_t1));
    CK(cudaEventElapsedTime(&res->ms_scatter, d->ev_t0, d->ev_s1));
    CK(cudaEventElapsedTime(&res->ms_staged, d->ev_t0, d->ev_h));
    CK(cudaEventElapsedTime(&res->ms_decoded, d->ev_t0, d->ev_d));
  }
  return P3_OK;
}

/* Floor of the gather: every rank r > 0 sends bytes_per_rank to rank 0 with nothing else going on (grouped receives from all
 * peers at once), `iters` times; *gbs = bytes that entered rank 0 per second (max over iterations is NOT taken: the mean). */
extern "C" int p3_dist_measure_ingest(p3_dist *d, uint64_t bytes_per_rank, int iters, float *ms_per_iter)
{
  if (!d || iters <= 0) return fail(P3_EINVAL, "bad argument");
  p3_ctx *c = d->c; const int W = d->world, R = d->rank;
  CK(cudaSetDevice(c->device));
  dbuf *scratch = &c->slot[c->cur_slot ^ 1].pcm;
  int rc = ensure(scratch, R == 0 ? bytes_per_rank * (uint64_t)(W > 1 ? W - 1 : 1) : bytes_per_rank);
  if (rc) return rc;
  for (int it = -1; it < iters; it++) {                      /* one untimed round first */
    if (it == 0) CK(cudaEventRecord(d->ev_t0, d->s_gather));
    if (R == 0) {
      NK(g_nccl.GroupStart());
      for (int r = 1; r < W; r++) NK(g_nccl.Recv((uint8_t *)scratch->p + (size_t)(r - 1) * bytes_per_rank, bytes_per_rank, ncclUint8, r, d->comm_g, d->s_gather));
      NK(g_nccl.GroupEnd());
    } else NK(g_nccl.Send(scratch->p, bytes_per_rank, ncclUint8, 0, d->comm_g, d->s_gather));
  }
  CK(cudaEventRecord(d->ev_t1, d->s_gather));
  CK(cudaStreamSynchronize(d->s_gather));
  float ms; CK(cudaEventElapsedTime(&ms, d->ev_t0, d->ev_t1));
  if (ms_per_iter) *ms_per_iter = ms / iters;
  return P3_OK;
}
