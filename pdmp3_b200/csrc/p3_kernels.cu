/* p3_kernels.cu -- hand-written sm_100a kernels for the MP3 Layer III granule decode path.
 *
 *   k_compact   K0  bit reservoir: header-stripped main-data stream    (Get_Main_Data 1096-1122)
 *   k_huffman   K1  scalefactors + Huffman      (Read_Main_L3 1376-1435, Read_Huffman 2051-2115,
 *                                                Huffman_Decode 1593-1643)
 *   k_requant   K2  requantize+reorder+stereo+antialias, fused
 *                                               (L3_Requantize 1829-1905, L3_Reorder 1786-1823,
 *                                                L3_Stereo 1911-1972, L3_Antialias 1706-1732)
 *   k_imdct     K3  IMDCT + window + overlap-add + frequency inversion
 *                                               (IMDCT_Win 1649-1700, L3_Hybrid_Synthesis 1752-1780,
 *                                                L3_Frequency_Inversion 1738-1746)
 *   k_polyphase K4  matrixing + 512-tap window + int16 pack
 *                                               (L3_Subband_Synthesis 1978-2045, Convert_Frame_S16 2307-2345)
 *
 * Arithmetic notes: the reference is compiled without FMA, sums run in index
 * order in fp32, MS stereo multiplies by a double constant, the final PCM scale is a double
 * multiply followed by truncation.  In P3_MODE_EXACT every one of these roundings is reproduced
 * with __fmul_rn/__fadd_rn (which nvcc never contracts), so PCM is bit-identical.
 */
#include "p3_device.cuh"
#include "p3_kernels.h"

/* =============================================================================================
 * K0: k_compact -- the bit reservoir of Get_Main_Data (pdmp3.c:1096-1122) for the whole batch at once.
 * Copies the main data of every frame (the bytes between the side info and the next header) into ONE contiguous,
 * header-stripped stream of big-endian 32-bit words in global memory: byte 512 of the stream is the first
 * main-data byte of the batch, bytes 0..511 are the reservoir carried over from the previous batch.  A frame whose
 * main_data_begin reaches back then simply starts main_data_begin bytes earlier in the same stream.
 * One warp per frame: whole destination words through two aligned loads + funnel shift, the ragged ends (a word
 * can straddle two frames' data) byte by byte.
 * ============================================================================================= */
extern "C" __global__ void __launch_bounds__(128)
k_compact(const uint8_t *__restrict__ raw, const p3_frame *__restrict__ frames, const uint8_t *__restrict__ tail,
          int64_t f_first, int64_t f_end, uint32_t *__restrict__ ms)
{
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *mb = reinterpret_cast<uint8_t *>(ms);
  if (f_first == 0 && blockIdx.x == 0)                     /* the 512 bytes in front of the batch */
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) mb[i ^ 3u] = tail[i];
  const int64_t fs = f_first + (int64_t)blockIdx.x * 4 + warp;
  if (fs >= f_end) return;
  const uint4 fa = __ldg(reinterpret_cast<const uint4 *>(frames + fs)), fb = __ldg(reinterpret_cast<const uint4 *>(frames + fs) + 1);
  const uint64_t main_off = (uint64_t)fa.x | (uint64_t)fa.y << 32, main_pos = (uint64_t)fa.z | (uint64_t)fa.w << 32;
  const uint64_t base0 = frames[0].main_pos;
  const int32_t n = (int32_t)(fb.x & 0xffffu);
  const uint8_t *src = raw + main_off;
  const uint64_t d0 = 512 + (main_pos - base0);            /* stream byte of the frame's first data byte */
  const int32_t lead = (int32_t)((4 - (d0 & 3)) & 3);     /* bytes up to the next word boundary */
  const int32_t nw = lead < n ? (n - lead) >> 2 : 0;       /* whole words */
  for (int32_t b = (int32_t)lane; b < min(lead, n); b += 32) mb[(d0 + b) ^ 3u] = src[b];
  for (int32_t b = lead + 4 * nw + (int32_t)lane; b < n; b += 32) mb[(d0 + b) ^ 3u] = src[b];
  const uintptr_t sa = (uintptr_t)(src + lead);
  const uint32_t *al = reinterpret_cast<const uint32_t *>(sa & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(sa & 3) * 8;
  uint32_t *dw = ms + ((d0 + lead) >> 2);
  /* every lane takes its aligned source words for the whole frame first (all loads in flight together: a frame is at
   * most 1441 bytes = 12 words per lane), the word after a lane's own comes from its neighbour by shuffle */
  uint32_t a[12];
  #pragma unroll
  for (int j = 0; j < 12; j++) a[j] = (int32_t)(lane + 32 * j) <= nw ? __ldcs(al + lane + 32 * j) : 0u;
  #pragma unroll
  for (int j = 0; j < 12; j++) {
    uint32_t nxt = __shfl_down_sync(0xffffffffu, a[j], 1);
    const uint32_t wrap = __shfl_sync(0xffffffffu, j + 1 < 12 ? a[j + 1 < 12 ? j + 1 : j] : 0u, 0);
    if (lane == 31) nxt = wrap;
    const int32_t k = (int32_t)lane + 32 * j;
    if (k < nw) dw[k] = __byte_perm(__funnelshift_r(a[j], nxt, sh), 0, 0x0123);   /* 4 stream bytes, first byte to the MSB */
  }
}

/* =============================================================================================
 * K1: Huffman.  One CTA per group of K1_FPB frames, one thread per granule-channel.  The group's slice of the
 * compact main-data stream (its frames' data plus the 512 reservoir bytes in front: contiguous and already in
 * big-endian words thanks to k_compact) is brought into shared memory by ONE bulk copy of the TMA engine
 * (cp.async.bulk + mbarrier) while the threads load the LUT and sort their parts.
 * ============================================================================================= */
#include "p3_k1.cuh"

extern "C" __global__ void __launch_bounds__(K1_THREADS)
k_huffman(const uint32_t *__restrict__ ms /* compact main-data stream */, const p3_frame *__restrict__ frames, const p3_gc *__restrict__ gcs,
          const p3_tables *__restrict__ T, int64_t f_first, int64_t f_end /*frames [f_first,f_end) decoded by this launch*/,
          uint32_t smem_words, int16_t *__restrict__ is_out, int32_t *__restrict__ count1_out, uint8_t *__restrict__ scf_out)
{
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t *sw = sm;                                     /* window of the main-data stream, big-endian words */
  uint32_t *ring = sm + smem_words;                      /* [4][K1_THREADS] output staging */
#ifdef K1_LUT_GLOBAL
  const uint16_t *lut = T->hlut;
#else
  uint16_t *lut = reinterpret_cast<uint16_t *>(ring + 4 * K1_THREADS);
#endif
  __shared__ __align__(8) unsigned long long s_bar;

  const int64_t F0 = f_first + (int64_t)blockIdx.x * K1_FPB;
  const int64_t F1 = min(F0 + (int64_t)K1_FPB, f_end);
  /* stream byte 512 = first main-data byte of the batch; the window starts (16-byte aligned) at or below the 512
   * reservoir bytes in front of frame F0 and ends behind frame F1-1 */
  const uint64_t base0 = frames[0].main_pos;
  const int64_t win0 = (int64_t)((frames[F0].main_pos - base0) & ~(uint64_t)15);          /* stream byte of window byte 0 */
  if (threadIdx.x == 0) {
    const int64_t end = 512 + (int64_t)(frames[F1 - 1].main_pos + frames[F1 - 1].main_size - base0);
    uint32_t bytes = (uint32_t)((end - win0 + 16 + 15) & ~(int64_t)15);
    if (bytes > smem_words * 4) bytes = smem_words * 4;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\nfence.mbarrier_init.release.cluster;" :: "r"(bar) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(sw)), "l"(reinterpret_cast<const uint8_t *>(ms) + win0), "r"(bytes), "r"(bar) : "memory");
  }

#ifndef K1_LUT_GLOBAL
  for (uint32_t i = threadIdx.x; i < (T->hlut_used + 1) / 2; i += blockDim.x)
    reinterpret_cast<uint32_t *>(lut)[i] = reinterpret_cast<const uint32_t *>(T->hlut)[i];
#endif

  /* Lanes of a warp run in lock step, so a warp takes as long as its longest part.  Parts are therefore
   * handed out sorted by big_values (bitonic sort of the group's 128 keys in shared memory): each warp
   * gets parts of similar length.  Outputs are addressed by part, so nothing downstream changes. */
  __shared__ uint32_t s_key[K1_THREADS];
  {
    const int64_t fk = F0 + (threadIdx.x >> 2);
    uint32_t bv = 0;
#ifdef K1_SORT_P23L
    if (fk < F1) bv = P3_GC_P23L(gcs[4 * fk + (threadIdx.x & 3)]) >> 3;     /* experiment: total bits of the part (pairs + quads) instead of the pair count */
#else
    if (fk < F1) bv = P3_GC_BIGV(gcs[4 * fk + (threadIdx.x & 3)]);
#endif
    s_key[threadIdx.x] = ((511u - bv) << 8) | threadIdx.x;          /* descending big_values, part index in the low byte */
    __syncthreads();
    for (uint32_t k = 2; k <= K1_THREADS; k <<= 1)
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        const uint32_t i = threadIdx.x, l = i ^ j;
        if (l > i) {
          const uint32_t a = s_key[i], b = s_key[l];
          if (((i & k) == 0) == (a > b)) { s_key[i] = b; s_key[l] = a; }
        }
        __syncthreads();
      }
  }
  {                                                        /* the window has landed (the sort above ended with a barrier, so s_bar is initialised for everyone) */
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    uint32_t ok = 0, spins = 0;
    do {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar) : "memory");
      if (!ok && ++spins > (1u << 24)) __trap();
    } while (!ok);
  }
#ifdef K1_COOP_TRIAL
  {                                                        /* trial: one WARP per part, the warp's 32 parts one after the other (p3_k1.cuh) */
    const uint32_t warp = threadIdx.x >> 5;
    const int64_t o_cta0 = (F0 - f_first) * 4;
    for (uint32_t k = 0; k < 32; k++) {
      const uint32_t gi0 = warp * 32 + k;
      const int64_t f0 = F0 + (gi0 >> 2);
      if (f0 >= F1) break;
      const uint32_t gr0 = (gi0 >> 1) & 1, ch0 = gi0 & 1;
      const int64_t o0 = o_cta0 + gi0;
      const p3_frame fr0 = frames[f0]; const p3_gc g0 = gcs[4 * f0 + 2 * gr0 + ch0];
      uint8_t *scf0 = scf_out + o0 * P3_SCF_STRIDE;
      if ((threadIdx.x & 31) < 4) reinterpret_cast<uint4 *>(scf0)[threadIdx.x & 31] = make_uint4(0, 0, 0, 0);
      __syncwarp();
      const uint32_t c1v = k1_decode_gc_coop(sw, lut, T, gcs, fr0, g0, f0, gr0, ch0, win0 - 512 + (int64_t)base0, reinterpret_cast<uint32_t *>(is_out + o0 * 576), scf0);
      if ((threadIdx.x & 31) == 0) count1_out[o0] = (int32_t)c1v;
    }
    return;
  }
#endif
  const uint32_t gi = s_key[threadIdx.x] & 0xffu;         /* granule-channel within the group handled by this thread */
  const int64_t f = F0 + (gi >> 2);
  const int64_t o_cta = (F0 - f_first) * 4;
  if (f < F1) {
    const uint32_t gr = (gi >> 1) & 1, ch = gi & 1;
    const int64_t o = o_cta + gi;
    const p3_frame fr = frames[f]; const p3_gc g = gcs[4 * f + 2 * gr + ch];
    k1_out ob; ob.ring = ring + threadIdx.x; ob.stride = K1_THREADS; ob.pw = 0; ob.dst = reinterpret_cast<uint4 *>(is_out + o * 576);
    /* scalefactor bytes go straight to this part's 64-byte row (byte stores merge in L2) */
    uint8_t *scf = scf_out + o * P3_SCF_STRIDE;
    for (int q = 0; q < P3_SCF_STRIDE / 16; q++) reinterpret_cast<uint4 *>(scf)[q] = make_uint4(0, 0, 0, 0);
    count1_out[o] = (int32_t)k1_decode_gc(sw, lut, T, gcs, fr, g, f, gr, ch, win0 - 512 + (int64_t)base0, ob, scf);
  }
}

/* =============================================================================================
 * K2: requantize + reorder + stereo + antialias, one CTA per granule (both channels: stereo
 * processing couples them).  Output lines are produced directly in reordered position, so the
 * reorder of short blocks costs nothing (the reference permutes through a scratch array).
 * ============================================================================================= */
__device__ __forceinline__ float k2_requant(const p3_tables *T, int v, uint32_t e2, int q)
{
  /* (t1*t2)*t3 with two separately rounded products (pdmp3.c:2132,2150) */
  float t3 = T->pow43[v < 0 ? -v : v];
  if (v < 0) t3 = -t3;
  return __fmul_rn(__fmul_rn(T->t1h[e2], T->t2[q + P3_T2_BIAS]), t3);
}

extern "C" __global__ void __launch_bounds__(K2_THREADS)
k_requant(const p3_frame *__restrict__ frames, const p3_gc *__restrict__ gcs, const p3_tables *__restrict__ T,
          int64_t f_first, int64_t f_end, const int16_t *__restrict__ is_in, const int32_t *__restrict__ count1,
          const uint8_t *__restrict__ scf, const p3_state *__restrict__ st_in, p3_state *__restrict__ st_out,
          float *__restrict__ xr_out)
{
  __shared__ float xs[2][576];
  __shared__ uint8_t sscf[2][P3_SCF_STRIDE];
  __shared__ int32_t sc1[2];

  const int64_t gidx = blockIdx.x;                        /* granule index within this launch */
  const int64_t f = f_first + (gidx >> 1);
  const uint32_t gr = (uint32_t)(gidx & 1);
  const p3_frame fr = frames[f];
  const uint32_t nch = fr.nch, sf = fr.sfreq;
  const int64_t o0 = gidx * 2;                            /* gc index of channel 0 within this launch */

  if (threadIdx.x < 2 * P3_SCF_STRIDE) sscf[threadIdx.x / P3_SCF_STRIDE][threadIdx.x % P3_SCF_STRIDE] = scf[o0 * P3_SCF_STRIDE + threadIdx.x];
  if (threadIdx.x < 2) {
    /* effective count1: a zero-length part leaves the reference's count1 stale (Q6, pdmp3.c:2057-2061);
     * the parser stored how many frames back the last non-empty part of this [gr][ch] slot lies */
    const uint32_t ch = threadIdx.x;
    const p3_gc g = gcs[4 * f + 2 * gr + ch];
    int32_t c = 0;
    if (ch < nch) {
      uint32_t back = g.w3;
      if (back == 0) c = count1[o0 + ch];
      else if ((int64_t)back <= f - f_first) c = count1[o0 + ch - 4 * (int64_t)back];
      else c = st_in->count1[gr][ch];                    /* older than this launch: carried state */
    }
    sc1[ch] = c;
    if (f == f_end - 1) st_out->count1[gr][ch] = c;
  }
  __syncthreads();

  for (uint32_t ch = 0; ch < nch; ch++) {
    const p3_gc g = gcs[4 * f + 2 * gr + ch];
    const bool is_short = P3_GC_WINSW(g) && P3_GC_BTYPE(g) == 2;
    const uint32_t first_short = is_short ? (P3_GC_MIXED(g) ? 36u : 0u) : 576u;
    const uint32_t mult = P3_GC_SCALE(g) ? 2u : 1u, pre = P3_GC_PREF(g);
    const int gg = (int)P3_GC_GAIN(g) - 210;
    const int16_t *isp = is_in + (o0 + ch) * 576;
    for (uint32_t d = threadIdx.x; d < 576; d += K2_THREADS) {
      float r;
      if (d >= first_short) {                             /* short-window line, taken from its bitstream position */
        const uint32_t s = T->reorder_src[sf][d];
        const uint32_t sw = T->line_sfbw_s[sf][s], sfb = sw & 15u, win = sw >> 4;
        const uint32_t sc = sfb < 12 ? sscf[ch][P3_SCF_S_OFF + 3 * sfb + win] : 0u;     /* pseudo band 12: ISO 0 (Q5) */
        r = k2_requant(T, isp[s], mult * sc, gg - 8 * (int)P3_GC_SBG(g, win));
      } else {
        const uint32_t sfb = T->line_sfb_l[sf][d];
        const uint32_t sc = sfb < 21 ? sscf[ch][sfb] + pre * T->pretab[sfb] : 0u;       /* pseudo band 21: ISO 0 (Q5) */
        r = k2_requant(T, isp[d], mult * sc, gg);
      }
      xs[ch][d] = r;
    }
  }
  __syncthreads();

  /* ---- stereo (pdmp3.c:1916-1971) ---- */
  if (nch == 2 && fr.mode == 1 && fr.mode_ext != 0) {
    const p3_gc g0 = gcs[4 * f + 2 * gr];
    const uint32_t c0 = (uint32_t)sc1[0], c1r = (uint32_t)sc1[1];
    /* ISO mode (P3_FRAME_ISO): MS up to max(count1); intensity positions from the RIGHT channel's scalefactors, position 7
     * (or above) = not intensity coded, such a band gets MS if that is on; short blocks multiply by the ratios too. */
    const bool iso = (fr.flags & P3_FRAME_ISO) != 0;
    const uint32_t isc = iso ? 1u : 0u;
    const uint32_t msn = (fr.mode_ext & 2) ? (iso ? (c0 > c1r ? c0 : c1r) : (c0 > c1r ? c1r : c0)) : 0u;   /* reference: min(count1), sic (pdmp3.c:1920) */
    const bool is_on = fr.mode_ext & 1;
    const bool sh0 = P3_GC_WINSW(g0) && P3_GC_BTYPE(g0) == 2;
    const uint32_t first_short0 = sh0 ? (P3_GC_MIXED(g0) ? 36u : 0u) : 576u;
    for (uint32_t i = threadIdx.x; i < 576; i += K2_THREADS) {
      float l = xs[0][i], r = xs[1][i];
      bool is_done = false;
      if (is_on && (iso || i >= msn)) {
        if (i >= first_short0) {
          /* short-block intensity (pdmp3.c:2190-2220) in REORDERED position: band sfb occupies
           * [3*s[sfb], 3*s[sfb+1]) and window `win` the win-th third of it (2201-2202) */
          const uint32_t sw = T->line_sfbw_s[sf][i], sfb = sw & 15u, win = sw >> 4;
          if (sfb < 12 && 3u * T->sfb_s[sf][sfb] >= c1r) {
            const uint32_t p = sscf[isc][P3_SCF_S_OFF + 3 * sfb + win];
            if (iso) { if (p < 7) { const float x = l; l = __fmul_rn(T->is_l[p], x); r = __fmul_rn(T->is_r[p], x); is_done = true; } }
            else if (p != 7) {
              /* Q4: assignment through an `unsigned` (pdmp3.c:2191,2212-2213) */
              float x = (float)(unsigned)(long long)l;
              l = x; r = x;
            }
          }
        } else {
          const uint32_t sfb = T->line_sfb_l[sf][i];
          const uint32_t lim = sh0 ? 8u : 21u;                                      /* mixed: long sfb 0..7 only (pdmp3.c:1944) */
          if (sfb < lim && T->sfb_l[sf][sfb] >= c1r) {
            const uint32_t p = sscf[isc][sfb];                                      /* reference: channel-0 scalefactor, sic (pdmp3.c:2163) */
            if (iso ? p < 7 : p != 7) { float x = l; l = __fmul_rn(T->is_l[p & 7], x); r = __fmul_rn(T->is_r[p & 7], x); is_done = true; }
          }
        }
      }
      if (i < msn && !is_done) {
        /* float sum times a double constant, rounded once to float (pdmp3.c:168,1923-1926) */
        float a = __fadd_rn(l, r), b = __fsub_rn(l, r);
        l = __double2float_rn(__dmul_rn((double)a, 0.70710678118654752440));
        r = __double2float_rn(__dmul_rn((double)b, 0.70710678118654752440));
      }
      xs[0][i] = l; xs[1][i] = r;
    }
    __syncthreads();
  }

  /* ---- antialias (pdmp3.c:1706-1732) ---- */
  for (uint32_t ch = 0; ch < nch; ch++) {
    const p3_gc g = gcs[4 * f + 2 * gr + ch];
    const bool sh = P3_GC_WINSW(g) && P3_GC_BTYPE(g) == 2;
    const uint32_t sblim = sh ? (P3_GC_MIXED(g) ? 2u : 1u) : 32u;
    for (uint32_t t = threadIdx.x; t < 31 * 8; t += K2_THREADS) {
      const uint32_t sb = 1 + (t >> 3), i = t & 7;
      if (sb < sblim) {
        const uint32_t li = 18 * sb - 1 - i, ui = 18 * sb + i;
        const float a = xs[ch][li], b = xs[ch][ui], cs = T->cs[i], ca = T->ca[i];
        xs[ch][li] = __fsub_rn(__fmul_rn(a, cs), __fmul_rn(b, ca));
        xs[ch][ui] = __fadd_rn(__fmul_rn(b, cs), __fmul_rn(a, ca));
      }
    }
  }
  __syncthreads();
  for (uint32_t ch = 0; ch < nch; ch++)
    for (uint32_t i = threadIdx.x; i < 576; i += K2_THREADS) xr_out[(o0 + ch) * 576 + i] = xs[ch][i];
}

/* =============================================================================================
 * K3: IMDCT + window + overlap-add + frequency inversion.  One CTA per granule-channel, one
 * thread per output sample (time index i = warp, subband sb = lane, so stores are coalesced in
 * the slot-major layout [i][sb] that the polyphase kernel consumes).
 * Overlap-add without a serial dependency: y[i] = firsthalf(g)[i] + secondhalf(g-1)[i]; the second
 * term is recomputed from the previous granule's spectrum (same 36 MACs per sample as forming all
 * 36 outputs of one granule), or taken from the carried state for the first granule of a launch.
 * ============================================================================================= */
__device__ __forceinline__ float k3_imdct_sample(const p3_tables *T, const float *in /*18 lines of one subband*/,
                                                 uint32_t bt, uint32_t p /*0..35*/)
{
  if (bt == 2) {                                          /* three 12-point transforms (pdmp3.c:1673-1686) */
    float acc = 0.0f;
    #pragma unroll
    for (int w = 0; w < 3; w++) {
      const int q = (int)p - 6 * w - 6;
      if (q >= 0 && q < 12) {
        float sum = 0.0f;
        #pragma unroll
        for (int m = 0; m < 6; m++) sum = __fadd_rn(sum, __fmul_rn(in[w + 3 * m], T->cos12[m][q]));
        acc = __fadd_rn(acc, __fmul_rn(sum, T->imdct_win[2][q]));
      }
    }
    return acc;
  }
  float sum = 0.0f;                                       /* 36-point (pdmp3.c:1689-1698) */
  #pragma unroll
  for (int m = 0; m < 18; m++) sum = __fadd_rn(sum, __fmul_rn(in[m], T->cos36[m][p]));
  return __fmul_rn(sum, T->imdct_win[bt][p]);
}

extern "C" __global__ void __launch_bounds__(K3_THREADS)
k_imdct(const p3_frame *__restrict__ frames, const p3_gc *__restrict__ gcs, const p3_tables *__restrict__ T,
        int64_t f_first, int64_t f_end, const float *__restrict__ xr, const p3_state *__restrict__ st_in,
        p3_state *__restrict__ st_out, float *__restrict__ y_out)
{
  __shared__ float cur[576], prv[576];
  const int64_t o = blockIdx.x;                           /* granule-channel within the launch */
  const int64_t gidx = o >> 1;                            /* granule within the launch */
  const uint32_t ch = (uint32_t)(o & 1);
  const int64_t f = f_first + (gidx >> 1);
  const uint32_t gr = (uint32_t)(gidx & 1);
  if (ch >= frames[f].nch) return;
  const p3_gc g = gcs[4 * f + 2 * gr + ch];
  const bool have_prev = gidx > 0;
  p3_gc gp = g;
  if (have_prev) { const int64_t gp_idx = gidx - 1; gp = gcs[4 * (f_first + (gp_idx >> 1)) + 2 * (gp_idx & 1) + ch]; }

  for (uint32_t i = threadIdx.x; i < 576; i += K3_THREADS) {
    cur[i] = xr[o * 576 + i];
    prv[i] = have_prev ? xr[(o - 2) * 576 + i] : st_in->store[ch][i];
  }
  __syncthreads();

  const uint32_t sb = threadIdx.x & 31, i = threadIdx.x >> 5;           /* i in 0..17 */
  const uint32_t bt = (P3_GC_WINSW(g) && P3_GC_MIXED(g) && sb < 2) ? 0u : P3_GC_BTYPE(g);      /* pdmp3.c:1769-1771 */
  const float a = k3_imdct_sample(T, cur + 18 * sb, bt, i);
  float b;
  if (have_prev) {
    const uint32_t btp = (P3_GC_WINSW(gp) && P3_GC_MIXED(gp) && sb < 2) ? 0u : P3_GC_BTYPE(gp);
    b = k3_imdct_sample(T, prv + 18 * sb, btp, i + 18);
  } else b = prv[18 * sb + i];
  float yv = __fadd_rn(a, b);                                            /* rawout[i] + store[i] (pdmp3.c:1775) */
  if ((sb & 1) && (i & 1)) yv = -yv;                                     /* frequency inversion (pdmp3.c:1741-1743) */
  y_out[o * 576 + i * 32 + sb] = yv;

  /* last granule of the launch: leave its second half behind as the next launch's overlap */
  if (gidx == 2 * (f_end - f_first) - 1)
    st_out->store[ch][18 * sb + i] = k3_imdct_sample(T, cur + 18 * sb, bt, i + 18);
}

/* =============================================================================================
 * K4: polyphase synthesis + PCM.  One CTA per K4_GRAN granules (both channels).
 * Phase A matrixes every slot of the chunk (plus nothing else: the 15 history slots are matrixed
 * from the previous granule's samples, or come from the carried state at the start of a launch):
 *   V(t)[i] = sum_j N[i][j] * S(t)[j]                               (pdmp3.c:2010-2014)
 * Phase B windows and sums without ever shifting a FIFO:
 *   pcm(t)[j] = sum_{k<16} D[32k+j] * V(t-k)[(k odd ? 32 : 0) + j]  (pdmp3.c:2015-2026)
 * ============================================================================================= */
extern "C" __global__ void __launch_bounds__(K4_THREADS)
k_polyphase(const p3_frame *__restrict__ frames, const p3_tables *__restrict__ T, int64_t f_first, int64_t f_end,
            const float *__restrict__ y, const p3_state *__restrict__ st_in, p3_state *__restrict__ st_out,
            int16_t *__restrict__ pcm)
{
  extern __shared__ float smf[];
  float *Nt = smf;                                        /* [32][64] transposed matrixing matrix: Nt[j][i] */
  float *D = Nt + 2048;                                   /* [512] */
  float *S = D + 512;                                     /* [2][15+K4_SLOTS][32] subband samples per slot */
  float *V = S + 2 * (15 + K4_SLOTS) * 32;                /* [2][15+K4_SLOTS][64] */
  const int NS = 15 + K4_SLOTS;

  const int64_t ngran = 2 * (f_end - f_first);
  const int64_t g0 = (int64_t)blockIdx.x * K4_GRAN;
  const int64_t g1 = min(g0 + (int64_t)K4_GRAN, ngran);
  const int nslots = (int)(g1 - g0) * 18;
  const uint32_t nch = frames[f_first].nch;

  for (int i = threadIdx.x; i < 2048; i += K4_THREADS) Nt[(i & 31) * 64 + (i >> 5)] = T->synth_n[i >> 5][i & 31];
  for (int i = threadIdx.x; i < 512; i += K4_THREADS) D[i] = T->synth_d[i];
  /* subband samples: chunk slots at rows 15.., history rows 0..14 from the previous granule */
  for (uint32_t ch = 0; ch < nch; ch++) {
    for (int e = threadIdx.x; e < (nslots + 15) * 32; e += K4_THREADS) {
      const int row = e >> 5, j = e & 31;                 /* row 0..14 history (slot -15..-1), 15.. chunk */
      const int64_t t = g0 * 18 + row - 15;               /* slot index within the launch */
      float v = 0.0f;
      if (t >= 0) { const int64_t gg = t / 18; const int ss = (int)(t % 18); v = y[(gg * 2 + ch) * 576 + ss * 32 + j]; }
      S[(ch * NS + row) * 32 + j] = v;
    }
  }
  __syncthreads();
  /* phase A */
  for (uint32_t ch = 0; ch < nch; ch++) {
    for (int e = threadIdx.x; e < (nslots + 15) * 64; e += K4_THREADS) {
      const int row = e >> 6, i = e & 63;
      const int64_t t = g0 * 18 + row - 15;
      float sum;
      if (t >= 0) {
        const float *s = S + (ch * NS + row) * 32;
        sum = 0.0f;
        #pragma unroll
        for (int j = 0; j < 32; j++) sum = __fadd_rn(sum, __fmul_rn(Nt[j * 64 + i], s[j]));
      } else sum = st_in->vhist[ch][(int)(-t) - 1][i];    /* before the launch: carried history */
      V[(ch * NS + row) * 64 + i] = sum;
    }
  }
  __syncthreads();
  /* phase B */
  for (int e = threadIdx.x; e < nslots * 32; e += K4_THREADS) {
    const int r = e >> 5, j = e & 31;                     /* chunk slot r, output sample j */
    const int64_t gg = g0 + r / 18;                       /* granule within the launch */
    const p3_frame &fr = frames[f_first + (gg >> 1)];
    int32_t out[2] = {0, 0};
    for (uint32_t ch = 0; ch < nch; ch++) {
      const float *v = V + (ch * NS + r + 15) * 64;
      float sum = 0.0f;
      #pragma unroll
      for (int k = 0; k < 16; k++) {
        const float u = __fmul_rn(v[-64 * k + ((k & 1) ? 32 : 0) + j], D[32 * k + j]);
        sum = __fadd_rn(sum, u);
      }
      const double d = __dmul_rn((double)sum, 32767.0);   /* double multiply, truncation (pdmp3.c:2028) */
      int32_t s = (d > -2147483649.0 && d < 2147483648.0) ? __double2int_rz(d) : (int32_t)0x80000000;   /* x86 cvttsd2si */
      s = s > 32767 ? 32767 : (s < -32767 ? -32767 : s);  /* clamp is +-32767 (pdmp3.c:2029-2030) */
      out[ch] = s;
    }
    if (fr.flags & P3_FRAME_DECODE) {
      const int64_t sample = (int64_t)fr.pcm_index * 1152 + (gg & 1) * 576 + (r % 18) * 32 + j;
      if (nch == 2) reinterpret_cast<uint32_t *>(pcm)[sample] = (uint32_t)(out[0] & 0xffff) | ((uint32_t)out[1] << 16);
      else pcm[sample] = (int16_t)out[0];
    }
  }
  /* the last chunk leaves the matrixed vectors of its last 15 slots as the next launch's history */
  if (g1 == ngran) {
    for (uint32_t ch = 0; ch < nch; ch++)
      for (int e = threadIdx.x; e < 15 * 64; e += K4_THREADS) {
        const int age = e >> 6, i = e & 63;               /* age+1 slots back from the end */
        st_out->vhist[ch][age][i] = V[(ch * NS + 15 + nslots - 1 - age) * 64 + i];
      }
  }
}

/* =============================================================================================
 * k_sideinfo -- Read_Audio_L3 (pdmp3.c:1129-1200) on the device: one thread per frame turns the 17/32 bytes of
 * side info in front of the frame's main data into the four p3_gc descriptors, scfsi and the validity flag
 * (SURVEY Q10), exactly what parse_side() of p3_parse.c produces on the host.  Every field sits at a fixed bit
 * position (a granule-channel is 59 bits whether or not window switching is on), so after the bytes are
 * realigned to big-endian words the extraction is compile-time shifts.  any_empty: set when a part has length 0
 * (the Q6 chain of k_q6_chain is needed).
 * ============================================================================================= */
__device__ __forceinline__ uint32_t si_get(const uint32_t (&w)[9], int pos, int n)   /* n <= 16, MSB first */
{
  const int i = pos >> 5, o = pos & 31;
  const uint32_t v = o ? __funnelshift_l(w[i + 1], w[i], o) : w[i];
  return v >> (32 - n);
}

/* bits of part 2 of a granule-channel, as part2_bits() of p3_parse.c (the empty-part rule of pdmp3.c:2057-2061) */
__device__ __forceinline__ uint32_t si_part2_bits(uint32_t sfc, uint32_t ws, uint32_t bt, uint32_t mixed, uint32_t gr, uint32_t scfsi4)
{
  /* slen1 / slen2 (pdmp3.c:530-533), 3 bits per entry */
  const uint32_t s1 = (uint32_t)((0x91b69224b000ull >> (3 * sfc)) & 7ull), s2 = (uint32_t)((0x69a2d1688688ull >> (3 * sfc)) & 7ull);
  if (ws && bt == 2) return mixed ? 17 * s1 + 18 * s2 : 18 * s1 + 18 * s2;
  if (gr == 0) scfsi4 = 0;
  return ((scfsi4 & 1) ? 0 : 6 * s1) + ((scfsi4 & 2) ? 0 : 5 * s1) + ((scfsi4 & 4) ? 0 : 5 * s2) + ((scfsi4 & 8) ? 0 : 5 * s2);
}

template <int NCH>
__device__ __forceinline__ void si_parse(const uint32_t (&w)[9], p3_frame &fr, p3_gc (&gc)[4], int &any_empty)
{
  constexpr int P0 = 9 + (NCH == 1 ? 5 : 3);
  fr.scfsi = (uint8_t)__brev(si_get(w, P0, 4 * NCH) << (32 - 4 * NCH));   /* first bit read = band 0 of ch 0 = bit 0 */
  uint32_t start = 0, bad = 0;
  #pragma unroll
  for (int k = 0; k < 4; k++) { gc[k].w0 = gc[k].w1 = gc[k].w2 = gc[k].w3 = 0; }
  #pragma unroll
  for (int gr = 0; gr < 2; gr++)
    #pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
      const int B = P0 + 4 * NCH + 59 * (gr * NCH + ch);
      const uint32_t p23l = si_get(w, B, 12), bigv = si_get(w, B + 12, 9), gain = si_get(w, B + 21, 8), sfc = si_get(w, B + 29, 4);
      const uint32_t ws = si_get(w, B + 33, 1);
      uint32_t bt = 0, mixed = 0, t0, t1, t2 = 0, s0 = 0, s1 = 0, s2 = 0, r0, r1;
      if (ws) {
        bt = si_get(w, B + 34, 2); mixed = si_get(w, B + 36, 1);
        t0 = si_get(w, B + 37, 5); t1 = si_get(w, B + 42, 5);
        s0 = si_get(w, B + 47, 3); s1 = si_get(w, B + 50, 3); s2 = si_get(w, B + 53, 3);
        r0 = (bt == 2 && !mixed) ? 8 : 7; r1 = 20 - r0;                       /* implicit (pdmp3.c:1181-1185) */
        if (bt == 0) bad = 1;
      } else {
        t0 = si_get(w, B + 34, 5); t1 = si_get(w, B + 39, 5); t2 = si_get(w, B + 44, 5);
        r0 = si_get(w, B + 49, 4); r1 = si_get(w, B + 53, 3);
        if (r0 + r1 + 2 > 22) bad = 1;
      }
      const uint32_t pre = si_get(w, B + 56, 1), scale = si_get(w, B + 57, 1), c1t = si_get(w, B + 58, 1);
      if (bigv > 288) bad = 1;
      p3_gc &g = gc[gr * 2 + ch];
      g.w0 = p23l | bigv << 12 | gain << 21 | pre << 29 | scale << 30 | c1t << 31;
      g.w1 = sfc | ws << 4 | bt << 5 | mixed << 7 | t0 << 8 | t1 << 13 | t2 << 18 | r0 << 23 | r1 << 27;
      g.w2 = s0 | s1 << 3 | s2 << 6 | (start & 0x3fffu) << 9;
      start += p23l ? p23l : ((fr.flags & P3_FRAME_ISO) ? 0u : si_part2_bits(sfc, ws, bt, mixed, gr, ((uint32_t)fr.scfsi >> (4 * ch)) & 15u));
      if (p23l == 0) any_empty = 1;
    }
  if (start > 8u * ((uint32_t)fr.main_begin + fr.main_size)) bad = 1;          /* parts overrun the frame's data */
  if (bad) fr.flags |= P3_FRAME_BAD;
}

extern "C" __global__ void __launch_bounds__(128)
k_sideinfo(const uint8_t *__restrict__ raw, p3_frame *__restrict__ frames, p3_gc *__restrict__ gcs, int64_t n_frames, int *__restrict__ any_empty)
{
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  p3_frame fr = frames[f];
  const uint32_t silen = fr.nch == 1 ? 17u : 32u;
  const uint8_t *si = raw + fr.main_off - silen;
  const uint32_t *al = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(si) & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(si) & 3) * 8;
  uint32_t l[10], w[9];
  #pragma unroll
  for (int k = 0; k < 10; k++) l[k] = __ldg(al + k);                            /* the staged stream carries 64 bytes of slack */
  #pragma unroll
  for (int k = 0; k < 9; k++) w[k] = __byte_perm(__funnelshift_r(l[k], l[k + 1], sh), 0, 0x0123);   /* stream order, first byte in the MSB */
  p3_gc gc[4]; int empty = 0;
  if (fr.nch == 1) si_parse<1>(w, fr, gc, empty); else si_parse<2>(w, fr, gc, empty);
  uint4 *out = reinterpret_cast<uint4 *>(gcs + 4 * f);
  #pragma unroll
  for (int k = 0; k < 4; k++) out[k] = make_uint4(gc[k].w0, gc[k].w1, gc[k].w2, 0u);
  frames[f].scfsi = fr.scfsi; frames[f].flags = fr.flags;
  if (empty) *any_empty = 1;
}

/* Q6 (pdmp3.c:2057-2061): a zero-length part leaves count1 of its [gr][ch] slot stale; w3 of such a part = how many
 * frames back the slot was last written (0x7fffffff: not in this batch -> the carried state).  Runs only when
 * k_sideinfo saw an empty part: one CTA, every thread scans a contiguous chunk of frames twice. */
extern "C" __global__ void __launch_bounds__(1024)
k_q6_chain(const p3_frame *__restrict__ frames, p3_gc *__restrict__ gcs, int64_t n_frames, const int *__restrict__ any_empty)
{
  if (!*any_empty || (frames[0].flags & P3_FRAME_ISO)) return;      /* ISO mode: an empty part has count1 = 0, nothing to chain */
  __shared__ int64_t s_last[1024][4];
  const int t = threadIdx.x;
  const int64_t per = (n_frames + 1023) / 1024, lo = (int64_t)t * per, hi = min(lo + per, n_frames);
  auto written = [&](int64_t f, int k) { return P3_GC_P23L(gcs[4 * f + k]) != 0 || (frames[f].flags & (P3_FRAME_NODATA | P3_FRAME_BAD)); };
  int64_t last[4] = {-1, -1, -1, -1};
  for (int64_t f = lo; f < hi; f++) for (int k = 0; k < 4; k++) if ((k & 1) < frames[f].nch && written(f, k)) last[k] = f;
  for (int k = 0; k < 4; k++) s_last[t][k] = last[k];
  __syncthreads();
  if (t < 4) { int64_t run = -1; for (int i = 0; i < 1024; i++) { const int64_t v = s_last[i][t]; s_last[i][t] = run; if (v >= 0) run = v; } }   /* exclusive scan */
  __syncthreads();
  for (int k = 0; k < 4; k++) last[k] = s_last[t][k];
  for (int64_t f = lo; f < hi; f++) for (int k = 0; k < 4; k++) {
    if ((k & 1) >= frames[f].nch) continue;
    if (written(f, k)) { last[k] = f; gcs[4 * f + k].w3 = 0; }
    else gcs[4 * f + k].w3 = last[k] >= 0 ? (uint32_t)(f - last[k]) : 0x7fffffffu;
  }
}
