/* p3_fused.cu -- FAST mode: requantize + reorder + stereo + antialias + IMDCT + polyphase + PCM in
 * ONE kernel (the whole of Decode_L3, pdmp3.c:1024-1060, plus Convert_Frame_S16, 2307-2345).
 *
 * Only the Huffman output (int16 spectra) is read and only int16 PCM is written; everything in
 * between lives in shared memory and registers.  One CTA walks a run of consecutive frames so the
 * two pieces of state the reference keeps in static arrays -- the IMDCT overlap (store, pdmp3.c:1755)
 * and the 16-slot polyphase FIFO (v_vec, pdmp3.c:1983) -- stay on chip; a run is primed by decoding
 * the frame in front of it without emitting PCM (halo of one frame).
 *
 * Transforms are evaluated with fast algorithms instead of the reference's O(N^2) loops:
 *   - 36-point IMDCT = 18-point DCT-IV plus sign/index symmetry (4x fewer MACs than pdmp3.c:1689-1698)
 *   - 64x32 matrixing = 32-point DCT-II (Lee) plus symmetry: V[i] is +-X[k] or 0, so only X[0..31] is kept
 *     (pdmp3.c:2010-2014 needs 2048 MACs per slot, this needs ~290 flops)
 *   - the 512-tap window (pdmp3.c:2015-2026) runs from registers with a sliding 33-slot history.
 * Sums are therefore taken in a different order than the reference: PCM differs by at most 1 LSB
 * (tests/test_gpu_fast.py); the stages up to antialias keep the reference's exact arithmetic.
 */
#include "p3_device.cuh"
#include "p3_kernels.h"
#include "p3_xform.cuh"

struct p3_fconst {
  float dct4[18][18];              /* cos(pi/18 (k+1/2)(m+1/2)) */
  float win[4][36];
  float swin[4][36];               /* win with the IMDCT symmetry signs folded in: -win for outputs 9..35 */
  float cos12[6][12];
  float cs[8], ca[8], is_l[8], is_r[8];
  float t1h[40];
  float t2[320];
  float pretab[24];
};
__constant__ p3_fconst FC;
static int p3_synthw_check_consts(const float *cs, const float *ca);   /* p3_synthw.cuh: its immediates against the table */

extern "C" int p3_fused_upload_consts(const p3_tables *T, const float *dct4)
{
  static p3_fconst h;
  memcpy(h.dct4, dct4, sizeof h.dct4);
  memcpy(h.win, T->imdct_win, sizeof h.win);
  for (int b = 0; b < 4; b++) for (int i = 0; i < 36; i++) h.swin[b][i] = i < 9 ? T->imdct_win[b][i] : -T->imdct_win[b][i];
  memcpy(h.cos12, T->cos12, sizeof h.cos12);
  memcpy(h.cs, T->cs, 32); memcpy(h.ca, T->ca, 32); memcpy(h.is_l, T->is_l, 32); memcpy(h.is_r, T->is_r, 32);
  memcpy(h.t1h, T->t1h, sizeof h.t1h); memcpy(h.t2, T->t2, sizeof h.t2);
  for (int i = 0; i < 24; i++) h.pretab[i] = (float)T->pretab[i];
  { int rc = p3_synthw_check_consts(h.cs, h.ca); if (rc) return rc; }
  return (int)cudaMemcpyToSymbol(FC, &h, sizeof h);
}

/* fast transforms (32-point DCT-II, 18-point DCT-IV), shared with the packed-stereo kernel: p3_xform.cuh */

/* three 12-point IMDCTs of a short block (pdmp3.c:1673-1686): raw[6w+6+p] += win2[p] * sum_m in[w+3m] cos12[m][p];
 * V = float (one channel) or f2 (both channels packed: the coefficients are the same for both) */
template <class V> __device__ __forceinline__ void imdct_short(const V (&in)[18], V (&raw)[36])
{
  #pragma unroll
  for (int q = 0; q < 36; q++) raw[q] = vzero(in[0]);
  #pragma unroll
  for (int w = 0; w < 3; w++)
    #pragma unroll
    for (int q = 0; q < 12; q++) {
      V sum = vzero(in[0]);
      #pragma unroll
      for (int m = 0; m < 6; m++) sum = vfma(in[w + 3 * m], FC.cos12[m][q], sum);
      raw[6 * w + 6 + q] = vfma(sum, FC.win[2][q], raw[6 * w + 6 + q]);
    }
}

#define FT 128                     /* threads per CTA */
#define XPITCH 33
#define XSLOTS 51                  /* 15 history slots + the 36 slots of one frame, linear */

/* per granule-channel parameters, unpacked once per frame by threads 0..3 */
struct gcpar {
  int32_t  c1;                     /* effective count1 (Q6) */
  int32_t  gg;                     /* global_gain - 210 */
  uint16_t first_short;            /* first line that uses short windows (576: none) */
  uint8_t  mult, pre, bt, mixed, ws, live;
  uint8_t  sbg8[3];                /* 8 * subblock_gain */
  uint8_t  sblim;                  /* antialias subband limit */
};

/* |is|^(4/3) with the sign, times the band scale: (t1*t2)*t3 with both products rounded (pdmp3.c:2132,2150) */
__device__ __forceinline__ float fq_requant(const float *__restrict__ pow43, int v, float scale)
{
  float t3 = __ldg(pow43 + (v < 0 ? -v : v));
  if (v < 0) t3 = -t3;
  return __fmul_rn(scale, t3);
}

/* shared-memory working set of the synthesis stages (one frame = 4 granule-channels at a time).
 * Side info, scalefactors and band scales are double buffered (slot = frame parity) so that a frame's
 * tables can be prepared while the previous frame is still in its transforms. */
struct synth_sm {
  float (*xs)[576];                  /* [4] spectra -> subband samples, in place */
  float (*tails)[2][576];            /* [3] IMDCT second halves: [0] granule 0, [1]/[2] granule 1 of odd/even frames */
  float (*xring)[XSLOTS][XPITCH];    /* [2] 32-point DCT of 15 history + 36 new time slots per channel */
  int16_t (*isbuf)[576];             /* [4] Huffman output of the frame about to be requantized */
  uint8_t *sfb_l, *sfbw_s; uint16_t *reo;   /* per-line helper tables of the batch's sample rate */
  gcpar (*par)[4]; float (*scale)[4][40]; uint8_t (*scf)[4][P3_SCF_STRIDE];   /* [2] slots */
  float (*t01)[2][576]; float (*t2)[576];   /* fused kernel: tails[0..1] in phase-local memory, tails[2] persistent */
  __device__ __forceinline__ float *tail(int idx, int ch) const
  {
    if (tails) return &tails[idx][ch][0];
    return idx == 2 ? &t2[ch][0] : &t01[idx][ch][0];
  }
};

#define SF_LOCALS \
  const int tid = threadIdx.x; (void)tid; \
  float (*xs)[576] = S.xs; float (*xring)[XSLOTS][XPITCH] = S.xring; int16_t (*isbuf)[576] = S.isbuf; \
  uint8_t *s_sfb_l = S.sfb_l, *s_sfbw_s = S.sfbw_s; uint16_t *s_reo = S.reo; gcpar *s_par = S.par[slot]; \
  float (*s_scale)[40] = S.scale[slot]; const uint8_t (*scf4)[P3_SCF_STRIDE] = S.scf[slot]; \
  (void)xs; (void)xring; (void)isbuf; (void)s_sfb_l; (void)s_sfbw_s; (void)s_reo; (void)s_par; (void)s_scale; (void)scf4;

/* per-line helper tables of sample rate `sf` (once per launch: a batch never mixes sample rates) */
__device__ __forceinline__ void sf_load_luts(const synth_sm &S, const p3_tables *__restrict__ T, uint32_t sf)
{
  for (int i = threadIdx.x; i < 576; i += FT) { S.sfb_l[i] = T->line_sfb_l[sf][i]; S.sfbw_s[i] = T->line_sfbw_s[sf][i]; S.reo[i] = T->reorder_src[sf][i]; }
}

/* Land one frame's inputs in shared memory slot `slot`: spectra (from the registers they were prefetched into),
 * scalefactor words, and the unpacked side info of the 4 granule-channels (threads 0..3; c1eff = effective count1, Q6). */
__device__ __forceinline__ void sf_land(const synth_sm &S, int slot, const uint32_t (&pre)[9], uint32_t scfword, const p3_gc &g, int32_t c1eff, uint32_t nch)
{
  const int tid = threadIdx.x;
  uint32_t *ib32 = reinterpret_cast<uint32_t *>(&S.isbuf[0][0]);
  #pragma unroll
  for (int k = 0; k < 9; k++) ib32[tid + FT * k] = pre[k];
  if (tid < 64) reinterpret_cast<uint32_t *>(&S.scf[slot][0][0])[tid] = scfword;
  if (tid < 4) {
    const uint32_t ch = tid & 1;
    gcpar p;
    const bool is_short = P3_GC_WINSW(g) && P3_GC_BTYPE(g) == 2;
    p.c1 = c1eff; p.gg = (int)P3_GC_GAIN(g) - 210;
    p.first_short = is_short ? (P3_GC_MIXED(g) ? 36 : 0) : 576;
    p.mult = P3_GC_SCALE(g) ? 2 : 1; p.pre = P3_GC_PREF(g); p.bt = P3_GC_BTYPE(g); p.mixed = P3_GC_MIXED(g); p.ws = P3_GC_WINSW(g);
    p.live = ch < nch;
    p.sbg8[0] = 8 * P3_GC_SBG(g, 0); p.sbg8[1] = 8 * P3_GC_SBG(g, 1); p.sbg8[2] = 8 * P3_GC_SBG(g, 2);
    p.sblim = is_short ? (P3_GC_MIXED(g) ? 2 : 1) : 32;
    S.par[slot][tid] = p;
  }
}

__device__ __forceinline__ void sf_scale(const synth_sm &S, int slot)
{
  SF_LOCALS
    /* band scales fl(t1*t2): 2^(-(scale?1:0.5)*(scalefac+preflag*pretab)) * 2^((gain-210-8*sbg)/4) (pdmp3.c:2127-2128,2144-2146).
     * long blocks: index = sfb (0..21); short: 3*sfb+win; mixed: long bands 0..7 sit in the unused short slots 0..7 */
    for (int e = tid; e < 4 * 40; e += FT) {
      const int gcl = e / 40, b = e % 40;
      const gcpar p = s_par[gcl];
      float v = 0.0f;
      if (p.live) {
        const bool longband = p.first_short == 576 ? b < 22 : (p.first_short == 36 && b < 8);
        if (longband) {
          const uint32_t sc = b < 21 ? scf4[gcl][b] + p.pre * (uint32_t)FC.pretab[b] : 0u;
          v = __fmul_rn(FC.t1h[p.mult * sc], FC.t2[p.gg + P3_T2_BIAS]);
        } else if (p.first_short < 576 && b < 39) {
          const int sfb = b / 3, win = b % 3;
          const uint32_t sc = sfb < 12 ? scf4[gcl][P3_SCF_S_OFF + 3 * sfb + win] : 0u;
          v = __fmul_rn(FC.t1h[p.mult * sc], FC.t2[p.gg - (int)p.sbg8[win] + P3_T2_BIAS]);
        }
      }
      s_scale[gcl][b] = v;
    }
}

template <int DUMMY>
__device__ __forceinline__ void sf_stageAB(const synth_sm &S, int slot, const p3_frame &fr, const p3_tables *__restrict__ T, uint32_t nch)
{
  SF_LOCALS
    /* ---- A+B: requantize + reorder (exact arithmetic of pdmp3.c:2121-2152) and stereo (pdmp3.c:1916-1971):
     *      a thread requantizes line d of BOTH channels of a granule and couples them in registers ---- */
    {
      const bool st_on = nch == 2 && fr.mode == 1 && fr.mode_ext != 0;
      const bool is_on = st_on && (fr.mode_ext & 1);
      #pragma unroll 1
      for (int gr = 0; gr < 2; gr++) {
        const gcpar p0 = s_par[2 * gr], p1 = s_par[2 * gr + 1];
        const int16_t *is0 = isbuf[2 * gr], *is1 = isbuf[2 * gr + 1];
        const float *sc0 = s_scale[2 * gr], *sc1 = s_scale[2 * gr + 1];
        const uint32_t cl = (uint32_t)p0.c1, c1r = (uint32_t)p1.c1;
        const bool iso = (fr.flags & P3_FRAME_ISO) != 0;        /* ISO semantics, see k_requant */
        const uint32_t isc = iso ? 1u : 0u;
        const uint32_t msn = (st_on && (fr.mode_ext & 2)) ? (iso ? (cl > c1r ? cl : c1r) : (cl > c1r ? c1r : cl)) : 0u;     /* reference: min(count1), sic (pdmp3.c:1920) */
        const uint32_t first_short0 = p0.first_short;
        const bool sh0 = first_short0 < 576;
        #pragma unroll
        for (int it = 0; it < 5; it++) {
          const uint32_t d = tid + FT * it;
          if (d >= 576) break;
          float l = 0.0f, r = 0.0f;
          if (p0.live) {
            if (d >= p0.first_short) { const uint32_t s = s_reo[d], sw = s_sfbw_s[s]; l = fq_requant(T->pow43, is0[s], sc0[3 * (sw & 15u) + (sw >> 4)]); }
            else l = fq_requant(T->pow43, is0[d], sc0[s_sfb_l[d]]);
          }
          if (p1.live) {
            if (d >= p1.first_short) { const uint32_t s = s_reo[d], sw = s_sfbw_s[s]; r = fq_requant(T->pow43, is1[s], sc1[3 * (sw & 15u) + (sw >> 4)]); }
            else r = fq_requant(T->pow43, is1[d], sc1[s_sfb_l[d]]);
          }
          bool is_done = false;
          if (is_on && (iso || d >= msn)) {
            if (d >= first_short0) {
              /* short-block intensity (pdmp3.c:2190-2220) in reordered position; Q4: assignment through an `unsigned` */
              const uint32_t sw = s_sfbw_s[d], sfb = sw & 15u, win = sw >> 4;
              if (sfb < 12 && 3u * T->sfb_s[fr.sfreq][sfb] >= c1r) {
                const uint32_t pp = scf4[2 * gr + isc][P3_SCF_S_OFF + 3 * sfb + win];
                if (iso) { if (pp < 7) { const float x = l; l = __fmul_rn(FC.is_l[pp], x); r = __fmul_rn(FC.is_r[pp], x); is_done = true; } }
                else if (pp != 7) { const float x = (float)(unsigned)(long long)l; l = x; r = x; }
              }
            } else {
              const uint32_t sfb = s_sfb_l[d], lim = sh0 ? 8u : 21u;                     /* mixed: long sfb 0..7 only (pdmp3.c:1944) */
              if (sfb < lim && T->sfb_l[fr.sfreq][sfb] >= c1r) {
                const uint32_t pp = scf4[2 * gr + isc][sfb];                             /* reference: channel-0 scalefactor, sic (pdmp3.c:2163) */
                if (iso ? pp < 7 : pp != 7) { const float x = l; l = __fmul_rn(FC.is_l[pp & 7], x); r = __fmul_rn(FC.is_r[pp & 7], x); is_done = true; }
              }
            }
          }
          if (d < msn && !is_done) {
            /* float sum times a double constant, rounded once to float (pdmp3.c:168,1923-1926) */
            const float a = __fadd_rn(l, r), b = __fsub_rn(l, r);
            l = __double2float_rn(__dmul_rn((double)a, 0.70710678118654752440));
            r = __double2float_rn(__dmul_rn((double)b, 0.70710678118654752440));
          }
          xs[2 * gr][d] = l; xs[2 * gr + 1][d] = r;
        }
      }
    }
}

/* the last 15 slots of the previous frame become the history of the next one (stage F of that frame must be over) */
__device__ __forceinline__ void sf_hist(const synth_sm &S)
{
  const int lane = threadIdx.x & 31;
  #pragma unroll
  for (int it = 0; it < 8; it++) {                          /* 30 rows of 32 floats over 4 warps */
    const int r = (threadIdx.x >> 5) + 4 * it;
    if (r < 30) { const int ch = r >= 15, sl = r - 15 * ch; S.xring[ch][sl][lane] = S.xring[ch][36 + sl][lane]; }
  }
}

__device__ __forceinline__ void sf_tapC(const synth_sm &S, int slot, float *xr_tap, bool store)
{
  SF_LOCALS
    /* ---- C: antialias (pdmp3.c:1706-1732).  Normally folded into stage D (below); as a separate pass only
     *      when the post-antialias spectra are tapped for the stage-level parity tests ---- */
    if (xr_tap) {
      #pragma unroll 1
      for (int gcl = 0; gcl < 4; gcl++) {
        const uint32_t sblim = s_par[gcl].live ? s_par[gcl].sblim : 0;
        #pragma unroll
        for (int it = 0; it < 2; it++) {
          const uint32_t t = tid + FT * it, sb = 1 + (t >> 3), i = t & 7;
          if (sb < sblim) {
            const uint32_t li = 18 * sb - 1 - i, ui = 18 * sb + i;
            const float a = xs[gcl][li], b = xs[gcl][ui];
            xs[gcl][li] = __fsub_rn(__fmul_rn(a, FC.cs[i]), __fmul_rn(b, FC.ca[i]));
            xs[gcl][ui] = __fadd_rn(__fmul_rn(b, FC.cs[i]), __fmul_rn(a, FC.ca[i]));
          }
        }
      }
      __syncthreads();
      if (store) for (int e = tid; e < 4 * 576; e += FT) xr_tap[e] = (&xs[0][0])[e];
    }

}

__device__ __forceinline__ void sf_stageD(const synth_sm &S, int slot, int n, const float *xr_tap)
{
  SF_LOCALS
    /* ---- D: antialias + IMDCT + window; first half in place, second half to the tail buffer.
     *      One warp = the 32 subbands of one granule-channel, so the butterflies across subband boundaries
     *      only need the neighbour lanes' lines: read, __syncwarp(), then write in place. ---- */
    {
      const uint32_t gcl = tid >> 5, sb = tid & 31, gr = gcl >> 1, ch = gcl & 1;
      const gcpar p = s_par[gcl];
      if (p.live) {
        const uint32_t bt = (p.ws && p.mixed && sb < 2) ? 0u : p.bt;
        float *x = &xs[gcl][18 * sb];
        float *tl = S.tail(gr == 0 ? 0 : 1 + (n & 1), ch) + 18 * sb;
        float in[18];
        #pragma unroll
        for (int m = 0; m < 18; m++) in[m] = x[m];
        if (!xr_tap) {
          /* boundary below (with subband sb-1) turns lines 0..7, boundary above (with sb+1) lines 17..10 */
          const bool lo = sb >= 1 && sb < p.sblim, hi = sb + 1 < p.sblim;
          #pragma unroll
          for (int i = 0; i < 8; i++) {
            const float below = lo ? x[-1 - i] : 0.0f, above = hi ? x[18 + i] : 0.0f;
            const float u = in[i], l = in[17 - i];
            if (lo) in[i] = __fadd_rn(__fmul_rn(u, FC.cs[i]), __fmul_rn(below, FC.ca[i]));       /* ub (pdmp3.c:1726) */
            if (hi) in[17 - i] = __fsub_rn(__fmul_rn(l, FC.cs[i]), __fmul_rn(above, FC.ca[i]));   /* lb (pdmp3.c:1725) */
          }
        }
        __syncwarp();
        if (bt != 2) {
          float t[18];
          dct4_18(in, t);
          #pragma unroll
          for (int k = 0; k < 9; k++) {                                    /* 36-point IMDCT from the DCT-IV by symmetry */
            tl[8 - k] = -t[k] * FC.win[bt][26 - k]; tl[9 + k] = -t[k] * FC.win[bt][27 + k];
            x[k] = t[9 + k] * FC.win[bt][k]; x[17 - k] = -t[9 + k] * FC.win[bt][17 - k];
          }
        } else {
          float raw[36];
          imdct_short<float>(in, raw);
          #pragma unroll
          for (int i = 0; i < 18; i++) { x[i] = raw[i]; tl[i] = raw[18 + i]; }
        }
      }
    }
}

__device__ __forceinline__ void sf_stageE(const synth_sm &S, int n, uint32_t nch, float *y_tap)
{
  const int slot = 0; SF_LOCALS
    /* ---- E: overlap-add + frequency inversion + 32-point DCT per time slot -> X ring ---- */
    if (tid < 72) {
      const uint32_t gr = tid / 36, r = tid % 36, ch = r / 18, ss = r % 18;
      if (ch < nch) {
        const float *cur = &xs[2 * gr + ch][0];
        const float *prv = gr == 0 ? S.tail(2 - (n & 1), ch) : S.tail(0, ch);
        float s[32];
        #pragma unroll
        for (int sb = 0; sb < 32; sb++) {
          float v = cur[18 * sb + ss] + prv[18 * sb + ss];                 /* rawout + store (pdmp3.c:1775) */
          s[sb] = ((sb & 1) && (ss & 1)) ? -v : v;                         /* frequency inversion (1741-1743) */
        }
        if (y_tap) {
          #pragma unroll
          for (int sb = 0; sb < 32; sb++) y_tap[(2 * gr + ch) * 576 + ss * 32 + sb] = s[sb];
        }
        dct2<32, float>(s);
        float *X = &xring[ch][15 + gr * 18 + ss][0];
        #pragma unroll
        for (int k = 0; k < 32; k++) X[k] = s[k];
      }
    }
}

template <bool SCRATCH>
__device__ __forceinline__ void sf_stageF(const synth_sm &S, const p3_frame &fr, uint32_t nch, bool emit, int16_t *__restrict__ pcm,
                                          const float (&ce)[8], const float (&co)[8], int ia, int ib)
{
  const int slot = 0; SF_LOCALS
  const int j = tid & 31, wgr = (tid >> 5) & 1, wch = tid >> 6;
    /* ---- F: 512-tap window from registers + PCM ---- */
    if (wch < (int)nch && emit) {
      const int64_t base = ((int64_t)fr.pcm_index * 1152 + wgr * 576 + j) * nch + wch;
      #pragma unroll
      for (int h = 0; h < 2; h++) {                                        /* two runs of 9 slots keep the history in 48 registers */
        float A[24], B[24];                                                /* A[i] = X(first-15+9h+i)[ia], B likewise with ib */
        const float *X0 = &xring[wch][wgr * 18 + 9 * h][0];
        #pragma unroll
        for (int i = 0; i < 24; i++) { A[i] = X0[i * XPITCH + ia]; B[i] = X0[i * XPITCH + ib]; }
        #pragma unroll
        for (int s9 = 0; s9 < 9; s9++) {
          float sum = 0.0f;
          #pragma unroll
          for (int k = 0; k < 8; k++) { sum = fmaf(ce[k], A[15 + s9 - 2 * k], sum); sum = fmaf(co[k], B[15 + s9 - 2 * k - 1], sum); }
          /* (int32)(sum*32767.0) with the scale folded into the window; x86 gives INT_MIN out of range */
          int32_t s = fabsf(sum) < 2147483648.0f ? __float2int_rz(sum) : (int32_t)0x80000000;
          s = max(-32767, min(32767, s));
          int16_t *dst = pcm + base + (int64_t)(9 * h + s9) * 32 * nch;
          if (SCRATCH) __stcs(dst, (int16_t)s);                                /* persistent kernel: do not displace the L2-resident spectra */
          else *dst = (int16_t)s;                                              /* the two channels of a sector meet in L2 */
        }
      }
    }
}

/* window coefficients of this thread's output column j, signs of the V<->X symmetry and the 32767 scale folded in:
 *   even k: V[j]:   j<16 -> +X[16+j];  j==16 -> 0;  j>16 -> -X[48-j]
 *   odd  k: V[32+j]: j<16 -> -X[16-j]; j>=16 -> -X[j-16]                       (see tools/proto/fast_transforms.py) */
__device__ __forceinline__ void synth_window_coeffs(const p3_tables *__restrict__ T, float (&ce)[8], float (&co)[8], int &ia, int &ib)
{
  const int j = threadIdx.x & 31;
  const float se = j < 16 ? 1.0f : (j == 16 ? 0.0f : -1.0f);
  ia = j < 16 ? 16 + j : (j == 16 ? 0 : 48 - j);
  ib = j < 16 ? 16 - j : j - 16;
  #pragma unroll
  for (int k = 0; k < 8; k++) { ce[k] = se * T->synth_d[64 * k + j] * 32767.0f; co[k] = -T->synth_d[64 * k + 32 + j] * 32767.0f; }
}

/* carried state -> shared memory (st == NULL: zero state, used when a run is primed by a warm-up frame) */
__device__ __forceinline__ void synth_load_state(const synth_sm &S, const p3_state *__restrict__ st)
{
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * 576; i += FT) {
    S.tail(2, i / 576)[i % 576] = st ? st->store[i / 576][i % 576] : 0.0f;
    S.tail(0, i / 576)[i % 576] = 0.0f; S.tail(1, i / 576)[i % 576] = 0.0f;
  }
  for (int i = tid; i < 2 * XSLOTS * XPITCH; i += FT) (&S.xring[0][0][0])[i] = 0.0f;
  __syncthreads();
  if (st)
    for (int i = tid; i < 2 * 15 * 32; i += FT) { int ch = i / 480, r = i % 480, s = r / 32, k = r % 32; S.xring[ch][36 + s][k] = st->xhist[ch][14 - s][k]; }
  __syncthreads();
}

/* shared memory -> carried state, after the last frame (iteration index `last`) of a launch */
__device__ __forceinline__ void synth_store_state(const synth_sm &S, int last, p3_state *__restrict__ st)
{
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * 576; i += FT) st->store[i / 576][i % 576] = S.tail(1 + (last & 1), i / 576)[i % 576];
  for (int i = tid; i < 2 * 15 * 32; i += FT) { int ch = i / 480, r = i % 480, age = r / 32, k = r % 32; st->xhist[ch][age][k] = S.xring[ch][50 - age][k]; }
}

extern "C" __global__ void __launch_bounds__(FT, 5)
k_synth_fast(const p3_frame *__restrict__ frames, const p3_gc *__restrict__ gcs, const p3_tables *__restrict__ T,
             int64_t f_first, int64_t f_end, int frames_per_cta,
             const int16_t *__restrict__ is_in, const int32_t *__restrict__ count1, const uint8_t *__restrict__ scf,
             const p3_state *__restrict__ st_in, p3_state *__restrict__ st_out, int16_t *__restrict__ pcm,
             float *__restrict__ xr_tap, float *__restrict__ y_tap)
{
  __shared__ float xs[4][576];
  __shared__ float tails[3][2][576];
  __shared__ float xring[2][XSLOTS][XPITCH];
  __shared__ __align__(16) int16_t isbuf[4][576];
  __shared__ uint8_t s_sfb_l[576], s_sfbw_s[576];
  __shared__ uint16_t s_reo[576];
  __shared__ __align__(16) uint8_t s_scf[2][4][P3_SCF_STRIDE];
  __shared__ gcpar s_par[2][4];
  __shared__ float s_scale[2][4][40];                    /* fl(t1*t2) per scalefactor band: long sfb 0..21, short 3*sfb+win */
  synth_sm S = {xs, tails, xring, isbuf, s_sfb_l, s_sfbw_s, s_reo, s_par, s_scale, s_scf, NULL, NULL};

  const int tid = threadIdx.x;
  const int64_t c0 = f_first + (int64_t)blockIdx.x * frames_per_cta;
  const int64_t c1 = min(c0 + (int64_t)frames_per_cta, f_end);
  const int warm = blockIdx.x > 0 ? 1 : 0;
  const uint32_t nch = frames[c0].nch;
  const uint32_t *isw = reinterpret_cast<const uint32_t *>(is_in);

  float ce[8], co[8]; int ia, ib;
  synth_window_coeffs(T, ce, co, ia, ib);
  sf_load_luts(S, T, frames[c0].sfreq);
  synth_load_state(S, warm ? NULL : st_in);

  /* Everything a frame needs from global memory is fetched one frame ahead into registers:
   *   pre[9]   this thread's share of the frame's spectra      pscf  its word of the scalefactors (threads 0..63)
   *   gq       side info of granule-channel `tid` (threads 0..3)    cq  its count1      frq  second half of p3_frame */
  uint32_t pre[9], pscf = 0; p3_gc gq = {0, 0, 0, 0}; int32_t cq = 0; uint4 frq;
  auto fetch = [&](int64_t f) {
    const int64_t o = (f - f_first) * 4;
    #pragma unroll
    for (int k = 0; k < 9; k++) pre[k] = __ldg(isw + o * 288 + tid + FT * k);
    if (tid < 64) pscf = __ldg(reinterpret_cast<const uint32_t *>(scf + o * P3_SCF_STRIDE) + tid);
    if (tid < 4) { gq = gcs[4 * f + tid]; cq = count1[o + tid]; }
    frq = __ldg(reinterpret_cast<const uint4 *>(frames + f) + 1);
  };
  /* effective count1 of granule-channel `tid` of frame f (Q6: an empty part keeps the slot's previous value) */
  auto eff_c1 = [&](int64_t f) -> int32_t {
    int32_t c = 0;
    if (tid < 4 && (uint32_t)(tid & 1) < nch) {
      const uint32_t back = gq.w3;
      if (back == 0) c = cq;
      else if ((int64_t)back <= f - f_first) c = count1[(f - f_first) * 4 + tid - 4 * (int64_t)back];
      else c = st_in->count1[tid >> 1][tid & 1];
    }
    if (tid < 4 && f == f_end - 1) st_out->count1[tid >> 1][tid & 1] = c;
    return c;
  };

  const int64_t fs = c0 - warm;
  fetch(fs);
  sf_land(S, 0, pre, pscf, gq, eff_c1(fs), nch);
  p3_frame fr; *(reinterpret_cast<uint4 *>(&fr) + 1) = frq;
  if (fs + 1 < c1) fetch(fs + 1);
  __syncthreads();
  sf_scale(S, 0);
  __syncthreads();

  /* Three barriers per frame.  While frame n is in its transforms, frame n+1 is landed in the other table
   * slot (after stage A+B has consumed the spectra buffer) and its band scales are computed (before stage E). */
  int n = 0;
  for (int64_t f = fs; f < c1; f++, n++) {
    const int slot = n & 1;
    const bool emit = !(warm && n == 0) && (fr.flags & P3_FRAME_DECODE);
    const int64_t o0 = (f - f_first) * 4;
    sf_stageAB<0>(S, slot, fr, T, nch);
    __syncthreads();
    /* tests only: separate antialias pass + tap (has its own barrier).  The warm-up frame of a run (decoded from zero state,
     * its overlap is not the real one) belongs to the previous CTA: that CTA writes its taps, this one must not. */
    const bool tap_here = !(warm && n == 0);
    if (xr_tap) sf_tapC(S, slot, xr_tap + o0 * 576, tap_here);
    sf_hist(S);
    p3_frame frn = fr;
    if (f + 1 < c1) {
      sf_land(S, slot ^ 1, pre, pscf, gq, eff_c1(f + 1), nch);
      *(reinterpret_cast<uint4 *>(&frn) + 1) = frq;
      if (f + 2 < c1) fetch(f + 2);
    }
    sf_stageD(S, slot, n, xr_tap);
    __syncthreads();
    if (f + 1 < c1) sf_scale(S, slot ^ 1);
    sf_stageE(S, n, nch, (y_tap && tap_here) ? y_tap + o0 * 576 : NULL);
    __syncthreads();
    sf_stageF<false>(S, fr, nch, emit, pcm, ce, co, ia, ib);
    fr = frn;
  }
  __syncthreads();
  if (c1 == f_end) synth_store_state(S, n - 1, st_out);
}

#include "p3_synthw.cuh"
