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
#include "p3_lee.inc"

struct p3_fconst {
  float dct4[18][18];              /* cos(pi/18 (k+1/2)(m+1/2)) */
  float win[4][36];
  float cos12[6][12];
  float cs[8], ca[8], is_l[8], is_r[8];
  float t1h[40];
  float t2[320];
  float pretab[24];
};
__constant__ p3_fconst FC;

extern "C" int p3_fused_upload_consts(const p3_tables *T, const float *dct4)
{
  static p3_fconst h;
  memcpy(h.dct4, dct4, sizeof h.dct4);
  memcpy(h.win, T->imdct_win, sizeof h.win); memcpy(h.cos12, T->cos12, sizeof h.cos12);
  memcpy(h.cs, T->cs, 32); memcpy(h.ca, T->ca, 32); memcpy(h.is_l, T->is_l, 32); memcpy(h.is_r, T->is_r, 32);
  memcpy(h.t1h, T->t1h, sizeof h.t1h); memcpy(h.t2, T->t2, sizeof h.t2);
  for (int i = 0; i < 24; i++) h.pretab[i] = (float)T->pretab[i];
  return (int)cudaMemcpyToSymbol(FC, &h, sizeof h);
}

/* ---- 32-point DCT-II, Lee's recursion, fully unrolled in registers ---------------------------- */
template <int N> struct Lee { static __device__ __forceinline__ const float *k(); };
#define LEE_TAB(N) template <> struct Lee<N> { static __device__ __forceinline__ float c(int i) { constexpr float t[N / 2] = P3_LEE##N; return t[i]; } };
LEE_TAB(32) LEE_TAB(16) LEE_TAB(8) LEE_TAB(4) LEE_TAB(2)

template <int N> __device__ __forceinline__ void dct2(float (&x)[N])
{
  if constexpr (N == 1) return;
  else {
    float a[N / 2], b[N / 2];
    #pragma unroll
    for (int i = 0; i < N / 2; i++) { a[i] = x[i] + x[N - 1 - i]; b[i] = (x[i] - x[N - 1 - i]) * Lee<N>::c(i); }
    dct2<N / 2>(a); dct2<N / 2>(b);
    #pragma unroll
    for (int i = 0; i < N / 2; i++) { x[2 * i] = a[i]; x[2 * i + 1] = (i + 1 < N / 2) ? b[i] + b[i + 1] : b[i]; }
  }
}


/* ---- 18-point DCT-IV via two 9-point DCT-IIs (tools/proto/fast_transforms.py) -------------------
 * y[m] = x[m] * 2cos(pi(2m+1)/72);  Y = DCT-II-18(y) by one Lee split into two DCT-II-9;
 * t[0] = Y[0]/2, t[k] = Y[k] - t[k-1].   ~125 flops instead of 324. */
__device__ __forceinline__ void dct9(const float (&x)[9], float (&X)[9])
{
  constexpr float C10 = 9.848077530e-01f, C20 = 9.396926208e-01f, C30 = 8.660254038e-01f, C40 = 7.660444431e-01f, C50 = 6.427876097e-01f, C70 = 3.420201433e-01f, C80 = 1.736481777e-01f;
  const float s0 = x[0] + x[8], s1 = x[1] + x[7], s2 = x[2] + x[6], s3 = x[3] + x[5], x4 = x[4];
  const float d0 = x[0] - x[8], d1 = x[1] - x[7], d2 = x[2] - x[6], d3 = x[3] - x[5];
  const float h1 = 0.5f * s1;
  X[0] = (s0 + s1) + (s2 + s3) + x4;
  X[2] = fmaf(s0, C20, fmaf(-s2, C80, fmaf(-s3, C40, h1 - x4)));
  X[4] = fmaf(s0, C40, fmaf(-s2, C20, fmaf(s3, C80, x4 - h1)));
  X[6] = fmaf(s0 + s2 + s3, 0.5f, -(s1 + x4));
  X[8] = fmaf(s0, C80, fmaf(s2, C40, fmaf(-s3, C20, x4 - h1)));
  const float e1 = d1 * C30;
  X[1] = fmaf(d0, C10, fmaf(d2, C50, fmaf(d3, C70, e1)));
  X[3] = (d0 - d2 - d3) * C30;
  X[5] = fmaf(d0, C50, fmaf(-d2, C70, fmaf(d3, C10, -e1)));
  X[7] = fmaf(d0, C70, fmaf(d2, C10, fmaf(-d3, C50, -e1)));
}

__device__ __forceinline__ void dct4_18(const float (&x)[18], float (&t)[18])
{
  constexpr float PRE[18] = {1.998096443e+00f, 1.982889723e+00f, 1.952592014e+00f, 1.907433901e+00f, 1.847759065e+00f, 1.774021666e+00f, 1.686782892e+00f, 1.586706681e+00f, 1.474554674e+00f, 1.351180415e+00f, 1.217522858e+00f, 1.074599217e+00f, 9.234972265e-01f, 7.653668647e-01f, 6.014115990e-01f, 4.328792279e-01f, 2.610523844e-01f, 8.723877473e-02f};
  constexpr float LEE[9] = {5.019099188e-01f, 5.176380902e-01f, 5.516889595e-01f, 6.103872944e-01f, 7.071067812e-01f, 8.717233978e-01f, 1.183100792e+00f, 1.931851653e+00f, 5.736856623e+00f};
  float a[9], b[9], A[9], B[9];
  #pragma unroll
  for (int m = 0; m < 9; m++) {
    const float u = x[m] * PRE[m], v = x[17 - m] * PRE[17 - m];
    a[m] = u + v; b[m] = (u - v) * LEE[m];
  }
  dct9(a, A); dct9(b, B);
  t[0] = 0.5f * A[0];
  #pragma unroll
  for (int k = 1; k < 18; k++) {
    const float Y = (k & 1) ? (k == 17 ? B[8] : B[k >> 1] + B[(k >> 1) + 1]) : A[k >> 1];
    t[k] = Y - t[k - 1];
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
  __shared__ __align__(16) uint8_t s_scf[4][P3_SCF_STRIDE];
  __shared__ gcpar s_par[4];
  __shared__ float s_scale[4][40];                       /* fl(t1*t2) per scalefactor band: long sfb 0..21, short 3*sfb+win */
  __shared__ uint32_t s_sfreq;

  const int tid = threadIdx.x;
  const int64_t c0 = f_first + (int64_t)blockIdx.x * frames_per_cta;
  const int64_t c1 = min(c0 + (int64_t)frames_per_cta, f_end);
  const int warm = blockIdx.x > 0 ? 1 : 0;
  const uint32_t nch = frames[c0].nch;

  /* window coefficients of this thread's output column j, signs of the V<->X symmetry folded in */
  const int j = tid & 31, wgr = (tid >> 5) & 1, wch = tid >> 6;
  float ce[8], co[8]; int ia, ib;
  {
    /* even k: V[j]:   j<16 -> +X[16+j];  j==16 -> 0;  j>16 -> -X[48-j]
       odd  k: V[32+j]: j<16 -> -X[16-j]; j>=16 -> -X[j-16]                       (see tools/proto/fast_transforms.py) */
    float se = j < 16 ? 1.0f : (j == 16 ? 0.0f : -1.0f);
    ia = j < 16 ? 16 + j : (j == 16 ? 0 : 48 - j);
    ib = j < 16 ? 16 - j : j - 16;
    #pragma unroll
    for (int k = 0; k < 8; k++) { ce[k] = se * T->synth_d[64 * k + j] * 32767.0f; co[k] = -T->synth_d[64 * k + 32 + j] * 32767.0f; }
  }

  /* carried state */
  for (int i = tid; i < 2 * 576; i += FT) {
    (&tails[2][0][0])[i] = warm ? 0.0f : st_in->store[i / 576][i % 576];
    (&tails[0][0][0])[i] = 0.0f; (&tails[1][0][0])[i] = 0.0f;
  }
  for (int i = tid; i < 2 * XSLOTS * XPITCH; i += FT) (&xring[0][0][0])[i] = 0.0f;
  __syncthreads();
  if (!warm)
    for (int i = tid; i < 2 * 15 * 32; i += FT) { int ch = i / 480, r = i % 480, s = r / 32, k = r % 32; xring[ch][36 + s][k] = st_in->xhist[ch][14 - s][k]; }
  if (tid == 0) s_sfreq = 0xffffffffu;

  /* software pipeline: the spectra of frame n+1 are fetched into registers while frame n is processed */
  uint32_t pre[9]; uint32_t pre_scf = 0;
  const uint32_t *isw = reinterpret_cast<const uint32_t *>(is_in);
  {
    const int64_t o0 = (c0 - warm - f_first) * 4;
    #pragma unroll
    for (int k = 0; k < 9; k++) pre[k] = __ldg(isw + o0 * 288 + tid + FT * k);
    if (tid < 64) pre_scf = __ldg(reinterpret_cast<const uint32_t *>(scf + o0 * P3_SCF_STRIDE) + tid);
  }
  __syncthreads();

  int n = 0;                                              /* frame iteration within this CTA */
  for (int64_t f = c0 - warm; f < c1; f++, n++) {
    const p3_frame fr = frames[f];
    const int64_t o0 = (f - f_first) * 4;
    const bool emit = !(warm && n == 0) && (fr.flags & P3_FRAME_DECODE);
    if (s_sfreq != fr.sfreq) {                            /* per-line helper tables of this sample rate */
      __syncthreads();
      for (int i = tid; i < 576; i += FT) { s_sfb_l[i] = T->line_sfb_l[fr.sfreq][i]; s_sfbw_s[i] = T->line_sfbw_s[fr.sfreq][i]; s_reo[i] = T->reorder_src[fr.sfreq][i]; }
      if (tid == 0) s_sfreq = fr.sfreq;
    }
    /* the last 15 slots of the previous frame become the history of this one (stage F of that frame must be done) */
    __syncthreads();
    for (int i = tid; i < 2 * 15 * 32; i += FT) { int ch = i / 480, r = i % 480, sl = r >> 5, k = r & 31; xring[ch][sl][k] = xring[ch][36 + sl][k]; }
    /* land the prefetched spectra / scalefactors in shared memory, start the next fetch */
    {
      uint32_t *ib32 = reinterpret_cast<uint32_t *>(&isbuf[0][0]);
      #pragma unroll
      for (int k = 0; k < 9; k++) ib32[tid + FT * k] = pre[k];
      if (tid < 64) reinterpret_cast<uint32_t *>(&s_scf[0][0])[tid] = pre_scf;
      if (f + 1 < c1) {
        #pragma unroll
        for (int k = 0; k < 9; k++) pre[k] = __ldg(isw + (o0 + 4) * 288 + tid + FT * k);
        if (tid < 64) pre_scf = __ldg(reinterpret_cast<const uint32_t *>(scf + (o0 + 4) * P3_SCF_STRIDE) + tid);
      }
    }
    if (tid < 4) {                                        /* unpack the side info of this granule-channel once */
      const uint32_t gr = tid >> 1, ch = tid & 1;
      const p3_gc g = gcs[4 * f + tid];
      gcpar p;
      int32_t c = 0;
      if (ch < nch) {                                     /* effective count1 (Q6), as in k_requant */
        const uint32_t back = g.w3;
        if (back == 0) c = count1[o0 + tid];
        else if ((int64_t)back <= f - f_first) c = count1[o0 + tid - 4 * (int64_t)back];
        else c = st_in->count1[gr][ch];
      }
      if (f == f_end - 1) st_out->count1[gr][ch] = c;
      const bool is_short = P3_GC_WINSW(g) && P3_GC_BTYPE(g) == 2;
      p.c1 = c; p.gg = (int)P3_GC_GAIN(g) - 210;
      p.first_short = is_short ? (P3_GC_MIXED(g) ? 36 : 0) : 576;
      p.mult = P3_GC_SCALE(g) ? 2 : 1; p.pre = P3_GC_PREF(g); p.bt = P3_GC_BTYPE(g); p.mixed = P3_GC_MIXED(g); p.ws = P3_GC_WINSW(g);
      p.live = ch < nch;
      p.sbg8[0] = 8 * P3_GC_SBG(g, 0); p.sbg8[1] = 8 * P3_GC_SBG(g, 1); p.sbg8[2] = 8 * P3_GC_SBG(g, 2);
      p.sblim = is_short ? (P3_GC_MIXED(g) ? 2 : 1) : 32;
      s_par[tid] = p;
    }
    __syncthreads();
    /* band scales fl(t1*t2): 2^(-(scale?1:0.5)*(scalefac+preflag*pretab)) * 2^((gain-210-8*sbg)/4) (pdmp3.c:2127-2128,2144-2146).
     * long blocks: index = sfb (0..21); short: 3*sfb+win; mixed: long bands 0..7 sit in the unused short slots 0..7 */
    for (int e = tid; e < 4 * 40; e += FT) {
      const int gcl = e / 40, b = e % 40;
      const gcpar p = s_par[gcl];
      float v = 0.0f;
      if (p.live) {
        const bool longband = p.first_short == 576 ? b < 22 : (p.first_short == 36 && b < 8);
        if (longband) {
          const uint32_t sc = b < 21 ? s_scf[gcl][b] + p.pre * (uint32_t)FC.pretab[b] : 0u;
          v = __fmul_rn(FC.t1h[p.mult * sc], FC.t2[p.gg + P3_T2_BIAS]);
        } else if (p.first_short < 576 && b < 39) {
          const int sfb = b / 3, win = b % 3;
          const uint32_t sc = sfb < 12 ? s_scf[gcl][P3_SCF_S_OFF + 3 * sfb + win] : 0u;
          v = __fmul_rn(FC.t1h[p.mult * sc], FC.t2[p.gg - (int)p.sbg8[win] + P3_T2_BIAS]);
        }
      }
      s_scale[gcl][b] = v;
    }
    __syncthreads();

    /* ---- A: requantize + reorder (exact arithmetic of pdmp3.c:2121-2152) ---- */
    #pragma unroll 1
    for (int gcl = 0; gcl < 4; gcl++) {
      const gcpar p = s_par[gcl];
      if (!p.live) { for (int d = tid; d < 576; d += FT) xs[gcl][d] = 0.0f; continue; }
      const int16_t *isp = isbuf[gcl];
      const float *scl = s_scale[gcl];
      #pragma unroll
      for (int it = 0; it < 5; it++) {
        const int d = tid + FT * it;
        if (d < 576) {
          float r;
          if (d >= p.first_short) {
            const uint32_t s = s_reo[d], sw = s_sfbw_s[s];
            r = fq_requant(T->pow43, isp[s], scl[3 * (sw & 15u) + (sw >> 4)]);
          } else r = fq_requant(T->pow43, isp[d], scl[s_sfb_l[d]]);
          xs[gcl][d] = r;
        }
      }
    }
    __syncthreads();

    /* ---- B: stereo (pdmp3.c:1916-1971), both granules ---- */
    if (nch == 2 && fr.mode == 1 && fr.mode_ext != 0) {
      #pragma unroll 1
      for (int gr = 0; gr < 2; gr++) {
        const uint32_t cl = (uint32_t)s_par[2 * gr].c1, c1r = (uint32_t)s_par[2 * gr + 1].c1;
        const uint32_t msn = (fr.mode_ext & 2) ? (cl > c1r ? c1r : cl) : 0u;
        const uint32_t first_short0 = s_par[2 * gr].first_short;
        const bool sh0 = first_short0 < 576;
        const bool is_on = fr.mode_ext & 1;
        #pragma unroll
        for (int it = 0; it < 5; it++) {
          const uint32_t i = tid + FT * it;
          if (i >= 576) break;
          float l = xs[2 * gr][i], r = xs[2 * gr + 1][i];
          if (i < msn) {
            float a = __fadd_rn(l, r), b = __fsub_rn(l, r);
            l = __double2float_rn(__dmul_rn((double)a, 0.70710678118654752440));
            r = __double2float_rn(__dmul_rn((double)b, 0.70710678118654752440));
            xs[2 * gr][i] = l; xs[2 * gr + 1][i] = r;
          } else if (is_on) {
            if (i >= first_short0) {
              const uint32_t sw = s_sfbw_s[i], sfb = sw & 15u, win = sw >> 4;
              if (sfb < 12 && 3u * T->sfb_s[fr.sfreq][sfb] >= c1r && s_scf[2 * gr][P3_SCF_S_OFF + 3 * sfb + win] != 7) {
                float x = (float)(unsigned)(long long)l;                               /* Q4 */
                xs[2 * gr][i] = x; xs[2 * gr + 1][i] = x;
              }
            } else {
              const uint32_t sfb = s_sfb_l[i], lim = sh0 ? 8u : 21u;
              if (sfb < lim && T->sfb_l[fr.sfreq][sfb] >= c1r) {
                const uint32_t pp = s_scf[2 * gr][sfb];
                if (pp != 7) { xs[2 * gr][i] = __fmul_rn(FC.is_l[pp & 7], l); xs[2 * gr + 1][i] = __fmul_rn(FC.is_r[pp & 7], l); }
              }
            }
          }
        }
      }
      __syncthreads();
    }

    /* ---- C: antialias (pdmp3.c:1706-1732): 31 boundaries x 8 butterflies per granule-channel ---- */
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
    if (xr_tap) for (int e = tid; e < 4 * 576; e += FT) xr_tap[o0 * 576 + e] = (&xs[0][0])[e];

    /* ---- D: IMDCT + window; first half in place, second half to the tail buffer ---- */
    {
      const uint32_t gcl = tid >> 5, sb = tid & 31, gr = gcl >> 1, ch = gcl & 1;
      const gcpar p = s_par[gcl];
      if (p.live) {
        const uint32_t bt = (p.ws && p.mixed && sb < 2) ? 0u : p.bt;
        float *x = &xs[gcl][18 * sb];
        float *tl = &tails[gr == 0 ? 0 : 1 + (n & 1)][ch][18 * sb];
        float in[18];
        #pragma unroll
        for (int m = 0; m < 18; m++) in[m] = x[m];
        if (bt != 2) {
          float t[18];
          dct4_18(in, t);
          #pragma unroll
          for (int k = 0; k < 9; k++) {                                    /* 36-point IMDCT from the DCT-IV by symmetry */
            tl[8 - k] = -t[k] * FC.win[bt][26 - k]; tl[9 + k] = -t[k] * FC.win[bt][27 + k];
            x[k] = t[9 + k] * FC.win[bt][k]; x[17 - k] = -t[9 + k] * FC.win[bt][17 - k];
          }
        } else {
          /* three 12-point transforms (pdmp3.c:1673-1686): raw[6w+6+p] += win2[p] * sum_m in[w+3m] cos12[m][p] */
          float raw[36];
          #pragma unroll
          for (int q = 0; q < 36; q++) raw[q] = 0.0f;
          #pragma unroll
          for (int w = 0; w < 3; w++)
            #pragma unroll
            for (int q = 0; q < 12; q++) {
              float sum = 0.0f;
              #pragma unroll
              for (int m = 0; m < 6; m++) sum = fmaf(in[w + 3 * m], FC.cos12[m][q], sum);
              raw[6 * w + 6 + q] += sum * FC.win[2][q];
            }
          #pragma unroll
          for (int i = 0; i < 18; i++) { x[i] = raw[i]; tl[i] = raw[18 + i]; }
        }
      }
    }
    __syncthreads();

    /* ---- E: overlap-add + frequency inversion + 32-point DCT per time slot -> X ring ---- */
    if (tid < 72) {
      const uint32_t gr = tid / 36, r = tid % 36, ch = r / 18, ss = r % 18;
      if (ch < nch) {
        const float *cur = &xs[2 * gr + ch][0];
        const float *prv = gr == 0 ? &tails[2 - (n & 1)][ch][0] : &tails[0][ch][0];
        float s[32];
        #pragma unroll
        for (int sb = 0; sb < 32; sb++) {
          float v = cur[18 * sb + ss] + prv[18 * sb + ss];                 /* rawout + store (pdmp3.c:1775) */
          s[sb] = ((sb & 1) && (ss & 1)) ? -v : v;                         /* frequency inversion (1741-1743) */
        }
        if (y_tap) {
          #pragma unroll
          for (int sb = 0; sb < 32; sb++) y_tap[(o0 + 2 * gr + ch) * 576 + ss * 32 + sb] = s[sb];
        }
        dct2<32>(s);
        float *X = &xring[ch][15 + gr * 18 + ss][0];
        #pragma unroll
        for (int k = 0; k < 32; k++) X[k] = s[k];
      }
    }
    __syncthreads();

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
          pcm[base + (int64_t)(9 * h + s9) * 32 * nch] = (int16_t)s;
        }
      }
    }
  }
  __syncthreads();
  if (c1 == f_end) {                                       /* leave the state for the next launch */
    const int last = n - 1;
    for (int i = tid; i < 2 * 576; i += FT) st_out->store[i / 576][i % 576] = tails[1 + (last & 1)][i / 576][i % 576];
    (void)last;
    for (int i = tid; i < 2 * 15 * 32; i += FT) { int ch = i / 480, r = i % 480, age = r / 32, k = r % 32; st_out->xhist[ch][age][k] = xring[ch][50 - age][k]; }
  }
}
