/* p3_synthw.cuh -- k_synth_warp: FAST-mode synthesis of STEREO streams, one autonomous warp per run of frames.
 * (included at the end of p3_fused.cu; shares its constant tables and scalar helpers)
 *
 * Same algorithm and operation order as k_synth_fast (requantize .. PCM of Decode_L3, pdmp3.c:1024-1060, plus
 * Convert_Frame_S16, 2307-2345); requantize / intensity stereo / antialias keep the reference's exact arithmetic, MS
 * stereo reproduces its double-precision product in packed fp32 (identical but for ~1e-7 of the values, by 1 ulp), the
 * transforms agree with k_synth_fast to rounding (ptxas contracts some packed mul+add pairs into FFMA2), the PCM
 * is within 1 LSB of the reference like that kernel's.  Organised around the warp:
 *
 *   - both channels of a granule travel together as one packed fp32 pair (f2) and every transform runs on
 *     Blackwell's packed FFMA2/FADD2/FMUL2: half the issue slots of the one-channel-per-thread kernel;
 *   - lane = subband for requantize / stereo / antialias / IMDCT, so the spectrum goes from the Huffman
 *     output to the IMDCT output in REGISTERS: the antialias butterflies reach the neighbour subbands with
 *     warp shuffles, the IMDCT overlap (the reference's static `store`, pdmp3.c:1755) never leaves the lane's
 *     registers; lane = time slot for the 32-point DCT and lane = output column for the 512-tap window;
 *   - the only shared-memory traffic is the [slot][subband] transpose in front of the DCT and the DCT outputs
 *     (a 36-slot ring: the 16-slot FIFO of the reference, pdmp3.c:1983) that the window stage slides over;
 *   - a warp never waits for another warp: no __syncthreads() in the frame loop, only __syncwarp();
 *   - the Huffman output of the next granules (2304 B of spectra + 128 B of scalefactors) is brought in by the
 *     TMA engine (cp.async.bulk, completion on an mbarrier) while the current granule is transformed, no registers involved.
 */
#ifndef SW_WPB
#define SW_WPB    4                 /* warps per CTA (each one independent) */
#endif
#ifndef SW_NBUF
#define SW_NBUF   1                 /* spectra buffers per warp: 1 = the next granule is fetched while this one is transformed (measured: long-block
                                       stream 5.34 -> 5.27 ms, mixed-block VBR 8.02 -> 7.78 ms against 2 buffers), 2 = fetched two granules ahead */
#endif
#ifndef SW_MINB
#define SW_MINB   3                 /* CTAs per SM the register allocation aims at */
#endif
#ifndef SW_MS_EXACT
#define SW_MS_EXACT 0               /* 1: MS stereo reproduces the reference's float-sum x DOUBLE-constant product bit for bit (4 packed instructions per line);
                                       0: one packed multiply by the nearest float to 1/sqrt 2 -- at most 1 ulp away on a fifth of the values, far below the
                                       1-LSB budget of FAST mode; saves 54 of the 842 packed instructions per granule (k_synth_fast and EXACT mode stay bit-exact) */
#endif
#define SW_PITCH  33                /* f2 elements per DCT row: 66 words -> conflict-free for row-per-lane and column-per-lane access */
#define SW_LUT_BYTES (1728 + 1440 + 656)  /* CTA-shared: reorder_src u16[576] | line_sfbw_s u8[576] | t1h f32[40] | t2 f32[320] | line_sfb_l u8[576] | sfb_l u16[24] | sfb_s u16[16] */

struct __align__(16) sw_warp_sm {
  uint8_t isb[SW_NBUF][2304];             /* [buffer][ch][576] int16: Huffman output of one granule (TMA destination) */
  uint8_t scf[SW_NBUF][128];              /* [buffer][ch][64]: its scalefactors (TMA destination) */
  f2 xr[36][SW_PITCH];              /* two blocks of 18 DCT rows; the block of the granule about to be transformed doubles as scratch */
  float scale[40][2];               /* band scales fl(t1*t2) of the current granule, [band][ch] */
  struct { uint4 fr; uint4 gc[4]; int32_t c1[4]; } desc[2];   /* [frame parity] second half of p3_frame, the 4 p3_gc, their count1 (TMA destination) */
  unsigned long long mbar[2], dbar[2];
};

__device__ __forceinline__ uint32_t sw_s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sw_mbar_init(unsigned long long *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(sw_s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sw_mbar_expect(unsigned long long *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sw_s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sw_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(sw_s32(dst)), "l"(src), "r"(bytes), "r"(sw_s32(bar)) : "memory");
}
__device__ __forceinline__ void sw_mbar_wait(unsigned long long *bar, uint32_t parity)
{
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(sw_s32(bar)), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();             /* a lost copy must not hang the device */
  } while (!ok);
}

/* per-channel parameters of a granule, identical in every lane (plain scalars: nothing here is ever indexed) */
struct sw_par {
  int32_t c1, gg; uint32_t first_short, mult, pre, bt, mixed, ws, sblim, sbg;   /* sbg: 8*subblock_gain[w] in byte w */
};
__device__ __forceinline__ sw_par sw_unpack(const uint4 g, int32_t c1eff)
{
  p3_gc q; q.w0 = g.x; q.w1 = g.y; q.w2 = g.z; q.w3 = g.w;
  sw_par p;
  const bool is_short = P3_GC_WINSW(q) && P3_GC_BTYPE(q) == 2;
  p.c1 = c1eff; p.gg = (int)P3_GC_GAIN(q) - 210;
  p.first_short = is_short ? (P3_GC_MIXED(q) ? 36u : 0u) : 576u;
  p.mult = P3_GC_SCALE(q) ? 2u : 1u; p.pre = P3_GC_PREF(q); p.bt = P3_GC_BTYPE(q); p.mixed = P3_GC_MIXED(q); p.ws = P3_GC_WINSW(q);
  p.sbg = (8 * P3_GC_SBG(q, 0)) | (8 * P3_GC_SBG(q, 1)) << 8 | (8 * P3_GC_SBG(q, 2)) << 16;
  p.sblim = is_short ? (P3_GC_MIXED(q) ? 2u : 1u) : 32u;
  return p;
}

/* int16 of (int32)(sum * 32767.0), scale folded into the window, clamped to +-32767 (pdmp3.c:2028-2030): a saturating
 * truncation to s16 and one packed max for both channels.  (Where the reference's own float->int32 conversion
 * overflows -- |sample| >= 2^31, undefined in C -- this saturates; k_synth_fast keeps the x86 result there.) */
__device__ __forceinline__ uint32_t sw_pcm2(f2 sum)
{
  int16_t l, r;
  asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(l) : "f"(f2_x(sum)));
  asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(r) : "f"(f2_y(sum)));
  const uint32_t v = (uint32_t)(uint16_t)l | ((uint32_t)(uint16_t)r << 16);
  return __vmaxs2(v, 0x80018001u);
}

/* One side of an antialias butterfly, in place and under a predicate: x = x*cs + nb*ca with the product nb*ca rounded
 * and the other one kept exact inside the fma (what ptxas makes of the separately written products anyway).  The
 * coefficients are immediates (cs/ca of pdmp3.c:573-574; p3_fused_upload_consts checks them against the table). */
#define SW_CS_LIST {0.857493f, 0.881742f, 0.949629f, 0.983315f, 0.995518f, 0.999161f, 0.999899f, 0.999993f}
#define SW_CA_LIST {-0.514496f, -0.471732f, -0.313377f, -0.181913f, -0.094574f, -0.040966f, -0.014199f, -0.003700f}
__device__ __forceinline__ float sw_cs(int i) { constexpr float t[8] = SW_CS_LIST; return t[i]; }
__device__ __forceinline__ float sw_ca(int i) { constexpr float t[8] = SW_CA_LIST; return t[i]; }
/* x = v under a predicate */
__device__ __forceinline__ void sw_mov_if(f2 &x, f2 v, bool on)
{
  asm("{\n .reg .pred p;\n setp.ne.s32 p, %2, 0;\n @p mov.b64 %0, %1;\n}" : "+l"(x.v) : "l"(v.v), "r"((int)on));
}
/* x = c - a under a predicate, as fma(a, -1, c) */
__device__ __forceinline__ void sw_fnma_if(f2 &x, f2 a, f2 c, bool on)
{
  asm("{\n .reg .pred p;\n .reg .b64 m;\n setp.ne.s32 p, %3, 0;\n mov.b64 m, {0fBF800000, 0fBF800000};\n @p fma.rn.f32x2 %0, %1, m, %2;\n}"
      : "+l"(x.v) : "l"(a.v), "l"(c.v), "r"((int)on));
}
__device__ __forceinline__ void sw_aa(f2 &x, f2 nb, float cs, float ca, bool on)
{
  asm("{\n .reg .pred p;\n .reg .b64 t, c;\n setp.ne.s32 p, %2, 0;\n mov.b64 c, {%4, %4};\n mul.rn.f32x2 t, %1, c;\n mov.b64 c, {%3, %3};\n @p fma.rn.f32x2 %0, %0, c, t;\n}"
      : "+l"(x.v) : "l"(nb.v), "r"((int)on), "f"(cs), "f"(ca));
}

/* The kernel body, instantiated six times.
 * ISO = the batch is flagged P3_FRAME_ISO (a batch is uniform in that: the parser flags every frame or none), so the default
 *       kernels carry no code for the switch.
 * CLS = content class of the CTA's frames.  Every launch sequence runs the kernels of all three classes over the same grid; a CTA
 *       first classifies the frames of its runs from the descriptors (one frame per lane, one __syncthreads_and per class) and
 *       returns at once unless they are of its kernel's class:
 *         2 LEAN  every granule has the same win_switch / block_type / mixed flag in both channels, no short block, and no frame
 *                 has the intensity bit set -- the usual case by far; the body has no test, branch, register or instruction-
 *                 cache line of the rare paths in its way: 6.0 -> 5.3 ms per 10^6 frames of the long-block benchmark stream;
 *         1 SAME  the two channels of every granule agree in those flags (short / mixed blocks and intensity stereo allowed):
 *                 the body lacks the one-channel-at-a-time IMDCT and antialias paths, 13 KB of code that cost the mixed-block
 *                 VBR stream 10 % (9.3 -> 8.3 ms) just by being in the kernel;
 *         0       anything.
 *       All three give the same bits for a frame they can decode (tests/test_gpu_fast.py compares every split with class 0
 *       alone).  The same split INSIDE one kernel was measured too: two bodies chosen granule by granule are 80 KB of code that
 *       the warps of an SM walk in different places, and the instruction caches thrash (mixed-block VBR stream 9.3 -> 16.2 ms,
 *       with a hysteresis 10.6 ms) -- hence whole CTAs and separate kernels. */
template <bool ISO, int CLS> __device__ __forceinline__ void
sw_body(const p3_frame *__restrict__ frames, const p3_gc *__restrict__ gcs, const p3_tables *__restrict__ T,
        int64_t f_first, int64_t f_end, int frames_per_warp,
        const int16_t *__restrict__ is_in, const int32_t *__restrict__ count1, const uint8_t *__restrict__ scf,
        const p3_state *__restrict__ st_in, p3_state *__restrict__ st_out, int16_t *__restrict__ pcm,
        const float *__restrict__ pow43s /* signed |is|^(4/3) table, indexable -8207..8207 */,
        int classify /* 0: every CTA is of class 0 (the other kernels return): the check that all kernels give the same bits */)
{
  constexpr bool LEAN = CLS == 2, SAME = CLS >= 1;
  extern __shared__ __align__(16) uint8_t sw_dsm[];
  uint16_t *s_reo = reinterpret_cast<uint16_t *>(sw_dsm);
  uint8_t *s_sfbw = sw_dsm + 1152;
  float *s_t1h = reinterpret_cast<float *>(sw_dsm + 1728), *s_t2 = s_t1h + 40;   /* per-lane indexed: shared memory, not the constant bank */
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  sw_warp_sm *W = reinterpret_cast<sw_warp_sm *>(sw_dsm + SW_LUT_BYTES) + warp;

  {                                                        /* content class of this CTA's frames (warm-up frames included) */
    const int64_t gw_ = (int64_t)blockIdx.x * SW_WPB + warp, c0_ = f_first + gw_ * frames_per_warp;
    bool same = classify != 0, simple = same;
    if (same && c0_ < f_end) {
      const int64_t hi = min(c0_ + (int64_t)frames_per_warp, f_end);
      for (int64_t f = c0_ - (gw_ > 0 ? 1 : 0) + lane; f < hi; f += 32) {
        const uint4 fq = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(frames + f) + 16);   /* size, begin, nch, mode, mode_ext, ... */
        if (((fq.y >> 8) & 0xffu) == 1u && ((fq.y >> 16) & 1u)) simple = false;          /* joint stereo with the intensity bit */
        const uint4 *g = reinterpret_cast<const uint4 *>(gcs + 4 * f);
        #pragma unroll
        for (int gr = 0; gr < 2; gr++) {
          const uint32_t a = (g[2 * gr].y >> 4) & 15u, bb = (g[2 * gr + 1].y >> 4) & 15u;  /* win_switch | block_type << 1 | mixed << 3 */
          if (a != bb) same = false;
          if ((a & 7u) == 5u) simple = false;                                              /* short: win_switch with block_type 2 */
        }
      }
    }
    const int all_same = __syncthreads_and(same), all_simple = __syncthreads_and(same && simple);
    if ((all_simple ? 2 : all_same ? 1 : 0) != CLS) return;
  }
  const uint32_t sf = frames[f_first].sfreq;               /* a batch never mixes sample rates */
  uint8_t *s_lsfb = sw_dsm + 3168;                              /* intensity stereo: long sfb of a line, band starts (long / short) */
  uint16_t *s_sfbl = reinterpret_cast<uint16_t *>(sw_dsm + 3168 + 576), *s_sfbs = s_sfbl + 24;
  for (uint32_t i = threadIdx.x; i < 576; i += blockDim.x) { s_reo[i] = T->reorder_src[sf][i]; s_sfbw[i] = T->line_sfbw_s[sf][i]; s_lsfb[i] = T->line_sfb_l[sf][i]; }
  if (threadIdx.x < 23) s_sfbl[threadIdx.x] = T->sfb_l[sf][threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 46) s_sfbs[threadIdx.x - 32] = T->sfb_s[sf][threadIdx.x - 32];
  for (uint32_t i = threadIdx.x; i < 360; i += blockDim.x) s_t1h[i] = i < 40 ? T->t1h[i] : T->t2[i - 40];
  __syncthreads();                                         /* the only CTA-wide barrier */

  const int64_t gw = (int64_t)blockIdx.x * SW_WPB + warp;
  const int64_t c0 = f_first + gw * frames_per_warp;
  if (c0 >= f_end) return;
  const int64_t c1 = min(c0 + (int64_t)frames_per_warp, f_end);
  const int warm = gw > 0 ? 1 : 0;
  const int64_t fs = c0 - warm;
  const uint32_t sb = lane;

  /* ---- per-lane constants ---- */
  float ce[8], co[8]; int ia, ib;
  synth_window_coeffs(T, ce, co, ia, ib);
  uint32_t sfbp[3] = {0, 0, 0};                            /* long-block sfb of this subband's 18 lines, 5 bits each */
  #pragma unroll
  for (int m = 0; m < 18; m++) sfbp[m / 6] |= (uint32_t)T->line_sfb_l[sf][18 * sb + m] << (5 * (m % 6));

  /* ---- carried state: IMDCT tails in registers, DCT history rows in the ring ---- */
  f2 tail[18];
  #pragma unroll
  for (int i = 0; i < 18; i++) tail[i] = warm ? f2_make(0.0f, 0.0f) : f2_make(st_in->store[0][18 * sb + i], st_in->store[1][18 * sb + i]);
  for (uint32_t i = lane; i < 36 * SW_PITCH; i += 32) (&W->xr[0][0])[i] = f2_make(0.0f, 0.0f);
  if (lane == 0) { sw_mbar_init(&W->mbar[0], 1); sw_mbar_init(&W->mbar[1], 1); sw_mbar_init(&W->dbar[0], 1); sw_mbar_init(&W->dbar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  if (!warm)
    for (int age = 1; age <= 15; age++) W->xr[36 - age][lane] = f2_make(st_in->xhist[0][age - 1][lane], st_in->xhist[1][age - 1][lane]);
  __syncwarp();

  const int32_t rel0 = (int32_t)(fs - f_first), nfr = (int32_t)(c1 - fs);   /* first frame (relative to the launch) and frames of this warp */
  const int32_t rel_last = (int32_t)(f_end - 1 - f_first);
  const p3_frame *fr_l = frames + f_first; const p3_gc *gc_l = gcs + 4 * f_first;
  auto issue = [&](int32_t rel, int gr, int b) {           /* TMA: spectra + scalefactors of granule (rel, gr) -> buffer b */
    if (lane == 0) {
      const int32_t o = rel * 4 + 2 * gr;
      sw_mbar_expect(&W->mbar[b], 2304 + 128);
      sw_bulk_g2s(W->isb[b], is_in + (size_t)o * 576, 2304, &W->mbar[b]);
      sw_bulk_g2s(W->scf[b], scf + (size_t)o * P3_SCF_STRIDE, 128, &W->mbar[b]);
    }
  };
  auto issue_desc = [&](int32_t rel, int d) {              /* TMA: the descriptors of frame `rel` -> descriptor buffer d */
    if (lane == 0) {
      sw_mbar_expect(&W->dbar[d], 96);
      sw_bulk_g2s(&W->desc[d].fr, reinterpret_cast<const uint8_t *>(fr_l + rel) + 16, 16, &W->dbar[d]);
      sw_bulk_g2s(&W->desc[d].gc[0], gc_l + 4 * (size_t)rel, 64, &W->dbar[d]);
      sw_bulk_g2s(&W->desc[d].c1[0], count1 + 4 * (size_t)rel, 16, &W->dbar[d]);
    }
  };
  issue_desc(rel0, 0); if (nfr > 1) issue_desc(rel0 + 1, 1);
  issue(rel0, 0, 0); if (SW_NBUF == 2) issue(rel0, 1, 1);

  #pragma unroll 1
  for (int32_t q = 0; q < 2 * nfr; q++) {                  /* q: granules done by this warp */
    {
      const int32_t gr = q & 1, fi = q >> 1, rel = rel0 + fi; /* frame index within the warp's run / relative to the launch */
      const uint32_t b = SW_NBUF == 2 ? (q & 1) : 0;
      f2 *blk = &W->xr[18 * gr][0], *prv = &W->xr[18 * (gr ^ 1)][0];
      if (gr == 0) sw_mbar_wait(&W->dbar[fi & 1], (fi >> 1) & 1);
      const uint4 frq = W->desc[fi & 1].fr, ga = W->desc[fi & 1].gc[2 * gr], gb = W->desc[fi & 1].gc[2 * gr + 1];
      const uint32_t mode = (frq.y >> 8) & 0xffu, mode_ext = (frq.y >> 16) & 0xffu;
      const bool st_on = mode == 1 && mode_ext != 0;
      const bool is_on = st_on && (mode_ext & 1);

      /* ---- this granule's parameters; effective count1 (Q6: an empty part keeps the slot's previous value) ---- */
      int32_t ce0 = W->desc[fi & 1].c1[2 * gr], ce1 = W->desc[fi & 1].c1[2 * gr + 1];
      if (ga.w) ce0 = (int32_t)ga.w <= rel ? count1[(rel - (int32_t)ga.w) * 4 + 2 * gr] : st_in->count1[gr][0];
      if (gb.w) ce1 = (int32_t)gb.w <= rel ? count1[(rel - (int32_t)gb.w) * 4 + 2 * gr + 1] : st_in->count1[gr][1];
      if (rel == rel_last && lane == 0) { st_out->count1[gr][0] = ce0; st_out->count1[gr][1] = ce1; }
      const sw_par p0 = sw_unpack(ga, ce0), p1 = sw_unpack(gb, ce1);

      sw_mbar_wait(&W->mbar[b], SW_NBUF == 2 ? (q >> 1) & 1 : q & 1);
      const uint8_t (*scf2)[P3_SCF_STRIDE] = reinterpret_cast<const uint8_t (*)[P3_SCF_STRIDE]>(W->scf[b]);

      /* ---- stereo ranges (pdmp3.c:1916-1971) and the choice of the body ---- */
      const uint32_t c1r = (uint32_t)p1.c1;
      constexpr bool iso = ISO;                                /* ISO semantics of MS / intensity stereo, see k_requant */
      const uint32_t msn_all = (st_on && (mode_ext & 2)) ? (iso ? max((uint32_t)p0.c1, c1r) : min((uint32_t)p0.c1, c1r)) : 0u;   /* reference: min(count1), sic (pdmp3.c:1920) */
      /* ISO mode with intensity stereo on: bands starting at or above the right channel's count1 may be intensity coded
       * instead, so only the lines below it take MS in registers; the line-by-line pass decides the rest */
      const uint32_t msn = (iso && is_on) ? min(msn_all, c1r) : msn_all;
      /* intensity stereo touches a band only if it starts at or above the right channel's count1: with count1 beyond the
       * start of the last eligible band there is nothing to do (the usual case at high bit rates) */
      const uint32_t first_short0 = p0.first_short;
      bool is_any = false;
      if (is_on) {                                            /* first line of the last band it can touch: long sfb 20, long sfb 7 (mixed), short sfb 11 (x3) */
        const uint32_t ll = s_sfbl[first_short0 == 576 ? 20 : 7], ls = 3u * s_sfbs[11];
        is_any = (first_short0 != 0 && ll >= c1r) || (first_short0 != 576 && ls >= c1r);
        if (iso && msn_all > msn) is_any = true;             /* MS lines at or above the right channel's count1 are done line by line as well */
      }
      /* ---- band scales fl(t1*t2) (pdmp3.c:2127-2128, 2144-2146); same table layout as k_synth_fast: long blocks
       *      index = sfb (0..21), short 3*sfb+win (0..38), mixed: long bands 0..7 sit in the unused short slots 0..7.
       *      The scalefactor byte of short band 3*sfb+win is byte 24 + 3*sfb + win of the row. ---- */
      if (LEAN || (p0.first_short == 576 && p1.first_short == 576)) {
        /* long blocks in both channels (the usual case): lane = sfb for both channels, one pass, one 8-byte store */
        if (lane < 22) {
          const uint32_t pretab_l = (lane >= 11 && lane < 21) ? (0xbfa55u >> (2 * (lane - 11))) & 3u : 0u;   /* pretab[sfb = lane] (pdmp3.c:2123) */
          const uint32_t s0 = lane < 21 ? scf2[0][lane] : 0u, s1 = lane < 21 ? scf2[1][lane] : 0u;
          const float v0 = __fmul_rn(s_t1h[p0.mult * (s0 + p0.pre * pretab_l)], s_t2[p0.gg + P3_T2_BIAS]);
          const float v1 = __fmul_rn(s_t1h[p1.mult * (s1 + p1.pre * pretab_l)], s_t2[p1.gg + P3_T2_BIAS]);
          *reinterpret_cast<float2 *>(&W->scale[lane][0]) = make_float2(v0, v1);
        }
      } else if constexpr (!LEAN) {
      #pragma unroll 1
      for (int k = 0; k < 3; k++) {
        const uint32_t e = lane + 32 * k;
        if (e < 80) {
          const uint32_t c = e >= 40, bnd = e - 40 * c;
          const uint32_t fsh = c ? p1.first_short : p0.first_short, mult = c ? p1.mult : p0.mult, pre = c ? p1.pre : p0.pre, sbg = c ? p1.sbg : p0.sbg;
          const int32_t gg = c ? p1.gg : p0.gg;
          const bool longband = fsh == 576 ? bnd < 22 : (fsh == 36 && bnd < 8);
          const bool active = longband || (fsh < 576 && bnd < 39);
          const bool has_sf = longband ? bnd < 21 : bnd < 36;
          const uint32_t win = bnd % 3;
          uint32_t sc = has_sf ? scf2[c][longband ? bnd : P3_SCF_S_OFF + bnd] : 0u;
          /* pretab (pdmp3.c:2123): sfb 11..20 = 1,1,1,1,2,2,3,3,3,2, two bits each */
          if (longband) sc += pre * ((bnd >= 11 && bnd < 21) ? (0xbfa55u >> (2 * (bnd - 11))) & 3u : 0u);
          const int32_t qq = gg - (longband ? 0 : (int32_t)((sbg >> (8 * win)) & 0xffu));
          const float v = __fmul_rn(s_t1h[mult * sc], s_t2[qq + P3_T2_BIAS]);
          W->scale[bnd][c] = active ? v : 0.0f;
        }
      }
      }
      __syncwarp();

      /* ---- A: requantize the 18 lines of subband `sb`, both channels (exact arithmetic of pdmp3.c:2121-2152) ---- */
      f2 in[18];
      {
        const uint32_t *w0p = reinterpret_cast<const uint32_t *>(W->isb[b]) + 9 * sb;
        const uint32_t *w1p = w0p + 288;
        const float2 *scl = reinterpret_cast<const float2 *>(&W->scale[0][0]);
        #pragma unroll
        for (int i = 0; i < 9; i++) {
          const uint32_t wa = w0p[i], wb = w1p[i];
          #pragma unroll
          for (int h = 0; h < 2; h++) {
            const int m = 2 * i + h;
            const float2 sc = scl[(sfbp[m / 6] >> (5 * (m % 6))) & 31u];
            const int va = h ? (int)wa >> 16 : (int)(int16_t)(wa & 0xffffu), vb = h ? (int)wb >> 16 : (int)(int16_t)(wb & 0xffffu);
            in[m] = vmul(f2_make(__ldg(pow43s + va), __ldg(pow43s + vb)), f2_make(sc.x, sc.y));   /* (t1*t2)*t3, t3 = sign*|is|^(4/3); one product, nothing to contract */
          }
        }
      }
      if constexpr (!LEAN)
      if (p0.first_short < 576 || p1.first_short < 576) {          /* a short (or mixed) block in either channel */
        /* lane = subband also here: line d = 18 sb + m of the REORDERED spectrum is bitstream line reorder_src[d], its scale that
         * of (sfb, window) of that line (pdmp3.c:1786-1823, 2140-2152); first_short is 0 or 36, i.e. whole subbands */
        const bool sh0 = 18 * sb >= p0.first_short, sh1 = 18 * sb >= p1.first_short;
        if (sh0 || sh1) {
          const int16_t *is0 = reinterpret_cast<const int16_t *>(W->isb[b]), *is1 = is0 + 576;
          const float2 *scl = reinterpret_cast<const float2 *>(&W->scale[0][0]);
          #pragma unroll
          for (int m = 0; m < 18; m++) {
            const uint32_t s = s_reo[18 * sb + m], sw = s_sfbw[s];
            const float2 sc = scl[3 * (sw & 15u) + (sw >> 4)];
            float l = f2_x(in[m]), r = f2_y(in[m]);
            if (sh0) l = __fmul_rn(sc.x, __ldg(pow43s + is0[s]));
            if (sh1) r = __fmul_rn(sc.y, __ldg(pow43s + is1[s]));
            in[m] = f2_make(l, r);
          }
        }
      }

      /* ---- B: stereo (pdmp3.c:1916-1971): MS in registers ---- */
      if (18 * sb < msn) {
        const int32_t msrem = (int32_t)msn - 18 * (int32_t)sb;     /* lines of this subband below msn */
        /* the reference multiplies the float sum by the DOUBLE constant 1/sqrt 2 and rounds once to float (pdmp3.c:168,
         * 1923-1926).  Same result from packed fp32: with C = Ch + Cl (two floats), p = fl(a*Ch), e = a*Ch - p (exact, one
         * fma), t = fl(a*Cl + e), result = fl(p + t).  It can differ from the double-precision product only when that
         * lies within 2^-47 of a rounding boundary (~1e-7 of the values; none in 2*10^7 random ones, tools/proto). */
        constexpr float Ch = 0.70710677f, Cl = 1.21016175e-08f;
        #pragma unroll
        for (int m = 0; m < 18; m++) {
          const float l = f2_x(in[m]), r = f2_y(in[m]);
          const f2 ad = f2_make(__fadd_rn(l, r), __fsub_rn(l, r));
          const f2 p = vmul(ad, Ch);
#if SW_MS_EXACT
          const f2 ne = vfma(ad, -Ch, p);                          /* -(a*Ch - p), exact */
          const f2 nt = vfma(ad, -Cl, ne);                         /* -(a*Cl + e) */
          sw_fnma_if(in[m], nt, p, m < msrem);                     /* p + t, written as an fma so that nothing can be contracted into it; in place, lines below msn only */
#else
          (void)Cl;
          sw_mov_if(in[m], p, m < msrem);                          /* lines below msn only */
#endif
        }
      }
      if constexpr (!LEAN)
      if (is_any) {                                            /* intensity stereo: line by line through the scratch block */
        f2 *scr = blk;                                           /* [576] */
        #pragma unroll
        for (int m = 0; m < 18; m++) scr[18 * sb + m] = in[m];
        __syncwarp();
        const bool sh0 = first_short0 < 576;
        const uint32_t isc = iso ? 1u : 0u;
        #pragma unroll 1
        for (uint32_t d = (c1r & ~31u) + lane; d < 576; d += 32) {   /* a line below count1 is in a band that starts below it */
          if (d < msn) continue;
          float l = f2_x(scr[d]), r = f2_y(scr[d]);
          bool is_done = false;
          if (d >= first_short0) {
            /* short-block intensity (pdmp3.c:2190-2220) in reordered position; Q4: assignment through an `unsigned` */
            const uint32_t sw = s_sfbw[d], sfb = sw & 15u, win = sw >> 4;
            if (sfb < 12 && 3u * s_sfbs[sfb] >= c1r) {
              const uint32_t pp = scf2[isc][P3_SCF_S_OFF + 3 * sfb + win];
              if (iso) { if (pp < 7) { const float x = l; l = __fmul_rn(FC.is_l[pp], x); r = __fmul_rn(FC.is_r[pp], x); is_done = true; } }
              else if (pp != 7) { const float x = (float)(unsigned)(long long)l; l = x; r = x; }
            }
          } else {
            const uint32_t sfb = s_lsfb[d], lim = sh0 ? 8u : 21u;                           /* mixed: long sfb 0..7 only (pdmp3.c:1944) */
            if (sfb < lim && s_sfbl[sfb] >= c1r) {
              const uint32_t pp = scf2[isc][sfb];                                           /* reference: channel-0 scalefactor, sic (pdmp3.c:2163) */
              if (iso ? pp < 7 : pp != 7) { const float x = l; l = __fmul_rn(FC.is_l[pp & 7], x); r = __fmul_rn(FC.is_r[pp & 7], x); is_done = true; }
            }
          }
          if (d < msn_all && !is_done) {                         /* ISO mode only (msn_all == msn otherwise): MS for a line that is not intensity coded */
            const float a = __fadd_rn(l, r), b = __fsub_rn(l, r);
            l = __double2float_rn(__dmul_rn((double)a, 0.70710678118654752440));
            r = __double2float_rn(__dmul_rn((double)b, 0.70710678118654752440));
          }
          scr[d] = f2_make(l, r);
        }
        __syncwarp();
        #pragma unroll
        for (int m = 0; m < 18; m++) in[m] = scr[18 * sb + m];
        __syncwarp();
      }

      /* every lane is done with this buffer: bring in the granule after the next one */
      __syncwarp();
      if (SW_NBUF == 2) { if (q + 2 < 2 * nfr) issue(rel + 1, gr, b); }
      else if (q + 1 < 2 * nfr) issue(rel + gr, gr ^ 1, 0);

      /* ---- C: antialias (pdmp3.c:1706-1732): the butterflies across a subband boundary take the neighbour lane's
       *      lines through shuffles; both shuffles of a pair happen before either line is rewritten ---- */
      {
        const bool lo0 = sb >= 1 && sb < p0.sblim, hi0 = sb + 1 < p0.sblim;
        const bool lo1 = sb >= 1 && sb < p1.sblim, hi1 = sb + 1 < p1.sblim;
        #pragma unroll
        for (int i = 0; i < 8; i++) {
          const float ux = f2_x(in[i]), uy = f2_y(in[i]), lx = f2_x(in[17 - i]), ly = f2_y(in[17 - i]);
          const float bx = __shfl_up_sync(0xffffffffu, lx, 1), by = __shfl_up_sync(0xffffffffu, ly, 1);       /* line 18sb-1-i */
          const float ax = __shfl_down_sync(0xffffffffu, ux, 1), ay = __shfl_down_sync(0xffffffffu, uy, 1);   /* line 18(sb+1)+i */
          if (SAME || p0.sblim == p1.sblim) {
            sw_aa(in[i], f2_make(bx, by), sw_cs(i), sw_ca(i), lo0);                              /* ub (pdmp3.c:1726) */
            sw_aa(in[17 - i], f2_make(ax, ay), sw_cs(i), -sw_ca(i), hi0);                        /* lb (pdmp3.c:1725) */
          } else if constexpr (!SAME) {
            const float nux = lo0 ? __fadd_rn(__fmul_rn(ux, FC.cs[i]), __fmul_rn(bx, FC.ca[i])) : ux;
            const float nuy = lo1 ? __fadd_rn(__fmul_rn(uy, FC.cs[i]), __fmul_rn(by, FC.ca[i])) : uy;
            const float nlx = hi0 ? __fsub_rn(__fmul_rn(lx, FC.cs[i]), __fmul_rn(ax, FC.ca[i])) : lx;
            const float nly = hi1 ? __fsub_rn(__fmul_rn(ly, FC.cs[i]), __fmul_rn(ay, FC.ca[i])) : ly;
            in[i] = f2_make(nux, nuy); in[17 - i] = f2_make(nlx, nly);
          }
        }
      }

      /* ---- D: IMDCT + window (pdmp3.c:1649-1700), overlap-add with the tail in registers (1774-1777),
       *      frequency inversion (1738-1746); the subband samples go to the ring transposed [slot][subband] ---- */
      {
        const uint32_t bt0 = (p0.ws && p0.mixed && sb < 2) ? 0u : p0.bt, bt1 = (p1.ws && p1.mixed && sb < 2) ? 0u : p1.bt;
        const float sgn = (sb & 1) ? -1.0f : 1.0f;
        if (LEAN || (bt0 != 2 && (SAME || bt0 == bt1))) {
          f2 t[18];
          dct4_18<f2>(in, t);
          /* 36-point IMDCT from the DCT-IV by symmetry (signs folded into swin): output p < 18 is overlap-added with
           * the tail and leaves at once, output 18 + p becomes the new tail */
          #pragma unroll
          for (int k = 0; k < 9; k++) {
            f2 ya = vadd(vmul(t[9 + k], FC.swin[bt0][k]), tail[k]);               /* rawout + store (pdmp3.c:1775) */
            f2 yb = vadd(vmul(t[9 + k], FC.swin[bt0][17 - k]), tail[17 - k]);
            if (k & 1) ya = vmul(ya, sgn); else yb = vmul(yb, sgn);               /* frequency inversion (1741-1743): odd slots of odd subbands */
            blk[k * SW_PITCH + sb] = ya; blk[(17 - k) * SW_PITCH + sb] = yb;
          }
          #pragma unroll
          for (int k = 0; k < 9; k++) { tail[8 - k] = vmul(t[k], FC.swin[bt0][26 - k]); tail[9 + k] = vmul(t[k], FC.swin[bt0][27 + k]); }
        }
        else if constexpr (!LEAN) {
        if (SAME || bt0 == bt1) {                                      /* short windows in both channels: three 12-point IMDCTs, packed */
          f2 raw[36];
          imdct_short<f2>(in, raw);
          #pragma unroll
          for (int ss = 0; ss < 18; ss++) {
            f2 y = vadd(raw[ss], tail[ss]);
            if (ss & 1) y = vmul(y, sgn);
            blk[ss * SW_PITCH + sb] = y;
            tail[ss] = raw[18 + ss];
          }
        }
        else if constexpr (!SAME) {                                 /* different block types in the two channels: one channel at a time */
          float *blkf = reinterpret_cast<float *>(blk);
          #pragma unroll 1
          for (int c = 0; c < 2; c++) {
            const uint32_t btc = c ? bt1 : bt0;
            float v[18], xo[18], to[18];
            #pragma unroll
            for (int m = 0; m < 18; m++) v[m] = c ? f2_y(in[m]) : f2_x(in[m]);
            if (btc != 2) {
              float t[18];
              dct4_18<float>(v, t);
              #pragma unroll
              for (int k = 0; k < 9; k++) {
                to[8 - k] = __fmul_rn(t[k], FC.swin[btc][26 - k]); to[9 + k] = __fmul_rn(t[k], FC.swin[btc][27 + k]);
                xo[k] = __fmul_rn(t[9 + k], FC.swin[btc][k]); xo[17 - k] = __fmul_rn(t[9 + k], FC.swin[btc][17 - k]);
              }
            } else {
              float raw[36];
              imdct_short<float>(v, raw);
              #pragma unroll
              for (int i = 0; i < 18; i++) { xo[i] = raw[i]; to[i] = raw[18 + i]; }
            }
            #pragma unroll
            for (int ss = 0; ss < 18; ss++) {
              const float y = __fadd_rn(xo[ss], c ? f2_y(tail[ss]) : f2_x(tail[ss]));
              blkf[2 * (ss * SW_PITCH + sb) + c] = (ss & 1) ? __fmul_rn(y, sgn) : y;
              tail[ss] = c ? f2_make(f2_x(tail[ss]), to[ss]) : f2_make(to[ss], f2_y(tail[ss]));
            }
          }
        }
        }
      }
      __syncwarp();

      /* ---- E: 32-point DCT of each time slot, in place in its ring row ---- */
      if (lane < 18) {
        f2 *row = blk + lane * SW_PITCH;
        f2 s[32];
        #pragma unroll
        for (int k = 0; k < 32; k++) s[k] = row[k];
        dct2<32, f2>(s);
        #pragma unroll
        for (int k = 0; k < 32; k++) row[k] = s[k];
      }
      __syncwarp();

      /* ---- F: 512-tap window from registers + PCM (pdmp3.c:2015-2041, 2307-2345) ---- */
      if (!(warm && fi == 0) && ((frq.z >> 8) & P3_FRAME_DECODE)) {
        uint32_t *out = reinterpret_cast<uint32_t *>(pcm) + ((size_t)frq.w * 1152 + gr * 576 + lane);
        #pragma unroll
        for (int h = 0; h < 2; h++) {                              /* two runs of 9 slots over a 24-entry sliding window */
          f2 sum[9];
          #pragma unroll
          for (int pass = 0; pass < 2; pass++) {                   /* even taps (column ia), then odd taps (column ib): half the live registers */
            f2 A[24];                                              /* entry i = slot 9h + i - 15 of this granule (negative: previous granule) */
            #pragma unroll
            for (int i = 1 - pass; i < 24 - pass; i++) {           /* even taps touch entries 1..23, odd taps 0..22 */
              const int rel = 9 * h + i - 15;
              const f2 *rowp = rel < 0 ? prv + (18 + rel) * SW_PITCH : blk + rel * SW_PITCH;
              A[i] = rowp[pass ? ib : ia];
            }
            #pragma unroll
            for (int s9 = 0; s9 < 9; s9++) {
              f2 acc = pass ? sum[s9] : f2_make(0.0f, 0.0f);
              #pragma unroll
              for (int k = 0; k < 8; k++) acc = pass ? vfma(A[15 + s9 - 2 * k - 1], co[k], acc) : vfma(A[15 + s9 - 2 * k], ce[k], acc);
              sum[s9] = acc;
            }
          }
          #pragma unroll
          for (int s9 = 0; s9 < 9; s9++) out[(9 * h + s9) * 32] = sw_pcm2(sum[s9]);
        }
      }
      __syncwarp();
      if (gr == 1 && fi + 2 < nfr) issue_desc(rel + 2, fi & 1);   /* both granules of the frame are done with this descriptor buffer */
    }
  }

  /* ---- carried state out, after the last frame of the launch ---- */
  if (c1 == f_end) {
    #pragma unroll
    for (int i = 0; i < 18; i++) { st_out->store[0][18 * sb + i] = f2_x(tail[i]); st_out->store[1][18 * sb + i] = f2_y(tail[i]); }
    for (int age = 1; age <= 15; age++) {
      const f2 v = W->xr[36 - age][lane];
      st_out->xhist[0][age - 1][lane] = f2_x(v); st_out->xhist[1][age - 1][lane] = f2_y(v);
    }
  }
}

#ifdef SW_MAXNREG                   /* a register cap instead of a CTA count: SW_WPB warps per CTA, SW_CPS CTAs per SM fit at SW_MAXNREG registers */
#define SW_BOUNDS __maxnreg__(SW_MAXNREG)
#ifndef SW_CPS
#define SW_CPS (65536 / (SW_MAXNREG * SW_WPB * 32))
#endif
#else
#define SW_BOUNDS __launch_bounds__(SW_WPB * 32, SW_MINB)
#define SW_CPS SW_MINB
#endif
#define SW_KERNEL(NAME, ISO, CLS) \
extern "C" __global__ void SW_BOUNDS \
NAME(const p3_frame *__restrict__ frames, const p3_gc *__restrict__ gcs, const p3_tables *__restrict__ T, int64_t f_first, int64_t f_end, int frames_per_warp, \
     const int16_t *__restrict__ is_in, const int32_t *__restrict__ count1, const uint8_t *__restrict__ scf, \
     const p3_state *__restrict__ st_in, p3_state *__restrict__ st_out, int16_t *__restrict__ pcm, const float *__restrict__ pow43s, int classify) \
{ sw_body<ISO, CLS>(frames, gcs, T, f_first, f_end, frames_per_warp, is_in, count1, scf, st_in, st_out, pcm, pow43s, classify); }
SW_KERNEL(k_synth_warp, false, 0)
SW_KERNEL(k_synth_warp_same, false, 1)
SW_KERNEL(k_synth_warp_lean, false, 2)
SW_KERNEL(k_synth_warp_iso, true, 0)
SW_KERNEL(k_synth_warp_iso_same, true, 1)
SW_KERNEL(k_synth_warp_iso_lean, true, 2)

static int p3_synthw_check_consts(const float *cs, const float *ca)
{
  const float k_cs[8] = SW_CS_LIST, k_ca[8] = SW_CA_LIST;
  for (int i = 0; i < 8; i++) if (cs[i] != k_cs[i] || ca[i] != k_ca[i]) return -1000 - i;
  return 0;
}

extern "C" size_t p3_synthw_smem_bytes(void) { return SW_LUT_BYTES + SW_WPB * sizeof(sw_warp_sm); }
extern "C" int p3_synthw_warps_per_cta(void) { return SW_WPB; }
extern "C" int p3_synthw_warps_per_sm(void) { return SW_WPB * SW_CPS; }      /* resident warps per SM (one wave = n_sm x this x frames per warp) */
