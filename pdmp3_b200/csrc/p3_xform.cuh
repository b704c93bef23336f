/* p3_xform.cuh -- the fast transforms of FAST mode, written once over a value type V:
 *   V = float : one channel per thread (k_synth_fast, mono streams and the tap path)
 *   V = f2    : both channels of a stereo granule packed in one 64-bit register pair and processed
 *               with Blackwell's packed fp32 instructions (FFMA2 / FADD2 / FMUL2 via PTX *.f32x2),
 *               which halves the issue slots of every transform stage (k_synth_warp).
 * Every operation is written as an explicitly rounded add / mul / fma.  nvcc keeps the scalar form as written;
 * ptxas still contracts some packed mul.rn.f32x2 + add.rn.f32x2 pairs into one FFMA2 (fewer roundings), so the
 * two instantiations agree to rounding error, not bit for bit (measured: tools/dbg/t_f2.cu).
 *
 * Replaces the O(N^2) loops of the reference: IMDCT_Win (pdmp3.c:1649-1700) and the 64x32 matrixing of
 * L3_Subband_Synthesis (pdmp3.c:2010-2014); derivations in tools/proto/fast_transforms.py. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "p3_lee.inc"

struct f2 { unsigned long long v; };                      /* {x = left/ch0 (low word), y = right/ch1 (high word)} */

__device__ __forceinline__ f2 f2_make(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float f2_x(f2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return x; }
__device__ __forceinline__ float f2_y(f2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return y; }

/* ---- explicitly rounded operations, same names for both value types ---- */
__device__ __forceinline__ float vadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float vsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float vmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float vfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ float vzero(float) { return 0.0f; }

__device__ __forceinline__ f2 vadd(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 vsub(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 vmul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 vfma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
/* scalar second operand: ptxas folds the {s, s} pair into the broadcast operand form of FMUL2 / FFMA2 */
__device__ __forceinline__ f2 vmul(f2 a, float s) { return vmul(a, f2_make(s, s)); }
__device__ __forceinline__ f2 vfma(f2 a, float s, f2 c) { return vfma(a, f2_make(s, s), c); }
__device__ __forceinline__ f2 vneg(f2 a) { return vmul(a, -1.0f); }
__device__ __forceinline__ f2 vzero(f2) { return f2_make(0.0f, 0.0f); }

/* ---- 32-point DCT-II, Lee's recursion, fully unrolled in registers: 80 mul + 209 add ---- */
template <int N> struct LeeTab;
#define P3_LEE_TAB(N) template <> struct LeeTab<N> { static __device__ __forceinline__ float c(int i) { constexpr float t[N / 2] = P3_LEE##N; return t[i]; } };
P3_LEE_TAB(32) P3_LEE_TAB(16) P3_LEE_TAB(8) P3_LEE_TAB(4) P3_LEE_TAB(2)

template <int N, class V> __device__ __forceinline__ void dct2(V (&x)[N])
{
  if constexpr (N == 1) return;
  else {
    V a[N / 2], b[N / 2];
    #pragma unroll
    for (int i = 0; i < N / 2; i++) { a[i] = vadd(x[i], x[N - 1 - i]); b[i] = vmul(vsub(x[i], x[N - 1 - i]), LeeTab<N>::c(i)); }
    dct2<N / 2, V>(a); dct2<N / 2, V>(b);
    #pragma unroll
    for (int i = 0; i < N / 2; i++) { x[2 * i] = a[i]; x[2 * i + 1] = (i + 1 < N / 2) ? vadd(b[i], b[i + 1]) : b[i]; }
  }
}

/* ---- 18-point DCT-IV via two 9-point DCT-IIs --------------------------------------------------
 * y[m] = x[m] * 2cos(pi(2m+1)/72);  Y = DCT-II-18(y) by one Lee split into two DCT-II-9;
 * t[0] = Y[0]/2, t[k] = Y[k] - t[k-1].   ~125 flops instead of 324. */
template <class V> __device__ __forceinline__ void dct9(const V (&x)[9], V (&X)[9])
{
  constexpr float C10 = 9.848077530e-01f, C20 = 9.396926208e-01f, C30 = 8.660254038e-01f, C40 = 7.660444431e-01f, C50 = 6.427876097e-01f, C70 = 3.420201433e-01f, C80 = 1.736481777e-01f;
  const V s0 = vadd(x[0], x[8]), s1 = vadd(x[1], x[7]), s2 = vadd(x[2], x[6]), s3 = vadd(x[3], x[5]), x4 = x[4];
  const V d0 = vsub(x[0], x[8]), d1 = vsub(x[1], x[7]), d2 = vsub(x[2], x[6]), d3 = vsub(x[3], x[5]);
  const V h1 = vmul(s1, 0.5f);
  const V hm = vsub(h1, x4), mh = vsub(x4, h1);
  X[0] = vadd(vadd(vadd(s0, s1), vadd(s2, s3)), x4);
  X[2] = vfma(s0, C20, vfma(s2, -C80, vfma(s3, -C40, hm)));
  X[4] = vfma(s0, C40, vfma(s2, -C20, vfma(s3, C80, mh)));
  X[6] = vfma(vadd(vadd(s0, s2), s3), 0.5f, vneg(vadd(s1, x4)));
  X[8] = vfma(s0, C80, vfma(s2, C40, vfma(s3, -C20, mh)));
  const V e1 = vmul(d1, C30), ne1 = vmul(d1, -C30);
  X[1] = vfma(d0, C10, vfma(d2, C50, vfma(d3, C70, e1)));
  X[3] = vmul(vsub(vsub(d0, d2), d3), C30);
  X[5] = vfma(d0, C50, vfma(d2, -C70, vfma(d3, C10, ne1)));
  X[7] = vfma(d0, C70, vfma(d2, C10, vfma(d3, -C50, ne1)));
}

template <class V> __device__ __forceinline__ void dct4_18(const V (&x)[18], V (&t)[18])
{
  constexpr float PRE[18] = {1.998096443e+00f, 1.982889723e+00f, 1.952592014e+00f, 1.907433901e+00f, 1.847759065e+00f, 1.774021666e+00f, 1.686782892e+00f, 1.586706681e+00f, 1.474554674e+00f, 1.351180415e+00f, 1.217522858e+00f, 1.074599217e+00f, 9.234972265e-01f, 7.653668647e-01f, 6.014115990e-01f, 4.328792279e-01f, 2.610523844e-01f, 8.723877473e-02f};
  constexpr float LEE[9] = {5.019099188e-01f, 5.176380902e-01f, 5.516889595e-01f, 6.103872944e-01f, 7.071067812e-01f, 8.717233978e-01f, 1.183100792e+00f, 1.931851653e+00f, 5.736856623e+00f};
  V a[9], b[9], A[9], B[9];
  #pragma unroll
  for (int m = 0; m < 9; m++) {
    const V u = vmul(x[m], PRE[m]), v = vmul(x[17 - m], PRE[17 - m]);
    a[m] = vadd(u, v); b[m] = vmul(vsub(u, v), LEE[m]);
  }
  dct9<V>(a, A); dct9<V>(b, B);
  t[0] = vmul(A[0], 0.5f);
  #pragma unroll
  for (int k = 1; k < 18; k++) {
    const V Y = (k & 1) ? (k == 17 ? B[8] : vadd(B[k >> 1], B[(k >> 1) + 1])) : A[k >> 1];
    t[k] = vsub(Y, t[k - 1]);
  }
}
