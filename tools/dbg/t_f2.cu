// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dbg/t_f2 tools/dbg/t_f2.cu ; run on a B200 (gpurun)
// packed (f2) vs scalar transforms on random data: must be bit-identical
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../pdmp3_b200/csrc/p3_xform.cuh"
__global__ void k(const float *in, float *out, int n)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n) return;
  const float *p = in + (size_t)t * 64;
  float a[32], b[32]; f2 c[32];
  for (int i = 0; i < 32; i++) { a[i] = p[i]; b[i] = p[32 + i]; c[i] = f2_make(a[i], b[i]); }
  dct2<32, float>(a); dct2<32, float>(b); dct2<32, f2>(c);
  float *o = out + (size_t)t * 256;
  for (int i = 0; i < 32; i++) { o[i] = a[i]; o[32 + i] = b[i]; o[64 + i] = f2_x(c[i]); o[96 + i] = f2_y(c[i]); }
  float x[18], y[18], tx[18], ty[18]; f2 z[18], tz[18];
  for (int i = 0; i < 18; i++) { x[i] = p[i] * 3.0f; y[i] = p[40 + i]; z[i] = f2_make(x[i], y[i]); }
  dct4_18<float>(x, tx); dct4_18<float>(y, ty); dct4_18<f2>(z, tz);
  for (int i = 0; i < 18; i++) { o[128 + i] = tx[i]; o[146 + i] = ty[i]; o[192 + i] = f2_x(tz[i]); o[210 + i] = f2_y(tz[i]); }
  // single ops
  f2 m = vmul(f2_make(p[0], p[1]), p[2]); o[170] = __fmul_rn(p[0], p[2]); o[171] = __fmul_rn(p[1], p[2]); o[234] = f2_x(m); o[235] = f2_y(m);
  f2 q = vfma(f2_make(p[3], p[4]), p[5], f2_make(p[6], p[7])); o[172] = __fmaf_rn(p[3], p[5], p[6]); o[173] = __fmaf_rn(p[4], p[5], p[7]); o[236] = f2_x(q); o[237] = f2_y(q);
  f2 r = vsub(f2_make(p[8], p[9]), f2_make(p[10], p[11])); o[174] = __fsub_rn(p[8], p[10]); o[175] = __fsub_rn(p[9], p[11]); o[238] = f2_x(r); o[239] = f2_y(r);
}
int main()
{
  int n = 1 << 16; size_t ni = (size_t)n * 64, no = (size_t)n * 256;
  float *h = (float *)malloc(ni * 4), *ho = (float *)malloc(no * 4);
  srand(1); for (size_t i = 0; i < ni; i++) h[i] = ((rand() & 0xffff) - 32768) / 3277.0f * ((rand() & 7) == 0 ? 1e-3f : 1.0f);
  float *d, *dd; cudaMalloc(&d, ni * 4); cudaMalloc(&dd, no * 4); cudaMemcpy(d, h, ni * 4, cudaMemcpyHostToDevice); cudaMemset(dd, 0, no * 4);
  k<<<n / 128, 128>>>(d, dd, n); cudaMemcpy(ho, dd, no * 4, cudaMemcpyDeviceToHost);
  long bad[5] = {0, 0, 0, 0, 0};
  for (int t = 0; t < n; t++) { float *o = ho + (size_t)t * 256;
    for (int i = 0; i < 64; i++) if (memcmp(&o[i], &o[64 + i], 4)) bad[0]++;
    for (int i = 0; i < 36; i++) if (memcmp(&o[128 + i], &o[192 + i], 4)) bad[1]++;
    for (int i = 0; i < 2; i++) { if (memcmp(&o[170 + i], &o[234 + i], 4)) bad[2]++; if (memcmp(&o[172 + i], &o[236 + i], 4)) bad[3]++; if (memcmp(&o[174 + i], &o[238 + i], 4)) bad[4]++; } }
  printf("mismatches: dct32 %ld  dct4_18 %ld  mul %ld  fma %ld  sub %ld  (err %s)\n", bad[0], bad[1], bad[2], bad[3], bad[4], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
