/* tc_matrix.cu -- the tensor-core trial north_star asks for: the polyphase matrixing stage of
 * L3_Subband_Synthesis (pdmp3.c:2010-2014; as a 32-point DCT-II it is stage E of k_synth_warp, p3_synthw.cuh) done
 *   (a) the way the product does it: Lee's fast DCT-32 on packed fp32 pairs (FFMA2/FADD2/FMUL2), lane = time slot, 18 of 32
 *       lanes busy, in place in the warp's DCT ring in shared memory;
 *   (b) on the 5th-generation tensor cores: tcgen05.mma kind::tf32, M = 128 (4 warps x 32 rows, 18 of them real time slots),
 *       N = 32 DCT outputs, K = 32 subbands, THREE passes per channel (A_hi B_hi + A_lo B_hi + A_hi B_lo: a single tf32 or
 *       bf16 pass is 16-64 LSB of int16 off), accumulators in TMEM, read back with tcgen05.ld so that lane =
 *       time slot holds the 32 outputs exactly where stage E leaves them, then to the ring like (a).
 * Both variants sit in the same harness, shaped like the stage's surroundings in k_synth_warp: every iteration each lane
 * (= subband) first deposits the 18 time-slot samples of both channels (what stage D does), and afterwards reads a window
 * column back (a stand-in for stage F), so that the stage's shared-memory traffic, barriers and latencies are all paid.
 * No global traffic inside the loop: this measures what the stage costs an SM, which is what decides (the fused kernel is
 * bound by issue slots and the FMA pipe, not by HBM).
 *
 * build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I pdmp3_b200/csrc -o tools/tc_trial/tc_matrix tools/tc_trial/tc_matrix.cu
 * run:   tools/tc_trial/tc_matrix [iterations]      (prints one JSON line; under ncu: -k regex:k_stage_)            */
#include "p3_xform.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define PITCH 33
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

/* the 18 samples a lane deposits in iteration `it`: cheap, different every iteration, nothing the compiler can hoist */
__device__ __forceinline__ f2 sample(const f2 seed, int slot, int it)
{
  const float s = 1.0f + 0.03125f * (float)((slot * 7 + it) & 15);
  return vmul(seed, s);
}

/* ------------------------------------------------------------------ (a) FFMA: Lee DCT-32 on 18 lanes ---- */
extern "C" __global__ void __launch_bounds__(128, 3)
k_stage_lee(const f2 *__restrict__ seeds, int iters, f2 *__restrict__ out /* [cta][warp][18][32] of the LAST iteration, or NULL */, float *__restrict__ sink)
{
  extern __shared__ __align__(16) uint8_t dsm[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  f2 *ring = reinterpret_cast<f2 *>(dsm) + (size_t)warp * 36 * PITCH;
  for (uint32_t i = lane; i < 36 * PITCH; i += 32) ring[i] = f2_make(0.0f, 0.0f);
  const f2 seed = seeds[(size_t)blockIdx.x * 128 + threadIdx.x];
  f2 acc = f2_make(0.0f, 0.0f);
  __syncwarp();
  #pragma unroll 1
  for (int it = 0; it < iters; it++) {
    f2 *blk = ring + (it & 1) * 18 * PITCH;
    #pragma unroll
    for (int s = 0; s < 18; s++) blk[s * PITCH + lane] = sample(seed, s, it);          /* stage D's stores: [slot][subband] */
    __syncwarp();
    if (lane < 18) {                                                                     /* stage E as in p3_synthw.cuh */
      f2 *row = blk + lane * PITCH;
      f2 x[32];
      #pragma unroll
      for (int k = 0; k < 32; k++) x[k] = row[k];
      dct2<32, f2>(x);
      #pragma unroll
      for (int k = 0; k < 32; k++) row[k] = x[k];
    }
    __syncwarp();
    #pragma unroll
    for (int s = 0; s < 18; s++) acc = vadd(acc, blk[s * PITCH + lane]);                 /* stand-in for stage F: a column of the ring */
    __syncwarp();
  }
  if (out) {
    const f2 *blk = ring + ((iters - 1) & 1) * 18 * PITCH;
    for (int s = 0; s < 18; s++) out[(((size_t)blockIdx.x * 4 + warp) * 18 + s) * 32 + lane] = blk[s * PITCH + lane];
  }
  if (sink && f2_x(acc) == 1.2345e33f) sink[0] = f2_y(acc);
}

/* ------------------------------------------------------------------ (b) tcgen05: 3 x tf32 GEMM ---- */
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 22)) __trap();                                           /* never hang the box */
  } while (!ok);
}
/* shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor) */
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
  return (uint64_t)((addr >> 4) & 0x3fffu) | (uint64_t)1 << 16 /* LBO (unused with a swizzle) */ | (uint64_t)(1024 >> 4) << 32 /* SBO */
       | (uint64_t)1 << 46 /* version: Blackwell */ | (uint64_t)2 << 61 /* SWIZZLE_128B */;
}
/* byte offset of element (row r, k) of a K-major SW128 tile of fp32 (Swizzle<3,4,3>) */
__device__ __host__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t k) { return (r >> 3) * 1024u + (r & 7u) * 128u + ((((k >> 2) ^ (r & 7u)) & 7u) << 4) + (k & 3u) * 4u; }

#define TC_TILE_A (128 * 128)              /* bytes: 128 rows x 32 fp32 */
#define TC_TILE_B (32 * 128)
#define TC_SMEM   (1024 + 4 * TC_TILE_A + 2 * TC_TILE_B + 4 * 36 * PITCH * 8)

extern "C" __global__ void __launch_bounds__(128, 2)
k_stage_tc(const f2 *__restrict__ seeds, const float *__restrict__ Bhi, const float *__restrict__ Blo /* [32][32] DCT matrix, C[n][k] */, int iters,
           f2 *__restrict__ out, float *__restrict__ sink)
{
  extern __shared__ __align__(16) uint8_t dsm[];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_tmem;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t *base = dsm + ((1024u - (s32(dsm) & 1023u)) & 1023u);                          /* the swizzle pattern is anchored at 1024-byte boundaries */
  uint8_t *tA = base;                                                                    /* [ch][hi|lo] tiles of 16 KB */
  uint8_t *tB = tA + 4 * TC_TILE_A;                                                      /* [hi|lo] tiles of 4 KB */
  f2 *ring = reinterpret_cast<f2 *>(tB + 2 * TC_TILE_B) + (size_t)warp * 36 * PITCH;
  for (uint32_t i = lane; i < 36 * PITCH; i += 32) ring[i] = f2_make(0.0f, 0.0f);
  for (uint32_t i = threadIdx.x; i < 4 * TC_TILE_A / 4; i += 128) reinterpret_cast<uint32_t *>(tA)[i] = 0;     /* rows 18..31 of every warp's group stay zero */
  for (uint32_t i = threadIdx.x; i < 1024; i += 128) {
    const uint32_t n = i >> 5, k = i & 31;
    *reinterpret_cast<float *>(tB + sw128(n, k)) = Bhi[i];
    *reinterpret_cast<float *>(tB + TC_TILE_B + sw128(n, k)) = Blo[i];
  }
  const uint32_t bar = s32(&s_bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {                                                                       /* 64 TMEM columns: D of channel 0 | D of channel 1 */
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(s32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  /* kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 32 (cute::UMMA::InstrDescriptor) */
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  const f2 seed = seeds[(size_t)blockIdx.x * 128 + threadIdx.x];
  f2 acc = f2_make(0.0f, 0.0f);
  #pragma unroll 1
  for (int it = 0; it < iters; it++) {
    f2 *blk = ring + (it & 1) * 18 * PITCH;
    /* stage D's stores, in the operand layout of the MMA: row = 32 warp + slot, k = subband; hi = the tf32 the hardware will
     * see (top 19 bits), lo = what it cuts off -- per channel, so four scalar stores per sample instead of one 64-bit store */
    #pragma unroll
    for (int s = 0; s < 18; s++) {
      const f2 v = sample(seed, s, it);
      const float x = f2_x(v), y = f2_y(v);
      const float xh = __uint_as_float(__float_as_uint(x) & 0xffffe000u), yh = __uint_as_float(__float_as_uint(y) & 0xffffe000u);
      const uint32_t o = sw128(32 * warp + s, lane);
      *reinterpret_cast<float *>(tA + 0 * TC_TILE_A + o) = xh; *reinterpret_cast<float *>(tA + 1 * TC_TILE_A + o) = x - xh;
      *reinterpret_cast<float *>(tA + 2 * TC_TILE_A + o) = yh; *reinterpret_cast<float *>(tA + 3 * TC_TILE_A + o) = y - yh;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                         /* generic-proxy stores -> visible to the tensor core's async proxy */
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      #pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        #pragma unroll
        for (int pass = 0; pass < 3; pass++) {                                           /* A_hi B_hi, A_lo B_hi, A_hi B_lo */
          const uint32_t a0 = s32(tA + (2 * ch + (pass == 1)) * TC_TILE_A), b0 = s32(tB + (pass == 2) * TC_TILE_B);
          #pragma unroll
          for (int k = 0; k < 4; k++) {                                                  /* K = 32 in steps of 8 tf32 (32 bytes inside the swizzled row) */
            const uint64_t da = smem_desc(a0 + 32 * k), db = smem_desc(b0 + 32 * k);
            const uint32_t accum = (pass | k) != 0;
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                         :: "r"(tmem + 32 * ch), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
    }
    mbar_wait(bar, (uint32_t)it & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    /* lane = time slot again: its 32 outputs of both channels, exactly where stage E leaves them */
    uint32_t r0[32], r1[32];
    const uint32_t ta = tmem + ((32u * warp) << 16);
#define LD32(R, ADDR) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(R[0]),"=r"(R[1]),"=r"(R[2]),"=r"(R[3]),"=r"(R[4]),"=r"(R[5]),"=r"(R[6]),"=r"(R[7]),"=r"(R[8]),"=r"(R[9]),"=r"(R[10]),"=r"(R[11]),"=r"(R[12]),"=r"(R[13]),"=r"(R[14]),"=r"(R[15]), \
        "=r"(R[16]),"=r"(R[17]),"=r"(R[18]),"=r"(R[19]),"=r"(R[20]),"=r"(R[21]),"=r"(R[22]),"=r"(R[23]),"=r"(R[24]),"=r"(R[25]),"=r"(R[26]),"=r"(R[27]),"=r"(R[28]),"=r"(R[29]),"=r"(R[30]),"=r"(R[31]) : "r"(ADDR) : "memory")
    LD32(r0, ta); LD32(r1, ta + 32);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (lane < 18) {
      f2 *row = blk + lane * PITCH;
      #pragma unroll
      for (int k = 0; k < 32; k++) row[k] = f2_make(__uint_as_float(r0[k]), __uint_as_float(r1[k]));
    }
    __syncwarp();
    #pragma unroll
    for (int s = 0; s < 18; s++) acc = vadd(acc, blk[s * PITCH + lane]);
    __syncwarp();
  }
  if (out) {
    const f2 *blk = ring + ((iters - 1) & 1) * 18 * PITCH;
    for (int s = 0; s < 18; s++) out[(((size_t)blockIdx.x * 4 + warp) * 18 + s) * 32 + lane] = blk[s * PITCH + lane];
  }
  if (sink && f2_x(acc) == 1.2345e33f) sink[0] = f2_y(acc);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem) : "memory");
}

int main(int argc, char **argv)
{
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int nsm = prop.multiProcessorCount;
  const int grid_lee = nsm * 3, grid_tc = nsm * 2, gmax = grid_lee;
  /* DCT-II matrix of dct2<32>: X[k] = sum_n x[n] cos(pi (2n+1) k / 64); split for the three tf32 passes */
  static float Bhi[1024], Blo[1024];
  for (int n = 0; n < 32; n++) for (int k = 0; k < 32; k++) {
    const double c = cos(3.14159265358979323846 * (2 * k + 1) * n / 64.0);                 /* row n = output index, column k = subband */
    float h = (float)c; uint32_t u; memcpy(&u, &h, 4); u &= 0xffffe000u; memcpy(&h, &u, 4);
    Bhi[n * 32 + k] = h; Blo[n * 32 + k] = (float)(c - (double)h);
  }
  float *dBhi, *dBlo, *dsink; f2 *dseed, *dout_a, *dout_b;
  CK(cudaMalloc(&dBhi, sizeof Bhi)); CK(cudaMalloc(&dBlo, sizeof Blo)); CK(cudaMalloc(&dsink, 16));
  CK(cudaMemcpy(dBhi, Bhi, sizeof Bhi, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBlo, Blo, sizeof Blo, cudaMemcpyHostToDevice));
  const size_t nseed = (size_t)gmax * 128, nout = (size_t)gmax * 4 * 18 * 32;
  float *hseed = (float *)malloc(nseed * 8);
  uint32_t rng = 12345u;
  for (size_t i = 0; i < 2 * nseed; i++) { rng = rng * 1664525u + 1013904223u; hseed[i] = ((int)(rng >> 8) - (1 << 23)) * (1.0f / (1 << 23)) * 0.3f; }
  CK(cudaMalloc(&dseed, nseed * 8)); CK(cudaMemcpy(dseed, hseed, nseed * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dout_a, nout * 8)); CK(cudaMalloc(&dout_b, nout * 8));
  const size_t smem_lee = 4 * 36 * PITCH * 8;
  CK(cudaFuncSetAttribute(k_stage_lee, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lee));
  CK(cudaFuncSetAttribute(k_stage_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  /* ---- numerics: one iteration of both on the same samples, against a double-precision DCT ---- */
  k_stage_lee<<<grid_tc, 128, smem_lee>>>(dseed, 1, dout_a, dsink); CK(cudaGetLastError());
  k_stage_tc<<<grid_tc, 128, TC_SMEM>>>(dseed, dBhi, dBlo, 1, dout_b, dsink); CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  const size_t ncmp = (size_t)grid_tc * 4 * 18 * 32;
  float *ha = (float *)malloc(ncmp * 8), *hb = (float *)malloc(ncmp * 8);
  CK(cudaMemcpy(ha, dout_a, ncmp * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hb, dout_b, ncmp * 8, cudaMemcpyDeviceToHost));
  double err_a = 0, err_b = 0, err_ab = 0, mx = 0;
  for (int cta = 0; cta < 8; cta++) for (int w = 0; w < 4; w++) for (int s = 0; s < 18; s++) for (int ch = 0; ch < 2; ch++) {
    double x[32];
    const float sc = 1.0f + 0.03125f * (float)((s * 7 + 0) & 15);
    for (int n = 0; n < 32; n++) x[n] = (double)(hseed[2 * ((size_t)cta * 128 + w * 32 + n) + ch] * sc);
    for (int k = 0; k < 32; k++) {
      double ref = 0; for (int n = 0; n < 32; n++) ref += x[n] * cos(3.14159265358979323846 * (2 * n + 1) * k / 64.0);
      const size_t o = 2 * ((((size_t)cta * 4 + w) * 18 + s) * 32 + k) + ch;
      err_a = fmax(err_a, fabs(ha[o] - ref)); err_b = fmax(err_b, fabs(hb[o] - ref)); mx = fmax(mx, fabs(ref));
    }
  }
  for (size_t i = 0; i < 2 * ncmp; i++) err_ab = fmax(err_ab, fabs((double)ha[i] - (double)hb[i]));
  /* ---- timing ---- */
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms_a = 0, ms_b = 0;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0)); k_stage_lee<<<grid_lee, 128, smem_lee>>>(dseed, iters, NULL, dsink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_a, e0, e1));
    CK(cudaEventRecord(e0)); k_stage_tc<<<grid_tc, 128, TC_SMEM>>>(dseed, dBhi, dBlo, iters, NULL, dsink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_b, e0, e1));
  }
  CK(cudaGetLastError());
  /* granules per SM: iterations x CTAs per SM x 4 warps; the 10^6-frame benchmark has 2 x 10^6 granules over nsm SMs */
  const double ns_a = ms_a * 1e6 / ((double)iters * 3 * 4), ns_b = ms_b * 1e6 / ((double)iters * 2 * 4);
  printf("{\"sms\": %d, \"iters\": %d, \"lee\": {\"ctas_per_sm\": 3, \"ms\": %.4f, \"ns_per_granule_per_sm\": %.2f, \"ms_per_1M_frames\": %.3f, \"max_abs_err\": %.3e},"
         " \"tcgen05_tf32x3\": {\"ctas_per_sm\": 2, \"smem_per_cta\": %d, \"ms\": %.4f, \"ns_per_granule_per_sm\": %.2f, \"ms_per_1M_frames\": %.3f, \"max_abs_err\": %.3e},"
         " \"max_abs_diff_between_them\": %.3e, \"max_abs_value\": %.3f}\n",
         nsm, iters, ms_a, ns_a, ns_a * 2e6 / nsm * 1e-6, err_a, TC_SMEM, ms_b, ns_b, ns_b * 2e6 / nsm * 1e-6, err_b, err_ab, mx);
  return 0;
}
