/* p3_device.cuh -- device-side building blocks shared by the sm_100a kernels.
 * Reference lines each block replaces are cited at the function. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "p3_tables.h"
#include "../../include/pdmp3_b200.h"

#define P3_SCF_STRIDE 64          /* bytes per granule-channel in the scalefactor buffer: l[21] pad s[12][3] */
#define P3_SCF_S_OFF  24

/* filter/decoder state carried from batch to batch (double buffered in the context).
 * Replaces the function-static arrays of the reference: store (pdmp3.c:1755), v_vec (1983),
 * and the stale count1 of Q6 (pdmp3.c:2057-2061). */
struct p3_state {
  float store[2][576];            /* second half of the last granule's IMDCT, [ch][sb*18+i]        */
  float vhist[2][15][64];         /* matrixed vectors of the last 15 slots, [ch][age-1][i], age 1 = most recent */
  int32_t count1[2][2];           /* effective count1 of the last frame, [gr][ch]                   */
  int32_t pad[4];
  float xhist[2][15][32];         /* FAST mode: 32-point DCT of the last 15 slots, [ch][age-1][k]   */
};

/* ---- MSB-first random bit access over big-endian words in shared memory (pdmp3.c:1489-1527) ---- */
__device__ __forceinline__ uint32_t p3_peek32(const uint32_t *sw, uint32_t bitpos)
{
  uint32_t i = bitpos >> 5;
  return __funnelshift_l(sw[i + 1], sw[i], bitpos & 31);
}
__device__ __forceinline__ uint32_t p3_getbits(const uint32_t *sw, uint32_t &bitpos, uint32_t n)
{
  if (n == 0) return 0;
  uint32_t v = p3_peek32(sw, bitpos) >> (32 - n);
  bitpos += n;
  return v;
}
