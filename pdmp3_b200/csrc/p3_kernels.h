/* p3_kernels.h -- launch geometry shared by p3_kernels.cu and p3_cabi.cu */
#pragma once
#include <stdint.h>
#include "p3_tables.h"
#include "../../include/pdmp3_b200.h"

#ifndef K1_FPB
#define K1_FPB      32                 /* frames per CTA in the Huffman kernel (32 or 64: the sort keys carry the part index in a byte) */
#endif
#define K1_THREADS  (K1_FPB * 4)       /* one thread per granule-channel       */
#ifdef K1_LUT_GLOBAL                    /* experiment (p3_k1.cuh): the LUT stays in global memory */
#define K1_LUT_SMEM(used) ((size_t)0)
#else
#define K1_LUT_SMEM(used) ((size_t)(used) * 2)
#endif
#define K2_THREADS  192                /* requantize kernel: one CTA per granule, 3 lines per thread */
#define K3_THREADS  576                /* IMDCT kernel: one CTA per granule-channel, one thread per output sample */
#define K4_GRAN     4                  /* polyphase kernel: granules per CTA   */
#define K4_SLOTS    (K4_GRAN * 18)
#define K4_THREADS  256

struct p3_state;

#ifdef __CUDACC__
extern "C" {
__global__ void k_compact(const uint8_t *raw, const p3_frame *frames, const uint8_t *tail, int64_t f_first, int64_t f_end, uint32_t *ms);
__global__ void k_huffman(const uint32_t *ms, const p3_frame *frames, const p3_gc *gcs, const p3_tables *T,
                          int64_t f_first, int64_t f_end, uint32_t smem_words, int16_t *is_out, int32_t *count1_out, uint8_t *scf_out);
__global__ void k_sideinfo(const uint8_t *raw, p3_frame *frames, p3_gc *gcs, int64_t n_frames, int *any_empty);
__global__ void k_q6_chain(const p3_frame *frames, p3_gc *gcs, int64_t n_frames, const int *any_empty);
__global__ void k_requant(const p3_frame *frames, const p3_gc *gcs, const p3_tables *T, int64_t f_first, int64_t f_end,
                          const int16_t *is_in, const int32_t *count1, const uint8_t *scf,
                          const p3_state *st_in, p3_state *st_out, float *xr_out);
__global__ void k_imdct(const p3_frame *frames, const p3_gc *gcs, const p3_tables *T, int64_t f_first, int64_t f_end,
                        const float *xr, const p3_state *st_in, p3_state *st_out, float *y_out);
__global__ void k_polyphase(const p3_frame *frames, const p3_tables *T, int64_t f_first, int64_t f_end,
                            const float *y, const p3_state *st_in, p3_state *st_out, int16_t *pcm);
}
#endif
