/* pdmp3.h -- the libmpg123-subset streaming API of technosaurus/PDMP3, served by the
 * B200-native decoder (libpdmp3_b200.so).  Drop-in for the declarations at
 * /root/reference/pdmp3.c:114-159: same seven prototypes, same return codes, same output
 * format (interleaved host-endian int16).  The handle is opaque here (the reference exposes
 * its struct only because it is a single-file library).
 *
 * Extension that stays inside the reference's signature: pdmp3_new()'s `decoder` string,
 * which the reference ignores (pdmp3.c:2351-2353), selects options, e.g.
 *     pdmp3_new("b200:ring=1073741824,device=0,mode=exact", &err)
 * `ring` = capacity of the input buffer in bytes (default 16384 = INBUF_SIZE, pdmp3.c:123,
 * which reproduces the reference's PDMP3_NO_SPACE behaviour exactly).  A large ring lets one
 * pdmp3_feed()+pdmp3_read() pair push a whole stream through the GPU in one batch.
 * `mode` = fast (default: fused kernel with fast transforms, PCM within 1 LSB of the reference) or
 * exact (direct-form transforms in the reference's summation order, PCM bit-identical).
 * `iso` = decode by ISO 11172-3 where pdmp3.c deviates from it (count1 table B, MS stereo range, intensity
 * stereo positions, empty parts: P3_FRAME_ISO in pdmp3_b200.h); without it the output is that of pdmp3.c.
 */
#ifndef PDMP3_H
#define PDMP3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PDMP3_OK           0
#define PDMP3_ERR         -1
#define PDMP3_NEED_MORE  -10
#define PDMP3_NEW_FORMAT -11
#define PDMP3_NO_SPACE     7

#define PDMP3_ENC_SIGNED_16 (0x080|0x040|0x10)

typedef struct pdmp3_handle pdmp3_handle;

pdmp3_handle *pdmp3_new(const char *decoder, int *error);                 /* pdmp3.c:2351 */
void pdmp3_delete(pdmp3_handle *id);                                      /* pdmp3.c:2360 */
int  pdmp3_open_feed(pdmp3_handle *id);                                   /* pdmp3.c:2369 */
int  pdmp3_feed(pdmp3_handle *id, const unsigned char *in, size_t size);  /* pdmp3.c:2391 */
int  pdmp3_read(pdmp3_handle *id, unsigned char *outmemory, size_t outsize, size_t *done); /* pdmp3.c:2431 */
int  pdmp3_decode(pdmp3_handle *id, const unsigned char *in, size_t insize,
                  unsigned char *out, size_t outsize, size_t *done);      /* pdmp3.c:2491 */
int  pdmp3_getformat(pdmp3_handle *id, long *rate, int *channels, int *encoding); /* pdmp3.c:2526 */

/* CLI entry of the reference (pdmp3.c:2540): decode the NULL-terminated list of files
 * ("-" = stdin) to <file>.raw / stdout.  The OSS /dev/dsp writer is out of scope. */
void pdmp3(char * const *mp3s);

#ifdef __cplusplus
}
#endif
#endif
