/* oracle/ref_bench.c -- TEST/BENCH INFRASTRUCTURE ONLY.
 * Times the unmodified reference decoder (linked from oracle/_ref/libpdmp3_ref.so) on an
 * in-memory stream with the loop of the reference CLI pdmp3() (pdmp3.c:2564-2584):
 * pdmp3_read() into a 16 KiB buffer, on NEED_MORE feed the next 4096 bytes.  The reference
 * keeps decoder state in function-static arrays (pdmp3.c:1755,1983) so N-core runs are N
 * forked processes over N contiguous byte shards split at frame offsets given by the caller.
 *
 * usage: ref_bench <stream file> <nprocs> <offsets file: (nprocs+1) little-endian u64 byte offsets>
 * prints one line: frames_total seconds_wall
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/wait.h>

typedef struct pdmp3_handle pdmp3_handle;
pdmp3_handle *pdmp3_new(const char *, int *);
void pdmp3_delete(pdmp3_handle *);
int pdmp3_open_feed(pdmp3_handle *);
int pdmp3_feed(pdmp3_handle *, const unsigned char *, size_t);
int pdmp3_read(pdmp3_handle *, unsigned char *, size_t, size_t *);

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static uint64_t decode_shard(const unsigned char *p, size_t n)
{
  unsigned char out[16384];
  pdmp3_handle *id = pdmp3_new(NULL, NULL);
  memset(id, 0, 65536 < 39912 ? 0 : 39912);   /* sizeof(pdmp3_handle) = 39912 on x86-64; zero stale state */
  pdmp3_open_feed(id);
  size_t fed = 0, done; uint64_t bytes = 0; int res;
  while ((res = pdmp3_read(id, out, sizeof out, &done)) != -1) {
    bytes += done;
    if (res == -10) {                       /* PDMP3_NEED_MORE */
      size_t k = n - fed; if (k > 4096) k = 4096;
      if (!k) break;
      pdmp3_feed(id, p + fed, k); fed += k;
    }
  }
  pdmp3_delete(id);
  return bytes / 4608;
}

int main(int argc, char **argv)
{
  if (argc < 4) { fprintf(stderr, "usage: %s stream nprocs offsets\n", argv[0]); return 2; }
  int np = atoi(argv[2]);
  FILE *f = fopen(argv[1], "rb"); if (!f) { perror(argv[1]); return 1; }
  fseek(f, 0, SEEK_END); size_t n = ftell(f); fseek(f, 0, SEEK_SET);
  unsigned char *buf = malloc(n); if (fread(buf, 1, n, f) != n) return 1; fclose(f);
  uint64_t *off = malloc(8 * (np + 1));
  f = fopen(argv[3], "rb"); if (!f || fread(off, 8, np + 1, f) != (size_t)(np + 1)) return 1; fclose(f);
  uint64_t *frames = mmap(NULL, 8 * np, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  double t0 = now();
  for (int i = 0; i < np; i++) {
    if (fork() == 0) { frames[i] = decode_shard(buf + off[i], off[i + 1] - off[i]); _exit(0); }
  }
  while (wait(NULL) > 0) {}
  double t1 = now();
  uint64_t tot = 0; for (int i = 0; i < np; i++) tot += frames[i];
  printf("%llu %.6f\n", (unsigned long long)tot, t1 - t0);
  return 0;
}
