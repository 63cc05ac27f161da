/* tools/zstd_mt.c — all-core libzstd 1.5.5 batch decode for bench.py's second CPU baseline (SURVEY.md §8d ii).
 * libzstd ships here as a shared object without headers: the two prototypes used are declared by hand and the
 * library is opened at run time.  Test/bench infrastructure only. */
#include <dlfcn.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>

typedef size_t (*decompress_fn)(void*, size_t, const void*, size_t);
typedef unsigned (*iserror_fn)(size_t);

struct job { const uint8_t* blob; const uint64_t* off; uint8_t* out; uint64_t n, chunk, total; int t, threads; decompress_fn dec; iserror_fn err; int bad; };

static void* work(void* p) {
  struct job* j = (struct job*)p;
  for (uint64_t i = (uint64_t)j->t; i < j->n; i += (uint64_t)j->threads) {
    uint64_t cap = j->chunk; if (i * j->chunk + cap > j->total) cap = j->total - i * j->chunk;
    size_t r = j->dec(j->out + i * j->chunk, cap, j->blob + j->off[i], j->off[i + 1] - j->off[i]);
    if (j->err(r) || r != cap) j->bad = 1;
  }
  return 0;
}

/* frame i = blob[off[i] .. off[i+1]) decodes to out[i*chunk ..]; returns 0 when every frame decoded to its full size */
int zmt_decompress(const uint8_t* blob, const uint64_t* off, uint64_t n, uint8_t* out, uint64_t chunk, uint64_t total, int threads) {
  void* h = dlopen("libzstd.so.1", RTLD_NOW);
  if (!h) return -1;
  decompress_fn dec = (decompress_fn)dlsym(h, "ZSTD_decompress"); iserror_fn err = (iserror_fn)dlsym(h, "ZSTD_isError");
  if (!dec || !err || threads < 1 || threads > 1024) return -2;
  pthread_t th[1024]; struct job jb[1024];
  for (int t = 0; t < threads; t++) {
    struct job j = {blob, off, out, n, chunk, total, t, threads, dec, err, 0}; jb[t] = j;
    pthread_create(&th[t], 0, work, &jb[t]);
  }
  int bad = 0;
  for (int t = 0; t < threads; t++) { pthread_join(th[t], 0); bad |= jb[t].bad; }
  return bad;
}
