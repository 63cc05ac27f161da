/* tools/corpus.c — deterministic synthetic corpora for tests and bench.
 * PRNG = the reference test suite's xorshift128+ (csharp/test/XorShift128Plus.cs:45-53).
 * A byte-identical pure-Python mirror lives in tools/corpus.py (checked by tests/test_corpus.py).
 * Test/bench infrastructure: not part of the product library. */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

typedef struct { uint64_t s0, s1; } rng_t;
static inline uint64_t rng_next(rng_t* r) {
  uint64_t x = r->s0, y = r->s1;
  r->s0 = y;
  x ^= x << 23;
  r->s1 = x ^ y ^ (x >> 17) ^ (y >> 26);
  return r->s1 + y;
}

/* bounded appender: never writes past cap, keeps counting */
typedef struct { uint8_t* p; uint64_t n, cap; } out_t;
static inline void put(out_t* o, uint8_t c) { if (o->n < o->cap) o->p[o->n] = c; o->n++; }
static void puts_(out_t* o, const char* s) { while (*s) put(o, (uint8_t)*s++); }
static void putu(out_t* o, uint64_t v, int width) { /* decimal, zero padded to width */
  char t[24]; int k = 0;
  do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (k < width) t[k++] = '0';
  while (k) put(o, (uint8_t)t[--k]);
}

static const char* LEVELS[16] = {"INFO","INFO","INFO","INFO","INFO","INFO","INFO","INFO",
                                 "DEBUG","DEBUG","DEBUG","DEBUG","WARN","WARN","ERROR","TRACE"};
static const char* COMPONENTS[8] = {"http.server","auth.session","cache.lru","net.pool",
                                    "sched.worker","rpc.client","runtime.gc","storage.wal"};
static const int MDAYS[12] = {31,28,31,30,31,30,31,31,30,31,30,31};

void corpus_log(uint8_t* dst, uint64_t n, uint64_t s0, uint64_t s1) {
  rng_t r = {s0, s1}; out_t o = {dst, 0, n};
  uint64_t t_ms = 0; /* since 2026-01-01T00:00:00.000Z, 365-day years */
  while (o.n < n) {
    t_ms += 1 + rng_next(&r) % 2000;
    uint64_t ms = t_ms % 1000, s = t_ms / 1000;
    uint64_t sec = s % 60, mi = (s / 60) % 60, h = (s / 3600) % 24, d = s / 86400;
    uint64_t year = 2026 + d / 365; d %= 365;
    int mo = 0; while (d >= (uint64_t)MDAYS[mo]) { d -= MDAYS[mo]; mo++; }
    putu(&o, year, 4); put(&o, '-'); putu(&o, mo + 1, 2); put(&o, '-'); putu(&o, d + 1, 2); put(&o, 'T');
    putu(&o, h, 2); put(&o, ':'); putu(&o, mi, 2); put(&o, ':'); putu(&o, sec, 2); put(&o, '.'); putu(&o, ms, 3);
    puts_(&o, "Z [");
    puts_(&o, LEVELS[rng_next(&r) % 16]); puts_(&o, "] ");
    puts_(&o, COMPONENTS[rng_next(&r) % 8]); puts_(&o, " - ");
    uint64_t a = rng_next(&r), b = rng_next(&r), c = rng_next(&r);
    switch (rng_next(&r) % 7) {
      case 0: puts_(&o, "request "); putu(&o, a % 1000000, 0); puts_(&o, " completed in "); putu(&o, b % 900, 0);
              puts_(&o, " ms status="); putu(&o, (c % 8) ? 200 : 500 + c % 4, 0); break;
      case 1: puts_(&o, "user "); putu(&o, a % 50000, 0); puts_(&o, " logged in from 10."); putu(&o, b % 256, 0);
              put(&o, '.'); putu(&o, (b >> 8) % 256, 0); put(&o, '.'); putu(&o, c % 256, 0); break;
      case 2: puts_(&o, "cache miss for key item:"); putu(&o, a % 100000, 0); puts_(&o, " shard="); putu(&o, b % 64, 0); break;
      case 3: puts_(&o, "connection "); putu(&o, a % 65536, 0); puts_(&o, " closed after "); putu(&o, b % 10000000, 0);
              puts_(&o, " bytes"); break;
      case 4: puts_(&o, "scheduled job "); putu(&o, a % 4096, 0); puts_(&o, " took "); putu(&o, b % 250000, 0);
              puts_(&o, " us queue_depth="); putu(&o, c % 128, 0); break;
      case 5: puts_(&o, "retry "); putu(&o, 1 + a % 5, 0); puts_(&o, "/5 for upstream svc-"); putu(&o, b % 32, 0); break;
      default: puts_(&o, "gc pause "); putu(&o, a % 120, 0); puts_(&o, " ms heap="); putu(&o, 512 + b % 7680, 0);
              puts_(&o, " MB"); break;
    }
    put(&o, '\n');
  }
}

static inline void le(out_t* o, uint64_t v, int bytes) { for (int i = 0; i < bytes; i++) put(o, (uint8_t)(v >> (8 * i))); }

void corpus_tick(uint8_t* dst, uint64_t n, uint64_t s0, uint64_t s1) {
  rng_t r = {s0, s1}; out_t o = {dst, 0, n};
  uint64_t ts = 1767225600000000ull; /* 2026-01-01 in us */
  int32_t price[64]; for (int i = 0; i < 64; i++) price[i] = 100000 + 1000 * i;
  while (o.n < n) {
    ts += 1 + rng_next(&r) % 5000;
    uint32_t sym = (uint32_t)(rng_next(&r) % 64);
    price[sym] += (int32_t)(rng_next(&r) % 11) - 5;
    uint32_t qty = 100 * (uint32_t)(1 + rng_next(&r) % 499);
    uint64_t v = rng_next(&r);
    le(&o, ts, 8); le(&o, sym, 4); le(&o, (uint32_t)price[sym], 4); le(&o, qty, 4);
    le(&o, v % 16, 2); le(&o, (v >> 8) & 1, 1); le(&o, 0, 1);
  }
}

void corpus_random(uint8_t* dst, uint64_t n, uint64_t s0, uint64_t s1) {
  rng_t r = {s0, s1}; out_t o = {dst, 0, n};
  while (o.n < n) le(&o, rng_next(&r), 8);
}

/* chunk i is log for i%10 in 0..4, tick for 5..7, random for 8..9; each kind is its own continuing stream */
void corpus_mixed(uint8_t* dst, uint64_t n, uint64_t chunk, uint64_t seed_shift) {
  uint64_t nchunks = (n + chunk - 1) / chunk, nl = 0, nt = 0, nr = 0;
  for (uint64_t i = 0; i < nchunks; i++) { uint64_t k = i % 10; if (k < 5) nl++; else if (k < 8) nt++; else nr++; }
  uint8_t* L = (uint8_t*)malloc(nl * chunk + 1); uint8_t* T = (uint8_t*)malloc(nt * chunk + 1); uint8_t* R = (uint8_t*)malloc(nr * chunk + 1);
  corpus_log(L, nl * chunk, 42 + seed_shift, 24); corpus_tick(T, nt * chunk, 43 + seed_shift, 25); corpus_random(R, nr * chunk, 44 + seed_shift, 26);
  uint64_t il = 0, it = 0, ir = 0;
  for (uint64_t i = 0; i < nchunks; i++) {
    uint64_t k = i % 10, off = i * chunk, len = off + chunk <= n ? chunk : n - off;
    const uint8_t* s = k < 5 ? L + (il++) * chunk : (k < 8 ? T + (it++) * chunk : R + (ir++) * chunk);
    memcpy(dst + off, s, len);
  }
  free(L); free(T); free(R);
}
