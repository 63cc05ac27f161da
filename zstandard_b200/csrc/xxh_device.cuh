// xxh_device.cuh — XXH64 (seed 0) over a byte range by a group of 4 lanes, one accumulator each
// (XxHash.cs:625-629 primes, :744-758 round, :1029-1093 update, :1105-1161 digest).
#pragma once
#include "zb_common.cuh"

namespace zb {

#define XP1 11400714785074694791ull
#define XP2 14029467366897019727ull
#define XP3 1609587929392839161ull
#define XP4 9650029242287828579ull
#define XP5 2870177450012600261ull
__device__ __forceinline__ u64 rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ u64 xxh_round(u64 acc, u64 in) { acc += in * XP2; acc = rotl64(acc, 31); return acc * XP1; }
__device__ __forceinline__ u64 xxh_merge(u64 acc, u64 v) { v = xxh_round(0, v); acc ^= v; return acc * XP1 + XP4; }
__device__ __forceinline__ u64 ldg64u(const u8* p) {   // unaligned 8-byte load from aligned words
  if ((((uintptr_t)p) & 7) == 0) return *(const u64*)p;
  const u32* w = (const u32*)((uintptr_t)p & ~(uintptr_t)3); u32 sh = ((u32)(uintptr_t)p & 3) * 8;
  u32 a0 = w[0], a1 = w[1], a2 = sh ? w[2] : 0;
  return ((u64)__funnelshift_r(a1, a2, sh) << 32) | __funnelshift_r(a0, a1, sh);
}

// DEEP: two batches of 16 loads alternate, so that 16 loads are in flight WHILE the other batch feeds the accumulator
// (for the few long frames of the block-parallel path, where a lane's chain over 32 768 stripes is the kernel's time)
template <bool DEEP>
static __device__ u64 xxh64_group(const u8* p, u64 len, u32 sub, unsigned gmask, u32 lead) {
  u64 h;
  const u8* tail = p;
  if (len >= 32) {
    u64 v = sub == 0 ? XP1 + XP2 : (sub == 1 ? XP2 : (sub == 2 ? 0 : 0 - XP1));
    const u64 stripes = len / 32;
    const u8* q = p + 8 * sub;
    u64 i = 0;
    if (DEEP && stripes >= 32) {
      u64 x[16], y[16];
#pragma unroll
      for (int k = 0; k < 16; k++) x[k] = ldg64u(q + 32 * k);
      q += 512; i = 16;
      for (; i + 32 <= stripes; i += 32) {
#pragma unroll
        for (int k = 0; k < 16; k++) y[k] = ldg64u(q + 32 * k);
#pragma unroll
        for (int k = 0; k < 16; k++) v = xxh_round(v, x[k]);
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = ldg64u(q + 512 + 32 * k);
#pragma unroll
        for (int k = 0; k < 16; k++) v = xxh_round(v, y[k]);
        q += 1024;
      }
#pragma unroll
      for (int k = 0; k < 16; k++) v = xxh_round(v, x[k]);
    }
    // the accumulator chain is serial; keep 16 independent loads in flight ahead of it
    for (; i + 16 <= stripes; i += 16) {
      u64 x[16];
#pragma unroll
      for (int k = 0; k < 16; k++) x[k] = ldg64u(q + 32 * k);
#pragma unroll
      for (int k = 0; k < 16; k++) v = xxh_round(v, x[k]);
      q += 512;
    }
    for (; i < stripes; i++) { v = xxh_round(v, ldg64u(q)); q += 32; }
    u64 v1 = __shfl_sync(gmask, v, lead), v2 = __shfl_sync(gmask, v, lead + 1), v3 = __shfl_sync(gmask, v, lead + 2), v4 = __shfl_sync(gmask, v, lead + 3);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
    tail = p + stripes * 32;
  } else h = XP5;   // seed(0) + P5
  h += len;
  const u8* end = p + len;
  while (tail + 8 <= end) { h ^= xxh_round(0, ldg64u(tail)); h = rotl64(h, 27) * XP1 + XP4; tail += 8; }
  if (tail + 4 <= end) { h ^= (u64)ld32(tail) * XP1; h = rotl64(h, 23) * XP2 + XP3; tail += 4; }
  while (tail < end) { h ^= (*tail) * XP5; h = rotl64(h, 11) * XP1; tail++; }
  h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
  return h;
}

// the digest from the four accumulators after len / 32 stripes of p[0..len) (XxHash.cs:1105-1161)
static __device__ u64 xxh64_finish(u64 v1, u64 v2, u64 v3, u64 v4, const u8* p, u64 len) {
  u64 h; const u8* tail = p;
  if (len >= 32) {
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
    tail = p + (len / 32) * 32;
  } else h = XP5;
  h += len;
  const u8* end = p + len;
  while (tail + 8 <= end) { h ^= xxh_round(0, ldg64u(tail)); h = rotl64(h, 27) * XP1 + XP4; tail += 8; }
  if (tail + 4 <= end) { h ^= (u64)ld32(tail) * XP1; h = rotl64(h, 23) * XP2 + XP3; tail += 4; }
  while (tail < end) { h ^= (*tail) * XP5; h = rotl64(h, 11) * XP1; tail++; }
  h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
  return h;
}

}  // namespace zb
