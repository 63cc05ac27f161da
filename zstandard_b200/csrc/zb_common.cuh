// zb_common.cuh — shared definitions for the libzstdb200 device code.
//
// Everything here is __host__ __device__ so that the per-thread parsers can also be compiled by g++ into
// the host simulator used by the CPU-side tests (tests/hostsim/); the product only ever runs them on the GPU.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define ZB_HD __host__ __device__ __forceinline__
#define ZB_D __device__ __forceinline__
#else
#define ZB_HD inline
#define ZB_D inline
#endif

namespace zb {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;
typedef int16_t s16;

// ---- result / error encoding: identical to the reference (csharp/src/ZStdErrors.cs:61-100) ----
enum : u32 {
  ZE_GENERIC = 1, ZE_prefix_unknown = 10, ZE_frameParameter_unsupported = 14, ZE_frameParameter_windowTooLarge = 16,
  ZE_corruption_detected = 20, ZE_checksum_wrong = 22, ZE_dictionary_corrupted = 30, ZE_dictionary_wrong = 32,
  ZE_tableLog_tooLarge = 44, ZE_maxSymbolValue_tooLarge = 46, ZE_maxSymbolValue_tooSmall = 48,
  ZE_workSpace_tooSmall = 66, ZE_dstSize_tooSmall = 70, ZE_srcSize_wrong = 72, ZE_maxCode = 120
};
ZB_HD u32 zerr(u32 code) { return 0u - code; }
ZB_HD bool is_err(u32 r) { return r > zerr(ZE_maxCode); }

// ---- format constants (csharp/src/ZStd.cs:386-416,1387-1389; ZStdInternal.cs:109-209) ----
static const u32 MAGIC = 0xFD2FB528u, MAGIC_SKIP = 0x184D2A50u;
static const u32 BLOCKSIZE_MAX = 1u << 17;
static const u32 LONGNBSEQ = 0x7F00;
static const u32 MaxLL = 35, MaxML = 52, MaxOff = 31, LLFSELog = 9, MLFSELog = 9, OffFSELog = 8;
static const u32 HUF_LOG_MAX = 12;

// ---- per-item record produced by the parse kernel and refined by later stages ----
enum : u32 { FI_CHECKSUM = 1, FI_FCS_KNOWN = 2, FI_DONE = 4 /* result[] already final, later stages skip the item */,
              FI_NEED_XXH = 8 /* set by the execute stage: verify the content checksum */ };
struct FrameInfo {
  u32 flags;
  u32 body_off;      // offset within the item of the first block header of its data frame
  u64 fcs;           // frame content size when FI_FCS_KNOWN
  u64 window;        // window size (== fcs for single-segment frames)
  // entropy-stage error records (first failing block of each stage; 0xFFFFFFFF = none)
  u32 huf_err_block, huf_err_code;
  u32 seq_err_block, seq_err_code, seq_err_index;   // seq_err_index: sequences decoded before the failure; 0xFFFFFFFF = header error
  // set by the execute stage for the checksum stage
  u32 trailer_off;   // offset of the 4-byte checksum within the item
  u32 decoded;       // bytes produced
};

// array view with a stride (bank-interleaved per-lane arrays in shared memory)
template <class T> struct Strided {
  T* p; u32 s;
  ZB_HD T& operator[](u32 i) const { return p[i * s]; }
};

// Pulls a global-memory line towards the SM ahead of a dependent load (no-op on the host simulator).
ZB_HD void prefetch_line(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// ---- unaligned little-endian loads ----
ZB_HD u32 ld16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
ZB_HD u32 ld24(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16); }
ZB_HD u32 ld32(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
ZB_HD u64 ld64(const u8* p) { return (u64)ld32(p) | ((u64)ld32(p + 4) << 32); }

ZB_HD u32 highbit(u32 v) {   // floor(log2(v)), v != 0 (csharp/src/BitStream.cs:205-215)
#if defined(__CUDA_ARCH__)
  return 31 - __clz(v);
#else
  return 31 - __builtin_clz(v);
#endif
}
ZB_HD u32 fshr(u32 lo, u32 hi, u32 s) {   // low 32 bits of (hi:lo) >> (s & 31)
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, s);
#else
  s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}

ZB_HD u32 fshl(u32 lo, u32 hi, u32 s) {   // high 32 bits of (hi:lo) << (s & 31)
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, s);
#else
  s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}
ZB_HD u32 shr_c(u32 x, u32 n) {   // x >> n with n clamped to 32 (so n == 32 yields 0), as PTX shr.b32 defines it
#if defined(__CUDA_ARCH__)
  u32 r; asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n)); return r;
#else
  return n >= 32 ? 0 : x >> n;
#endif
}

// ---- backward bit cursor -------------------------------------------------------------------------
// A zstd bitstream of n bytes is read from its last byte towards its first (csharp/src/BitStream.cs:322-497).
// We address it by P = number of still-unread bits counted from the first byte; reading k bits takes stream
// bits [P-k, P) and leaves P-k.  Bits below 0 read as zero (the reference's container shifts zeros in) and a
// negative P is the reference's "overflow" status.  Loads are 4-byte aligned words relative to `words`.
struct BitCursor {
  const u32* words;   // 4-byte aligned address at or below the first stream byte
  i32 gofs;           // bit offset of stream bit 0 relative to words[0] (0, 8, 16 or 24)
  i32 P;              // unread bits
};

// initialise from (src, n); returns false if n == 0 or the final byte has no end mark (BitStream.cs:324-339, 369-372)
ZB_HD bool bc_init(BitCursor& c, const u8* src, u32 n) {
  if (n < 1) return false;
  u8 last = src[n - 1];
  if (last == 0) return false;
  uintptr_t a = (uintptr_t)src;
  c.words = (const u32*)(a & ~(uintptr_t)3);
  c.gofs = (i32)(a & 3) * 8;
  c.P = (i32)(n * 8) - (i32)(8 - highbit(last));
  return true;
}

// 64 stream bits ending at P, left aligned: bit 63 of the result is stream bit P-1.  Bits below stream bit 0
// are zero.  Requires P >= 0 for meaningful data; for P < 64 a byte-safe path is used so that nothing below
// `words` or before the stream is ever dereferenced.
ZB_HD u64 bc_window64(const BitCursor& c, i32 P) {
  if (P >= 64) {
    i32 g = c.gofs + P - 64;          // bit offset (>= 0) of the window's lowest bit
    const u32* w = c.words + (g >> 5);
    u32 s = (u32)g & 31;
    u32 w0 = w[0], w1 = w[1], w2 = s ? w[2] : 0;
    u32 lo = fshr(w0, w1, s), hi = fshr(w1, w2, s);
    return ((u64)hi << 32) | lo;
  }
  if (P <= 0) return 0;
  // tail: gather the ceil(P/8) remaining bytes
  const u8* b = (const u8*)c.words + (c.gofs >> 3);
  u64 v = 0; i32 nb = (P + 7) >> 3;
  for (i32 i = 0; i < nb; i++) v |= (u64)b[i] << (8 * i);
  return v << (64 - P);   // bits at and above P (end mark, padding) fall off the top
}

// ---- per-thread read-ahead of a backward bitstream through shared memory ---------------------------
// Every lane of the entropy kernels walks its own bitstream, so its loads can never coalesce with its
// neighbours', and L1 is sectored: read straight from global memory nearly every step waits on an L2/HBM
// sector.  BitRing keeps the 64 bytes around the cursor in a 16-word ring (ring[(word & 15) * stride], lane-
// interleaved in shared memory => conflict-free) and fetches the next lower 16-byte chunk one refill ahead
// into registers, so the dependent chain only ever sees shared-memory latency.
struct U4 { u32 x, y, z, w; };
ZB_HD U4 load_chunk16(const u8* p) {   // p is 16-byte aligned
#if defined(__CUDA_ARCH__)
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  return U4{v.x, v.y, v.z, v.w};
#else
  U4 v; memcpy(&v, p, 16); return v;
#endif
}

struct BitRing {
  u32* ring; u32 stride;
  const u8* base16;      // 16-byte aligned address at or below the first stream byte
  i32 gofs;              // bit offset of stream bit 0 relative to base16 (0..120)
  i32 lowChunk;          // ring holds chunks lowChunk .. lowChunk+3 (chunk k = bytes [16k, 16k+16) from base16)
  U4 pend;               // chunk lowChunk-1, requested at the previous refill (zeros when below the stream)
};

ZB_HD void ring_store(const BitRing& r, i32 chunk, const U4& v) {
  const u32 w = ((u32)chunk * 4) & 15;
  r.ring[(w + 0) * r.stride] = v.x; r.ring[(w + 1) * r.stride] = v.y;
  r.ring[(w + 2) * r.stride] = v.z; r.ring[(w + 3) * r.stride] = v.w;
}
ZB_HD U4 ring_fetch(const BitRing& r, i32 chunk) {
  if (chunk < 0) return U4{0, 0, 0, 0};
  return load_chunk16(r.base16 + (size_t)chunk * 16);
}
// Prepares the ring for a stream of `nbits` total bits (including the end-mark padding) starting at src.
// Only valid when the cursor starts at P >= 128.
ZB_HD void ring_init(BitRing& r, u32* ring, u32 stride, const u8* src, u32 nbytes) {
  const uintptr_t a = (uintptr_t)src;
  r.ring = ring; r.stride = stride;
  r.base16 = (const u8*)(a & ~(uintptr_t)15);
  r.gofs = (i32)(a & 15) * 8;
  const i32 ctop = (r.gofs + (i32)nbytes * 8 - 1) >> 7;
  r.lowChunk = ctop - 3;
  for (i32 k = 0; k < 4; k++) ring_store(r, r.lowChunk + k, ring_fetch(r, r.lowChunk + k));
  r.pend = ring_fetch(r, r.lowChunk - 1);
}
// 64 stream bits ending at P (P >= 64 and inside the ring's coverage): lo = bits [P-64, P-32), hi = [P-32, P)
ZB_HD void ring_window(const BitRing& r, i32 P, u32& lo, u32& hi) {
  const i32 g = r.gofs + P - 64;
  const u32 wi = (u32)(g >> 5), sh = (u32)g & 31;
  const u32 w0 = r.ring[(wi & 15) * r.stride], w1 = r.ring[((wi + 1) & 15) * r.stride], w2 = r.ring[((wi + 2) & 15) * r.stride];
  lo = fshr(w0, w1, sh); hi = fshr(w1, w2, sh);   // sh == 0 ignores w2
}
// Call after the cursor moved to Pn (by at most 127 bits since the last call): keeps one chunk of margin below
// the window and at most one chunk in flight.
ZB_HD void ring_advance(BitRing& r, i32 Pn) {
  const i32 wc = (r.gofs + Pn - 64) >> 7;
  if (wc <= r.lowChunk) {
    r.lowChunk -= 1;
    ring_store(r, r.lowChunk, r.pend);
    r.pend = ring_fetch(r, r.lowChunk - 1);
  }
}

}  // namespace zb
