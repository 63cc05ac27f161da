// decode_kernels.cu — sm_100a kernels of the batched zstd frame decoder.
//
// Pipeline per batch (all on one stream unless the host overlaps sub-batches):
//   k_parse : thread / item     frame header, skippable frames, early verdicts      (ZStdDecompress.cs:2096-2160, 389-499)
//   k_huf   : 4 lanes / frame   Huffman literals -> literal scratch                  (HufDecompress.cs:117-358)
//   k_seq   : thread / frame    FSE tables + sequence bitstream -> 8-byte records    (ZStdDecompress.cs:958-1180, 1443-1608)
//   k_exec  : warp / frame      literal copy + match copy into dst, raw/RLE blocks   (ZStdDecompress.cs:1212-1352, 1599-1605, 2033-2091)
//   k_xxh   : 4 lanes / frame   XXH64 content checksum                               (XxHash.cs:896-1161)
#include "zb_decode.cuh"
#include "decode_kernels.cuh"
#include "xxh_device.cuh"
#ifdef ZB_EXEC_DEBUG
#include <stdio.h>
#endif

namespace zb {

// -----------------------------------------------------------------------------------------------
// scratch addressing shared by the stages (see DESIGN.md "HBM layout")
// -----------------------------------------------------------------------------------------------
// A data frame of item f writes at dst_off[f] + out_base and may use dst_cap[f] - out_base bytes (out_base > 0 only
// for the second and later data frames of one item); its scratch regions follow from that position, so they stay
// inside the item's share of the arenas.
__device__ __forceinline__ u64 frame_dst_off(const DecodeArgs& a, u32 f, const FrameInfo& fi) { return a.dst_off[f] + fi.out_base; }
__device__ __forceinline__ u32 frame_cap(const DecodeArgs& a, u32 f, const FrameInfo& fi) { return a.dst_cap[f] - fi.out_base; }
__device__ __forceinline__ u8* lit_region(const DecodeArgs& a, u32 f, const FrameInfo& fi) { return a.lit_arena + (frame_dst_off(a, f, fi) & ~15ull) + 64ull * (a.item_base + f); }
__device__ __forceinline__ u64 lit_capacity(u32 cap) { return (u64)cap + 40; }
__device__ __forceinline__ SeqRec* seq_region(const DecodeArgs& a, u32 f, const FrameInfo& fi) { return a.seq_arena + 2 * (frame_dst_off(a, f, fi) / 3) + 32ull * (a.item_base + f); }

// =================================================================================================
// k_parse
// =================================================================================================
__global__ void __launch_bounds__(128) k_parse(DecodeArgs a) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  FrameInfo fi; u32 r = 0; u32 start = 0, outBase = 0;
  if (a.pass) {
    // later passes only touch items whose previous data frame decoded cleanly and is followed by another one
    const FrameInfo prev = a.info[i];
    if ((prev.flags & FI_DONE) || prev.next_off == 0 || is_err(a.result[i])) { a.info[i].flags = FI_DONE; return; }
    start = prev.next_off; outBase = prev.out_base + prev.decoded;
  }
  bool go = parse_item(a.src_base + a.src_off[i], a.src_size[i], fi, &r, start, outBase);
  a.info[i] = fi;
  if (!go) a.result[i] = r;
}

// =================================================================================================
// k_huf : one warp per CTA, 8 frames per warp, lanes 4f..4f+3 own the 4 streams of frame f
// =================================================================================================
struct HufSmem {
  __align__(16) u16 table[8][1 << HUF_TABLE_LOG];   // also lent as HufFseScratch while a frame's weights are decoded
  HufBuildWk wk[8];
  __align__(16) u32 ring[32][ZB_RING_WORDS + 4];   // per-lane (= per-stream) bitstream read-ahead (BitRing); a folded
                                                   // log-12 table keeps its 256-byte side table in the group's first ring
};
static_assert(sizeof(HufFseScratch) <= sizeof(u16) << HUF_TABLE_LOG, "FSE scratch must fit the decode table it borrows");
static_assert((ZB_RING_WORDS + 4) * 4 >= 256, "side table must fit one lane's ring");

__global__ void __launch_bounds__(32) k_huf(DecodeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HufSmem& sm = *reinterpret_cast<HufSmem*>(smem_raw);
  const u32 lane = threadIdx.x, sub = lane & 3, slot = lane >> 2;
  const u32 f = blockIdx.x * 8 + slot;
  const unsigned gmask = 0xFu << (slot * 4);
  bool active = f < a.n;
  FrameInfo fi;
  if (active) { fi = a.info[f]; if (fi.flags & FI_DONE) active = false; }
  if (!active) return;   // whole 4-lane group leaves together; group syncs below use gmask
  const u8* src = a.src_base + a.src_off[f]; const u32 size = a.src_size[f];
  u8* lit = lit_region(a, f, fi); const u64 litCap = lit_capacity(frame_cap(a, f, fi));
  u16* dt = sm.table[slot]; HufBuildWk& wk = sm.wk[slot];
  u8* const sideMem = (u8*)&sm.ring[slot * 4][0]; const u8* side = nullptr;
  u8* const slotMem = (u8*)&sm.ring[slot * 4 + 1][0];   // per-symbol rank within its weight: dead once the table is filled
  u32 pos = fi.body_off, blk = 0; u64 litRun = 0;
  u32 tableLog = 0; bool haveTable = false;
  u32 errBlock = 0xFFFFFFFFu, errCode = 0;
  while (true) {
    BlockHdr bh;
    if (read_block_hdr(src + pos, size - pos, bh)) break;
    pos += 3;
    if (bh.type == 2) {
      const u8* bp = src + pos; u32 bsz = bh.csize;
      if (bsz >= BLOCKSIZE_MAX) break;
      LitHdr lh; bool needs;
      if (read_lit_hdr(bp, bsz, lh, &needs)) break;
      if (lh.type >= 2) {
        if (lh.type == 3 && !haveTable) break;                                     // dictionary_corrupted, reported by k_exec
        bool ok = true;
        const u8* body = bp + lh.lhSize; u32 bodySize = lh.litCSize;
        const bool dry = litRun + lh.litSize + 3 > litCap;                         // cannot be stored: validate only (HUF_DRY)
        if (lh.type == 2) {
          if (!lh.single && (lh.litSize == 0 || bodySize == 0)) ok = false;       // HufDecompress.cs:1211-1212
          u32 hdr = 0, nbSym = 0, tl = 0;
          if (ok) {
            u32 e = 0;
            if (sub == 0) {
              e = huf_read_weights(body, bodySize, wk, *reinterpret_cast<HufFseScratch*>(dt), slotMem, &hdr, &tl, &nbSym);
              if (!e && hdr >= bodySize) e = ZE_srcSize_wrong;                     // HufDecompress.cs:1193
            }
            __syncwarp(gmask);
            e = __shfl_sync(gmask, e, slot * 4); hdr = __shfl_sync(gmask, hdr, slot * 4);
            tl = __shfl_sync(gmask, tl, slot * 4); nbSym = __shfl_sync(gmask, nbSym, slot * 4);
            if (e) ok = false;
          }
          if (ok) {
            __syncwarp(gmask);                                                     // lane 0's scratch use of dt is over
            huf_fill_table(dt, sideMem, wk, slotMem, tl, nbSym, sub, 4);
            __syncwarp(gmask);
            tableLog = tl; haveTable = true; side = tl > HUF_TABLE_LOG ? sideMem : nullptr;
            body += hdr; bodySize -= hdr;
          }
        }
        if (ok) {
          bool good = true;
          if (lh.single) {
            if (sub == 0) good = dry ? huf_check_stream(body, bodySize, lh.litSize, dt, tableLog, side)
                                     : huf_decode_stream(body, bodySize, lit + litRun, lh.litSize, dt, tableLog, &sm.ring[lane][0], side);   // HufDecompress.cs:247-264
          } else {
            HufStream st;
            good = huf_split4(body, bodySize, lh.litSize, sub, st);
            if (good) good = dry ? huf_check_stream(st.src, st.len, st.count, dt, tableLog, side)
                                 : huf_decode_stream(st.src, st.len, lit + litRun + st.outOfs, st.count, dt, tableLog, &sm.ring[lane][0], side);
          }
          unsigned okmask = __ballot_sync(gmask, good);
          if ((okmask & gmask) != gmask) ok = false;
        }
        if (!ok || dry) { errBlock = blk; errCode = ok ? HUF_DRY : ZE_corruption_detected; break; }
        litRun += lh.litSize;
      }
    }
    pos += bh.csize; blk++;
    if (bh.last) break;
  }
  if (sub == 0 && errBlock != 0xFFFFFFFFu) { a.info[f].huf_err_block = errBlock; a.info[f].huf_err_code = errCode; }
}

// =================================================================================================
// k_seq : one warp per CTA, one frame per lane, tables bank-interleaved across lanes
// =================================================================================================
struct SeqSmem {
  u16 ll[512][32];       // lane-interleaved: cell[state][lane]
  u16 ml[512][32];
  u16 of[256][32];
  u16 defLL[64], defOF[32], defML[64];
  u32 llInfo[36], mlInfo[53];
  s16 norm[53][32];      // per-lane scratch of the table builder, lane-interleaved like the tables
  u16 next[53][32];
  __align__(16) u32 ring[32][ZB_RING_WORDS + 4];   // per-lane bitstream read-ahead (BitRing), skewed by 4 banks per lane
};

__global__ void __launch_bounds__(32) k_seq(DecodeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SeqSmem& sm = *reinterpret_cast<SeqSmem*>(smem_raw);
  const u32 lane = threadIdx.x;
  // predefined tables + info LUTs, built once per CTA
  if (lane < 3) {
    s16* norm = &sm.norm[0][lane]; u16* next = &sm.next[0][lane];
    Strided<s16> nv{norm, 32}; Strided<u16> sn{next, 32};
    if (lane == 0) { for (int i = 0; i < 36; i++) nv[i] = kLLnorm[i]; build_seq_table(sm.defLL, 1, nv, 35, 6, sn); }
    if (lane == 1) { for (int i = 0; i < 29; i++) nv[i] = kOFnorm[i]; build_seq_table(sm.defOF, 1, nv, 28, 5, sn); }
    if (lane == 2) { for (int i = 0; i < 53; i++) nv[i] = kMLnorm[i]; build_seq_table(sm.defML, 1, nv, 52, 6, sn); }
  }
  for (u32 i = lane; i < 36; i += 32) sm.llInfo[i] = ll_info(i);
  for (u32 i = lane; i < 53; i += 32) sm.mlInfo[i] = ml_info(i);
  __syncwarp();
  const u32 f = blockIdx.x * 32 + lane;
  if (f >= a.n) return;
  FrameInfo fi = a.info[f];
  if (fi.flags & FI_DONE) return;
  SeqTableSet T;
  T.space[KIND_LL] = &sm.ll[0][lane]; T.space[KIND_ML] = &sm.ml[0][lane]; T.space[KIND_OF] = &sm.of[0][lane]; T.stride = 32;
  T.defs[KIND_LL] = sm.defLL; T.defs[KIND_OF] = sm.defOF; T.defs[KIND_ML] = sm.defML;
  SeqFrameOut res;
  seq_decode_frame(a.src_base + a.src_off[f], a.src_size[f], fi.body_off, fi.window, T, seq_region(a, f, fi), seq_capacity(frame_cap(a, f, fi)), res, sm.llInfo, sm.mlInfo,
                   Strided<s16>{&sm.norm[0][lane], 32}, Strided<u16>{&sm.next[0][lane], 32}, &sm.ring[lane][0]);
  if (res.err_block != 0xFFFFFFFFu) {
    a.info[f].seq_err_block = res.err_block; a.info[f].seq_err_code = res.err_code; a.info[f].seq_err_index = res.err_index;
  }
}

// =================================================================================================
// k_exec : one CTA per frame, the output window assembled in shared memory
// =================================================================================================
// Sequence execution is the reference's ExecSequence loop (:1265-1352, :1582-1605) re-expressed for a CTA whose
// shared memory holds the segment [segLo, segLo + EX_SEG) of the frame's output (all of it for frames <= 64 KiB):
//  * match sources are read from shared memory (29-cycle LDS instead of an L2/HBM round trip per dependency hop;
//    nothing of the window is ever re-read from DRAM); a finished segment leaves with one TMA bulk store
//    (cp.async.bulk.global.shared::cta), 16-byte aligned on both sides because the window is laid out with the
//    destination's alignment; the head of each block's literals arrives by a TMA bulk load.
//  * records are self-describing (zb_decode.cuh: output position, literal position, offset, lengths), so the
//    EXW execute warps take groups of 32 records round-robin with no prefix sums and no positional chain: lane =
//    sequence.  A lane copies its literals (no dependency) and then its match in straight-line pieces of <= 16
//    bytes (6 LDS.32, 5 SHF, aligned word stores, byte stores only in the first and last word).
//  * dependencies are exact and out of order: a match may run once every sequence whose output overlaps its
//    source has finished.  Which sequences those are follows from a position -> sequence map (one entry per 32
//    output bytes, written by the sequences themselves when their records are loaded); whether they have finished
//    from one word of done bits per group, published by the group's warp after every round.  On log text this
//    needs ~100 dependent rounds per 64 KiB frame (the depth of the dependency graph) where an in-order
//    watermark needed ~330.
// Frames larger than the window are executed segment by segment: copies are clipped to the segment, sources
// below it are read back from global memory (bulk stores of earlier segments are complete by then).
#define FULLMASK 0xFFFFFFFFu
#define EX_SEG 65536u                 // bytes of output held in shared memory
#define EX_PAD 32u                    // slack before the window (source loads reach below a copy's first source byte)
#define EXW 8u                        // execute warps per CTA
#define EX_LANE_MAX 64u               // longer runs are copied by the whole warp
#define EX_ERR_NONE 0xFFFFFFFFu
#define EX_LIT_STAGE 4096u            // bytes of a block's literals staged in shared memory
#define EX_CHUNK_LOG 5u               // position -> sequence map: one entry per 32 output bytes
#define EX_MAP_N (EX_SEG >> EX_CHUNK_LOG)
#define EX_EPOCH_GROUPS 512u          // groups per pass: sequence numbers within a pass fit 14 bits (0xFFFF = no entry yet)

struct ExecSmem {
  __align__(16) u8 win[EX_PAD + 16 + EX_SEG + 32];
  __align__(16) u8 lit[EX_LIT_STAGE + 32];
  __align__(16) u16 map[EX_MAP_N + 8];     // map[c] = pass-relative number of the sequence that covers position segLo + 32 c
  __align__(16) u32 done[EX_EPOCH_GROUPS]; // done[g - gFirst]: bit l = sequence 32 (g - gFirst) + l has finished
  __align__(8) u64 litBar;                 // the staged literals have landed
  u32 errKey;     // min over failing records of (record index << 8 | error code)
  u32 nextFirst;  // first group with output beyond the current segment
#ifdef ZB_EXEC_DEBUG
  unsigned long long dbg[EXW][8];
#endif
};

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ u32 lds16(u32 a) { u32 v; asm volatile("ld.volatile.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ u32 lds8(u32 a) { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(u32 a, u32 lo, u32 hi) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(lo), "r"(hi) : "memory"); }
__device__ __forceinline__ void sts16(u32 a, u32 v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(u32 a, u32 v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ u32 ld_volatile_s(u32 a) { u32 v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void st_volatile_s(u32 a, u32 v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// The window: shared address of position p is s0 + (p - segLo); s0 is congruent to the destination address of
// segLo modulo 16, so that 16-byte granules of the window are 16-byte granules of the output.
struct Win {
  u32 s0, segLo, segHi;
  u8* g;        // global address of output position 0
};
__device__ __forceinline__ u32 waddr(const Win& w, u32 pos) { return w.s0 + (pos - w.segLo); }

// ---- the hot copy: n <= 16 bytes within shared memory, straight line ----
// bytes [A, A + n) <- [A - off, A - off + n), 1 <= n <= 16, source and destination do not overlap (off >= n).  The
// destination is cut into its (at most 5) aligned words; each is rebuilt from two aligned source words.  Only the
// first and the last word can be partial; their bytes go out as u8 / u16 stores chosen by the low address bits.
__device__ __forceinline__ void small_copy_s(u32 A, u32 n, u32 off) {
  const u32 W0 = A & ~3u, blo = A & 3u, last = (blo + n + 3 - 4) >> 2;      // last = index of the last destination word (0..4)
  const u32 Sa = W0 - off, sh = (Sa & 3u) * 8, Sb = Sa & ~3u;
  const u32 x0 = lds32(Sb), x1 = lds32(Sb + 4), x2 = lds32(Sb + 8), x3 = lds32(Sb + 12), x4 = lds32(Sb + 16), x5 = lds32(Sb + 20);
  const u32 v0 = __funnelshift_r(x0, x1, sh), v1 = __funnelshift_r(x1, x2, sh), v2 = __funnelshift_r(x2, x3, sh),
            v3 = __funnelshift_r(x3, x4, sh), v4 = __funnelshift_r(x4, x5, sh);
  if (last == 0) {
    if (n == 4) sts32(W0, v0);
    else { const u32 t = v0 >> (8 * blo); sts8(A, t); if (n >= 2) sts8(A + 1, t >> 8); if (n >= 3) sts8(A + 2, t >> 16); }
    return;
  }
  if (blo == 0) sts32(W0, v0);
  else { if (blo & 1) sts8(A, v0 >> (8 * blo)); if (blo <= 2) sts16(W0 + 2, v0 >> 16); }
  if (last > 1) sts32(W0 + 4, v1);
  if (last > 2) sts32(W0 + 8, v2);
  if (last > 3) sts32(W0 + 12, v3);
  const u32 vl = last == 1 ? v1 : (last == 2 ? v2 : (last == 3 ? v3 : v4));
  const u32 Wl = W0 + 4 * last, r = (A + n) & 3u;
  if (r == 0) sts32(Wl, vl);
  else { if (r & 2) sts16(Wl, vl); if (r & 1) sts8(Wl + (r & 2), vl >> (8 * (r & 2))); }
}
// a run of any length by one lane: pieces of <= 16 bytes, every piece after the first starts on a word boundary.
// Byte-serial meaning for self-overlapping matches (off < len): the run is periodic, so a piece may be taken from any
// multiple m of the period behind it — m grows with the bytes already written until pieces reach 16 bytes.
__device__ __forceinline__ void lane_copy_s(u32 A, u32 len, u32 off) {
  if (off >= len) {
    u32 n = 16 - (A & 3u); n = len < n ? len : n;
    small_copy_s(A, n, off); A += n; len -= n;
    while (len) { n = len < 16 ? len : 16; small_copy_s(A, n, off); A += n; len -= n; }
  } else {
    u32 m = off, done = 0;
    while (done < len) {
      while (m <= done) m += off;
      u32 n = len - done; n = n < m ? n : m; n = n < 16 ? n : 16;
      small_copy_s(A + done, n, m); done += n;
    }
  }
}
// the same run shared by several threads: piece p = first, first + stride, ...  (off >= len)
__device__ __forceinline__ void strided_copy_s(u32 A, u32 len, u32 off, u32 first, u32 stride) {
  u32 n0 = 16 - (A & 3u); n0 = len < n0 ? len : n0;
  const u32 nP = 1 + ((len - n0 + 15) >> 4);
  for (u32 p = first; p < nP; p += stride) {
    const u32 a = p ? A + n0 + 16 * (p - 1) : A, left = A + len - a;
    small_copy_s(a, p ? (left < 16 ? left : 16) : n0, off);
  }
}
// one match by the whole warp, any offset / length relation: a self-overlapping match is built by doubling — each step
// copies the periodic run produced so far (a multiple of the period) behind itself
__device__ __noinline__ void warp_match_s(u32 A, u32 len, u32 off, u32 lane) {
  if (off >= len) { strided_copy_s(A, len, off, lane, 32); return; }
  u32 done = 0;
  while (done < len) {
    const u32 span = done + off, n = len - done < span ? len - done : span;
    strided_copy_s(A + done, n, span, lane, 32);
    __syncwarp();
    done += n;
  }
}

// ---- the cold copies (raw / RLE blocks, last literals, literals beyond the staged head, sources in earlier segments):
//      8-byte destination-aligned units, source in global memory or a fill value ----
// bytes [b0, b1) of the word at shared address a (0 <= b0, b1 <= 4; empty when b0 >= b1)
__device__ __forceinline__ void store_word(u32 a, u32 v, u32 b0, u32 b1) {
  if (b0 == 0 && b1 == 4) { sts32(a, v); return; }
  if (b0 == 0 && b1 > 0) sts8(a, v);
  if (b0 <= 1 && b1 > 1) sts8(a + 1, v >> 8);
  if (b0 <= 2 && b1 > 2) sts8(a + 2, v >> 16);
  if (b0 <= 3 && b1 > 3) sts8(a + 3, v >> 24);
}
// bytes [blo, bhi) of the 8-byte unit at shared address a (8-byte aligned)
__device__ __forceinline__ void store_unit(u32 a, u32 lo, u32 hi, u32 blo, u32 bhi) {
  if (blo == 0 && bhi == 8) { sts64(a, lo, hi); return; }
  store_word(a, lo, blo, bhi < 4 ? bhi : 4);
  store_word(a + 4, hi, blo > 4 ? blo - 4 : 0, bhi > 4 ? bhi - 4 : 0);
}
// the bytes at [p, p + 8) of global memory of which only [blo, bhi) are needed; aligned words that hold none of the
// needed bytes are not touched (p may point below the buffer for a head unit).  CG: bypass L1 (used for output
// bytes written earlier by bulk stores, which L1 does not see).
template <bool CG>
__device__ __forceinline__ void load8_g(const u8* p, u32 blo, u32 bhi, u32& lo, u32& hi) {
  const u32 sb = (u32)(uintptr_t)p & 3;
  const u32* w = reinterpret_cast<const u32*>(p - sb);
  const i32 a0 = -(i32)sb;                       // unit-byte index of word 0's first byte
  u32 w0 = 0, w1 = 0, w2 = 0;
  if ((i32)bhi > a0 && (i32)blo < a0 + 4) w0 = CG ? __ldcg(w) : __ldg(w);
  if ((i32)bhi > a0 + 4 && (i32)blo < a0 + 8) w1 = CG ? __ldcg(w + 1) : __ldg(w + 1);
  if ((i32)bhi > a0 + 8 && (i32)blo < a0 + 12) w2 = CG ? __ldcg(w + 2) : __ldg(w + 2);
  lo = __funnelshift_r(w0, w1, sb * 8); hi = __funnelshift_r(w1, w2, sb * 8);
}
// bytes [A, A + len) <- src[0 .. len): unit `first`, `first + stride`, ... (one lane: 0 / 1, warp: lane / 32, CTA: tid / threads)
template <bool CG>
__device__ __forceinline__ void strided_copy_g(u32 A, u32 len, const u8* src, u32 first, u32 stride) {
  const u32 E = A + len, D0 = A & ~7u;
  for (u32 D = D0 + 8 * first; D < E; D += 8 * stride) {
    const u32 blo = D < A ? A - D : 0, left = E - D, bhi = left < 8 ? left : 8;
    u32 lo, hi; load8_g<CG>(src + (ptrdiff_t)(i32)(D - A), blo, bhi, lo, hi);   // negative for the head unit
    store_unit(D, lo, hi, blo, bhi);
  }
}
__device__ __forceinline__ void strided_fill(u32 A, u32 len, u32 v4, u32 first, u32 stride) {
  const u32 E = A + len;
  for (u32 D = (A & ~7u) + 8 * first; D < E; D += 8 * stride) {
    const u32 blo = D < A ? A - D : 0, left = E - D, bhi = left < 8 ? left : 8;
    store_unit(D, v4, v4, blo, bhi);
  }
}

// Writes the window's positions [from, to) to global memory: the 16-byte aligned middle as one TMA bulk store, the
// ragged ends (neighbouring frames may share those granules) by byte stores.  Called by the whole CTA; on return the
// bytes are in global memory and the window may be reused.
__device__ __forceinline__ void flush_window(const Win& w, u32 from, u32 to, u32 tid) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of this thread -> visible to the async proxy
  __syncthreads();
  if (to > from) {
    u8* g = w.g + from; const u32 s = waddr(w, from), n = to - from;
    u32 head = (16u - ((u32)(uintptr_t)g & 15u)) & 15u; if (head > n) head = n;
    const u32 body = (n - head) & ~15u, tail = n - head - body;
    if (tid < head) g[tid] = (u8)lds8(s + tid);
    else if (tid >= 32 && tid - 32 < tail) g[head + body + tid - 32] = (u8)lds8(s + head + body + tid - 32);
    if (tid == 64 && body) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + head), "r"(s + head), "r"(body) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
    }
  }
  __syncthreads();
}

// Appends n bytes at output position op through the window, rolling it over (flush + next segment) whenever it fills.
// kind 0: bytes from global memory `src`; kind 1: the byte value `fill`.  Called by the whole CTA.
__device__ __forceinline__ void cta_append(Win& w, u32& flushed, u32 op, u32 n, int kind, const u8* src, u32 fill, u32 tid) {
  while (n) {
    if (op == w.segHi) { flush_window(w, flushed, op, tid); flushed = op; w.segLo = op; w.segHi = op + EX_SEG; }
    const u32 room = w.segHi - op, take = n < room ? n : room;
    if (kind == 0) strided_copy_g<false>(waddr(w, op), take, src, tid, EXEC_THREADS);
    else strided_fill(waddr(w, op), take, fill * 0x01010101u, tid, EXEC_THREADS);
    op += take; n -= take; src += take;
  }
}

// ---- mbarrier / TMA primitives ----
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(u32 bar, u32 parity) {
  u32 ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// TMA 1-D bulk load global -> shared (16-byte aligned on both sides, size a multiple of 16), completion on `bar`
__device__ __forceinline__ void bulk_load(u32 dstS, const void* src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dstS), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

#ifdef ZB_EXEC_DEBUG
#define DBG_T0 const long long t0_ = clock64()
#define DBG_ACC(x) x += clock64() - t0_
#define DBG_INC(x) x++
#else
#define DBG_T0
#define DBG_ACC(x)
#define DBG_INC(x)
#endif

// One pass of the execute warps over the groups [gFirst, gLimit) of a block's records for the current window
// segment.  recs = the block's first record (after its header), blockBase = output position of the block's first
// byte, passBase = every output byte below it is final (the window's start or where the pass's first sequence
// begins), litS / litStaged = shared address and size of the staged head of the block's literals (phase litParity
// of sm.litBar).
__device__ __forceinline__ void exec_groups(ExecSmem& sm, const Win& w, const SeqRec* __restrict__ recs, u32 nRecs, u32 blockBase, u32 cap,
                                            const u8* lit, u32 litSize, u32 litS, u32 litStaged, bool isRle, u32 rleByte, bool dry,
                                            u32 gFirst, u32 gLimit, u32 passBase, u32 litParity, u32 warp, u32 lane) {
  const u32 sE = smem_u32(&sm.errKey), sMap = smem_u32(&sm.map[0]), sDone = smem_u32(&sm.done[0]);
  const uint4* __restrict__ rv = reinterpret_cast<const uint4*>(recs);
  const uint4 zero = make_uint4(0, 0, 0, 0);
  u32 g = gFirst + warp;
  uint4 cur = (g < gLimit && g * 32 + lane < nRecs) ? __ldg(rv + g * 32 + lane) : zero;
  bool litReady = litStaged == 0;
#ifdef ZB_EXEC_DEBUG
  unsigned long long dGroups = 0, dRounds = 0, dLit = 0, dPoll = 0, dCopy = 0, dTot = clock64(), dPolls = 0, dSetup = 0;
#endif
  for (; g < gLimit; g += EXW) {
    const u32 gn = g + EXW;
    const uint4 nxt = (gn < gLimit && gn * 32 + lane < nRecs) ? __ldg(rv + gn * 32 + lane) : zero;   // in flight while this group runs
    const u32 idx = g * 32 + lane; const bool valid = idx < nRecs;
    const u32 ll = cur.w & 0x1FFFF, ml = (cur.w >> 17) | (((cur.y >> 18) & 3) << 15), lpos = cur.y & 0x3FFFF, off = cur.z;
    // checks in the reference's order (:1278, :1279, :1290-1294)
    const u64 start64 = (u64)blockBase + cur.x, end64 = start64 + ll + ml;
    const bool e1 = valid && end64 > cap;
    const bool e2 = valid && lpos + ll > litSize;
    const bool e3 = valid && (u64)off > start64 + ll;
    const unsigned bad = __ballot_sync(FULLMASK, e1 | e2 | e3);
    if (bad) {
      const u32 first = (u32)__ffs(bad) - 1;
      const u32 code = __shfl_sync(FULLMASK, e1 ? (u32)ZE_dstSize_tooSmall : (u32)ZE_corruption_detected, first);
      if (lane == 0) atomicMin(&sm.errKey, ((g * 32 + first) << 8) | code);
      return;                                                   // the block fails; polling warps see errKey and leave
    }
    if (ld_volatile_s(sE) != EX_ERR_NONE) return;
    const u32 nValid = nRecs - g * 32 < 32 ? nRecs - g * 32 : 32;
    const u32 start = (u32)start64, mpos = start + ll, end = mpos + ml;
    const u32 G0 = __shfl_sync(FULLMASK, start, 0), Gend = __shfl_sync(FULLMASK, end, nValid - 1);
    if (Gend > w.segHi && lane == 0) atomicMin(&sm.nextFirst, g);   // first group of the next segment's pass
    if (G0 >= w.segHi) return;                                  // this and all later groups belong to later segments
    if (dry) { cur = nxt; continue; }                           // literals were not stored: checks only
    const u32 gr = g - gFirst, seqNo = gr * 32 + lane;          // pass-relative group / sequence number
    // ---- position -> sequence map: one entry for every 32-byte chunk whose first byte this sequence covers (the
    //      sequence that covers passBase also takes the chunk passBase lies in) ----
    {
      DBG_T0;
      const u32 lo = start > passBase ? start : passBase, hi = end < w.segHi ? end : w.segHi;
      const bool any = valid && hi > lo;
      const u32 c0 = start <= passBase ? (lo - w.segLo) >> EX_CHUNK_LOG : (lo - w.segLo + 31) >> EX_CHUNK_LOG;
      const u32 c1 = any ? (hi - 1 - w.segLo) >> EX_CHUNK_LOG : 0;
      const u32 cnt = (any && c1 >= c0) ? c1 - c0 + 1 : 0;
      if (cnt && cnt <= 4) { sts16(sMap + 2 * c0, seqNo); if (cnt > 1) sts16(sMap + 2 * c0 + 2, seqNo); if (cnt > 2) sts16(sMap + 2 * c0 + 4, seqNo); if (cnt > 3) sts16(sMap + 2 * c0 + 6, seqNo); }
      unsigned big = __ballot_sync(FULLMASK, cnt > 4);
      while (big) {
        const u32 j = (u32)__ffs(big) - 1; big &= big - 1;
        const u32 cj = __shfl_sync(FULLMASK, c0, j), nj = __shfl_sync(FULLMASK, cnt, j), sj = __shfl_sync(FULLMASK, seqNo, j);
        for (u32 c = lane; c < nj; c += 32) sts16(sMap + 2 * (cj + c), sj);
      }
      DBG_ACC(dSetup);
    }
    // ---- literals (no dependency): clipped to the segment ----
    {
      DBG_T0;
      const u32 p0 = start > w.segLo ? start : w.segLo, p1 = mpos < w.segHi ? mpos : w.segHi;
      const u32 n = (valid && p1 > p0) ? p1 - p0 : 0;
      const u32 A = waddr(w, p0), lp = lpos + (p0 - start);
      const bool staged = lp + n <= litStaged;
      if (!litReady) {
        while (!mbar_try_wait(smem_u32(&sm.litBar), litParity)) { if (ld_volatile_s(sE) != EX_ERR_NONE) return; }
        litReady = true;
      }
      if (n && n <= EX_LANE_MAX && !isRle && staged) lane_copy_s(A, n, A - (litS + lp));
      const unsigned slow = __ballot_sync(FULLMASK, n && (n > EX_LANE_MAX || isRle || !staged));
      if (slow) {
        const u8* ls = lit + lp;
        if (n && n <= EX_LANE_MAX && (isRle || !staged)) { if (isRle) strided_fill(A, n, rleByte * 0x01010101u, 0, 1); else strided_copy_g<false>(A, n, ls, 0, 1); }
        unsigned big = __ballot_sync(FULLMASK, n > EX_LANE_MAX);
        while (big) {
          const u32 j = (u32)__ffs(big) - 1; big &= big - 1;
          const u32 Aj = __shfl_sync(FULLMASK, A, j), nj = __shfl_sync(FULLMASK, n, j);
          const u64 lj = __shfl_sync(FULLMASK, (u64)(uintptr_t)ls, j);
          if (isRle) strided_fill(Aj, nj, rleByte * 0x01010101u, lane, 32); else strided_copy_g<false>(Aj, nj, (const u8*)(uintptr_t)lj, lane, 32);
        }
      }
      DBG_ACC(dLit);
    }
    // ---- matches: clip to the segment; the part whose source lies below the segment comes from global memory ----
    u32 q0 = mpos > w.segLo ? mpos : w.segLo; const u32 q1 = end < w.segHi ? end : w.segHi;
    bool hasM = valid && q1 > q0 && off != 0;                   // offset 0 (corrupted input only) copies bytes onto themselves
    if (w.segLo) {                                              // only frames larger than the window
      const bool below = hasM && q0 - off < w.segLo;            // q0 >= off (check e3), so no wrap
      unsigned gm = __ballot_sync(FULLMASK, below);
      while (gm) {
        const u32 j = (u32)__ffs(gm) - 1; gm &= gm - 1;
        const u32 qj = __shfl_sync(FULLMASK, q0, j), ej = __shfl_sync(FULLMASK, q1, j), oj = __shfl_sync(FULLMASK, off, j);
        const u32 split = oj < ej - w.segLo ? w.segLo + oj : ej;   // positions below `split` have their source below the segment
        strided_copy_g<true>(waddr(w, qj), split - qj, w.g + (qj - oj), lane, 32);
      }
      if (below) { q0 = off < q1 - w.segLo ? w.segLo + off : q1; hasM = q1 > q0; }
    }
    __syncwarp();
    // ---- window matches, each as soon as the sequences its source overlaps have finished ----
    const u32 srcLo = q0 - off;                                 // >= segLo here
    const u32 srcHi = (q1 - off < q0) ? q1 - off : q0;          // source bytes outside the match's own output
    const u32 mlen = q1 - q0, A = waddr(w, q0);
    unsigned pend = __ballot_sync(FULLMASK, hasM);
    unsigned mine = ~pend;                                      // finished sequences of this group (lanes without a window match: at once)
    if (lane == 0) st_volatile_s(sDone + 4 * gr, mine);
    // dependency range [depA, depB) in pass-relative sequence numbers, resolved from the map when its entries exist
    const bool needDep = hasM && srcHi > passBase;
    const u32 needLo = srcLo > passBase ? srcLo : passBase;
    const u32 mA = sMap + 2 * ((needLo - w.segLo) >> EX_CHUNK_LOG), mB = sMap + 2 * (((srcHi - 1 - w.segLo) >> EX_CHUNK_LOG) + 1);
    u32 depA = 0, depB = 0; bool resolved = !needDep;
    while (pend) {
      DBG_T0;
      bool ready = (pend >> lane) & 1;
      if (ready && !resolved) {
        const u32 a = lds16(mA);
        if (a == 0xFFFFu) ready = false;
        else {
          const u32 nx = lds16(mB);
          depA = a; depB = nx + 1 < seqNo ? nx + 1 : seqNo;     // never beyond the sequence itself (nx = 0xFFFF: not mapped yet)
          resolved = nx != 0xFFFFu;                             // a later poll may find a tighter range
        }
      }
      if (ready && depB > depA) {
        const u32 wa = depA >> 5, wb = (depB - 1) >> 5;
        const u32 ma = 0xFFFFFFFFu << (depA & 31), mb = 0xFFFFFFFFu >> (31 - ((depB - 1) & 31));
        if (wa == wb) { const u32 m = ma & mb; ready = (ld_volatile_s(sDone + 4 * wa) & m) == m; }
        else {
          ready = (ld_volatile_s(sDone + 4 * wa) & ma) == ma && (ld_volatile_s(sDone + 4 * wb) & mb) == mb;
          for (u32 x = wa + 1; ready && x < wb; x++) ready = ld_volatile_s(sDone + 4 * x) == 0xFFFFFFFFu;
        }
      }
      const unsigned R = __ballot_sync(FULLMASK, ready);
      DBG_ACC(dPoll); DBG_INC(dPolls);
      if (!R) { if (ld_volatile_s(sE) != EX_ERR_NONE) return; continue; }
      {
        DBG_T0; DBG_INC(dRounds);
        const bool longM = ready && mlen > EX_LANE_MAX;
        if (ready && !longM) lane_copy_s(A, mlen, off);
        unsigned big = __ballot_sync(FULLMASK, longM);
        while (big) {
          const u32 j = (u32)__ffs(big) - 1; big &= big - 1;
          warp_match_s(__shfl_sync(FULLMASK, A, j), __shfl_sync(FULLMASK, mlen, j), __shfl_sync(FULLMASK, off, j), lane);
        }
        __syncwarp();
        pend &= ~R; mine |= R;
        if (lane == 0) st_volatile_s(sDone + 4 * gr, mine);
        DBG_ACC(dCopy);
      }
    }
    DBG_INC(dGroups);
    cur = nxt;
  }
#ifdef ZB_EXEC_DEBUG
  if (lane == 0) { unsigned long long* d = sm.dbg[warp]; d[0] += dGroups; d[1] += dRounds; d[2] += dLit; d[3] += dPoll; d[4] += dCopy; d[5] += clock64() - dTot; d[6] += dPolls; d[7] += dSetup; }
#endif
}

__global__ void __launch_bounds__(EXEC_THREADS, 3) k_exec(DecodeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ExecSmem& sm = *reinterpret_cast<ExecSmem*>(smem_raw);
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 f = blockIdx.x;
  FrameInfo fi = a.info[f];
  if (fi.flags & FI_DONE) return;
  const u8* src = a.src_base + a.src_off[f]; const u32 size = a.src_size[f];
  u8* dst = a.dst_base + frame_dst_off(a, f, fi); const u32 cap = frame_cap(a, f, fi);
  const u8* litScratch = lit_region(a, f, fi);
  const SeqRec* recs = seq_region(a, f, fi); const u32 recCap = (u32)seq_capacity(cap);
  Win w; w.g = dst; w.segLo = 0; w.segHi = EX_SEG; w.s0 = smem_u32(sm.win) + EX_PAD + ((u32)(uintptr_t)dst & 15u);
  u32 flushed = 0;                                                  // output positions below are in global memory
  if (tid == 0) { sm.errKey = EX_ERR_NONE; mbar_init(smem_u32(&sm.litBar), 1); }
#ifdef ZB_EXEC_DEBUG
  if (tid < EXW * 8) (&sm.dbg[0][0])[tid] = 0;
  const long long tFrame = clock64();
#endif
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  u32 pos = fi.body_off, blk = 0, op = 0, litRun = 0, recRun = 0;   // op <= cap < 2^32 throughout: checks are written against cap - op
  bool litEntropy = false, dry = false; u32 err = 0;
  u32 litGen = 0;                                                   // literal stagings so far (phase of sm.litBar)
  while (true) {
    BlockHdr bh;
    err = read_block_hdr(src + pos, size - pos, bh);
    if (err) break;
    pos += 3;
    if (bh.type == 0) {                                                            // raw, :662-667
      if (bh.csize > cap - op) { err = ZE_dstSize_tooSmall; break; }
      cta_append(w, flushed, op, bh.csize, 0, src + pos, 0, tid); op += bh.csize;
    } else if (bh.type == 1) {                                                     // RLE, :1945-1950
      if (bh.orig > cap - op) { err = ZE_dstSize_tooSmall; break; }
      cta_append(w, flushed, op, bh.orig, 1, nullptr, src[pos], tid); op += bh.orig;
    } else {
      const u8* bp = src + pos; const u32 bsz = bh.csize;
      if (bsz >= BLOCKSIZE_MAX) { err = ZE_srcSize_wrong; break; }                 // :1880
      LitHdr lh; bool needs;
      u32 e = read_lit_hdr(bp, bsz, lh, &needs);
      if (needs && !litEntropy) { err = ZE_dictionary_corrupted; break; }          // :696-697
      if (e) { err = e; break; }
      const u8* lit; u32 rleByte = 0; bool isRle = false;
      if (lh.type >= 2) {
        if (fi.huf_err_block == blk) {
          if (fi.huf_err_code != HUF_DRY) { err = fi.huf_err_code; break; }        // :742 (all Huffman failures -> corruption_detected)
          dry = true;                                                              // literals were not stored: checks only, the block must fail
        }
        litEntropy = true; lit = litScratch + litRun; litRun += lh.litSize;
      } else if (lh.type == 0) lit = bp + lh.lhSize;
      else { lit = nullptr; isRle = true; rleByte = bp[lh.lhSize]; }
      const u32 litSize = lh.litSize;
      const u8* sp = bp + lh.consumed; const u32 ssz = bsz - lh.consumed;
      u32 nbSeq, modes, hdr;
      e = read_seq_count(sp, ssz, &nbSeq, &modes, &hdr);
      if (e) { err = e; break; }
      if (fi.seq_err_block == blk && fi.seq_err_index == 0xFFFFFFFFu) { err = fi.seq_err_code; break; }
      u32 litPos = 0;
      if (nbSeq) {
        // the block's header record: how many records follow and what they produce (zb_decode.cuh)
        u32 nRecs = 0, outBytes = 0, litBytes = 0;
        if (recRun < recCap) { const uint4 h = __ldg(reinterpret_cast<const uint4*>(recs + recRun)); nRecs = h.x; outBytes = h.y; litBytes = h.z; }
        const SeqRec* r = recs + recRun + 1;
        const u32 blockBase = op, nGroups = (nRecs + 31) >> 5;
        // the head of the block's literals is staged in shared memory by one TMA bulk load; lanes whose literals lie
        // beyond it read global memory
        const u32 litAl = isRle ? 0 : (u32)(uintptr_t)lit & 15u;
        u32 litStaged = 0;
        if (!isRle && !dry && nRecs) { const u32 want = litAl + litSize; litStaged = (want < EX_LIT_STAGE ? want : EX_LIT_STAGE); }
        const u32 litStageBytes = (litStaged + 15u) & ~15u;                          // reads <= 15 bytes past the literals, inside their 16-byte granule
        litStaged = litStaged > litAl ? litStaged - litAl : 0;                     // literal bytes [0, litStaged) are in shared memory
        u32 gFirst = 0;
        bool first = true;
        while (gFirst < nGroups) {
          const u32 gLimit = nGroups - gFirst < EX_EPOCH_GROUPS ? nGroups : gFirst + EX_EPOCH_GROUPS;
          // every output byte below passBase is final: the bytes of earlier blocks, segments and passes
          const u32 firstStart = blockBase + __ldg(&r[gFirst * 32].x);             // < 2^32: the pass's first sequence passed its checks or fails them now
          const u32 passBase = firstStart > w.segLo ? firstStart : w.segLo;
          __syncthreads();
          for (u32 i = tid; i < (EX_MAP_N + 8) / 2; i += EXEC_THREADS) reinterpret_cast<u32*>(sm.map)[i] = 0xFFFFFFFFu;
          for (u32 i = tid; i < EX_EPOCH_GROUPS; i += EXEC_THREADS) sm.done[i] = 0;
          if (tid == 0) sm.nextFirst = 0xFFFFFFFFu;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic accesses to sm.lit before the TMA write
          __syncthreads();
          if (tid == 0 && first && litStaged) {
            mbar_expect_tx(smem_u32(&sm.litBar), litStageBytes);
            bulk_load(smem_u32(sm.lit), lit - litAl, litStageBytes, smem_u32(&sm.litBar));
          }
          first = false;
          exec_groups(sm, w, r, nRecs, blockBase, cap, lit, litSize, smem_u32(sm.lit) + litAl, litStaged, isRle, rleByte, dry, gFirst, gLimit, passBase, litGen & 1, warp, lane);
          __syncthreads();
          if (sm.errKey != EX_ERR_NONE) break;
          const u32 nf = sm.nextFirst;
          if (nf < gLimit) {                                                       // the segment is full: flush it, go on from the group that crossed its end
            flush_window(w, flushed, w.segHi, tid); flushed = w.segHi; w.segLo = w.segHi; w.segHi += EX_SEG;
            gFirst = nf;
          } else gFirst = gLimit;
        }
        if (sm.errKey != EX_ERR_NONE) { err = sm.errKey & 0xFF; break; }
        if (litStaged) litGen++;
        recRun += 1 + nRecs;
        if (fi.seq_err_block == blk) { err = fi.seq_err_code; break; }             // :1594 after the decodable prefix
        op += outBytes; litPos = litBytes;
      }
      // last literals (:1599-1605)
      const u32 lastLL = litSize - litPos;
      if (lastLL > cap - op || dry) { err = ZE_dstSize_tooSmall; break; }
      cta_append(w, flushed, op, lastLL, isRle ? 1 : 0, isRle ? nullptr : lit + litPos, rleByte, tid);
      op += lastLL;
    }
    pos += bh.csize; blk++;
    if (bh.last) break;
  }
#ifdef ZB_EXEC_DEBUG
  __syncthreads();
  const long long tBeforeFlush = clock64();
#endif
  // what the window still holds leaves even when the frame failed: the reference, too, leaves its partial output behind
  flush_window(w, flushed, op > flushed ? op : flushed, tid);
#ifdef ZB_EXEC_DEBUG
  if ((f == 1000 || f == 3000) && tid < EXW) {
    const unsigned long long* d = sm.dbg[tid];
    printf("k_exec stats frame %u warp %u: groups %llu rounds %llu polls %llu | cycles setup %llu lit %llu poll %llu copy %llu exec_groups %llu | frame %lld flush %lld\n", f, tid, d[0], d[1], d[6], d[7], d[2], d[3], d[4], d[5],
           (long long)(clock64() - tFrame), (long long)(clock64() - tBeforeFlush));
  }
#endif
  // ---- frame epilogue (:2069-2085) and the multi-frame loop tail (:2111-2157) ----
  bool needXxh = false; u32 trailer = 0;
  if (!err) {
    if ((fi.flags & FI_FCS_KNOWN) && op != fi.fcs) err = ZE_corruption_detected;
    else if (fi.flags & FI_CHECKSUM) {
      if (size - pos < 4) err = ZE_checksum_wrong; else { trailer = pos; pos += 4; needXxh = true; }
    }
  }
  u32 tailErr = 0, nextOff = 0;
  if (!err) {
    while (true) {
      u32 rem = size - pos;
      if (rem < 5) { if (rem) tailErr = ZE_srcSize_wrong; break; }
      u32 magic = ld32(src + pos);
      if (magic == MAGIC) { nextOff = pos; break; }                                // another data frame: next pass (:2111-2153)
      if ((magic & 0xFFFFFFF0u) != MAGIC_SKIP) { tailErr = ZE_prefix_unknown; break; }
      if (rem < 8) { tailErr = ZE_srcSize_wrong; break; }
      u32 skip = ld32(src + pos + 4) + 8u;
      if (rem < skip) { tailErr = ZE_srcSize_wrong; break; }
      pos += skip;
    }
  }
  if (tid == 0) {
    u32 res = err ? zerr(err) : (tailErr ? zerr(tailErr) : fi.out_base + (u32)op);
    a.result[f] = res;                                                             // provisional while next_off != 0
    a.info[f].trailer_off = trailer; a.info[f].decoded = (u32)op; a.info[f].next_off = nextOff;
    a.info[f].flags = fi.flags | (needXxh ? FI_NEED_XXH : 0);
    if (nextOff && a.more) atomicAdd(a.more, 1u);
  }
}

// =================================================================================================
// k_xxh : XXH64(seed 0) of the decoded bytes, 4 lanes per frame (one accumulator each)
// =================================================================================================
__global__ void __launch_bounds__(128) k_xxh(DecodeArgs a) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 f = t >> 2, sub = t & 3, lane = threadIdx.x & 31;
  const unsigned gmask = 0xFu << (lane & ~3u);
  if (f >= a.n) return;
  FrameInfo fi = a.info[f];
  if (!(fi.flags & FI_NEED_XXH)) return;
  const u8* dst = a.dst_base + frame_dst_off(a, f, fi);
  u64 h = xxh64_group(dst, fi.decoded, sub, gmask, lane & ~3u);
  if (sub == 0) {
    const u8* src = a.src_base + a.src_off[f];
    if ((u32)h != ld32(src + fi.trailer_off)) a.result[f] = zerr(ZE_checksum_wrong);   // :2078-2082
  }
}


// =================================================================================================
// launch
// =================================================================================================
size_t decode_lit_arena_bytes(u64 max_dst_bytes, u64 max_items) { return (size_t)(max_dst_bytes + 64 * (max_items + 2) + 256); }
size_t decode_seq_arena_bytes(u64 max_dst_bytes, u64 max_items) { return (size_t)((2 * (max_dst_bytes / 3) + 32 * (max_items + 2) + 64) * sizeof(SeqRec)); }

cudaError_t decode_configure() {
  cudaError_t e = cudaFuncSetAttribute(k_huf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HufSmem));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_exec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExecSmem));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_exec, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   // 3 CTAs of 74 KB per SM
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SeqSmem));
}

const char* const kDecodeKernelNames[DECODE_KERNELS] = {"k_parse", "k_huf", "k_seq", "k_exec", "k_xxh"};

// The pipeline in two halves (entropy stages, execute stages).  They were split to overlap slices on different
// streams; that measured slower (the execute warps starve the lone FSE warps of an SM), so decode_launch runs both
// on one stream.
cudaError_t decode_launch_entropy(const DecodeArgs& a, cudaStream_t st, int* launches, cudaEvent_t* marks) {
  if (a.n == 0) return cudaSuccess;
  if (marks) cudaEventRecord(marks[0], st);
  k_parse<<<(a.n + 127) / 128, 128, 0, st>>>(a);
  if (marks) cudaEventRecord(marks[1], st);
  k_huf<<<(a.n + 7) / 8, 32, sizeof(HufSmem), st>>>(a);
  if (marks) cudaEventRecord(marks[2], st);
  k_seq<<<(a.n + 31) / 32, 32, sizeof(SeqSmem), st>>>(a);
  if (marks) cudaEventRecord(marks[3], st);
  if (launches) *launches += 3;
  return cudaGetLastError();
}
cudaError_t decode_launch_exec(const DecodeArgs& a, cudaStream_t st, int* launches, cudaEvent_t* marks) {
  if (a.n == 0) return cudaSuccess;
  k_exec<<<a.n, EXEC_THREADS, sizeof(ExecSmem), st>>>(a);
  if (marks) cudaEventRecord(marks[4], st);
  k_xxh<<<(a.n * 4 + 127) / 128, 128, 0, st>>>(a);
  if (marks) cudaEventRecord(marks[5], st);
  if (launches) *launches += 2;
  return cudaGetLastError();
}
cudaError_t decode_launch(const DecodeArgs& a, cudaStream_t st, int* launches, cudaEvent_t* marks) {
  cudaError_t e = decode_launch_entropy(a, st, launches, marks);
  if (e != cudaSuccess) return e;
  return decode_launch_exec(a, st, launches, marks);
}

}  // namespace zb
