// zb_decode.cuh — per-thread bodies of the entropy stages of the GPU decoder.
//
//   seq_decode_frame : one thread walks one frame's blocks, builds the LL/OF/ML tables in (bank-interleaved)
//                      shared memory and turns every sequences bitstream into 8-byte records in HBM scratch.
//   huf_*            : one thread per Huffman stream (4 per frame) decodes literals into HBM scratch from a
//                      shared-memory decode table built by the frame's first lane.
//
// They are __host__ __device__ so that tests/hostsim can replay exactly this code on the CPU.
#pragma once
#include "zb_format.cuh"

namespace zb {

// 8-byte sequence record consumed by the execute stage.
//   x = offset (>= 1 for a match; 0 marks the end of a block's records)
//   y = litLength | matchLength << 16       (both < 65536; longer ones are split into several records,
//                                            a literals-only piece has matchLength 0 and offset 1)
struct SeqRec { u32 x, y; };

// Capacity rule shared with the host side: records for a frame whose output capacity is `cap` bytes.
// Every record with a match yields >= 3 bytes and each block adds one terminator.
ZB_HD u64 seq_capacity(u64 cap) { return 2 * (cap / 3) + 24; }

struct SeqTableSet {
  u32* space[3];            // lane-private cells for LL, OF, ML (index with *stride)
  u32 stride;
  const u32* defs[3];       // predefined tables, stride 1
  const u32* cur[3]; u32 curStride[3]; u32 log[3];
};

struct SeqFrameOut {
  u32 err_block, err_code, err_index;
};

ZB_HD void seq_emit(SeqRec* out, u64& n, u64 cap, bool& overflow, u32 off, u32 ll, u32 ml) {
  while (ll > 65535) { if (n < cap) { out[n].x = 1; out[n].y = 65535; } else overflow = true; n++; ll -= 65535; }
  while (ml > 65535) { if (n < cap) { out[n].x = off; out[n].y = ll | (65535u << 16); } else overflow = true; n++; ll = 0; ml -= 65535; }
  if (n < cap) { out[n].x = off; out[n].y = ll | (ml << 16); } else overflow = true;
  n++;
}

// Walks the frame at item `src` (size bytes, first block header at body_off) and decodes every compressed
// block's sequences.  Stops silently at structural errors that the execute stage will report itself from the
// same headers; records entropy-level failures in `res`.
// llBase/mlBase: base-value tables (any address space readable by this thread).
ZB_HD void seq_decode_frame(const u8* src, u32 size, u32 body_off, SeqTableSet& T, SeqRec* out, u64 cap, SeqFrameOut& res,
                            const u32* llBase, const u32* mlBase) {
  res.err_block = 0xFFFFFFFFu; res.err_code = 0; res.err_index = 0;
  u32 pos = body_off, blk = 0;
  u32 rep0 = 1, rep1 = 4, rep2 = 8;                      // ZStdInternal.cs:111, ZStdDecompress.cs:2492
  bool haveRepeat = false;
  u64 n = 0; bool overflow = false;
  s16 norm[53]; u16 symbolNext[53];
  for (int k = 0; k < 3; k++) { T.cur[k] = T.space[k]; T.curStride[k] = T.stride; T.log[k] = 0; }
  while (true) {
    BlockHdr bh;
    if (read_block_hdr(src + pos, size - pos, bh)) return;
    pos += 3;
    if (bh.type == 2) {
      const u8* bp = src + pos; u32 bsz = bh.csize;
      if (bsz >= BLOCKSIZE_MAX) return;                                            // :1880
      LitHdr lh; bool needs;
      if (read_lit_hdr(bp, bsz, lh, &needs)) return;
      const u8* sp = bp + lh.consumed; u32 ssz = bsz - lh.consumed;
      u32 nbSeq, modes, hdr;
      u32 e = read_seq_count(sp, ssz, &nbSeq, &modes, &hdr);
      if (e) { res.err_block = blk; res.err_code = e; res.err_index = 0xFFFFFFFFu; return; }
      if (nbSeq) {
        // tables in the reference's order LL, OF, ML (:1149-1176); any failure is corruption_detected
        const int kinds[3] = {KIND_LL, KIND_OF, KIND_ML};
        const u32 modeOf[3] = {(modes >> 4) & 3, (modes >> 2) & 3, modes & 3};
        for (int k = 0; k < 3 && !e; k++) {
          int kind = kinds[k]; u32 used, lg = T.log[kind]; bool isDef = false;
          u32 m = modeOf[k];
          e = read_seq_table(m, kind, sp + hdr, ssz - hdr, T.space[kind], T.stride, &lg, &isDef, haveRepeat, &used, norm, symbolNext);
          if (!e) {
            hdr += used;
            if (m != 3) {
              T.log[kind] = lg;
              if (isDef) { T.cur[kind] = T.defs[kind]; T.curStride[kind] = 1; }
              else { T.cur[kind] = T.space[kind]; T.curStride[kind] = T.stride; }
            }
          }
        }
        if (e) { res.err_block = blk; res.err_code = ZE_corruption_detected; res.err_index = 0xFFFFFFFFu; return; }
        haveRepeat = true;                                                         // fseEntropy = 1 (:1575)
        // ---- bitstream ----
        BitCursor c;
        u32 decoded = 0; bool bad = false;
        if (!bc_init(c, sp + hdr, ssz - hdr)) bad = true;                          // :1577 -> corruption_detected
        if (!bad) {
          i32 P = c.P;
          const u32 *tLL = T.cur[KIND_LL], *tOF = T.cur[KIND_OF], *tML = T.cur[KIND_ML];
          const u32 sLLs = T.curStride[KIND_LL], sOFs = T.curStride[KIND_OF], sMLs = T.curStride[KIND_ML];
          u32 stLL, stOF, stML;
          { u64 w = bc_window64(c, P); u32 lg = T.log[KIND_LL]; stLL = lg ? (u32)(w >> (64 - lg)) : 0; w <<= lg; P -= (i32)lg;
            lg = T.log[KIND_OF]; stOF = lg ? (u32)(w >> (64 - lg)) : 0; w <<= lg; P -= (i32)lg;
            lg = T.log[KIND_ML]; stML = lg ? (u32)(w >> (64 - lg)) : 0; P -= (i32)lg; }   // :1578-1580 (<= 26 bits)
          for (u32 i = 0; i < nbSeq; i++) {
            if (P < 0) { bad = true; break; }                                      // loop test :1582 (overflow)
            u32 cLL = tLL[stLL * sLLs], cOF = tOF[stOF * sOFs], cML = tML[stML * sMLs];
            u32 llBits = (cLL >> 14) & 31, mlBits = (cML >> 14) & 31, ofBits = (cOF >> 14) & 31;
            u32 llSym = cLL >> 19, mlSym = cML >> 19;
            u64 w = bc_window64(c, P);
            u32 ofv = ofBits ? (u32)(w >> (64 - ofBits)) : 0; w <<= ofBits;
            u32 mlv = mlBits ? (u32)(w >> (64 - mlBits)) : 0; w <<= mlBits;
            u32 llv = llBits ? (u32)(w >> (64 - llBits)) : 0;
            P -= (i32)(ofBits + mlBits + llBits);
            if (P < 0) { bad = true; break; }     // values came from beyond the stream start (over-read)
            u32 offset = ofBits ? of_base(ofBits) + ofv : 0;                       // :1487-1507
            if (ofBits <= 1) {                                                     // :1509-1524
              offset += (llSym == 0);
              if (offset) {
                u32 temp = offset == 3 ? rep0 - 1 : (offset == 1 ? rep1 : rep2);   // prevOffset[offset], :1514
                temp += !temp;
                if (offset != 1) rep2 = rep1;
                rep1 = rep0; rep0 = offset = temp;
              } else offset = rep0;
            } else { rep2 = rep1; rep1 = rep0; rep0 = offset; }                    // :1527-1529
            u32 ml = mlBase[mlSym] + mlv, ll = llBase[llSym] + llv;
            seq_emit(out, n, cap, overflow, offset, ll, ml);
            decoded++;
            // state update LL, ML, OF (:1547-1550); past the last sequence these bits do not exist
            u32 nLL = (cLL >> 10) & 15, nML = (cML >> 10) & 15, nOF = (cOF >> 10) & 15;
            u64 w2 = bc_window64(c, P);
            stLL = (cLL & 0x3FF) + (nLL ? (u32)(w2 >> (64 - nLL)) : 0); w2 <<= nLL;
            stML = (cML & 0x3FF) + (nML ? (u32)(w2 >> (64 - nML)) : 0); w2 <<= nML;
            stOF = (cOF & 0x3FF) + (nOF ? (u32)(w2 >> (64 - nOF)) : 0);
            P -= (i32)(nLL + nML + nOF);
          }
        }
        // terminator
        if (n < cap) { out[n].x = 0; out[n].y = 0; } else overflow = true;
        n++;
        if (overflow) { res.err_block = blk; res.err_code = ZE_corruption_detected; res.err_index = 0xFFFFFFFFu; return; }   // records unusable
        if (bad) { res.err_block = blk; res.err_code = ZE_corruption_detected; res.err_index = decoded; return; }
      }
    }
    pos += bh.csize; blk++;
    if (bh.last) return;
  }
}

// ---------------------------------------------------------------------------------------------------
// Huffman literals
// ---------------------------------------------------------------------------------------------------
// Stream layout of a literals section body (after the optional weight header), HufDecompress.cs:266-307:
// 4 streams: 6-byte jump table then the streams; stream k (k<3) regenerates seg=(n+3)/4 bytes at out+k*seg,
// the 4th regenerates max(0, n-3*seg) at out+3*seg.  1 stream: the whole body regenerates n bytes.
struct HufStream { const u8* src; u32 len; u32 outOfs; u32 count; };

// Returns false when the jump table is inconsistent (corruption_detected in the reference).
ZB_HD bool huf_split4(const u8* body, u32 bodySize, u32 n, u32 lane, HufStream& s) {
  if (bodySize < 10) return false;                                                 // :269
  u32 l1 = ld16(body), l2 = ld16(body + 2), l3 = ld16(body + 4);
  u32 l4 = bodySize - (l1 + l2 + l3 + 6);
  if (l4 > bodySize) return false;                                                 // :303
  u32 seg = (n + 3) / 4;
  u32 start = 6, len = l1;
  if (lane == 1) { start = 6 + l1; len = l2; } else if (lane == 2) { start = 6 + l1 + l2; len = l3; } else if (lane == 3) { start = 6 + l1 + l2 + l3; len = l4; }
  s.src = body + start; s.len = len; s.outOfs = lane * seg;
  s.count = lane < 3 ? seg : (n > 3 * seg ? n - 3 * seg : 0);
  return true;
}

// Decodes `count` symbols of one backward stream into out[0..count).  true iff the stream was consumed
// exactly (EndOfDStream, HufDecompress.cs:350-353 / :261) — which also implies it was never over-read.
ZB_HD bool huf_decode_stream(const u8* src, u32 len, u8* out, u32 count, const u16* dt, u32 tableLog) {
  BitCursor c;
  if (!bc_init(c, src, len)) return false;                                         // InitDStream errors :304-307
  i32 P = c.P;
  u32 left = count;
  const u32 sh = 64 - tableLog;
  // head: reach 4-byte alignment of the output
  while (left && ((uintptr_t)out & 3)) {
    u64 w = bc_window64(c, P);
    u32 cell = dt[(u32)(w >> sh)];
    *out++ = (u8)cell; P -= (i32)(cell >> 8); left--;
  }
  while (left >= 4) {
    u64 w = bc_window64(c, P);
    u32 c0 = dt[(u32)(w >> sh)]; w <<= (c0 >> 8);
    u32 c1 = dt[(u32)(w >> sh)]; w <<= (c1 >> 8);
    u32 c2 = dt[(u32)(w >> sh)]; w <<= (c2 >> 8);
    u32 c3 = dt[(u32)(w >> sh)];
    P -= (i32)((c0 >> 8) + (c1 >> 8) + (c2 >> 8) + (c3 >> 8));
    *(u32*)out = (c0 & 0xFF) | ((c1 & 0xFF) << 8) | ((c2 & 0xFF) << 16) | ((c3 & 0xFF) << 24);
    out += 4; left -= 4;
  }
  while (left) {
    u64 w = bc_window64(c, P);
    u32 cell = dt[(u32)(w >> sh)];
    *out++ = (u8)cell; P -= (i32)(cell >> 8); left--;
  }
  return P == 0;
}

}  // namespace zb
