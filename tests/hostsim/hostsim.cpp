// tests/hostsim/hostsim.cpp — CPU replay of the GPU decoder's per-thread stages (TEST AID ONLY).
//
// The build container has no GPU.  The entropy stages of the CUDA decoder are written as __host__ __device__
// functions (zstandard_b200/csrc/zb_format.cuh, zb_decode.cuh); this file compiles exactly those functions
// with g++ and drives them the way the kernels do (k_parse -> k_huf lanes -> k_seq -> a serial stand-in for the
// warp-cooperative k_exec) so that tests/test_hostsim.py can compare them with the oracle before any GPU time
// is spent.  It is never linked into libzstdb200.so.
#include <cstring>
#include <vector>
#include "../../zstandard_b200/csrc/zb_decode.cuh"
#include "../../zstandard_b200/csrc/zb_encode.cuh"

using namespace zb;

namespace {

struct Sim {
  std::vector<u16> ll, ml, of;       // lane-private tables, stride 1 here
  u16 defLL[64], defOF[32], defML[64];
  u32 llInfo[36], mlInfo[53];
  Sim() : ll(512), ml(512), of(256) {
    u16 sn[53]; s16 norm[53];
    for (int i = 0; i < 36; i++) norm[i] = kLLnorm[i]; build_seq_table(defLL, 1, norm, 35, 6, sn);
    for (int i = 0; i < 29; i++) norm[i] = kOFnorm[i]; build_seq_table(defOF, 1, norm, 28, 5, sn);
    for (int i = 0; i < 53; i++) norm[i] = kMLnorm[i]; build_seq_table(defML, 1, norm, 52, 6, sn);
    for (u32 i = 0; i < 36; i++) llInfo[i] = ll_info(i);
    for (u32 i = 0; i < 53; i++) mlInfo[i] = ml_info(i);
  }
};

// mirrors k_huf for one frame; returns first failing block / code through fi
void sim_huf(const u8* src, u32 size, FrameInfo& fi, u8* lit, u64 litCap) {
  static thread_local u16 dt[1 << HUF_LOG_MAX]; static thread_local HufBuildWk wk; alignas(16) u32 ringBuf[ZB_RING_WORDS];
  u32 pos = fi.body_off, blk = 0; u64 litRun = 0; u32 tableLog = 0; bool haveTable = false;
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
        if (lh.type == 3 && !haveTable) break;
        bool ok = true; u32 code = ZE_corruption_detected;
        const u8* body = bp + lh.lhSize; u32 bodySize = lh.litCSize;
        if (litRun + lh.litSize + 3 > litCap) { ok = false; code = ZE_dstSize_tooSmall; }
        if (ok && lh.type == 2) {
          if (!lh.single && (lh.litSize == 0 || bodySize == 0)) ok = false;
          u32 hdr = 0, nbSym = 0, tl = 0;
          if (ok) {
            u32 e = huf_read_weights(body, bodySize, wk, &hdr, &tl, &nbSym);
            if (!e && hdr >= bodySize) e = ZE_srcSize_wrong;
            if (e) ok = false;
          }
          if (ok) {
            for (u32 sub = 0; sub < 4; sub++) huf_fill_table(dt, wk, tl, nbSym, sub, 4);
            tableLog = tl; haveTable = true; body += hdr; bodySize -= hdr;
          }
        }
        if (ok) {
          if (lh.single) ok = huf_decode_stream(body, bodySize, lit + litRun, lh.litSize, dt, tableLog, ringBuf);
          else for (u32 sub = 0; sub < 4; sub++) {
            HufStream st; bool good = huf_split4(body, bodySize, lh.litSize, sub, st);
            if (good) good = huf_decode_stream(st.src, st.len, lit + litRun + st.outOfs, st.count, dt, tableLog, ringBuf);
            if (!good) ok = false;
          }
        }
        if (!ok) { fi.huf_err_block = blk; fi.huf_err_code = code; break; }
        litRun += lh.litSize;
      }
    }
    pos += bh.csize; blk++;
    if (bh.last) break;
  }
}

// serial stand-in for k_exec: same checks in the same order, byte-serial copies; no checksum verification
u32 sim_exec(const u8* src, u32 size, const FrameInfo& fi, u8* dst, u64 cap, const u8* litScratch, const SeqRec* recs, bool* needXxh, u32* trailerOff) {
  u32 pos = fi.body_off, blk = 0; u64 op = 0, litRun = 0, recRun = 0; bool litEntropy = false; u32 err = 0;
  while (true) {
    BlockHdr bh;
    err = read_block_hdr(src + pos, size - pos, bh);
    if (err) break;
    pos += 3;
    if (bh.type == 0) { if (bh.csize > cap - op) { err = ZE_dstSize_tooSmall; break; } memcpy(dst + op, src + pos, bh.csize); op += bh.csize; }
    else if (bh.type == 1) { if (bh.orig > cap - op) { err = ZE_dstSize_tooSmall; break; } memset(dst + op, src[pos], bh.orig); op += bh.orig; }
    else {
      const u8* bp = src + pos; const u32 bsz = bh.csize;
      if (bsz >= BLOCKSIZE_MAX) { err = ZE_srcSize_wrong; break; }
      LitHdr lh; bool needs;
      u32 e = read_lit_hdr(bp, bsz, lh, &needs);
      if (needs && !litEntropy) { err = ZE_dictionary_corrupted; break; }
      if (e) { err = e; break; }
      const u8* lit = nullptr; u32 rleByte = 0; bool isRle = false;
      if (lh.type >= 2) { if (fi.huf_err_block == blk) { err = fi.huf_err_code; break; } litEntropy = true; lit = litScratch + litRun; litRun += lh.litSize; }
      else if (lh.type == 0) lit = bp + lh.lhSize;
      else { isRle = true; rleByte = bp[lh.lhSize]; }
      const u32 litSize = lh.litSize;
      const u8* sp = bp + lh.consumed; const u32 ssz = bsz - lh.consumed;
      u32 nbSeq, modes, hdr;
      e = read_seq_count(sp, ssz, &nbSeq, &modes, &hdr);
      if (e) { err = e; break; }
      if (fi.seq_err_block == blk && fi.seq_err_index == 0xFFFFFFFFu) { err = fi.seq_err_code; break; }
      u64 litPos = 0;
      if (nbSeq) {
        const SeqRec* r = recs + recRun;
        for (; r->x != 0; r++) {
          u32 ll = r->y & 0xFFFF, ml = r->y >> 16, off = r->x;
          if (op + ll + ml > cap) { err = ZE_dstSize_tooSmall; break; }
          if (litPos + ll > litSize) { err = ZE_corruption_detected; break; }
          if ((u64)off > op + ll) { err = ZE_corruption_detected; break; }
          for (u32 i = 0; i < ll; i++) dst[op + i] = isRle ? (u8)rleByte : lit[litPos + i];
          op += ll; litPos += ll;
          for (u32 i = 0; i < ml; i++) dst[op + i] = dst[op + i - off];
          op += ml;
        }
        if (err) break;
        recRun = (u64)(r - recs) + 1;
        if (fi.seq_err_block == blk) { err = fi.seq_err_code; break; }
      }
      u64 lastLL = litSize - litPos;
      if (lastLL > cap - op) { err = ZE_dstSize_tooSmall; break; }
      for (u64 i = 0; i < lastLL; i++) dst[op + i] = isRle ? (u8)rleByte : lit[litPos + i];
      op += lastLL;
    }
    pos += bh.csize; blk++;
    if (bh.last) break;
  }
  *needXxh = false; *trailerOff = 0;
  if (!err) {
    if ((fi.flags & FI_FCS_KNOWN) && op != fi.fcs) err = ZE_corruption_detected;
    else if (fi.flags & FI_CHECKSUM) { if (size - pos < 4) err = ZE_checksum_wrong; else { *trailerOff = pos; pos += 4; *needXxh = true; } }
  }
  u32 tailErr = 0;
  if (!err) while (true) {
    u32 rem = size - pos;
    if (rem < 5) { if (rem) tailErr = ZE_srcSize_wrong; break; }
    u32 magic = ld32(src + pos);
    if (magic == MAGIC) { tailErr = ZE_GENERIC; break; }
    if ((magic & 0xFFFFFFF0u) != MAGIC_SKIP) { tailErr = ZE_prefix_unknown; break; }
    if (rem < 8) { tailErr = ZE_srcSize_wrong; break; }
    u32 skip = ld32(src + pos + 4) + 8u;
    if (rem < skip) { tailErr = ZE_srcSize_wrong; break; }
    pos += skip;
  }
  return err ? zerr(err) : (tailErr ? zerr(tailErr) : (u32)op);
}

}  // namespace

extern "C" uint32_t hostsim_decompress(uint8_t* dst, uint32_t cap, const uint8_t* src_in, uint32_t size, uint32_t* trailer_off, int* need_xxh) {
  // the device reads whole aligned words around the streams: give the copy slack on both sides
  std::vector<u8> padded(size + 32, 0);
  u8* src = padded.data() + 16;
  memcpy(src, src_in, size);
  FrameInfo fi; u32 r = 0;
  *need_xxh = 0; *trailer_off = 0;
  if (!parse_item(src, size, fi, &r)) return r;
  std::vector<u8> lit((size_t)cap + 64);
  std::vector<SeqRec> recs(seq_capacity(cap) + 40);
  sim_huf(src, size, fi, lit.data(), (u64)cap + 40);
  static thread_local Sim* sim = nullptr; if (!sim) sim = new Sim();
  SeqTableSet T;
  T.space[KIND_LL] = sim->ll.data(); T.space[KIND_ML] = sim->ml.data(); T.space[KIND_OF] = sim->of.data(); T.stride = 1;
  T.defs[KIND_LL] = sim->defLL; T.defs[KIND_OF] = sim->defOF; T.defs[KIND_ML] = sim->defML;
  SeqFrameOut res;
  s16 normBuf[53]; u16 nextBuf[53]; alignas(16) u32 ringBuf[ZB_RING_WORDS];
  seq_decode_frame(src, size, fi.body_off, T, recs.data(), seq_capacity(cap), res, sim->llInfo, sim->mlInfo, normBuf, nextBuf, ringBuf);
  if (res.err_block != 0xFFFFFFFFu) { fi.seq_err_block = res.err_block; fi.seq_err_code = res.err_code; fi.seq_err_index = res.err_index; }
  bool nx; u32 tr;
  static u8 dummy[8];
  u32 out = sim_exec(src, size, fi, dst ? dst : dummy, cap, lit.data(), recs.data(), &nx, &tr);
  *need_xxh = nx; *trailer_off = tr;
  return out;
}

// debug surface: literals + records of the first frame of an item (no execution)
extern "C" uint32_t hostsim_stages(const uint8_t* src_in, uint32_t size, uint32_t cap, uint8_t* lit_out, uint32_t* rec_out, uint32_t max_recs, uint32_t* info_out) {
  std::vector<u8> padded(size + 32, 0);
  u8* src = padded.data() + 16;
  memcpy(src, src_in, size);
  FrameInfo fi; u32 r = 0;
  if (!parse_item(src, size, fi, &r)) return r;
  std::vector<u8> lit((size_t)cap + 64);
  std::vector<SeqRec> recs(seq_capacity(cap) + 40);
  sim_huf(src, size, fi, lit.data(), (u64)cap + 40);
  static thread_local Sim* sim = nullptr; if (!sim) sim = new Sim();
  SeqTableSet T;
  T.space[KIND_LL] = sim->ll.data(); T.space[KIND_ML] = sim->ml.data(); T.space[KIND_OF] = sim->of.data(); T.stride = 1;
  T.defs[KIND_LL] = sim->defLL; T.defs[KIND_OF] = sim->defOF; T.defs[KIND_ML] = sim->defML;
  SeqFrameOut res;
  s16 normBuf[53]; u16 nextBuf[53]; alignas(16) u32 ringBuf[ZB_RING_WORDS];
  seq_decode_frame(src, size, fi.body_off, T, recs.data(), seq_capacity(cap), res, sim->llInfo, sim->mlInfo, normBuf, nextBuf, ringBuf);
  memcpy(lit_out, lit.data(), cap);
  u32 n = (u32)std::min<size_t>(max_recs, recs.size());
  memcpy(rec_out, recs.data(), (size_t)n * 8);
  info_out[0] = fi.huf_err_block; info_out[1] = fi.huf_err_code; info_out[2] = res.err_block; info_out[3] = res.err_code; info_out[4] = res.err_index;
  return 0;
}

// ---- encoder replay: one frame through zb_encode.cuh's encode_frame (checksum slot left for the caller) ----
extern "C" uint32_t hostsim_compress(uint8_t* dst, uint32_t cap, const uint8_t* src_in, uint32_t size, int level, int checksum) {
  std::vector<u8> padded(size + 64, 0);
  u8* src = padded.data() + 16;
  if (size) memcpy(src, src_in, size);
  const u32 seqCap = BLOCKSIZE_MAX / 4 + 64;
  std::vector<u32> table(enc_table_words(level)); std::vector<u8> lits(BLOCKSIZE_MAX + 64); std::vector<u32> seqs(2 * seqCap);
  std::vector<u8> codes(3 * seqCap); std::vector<u16> ct(3 * 514 + 16); std::vector<u8> tmp(BLOCKSIZE_MAX + 4096);
  EncScratch sc; sc.table = table.data(); sc.lits = lits.data(); sc.seqs = seqs.data(); sc.seqCap = seqCap; sc.codes = codes.data();
  sc.ctables = ct.data(); sc.tmp = tmp.data();
  return encode_frame(src, size, dst, cap, level, checksum, sc);
}
