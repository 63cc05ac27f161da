// tests/hostsim/hostsim.cpp — CPU replay of the GPU decoder's per-thread stages (TEST AID ONLY).
//
// The build container has no GPU.  The entropy stages of the CUDA decoder are written as __host__ __device__
// functions (zstandard_b200/csrc/zb_format.cuh, zb_decode.cuh); this file compiles exactly those functions
// with g++ and drives them the way the kernels do (k_parse -> k_huf lanes -> k_seq -> a serial stand-in for the
// warp-cooperative k_exec) so that tests/test_hostsim.py can compare them with the oracle before any GPU time
// is spent.  It is never linked into libzstdb200.so.
#include <stdio.h>
#include <stdlib.h>
#include <cstring>
#include <cstdlib>
#include <vector>
#include "../../zstandard_b200/csrc/zb_decode.cuh"
#include "../../zstandard_b200/csrc/zb_blocks.cuh"
#include "../../zstandard_b200/csrc/zb_encode.cuh"
#include "serial_encoder.h"

using namespace zb;

namespace {

struct Sim {
  std::vector<u16> ll, ml, of;       // lane-private tables, stride 1 here
  u16 defLL[64], defOF[32], defML[64];
  u32 llInfo[36], mlInfo[53];
  Sim() : ll(512), ml(512), of(256) {
    u16 sn[53]; s16 norm[53];
    for (int i = 0; i < 36; i++) norm[i] = kLLnorm[i];
    build_seq_table(defLL, 1, norm, 35, 6, sn);
    for (int i = 0; i < 29; i++) norm[i] = kOFnorm[i];
    build_seq_table(defOF, 1, norm, 28, 5, sn);
    for (int i = 0; i < 53; i++) norm[i] = kMLnorm[i];
    build_seq_table(defML, 1, norm, 52, 6, sn);
    for (u32 i = 0; i < 36; i++) llInfo[i] = ll_info(i);
    for (u32 i = 0; i < 53; i++) mlInfo[i] = ml_info(i);
  }
};

// mirrors k_huf for one frame; returns first failing block / code through fi
// the replay's dictionary (hostsim_set_dict): what zstdb200_load_dictionary prepares on the device
static DictState g_dict; static std::vector<u8> g_dictBytes; static bool g_haveDict = false;
static u32 g_seqCap = 9; static int g_deferred = 0;   // hostsim_set_seq_cap: table log the sequence stage's first attempt has room for;
static int g_classFrames[3] = {0, 0, 0};              // 0 = k_parse's own choice (first_block_classes): frames per class (full size, A, B)
static u32 g_rootLog = HUF_ROOT_SMALL;   // hostsim_set_huf_root: the Huffman kernels run with root tables of 2^9 (small frames) or 2^11 cells
static const DictState* cur_dict() { return g_haveDict ? &g_dict : nullptr; }

void sim_huf(const u8* src, u32 size, FrameInfo& fi, u8* lit, u64 litCap) {
  // as the kernels: a root table (shared memory there) plus the full table for the codes longer than the root log
  alignas(16) static thread_local u16 root[1 << HUF_TABLE_LOG], fullMem[1 << HUF_TABLE_LOG]; static thread_local HufBuildWk wk; alignas(16) u32 ringBuf[ZB_RING_WORDS];
  static thread_local u8 sideMem[256], slotMem[256]; static thread_local HufFseScratch fs;
  HufTabs T{root, g_rootLog, fullMem, sideMem, 0};
  u32 pos = fi.body_off, blk = 0; u64 litRun = 0; bool haveTable = false;
  if (const DictState* ds = cur_dict()) if (ds->hasEntropy) {
    for (u32 sub = 0; sub < 4; sub++) huf_root_from_full(root, g_rootLog, ds->huf, ds->hufLog, sub, 4);
    T.full = ds->huf; T.side = ds->hufSide; T.log = ds->hufLog; haveTable = true;
  }
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
        bool ok = true;
        const u8* body = bp + lh.lhSize; u32 bodySize = lh.litCSize;
        const bool dry = litRun + lh.litSize + 3 > litCap;
        if (lh.type == 2) {
          if (!lh.single && (lh.litSize == 0 || bodySize == 0)) ok = false;
          u32 hdr = 0, nbSym = 0, tl = 0;
          if (ok) {
            u32 e = huf_read_weights(body, bodySize, wk, fs, slotMem, &hdr, &tl, &nbSym);
            if (!e && hdr >= bodySize) e = ZE_srcSize_wrong;
            if (e) ok = false;
          }
          if (ok) {
            for (u32 sub = 0; sub < 4; sub++) huf_fill_root(root, g_rootLog, wk, slotMem, tl, nbSym, sub, 4);
            if (tl > g_rootLog) for (u32 sub = 0; sub < 4; sub++) huf_fill_table(fullMem, sideMem, wk, slotMem, tl, nbSym, sub, 4, g_rootLog);
            T.full = fullMem; T.side = sideMem; T.log = tl;
            haveTable = true; body += hdr; bodySize -= hdr;
          }
        }
        if (ok) {
          if (lh.single) ok = dry ? huf_check_stream(body, bodySize, lh.litSize, T) : huf_decode_stream(body, bodySize, lit + litRun, lh.litSize, T, ringBuf);
          else for (u32 sub = 0; sub < 4; sub++) {
            HufStream st; bool good = huf_split4(body, bodySize, lh.litSize, sub, st);
            if (good) good = dry ? huf_check_stream(st.src, st.len, st.count, T) : huf_decode_stream(st.src, st.len, lit + litRun + st.outOfs, st.count, T, ringBuf);
            if (!good) ok = false;
          }
        }
        if (!ok || dry) { fi.huf_err_block = blk; fi.huf_err_code = ok ? HUF_DRY : ZE_corruption_detected; break; }
        litRun += lh.litSize;
      }
    }
    pos += bh.csize; blk++;
    if (bh.last) break;
  }
}

// serial stand-in for k_exec: same checks in the same order, byte-serial copies; no checksum verification
u32 sim_exec(const u8* src, u32 size, const FrameInfo& fi, u8* dst, u64 cap, const u8* litScratch, const SeqRec* recs, bool* needXxh, u32* trailerOff, u32* nextOff, u32* decoded) {
  u32 pos = fi.body_off, blk = 0; u64 op = 0, litRun = 0, recRun = 0; bool litEntropy = false, dry = false; u32 err = 0;
  const DictState* ds = cur_dict();
  const u32 dictContent = ds ? ds->contentSize : 0;
  const u8* const dictEnd = ds ? g_dictBytes.data() + ds->contentOff + dictContent : nullptr;
  if (ds && ds->hasEntropy) litEntropy = true;
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
      if (lh.type >= 2) { if (fi.huf_err_block == blk) { if (fi.huf_err_code != HUF_DRY) { err = fi.huf_err_code; break; } dry = true; } litEntropy = true; lit = litScratch + litRun; litRun += lh.litSize; }
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
        // header record (count, output bytes, literal bytes), then self-describing records (zb_decode.cuh)
        u32 nRecs = 0;
        if (recRun < seq_capacity(cap)) nRecs = recs[recRun].x;
        const SeqRec* r = recs + recRun + 1;
        const u64 blockBase = op;
        for (u32 k = 0; k < nRecs; k++, r++) {
          const u32 ll = rec_ll(*r), ml = rec_ml(*r), off = r->z;
          const u64 start = blockBase + r->x;
          if (start != op && start <= cap) { fprintf(stderr, "hostsim: record position %llu != %llu\n", (unsigned long long)start, (unsigned long long)op); abort(); }
          if (start + ll + ml > cap) { err = ZE_dstSize_tooSmall; break; }
          if ((u64)rec_lpos(*r) + ll > litSize) { err = ZE_corruption_detected; break; }
          if ((u64)off > start + ll + dictContent) { err = ZE_corruption_detected; break; }
          if (rec_lpos(*r) != litPos) { fprintf(stderr, "hostsim: literal position mismatch\n"); abort(); }
          if (!dry) for (u32 i = 0; i < ll; i++) dst[op + i] = isRle ? (u8)rleByte : lit[litPos + i];
          op += ll; litPos += ll;
          if (!dry) for (u32 i = 0; i < ml; i++) dst[op + i] = (u64)off > op + i ? dictEnd[(i64)(op + i) - (i64)off] : dst[op + i - off];
          op += ml;
        }
        if (err) break;
        recRun += 1 + (u64)nRecs;
        if (fi.seq_err_block == blk) { err = fi.seq_err_code; break; }
      }
      u64 lastLL = litSize - litPos;
      if (lastLL > cap - op || dry) { err = ZE_dstSize_tooSmall; break; }
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
  u32 tailErr = 0; *decoded = (u32)op;
  if (!err) while (true) {
    u32 rem = size - pos;
    if (rem < 5) { if (rem) tailErr = ZE_srcSize_wrong; break; }
    u32 magic = ld32(src + pos);
    if (magic == MAGIC) { *nextOff = pos; break; }
    if ((magic & 0xFFFFFFF0u) != MAGIC_SKIP) { tailErr = ZE_prefix_unknown; break; }
    if (rem < 8) { tailErr = ZE_srcSize_wrong; break; }
    u32 skip = ld32(src + pos + 4) + 8u;
    if (rem < skip) { tailErr = ZE_srcSize_wrong; break; }
    pos += skip;
  }
  return err ? zerr(err) : (tailErr ? zerr(tailErr) : fi.out_base + (u32)op);
}

// ---- block-parallel path (zb_blocks.cuh): k_huf_blk / k_seq_blk / k_exec_big, one unit at a time ----
static int g_par = 1;            // hostsim_set_par: 0 = every frame on the frame-serial replay
static int g_parFrames = 0;      // data frames that took the block-parallel replay (hostsim_par_frames)

// mirrors k_huf_blk for one unit
void sim_huf_unit(const u8* src, BlockUnit& u, const BlockUnit* frameUnits, u8* litRegion) {
  alignas(16) static thread_local u16 root[1 << HUF_TABLE_LOG], fullMem[1 << HUF_TABLE_LOG]; static thread_local HufBuildWk wk; alignas(16) u32 ringBuf[ZB_RING_WORDS];
  static thread_local u8 sideMem[256], slotMem[256]; static thread_local HufFseScratch fs;
  HufTabs T{root, g_rootLog, fullMem, sideMem, 0};
  const u8* bp = src + u.body;
  LitHdr lh; bool needs;
  read_lit_hdr(bp, u.csize, lh, &needs);
  if (lh.type < 2) return;
  u8* lit = litRegion + u.lit_off;
  bool ok = true;
  const u8* body = bp + lh.lhSize; u32 bodySize = lh.litCSize;
  if (lh.type == 3 && u.huf_def == DEF_DICT) {
    const DictState* ds = cur_dict();
    for (u32 sub = 0; sub < 4; sub++) huf_root_from_full(root, g_rootLog, ds->huf, ds->hufLog, sub, 4);
    T.full = ds->huf; T.side = ds->hufSide; T.log = ds->hufLog;
  } else {
    const u8* tb = body; u32 tbSize = bodySize;
    if (lh.type == 3) {
      const BlockUnit& d = frameUnits[u.huf_def];
      LitHdr lhD; bool n2;
      read_lit_hdr(src + d.body, d.csize, lhD, &n2);
      tb = src + d.body + lhD.lhSize; tbSize = lhD.litCSize;
    } else if (!lh.single && (lh.litSize == 0 || bodySize == 0)) ok = false;
    u32 hdr = 0, nbSym = 0, tl = 0;
    if (ok) {
      u32 e = huf_read_weights(tb, tbSize, wk, fs, slotMem, &hdr, &tl, &nbSym);
      if (!e && hdr >= tbSize) e = ZE_srcSize_wrong;
      if (e) ok = false;
    }
    if (ok) {
      for (u32 sub = 0; sub < 4; sub++) huf_fill_root(root, g_rootLog, wk, slotMem, tl, nbSym, sub, 4);
      if (tl > g_rootLog) for (u32 sub = 0; sub < 4; sub++) huf_fill_table(fullMem, sideMem, wk, slotMem, tl, nbSym, sub, 4, g_rootLog);
      T.log = tl;
      if (lh.type == 2) { body += hdr; bodySize -= hdr; }
    }
  }
  if (ok) {
    if (lh.single) ok = huf_decode_stream(body, bodySize, lit, lh.litSize, T, ringBuf);
    else for (u32 sub = 0; sub < 4; sub++) {
      HufStream st; bool good = huf_split4(body, bodySize, lh.litSize, sub, st);
      if (good) good = huf_decode_stream(st.src, st.len, lit + st.outOfs, st.count, T, ringBuf);
      if (!good) ok = false;
    }
  }
  if (!ok) u.huf_err = ZE_corruption_detected;
}

// serial stand-in for k_exec_big: the same checks in block order, symbolic offsets resolved against the running history
u32 sim_exec_par(const u8* src, u32 size, const FrameInfo& fi, const BlockUnit* units, u8* dst, u64 cap, const u8* litScratch, const SeqRec* recs, bool* needXxh, u32* trailerOff, u32* nextOff, u32* decoded) {
  u32 pos = fi.body_off, cblk = 0; u64 op = 0; u32 err = 0;
  const DictState* ds = cur_dict();
  const u32 dictContent = ds ? ds->contentSize : 0;
  const u8* const dictEnd = ds ? g_dictBytes.data() + ds->contentOff + dictContent : nullptr;
  u32 R[3] = {1, 4, 8};
  if (ds) { R[0] = ds->rep[0]; R[1] = ds->rep[1]; R[2] = ds->rep[2]; }
  while (true) {
    BlockHdr bh;
    err = read_block_hdr(src + pos, size - pos, bh);
    if (err) break;
    pos += 3;
    if (bh.type == 0) { if (bh.csize > cap - op) { err = ZE_dstSize_tooSmall; break; } memcpy(dst + op, src + pos, bh.csize); op += bh.csize; }
    else if (bh.type == 1) { if (bh.orig > cap - op) { err = ZE_dstSize_tooSmall; break; } memset(dst + op, src[pos], bh.orig); op += bh.orig; }
    else {
      const BlockUnit& bu = units[cblk++];
      if (bu.body != pos || bu.csize != bh.csize) { fprintf(stderr, "hostsim: unit / block mismatch\n"); abort(); }
      const u8* bp = src + pos; const u32 bsz = bh.csize;
      LitHdr lh; bool needs;
      read_lit_hdr(bp, bsz, lh, &needs);
      const u8* lit = nullptr; u32 rleByte = 0; bool isRle = false;
      if (lh.type >= 2) { if (bu.huf_err) { err = bu.huf_err; break; } lit = litScratch + bu.lit_off; }
      else if (lh.type == 0) lit = bp + lh.lhSize;
      else { isRle = true; rleByte = bp[lh.lhSize]; }
      const u32 litSize = lh.litSize;
      const u8* sp = bp + lh.consumed; const u32 ssz = bsz - lh.consumed;
      u32 nbSeq, modes, hdr;
      read_seq_count(sp, ssz, &nbSeq, &modes, &hdr);
      if (bu.seq_err_code && bu.seq_err_index == 0xFFFFFFFFu) { err = bu.seq_err_code; break; }
      u64 litPos = 0;
      if (nbSeq) {
        const u32 nRecs = recs[bu.rec_off].x;
        const SeqRec* r = recs + bu.rec_off + 1;
        const u64 blockBase = op;
        for (u32 k = 0; k < nRecs; k++, r++) {
          const u32 ll = rec_ll(*r), ml = rec_ml(*r), off = repsym_resolve(r->z, rec_tag(*r), R);
          const u64 start = blockBase + r->x;
          if (start != op && start <= cap) { fprintf(stderr, "hostsim: record position %llu != %llu\n", (unsigned long long)start, (unsigned long long)op); abort(); }
          if (start + ll + ml > cap) { err = ZE_dstSize_tooSmall; break; }
          if ((u64)rec_lpos(*r) + ll > litSize) { err = ZE_corruption_detected; break; }
          if ((u64)off > start + ll + dictContent) { err = ZE_corruption_detected; break; }
          for (u32 i = 0; i < ll; i++) dst[op + i] = isRle ? (u8)rleByte : lit[litPos + i];
          op += ll; litPos += ll;
          for (u32 i = 0; i < ml; i++) dst[op + i] = (u64)off > op + i ? dictEnd[(i64)(op + i) - (i64)off] : dst[op + i - off];
          op += ml;
        }
        if (err) break;
        if (bu.seq_err_code) { err = bu.seq_err_code; break; }
        op = blockBase + recs[bu.rec_off].y; litPos = recs[bu.rec_off].z;
        const u32 n0 = repsym_resolve(bu.rep[0], bu.rep_sym & 7, R), n1 = repsym_resolve(bu.rep[1], (bu.rep_sym >> 3) & 7, R), n2 = repsym_resolve(bu.rep[2], (bu.rep_sym >> 6) & 7, R);
        R[0] = n0; R[1] = n1; R[2] = n2;
      }
      u64 lastLL = litSize - litPos;
      if (lastLL > cap - op) { err = ZE_dstSize_tooSmall; break; }
      for (u64 i = 0; i < lastLL; i++) dst[op + i] = isRle ? (u8)rleByte : lit[litPos + i];
      op += lastLL;
    }
    pos += bh.csize;
    if (bh.last) break;
  }
  *needXxh = false; *trailerOff = 0;
  if (!err) {
    if ((fi.flags & FI_FCS_KNOWN) && op != fi.fcs) err = ZE_corruption_detected;
    else if (fi.flags & FI_CHECKSUM) { if (size - pos < 4) err = ZE_checksum_wrong; else { *trailerOff = pos; pos += 4; *needXxh = true; } }
  }
  u32 tailErr = 0; *decoded = (u32)op;
  if (!err) while (true) {
    u32 rem = size - pos;
    if (rem < 5) { if (rem) tailErr = ZE_srcSize_wrong; break; }
    u32 magic = ld32(src + pos);
    if (magic == MAGIC) { *nextOff = pos; break; }
    if ((magic & 0xFFFFFFF0u) != MAGIC_SKIP) { tailErr = ZE_prefix_unknown; break; }
    if (rem < 8) { tailErr = ZE_srcSize_wrong; break; }
    u32 skip = ld32(src + pos + 4) + 8u;
    if (rem < skip) { tailErr = ZE_srcSize_wrong; break; }
    pos += skip;
  }
  return err ? zerr(err) : (tailErr ? zerr(tailErr) : fi.out_base + (u32)op);
}

}  // namespace

// One pass per data frame of the item, as the host loop around the kernels does (api.cu, decode_more_passes).
// xxh: XXH64 implementation to stand in for k_xxh (the test passes the oracle's); null = checksums are not verified
// and trailer_off / need_xxh describe the LAST data frame, *last_base where its output starts within dst.
typedef uint64_t (*xxh_fn)(const void*, size_t, uint64_t);
extern "C" uint32_t hostsim_decompress2(uint8_t* dst_in, uint32_t capAll, const uint8_t* src_in, uint32_t size, uint32_t* trailer_off, int* need_xxh, uint32_t* last_base, xxh_fn xxh) {
  // the device reads whole aligned words around the streams: give the copy slack on both sides
  std::vector<u8> padded(size + 32, 0);
  u8* src = padded.data() + 16;
  memcpy(src, src_in, size);
  *need_xxh = 0; *trailer_off = 0; *last_base = 0;
  static u8 dummy[8];
  u32 start = 0, outBase = 0, out = 0;
  while (true) {
    FrameInfo fi; u32 r = 0;
    if (!parse_item(src, size, fi, &r, start, outBase, cur_dict() ? cur_dict()->err : 0, cur_dict() ? cur_dict()->dictID : 0)) return r;
    const u32 cap = capAll - fi.out_base; u8* dst = dst_in ? dst_in + fi.out_base : dummy;
    std::vector<u8> lit((size_t)cap + 64);
    std::vector<SeqRec> recs(seq_capacity(cap) + 40);
    static thread_local Sim* sim = nullptr; if (!sim) sim = new Sim();
    SeqTableSet T;
    T.space[KIND_LL] = sim->ll.data(); T.space[KIND_ML] = sim->ml.data(); T.space[KIND_OF] = sim->of.data(); T.stride = 1;
    T.defs[KIND_LL] = sim->defLL; T.defs[KIND_OF] = sim->defOF; T.defs[KIND_ML] = sim->defML;
    s16 normBuf[53]; u16 nextBuf[53]; alignas(16) u32 ringBuf[ZB_RING_WORDS];
    bool nx; u32 tr, nextOff = 0, produced = 0;
    // k_parse's decision: structurally sound multi-block frames become units (zb_blocks.cuh)
    const u32 maxU = cap / PAR_UNIT_BYTES + PAR_UNIT_SLACK;
    const u32 nu = g_par ? par_walk(src, size, fi.body_off, (u64)cap + 40, seq_capacity(cap), cur_dict(), maxU, nullptr, 0) : 0;
    if (nu) {
      g_parFrames++;
      std::vector<BlockUnit> units(nu);
      par_walk(src, size, fi.body_off, (u64)cap + 40, seq_capacity(cap), cur_dict(), maxU, units.data(), 0);
      for (u32 w = 0; w < nu; w++) sim_huf_unit(src, units[w], units.data(), lit.data());
      for (u32 w = 0; w < nu; w++) {
        UnitEmitter em; em.init(recs.data() + units[w].rec_off, sim->llInfo, sim->mlInfo);
        seq_decode_unit(src, units[w], units.data(), fi.window, T, em, sim->llInfo, sim->mlInfo, normBuf, nextBuf, ringBuf, cur_dict());
        if (em.dead) { units[w].seq_err_code = em.err_code; units[w].seq_err_index = em.err_index; }
        else if (em.n) { units[w].rep[0] = em.h.v0; units[w].rep[1] = em.h.v1; units[w].rep[2] = em.h.v2; units[w].rep_sym = em.h.t0 | (em.h.t1 << 3) | (em.h.t2 << 6); }
      }
      out = sim_exec_par(src, size, fi, units.data(), dst, cap, lit.data(), recs.data(), &nx, &tr, &nextOff, &produced);
    } else {
    sim_huf(src, size, fi, lit.data(), (u64)cap + 40);
    SeqEmitter em; em.init(recs.data(), seq_capacity(cap), sim->llInfo, sim->mlInfo);      // the finishing half, plugged in directly
    if (cur_dict()) em.set_reps(cur_dict()->rep);
    // the kernel instantiation with small tables first (k_seq_t<1 / 2>), the full-size one if it hands the frame over
    const bool dictTables = cur_dict() && cur_dict()->hasEntropy;
    if (g_seqCap == 0) {                                                           // as k_parse + the three k_seq_t instantiations do
      bool few; u32 cls = first_block_classes(src + fi.body_off, size - fi.body_off, 512, 2048, &few);
      if (dictTables) cls = 0;
      g_classFrames[cls]++;
      if (cls == 1) { T.cap[KIND_LL] = 6; T.cap[KIND_OF] = 6; T.cap[KIND_ML] = 7; }
      else if (cls == 2) { T.cap[KIND_LL] = T.cap[KIND_OF] = T.cap[KIND_ML] = 8; }
    } else if (g_seqCap < 9 && !dictTables) { T.cap[0] = T.cap[1] = T.cap[2] = g_seqCap; }
    seq_decode_frame(src, size, fi.body_off, fi.window, T, em, sim->llInfo, sim->mlInfo, normBuf, nextBuf, ringBuf, cur_dict());
    if (em.deferred) {
      g_deferred++;
      T.cap[0] = T.cap[1] = T.cap[2] = 9;
      em.init(recs.data(), seq_capacity(cap), sim->llInfo, sim->mlInfo);
      if (cur_dict()) em.set_reps(cur_dict()->rep);
      seq_decode_frame(src, size, fi.body_off, fi.window, T, em, sim->llInfo, sim->mlInfo, normBuf, nextBuf, ringBuf, cur_dict());
    }
    const SeqFrameOut res = em.res;
    if (res.err_block != 0xFFFFFFFFu) { fi.seq_err_block = res.err_block; fi.seq_err_code = res.err_code; fi.seq_err_index = res.err_index; }
    out = sim_exec(src, size, fi, dst, cap, lit.data(), recs.data(), &nx, &tr, &nextOff, &produced);
    }
    *need_xxh = nx; *trailer_off = tr; *last_base = fi.out_base;
    // k_xxh runs whenever the frame itself decoded (even if what follows it in the item is malformed) and overrides the result
    if (nx && xxh && dst_in && (u32)xxh(dst, produced, 0) != ld32(src + tr)) return zerr(ZE_checksum_wrong);
    if (is_err(out) || !nextOff) return out;
    start = nextOff; outBase = out;
  }
}
extern "C" void hostsim_set_par(int on) { g_par = on; }
extern "C" void hostsim_set_huf_root(int rootLog) { g_rootLog = (u32)rootLog; }
extern "C" void hostsim_set_seq_cap(int capLog) { g_seqCap = (u32)capLog; }
extern "C" int hostsim_deferred() { return g_deferred; }
extern "C" int hostsim_class_frames(int cls) { return g_classFrames[cls]; }
extern "C" int hostsim_par_frames() { return g_parFrames; }
extern "C" uint32_t hostsim_decompress(uint8_t* dst, uint32_t cap, const uint8_t* src_in, uint32_t size, uint32_t* trailer_off, int* need_xxh) {
  uint32_t lastBase;
  return hostsim_decompress2(dst, cap, src_in, size, trailer_off, need_xxh, &lastBase, nullptr);
}

// debug surface: literals + records of the first frame of an item (no execution)
extern "C" uint32_t hostsim_stages(const uint8_t* src_in, uint32_t size, uint32_t cap, uint8_t* lit_out, uint32_t* rec_out, uint32_t max_recs, uint32_t* info_out) {
  std::vector<u8> padded(size + 32, 0);
  u8* src = padded.data() + 16;
  memcpy(src, src_in, size);
  FrameInfo fi; u32 r = 0;
  if (!parse_item(src, size, fi, &r, 0, 0, cur_dict() ? cur_dict()->err : 0, cur_dict() ? cur_dict()->dictID : 0)) return r;
  std::vector<u8> lit((size_t)cap + 64);
  std::vector<SeqRec> recs(seq_capacity(cap) + 40);
  sim_huf(src, size, fi, lit.data(), (u64)cap + 40);
  static thread_local Sim* sim = nullptr; if (!sim) sim = new Sim();
  SeqTableSet T;
  T.space[KIND_LL] = sim->ll.data(); T.space[KIND_ML] = sim->ml.data(); T.space[KIND_OF] = sim->of.data(); T.stride = 1;
  T.defs[KIND_LL] = sim->defLL; T.defs[KIND_OF] = sim->defOF; T.defs[KIND_ML] = sim->defML;
  s16 normBuf[53]; u16 nextBuf[53]; alignas(16) u32 ringBuf[ZB_RING_WORDS];
  SeqEmitter em; em.init(recs.data(), seq_capacity(cap), sim->llInfo, sim->mlInfo);
  if (cur_dict()) em.set_reps(cur_dict()->rep);
  seq_decode_frame(src, size, fi.body_off, fi.window, T, em, sim->llInfo, sim->mlInfo, normBuf, nextBuf, ringBuf, cur_dict());
  const SeqFrameOut res = em.res;
  memcpy(lit_out, lit.data(), cap);
  u32 n = (u32)std::min<size_t>(max_recs, recs.size());
  memcpy(rec_out, recs.data(), (size_t)n * sizeof(SeqRec));
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

// ---- lock-step CPU emulation of the warp-parallel match finder (k_enc_match in encode_kernels.cu) ----
// Same algorithm, lanes emulated by loops; used to study ratio without a GPU and as a second replay in tests.
namespace {
struct WarpMatcher {
  std::vector<u16> tab; u32 hlogL, hlogS, mls; bool dfast;
  const u8* src; u32 size; u32 rep1 = 1, rep2 = 4;
  std::vector<u32> seqs; std::vector<u8> lits;
  static u32 h64(u64 v, u32 hlog, u32 mls) {
    if (mls >= 8) return (u32)((v * 0xCF1BBCDCB7A56463ull) >> (64 - hlog));
    const u32 lo = (u32)v, hi = (u32)(v >> 32);                 // 5 / 6 bytes: two 32-bit multiplies (as k_enc_match)
    return (lo * 2654435761u + (hi & (mls == 6 ? 0xFFFFu : 0xFFu)) * 2246822519u) >> (32 - hlog);
  }
  static bool candFrom(u32 p, u32 e, u32& c) { u32 d = (p - e) & 0xFFFF; if (!d) d = 0x10000; c = p - d; return d <= p; }
  u32 extend(u32 a, u32 off, u32 end) { u32 n = 0; while (a + n < end && src[a + n] == src[a + n - off]) n++; return n; }
  bool operator()(SeqStore& st, u32 blk, const u8* bstart, u32 bsize, bool isRle) {
    if (isRle || bsize < 64) return false;
    seqs.assign(2 * (BLOCKSIZE_MAX / 4 + 64), 0); lits.assign(BLOCKSIZE_MAX + 64, 0);
    u32 nseq = 0, nlits = 0;
    const u32 bpos = (u32)(bstart - src), bend = bpos + bsize;
    if (blk > 0) { rep1 = 0; rep2 = 0; }
    u16* tabL = tab.data(); u16* tabS = tab.data() + (1u << hlogL);
    {
      // every block is an independent unit of the match kernel: cleared tables, primed with the 16 KiB before the block
      std::fill(tab.begin(), tab.end(), (u16)0);
      const u32 b0 = (u32)(bstart - src);
      if (blk > 0) for (u32 q = b0 - 16384u; q < b0; q++) {        // ascending: the last position of a hash value is the one kept
        const u64 v = ld64(src + q);
        tabL[h64(v, hlogL, dfast ? 8 : mls)] = (u16)q; if (dfast) tabS[h64(v, hlogS, mls)] = (u16)q;
      }
    }
    u32 anchor = bpos; const u32 ilimit = bend - 8; u32 p0 = bpos + (bpos == 0 ? 1 : 0);
    while (p0 < ilimit) {
      const u32 step = 1 + ((p0 - anchor) >> 8);
      u32 p[32], hL[32], hS[32], candL[32], candS[32]; bool act[32], vL[32], vS[32], okL[32], okS[32], rok[32]; u64 v[32];
      for (u32 l = 0; l < 32; l++) {
        p[l] = p0 + l * step; act[l] = p[l] < ilimit; v[l] = act[l] ? ld64(src + p[l]) : 0;
        hL[l] = act[l] ? h64(v[l], hlogL, dfast ? 8 : mls) : (0x80000000u | l);
        hS[l] = act[l] ? h64(v[l], hlogS, mls) : (0x80000000u | l);
      }
      for (u32 l = 0; l < 32; l++) {
        i32 low = -1; for (i32 k = (i32)l - 1; k >= 0; k--) if (hL[k] == hL[l]) { low = k; break; }
        if (low >= 0) { candL[l] = p[low]; vL[l] = act[l]; } else vL[l] = act[l] && candFrom(p[l], tabL[hL[l] & 0xFFFFFF], candL[l]);
        vS[l] = false;
        if (dfast) { low = -1; for (i32 k = (i32)l - 1; k >= 0; k--) if (hS[k] == hS[l]) { low = k; break; }
          if (low >= 0) { candS[l] = p[low]; vS[l] = act[l]; } else vS[l] = act[l] && candFrom(p[l], tabS[hS[l] & 0xFFFFFF], candS[l]); }
      }
      i32 fl = -1;
      for (u32 l = 0; l < 32; l++) {
        okL[l] = vL[l] && (dfast ? ld64(src + candL[l]) == v[l] : ld32(src + candL[l]) == (u32)v[l]);
        okS[l] = dfast && vS[l] && ld32(src + candS[l]) == (u32)v[l];
        rok[l] = act[l] && rep1 != 0 && rep1 <= p[l] && ld32(src + p[l] - rep1) == (u32)v[l];
        if (fl < 0 && (okL[l] || okS[l] || rok[l])) fl = (i32)l;
      }
      // a repeat-offset match a few bytes further on beats a table match here (cheaper to code)
      if (fl >= 0 && !rok[fl]) { const int look = 3; for (i32 l = fl + 1; l <= fl + look && l < 32; l++) if (rok[l]) { fl = l; break; } }
      auto insert = [&](u32 limit) {   // ascending lanes: last writer wins; nothing at or beyond the restart position
        for (u32 l = 0; l < 32; l++) if (act[l] && p[l] < limit) { tabL[hL[l]] = (u16)p[l]; if (dfast) tabS[hS[l]] = (u16)p[l]; } };
      if (fl < 0) { insert(0xFFFFFFFFu); p0 += 32 * step; continue; }
      u32 pos = p[fl]; const u32 kind = rok[fl] ? 0 : (okL[fl] ? 1 : 2);
      u32 cnd = kind == 0 ? pos - rep1 : (kind == 1 ? candL[fl] : candS[fl]);
      const u32 off = pos - cnd;
      u32 mlen = 4 + extend(pos + 4, off, bend);
      while (pos > anchor && cnd > 0 && src[pos - 1] == src[cnd - 1]) { pos--; cnd--; mlen++; }
      insert(pos + mlen);
      const u32 ll = pos - anchor; u32 offBase;
      if (kind == 0 && ll > 0) offBase = 1; else { offBase = off + 3; rep2 = rep1; rep1 = off; }
      memcpy(lits.data() + nlits, src + anchor, ll);
      seqs[2 * nseq] = (ll & 0xFFFF) | (((mlen - 3) & 0xFFFF) << 16);
      seqs[2 * nseq + 1] = (offBase & 0x3FFFFFFFu) | ((((mlen - 3) >> 16) & 1) << 30) | (((ll >> 16) & 1) << 31);
      nlits += ll; nseq++;
      anchor = pos + mlen; p0 = anchor;
      if (p0 <= ilimit) {

        while (rep2 != 0 && p0 <= ilimit && ld32(src + p0) == ld32(src + p0 - rep2)) {
          const u32 rlen = 4 + extend(p0 + 4, rep2, bend);
          std::swap(rep1, rep2);
          tabL[h64(ld64(src + p0), hlogL, dfast ? 8 : mls)] = (u16)p0; if (dfast) tabS[h64(ld64(src + p0), hlogS, mls)] = (u16)p0;
          seqs[2 * nseq] = ((rlen - 3) & 0xFFFF) << 16; seqs[2 * nseq + 1] = 1u | ((((rlen - 3) >> 16) & 1) << 30);
          nseq++; p0 += rlen; anchor = p0;
        }
      }
    }
    const u32 ll = bend - anchor; memcpy(lits.data() + nlits, src + anchor, ll); nlits += ll;
    st.seqs = seqs.data(); st.n = nseq; st.cap = BLOCKSIZE_MAX / 4 + 64; st.lits = lits.data(); st.nlits = nlits;
    return true;
  }
  void done(bool) {}
};
}  // namespace

// batch_max: the largest chunk of the call the frame is part of (the kernels choose their table sizes per call)
extern "C" uint32_t hostsim_compress_warp2(uint8_t* dst, uint32_t cap, const uint8_t* src_in, uint32_t size, int level, int checksum, uint32_t batch_max) {
  std::vector<u8> padded(size + 64, 0);
  u8* src = padded.data() + 16;
  if (size) memcpy(src, src_in, size);
  const bool big = batch_max > BLOCKSIZE_MAX;
  WarpMatcher m; m.src = src; m.size = size; m.dfast = level >= 3; m.hlogL = getenv("HS_HLOGL") ? atoi(getenv("HS_HLOGL")) : enc_hlog_long(level, big); m.hlogS = getenv("HS_HLOGS") ? atoi(getenv("HS_HLOGS")) : enc_hlog_short(level, big); m.mls = level <= 1 ? 6 : 5;
  m.tab.assign((1u << m.hlogL) + (1u << m.hlogS), 0);
  const u32 seqCap = BLOCKSIZE_MAX / 4 + 64;
  std::vector<u8> codes(3 * seqCap), sym(4096); std::vector<u16> ct(3 * 514 + 16);
  return encode_frame_with(src, size, dst, cap, level, checksum, codes.data(), ct.data(), sym.data(), m);
}
extern "C" uint32_t hostsim_compress_warp(uint8_t* dst, uint32_t cap, const uint8_t* src_in, uint32_t size, int level, int checksum) {
  return hostsim_compress_warp2(dst, cap, src_in, size, level, checksum, size);
}

// ---- dictionary of the replay (the device's zstdb200_load_dictionary): size 0 removes it ----
extern "C" void hostsim_set_dict(const uint8_t* dict, uint32_t size) {
  g_haveDict = false;
  if (!dict || size == 0) return;
  g_dictBytes.assign(dict, dict + size); g_dictBytes.resize(size + 64, 0);
  static HufBuildWk wk; static HufFseScratch fs; static u8 slot[260]; static s16 norm[64]; static u16 next[64];
  dict_load(g_dictBytes.data(), size, g_dict, wk, fs, slot, norm, next);
  g_haveDict = true;
}
