// serial_encoder.h — serial CPU replay of the GPU encoder (test infrastructure, NOT part of the product).
//
// zstandard_b200/csrc/zb_encode.cuh holds the entropy-stage building blocks the kernels use (bit writer, FSE / Huffman
// table construction, header writers).  This file adds what only the CPU replay needs: serial "fast" / "double-fast"
// match finders, the serial literals / sequences section writers and the frame loop — so that tests can (a) produce
// frames with the same entropy code on a GPU-less box and (b) check the kernels' warp-parallel writers against a
// straightforward serial statement of the same format rules (byte equality, tests/test_encode_gpu.py).
#pragma once
#include "../../zstandard_b200/csrc/zb_encode.cuh"

namespace zb {

// ------------------------------------------------------------------------------------------------
// parameters per level (libzstd's table for sources <= 128 KiB; hash logs shrink with the input)
// ------------------------------------------------------------------------------------------------
struct EncParams { u32 hashLog, chainLog /* short table of double-fast */, minMatch; bool dfast; };

ZB_HD EncParams enc_params(int level, u32 srcSize) {
  EncParams p;
  if (level <= 1) { p.hashLog = 13; p.chainLog = 0; p.minMatch = 6; p.dfast = false; }
  else if (level == 2) { p.hashLog = 15; p.chainLog = 0; p.minMatch = 5; p.dfast = false; }
  else { p.hashLog = 16; p.chainLog = 15; p.minMatch = 5; p.dfast = true; }
  u32 srcLog = srcSize < 64 ? 6 : highbit(srcSize - 1) + 1;
  if (p.hashLog > srcLog + 1) p.hashLog = srcLog + 1;
  if (p.chainLog > srcLog) p.chainLog = srcLog;
  return p;
}

ZB_HD u32 enc_table_words(int level) { return level <= 1 ? (1u << 13) : (level == 2 ? (1u << 15) : (1u << 16) + (1u << 15)); }

// per-frame scratch layout (all in HBM; addressed by the kernel from the item index)
struct EncScratch {
  u32* table;      // enc_table_words(level) position entries (0 = empty; positions are stored +1)
  u8* lits;        // BLOCKSIZE_MAX bytes
  u32* seqs;       // per sequence: litLength, matchLength-3 | offCode... packed as 2 words (see seq_push)
  u32 seqCap;      // sequences
  u8* codes;       // 3 * seqCap bytes: llCode, ofCode, mlCode per sequence
  u16* ctables;    // FSE state tables scratch: 3 * 512 u16
  u8* tmp;         // BLOCKSIZE_MAX + 1024 bytes: block assembled here before the raw/compressed decision
};

ZB_HD size_t enc_scratch_bytes_per_frame(int level) {
  return (size_t)enc_table_words(level) * 4 + BLOCKSIZE_MAX + (size_t)(BLOCKSIZE_MAX / 4 + 64) * (8 + 3) + 3 * 512 * 2 + BLOCKSIZE_MAX + 2048;
}

// hashes of the first mls bytes at p (zstd's multiplicative hashes)
ZB_HD u32 hash_bytes(const u8* p, u32 hlog, u32 mls) {
  if (mls >= 8) return (u32)((rd64u(p) * 0xCF1BBCDCB7A56463ull) >> (64 - hlog));
  if (mls == 7) return (u32)(((rd64u(p) << 8) * 0xCF1BBCDCBFA563ull) >> (64 - hlog));
  if (mls == 6) return (u32)(((rd64u(p) << 16) * 0xCF1BBCDCBF9Bull) >> (64 - hlog));
  if (mls == 5) return (u32)(((rd64u(p) << 24) * 0xCF1BBCDCBBull) >> (64 - hlog));
  return (rd32u(p) * 2654435761u) >> (32 - hlog);
}

// number of equal bytes at a and b, both readable up to `end` on a's side
ZB_HD u32 count_match(const u8* a, const u8* b, const u8* aend) {
  const u8* s = a;
  while (a + 8 <= aend) {
    u64 d = rd64u(a) ^ rd64u(b);
    if (d) {
#if defined(__CUDA_ARCH__)
      return (u32)(a - s) + ((u32)__ffsll((long long)d) - 1) / 8;
#else
      return (u32)(a - s) + (u32)__builtin_ctzll(d) / 8;
#endif
    }
    a += 8; b += 8;
  }
  while (a < aend && *a == *b) { a++; b++; }
  return (u32)(a - s);
}

// offBase: 1..3 = repeat codes, >= 4 = offset + 3 (zstd's "offBase" convention)
ZB_HD void seq_push(SeqStore& st, const u8* litSrc, u32 ll, u32 offBase, u32 ml) {
  for (u32 i = 0; i < ll; i++) st.lits[st.nlits + i] = litSrc[i];
  st.nlits += ll;
  st.seqs[2 * st.n] = (ll & 0xFFFF) | (((ml - 3) & 0xFFFF) << 16);   // both lengths need 17 bits: bit 16 of each lives in word 1
  st.seqs[2 * st.n + 1] = (offBase & 0x3FFFFFFFu) | ((((ml - 3) >> 16) & 1) << 30) | (((ll >> 16) & 1) << 31);
  st.n++;
}

// ------------------------------------------------------------------------------------------------
// match finders.  base = frame start, [istart, iend) = the block; window = everything since base.
// rep[0..1] carried across blocks.  Positions in the tables are (index from base) + 1, 0 = empty.
// ------------------------------------------------------------------------------------------------
// "fast" strategy in its pipelined form: two positions are probed per round, the repeat offset is tried two
// bytes ahead *before* the hash candidate of the current position, and the stride grows by one every 128 bytes
// without a match (incompressible runs are skimmed).  The look-ahead position ip1 is only entered into the
// table while the stride is small: with a large stride it can lie beyond the end of the match just found, and
// an entry at or after the restart position would later be found as its own candidate (offset 0).
ZB_HD void match_fast(SeqStore& st, u32* table, u32 hlog, u32 mls, const u8* base, const u8* istart, const u8* iend, u32 rep[2]) {
  const u8* ip0 = istart; const u8* anchor = istart;
  const u8* const ilimit = iend - 8;
  u32 off1 = rep[0], off2 = rep[1], saved = 0;
  const bool run = iend - istart >= 16;
  if (run && ip0 == base) ip0++;
  { u32 maxRep = (u32)(ip0 - base); if (off2 > maxRep) { saved = off2; off2 = 0; } if (off1 > maxRep) { saved = off1; off1 = 0; } }
  while (run) {
    u32 step = 2; const u8* nextStep = ip0 + 128;
    const u8 *ip1 = ip0 + 1, *ip2 = ip0 + step, *ip3 = ip2 + 1;
    if (ip3 >= ilimit) break;
    u32 hash0 = hash_bytes(ip0, hlog, mls), hash1 = hash_bytes(ip1, hlog, mls);
    u32 idx = table[hash0], cur0 = 0, found = 0, mlen = 0, offBase = 0;
    const u8* match0 = nullptr;
    do {
      const u32 rval = off1 ? rd32u(ip2 - off1) : 0;
      cur0 = (u32)(ip0 - base); table[hash0] = cur0 + 1;
      if (off1 > 0 && rd32u(ip2) == rval) {                         // repeat offset two bytes ahead
        ip0 = ip2; match0 = ip0 - off1; mlen = ip0[-1] == match0[-1]; ip0 -= mlen; match0 -= mlen; offBase = 1; mlen += 4;
        table[hash1] = (u32)(ip1 - base) + 1; found = 1; break;
      }
      if (idx != 0 && rd32u(base + idx - 1) == rd32u(ip0)) { if (step <= 4) table[hash1] = (u32)(ip1 - base) + 1; found = 2; break; }
      idx = table[hash1]; hash0 = hash1; hash1 = hash_bytes(ip2, hlog, mls);
      ip0 = ip1; ip1 = ip2; ip2 = ip3;
      cur0 = (u32)(ip0 - base); table[hash0] = cur0 + 1;
      if (idx != 0 && rd32u(base + idx - 1) == rd32u(ip0)) { if (step <= 4) table[hash1] = (u32)(ip1 - base) + 1; found = 2; break; }
      idx = table[hash1]; hash0 = hash1; hash1 = hash_bytes(ip2, hlog, mls);
      ip0 = ip1; ip1 = ip2; ip2 = ip0 + step; ip3 = ip1 + step;
      if (ip2 >= nextStep) { step++; nextStep += 128; }
    } while (ip3 < ilimit);
    if (!found) break;
    if (found == 2) {
      match0 = base + idx - 1; off2 = off1; off1 = (u32)(ip0 - match0); offBase = off1 + 3; mlen = 4;
      while (ip0 > anchor && match0 > base && ip0[-1] == match0[-1]) { ip0--; match0--; mlen++; }   // catch up
    }
    mlen += count_match(ip0 + mlen, match0 + mlen, iend);
    seq_push(st, anchor, (u32)(ip0 - anchor), offBase, mlen);
    ip0 += mlen; anchor = ip0;
    if (ip0 <= ilimit) {
      table[hash_bytes(base + cur0 + 2, hlog, mls)] = cur0 + 2 + 1;
      table[hash_bytes(ip0 - 2, hlog, mls)] = (u32)(ip0 - 2 - base) + 1;
      while (off2 > 0 && ip0 <= ilimit && rd32u(ip0) == rd32u(ip0 - off2)) {   // immediate repeat of the older offset
        const u32 rlen = count_match(ip0 + 4, ip0 + 4 - off2, iend) + 4;
        { u32 t = off2; off2 = off1; off1 = t; }
        table[hash_bytes(ip0, hlog, mls)] = (u32)(ip0 - base) + 1;
        seq_push(st, anchor, 0, 1, rlen);
        ip0 += rlen; anchor = ip0;
      }
    }
  }
  rep[0] = off1 ? off1 : saved; rep[1] = off2 ? off2 : saved;
  { u32 ll = (u32)(iend - anchor); for (u32 i = 0; i < ll; i++) st.lits[st.nlits + i] = anchor[i]; st.nlits += ll; }
}

ZB_HD void match_dfast(SeqStore& st, u32* hashLong, u32 hlogL, u32* hashSmall, u32 hlogS, u32 mls, const u8* base, const u8* istart,
                       const u8* iend, u32 rep[2]) {
  const u8* ip = istart; const u8* anchor = istart;
  const u8* const ilimit = iend - 8;
  u32 off1 = rep[0], off2 = rep[1], saved = 0;
  const bool run = iend - istart >= 16;
  if (run && ip == base) ip++;
  { u32 maxRep = (u32)(ip - base); if (off2 > maxRep) { saved = off2; off2 = 0; } if (off1 > maxRep) { saved = off1; off1 = 0; } }
  while (run && ip < ilimit) {
    u32 mlen;
    const u32 h2 = hash_bytes(ip, hlogL, 8), h = hash_bytes(ip, hlogS, mls);
    const u32 cur = (u32)(ip - base);
    const u32 miL = hashLong[h2], miS = hashSmall[h];
    hashLong[h2] = hashSmall[h] = cur + 1;
    if (off1 > 0 && rd32u(ip + 1 - off1) == rd32u(ip + 1)) {
      mlen = count_match(ip + 1 + 4, ip + 1 + 4 - off1, iend) + 4;
      ip++;
      seq_push(st, anchor, (u32)(ip - anchor), 1, mlen);
    } else {
      u32 offset; const u8* match;
      const u8* mL = base + miL - 1; const u8* mS = base + miS - 1;
      if (miL != 0 && rd64u(mL) == rd64u(ip)) {
        mlen = count_match(ip + 8, mL + 8, iend) + 8; match = mL;
        while (ip > anchor && match > base && ip[-1] == match[-1]) { ip--; match--; mlen++; }
      } else if (miS != 0 && rd32u(mS) == rd32u(ip)) {
        // a short match: try the long table one position later first
        const u32 hl3 = hash_bytes(ip + 1, hlogL, 8);
        const u32 mi3 = hashLong[hl3];
        hashLong[hl3] = cur + 1 + 1;
        const u8* m3 = base + mi3 - 1;
        if (mi3 != 0 && rd64u(m3) == rd64u(ip + 1)) {
          mlen = count_match(ip + 9, m3 + 8, iend) + 8; ip++; match = m3;
          while (ip > anchor && match > base && ip[-1] == match[-1]) { ip--; match--; mlen++; }
        } else {
          mlen = count_match(ip + 4, mS + 4, iend) + 4; match = mS;
          while (ip > anchor && match > base && ip[-1] == match[-1]) { ip--; match--; mlen++; }
        }
      } else { ip += ((ip - anchor) >> 8) + 1; continue; }
      offset = (u32)(ip - match);
      off2 = off1; off1 = offset;
      seq_push(st, anchor, (u32)(ip - anchor), offset + 3, mlen);
    }
    ip += mlen; anchor = ip;
    if (ip <= ilimit) {
      hashLong[hash_bytes(base + cur + 2, hlogL, 8)] = hashSmall[hash_bytes(base + cur + 2, hlogS, mls)] = cur + 2 + 1;
      hashLong[hash_bytes(ip - 2, hlogL, 8)] = hashSmall[hash_bytes(ip - 2, hlogS, mls)] = (u32)(ip - 2 - base) + 1;
      while (ip <= ilimit && off2 > 0 && rd32u(ip) == rd32u(ip - off2)) {
        const u32 rlen = count_match(ip + 4, ip + 4 - off2, iend) + 4;
        { u32 t = off2; off2 = off1; off1 = t; }
        hashSmall[hash_bytes(ip, hlogS, mls)] = hashLong[hash_bytes(ip, hlogL, 8)] = (u32)(ip - base) + 1;
        seq_push(st, anchor, 0, 1, rlen);
        ip += rlen; anchor = ip;
      }
    }
  }
  rep[0] = off1 ? off1 : saved; rep[1] = off2 ? off2 : saved;
  { u32 ll = (u32)(iend - anchor); for (u32 i = 0; i < ll; i++) st.lits[st.nlits + i] = anchor[i]; st.nlits += ll; }
}

ZB_HD bool huf_build(HufEnc& he, const u32* count, u32 maxSym, u32 maxBits, HufBuildScratch& sc) {
  const u32 n = huf_sort_symbols(sc.order, count, maxSym);
  return huf_build_sorted(he, count, n, maxBits, sc);
}

// literals section (DecodeLiteralsBlock ZStdDecompress.cs:683-821).  Returns bytes written (0 = no room).
ZB_HD u32 enc_literals(u8* out, u32 cap, const u8* lits, u32 n, u16* stateScratch, u8* symScratch) {
  auto raw = [&]() -> u32 {
    const u32 lh = n < 32 ? 1 : (n < 4096 ? 2 : 3);
    if (lh + n > cap) return 0;
    if (lh == 1) out[0] = (u8)(n << 3); else if (lh == 2) { u32 v = (1u << 2) | (n << 4); out[0] = (u8)v; out[1] = (u8)(v >> 8); }
    else { u32 v = (3u << 2) | (n << 4); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); }
    for (u32 i = 0; i < n; i++) out[lh + i] = lits[i];
    return lh + n;
  };
  if (n < 64) return raw();
  u32 count[256]; for (u32 i = 0; i < 256; i++) count[i] = 0;
  for (u32 i = 0; i < n; i++) count[lits[i]]++;
  u32 maxSym = 255; while (maxSym > 0 && !count[maxSym]) maxSym--;
  u32 largest = 0; for (u32 s = 0; s <= maxSym; s++) if (count[s] > largest) largest = count[s];
  if (largest == n) {   // rle literals
    const u32 lh = n < 32 ? 1 : (n < 4096 ? 2 : 3);
    if (lh + 1 > cap) return 0;
    if (lh == 1) out[0] = (u8)(1 | (n << 3)); else if (lh == 2) { u32 v = 1 | (1u << 2) | (n << 4); out[0] = (u8)v; out[1] = (u8)(v >> 8); }
    else { u32 v = 1 | (3u << 2) | (n << 4); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); }
    out[lh] = lits[0];
    return lh + 1;
  }
  if (largest <= (n >> 7) + 4) return raw();   // too flat to be worth it
  HufEnc he; HufBuildScratch hsc;
  u32 maxBits = fse_optimal_log(11, n, maxSym, 1); if (maxBits > 11) maxBits = 11;
  if (!huf_build(he, count, maxSym, maxBits, hsc)) return raw();
  const bool single = n < 256;
  const u32 lhSize = 3 + (n >= 1024) + (n >= 16384);
  if (lhSize + 8 > cap) return 0;
  u8* body = out + lhSize; const u32 bodyCap = cap - lhSize;
  u32 hdr = huf_write_header(body, bodyCap, he, stateScratch, symScratch);
  if (!hdr) return raw();
  u32 csize = hdr;
  if (single) {
    u32 s = huf_encode_stream(body + csize, bodyCap - csize, lits, n, he);
    if (!s) return raw();
    csize += s;
  } else {
    const u32 seg = (n + 3) / 4;
    if (csize + 6 > bodyCap) return raw();
    u8* jump = body + csize; csize += 6;
    for (u32 k = 0; k < 4; k++) {
      const u32 from = k * seg, len = k < 3 ? seg : n - 3 * seg;
      u32 s = huf_encode_stream(body + csize, bodyCap - csize, lits + from, len, he);
      if (!s || s > 65535) return raw();
      if (k < 3) { jump[2 * k] = (u8)s; jump[2 * k + 1] = (u8)(s >> 8); }
      csize += s;
    }
  }
  const u32 minGain = (n >> 6) + 2;
  if (csize + minGain >= n) return raw();
  // header: type 2 (compressed), size format by lhSize
  if (lhSize == 3) { u32 v = 2 | ((single ? 0u : 1u) << 2) | (n << 4) | (csize << 14); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); }
  else if (lhSize == 4) { u32 v = 2 | (2u << 2) | (n << 4) | (csize << 18); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); out[3] = (u8)(v >> 24); }
  else { u32 v = 2 | (3u << 2) | (n << 4) | (csize << 22); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); out[3] = (u8)(v >> 24); out[4] = (u8)(csize >> 10); }
  return lhSize + csize;
}

ZB_HD u32 enc_seq_table(FseCTable& ct, u16* stateTable, u8* symScratch, const u8* codes, u32 nbSeq, const SeqKind& k, int level,
                        u8* out, u32 cap, u32* used) {
  u32 count[53]; for (u32 i = 0; i <= k.maxSym; i++) count[i] = 0;
  for (u32 i = 0; i < nbSeq; i++) count[codes[i]]++;
  return enc_seq_table_counts(ct, stateTable, symScratch, count, codes[nbSeq - 1], nbSeq, k, level, out, cap, used);
}

// bitstream: sequences last to first; per sequence the decoder reads offset, matchLength, litLength extra
// bits, then the LL, ML, OF state bits (:1504-1550) — so we write them in the opposite order.
// Returns the end of the stream or nullptr when out of room.
// The encoder's running state, so that callers can feed the sequences in pieces (the GPU kernel stages them
// through shared memory a chunk at a time).
struct SeqBits { BitWriter w; u32 sLL, sOF, sML; };

ZB_HD void seqbits_first(SeqBits& b, u8* out, u8* end, const FseCTable& ctLL, const FseCTable& ctOF, const FseCTable& ctML,
                         u32 ll, u32 ob, u32 mlm3, u32 lc, u32 oc, u32 mc) {   // the last sequence of the block
  bw_init(b.w, out, end);
  fse_init_state(ctML, b.sML, mc); fse_init_state(ctOF, b.sOF, oc); fse_init_state(ctLL, b.sLL, lc);
  bw_add(b.w, ll - kLLbase[lc], kLLbits[lc]);
  bw_add(b.w, mlm3 + 3 - kMLbase[mc], kMLbits[mc]); bw_flush(b.w);
  bw_add(b.w, ob - (1u << oc), oc); bw_flush(b.w);
}

ZB_HD void seqbits_next(SeqBits& b, const FseCTable& ctLL, const FseCTable& ctOF, const FseCTable& ctML,
                        u32 ll, u32 ob, u32 mlm3, u32 lc, u32 oc, u32 mc) {    // the others, last to first
  fse_encode(b.w, ctOF, b.sOF, oc); fse_encode(b.w, ctML, b.sML, mc); bw_flush(b.w);
  fse_encode(b.w, ctLL, b.sLL, lc);
  bw_add(b.w, ll - kLLbase[lc], kLLbits[lc]); bw_flush(b.w);
  bw_add(b.w, mlm3 + 3 - kMLbase[mc], kMLbits[mc]); bw_flush(b.w);
  bw_add(b.w, ob - (1u << oc), oc); bw_flush(b.w);
}

ZB_HD u8* seqbits_finish(SeqBits& b, const FseCTable& ctLL, const FseCTable& ctOF, const FseCTable& ctML) {
  fse_flush_state(b.w, ctML, b.sML); fse_flush_state(b.w, ctOF, b.sOF); fse_flush_state(b.w, ctLL, b.sLL);
  return bw_close(b.w);
}

ZB_HD u8* enc_seq_bitstream(u8* out, u8* end, const SeqStore& st, const u8* llc, const u8* ofc, const u8* mlc,
                            const FseCTable& ctLL, const FseCTable& ctOF, const FseCTable& ctML) {
  const u32 nbSeq = st.n;
  SeqBits b;
  { const u32 i = nbSeq - 1; u32 ll, ob, mlm3; seq_get(st, i, ll, ob, mlm3);
    seqbits_first(b, out, end, ctLL, ctOF, ctML, ll, ob, mlm3, llc[i], ofc[i], mlc[i]); }
  for (i32 n = (i32)nbSeq - 2; n >= 0; n--) {
    u32 ll, ob, mlm3; seq_get(st, (u32)n, ll, ob, mlm3);
    seqbits_next(b, ctLL, ctOF, ctML, ll, ob, mlm3, llc[n], ofc[n], mlc[n]);
  }
  return seqbits_finish(b, ctLL, ctOF, ctML);
}

// Returns bytes written, 0 on failure (caller falls back to a raw block).
ZB_HD u32 enc_sequences(u8* out, u32 cap, const SeqStore& st, u8* codes, u16* ctables, u8* symScratch, int level) {
  const u32 nbSeq = st.n;
  if (cap < 4) return 0;
  u32 op = enc_seq_count_header(out, nbSeq);
  if (nbSeq == 0) return op;
  u8 *llc = codes, *ofc = codes + st.cap, *mlc = codes + 2 * st.cap;
  for (u32 i = 0; i < nbSeq; i++) {
    u32 ll, ob, mlm3; seq_get(st, i, ll, ob, mlm3);
    llc[i] = (u8)ll_code(ll); ofc[i] = (u8)highbit(ob); mlc[i] = (u8)ml_code(mlm3);
  }
  const SeqKind kLL = seq_kind(KIND_LL), kOF = seq_kind(KIND_OF), kML = seq_kind(KIND_ML);
  u8* modeByte = out + op++; u32 used;
  FseCTable ctLL, ctOF, ctML;
  const u32 mLL = enc_seq_table(ctLL, ctables, symScratch, llc, nbSeq, kLL, level, out + op, cap - op, &used); if (mLL == 0xFF) return 0; op += used;
  const u32 mOF = enc_seq_table(ctOF, ctables + 514, symScratch, ofc, nbSeq, kOF, level, out + op, cap - op, &used); if (mOF == 0xFF) return 0; op += used;
  const u32 mML = enc_seq_table(ctML, ctables + 1028, symScratch, mlc, nbSeq, kML, level, out + op, cap - op, &used); if (mML == 0xFF) return 0; op += used;
  *modeByte = (u8)((mLL << 6) | (mOF << 4) | (mML << 2));
  u8* e = enc_seq_bitstream(out + op, out + cap, st, llc, ofc, mlc, ctLL, ctOF, ctML);
  if (!e) return 0;
  return (u32)(e - out);
}

// ------------------------------------------------------------------------------------------------
// frame assembly (frame header ZStdDecompress.cs:421-499, block header :646-659)
// ------------------------------------------------------------------------------------------------
// Writes a complete frame for src[0..size) at dst (capacity cap).  Returns the frame size without the
// 4-byte content checksum slot (the checksum kernel fills it), or an error code.
//
// `blockSeqs(st, blockIndex, bstart, bsize)` supplies the block's sequence store: either by running a match
// finder right here (SerialMatcher: the thread-per-frame replay used by tests/hostsim) or by pointing at what
// an emulation of the warp-parallel match kernel (WarpMatcher, tests/hostsim).  k_enc_entropy mirrors this function
// warp-wide on what k_enc_match left in HBM (encode_kernels.cu).
// Block bodies are written straight into dst and replaced by a raw copy when they do not pay.
template <class BlockSeqs>
ZB_HD u32 encode_frame_with(const u8* src, u32 size, u8* dst, u32 cap, int level, int checksum, u8* codes, u16* ctables, u8* symScratch,
                            BlockSeqs& blockSeqs) {
  u32 op = 0;
  const u32 fcsCode = size < 256 ? 0 : (size < 65536 + 256 ? 1 : 2);
  const u32 fhs = 4 + 1 + (fcsCode == 0 ? 1 : (fcsCode == 1 ? 2 : 4));
  if (cap < fhs + 3 + (checksum ? 4 : 0)) return zerr(ZE_dstSize_tooSmall);
  dst[0] = 0x28; dst[1] = 0xB5; dst[2] = 0x2F; dst[3] = 0xFD;
  dst[4] = (u8)((fcsCode << 6) | (1u << 5) | (checksum ? 4 : 0));     // single segment, no dictionary
  if (fcsCode == 0) dst[5] = (u8)size;
  else if (fcsCode == 1) { const u32 v = size - 256; dst[5] = (u8)v; dst[6] = (u8)(v >> 8); }
  else { dst[5] = (u8)size; dst[6] = (u8)(size >> 8); dst[7] = (u8)(size >> 16); dst[8] = (u8)(size >> 24); }
  op = fhs;
  const u32 tail = checksum ? 4 : 0;
  u32 pos = 0, blk = 0;
  do {
    const u32 bsize = size - pos < BLOCKSIZE_MAX ? size - pos : BLOCKSIZE_MAX;
    const u32 last = pos + bsize == size;
    const u8* bstart = src + pos;
    if (op + 3 + tail > cap) return zerr(ZE_dstSize_tooSmall);
    bool rle = bsize > 0;
    for (u32 i = 1; i < bsize && rle; i++) if (bstart[i] != bstart[0]) rle = false;
    SeqStore st; st.n = 0; st.nlits = 0; st.seqs = nullptr; st.lits = nullptr; st.cap = 0;
    const bool haveSeqs = blockSeqs(st, blk, bstart, bsize, rle && bsize >= 2);   // always called: keeps matcher state in step
    if (rle && bsize >= 2) {
      if (op + 4 + tail > cap) return zerr(ZE_dstSize_tooSmall);
      const u32 h = last | (1u << 1) | (bsize << 3);
      dst[op] = (u8)h; dst[op + 1] = (u8)(h >> 8); dst[op + 2] = (u8)(h >> 16); dst[op + 3] = bstart[0]; op += 4;
    } else {
      u32 csize = 0; bool compressed = false;
      if (haveSeqs) {
        u32 room = bsize - 1 < BLOCKSIZE_MAX - 1 ? bsize - 1 : BLOCKSIZE_MAX - 1;   // must beat raw and stay < 128 KiB (:1880)
        const u32 avail = cap - op - 3 - tail;
        if (room > avail) room = avail;
        u8* body = dst + op + 3;
        const u32 l = enc_literals(body, room, st.lits, st.nlits, ctables, symScratch);
        if (l) {
          const u32 s = enc_sequences(body + l, room - l, st, codes, ctables, symScratch, level);
          if (s && l + s < bsize) { csize = l + s; compressed = true; }
        }
        blockSeqs.done(compressed);
      }
      if (compressed) {
        const u32 h = last | (2u << 1) | (csize << 3);
        dst[op] = (u8)h; dst[op + 1] = (u8)(h >> 8); dst[op + 2] = (u8)(h >> 16); op += 3 + csize;
      } else {
        if (op + 3 + bsize + tail > cap) return zerr(ZE_dstSize_tooSmall);
        const u32 h = last | (0u << 1) | (bsize << 3);
        dst[op] = (u8)h; dst[op + 1] = (u8)(h >> 8); dst[op + 2] = (u8)(h >> 16); op += 3;
        for (u32 i = 0; i < bsize; i++) dst[op + i] = bstart[i];
        op += bsize;
      }
    }
    pos += bsize; blk++;
  } while (pos < size);
  return op;
}

// match finding inside the encoding thread (thread-per-frame replay)
struct SerialMatcher {
  const EncScratch& sc; EncParams pr; const u8* base; u32 rep[2], savedRep[2];
  ZB_HD SerialMatcher(const EncScratch& s, int level, const u8* src, u32 size) : sc(s), pr(enc_params(level, size)), base(src) {
    const u32 tw = (1u << pr.hashLog) + (pr.dfast ? (1u << pr.chainLog) : 0);
    for (u32 i = 0; i < tw; i++) sc.table[i] = 0;
    rep[0] = 1; rep[1] = 4; savedRep[0] = 1; savedRep[1] = 4;
  }
  ZB_HD bool operator()(SeqStore& st, u32, const u8* bstart, u32 bsize, bool isRle) {
    if (isRle || bsize < 64) return false;
    st.seqs = sc.seqs; st.n = 0; st.cap = sc.seqCap; st.lits = sc.lits; st.nlits = 0;
    savedRep[0] = rep[0]; savedRep[1] = rep[1];
    if (pr.dfast) match_dfast(st, sc.table, pr.hashLog, sc.table + (1u << pr.hashLog), pr.chainLog, pr.minMatch, base, bstart, bstart + bsize, rep);
    else match_fast(st, sc.table, pr.hashLog, pr.minMatch, base, bstart, bstart + bsize, rep);
    return true;
  }
  ZB_HD void done(bool compressed) { if (!compressed) { rep[0] = savedRep[0]; rep[1] = savedRep[1]; } }   // a raw block leaves the decoder's history alone
};

ZB_HD u32 encode_frame(const u8* src, u32 size, u8* dst, u32 cap, int level, int checksum, const EncScratch& sc) {
  SerialMatcher m(sc, level, src, size);
  return encode_frame_with(src, size, dst, cap, level, checksum, sc.codes, sc.ctables, sc.tmp, m);
}

}  // namespace zb
