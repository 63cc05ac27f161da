// zb_encode.cuh — per-thread body of the GPU zstd frame encoder (levels 1-3).
//
// The reference repository ships no compressor: what binds this code is the reference
// *decoder's* accept set — every structure written here is one that csharp/src/ZStdDecompress.cs,
// HufDecompress.cs, EntropyCommon.cs and FseDecompress.cs parse (citations at each writer) — plus the ratio
// band against libzstd at the same level.  The match finders follow the published zstd "fast" (levels 1-2) and
// "double-fast" (level 3) strategies: greedy parse over one (or a long+short pair of) position hash table(s),
// repeat-offset probing, step acceleration through incompressible runs.
//
// One GPU thread encodes one frame: frames are independent (the path shards by frame), a frame's blocks are
// sequentially dependent through the window, the repeat offsets and the hash tables.  Everything is
// __host__ __device__ so that tests/hostsim replays the same code on the CPU.
#pragma once
#include "zb_format.cuh"

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

ZB_HD u64 rd64u(const u8* p) { return ld64(p); }
ZB_HD u32 rd32u(const u8* p) { return ld32(p); }

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

// ------------------------------------------------------------------------------------------------
// sequence store of one block
// ------------------------------------------------------------------------------------------------
struct SeqStore {
  u32* seqs; u32 n, cap;       // word0 = litLength, word1 = (matchLength - 3) | offCodeLow... see push
  u8* lits; u32 nlits;
};
// offBase: 1..3 = repeat codes, >= 4 = offset + 3 (zstd's "offBase" convention)
ZB_HD void seq_push(SeqStore& st, const u8* litSrc, u32 ll, u32 offBase, u32 ml) {
  for (u32 i = 0; i < ll; i++) st.lits[st.nlits + i] = litSrc[i];
  st.nlits += ll;
  st.seqs[2 * st.n] = (ll & 0xFFFF) | (((ml - 3) & 0xFFFF) << 16);   // both lengths need 17 bits: bit 16 of each lives in word 1
  st.seqs[2 * st.n + 1] = (offBase & 0x3FFFFFFFu) | ((((ml - 3) >> 16) & 1) << 30) | (((ll >> 16) & 1) << 31);
  st.n++;
}
ZB_HD void seq_get(const SeqStore& st, u32 i, u32& ll, u32& offBase, u32& mlm3) {
  u32 a = st.seqs[2 * i], b = st.seqs[2 * i + 1];
  ll = (a & 0xFFFF) | ((b >> 31) << 16);
  mlm3 = (a >> 16) | (((b >> 30) & 1) << 16);
  offBase = b & 0x3FFFFFFFu;
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

// ------------------------------------------------------------------------------------------------
// forward bit writer (the decoder reads it backwards: BitStream.cs:322-497).  Bits are appended LSB first.
// ------------------------------------------------------------------------------------------------
// Forward bit writer that stores whole aligned 32-bit words: p is 4-byte aligned, acc holds the nb pending bits
// (when the stream starts off alignment, the bytes already in memory before it ride along as the first pending
// bits and are stored back unchanged).  Contract: at most 32 bits are added between two bw_flush calls.
struct BitWriter { u8* p; u8* end; u64 acc; u32 nb; bool ovf; };
ZB_HD u32 bw_load32(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
ZB_HD void bw_init(BitWriter& w, u8* p, u8* end) {
  const u32 a = (u32)((uintptr_t)p & 3);
  w.p = p - a; w.end = end; w.nb = 8 * a; w.ovf = false;
  w.acc = a ? (bw_load32(w.p) & ((1u << (8 * a)) - 1)) : 0;
}
ZB_HD void bw_add(BitWriter& w, u32 v, u32 n) { w.acc |= (u64)(v & ((n >= 32) ? 0xFFFFFFFFu : ((1u << n) - 1))) << w.nb; w.nb += n; }
ZB_HD void bw_flush(BitWriter& w) {   // keeps < 32 bits pending
  if (w.nb >= 32) {
    const u32 v = (u32)w.acc;
    if (w.p + 4 <= w.end) {
#if defined(__CUDA_ARCH__)
      *reinterpret_cast<u32*>(w.p) = v;
#else
      w.p[0] = (u8)v; w.p[1] = (u8)(v >> 8); w.p[2] = (u8)(v >> 16); w.p[3] = (u8)(v >> 24);
#endif
    } else { for (u32 k = 0; k < 4; k++) { if (w.p + k < w.end) w.p[k] = (u8)(v >> (8 * k)); else w.ovf = true; } }
    w.p += 4; w.acc >>= 32; w.nb -= 32;
  }
}
// end mark: a 1 bit, then zero padding (InitDStream locates it through the highest set bit of the last byte)
ZB_HD u8* bw_close(BitWriter& w) {
  bw_add(w, 1, 1); bw_flush(w);
  const u32 bytes = (w.nb + 7) >> 3;
  for (u32 k = 0; k < bytes; k++) { if (w.p + k < w.end) w.p[k] = (u8)(w.acc >> (8 * k)); else w.ovf = true; }
  w.p += bytes; w.nb = 0;
  return w.ovf ? nullptr : w.p;
}

// ------------------------------------------------------------------------------------------------
// FSE: normalisation, header, encoding table (accepted by ReadNCount EntropyCommon.cs:79-188 and
// BuildFSETable ZStdDecompress.cs:958-1034 / BuildDTable FseDecompress.cs:111-181)
// ------------------------------------------------------------------------------------------------
ZB_HD u32 fse_optimal_log(u32 maxLog, u32 total, u32 maxSym, u32 minus = 2) {
  u32 maxBitsSrc = highbit(total - 1) - minus;
  u32 minBitsSrc = highbit(total) + 1, minBitsSym = highbit(maxSym) + 2;
  u32 minBits = minBitsSrc < minBitsSym ? minBitsSrc : minBitsSym;
  u32 tl = maxLog;
  if (maxBitsSrc < tl) tl = maxBitsSrc;
  if (minBits > tl) tl = minBits;
  if (tl < 5) tl = 5;
  if (tl > 12) tl = 12;
  return tl;
}

// proportional normalisation to a sum of 1 << tableLog; every present symbol gets >= 1 (-1 marks "less than
// one" and costs a full tableLog-bit state, as in the format)
ZB_HD bool fse_normalize(s16* norm, u32 tableLog, const u32* count, u32 total, u32 maxSym) {
  const u64 scale = 62 - tableLog, step = ((u64)1 << 62) / total, vStep = (u64)1 << (scale - 20);
  const u32 lowThreshold = total >> tableLog;
  i32 stillToDistribute = 1 << tableLog; u32 largest = 0; s16 largestP = 0;
  static const u32 rtb[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
  for (u32 s = 0; s <= maxSym; s++) {
    if (count[s] == total) return false;   // rle: caller handles
    if (count[s] == 0) { norm[s] = 0; continue; }
    if (count[s] <= lowThreshold) { norm[s] = -1; stillToDistribute--; }
    else {
      s16 proba = (s16)((count[s] * step) >> scale);
      if (proba < 8) { u64 restToBeat = vStep * rtb[proba]; proba += (count[s] * step) - ((u64)proba << scale) > restToBeat; }
      if (proba > largestP) { largestP = proba; largest = s; }
      norm[s] = proba; stillToDistribute -= proba;
    }
  }
  if (-stillToDistribute >= (norm[largest] >> 1)) {
    // corner case: redistribute with the secondary method
    const s16 NOT_YET = -2; u32 distributed = 0; u32 ToDistribute;
    const u32 lowOne = (u32)((total * 3ull) >> (tableLog + 1));
    u32 tot = total;
    for (u32 s = 0; s <= maxSym; s++) {
      if (count[s] == 0) { norm[s] = 0; continue; }
      if (count[s] <= lowThreshold) { norm[s] = -1; distributed++; tot -= count[s]; continue; }
      if (count[s] <= lowOne) { norm[s] = 1; distributed++; tot -= count[s]; continue; }
      norm[s] = NOT_YET;
    }
    ToDistribute = (1u << tableLog) - distributed;
    if (ToDistribute == 0) return true;
    if ((tot / ToDistribute) > lowOne) {
      const u32 lowOne2 = (u32)((tot * 3ull) / (ToDistribute * 2));
      for (u32 s = 0; s <= maxSym; s++) if (norm[s] == NOT_YET && count[s] <= lowOne2) { norm[s] = 1; distributed++; tot -= count[s]; }
      ToDistribute = (1u << tableLog) - distributed;
    }
    if (distributed == maxSym + 1) {
      u32 maxV = 0, maxC = 0;
      for (u32 s = 0; s <= maxSym; s++) if (count[s] > maxC) { maxV = s; maxC = count[s]; }
      norm[maxV] += (s16)ToDistribute; return true;
    }
    if (tot == 0) { for (u32 s = 0; ToDistribute > 0; s = (s + 1) % (maxSym + 1)) if (norm[s] > 0) { ToDistribute--; norm[s]++; } return true; }
    {
      const u64 vStepLog = 62 - tableLog, mid = ((u64)1 << (vStepLog - 1)) - 1;
      const u64 rStep = ((((u64)1 << vStepLog) * ToDistribute) + mid) / tot;
      u64 tmpTotal = mid;
      for (u32 s = 0; s <= maxSym; s++) if (norm[s] == NOT_YET) {
        const u64 end = tmpTotal + (count[s] * rStep);
        const u32 sStart = (u32)(tmpTotal >> vStepLog), sEnd = (u32)(end >> vStepLog), weight = sEnd - sStart;
        if (weight < 1) return false;
        norm[s] = (s16)weight; tmpTotal = end;
      }
    }
  } else norm[largest] += (s16)stillToDistribute;
  return true;
}

// header writer, the inverse of ReadNCount; returns bytes written or 0 when out of room
ZB_HD u32 fse_write_ncount(u8* out, u32 cap, const s16* norm, u32 maxSym, u32 tableLog) {
  u32 op = 0; const i32 tableSize = 1 << tableLog;
  i32 remaining = tableSize + 1, threshold = tableSize, nbBits = (i32)tableLog + 1;
  u32 bitStream = tableLog - 5; i32 bitCount = 4; u32 symbol = 0; const u32 alphabet = maxSym + 1; bool previousIs0 = false;
  while (symbol < alphabet && remaining > 1) {
    if (previousIs0) {
      u32 start = symbol;
      while (symbol < alphabet && !norm[symbol]) symbol++;
      if (symbol == alphabet) break;
      while (symbol >= start + 24) {
        start += 24; bitStream += 0xFFFFu << bitCount;
        if (op + 2 > cap) return 0;
        out[op] = (u8)bitStream; out[op + 1] = (u8)(bitStream >> 8); op += 2; bitStream >>= 16;
      }
      while (symbol >= start + 3) { start += 3; bitStream += 3u << bitCount; bitCount += 2; }
      bitStream += (symbol - start) << bitCount; bitCount += 2;
      if (bitCount > 16) { if (op + 2 > cap) return 0; out[op] = (u8)bitStream; out[op + 1] = (u8)(bitStream >> 8); op += 2; bitStream >>= 16; bitCount -= 16; }
    }
    {
      i32 count = norm[symbol++];
      const i32 max = (2 * threshold - 1) - remaining;
      remaining -= count < 0 ? -count : count;
      count++;
      if (count >= threshold) count += max;
      bitStream += (u32)count << bitCount; bitCount += nbBits; bitCount -= (count < max);
      previousIs0 = (count == 1);
      if (remaining < 1) return 0;
      while (remaining < threshold) { nbBits--; threshold >>= 1; }
    }
    if (bitCount > 16) { if (op + 2 > cap) return 0; out[op] = (u8)bitStream; out[op + 1] = (u8)(bitStream >> 8); op += 2; bitStream >>= 16; bitCount -= 16; }
  }
  if (remaining != 1) return 0;
  if (op + 2 > cap) return 0;
  out[op] = (u8)bitStream; out[op + 1] = (u8)(bitStream >> 8);
  op += (u32)(bitCount + 7) / 8;
  return op;
}

// encoding table: stateTable[tableSize] + per symbol (deltaNbBits, deltaFindState)
struct FseCTable { u16* stateTable; u32 tableLog; u32 deltaNbBits[64]; i32 deltaFindState[64]; };   // alphabets here have <= 53 symbols
ZB_HD void fse_build_ctable(FseCTable& ct, u16* stateTable, const s16* norm, u32 maxSym, u32 tableLog, u8* tableSymbol /* tableSize bytes */) {
  const u32 tableSize = 1u << tableLog, mask = tableSize - 1, step = (tableSize >> 1) + (tableSize >> 3) + 3;
  u32 cumul[258]; u32 high = tableSize - 1;
  ct.stateTable = stateTable; ct.tableLog = tableLog;
  cumul[0] = 0;
  for (u32 u = 1; u <= maxSym + 1; u++) {
    if (norm[u - 1] == -1) { cumul[u] = cumul[u - 1] + 1; tableSymbol[high--] = (u8)(u - 1); }
    else cumul[u] = cumul[u - 1] + (u32)norm[u - 1];
  }
  cumul[maxSym + 2] = tableSize + 1;
  { u32 pos = 0;
    for (u32 s = 0; s <= maxSym; s++)
      for (i32 i = 0; i < norm[s]; i++) { tableSymbol[pos] = (u8)s; pos = (pos + step) & mask; while (pos > high) pos = (pos + step) & mask; } }
  for (u32 u = 0; u < tableSize; u++) { u8 s = tableSymbol[u]; stateTable[cumul[s]++] = (u16)(tableSize + u); }
  { u32 total = 0;
    for (u32 s = 0; s <= maxSym; s++) {
      const i32 c = norm[s];
      if (c == 0) { ct.deltaNbBits[s] = ((tableLog + 1) << 16) - (1u << tableLog); ct.deltaFindState[s] = 0; }
      else if (c == -1 || c == 1) { ct.deltaNbBits[s] = (tableLog << 16) - (1u << tableLog); ct.deltaFindState[s] = (i32)total - 1; total++; }
      else {
        const u32 maxBitsOut = tableLog - highbit((u32)c - 1), minStatePlus = (u32)c << maxBitsOut;
        ct.deltaNbBits[s] = (maxBitsOut << 16) - minStatePlus; ct.deltaFindState[s] = (i32)total - c; total += (u32)c;
      }
    } }
}
ZB_HD void fse_build_ctable_rle(FseCTable& ct, u16* stateTable, u32 sym) {
  ct.stateTable = stateTable; ct.tableLog = 0; stateTable[0] = 0; stateTable[1] = 0;
  ct.deltaNbBits[sym] = 0; ct.deltaFindState[sym] = 0;
}
ZB_HD void fse_init_state(const FseCTable& ct, u32& state, u32 sym) {   // first symbol costs no bits
  const u32 nbBitsOut = (ct.deltaNbBits[sym] + (1u << 15)) >> 16;
  const u32 v = (nbBitsOut << 16) - ct.deltaNbBits[sym];
  state = ct.stateTable[(i32)(v >> nbBitsOut) + ct.deltaFindState[sym]];
}
ZB_HD void fse_encode(BitWriter& w, const FseCTable& ct, u32& state, u32 sym) {
  const u32 nbBitsOut = (state + ct.deltaNbBits[sym]) >> 16;
  bw_add(w, state, nbBitsOut);
  state = ct.stateTable[(i32)(state >> nbBitsOut) + ct.deltaFindState[sym]];
}
ZB_HD void fse_flush_state(BitWriter& w, const FseCTable& ct, u32 state) { bw_add(w, state, ct.tableLog); bw_flush(w); }

// ------------------------------------------------------------------------------------------------
// Huffman: length-limited code from a histogram, weight header, 1/4-stream encoding
// (accepted by ReadStats EntropyCommon.cs:198-269, HUF_readDTableX2 HufDecompress.cs:117-180,
//  4-stream layout HufDecompress.cs:266-307)
// ------------------------------------------------------------------------------------------------
struct HufCode { u16 val; u8 nbBits; };
struct HufEnc { HufCode code[256]; u8 weight[256]; u32 maxSym; u32 tableLog; };

// Builds code lengths <= maxBits for symbols with count > 0 (needs >= 2 distinct symbols).  Package-free
// heuristic: build an optimal tree with two queues over the sorted symbols, then repair over-long codes the
// way zstd's HUF_setMaxHeight does (pay back the Kraft debt on the longest cheap symbols).
// Working arrays of huf_build (4 KB): callers choose where they live (the GPU kernel lends shared memory).
// hash-table logs of the warp-parallel match finder (k_enc_match and its lock-step emulation in tests/hostsim)
ZB_HD u32 enc_hlog_long(int level) { return level == 2 ? 13 : (level >= 3 ? 11 : 12); }
ZB_HD u32 enc_hlog_short(int /*level*/) { return 12; }

struct HufBuildScratch { u32 nodeCount[512]; u16 parent[512]; u16 order[256]; u8 depth[512]; };

// Stable sort of the used symbols by count ascending (ties: symbol ascending) -> order[0..n); returns n.
ZB_HD u32 huf_sort_symbols(u16* order, const u32* count, u32 maxSym) {
  u32 n = 0;
  for (u32 s = 0; s <= maxSym; s++) if (count[s]) order[n++] = (u16)s;
  for (u32 i = 1; i < n; i++) { u16 v = order[i]; u32 c = count[v]; u32 j = i; while (j > 0 && count[order[j - 1]] > c) { order[j] = order[j - 1]; j--; } order[j] = v; }
  return n;
}

// sc.order[0..n) holds the used symbols sorted as huf_sort_symbols does
ZB_HD bool huf_build_sorted(HufEnc& he, const u32* count, u32 n, u32 maxBits, HufBuildScratch& sc) {
  if (n < 2) return false;
  u16* const order = sc.order; u32* const nodeCount = sc.nodeCount; u16* const parent = sc.parent; u8* const depth = sc.depth;
  // two-queue Huffman: nodes 0..n-1 leaves (sorted), n..2n-2 internal
  for (u32 i = 0; i < n; i++) nodeCount[i] = count[order[i]];
  u32 leaf = 0, inner = n, next = n;
  while (next < 2 * n - 1) {
    u32 pick[2];
    for (int k = 0; k < 2; k++) {
      if (leaf < n && (inner >= next || nodeCount[leaf] <= nodeCount[inner])) pick[k] = leaf++; else pick[k] = inner++;
    }
    nodeCount[next] = nodeCount[pick[0]] + nodeCount[pick[1]];
    parent[pick[0]] = parent[pick[1]] = (u16)next; next++;
  }
  depth[2 * n - 2] = 0;
  for (i32 i = (i32)(2 * n - 3); i >= 0; i--) depth[i] = depth[parent[i]] + 1;
  // enforce the length limit: clamp, then repay the Kraft excess (in units of 2^-maxBits) by lengthening the
  // codes that are closest to the limit, and hand any overshoot back by shortening codes at the limit
  u32 largest = 0; for (u32 i = 0; i < n; i++) if (depth[i] > largest) largest = depth[i];
  if (largest > maxBits) {
    i64 debt = -((i64)1 << maxBits);
    for (u32 i = 0; i < n; i++) { if (depth[i] > maxBits) depth[i] = (u8)maxBits; debt += (i64)1 << (maxBits - depth[i]); }
    while (debt > 0) {
      i32 best = -1; i64 bestGain = 0;
      for (u32 i = 0; i < n; i++) {            // ascending count: the cheapest symbol of a depth comes first
        if (depth[i] >= maxBits) continue;
        const i64 gain = (i64)1 << (maxBits - depth[i] - 1);
        if (gain <= debt && gain > bestGain) { best = (i32)i; bestGain = gain; }
      }
      if (best < 0) {                          // every step overshoots: take the smallest one
        for (u32 i = 0; i < n; i++) if (depth[i] < maxBits) { const i64 gain = (i64)1 << (maxBits - depth[i] - 1); if (best < 0 || gain < bestGain) { best = (i32)i; bestGain = gain; } }
        if (best < 0) return false;
      }
      depth[best]++; debt -= bestGain;
    }
    while (debt < 0) {                         // refund one unit at a time: the most frequent code at the limit gets shorter
      i32 pick = -1;
      for (i32 i = (i32)n - 1; i >= 0; i--) if (depth[i] == maxBits) { pick = i; break; }
      if (pick < 0) return false;
      depth[pick]--; debt++;
    }
  }
  // the decoder derives the last weight, so the Kraft sum must be exactly 1 and tableLog = max depth
  u32 tl = 0; for (u32 i = 0; i < n; i++) if (depth[i] > tl) tl = depth[i];
  { u64 kraft = 0; for (u32 i = 0; i < n; i++) kraft += 1ull << (tl - depth[i]); if (kraft != (1ull << tl)) return false; }
  for (u32 s = 0; s < 256; s++) { he.code[s].nbBits = 0; he.code[s].val = 0; he.weight[s] = 0; }
  for (u32 i = 0; i < n; i++) { he.code[order[i]].nbBits = depth[i]; he.weight[order[i]] = (u8)(tl + 1 - depth[i]); }
  he.tableLog = tl; he.maxSym = order[0];
  for (u32 i = 0; i < n; i++) if (order[i] > he.maxSym) he.maxSym = order[i];
  // canonical values matching the decoder's table fill: within a weight, symbols ascending; weights ascending
  // start at rankStart (HufDecompress.cs:151-177): code value = cell index >> (tableLog - nbBits)
  u32 rankCount[16]; for (u32 i = 0; i < 16; i++) rankCount[i] = 0;
  for (u32 s = 0; s <= he.maxSym; s++) rankCount[he.weight[s]]++;
  u32 rankStart[16]; u32 nextStart = 0;
  for (u32 wv = 1; wv <= tl; wv++) { rankStart[wv] = nextStart; nextStart += rankCount[wv] << (wv - 1); }
  for (u32 s = 0; s <= he.maxSym; s++) {
    const u32 wv = he.weight[s]; if (!wv) continue;
    he.code[s].val = (u16)(rankStart[wv] >> (wv - 1));
    rankStart[wv] += 1u << (wv - 1);
  }
  return true;
}
ZB_HD bool huf_build(HufEnc& he, const u32* count, u32 maxSym, u32 maxBits, HufBuildScratch& sc) {
  const u32 n = huf_sort_symbols(sc.order, count, maxSym);
  return huf_build_sorted(he, count, n, maxBits, sc);
}

// weight header: FSE-compressed weights when that is smaller, else 4-bit nibbles (needs maxSym <= 128 there)
ZB_HD u32 huf_write_header(u8* out, u32 cap, const HufEnc& he, u16* stateScratch, u8* symScratch) {
  const u32 nw = he.maxSym;   // weights of symbols 0..maxSym-1; the last one is implied
  if (nw == 0) return 0;
  // try FSE
  u32 fseSize = 0; alignas(4) u8 fseBuf[160];
  if (nw > 1) {
    u32 cnt[13]; for (u32 i = 0; i < 13; i++) cnt[i] = 0;
    u32 maxW = 0; for (u32 s = 0; s < nw; s++) { cnt[he.weight[s]]++; if (he.weight[s] > maxW) maxW = he.weight[s]; }
    bool single = false; for (u32 i = 0; i <= maxW; i++) if (cnt[i] == nw) single = true;
    if (!single && nw > 1) {
      u32 tl = fse_optimal_log(6, nw, maxW);
      if (tl > 6) tl = 6;
      s16 norm[13];
      if (fse_normalize(norm, tl, cnt, nw, maxW)) {
        u32 h = fse_write_ncount(fseBuf, 150, norm, maxW, tl);
        if (h) {
          FseCTable ct; fse_build_ctable(ct, stateScratch, norm, maxW, tl, symScratch);
          BitWriter w; bw_init(w, fseBuf + h, fseBuf + 150);
          // two interleaved states, last symbol first (decoder: FseDecompress.cs:233-295)
          i32 ip = (i32)nw; u32 s1, s2;
          if (nw & 1) { fse_init_state(ct, s1, he.weight[--ip]); fse_init_state(ct, s2, he.weight[--ip]); fse_encode(w, ct, s1, he.weight[--ip]); bw_flush(w); }
          else { fse_init_state(ct, s2, he.weight[--ip]); fse_init_state(ct, s1, he.weight[--ip]); }
          while (ip > 0) {
            fse_encode(w, ct, s2, he.weight[--ip]);
            if (ip > 0) fse_encode(w, ct, s1, he.weight[--ip]);
            bw_flush(w);
          }
          fse_flush_state(w, ct, s2); fse_flush_state(w, ct, s1);
          u8* e = bw_close(w);
          if (e) fseSize = (u32)(e - fseBuf);
        }
      }
    }
  }
  if (fseSize > 1 && fseSize < nw / 2 && fseSize < 128) {
    if (1 + fseSize > cap) return 0;
    out[0] = (u8)fseSize; for (u32 i = 0; i < fseSize; i++) out[1 + i] = fseBuf[i];
    return 1 + fseSize;
  }
  if (nw > 128) return 0;
  const u32 bytes = (nw + 1) / 2;
  if (1 + bytes > cap) return 0;
  out[0] = (u8)(128 + (nw - 1));
  for (u32 n = 0; n < nw; n += 2) out[1 + n / 2] = (u8)((he.weight[n] << 4) | (n + 1 < nw ? he.weight[n + 1] : 0));
  return 1 + bytes;
}

// one stream: symbols are written last-to-first so that the backward reader yields them in order
ZB_HD u32 huf_encode_stream(u8* out, u32 cap, const u8* src, u32 n, const HufEnc& he) {
  BitWriter w; bw_init(w, out, out + cap);
  for (i32 i = (i32)n - 1; i >= 0; i--) { const HufCode c = he.code[src[i]]; bw_add(w, c.val, c.nbBits); bw_flush(w); }
  u8* e = bw_close(w);
  return e ? (u32)(e - out) : 0;
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

// ------------------------------------------------------------------------------------------------
// sequences section (DecodeSeqHeaders :1110-1180, DecodeSequence :1473-1553)
// ------------------------------------------------------------------------------------------------
ZB_HD u32 ll_code(u32 ll) {
  if (ll < 16) return ll;
  if (ll < 64) { const u8 t[48] = {16,16,17,17,18,18,19,19,20,20,20,20,21,21,21,21,22,22,22,22,22,22,22,22,23,23,23,23,23,23,23,23,
                                   24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24}; return t[ll - 16]; }
  return highbit(ll) + 19;
}
ZB_HD u32 ml_code(u32 mlm3) {
  if (mlm3 < 32) return mlm3;
  if (mlm3 < 128) { const u8 t[96] = {32,32,33,33,34,34,35,35,36,36,36,36,37,37,37,37,38,38,38,38,38,38,38,38,39,39,39,39,39,39,39,39,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,41,41,41,41,41,41,41,41,41,41,41,41,41,41,41,41,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42};
                     return t[mlm3 - 32]; }
  return highbit(mlm3) + 36;
}

struct SeqKind { u32 maxSym, maxLog, defLog; const s16* defNorm; };

// chooses the table mode for one symbol kind and emits its description; returns the mode or 0xFF on failure
// count[] = histogram of the nbSeq codes of one kind (it is modified), lastCode = code of the last sequence
ZB_HD u32 enc_seq_table_counts(FseCTable& ct, u16* stateTable, u8* symScratch, u32* count, u32 lastCode, u32 nbSeq, const SeqKind& k, int level,
                               u8* out, u32 cap, u32* used) {
  u32 maxSym = k.maxSym; while (maxSym > 0 && !count[maxSym]) maxSym--;
  u32 most = 0; for (u32 s = 0; s <= maxSym; s++) if (count[s] > most) most = count[s];
  *used = 0;
  s16 norm[53];
  const bool defOk = maxSym <= (k.defLog == 5 ? 28u : k.maxSym);
  if (most == nbSeq) {
    if (defOk && nbSeq <= 2) { for (u32 s = 0; s <= k.maxSym; s++) norm[s] = s < (k.defLog == 5 ? 29u : k.maxSym + 1) ? k.defNorm[s] : 0;
      fse_build_ctable(ct, stateTable, norm, k.defLog == 5 ? 28 : k.maxSym, k.defLog, symScratch); return 0; }
    if (cap < 1) return 0xFF;
    out[0] = (u8)maxSym; *used = 1; fse_build_ctable_rle(ct, stateTable, maxSym); return 1;
  }
  if (defOk) {
    const u32 mult = 10 - (level <= 1 ? 1 : (level == 2 ? 1 : 2)), dynMin = ((1u << k.defLog) * mult) >> 3;
    if (nbSeq < dynMin || most < (nbSeq >> (k.defLog - 1))) {
      for (u32 s = 0; s <= k.maxSym; s++) norm[s] = s < (k.defLog == 5 ? 29u : k.maxSym + 1) ? k.defNorm[s] : 0;
      fse_build_ctable(ct, stateTable, norm, k.defLog == 5 ? 28 : k.maxSym, k.defLog, symScratch); return 0;
    }
  }
  u32 tl = fse_optimal_log(k.maxLog, nbSeq, maxSym);
  u32 total = nbSeq;
  if (count[lastCode] > 1) { count[lastCode]--; total--; }   // the last symbol costs no bits
  if (!fse_normalize(norm, tl, count, total, maxSym)) return 0xFF;
  u32 h = fse_write_ncount(out, cap, norm, maxSym, tl);
  if (!h) return 0xFF;
  *used = h;
  fse_build_ctable(ct, stateTable, norm, maxSym, tl, symScratch);
  return 2;
}
ZB_HD u32 enc_seq_table(FseCTable& ct, u16* stateTable, u8* symScratch, const u8* codes, u32 nbSeq, const SeqKind& k, int level,
                        u8* out, u32 cap, u32* used) {
  u32 count[53]; for (u32 i = 0; i <= k.maxSym; i++) count[i] = 0;
  for (u32 i = 0; i < nbSeq; i++) count[codes[i]]++;
  return enc_seq_table_counts(ct, stateTable, symScratch, count, codes[nbSeq - 1], nbSeq, k, level, out, cap, used);
}

ZB_HD u32 enc_seq_count_header(u8* out, u32 nbSeq) {
  u32 op = 0;
  if (nbSeq < 128) out[op++] = (u8)nbSeq;
  else if (nbSeq < LONGNBSEQ) { out[op++] = (u8)((nbSeq >> 8) + 0x80); out[op++] = (u8)nbSeq; }
  else { out[op++] = 0xFF; out[op++] = (u8)(nbSeq - LONGNBSEQ); out[op++] = (u8)((nbSeq - LONGNBSEQ) >> 8); }
  return op;
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

ZB_HD SeqKind seq_kind(int kind) {
  if (kind == KIND_LL) return SeqKind{MaxLL, LLFSELog, 6, kLLnorm};
  if (kind == KIND_OF) return SeqKind{MaxOff, OffFSELog, 5, kOFnorm};
  return SeqKind{MaxML, MLFSELog, 6, kMLnorm};
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
