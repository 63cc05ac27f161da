// encode_kernels.cu — sm_100a kernels of the batched zstd frame encoder.
//
//   k_enc_match<DFAST> : warp / frame    match finder: 32 consecutive (or strided) positions per step, one per lane;
//                                        position hash table(s) in shared memory (u16 entries), in-window duplicate
//                                        hashes resolved with __match_any_sync, candidates verified in parallel, the
//                                        first hit wins (greedy parse), match length by warp ballot; sequence store
//                                        and literals go to HBM scratch
//   k_enc_entropy      : thread / frame  Huffman + FSE table construction, bitstream encode, block + frame assembly
//                                        (zb_encode.cuh) from the stored sequences
//   k_enc_xxh          : 4 lanes / frame XXH64 content checksum appended to the frame (xxh_device.cuh)
//
// Frames are independent, so there is no inter-warp communication.  HBM scratch of the match stage is addressed
// from src_off (no scan): literals of item f at lit + src_off[f] + 64 f, sequences (8-byte entries) at
// seq + src_off[f]/4 + 64 (src_off[f] >> 17) + 128 f, block metadata at meta + (src_off[f] >> 17) + f.
#include "encode_kernels.cuh"
#include "zb_encode.cuh"
#include "xxh_device.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace zb {

#define FULLMASK 0xFFFFFFFFu
static constexpr u32 kBlockSeqCap = BLOCKSIZE_MAX / 4 + 64;

struct BlockMeta { u32 nseq, nlits; };

__device__ __forceinline__ u8* lit_base(const EncodeArgs& a, const EncodeScratch& s, u32 f) { return s.lit + a.src_off[f] + 64ull * (a.item_base + f); }
__device__ __forceinline__ u32* seq_base(const EncodeArgs& a, const EncodeScratch& s, u32 f) {
  const u64 o = a.src_off[f];
  return s.seq + 2 * (o / 4 + 64 * (o >> 17) + 128ull * (a.item_base + f));
}
__device__ __forceinline__ BlockMeta* meta_base(const EncodeArgs& a, const EncodeScratch& s, u32 f) {
  return (BlockMeta*)s.meta + (a.src_off[f] >> 17) + (a.item_base + f);
}

// unaligned loads built from aligned words (may touch up to 3 bytes past the last byte asked for)
__device__ __forceinline__ u64 ldu64(const u8* p) {
  const u32* w = (const u32*)((uintptr_t)p & ~(uintptr_t)3); const u32 sh = ((u32)(uintptr_t)p & 3) * 8;
  const u32 a0 = w[0], a1 = w[1], a2 = sh ? w[2] : 0;
  return ((u64)__funnelshift_r(a1, a2, sh) << 32) | __funnelshift_r(a0, a1, sh);
}
__device__ __forceinline__ u32 ldu32(const u8* p) {
  const u32* w = (const u32*)((uintptr_t)p & ~(uintptr_t)3); const u32 sh = ((u32)(uintptr_t)p & 3) * 8;
  const u32 a0 = w[0], a1 = sh ? w[1] : 0;
  return __funnelshift_r(a0, a1, sh);
}
__device__ __forceinline__ u32 hash64(u64 v, u32 hlog, u32 mls) {
  if (mls >= 8) return (u32)((v * 0xCF1BBCDCB7A56463ull) >> (64 - hlog));
  if (mls == 6) return (u32)(((v << 16) * 0xCF1BBCDCBF9Bull) >> (64 - hlog));
  return (u32)(((v << 24) * 0xCF1BBCDCBBull) >> (64 - hlog));   // 5 bytes
}

// candidate position from a 16-bit table entry: the most recent position below p with those low 16 bits
__device__ __forceinline__ bool cand_from_entry(u32 p, u32 e, u32& cand) {
  u32 d = (p - e) & 0xFFFF; if (d == 0) d = 0x10000;
  cand = p - d;
  return d <= p;
}

// number of equal bytes of src[a..) and src[a-off..), a + n <= end, counted by the whole warp
__device__ __forceinline__ u32 warp_extend(const u8* src, u32 a, u32 off, u32 end, u32 lane) {
  u32 n = 0;
  while (true) {
    const u32 i = a + n + lane;
    const bool eq = i < end && src[i] == src[i - off];
    const unsigned m = __ballot_sync(FULLMASK, eq);
    if (m == FULLMASK) { n += 32; continue; }
    return n + (u32)__ffs(~m) - 1;
  }
}

template <bool DFAST>
__global__ void __launch_bounds__(32) k_enc_match(EncodeArgs a, EncodeScratch sc, u32 hlogL, u32 hlogS, u32 mls) {
  extern __shared__ __align__(16) u16 tab[];
  u16* const tabL = tab; u16* const tabS = tab + (1u << hlogL);
  const u32 lane = threadIdx.x;
  const u32 tw = (1u << hlogL) + (DFAST ? (1u << hlogS) : 0);
  for (u32 f = blockIdx.x; f < a.n; f += gridDim.x) {
    const u8* src = a.src_base + a.src_off[f]; const u32 size = a.src_size[f];
    u8* const lits0 = lit_base(a, sc, f); u32* const seqs0 = seq_base(a, sc, f); BlockMeta* const meta = meta_base(a, sc, f);
    for (u32 i = lane; i < tw / 2; i += 32) ((u32*)tab)[i] = 0;
    __syncwarp();
    u32 rep1 = 1, rep2 = 4;
    u32 blk = 0, bpos = 0;
    do {
      const u32 bsize = size - bpos < BLOCKSIZE_MAX ? size - bpos : BLOCKSIZE_MAX;
      const u32 bend = bpos + bsize;
      u8* const lits = lits0 + (size_t)blk * BLOCKSIZE_MAX; u32* const seqs = seqs0 + 2 * (size_t)blk * kBlockSeqCap;
      u32 nseq = 0, nlits = 0;
      // all-equal block? (RLE block: the entropy stage handles it, nothing to match)
      bool rle = bsize >= 2;
      if (rle) {
        const u8 b0 = src[bpos];
        for (u32 i0 = bpos; i0 < bend && rle; i0 += 32) { const u32 i = i0 + lane; if (__ballot_sync(FULLMASK, i < bend && src[i] != b0)) rle = false; }
      }
      if (!rle && bsize >= 64) {
        // repeat offsets are only trusted when established inside this block (or the frame's initial 1, 4): an
        // earlier block may end up stored raw, which leaves the decoder's history untouched (DecodeSequence :1576-1596)
        if (blk > 0) { rep1 = 0; rep2 = 0; }
        u32 anchor = bpos; const u32 ilimit = bend - 8;
        u32 p0 = bpos + (bpos == 0 ? 1 : 0);
        while (p0 < ilimit) {
          const u32 step = 1 + ((p0 - anchor) >> 8);                 // skim incompressible runs
          const u32 p = p0 + lane * step;
          const bool act = p < ilimit;
          const u64 v = act ? ldu64(src + p) : 0;
          const u32 hL = act ? hash64(v, hlogL, DFAST ? 8 : mls) : (0x80000000u | lane);
          const u32 eL = act ? tabL[hL] : 0;
          const unsigned mL = __match_any_sync(FULLMASK, hL);
          const unsigned lowL = mL & ((1u << lane) - 1);
          u32 candL; bool validL;
          if (lowL) { candL = p0 + (31 - __clz(lowL)) * step; validL = act; } else validL = cand_from_entry(p, eL, candL) && act;
          u32 hS = 0, candS = 0; bool validS = false; unsigned mS = 0;
          if (DFAST) {
            hS = act ? hash64(v, hlogS, mls) : (0x80000000u | lane);
            const u32 eS = act ? tabS[hS] : 0;
            mS = __match_any_sync(FULLMASK, hS);
            const unsigned lowS = mS & ((1u << lane) - 1);
            if (lowS) { candS = p0 + (31 - __clz(lowS)) * step; validS = act; } else validS = cand_from_entry(p, eS, candS) && act;
          }
          bool okL = false, okS = false;
          if (validL) okL = DFAST ? (ldu64(src + candL) == v) : (ldu32(src + candL) == (u32)v);
          if (DFAST && validS) okS = ldu32(src + candS) == (u32)v;
          const bool rok = act && rep1 != 0 && rep1 <= p && ldu32(src + p - rep1) == (u32)v;
          const unsigned fRep = __ballot_sync(FULLMASK, rok);
          const unsigned found = __ballot_sync(FULLMASK, okL | okS) | fRep;
          if (!found) {                                                // no match in the window: enter every position, move on
            if (act && (mL >> lane) == 1) tabL[hL] = (u16)p;          // the last position of each hash group is the one kept
            if (DFAST && act && (mS >> lane) == 1) tabS[hS] = (u16)p;
            __syncwarp();
            p0 += 32 * step; continue;
          }
          u32 fl = (u32)__ffs(found) - 1;
          // a repeat-offset match up to 3 positions further on beats a table match here (it is cheaper to code)
          if (!((fRep >> fl) & 1)) { const unsigned nearRep = fRep & (0xEu << fl); if (nearRep) fl = (u32)__ffs(nearRep) - 1; }
          u32 pos = p0 + fl * step;
          const u32 kind = rok ? 0 : (okL ? 1 : 2);                   // at the winning lane: repeat offset > long/main table > short table
          const u32 kf = __shfl_sync(FULLMASK, kind, fl);
          const u32 cf = __shfl_sync(FULLMASK, kind == 1 ? candL : candS, fl);
          u32 cnd = kf == 0 ? pos - rep1 : cf;
          const u32 off = pos - cnd;
          u32 mlen = 4 + warp_extend(src, pos + 4, off, bend, lane);
          while (pos > anchor && cnd > 0 && src[pos - 1] == src[cnd - 1]) { pos--; cnd--; mlen++; }   // catch up
 {
            // enter the window's positions up to the end of the match: nothing at or beyond the restart position may
            // go in, or it would later be found as its own candidate
            const bool ins = act && p < pos + mlen;
            const unsigned insMask = __ballot_sync(FULLMASK, ins);
            if (ins && ((mL & insMask) >> lane) == 1) tabL[hL] = (u16)p;
            if (DFAST && ins && ((mS & insMask) >> lane) == 1) tabS[hS] = (u16)p;
            __syncwarp();
          }
          const u32 ll = pos - anchor;
          u32 offBase;
          if (kf == 0 && ll > 0) offBase = 1;                          // repeat offset 1; history unchanged
          else { offBase = off + 3; rep2 = rep1; rep1 = off; }         // (a repeat right after a match must be spelled out: with
                                                                       //  litLength 0 the code 1 means the *second* offset, :1511)
          for (u32 i = lane; i < ll; i += 32) lits[nlits + i] = src[anchor + i];
          if (lane == 0) {
            seqs[2 * nseq] = (ll & 0xFFFF) | (((mlen - 3) & 0xFFFF) << 16);
            seqs[2 * nseq + 1] = (offBase & 0x3FFFFFFFu) | ((((mlen - 3) >> 16) & 1) << 30) | (((ll >> 16) & 1) << 31);
          }
          nlits += ll; nseq++;
          anchor = pos + mlen; p0 = anchor;
          if (p0 <= ilimit) {
            if (lane == 0) { const u32 q = p0 - 2; tabL[hash64(ldu64(src + q), hlogL, DFAST ? 8 : mls)] = (u16)q; if (DFAST) tabS[hash64(ldu64(src + q), hlogS, mls)] = (u16)q; }
            __syncwarp();
            // immediate repeat of the older offset (litLength 0, code 1 == second offset)
            while (rep2 != 0 && p0 <= ilimit && ldu32(src + p0) == ldu32(src + p0 - rep2)) {
              const u32 rlen = 4 + warp_extend(src, p0 + 4, rep2, bend, lane);
              { const u32 t = rep2; rep2 = rep1; rep1 = t; }
              if (lane == 0) {
                tabL[hash64(ldu64(src + p0), hlogL, DFAST ? 8 : mls)] = (u16)p0;
                if (DFAST) tabS[hash64(ldu64(src + p0), hlogS, mls)] = (u16)p0;
                seqs[2 * nseq] = ((rlen - 3) & 0xFFFF) << 16;
                seqs[2 * nseq + 1] = 1u | ((((rlen - 3) >> 16) & 1) << 30);
              }
              __syncwarp();
              nseq++;
              p0 += rlen; anchor = p0;
            }
          }
        }
        // last literals
        const u32 ll = bend - anchor;
        for (u32 i = lane; i < ll; i += 32) lits[nlits + i] = src[anchor + i];
        nlits += ll;
      }
      if (lane == 0) { meta[blk].nseq = nseq; meta[blk].nlits = nlits; }
      bpos = bend; blk++;
    } while (bpos < size);
  }
}

// sequences left by k_enc_match
struct StoredMatcher {
  u8* lits0; u32* seqs0; const BlockMeta* meta;
  __host__ __device__ bool operator()(SeqStore& st, u32 blk, const u8*, u32 bsize, bool isRle) {
    if (isRle || bsize < 64) return false;
    st.seqs = seqs0 + 2 * (size_t)blk * kBlockSeqCap; st.n = meta[blk].nseq; st.cap = kBlockSeqCap;
    st.lits = lits0 + (size_t)blk * BLOCKSIZE_MAX; st.nlits = meta[blk].nlits;
    return true;
  }
  __host__ __device__ void done(bool) {}
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
__host__ __device__ inline size_t slot_bytes() { return align16((size_t)kBlockSeqCap * 3) + align16(3 * 514 * 2 + 32) + 1024; }

__global__ void __launch_bounds__(64) k_enc_entropy(EncodeArgs a, EncodeScratch sc, u32 slot0, u32 slots) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= slots) return;
  u8* p = sc.slots + (size_t)(slot0 + t) * slot_bytes();
  u8* codes = p; p += align16((size_t)kBlockSeqCap * 3);
  u16* ctables = (u16*)p; p += align16(3 * 514 * 2 + 32);
  u8* sym = p;
  for (u32 f = t; f < a.n; f += slots) {
    StoredMatcher m{lit_base(a, sc, f), seq_base(a, sc, f), meta_base(a, sc, f)};
    a.result[f] = encode_frame_with(a.src_base + a.src_off[f], a.src_size[f], a.dst_base + a.dst_off[f], a.dst_cap[f], a.level, a.checksum,
                                    codes, ctables, sym, m);
  }
}

__global__ void __launch_bounds__(128) k_enc_xxh(EncodeArgs a) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 f = t >> 2, sub = t & 3, lane = threadIdx.x & 31;
  const unsigned gmask = 0xFu << (lane & ~3u);
  if (f >= a.n) return;
  const u32 r = a.result[f];
  if (is_err(r)) return;
  const u64 h = xxh64_group(a.src_base + a.src_off[f], a.src_size[f], sub, gmask, lane & ~3u);
  if (sub == 0) {
    u8* p = a.dst_base + a.dst_off[f] + r;   // encode_frame_with reserved these 4 bytes
    p[0] = (u8)h; p[1] = (u8)(h >> 8); p[2] = (u8)(h >> 16); p[3] = (u8)(h >> 24);
    a.result[f] = r + 4;
  }
}

size_t encode_bound(size_t srcSize) { return srcSize + (srcSize >> 8) + 32 + 3 * ((srcSize >> 17) + 1); }

static constexpr u32 kSlotsPerStream = 148 * 32;
static constexpr u32 kStreamSlots = 4;

cudaError_t encode_alloc(EncodeScratch& s, size_t maxBatchBytes, size_t maxItems) {
  s = EncodeScratch();
  s.maxBytes = maxBatchBytes + 16 * maxItems; s.maxItems = maxItems;   // src offsets are 16-byte aligned in the host path
  return cudaSuccess;
}
void encode_free(EncodeScratch& s) {
  if (s.lit) cudaFree(s.lit); if (s.seq) cudaFree(s.seq); if (s.meta) cudaFree(s.meta); if (s.slots) cudaFree(s.slots);
  s.lit = nullptr; s.seq = nullptr; s.meta = nullptr; s.slots = nullptr;
}

static cudaError_t encode_lazy_alloc(EncodeScratch& s) {
  if (s.lit) return cudaSuccess;
  const size_t B = s.maxBytes, N = s.maxItems + 2;
  cudaError_t e;
  if ((e = cudaMalloc(&s.lit, B + 64 * N + 256)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&s.seq, (B / 4 + 64 * ((B >> 17) + 1) + 128 * N + 64) * 8)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&s.meta, ((B >> 17) + N + 8) * sizeof(BlockMeta))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&s.slots, (size_t)kSlotsPerStream * kStreamSlots * slot_bytes())) != cudaSuccess) return e;
  int dev = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&s.sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(k_enc_match<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_enc_match<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  return cudaSuccess;
}

cudaError_t encode_launch(const EncodeArgs& a, EncodeScratch& s, cudaStream_t st, int* launches) {
  if (a.n == 0) return cudaSuccess;
  cudaError_t e = encode_lazy_alloc(s);
  if (e != cudaSuccess) return e;
  // match stage: table sizes per level (u16 entries); as many resident warps as shared memory allows
  const bool dfast = a.level >= 3;
  const u32 hlogL = a.level <= 1 ? 13 : 15, hlogS = 14, mls = a.level <= 1 ? 6 : 5;
  const size_t smem = ((size_t)(1u << hlogL) + (dfast ? (1u << hlogS) : 0)) * 2;
  u32 perSm = (u32)((220 * 1024) / smem); if (perSm > 16) perSm = 16;
  u32 grid = (u32)s.sms * perSm; if (grid > a.n) grid = a.n;
  static const bool timing = getenv("ZSTDB200_ENC_TIMING") != nullptr;   // debug aid: per-kernel times on stderr
  cudaEvent_t ev[3];
  if (timing) { for (auto& x : ev) cudaEventCreate(&x); cudaEventRecord(ev[0], st); }
  if (dfast) k_enc_match<true><<<grid, 32, smem, st>>>(a, s, hlogL, hlogS, mls);
  else k_enc_match<false><<<grid, 32, smem, st>>>(a, s, hlogL, hlogS, mls);
  const u32 part = a.stream_slot % kStreamSlots;
  const u32 slots = a.n < kSlotsPerStream ? a.n : kSlotsPerStream;
  if (timing) cudaEventRecord(ev[1], st);
  k_enc_entropy<<<(slots + 63) / 64, 64, 0, st>>>(a, s, part * kSlotsPerStream, slots);
  if (timing) {
    cudaEventRecord(ev[2], st); cudaEventSynchronize(ev[2]);
    float m1 = 0, m2 = 0; cudaEventElapsedTime(&m1, ev[0], ev[1]); cudaEventElapsedTime(&m2, ev[1], ev[2]);
    fprintf(stderr, "[zstdb200] encode level %d n %u: k_enc_match %.3f ms (grid %u, smem %zu), k_enc_entropy %.3f ms (slots %u)\n", a.level, a.n, m1, grid, smem, m2, slots);
    for (auto& x : ev) cudaEventDestroy(x);
  }
  if (launches) *launches += 2;
  if (a.checksum) { k_enc_xxh<<<(a.n * 4 + 127) / 128, 128, 0, st>>>(a); if (launches) *launches += 1; }
  return cudaGetLastError();
}

}  // namespace zb
