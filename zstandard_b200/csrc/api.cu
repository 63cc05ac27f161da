// api.cu — host side of libzstdb200: contexts, device arenas, staging, sharding, the C ABI of include/zstdb200.h.
//
// Host responsibilities: build flat descriptor tables, shard items over the context's GPUs by
// bytes (no collective: frames are independent, ZStdDecompress.cs:2478-2499 resets all state per frame), move
// bytes host<->device and launch the kernels.  No decoding or encoding arithmetic happens on the host.
//
// Data movement.  A batch is cut into sub-batches that fit the device arenas; a sub-batch is cut into slices that
// are dealt round-robin to NSTREAMS streams, so the H2D copy of one slice, the kernels of another and the D2H
// copy of a third overlap.  When the caller's buffers are already laid out back to back (the layout a managed host gets
// from one pinned array plus offsets) they are DMA'd directly; otherwise items are gathered/scattered through the
// context's pinned staging.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zstdb200.h"
#include "decode_kernels.cuh"
#include "encode_kernels.cuh"

using namespace zb;

namespace {

constexpr int NSTREAMS = 8;
static_assert(NSTREAMS <= ENC_STREAM_PARTS, "one encoder slot partition per stream");
constexpr size_t SLICE_BYTES_DEFAULT = 32u << 20;   // uncompressed bytes per slice (several slices per stream per GiB)
constexpr size_t MIN_SLICE_ITEMS_DECODE = 1024;      // and at least this many frames (decode kernels: one frame-time per launch)
constexpr size_t MIN_SLICE_ITEMS_ENCODE = 512;       // the match kernel fills the GPU with 2-4 thousand frames: start early

// tuning aids (not part of the ABI): ZSTDB200_STREAMS = streams actually used (1..NSTREAMS), ZSTDB200_SLICE_MB
int env_int(const char* name, int dflt, int lo, int hi) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}

struct Device {
  int id = 0;
  cudaStream_t stream[NSTREAMS] = {};
  cudaEvent_t descEv = nullptr;                 // descriptors of the current sub-batch are on the device
  // device memory
  u8 *d_src = nullptr, *d_dst = nullptr;       // staging for the host-pointer API
  u64 *d_srcOff = nullptr, *d_dstOff = nullptr; u32 *d_srcSize = nullptr, *d_dstCap = nullptr, *d_result = nullptr;
  FrameInfo* d_info = nullptr; u8* d_lit = nullptr; SeqRec* d_seq = nullptr;
  u32* d_more = nullptr;                        // per-stream counters of items with another data frame to decode (DecodeArgs::more)
  EncodeScratch enc;                            // encoder arenas (encode_kernels.cuh)
  // pinned host memory
  u8 *h_src = nullptr, *h_dst = nullptr;
  u64 *h_srcOff = nullptr, *h_dstOff = nullptr; u32 *h_srcSize = nullptr, *h_dstCap = nullptr, *h_result = nullptr;
  u32* h_more = nullptr;
};

}  // namespace

struct zstdb200_ctx {
  std::vector<Device> dev;
  size_t maxBatch = 0, maxItems = 0, srcCap = 0, dstSpan = 0;
  std::string err;
  uint64_t launches = 0;
};

namespace {

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512]; snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->err = b_; return 1;                                                                     \
    }                                                                                              \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int alloc_device(zstdb200_ctx* ctx, Device& d) {
  CK(cudaSetDevice(d.id));
  for (auto& s : d.stream) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&d.descEv, cudaEventDisableTiming));
  const size_t items = ctx->maxItems;
  // d_src / d_dst are the codec's input / output staging in both directions: compressed frames of a full batch can
  // be larger than the batch (compress bounds), so both buffers take the larger size
  CK(cudaMalloc(&d.d_src, ctx->srcCap + 256));
  CK(cudaMalloc(&d.d_dst, ctx->srcCap + 256));
  CK(cudaMalloc(&d.d_srcOff, items * 8)); CK(cudaMalloc(&d.d_dstOff, items * 8));
  CK(cudaMalloc(&d.d_srcSize, items * 4)); CK(cudaMalloc(&d.d_dstCap, items * 4)); CK(cudaMalloc(&d.d_result, items * 4));
  CK(cudaMalloc(&d.d_info, items * sizeof(FrameInfo)));
  CK(cudaMalloc(&d.d_more, NSTREAMS * 4)); CK(cudaMallocHost(&d.h_more, NSTREAMS * 4));
  CK(cudaMalloc(&d.d_lit, decode_lit_arena_bytes(ctx->dstSpan, items)));
  CK(cudaMalloc(&d.d_seq, decode_seq_arena_bytes(ctx->dstSpan, items)));
  CK(cudaMallocHost(&d.h_src, ctx->srcCap + 256)); CK(cudaMallocHost(&d.h_dst, ctx->srcCap + 256));
  CK(cudaMallocHost(&d.h_srcOff, items * 8)); CK(cudaMallocHost(&d.h_dstOff, items * 8));
  CK(cudaMallocHost(&d.h_srcSize, items * 4)); CK(cudaMallocHost(&d.h_dstCap, items * 4)); CK(cudaMallocHost(&d.h_result, items * 4));
  CK(decode_configure());
  cudaError_t ee = encode_alloc(d.enc, ctx->maxBatch, items);
  if (ee != cudaSuccess) { ctx->err = std::string("encoder arena allocation failed: ") + cudaGetErrorString(ee); return 1; }
  return 0;
}

void free_device(Device& d) {
  cudaSetDevice(d.id);
  for (auto& s : d.stream) if (s) cudaStreamSynchronize(s);
  cudaFree(d.d_src); cudaFree(d.d_dst); cudaFree(d.d_srcOff); cudaFree(d.d_dstOff); cudaFree(d.d_srcSize); cudaFree(d.d_dstCap);
  cudaFree(d.d_result); cudaFree(d.d_info); cudaFree(d.d_lit); cudaFree(d.d_seq); cudaFree(d.d_more); cudaFreeHost(d.h_more);
  encode_free(d.enc);
  cudaFreeHost(d.h_src); cudaFreeHost(d.h_dst); cudaFreeHost(d.h_srcOff); cudaFreeHost(d.h_dstOff); cudaFreeHost(d.h_srcSize);
  cudaFreeHost(d.h_dstCap); cudaFreeHost(d.h_result);
  for (auto& s : d.stream) if (s) cudaStreamDestroy(s);
  if (d.descEv) cudaEventDestroy(d.descEv);
}

struct Range { size_t lo, hi; };

// Greedy split of [0, n) into consecutive sub-batches that respect the per-device arena limits.
template <class FI, class FO>
std::vector<Range> make_subbatches(size_t n, size_t maxIn, size_t maxOut, size_t maxItems, FI inBytes, FO outBytes, bool* tooBig) {
  std::vector<Range> r; size_t lo = 0; *tooBig = false;
  while (lo < n) {
    size_t in = 0, out = 0, hi = lo;
    while (hi < n && hi - lo < maxItems) {
      size_t a = align_up(inBytes(hi), 16), b = align_up(outBytes(hi), 16);   // exactly what run_subbatch lays out
      if (in + a > maxIn || out + b > maxOut) break;
      in += a; out += b; hi++;
    }
    if (hi == lo) { *tooBig = true; return r; }   // a single item exceeds the arena
    r.push_back({lo, hi}); lo = hi;
  }
  return r;
}

enum class Op { Decompress, Compress };

// Items that hold more than one data frame (DecompressMultiFrame, ZStdDecompress.cs:2096-2160): after the first pass
// the execute stage has counted them in *more; every further pass decodes the next data frame of each of them.
// Synchronises `st` once per pass (the count has to reach the host).  `a` covers the same items as the first pass.
cudaError_t decode_more_passes(Device& d, DecodeArgs a, u32 slot, cudaStream_t st, int* launches) {
  cudaError_t e;
  while (true) {
    a.pass++; a.more = d.d_more + slot;
    if ((e = cudaMemsetAsync(a.more, 0, 4, st)) != cudaSuccess) return e;
    if ((e = decode_launch(a, st, launches)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(d.h_more + slot, a.more, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    if (d.h_more[slot] == 0) return cudaSuccess;
  }
}

struct Job {
  Op op; int level, checksum;
  const void* const* src; const uint32_t* srcSize; void* const* dst; const uint32_t* dstCap; uint32_t* result;
};

// One sub-batch [lo, hi) on one device: slices over NSTREAMS streams, direct DMA where the layout allows.
// Returns "" or an error description.
std::string run_subbatch(zstdb200_ctx* ctx, Device& d, const Job& j, size_t lo, size_t hi, uint64_t* launches) {
  const size_t m = hi - lo;
  auto fail = [&](const char* what, cudaError_t e) { return std::string(what) + ": " + cudaGetErrorString(e); };
  // ---- layout analysis ----
  bool srcDirect = true, dstDirect = true;
  for (size_t k = 0; k + 1 < m && (srcDirect || dstDirect); k++) {
    const size_t i = lo + k;
    const u8 *s0 = (const u8*)j.src[i], *s1 = (const u8*)j.src[i + 1];
    if (!(s1 >= s0 + j.srcSize[i] && s1 <= s0 + j.srcSize[i] + 64)) srcDirect = false;
    const u8 *d0 = (const u8*)j.dst[i], *d1 = (const u8*)j.dst[i + 1];
    if (d1 != d0 + j.dstCap[i]) dstDirect = false;
  }
  const u8* srcBase = (const u8*)j.src[lo]; u8* dstBase = (u8*)j.dst[lo];
  if (srcDirect) {
    const size_t span = (size_t)((const u8*)j.src[hi - 1] - srcBase) + j.srcSize[hi - 1];
    const size_t lim = j.op == Op::Decompress ? ctx->srcCap : ctx->dstSpan;
    if (span > lim || !srcBase) srcDirect = false;
  }
  if (dstDirect) {
    const size_t span = (size_t)((u8*)j.dst[hi - 1] - dstBase) + j.dstCap[hi - 1];
    const size_t lim = j.op == Op::Decompress ? ctx->dstSpan : ctx->srcCap;
    if (span > lim || !dstBase) dstDirect = false;
  }
  // ---- descriptors (device offsets) ----
  size_t in = 0, out = 0;
  for (size_t k = 0; k < m; k++) {
    const size_t i = lo + k;
    d.h_srcSize[k] = j.srcSize[i]; d.h_dstCap[k] = j.dstCap[i];
    d.h_srcOff[k] = srcDirect ? (u64)((const u8*)j.src[i] - srcBase) : in;
    d.h_dstOff[k] = dstDirect ? (u64)((u8*)j.dst[i] - dstBase) : out;
    in += align_up(j.srcSize[i], 16); out += align_up(j.dstCap[i], 16);
  }
  // ---- slices ----
  static const int nStreams = env_int("ZSTDB200_STREAMS", NSTREAMS, 1, NSTREAMS);
  static const size_t SLICE_BYTES = (size_t)env_int("ZSTDB200_SLICE_MB", (int)(SLICE_BYTES_DEFAULT >> 20), 1, 4096) << 20;
  std::vector<Range> slices;
  {
    size_t a = 0;
    while (a < m) {
      size_t b = a, bytes = 0;
      // a slice is >= SLICE_BYTES and >= MIN_SLICE_ITEMS frames: the kernels take one frame-time however few frames
      // they are given, so slices of a few large frames would serialise on that latency
      static const int envMin = env_int("ZSTDB200_MIN_SLICE_ITEMS", 0, 0, 1 << 30);
      const size_t minItems = envMin ? (size_t)envMin : (j.op == Op::Decompress ? MIN_SLICE_ITEMS_DECODE : MIN_SLICE_ITEMS_ENCODE);
      // The first slices are smaller: output can only start to leave once a slice's kernels are done (one frame-time
      // however small the slice), and the D2H engine then has to be fed without a gap while the big slices finish.
      const size_t k = slices.size(), shrink = j.op == Op::Decompress ? (k < 2 ? 4 : (k < 4 ? 2 : 1)) : 1;
      const size_t wantItems = std::max<size_t>(1, minItems / shrink), wantBytes = SLICE_BYTES / shrink;
      while (b < m && (b - a < wantItems || bytes + std::max(d.h_dstCap[b], d.h_srcSize[b]) <= wantBytes)) { bytes += std::max(d.h_dstCap[b], d.h_srcSize[b]); b++; }
      slices.push_back({a, b}); a = b;
    }
  }
  cudaError_t e;
  // Small copies cost the copy engines about as much as a megabyte each, and they queue between the slices'
  // payload copies: descriptors go up once per sub-batch, results and counters come back once at its end.
  {
    cudaStream_t s0 = d.stream[0];
    e = cudaMemsetAsync(d.d_more, 0, NSTREAMS * 4, s0); if (e) return fail("memset", e);
    e = cudaMemcpyAsync(d.d_srcOff, d.h_srcOff, m * 8, cudaMemcpyHostToDevice, s0); if (e) return fail("H2D desc", e);
    e = cudaMemcpyAsync(d.d_dstOff, d.h_dstOff, m * 8, cudaMemcpyHostToDevice, s0); if (e) return fail("H2D desc", e);
    e = cudaMemcpyAsync(d.d_srcSize, d.h_srcSize, m * 4, cudaMemcpyHostToDevice, s0); if (e) return fail("H2D desc", e);
    e = cudaMemcpyAsync(d.d_dstCap, d.h_dstCap, m * 4, cudaMemcpyHostToDevice, s0); if (e) return fail("H2D desc", e);
    e = cudaEventRecord(d.descEv, s0); if (e) return fail("event", e);
    for (int k = 1; k < NSTREAMS; k++) { e = cudaStreamWaitEvent(d.stream[k], d.descEv, 0); if (e) return fail("event wait", e); }
  }
  static const bool trace = getenv("ZSTDB200_TRACE") != nullptr;   // per-slice timeline on stderr (tuning aid)
  std::vector<cudaEvent_t> tev;
  if (trace) { tev.resize(1 + 3 * slices.size()); for (auto& x : tev) cudaEventCreate(&x); cudaEventRecord(tev[0], d.stream[0]); }
  for (size_t s = 0; s < slices.size(); s++) {
    const size_t a = slices[s].lo, b = slices[s].hi, cnt = b - a;
    cudaStream_t st = d.stream[s % nStreams];
    // input bytes of the slice
    const size_t inLo = d.h_srcOff[a], inHi = d.h_srcOff[b - 1] + d.h_srcSize[b - 1];
    if (srcDirect) {
      if (inHi > inLo) { e = cudaMemcpyAsync(d.d_src + inLo, srcBase + inLo, inHi - inLo, cudaMemcpyHostToDevice, st); if (e) return fail("H2D src", e); }
    } else {
      for (size_t k = a; k < b; k++) if (d.h_srcSize[k]) memcpy(d.h_src + d.h_srcOff[k], j.src[lo + k], d.h_srcSize[k]);
      if (inHi > inLo) { e = cudaMemcpyAsync(d.d_src + inLo, d.h_src + inLo, inHi - inLo, cudaMemcpyHostToDevice, st); if (e) return fail("H2D src", e); }
    }
    if (trace) cudaEventRecord(tev[1 + 3 * s], st);
    int nl = 0;
    if (j.op == Op::Decompress) {
      DecodeArgs ar{d.d_src, d.d_srcOff + a, d.d_srcSize + a, d.d_dst, d.d_dstOff + a, d.d_dstCap + a, d.d_result + a, (u32)cnt, (u32)a,
                    d.d_info + a, d.d_lit, d.d_seq, 0, d.d_more + (s % nStreams)};
      e = decode_launch(ar, st, &nl);
    } else {
      EncodeArgs ar{d.d_src, d.d_srcOff + a, d.d_srcSize + a, d.d_dst, d.d_dstOff + a, d.d_dstCap + a, d.d_result + a, (u32)cnt, (u32)a,
                    j.level, j.checksum, (u32)(s % nStreams)};
      e = encode_launch(ar, d.enc, st, &nl);
    }
    *launches += nl;
    if (e) return fail("kernel launch", e);
    if (trace) cudaEventRecord(tev[2 + 3 * s], st);
    const size_t outLo = d.h_dstOff[a], outHi = d.h_dstOff[b - 1] + d.h_dstCap[b - 1];
    if (outHi > outLo) {
      u8* hostDst = dstDirect ? dstBase + outLo : d.h_dst + outLo;
      e = cudaMemcpyAsync(hostDst, d.d_dst + outLo, outHi - outLo, cudaMemcpyDeviceToHost, st); if (e) return fail("D2H dst", e);
    }
    if (trace) cudaEventRecord(tev[3 + 3 * s], st);
  }
  for (auto& st : d.stream) { e = cudaStreamSynchronize(st); if (e) return fail("stream sync", e); }
  if (trace) {
    for (size_t s = 0; s < slices.size(); s++) {
      float t1 = 0, t2 = 0, t3 = 0;
      cudaEventElapsedTime(&t1, tev[0], tev[1 + 3 * s]); cudaEventElapsedTime(&t2, tev[0], tev[2 + 3 * s]); cudaEventElapsedTime(&t3, tev[0], tev[3 + 3 * s]);
      fprintf(stderr, "[zstdb200] slice %2zu items %6zu: h2d done %7.3f  kernels done %7.3f  d2h done %7.3f ms\n", s, slices[s].hi - slices[s].lo, t1, t2, t3);
    }
    for (auto& x : tev) cudaEventDestroy(x);
  }
  e = cudaMemcpyAsync(d.h_result, d.d_result, m * 4, cudaMemcpyDeviceToHost, d.stream[0]); if (e) return fail("D2H result", e);
  e = cudaMemcpyAsync(d.h_more, d.d_more, NSTREAMS * 4, cudaMemcpyDeviceToHost, d.stream[0]); if (e) return fail("D2H counters", e);
  e = cudaStreamSynchronize(d.stream[0]); if (e) return fail("stream sync", e);
  u32 anyMore = 0;
  for (int k = 0; k < NSTREAMS; k++) anyMore |= d.h_more[k];
  if (j.op == Op::Decompress && anyMore) {
    // rare path: some items hold several data frames.  Everything is still resident: decode the remaining frames
    // pass by pass over the whole sub-batch, then fetch results and output again.
    cudaStream_t st = d.stream[0]; int nl = 0;
    DecodeArgs ar{d.d_src, d.d_srcOff, d.d_srcSize, d.d_dst, d.d_dstOff, d.d_dstCap, d.d_result, (u32)m, 0, d.d_info, d.d_lit, d.d_seq, 0, nullptr};
    e = decode_more_passes(d, ar, 0, st, &nl); *launches += nl;
    if (e) return fail("multi-frame passes", e);
    e = cudaMemcpyAsync(d.h_result, d.d_result, m * 4, cudaMemcpyDeviceToHost, st); if (e) return fail("D2H result", e);
    const size_t outLo = d.h_dstOff[0], outHi = d.h_dstOff[m - 1] + d.h_dstCap[m - 1];
    if (outHi > outLo) { e = cudaMemcpyAsync(dstDirect ? dstBase + outLo : d.h_dst + outLo, d.d_dst + outLo, outHi - outLo, cudaMemcpyDeviceToHost, st); if (e) return fail("D2H dst", e); }
    e = cudaStreamSynchronize(st); if (e) return fail("stream sync", e);
  }
  for (size_t k = 0; k < m; k++) {
    const size_t i = lo + k; const u32 r = d.h_result[k];
    j.result[i] = r;
    // staged output: copy what was produced; on error the reference leaves dst partially written, we copy nothing
    if (!dstDirect && !is_err(r) && r) memcpy(j.dst[i], d.h_dst + d.h_dstOff[k], std::min<u32>(r, j.dstCap[i]));
  }
  return "";
}

int run_host_batch(zstdb200_ctx* ctx, const Job& j, size_t n) {
  if (n == 0) return 0;
  if (!j.src || !j.srcSize || !j.dst || !j.dstCap || !j.result) { ctx->err = "null argument"; return 1; }
  // staging limits.  decode: in = frames (d_src, srcCap), out = content (dstSpan: what the literal / sequence arenas
  // are sized for); encode: in = raw chunks (dstSpan: what the encoder arenas are sized for), out = frames (d_dst, srcCap)
  const size_t maxIn = j.op == Op::Decompress ? ctx->srcCap : ctx->dstSpan;
  const size_t maxOut = j.op == Op::Decompress ? ctx->dstSpan : ctx->srcCap;
  bool tooBig = false;
  std::vector<Range> subs = make_subbatches(n, maxIn, maxOut, ctx->maxItems,
      [&](size_t i) { return (size_t)j.srcSize[i]; }, [&](size_t i) { return (size_t)j.dstCap[i]; }, &tooBig);
  if (tooBig) { ctx->err = "an item is larger than the context's max_batch_bytes"; return 1; }
  // a single sub-batch on a multi-GPU context is re-cut so that every device gets a share
  const size_t nd = ctx->dev.size();
  if (nd > 1 && subs.size() < nd) {
    std::vector<Range> cut;
    for (auto& r : subs) {
      // equal shares of bytes (in + out), not of items: a raw-block frame costs a copy, a text frame a full decode
      const size_t parts = std::min(nd, r.hi - r.lo);
      uint64_t total = 0;
      for (size_t i = r.lo; i < r.hi; i++) total += (uint64_t)j.srcSize[i] + j.dstCap[i] + 1;
      size_t lo = r.lo; uint64_t acc = 0;
      for (size_t p = 0; p < parts; p++) {
        size_t hi = lo;
        const uint64_t want = total * (p + 1) / parts;
        while (hi < r.hi && (acc < want || hi == lo) && (r.hi - hi) > (parts - 1 - p)) { acc += (uint64_t)j.srcSize[hi] + j.dstCap[hi] + 1; hi++; }
        if (p + 1 == parts) hi = r.hi;
        cut.push_back({lo, hi}); lo = hi;
      }
    }
    subs.swap(cut);
  }
  std::vector<std::string> errs(nd);
  std::vector<uint64_t> launches(nd, 0);
  auto worker = [&](size_t di) {
    Device& d = ctx->dev[di];
    cudaError_t e = cudaSetDevice(d.id);
    if (e != cudaSuccess) { errs[di] = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return; }
    for (size_t s = di; s < subs.size(); s += nd) {
      errs[di] = run_subbatch(ctx, d, j, subs[s].lo, subs[s].hi, &launches[di]);
      if (!errs[di].empty()) return;
    }
  };
  if (nd == 1 || subs.size() == 1) { for (size_t di = 0; di < nd; di++) worker(di); }
  else { std::vector<std::thread> th; for (size_t di = 0; di < nd; di++) th.emplace_back(worker, di); for (auto& t : th) t.join(); }
  for (size_t di = 0; di < nd; di++) ctx->launches += launches[di];
  for (size_t di = 0; di < nd; di++) if (!errs[di].empty()) { ctx->err = "device " + std::to_string(ctx->dev[di].id) + ": " + errs[di]; return 1; }
  return 0;
}

}  // namespace

extern "C" {

const char* zstdb200_version(void) { return "zstdb200 0.1 (sm_100a)"; }

int zstdb200_create(zstdb200_ctx** out, const int* devices, int n_devices, size_t max_batch_bytes) {
  if (!out) return 1;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return 2;   // no CUDA device: there is no CPU fallback
  zstdb200_ctx* ctx = new zstdb200_ctx();
  if (max_batch_bytes < (1u << 20)) max_batch_bytes = 1u << 20;
  ctx->maxBatch = align_up(max_batch_bytes, 4096);
  ctx->maxItems = std::max<size_t>(65536, ctx->maxBatch / 1024);
  ctx->dstSpan = ctx->maxBatch + 16 * ctx->maxItems;
  ctx->srcCap = ctx->maxBatch + ctx->maxBatch / 128 + 128 * ctx->maxItems;   // >= sum of compress bounds of a full batch
  int one = 0;
  if (!devices || n_devices <= 0) { devices = &one; n_devices = 1; }
  for (int i = 0; i < n_devices; i++) {
    if (devices[i] < 0 || devices[i] >= count) { delete ctx; return 3; }
    Device d; d.id = devices[i]; ctx->dev.push_back(d);
  }
  for (auto& d : ctx->dev)
    if (alloc_device(ctx, d)) { fprintf(stderr, "zstdb200_create: %s\n", ctx->err.c_str()); for (auto& x : ctx->dev) free_device(x); delete ctx; return 4; }
  *out = ctx;
  return 0;
}

void zstdb200_destroy(zstdb200_ctx* ctx) {
  if (!ctx) return;
  for (auto& d : ctx->dev) free_device(d);
  delete ctx;
}

const char* zstdb200_last_error(const zstdb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
size_t zstdb200_max_items(const zstdb200_ctx* ctx) { return ctx ? ctx->maxItems : 0; }
int zstdb200_device_count(const zstdb200_ctx* ctx) { return ctx ? (int)ctx->dev.size() : 0; }
uint64_t zstdb200_kernel_launches(const zstdb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int zstdb200_is_error(uint32_t code) { return is_err(code); }

// Host-only frame-header parse, ZStdDecompress.cs:518-531, 590-622 (same rules as parse_item in zb_format.cuh).
uint64_t zstdb200_get_decompressed_size(const void* src, uint32_t srcSize) {
  const u8* p = (const u8*)src;
  if (!p || srcSize < 5) return 0;
  u32 magic = ld32(p);
  if (magic != MAGIC) return 0;                       // skippable frame -> 0, unknown prefix -> 0
  u32 fhd = p[4], did = fhd & 3, single = (fhd >> 5) & 1, fcsId = fhd >> 6;
  u32 fhs = 5 + (single ? 0 : 1) + (did == 3 ? 4 : did) + (fcsId == 0 ? 0 : (1u << fcsId)) + ((single && fcsId == 0) ? 1 : 0);
  if (srcSize < fhs) return 0;
  if (fhd & 0x08) return 0;
  u32 pos = 5;
  if (!single) { u32 wl = p[pos++]; if ((wl >> 3) + 10 > 30) return 0; }
  pos += did == 3 ? 4 : did;
  u64 fcs;
  if (fcsId == 0) { if (!single) return 0; fcs = p[pos]; }
  else if (fcsId == 1) fcs = ld16(p + pos) + 256;
  else if (fcsId == 2) fcs = ld32(p + pos);
  else fcs = ld64(p + pos);
  return fcs >= 0xFFFFFFFFFFFFFFFEull ? 0 : fcs;
}

int zstdb200_decompress_batch(zstdb200_ctx* ctx, const void* const* src, const uint32_t* srcSize,
                              void* const* dst, const uint32_t* dstCap, uint32_t* result, size_t n) {
  if (!ctx) return 1;
  ctx->err.clear();
  Job j{Op::Decompress, 0, 0, src, srcSize, dst, dstCap, result};
  return run_host_batch(ctx, j, n);
}

uint32_t zstdb200_decompress(zstdb200_ctx* ctx, void* dst, uint32_t dstCapacity, const void* src, uint32_t srcSize) {
  uint32_t r = zerr(ZE_GENERIC);
  const void* s = src; void* d = dst;
  static unsigned char dummy[16];
  if (!d) d = dummy;
  if (!s) s = dummy;
  if (zstdb200_decompress_batch(ctx, &s, &srcSize, &d, &dstCapacity, &r, 1)) return zerr(ZE_GENERIC);
  return r;
}

// Device-resident decode: one launch sequence over the whole batch on the caller's stream (slicing a resident batch
// over several streams was measured slower: whole-batch kernels fill the GPU better).  The count of items that
// hold a further data frame has to reach the host, so this entry point synchronises the stream once per pass.
static int decode_device(zstdb200_ctx* ctx, Device& d, const DecodeArgs& a, cudaStream_t user, cudaEvent_t* marks) {
  int nl = 0;
  DecodeArgs b = a; b.pass = 0; b.more = d.d_more;
  CK(cudaMemsetAsync(b.more, 0, 4, user));
  CK(decode_launch(b, user, &nl, marks));
  CK(cudaMemcpyAsync(d.h_more, b.more, 4, cudaMemcpyDeviceToHost, user));
  CK(cudaStreamSynchronize(user));
  if (d.h_more[0]) CK(decode_more_passes(d, b, 0, user, &nl));
  ctx->launches += nl;
  return 0;
}

int zstdb200_decompress_batch_device(zstdb200_ctx* ctx, int device_index, const void* src_base, const uint64_t* src_off,
                                     const uint32_t* src_size, void* dst_base, const uint64_t* dst_off, const uint32_t* dst_cap,
                                     uint32_t* result, size_t n, void* stream) {
  if (!ctx) return 1;
  ctx->err.clear();
  if (device_index < 0 || device_index >= (int)ctx->dev.size()) { ctx->err = "bad device_index"; return 1; }
  if (n > ctx->maxItems) { ctx->err = "n exceeds zstdb200_max_items"; return 1; }
  Device& d = ctx->dev[device_index];
  CK(cudaSetDevice(d.id));
  DecodeArgs a{(const u8*)src_base, src_off, src_size, (u8*)dst_base, dst_off, dst_cap, result, (u32)n, 0, d.d_info, d.d_lit, d.d_seq, 0, nullptr};
  return decode_device(ctx, d, a, stream ? (cudaStream_t)stream : d.stream[0], nullptr);
}

// Same work as one launch sequence with CUDA events between the kernels: per-kernel device times for the
// bench's roofline block.
int zstdb200_decompress_batch_device_timed(zstdb200_ctx* ctx, int device_index, const void* src_base, const uint64_t* src_off,
                                           const uint32_t* src_size, void* dst_base, const uint64_t* dst_off, const uint32_t* dst_cap,
                                           uint32_t* result, size_t n, void* stream, float* kernel_ms, int max_kernels) {
  if (!ctx) return 1;
  ctx->err.clear();
  if (device_index < 0 || device_index >= (int)ctx->dev.size()) { ctx->err = "bad device_index"; return 1; }
  if (n > ctx->maxItems) { ctx->err = "n exceeds zstdb200_max_items"; return 1; }
  Device& d = ctx->dev[device_index];
  CK(cudaSetDevice(d.id));
  cudaStream_t st = stream ? (cudaStream_t)stream : d.stream[0];
  cudaEvent_t ev[DECODE_KERNELS + 1];
  for (auto& e : ev) CK(cudaEventCreate(&e));
  DecodeArgs a{(const u8*)src_base, src_off, src_size, (u8*)dst_base, dst_off, dst_cap, result, (u32)n, 0, d.d_info, d.d_lit, d.d_seq, 0, nullptr};
  if (decode_device(ctx, d, a, st, ev)) return 1;
  CK(cudaStreamSynchronize(st));
  for (int k = 0; k < DECODE_KERNELS && k < max_kernels; k++) CK(cudaEventElapsedTime(&kernel_ms[k], ev[k], ev[k + 1]));
  for (auto& e : ev) cudaEventDestroy(e);
  return 0;
}
const char* zstdb200_decode_kernel_name(int k) { return (k >= 0 && k < DECODE_KERNELS) ? kDecodeKernelNames[k] : ""; }

size_t zstdb200_compress_bound(size_t srcSize) { return encode_bound(srcSize); }

int zstdb200_compress_batch(zstdb200_ctx* ctx, int level, int checksum, const void* const* src, const uint32_t* srcSize,
                            void* const* dst, const uint32_t* dstCap, uint32_t* result, size_t n) {
  if (!ctx) return 1;
  ctx->err.clear();
  if (level < 1 || level > 3) { ctx->err = "level must be 1..3"; return 1; }
  Job j{Op::Compress, level, checksum, src, srcSize, dst, dstCap, result};
  return run_host_batch(ctx, j, n);
}

uint32_t zstdb200_compress(zstdb200_ctx* ctx, int level, int checksum, void* dst, uint32_t dstCapacity, const void* src, uint32_t srcSize) {
  uint32_t r = zerr(ZE_GENERIC);
  const void* s = src; void* d = dst;
  static unsigned char dummy[16];
  if (!s) s = dummy;
  if (!d) d = dummy;
  if (zstdb200_compress_batch(ctx, level, checksum, &s, &srcSize, &d, &dstCapacity, &r, 1)) return zerr(ZE_GENERIC);
  return r;
}

int zstdb200_compress_batch_device(zstdb200_ctx* ctx, int device_index, int level, int checksum, const void* src_base,
                                   const uint64_t* src_off, const uint32_t* src_size, void* dst_base, const uint64_t* dst_off,
                                   const uint32_t* dst_cap, uint32_t* result, size_t n, void* stream) {
  if (!ctx) return 1;
  ctx->err.clear();
  if (device_index < 0 || device_index >= (int)ctx->dev.size()) { ctx->err = "bad device_index"; return 1; }
  if (level < 1 || level > 3) { ctx->err = "level must be 1..3"; return 1; }
  if (n > ctx->maxItems) { ctx->err = "n exceeds zstdb200_max_items"; return 1; }
  Device& d = ctx->dev[device_index];
  CK(cudaSetDevice(d.id));
  EncodeArgs a{(const u8*)src_base, src_off, src_size, (u8*)dst_base, dst_off, dst_cap, result, (u32)n, 0, level, checksum, ENC_EXCLUSIVE};
  int nl = 0;
  CK(encode_launch(a, d.enc, stream ? (cudaStream_t)stream : d.stream[0], &nl));
  ctx->launches += nl;
  return 0;
}

// zstdb200_compress_batch_device with CUDA events between its kernels (see zstdb200_decompress_batch_device_timed).
int zstdb200_compress_batch_device_timed(zstdb200_ctx* ctx, int device_index, int level, int checksum, const void* src_base,
                                         const uint64_t* src_off, const uint32_t* src_size, void* dst_base, const uint64_t* dst_off,
                                         const uint32_t* dst_cap, uint32_t* result, size_t n, void* stream, float* kernel_ms, int max_kernels) {
  if (!ctx) return 1;
  ctx->err.clear();
  if (device_index < 0 || device_index >= (int)ctx->dev.size()) { ctx->err = "bad device_index"; return 1; }
  if (level < 1 || level > 3) { ctx->err = "level must be 1..3"; return 1; }
  if (n > ctx->maxItems) { ctx->err = "n exceeds zstdb200_max_items"; return 1; }
  Device& d = ctx->dev[device_index];
  CK(cudaSetDevice(d.id));
  cudaStream_t st = stream ? (cudaStream_t)stream : d.stream[0];
  cudaEvent_t ev[ENCODE_KERNELS + 1];
  for (auto& e : ev) CK(cudaEventCreate(&e));
  EncodeArgs a{(const u8*)src_base, src_off, src_size, (u8*)dst_base, dst_off, dst_cap, result, (u32)n, 0, level, checksum, ENC_EXCLUSIVE};
  int nl = 0;
  CK(encode_launch(a, d.enc, st, &nl, ev));
  ctx->launches += nl;
  CK(cudaStreamSynchronize(st));
  for (int k = 0; k < ENCODE_KERNELS && k < max_kernels; k++) CK(cudaEventElapsedTime(&kernel_ms[k], ev[k], ev[k + 1]));
  for (auto& e : ev) cudaEventDestroy(e);
  return 0;
}
const char* zstdb200_encode_kernel_name(int k) { return (k >= 0 && k < ENCODE_KERNELS) ? kEncodeKernelNames[k] : ""; }

void* zstdb200_host_alloc(size_t bytes) { void* p = nullptr; return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr; }
void zstdb200_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
