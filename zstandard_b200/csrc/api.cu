// api.cu — host side of libzstdb200: contexts, device arenas, staging, sharding, the C ABI of include/zstdb200.h.
//
// Host responsibilities: build flat descriptor tables, shard items over the context's GPUs by
// bytes (no collective: frames are independent, ZStdDecompress.cs:2478-2499 resets all state per frame), move
// bytes host<->device and launch the kernels.  No decoding or encoding arithmetic happens on the host.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
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

struct Device {
  int id = 0;
  cudaStream_t stream = nullptr;
  // device memory
  u8 *d_src = nullptr, *d_dst = nullptr;       // staging for the host-pointer API
  u64 *d_srcOff = nullptr, *d_dstOff = nullptr; u32 *d_srcSize = nullptr, *d_dstCap = nullptr, *d_result = nullptr;
  FrameInfo* d_info = nullptr; u8* d_lit = nullptr; SeqRec* d_seq = nullptr;
  EncodeScratch enc;                            // encoder arenas (encode_kernels.cuh)
  // pinned host memory
  u8 *h_src = nullptr, *h_dst = nullptr;
  u64 *h_srcOff = nullptr, *h_dstOff = nullptr; u32 *h_srcSize = nullptr, *h_dstCap = nullptr, *h_result = nullptr;
};

}  // namespace

struct zstdb200_ctx {
  std::vector<Device> dev;
  size_t maxBatch = 0, maxItems = 0, srcCap = 0;
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
  CK(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
  const size_t items = ctx->maxItems;
  CK(cudaMalloc(&d.d_src, ctx->srcCap + 256));
  CK(cudaMalloc(&d.d_dst, ctx->maxBatch + 16 * items + 256));
  CK(cudaMalloc(&d.d_srcOff, items * 8)); CK(cudaMalloc(&d.d_dstOff, items * 8));
  CK(cudaMalloc(&d.d_srcSize, items * 4)); CK(cudaMalloc(&d.d_dstCap, items * 4)); CK(cudaMalloc(&d.d_result, items * 4));
  CK(cudaMalloc(&d.d_info, items * sizeof(FrameInfo)));
  const u64 dstSpan = ctx->maxBatch + 16 * items;
  CK(cudaMalloc(&d.d_lit, decode_lit_arena_bytes(dstSpan, items)));
  CK(cudaMalloc(&d.d_seq, decode_seq_arena_bytes(dstSpan, items)));
  CK(cudaMallocHost(&d.h_src, ctx->srcCap + 256)); CK(cudaMallocHost(&d.h_dst, ctx->maxBatch + 16 * items + 256));
  CK(cudaMallocHost(&d.h_srcOff, items * 8)); CK(cudaMallocHost(&d.h_dstOff, items * 8));
  CK(cudaMallocHost(&d.h_srcSize, items * 4)); CK(cudaMallocHost(&d.h_dstCap, items * 4)); CK(cudaMallocHost(&d.h_result, items * 4));
  CK(decode_configure());
  cudaError_t ee = encode_alloc(d.enc, ctx->maxBatch, items);
  if (ee != cudaSuccess) { ctx->err = std::string("encoder arena allocation failed: ") + cudaGetErrorString(ee); return 1; }
  return 0;
}

void free_device(Device& d) {
  cudaSetDevice(d.id);
  if (d.stream) cudaStreamSynchronize(d.stream);
  cudaFree(d.d_src); cudaFree(d.d_dst); cudaFree(d.d_srcOff); cudaFree(d.d_dstOff); cudaFree(d.d_srcSize); cudaFree(d.d_dstCap);
  cudaFree(d.d_result); cudaFree(d.d_info); cudaFree(d.d_lit); cudaFree(d.d_seq);
  encode_free(d.enc);
  cudaFreeHost(d.h_src); cudaFreeHost(d.h_dst); cudaFreeHost(d.h_srcOff); cudaFreeHost(d.h_dstOff); cudaFreeHost(d.h_srcSize);
  cudaFreeHost(d.h_dstCap); cudaFreeHost(d.h_result);
  if (d.stream) cudaStreamDestroy(d.stream);
}

// A sub-batch of items [lo, hi) assigned to one device.
struct Range { size_t lo, hi; };

// Greedy split of [0, n) into consecutive sub-batches that respect the per-device arena limits.
// inBytes(i)/outBytes(i): staging bytes needed by item i on the input / output side.
template <class FI, class FO>
std::vector<Range> make_subbatches(size_t n, size_t maxIn, size_t maxOut, size_t maxItems, FI inBytes, FO outBytes, bool* tooBig) {
  std::vector<Range> r; size_t lo = 0; *tooBig = false;
  while (lo < n) {
    size_t in = 0, out = 0, hi = lo;
    while (hi < n && hi - lo < maxItems) {
      size_t a = align_up(inBytes(hi), 16), b = align_up(outBytes(hi), 16);
      if (in + a > maxIn || out + b > maxOut) break;
      in += a; out += b; hi++;
    }
    if (hi == lo) { *tooBig = true; return r; }   // a single item exceeds the arena
    r.push_back({lo, hi}); lo = hi;
  }
  return r;
}

enum class Op { Decompress, Compress };

// Runs one op over items [0,n) with host pointers: sub-batches are dealt round-robin to devices; each device
// handles its sub-batches in order on its own stream.  One host thread per device drives staging copies.
int run_host_batch(zstdb200_ctx* ctx, Op op, int level, int checksum, const void* const* src, const uint32_t* srcSize,
                   void* const* dst, const uint32_t* dstCap, uint32_t* result, size_t n) {
  if (n == 0) return 0;
  if (!src || !srcSize || !dst || !dstCap || !result) { ctx->err = "null argument"; return 1; }
  // staging limits: decode: in = compressed (srcCap), out = raw (maxBatch); encode: in = raw (maxBatch), out = frames (srcCap)
  const size_t maxIn = op == Op::Decompress ? ctx->srcCap : ctx->maxBatch;
  const size_t maxOut = op == Op::Decompress ? ctx->maxBatch : ctx->srcCap;
  bool tooBig = false;
  std::vector<Range> subs = make_subbatches(n, maxIn, maxOut, ctx->maxItems,
      [&](size_t i) { return (size_t)srcSize[i]; }, [&](size_t i) { return (size_t)dstCap[i]; }, &tooBig);
  if (tooBig) { ctx->err = "an item is larger than the context's max_batch_bytes"; return 1; }
  const size_t nd = ctx->dev.size();
  std::vector<std::string> errs(nd);
  std::vector<uint64_t> launches(nd, 0);
  auto worker = [&](size_t di) {
    Device& d = ctx->dev[di];
    auto fail = [&](const char* what, cudaError_t e) { errs[di] = std::string(what) + ": " + cudaGetErrorString(e); };
    cudaError_t e = cudaSetDevice(d.id);
    if (e != cudaSuccess) { fail("cudaSetDevice", e); return; }
    for (size_t s = di; s < subs.size(); s += nd) {
      const Range rg = subs[s]; const size_t m = rg.hi - rg.lo;
      // ---- gather into pinned staging, build descriptors ----
      size_t in = 0, out = 0;
      for (size_t k = 0; k < m; k++) {
        const size_t i = rg.lo + k;
        d.h_srcOff[k] = in; d.h_srcSize[k] = srcSize[i]; d.h_dstOff[k] = out; d.h_dstCap[k] = dstCap[i];
        if (srcSize[i]) memcpy(d.h_src + in, src[i], srcSize[i]);
        in += align_up(srcSize[i], 16); out += align_up(dstCap[i], 16);
      }
      e = cudaMemcpyAsync(d.d_src, d.h_src, in, cudaMemcpyHostToDevice, d.stream); if (e) { fail("H2D src", e); return; }
      e = cudaMemcpyAsync(d.d_srcOff, d.h_srcOff, m * 8, cudaMemcpyHostToDevice, d.stream); if (e) { fail("H2D", e); return; }
      e = cudaMemcpyAsync(d.d_dstOff, d.h_dstOff, m * 8, cudaMemcpyHostToDevice, d.stream); if (e) { fail("H2D", e); return; }
      e = cudaMemcpyAsync(d.d_srcSize, d.h_srcSize, m * 4, cudaMemcpyHostToDevice, d.stream); if (e) { fail("H2D", e); return; }
      e = cudaMemcpyAsync(d.d_dstCap, d.h_dstCap, m * 4, cudaMemcpyHostToDevice, d.stream); if (e) { fail("H2D", e); return; }
      int nl = 0;
      if (op == Op::Decompress) {
        DecodeArgs a{d.d_src, d.d_srcOff, d.d_srcSize, d.d_dst, d.d_dstOff, d.d_dstCap, d.d_result, (u32)m, d.d_info, d.d_lit, d.d_seq};
        e = decode_launch(a, d.stream, &nl);
      } else {
        EncodeArgs a{d.d_src, d.d_srcOff, d.d_srcSize, d.d_dst, d.d_dstOff, d.d_dstCap, d.d_result, (u32)m, level, checksum};
        e = encode_launch(a, d.enc, d.stream, &nl);
      }
      launches[di] += nl;
      if (e) { fail("kernel launch", e); return; }
      e = cudaMemcpyAsync(d.h_result, d.d_result, m * 4, cudaMemcpyDeviceToHost, d.stream); if (e) { fail("D2H", e); return; }
      e = cudaMemcpyAsync(d.h_dst, d.d_dst, out, cudaMemcpyDeviceToHost, d.stream); if (e) { fail("D2H dst", e); return; }
      e = cudaStreamSynchronize(d.stream); if (e) { fail("stream sync", e); return; }
      // ---- scatter ----
      for (size_t k = 0; k < m; k++) {
        const size_t i = rg.lo + k; const u32 r = d.h_result[k];
        result[i] = r;
        // on error the reference leaves dst partially written; we copy nothing
        if (!is_err(r) && r) memcpy(dst[i], d.h_dst + d.h_dstOff[k], std::min<u32>(r, dstCap[i]));
      }
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
  ctx->srcCap = ctx->maxBatch + ctx->maxBatch / 128 + 64 * ctx->maxItems;   // >= sum of compress bounds of a full batch
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

// Host-only frame-header parse, ZStdDecompress.cs:518-531, 590-622 (same function the kernels use).
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
  return run_host_batch(ctx, Op::Decompress, 0, 0, src, srcSize, dst, dstCap, result, n);
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

int zstdb200_decompress_batch_device(zstdb200_ctx* ctx, int device_index, const void* src_base, const uint64_t* src_off,
                                     const uint32_t* src_size, void* dst_base, const uint64_t* dst_off, const uint32_t* dst_cap,
                                     uint32_t* result, size_t n, void* stream) {
  if (!ctx) return 1;
  ctx->err.clear();
  if (device_index < 0 || device_index >= (int)ctx->dev.size()) { ctx->err = "bad device_index"; return 1; }
  if (n > ctx->maxItems) { ctx->err = "n exceeds zstdb200_max_items"; return 1; }
  Device& d = ctx->dev[device_index];
  CK(cudaSetDevice(d.id));
  DecodeArgs a{(const u8*)src_base, src_off, src_size, (u8*)dst_base, dst_off, dst_cap, result, (u32)n, d.d_info, d.d_lit, d.d_seq};
  int nl = 0;
  CK(decode_launch(a, stream ? (cudaStream_t)stream : d.stream, &nl));
  ctx->launches += nl;
  return 0;
}

// Same launch with CUDA events between the kernels: per-kernel device times for the bench's roofline block.
int zstdb200_decompress_batch_device_timed(zstdb200_ctx* ctx, int device_index, const void* src_base, const uint64_t* src_off,
                                           const uint32_t* src_size, void* dst_base, const uint64_t* dst_off, const uint32_t* dst_cap,
                                           uint32_t* result, size_t n, void* stream, float* kernel_ms, int max_kernels) {
  if (!ctx) return 1;
  ctx->err.clear();
  if (device_index < 0 || device_index >= (int)ctx->dev.size()) { ctx->err = "bad device_index"; return 1; }
  if (n > ctx->maxItems) { ctx->err = "n exceeds zstdb200_max_items"; return 1; }
  Device& d = ctx->dev[device_index];
  CK(cudaSetDevice(d.id));
  cudaStream_t st = stream ? (cudaStream_t)stream : d.stream;
  cudaEvent_t ev[DECODE_KERNELS + 1];
  for (auto& e : ev) CK(cudaEventCreate(&e));
  DecodeArgs a{(const u8*)src_base, src_off, src_size, (u8*)dst_base, dst_off, dst_cap, result, (u32)n, d.d_info, d.d_lit, d.d_seq};
  int nl = 0;
  CK(decode_launch(a, st, &nl, ev));
  ctx->launches += nl;
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
  return run_host_batch(ctx, Op::Compress, level, checksum, src, srcSize, dst, dstCap, result, n);
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
  EncodeArgs a{(const u8*)src_base, src_off, src_size, (u8*)dst_base, dst_off, dst_cap, result, (u32)n, level, checksum};
  int nl = 0;
  CK(encode_launch(a, d.enc, stream ? (cudaStream_t)stream : d.stream, &nl));
  ctx->launches += nl;
  return 0;
}

void* zstdb200_host_alloc(size_t bytes) { void* p = nullptr; return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr; }
void zstdb200_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
