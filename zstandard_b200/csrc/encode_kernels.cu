// encode_kernels.cu — sm_100a kernels of the batched zstd frame encoder.
//
//   k_encode  : thread / frame   match finding (fast / double-fast), sequence store, Huffman + FSE table
//                                construction, bitstream encode, block + frame assembly      (zb_encode.cuh)
//   k_enc_xxh : 4 lanes / frame  XXH64 content checksum appended to the frame                  (xxh_device.cuh)
//
// Every thread owns a scratch slot in HBM (hash tables, literal buffer, sequence store, block staging) and
// walks the frames slot, slot + slots, ...  Frames are independent, so there is no inter-thread communication.
#include "encode_kernels.cuh"
#include "zb_encode.cuh"
#include "xxh_device.cuh"

namespace zb {

static constexpr u32 kSeqCap = BLOCKSIZE_MAX / 4 + 64;

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
// slot layout (level-3 sizes so that one arena serves every level)
__host__ __device__ inline size_t slot_bytes() {
  return align16((size_t)enc_table_words(3) * 4) + align16(BLOCKSIZE_MAX + 64) + align16((size_t)kSeqCap * 8) + align16((size_t)kSeqCap * 3) +
         align16(3 * 514 * 2 + 32) + align16(BLOCKSIZE_MAX + 4096);
}
__device__ inline EncScratch slot_view(u8* p) {
  EncScratch sc;
  sc.table = (u32*)p; p += align16((size_t)enc_table_words(3) * 4);
  sc.lits = p; p += align16(BLOCKSIZE_MAX + 64);
  sc.seqs = (u32*)p; p += align16((size_t)kSeqCap * 8);
  sc.seqCap = kSeqCap;
  sc.codes = p; p += align16((size_t)kSeqCap * 3);
  sc.ctables = (u16*)p; p += align16(3 * 514 * 2 + 32);
  sc.tmp = p;
  return sc;
}

__global__ void __launch_bounds__(64) k_encode(EncodeArgs a, u8* arena, size_t slotBytes, u32 slots) {
  const u32 slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= slots) return;
  const EncScratch sc = slot_view(arena + (size_t)slot * slotBytes);
  for (u32 f = slot; f < a.n; f += slots)
    a.result[f] = encode_frame(a.src_base + a.src_off[f], a.src_size[f], a.dst_base + a.dst_off[f], a.dst_cap[f], a.level, a.checksum, sc);
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
    u8* p = a.dst_base + a.dst_off[f] + r;   // encode_frame reserved these 4 bytes
    p[0] = (u8)h; p[1] = (u8)(h >> 8); p[2] = (u8)(h >> 16); p[3] = (u8)(h >> 24);
    a.result[f] = r + 4;
  }
}

size_t encode_bound(size_t srcSize) { return srcSize + (srcSize >> 8) + 32 + 3 * ((srcSize >> 17) + 1); }

cudaError_t encode_alloc(EncodeScratch& s, size_t, size_t) { s.arena = nullptr; s.slots = 0; s.slotBytes = slot_bytes(); return cudaSuccess; }
void encode_free(EncodeScratch& s) { if (s.arena) cudaFree(s.arena); s.arena = nullptr; s.slots = 0; }

cudaError_t encode_launch(const EncodeArgs& a, EncodeScratch& s, cudaStream_t st, int* launches) {
  if (a.n == 0) return cudaSuccess;
  if (!s.arena) {
    // one slot per resident thread: 148 SMs x 2 warps x 32 lanes x 2 = 18944 slots (~1 MiB each)
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    u32 want = (u32)sms * 128;
    size_t freeB = 0, totalB = 0; cudaMemGetInfo(&freeB, &totalB);
    while (want > 64 && (size_t)want * s.slotBytes > freeB / 2) want /= 2;
    cudaError_t e = cudaMalloc(&s.arena, (size_t)want * s.slotBytes);
    if (e != cudaSuccess) return e;
    s.slots = want;
  }
  const u32 slots = a.n < s.slots ? a.n : s.slots;
  k_encode<<<(slots + 63) / 64, 64, 0, st>>>(a, s.arena, s.slotBytes, slots);
  if (launches) *launches += 1;
  if (a.checksum) { k_enc_xxh<<<(a.n * 4 + 127) / 128, 128, 0, st>>>(a); if (launches) *launches += 1; }
  return cudaGetLastError();
}

}  // namespace zb
