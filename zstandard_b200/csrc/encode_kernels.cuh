// encode_kernels.cuh — launch interface of the encoder kernels (encode_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "zb_common.cuh"

namespace zb {

// All pointers are device pointers.  Item i: src_base[src_off[i] .. +src_size[i]) -> one zstd frame written at
// dst_base[dst_off[i] ..] (at most dst_cap[i] bytes); result[i] = frame size or an error code.
struct EncodeArgs {
  const u8* src_base; const u64* src_off; const u32* src_size;
  u8* dst_base; const u64* dst_off; const u32* dst_cap;
  u32* result; u32 n;
  u32 item_base;   // unused by the encoder's slot-based scratch; kept symmetric with DecodeArgs
  int level, checksum;
};

// Per-device scratch owned by the context: `slots` independent work areas; a GPU thread owns one slot and
// encodes frames slot, slot + slots, ... with it.
struct EncodeScratch {
  u8* arena = nullptr;
  size_t slotBytes = 0; u32 slots = 0;
};

size_t encode_bound(size_t srcSize);
cudaError_t encode_alloc(EncodeScratch& s, size_t maxBatchBytes, size_t maxItems);   // lazy: allocates on first launch
void encode_free(EncodeScratch& s);
cudaError_t encode_launch(const EncodeArgs& a, EncodeScratch& s, cudaStream_t st, int* launches);

}  // namespace zb
