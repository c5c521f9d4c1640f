// transform.cuh -- launch interface of the normalisation kernels (transform.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "olm_format.h"
#include "scan.cuh"

namespace olm {

// Summary of one 16 KiB block of a window (transform.cu count / resolve / write passes).
struct alignas(16) TfBlock {
  uint32_t count;    // kept bytes of the block assuming no whitespace run is carried in
  uint32_t out_base; // resolved: offset of the block's first kept byte in the normalised window
  uint32_t flags;    // bit 0 has a non-skipped byte, 1 its last one is whitespace, 2 its first one is
                     // whitespace, 8 resolved carry-in, 16..23 mapped value of the last non-skipped byte
  uint32_t _pad;
};
constexpr uint32_t kTfBlocksPerWindow = kWindowBytes / 16384;

struct TransformParams {
  const uint8_t *src;   // source bytes (device)
  uint64_t src_off;     // offset of this batch's first window inside src
  uint64_t src_len;     // source bytes of this batch (windows of 4 MiB, the last may be short)
  uint8_t *norm;        // normalised windows, `win_stride` apart, starting at norm_off
  uint64_t norm_off;
  uint64_t win_stride;
  uint32_t *map;        // kWindowBytes entries per window, or nullptr (case folding only)
  WindowDesc *windows;  // one per window of the batch
  TfBlock *blocks;      // kTfBlocksPerWindow per window of the batch (scratch)
  uint8_t *ghost;       // kWindowBytes + 1 bytes, image of the reference's scratch buffer
  uint32_t flags;       // header flags of the store
};

cudaError_t transform_launch(const TransformParams &p, uint32_t n_windows, bool need_tails, int sms,
                             cudaStream_t stream, uint32_t *launches);

} // namespace olm
