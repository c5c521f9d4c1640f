// transform.cuh -- launch interface of the normalisation kernels (transform.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "olm_format.h"
#include "scan.cuh"

namespace olm {

// Summary of one 4 KiB block of a window (transform.cu count / resolve passes).
struct alignas(16) TfBlock {
  uint32_t count;    // kept bytes of the block assuming no whitespace run is carried in
  uint32_t out_base; // resolved: normalised index of the block's first kept byte
  uint32_t flags;    // bit 0 has a non-skipped byte, 1 its last one is whitespace, 2 its first one is
                     // whitespace, 8 resolved carry-in, 16..23 mapped value of the last non-skipped byte
  uint32_t _pad;
};
constexpr uint32_t kTfBlocksPerWindow = kWindowBytes / 4096;

struct TransformParams {
  const uint8_t *src;   // source bytes (device)
  uint64_t src_off;     // offset of the launch's first window inside src
  uint64_t src_len;     // source bytes of the launch (windows of 4 MiB, the last may be short)
  WindowDesc *windows;  // one per window of the launch
  TfBlock *blocks;      // kTfBlocksPerWindow per window (scratch; stores that drop bytes and have 2..4 byte patterns)
  uint2 *visible;       // one per window (scratch, same stores)
  uint4 *plan;          // one per window (scratch, same stores): where a window's stale-tail byte is found
  uint8_t *ghost;       // kWindowBytes + 1 bytes, image of the reference's scratch buffer
  uint32_t flags;       // header flags of the store
};

// Window descriptors of a launch (see transform.cu for who needs what).  need_tails: the store has
// 2..4 byte patterns -- tails are resolved and the ghost image is brought up to date.
cudaError_t window_descs_launch(const TransformParams &p, uint32_t n_windows, bool need_tails, int sms,
                                cudaStream_t stream, uint32_t *launches);
// Stores that drop bytes and have 2..4 byte patterns, launches that did NOT need the descriptors:
// the ghost image is brought up to date AFTER the scan, from the window extents the scan counted
// (extents[w] = kept bytes of window w); only the windows that stay visible are normalised again.
cudaError_t ghost_update_launch(const TransformParams &p, uint32_t n_windows, const uint32_t *extents, int sms,
                                cudaStream_t stream, uint32_t *launches);

} // namespace olm
