// transform.cuh -- launch interface of the normalisation kernels (transform.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "olm_format.h"
#include "scan.cuh"

namespace olm {

struct TransformParams {
  const uint8_t *src;   // source bytes (device)
  uint64_t src_off;     // offset of this batch's first window inside src
  uint64_t src_len;     // source bytes of this batch (windows of 4 MiB, the last may be short)
  uint8_t *norm;        // normalised windows, `win_stride` apart, starting at norm_off
  uint64_t norm_off;
  uint64_t win_stride;
  uint32_t *map;        // kWindowBytes entries per window, or nullptr (case folding only)
  WindowDesc *windows;  // one per window of the batch
  uint8_t *ghost;       // kWindowBytes + 1 bytes, image of the reference's scratch buffer
  uint32_t flags;       // header flags of the store
};

cudaError_t transform_launch(const TransformParams &p, uint32_t n_windows, bool need_tails, int sms,
                             cudaStream_t stream, uint32_t *launches);

} // namespace olm
