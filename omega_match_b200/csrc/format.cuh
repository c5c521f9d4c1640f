// format.cuh -- launch interface of the result listing kernels (format.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "scan.cuh"

namespace olm {

// scratch for n records: the line lengths and the block sums of their prefix
size_t format_scratch_bytes(uint64_t n);
// lengths of the n lines "offset:bytes\n" (main.c:89-133) and their sum -> *d_total (device);
// `hay` is the device address of the haystack byte with offset hay_off0
cudaError_t format_lengths_launch(const Record *rec, uint64_t n, const uint8_t *hay, uint64_t hay_off0, void *scratch,
                                  unsigned long long *d_total, cudaStream_t st, uint32_t *launches);
// the lines themselves, after format_lengths_launch with the same scratch
cudaError_t format_write_launch(const Record *rec, uint64_t n, const uint8_t *hay, uint64_t hay_off0, void *scratch,
                                uint8_t *text, uint64_t text_cap, cudaStream_t st, uint32_t *launches);

} // namespace olm
