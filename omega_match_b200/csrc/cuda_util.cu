// cuda_util.cu -- small CUDA-facing entry points of the C ABI that do not belong to a matcher.
#include <cuda_runtime.h>

#include "../../include/olm_b200.h"

extern "C" {

int olm_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void *olm_cuda_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) return nullptr;
  return p;
}

void olm_cuda_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

} // extern "C"
