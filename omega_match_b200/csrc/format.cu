// format.cu -- the result listing of the reference's CLI, produced on the GPU.
//
// `olm match` prints one line per match, "offset:matched bytes\n", with
// snprintf("%zu:%.*s\n", offset, len, match) (omega_match/main.c:89-133) -- so the bytes of a line
// stop at the first NUL inside the match.  With the haystack and the records already in HBM the
// listing is three small kernels: length of every line, exclusive prefix over the lengths (block
// sums, one block over the sums, down-sweep), and the lines written at their offsets.  The caller
// copies one contiguous text buffer to the host instead of walking the records.
#include "format.cuh"

namespace olm {

namespace {

constexpr int kFmtThreads = 256;
constexpr uint32_t kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t dec_digits(unsigned long long v) {
  uint32_t n = 1;
  while (v >= 10ull) {
    v /= 10ull;
    ++n;
  }
  return n;
}
// bytes of the match that "%.*s" prints: up to the first NUL
__device__ __forceinline__ uint32_t printed_len(const uint8_t *p, uint32_t len) {
  uint32_t n = 0;
  while (n < len && p[n] != 0) ++n;
  return n;
}

// line_len[i] and the sum of every block of kFmtThreads lines
__global__ void __launch_bounds__(kFmtThreads) format_len_kernel(const Record *rec, uint64_t n, const uint8_t *hay,
                                                                 uint64_t hay_off0, uint32_t *line_len,
                                                                 unsigned long long *block_sum) {
  __shared__ unsigned long long s_warp[kFmtThreads / 32];
  const uint64_t i = (uint64_t)blockIdx.x * kFmtThreads + threadIdx.x;
  uint32_t l = 0;
  if (i < n) {
    const Record r = rec[i];
    l = dec_digits(r.offset) + 1u + printed_len(hay + (r.offset - hay_off0), r.len) + 1u;
    line_len[i] = l;
  }
  unsigned long long s = l;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(kFull, s, d);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kFmtThreads / 32; ++w) t += s_warp[w];
    block_sum[blockIdx.x] = t;
  }
}

// block_sum[b] <- sum of the blocks before b; *total <- everything (one CTA)
__global__ void __launch_bounds__(1024, 1) format_scan_kernel(unsigned long long *block_sum, uint64_t n_blocks,
                                                              unsigned long long *total) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint64_t b0 = 0; b0 < n_blocks; b0 += 1024) {
    const unsigned long long mine = b0 + tid < n_blocks ? block_sum[b0 + tid] : 0ull;
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long t = __shfl_up_sync(kFull, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long before = s_carry;
    for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
    if (b0 + tid < n_blocks) block_sum[b0 + tid] = before + incl - mine;
    __syncthreads();
    if (tid == 1023) s_carry = before + incl;
    __syncthreads();
  }
  if (tid == 0) *total = s_carry;
}

__global__ void __launch_bounds__(kFmtThreads) format_write_kernel(const Record *rec, uint64_t n, const uint8_t *hay,
                                                                   uint64_t hay_off0, const uint32_t *line_len,
                                                                   const unsigned long long *block_base, uint8_t *text,
                                                                   uint64_t text_cap) {
  __shared__ unsigned long long s_warp[kFmtThreads / 32];
  const uint64_t i = (uint64_t)blockIdx.x * kFmtThreads + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t l = i < n ? line_len[i] : 0u;
  unsigned long long incl = l;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(kFull, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned long long at = block_base[blockIdx.x] + incl - l;
  for (uint32_t w = 0; w < warp; ++w) at += s_warp[w];
  if (i >= n || at + l > text_cap) return;
  const Record r = rec[i];
  uint8_t *o = text + at;
  const uint32_t nd = dec_digits(r.offset);
  unsigned long long v = r.offset;
  for (uint32_t k = nd; k-- > 0;) {
    o[k] = (uint8_t)('0' + v % 10ull);
    v /= 10ull;
  }
  o[nd] = ':';
  const uint8_t *p = hay + (r.offset - hay_off0);
  const uint32_t body = l - nd - 2u;
  for (uint32_t k = 0; k < body; ++k) o[nd + 1 + k] = p[k];
  o[l - 1] = '\n';
}

} // namespace

size_t format_scratch_bytes(uint64_t n) {
  const uint64_t blocks = (n + kFmtThreads - 1) / kFmtThreads;
  return size_t(n) * 4 + size_t(blocks + 2) * 8 + 64;
}

cudaError_t format_lengths_launch(const Record *rec, uint64_t n, const uint8_t *hay, uint64_t hay_off0, void *scratch,
                                  unsigned long long *d_total, cudaStream_t st, uint32_t *launches) {
  const uint64_t blocks = (n + kFmtThreads - 1) / kFmtThreads;
  uint32_t *line_len = static_cast<uint32_t *>(scratch);
  unsigned long long *block_sum = reinterpret_cast<unsigned long long *>(static_cast<uint8_t *>(scratch) + ((size_t(n) * 4 + 15) & ~size_t(15)));
  format_len_kernel<<<(unsigned)blocks, kFmtThreads, 0, st>>>(rec, n, hay, hay_off0, line_len, block_sum);
  format_scan_kernel<<<1, 1024, 0, st>>>(block_sum, blocks, d_total);
  *launches += 2;
  return cudaGetLastError();
}

cudaError_t format_write_launch(const Record *rec, uint64_t n, const uint8_t *hay, uint64_t hay_off0, void *scratch,
                                uint8_t *text, uint64_t text_cap, cudaStream_t st, uint32_t *launches) {
  const uint64_t blocks = (n + kFmtThreads - 1) / kFmtThreads;
  const uint32_t *line_len = static_cast<const uint32_t *>(scratch);
  const unsigned long long *block_sum =
      reinterpret_cast<const unsigned long long *>(static_cast<const uint8_t *>(scratch) + ((size_t(n) * 4 + 15) & ~size_t(15)));
  format_write_kernel<<<(unsigned)blocks, kFmtThreads, 0, st>>>(rec, n, hay, hay_off0, line_len, block_sum, text, text_cap);
  *launches += 1;
  return cudaGetLastError();
}

} // namespace olm
