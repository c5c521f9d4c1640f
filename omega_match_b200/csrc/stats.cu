// stats.cu -- see stats.cuh.  One thread per start position, grid-stride; the five counters are
// kept in registers, reduced per warp at the end and added with one atomic per warp and counter.
#include "stats.cuh"

#include "olm_classes.h"
#include "olm_format.h"

namespace olm {

namespace {

constexpr int kStatsThreads = 256;

__device__ __forceinline__ uint32_t fmix32(uint32_t g) { // hash.h:13-21
  g ^= g >> 16;
  g *= 0x85ebca6bu;
  g ^= g >> 13;
  g *= 0xc2b2ae35u;
  g ^= g >> 16;
  return g;
}

// bloom_filter_query, bloom.c:51-64
__device__ __forceinline__ bool bloom_has(const StatsTables &t, uint32_t g) {
  const uint32_t h1 = fmix32(g), h2 = g * 0x9e3779b1u;
#pragma unroll
  for (uint32_t i = 0; i < 3; ++i) {
    const uint32_t bp = (h1 + i * h2) & t.bloom_mask;
    if (!((__ldg(t.bloom + (bp >> 6)) >> (bp & 63)) & 1ull)) return false;
  }
  return true;
}

// probe_bucket, hash_table.c:91-109, as an exact map: index into lens[] of the gram's bucket or -1
__device__ __forceinline__ int64_t bucket_of(const StatsTables &t, uint32_t g) {
  for (uint32_t i = (g * kHashMul) >> t.map_shift;; i = (i + 1) & t.map_mask) {
    const uint2 e = __ldg(t.map + i);
    if (e.y == 0) return -1;
    if (e.x == g) return (int64_t)e.y - 1;
  }
}

// blockIdx.y = segment: 0 in plain/shard mode, the window of the batch in window mode
__global__ void __launch_bounds__(kStatsThreads) stats_kernel(const __grid_constant__ ScanParams P,
                                                              const __grid_constant__ StatsTables S,
                                                              unsigned long long *out) {
  // positions [first, last) of a segment of `n` bytes whose position 0 is P.buf[base]
  unsigned long long first, last, n;
  long long base;
  if (P.flags & kWindowMode) {
    const WindowDesc wd = P.windows[blockIdx.y];
    first = 0;
    last = n = wd.norm_len;
    base = (long long)(P.win_buf_off + (unsigned long long)blockIdx.y * P.win_stride);
  } else {
    first = P.scan_begin;
    n = P.seg_len;
    last = P.scan_end < n ? P.scan_end : n;
    base = P.seg_buf_off;
  }
  const bool wb = P.flags & kWordBoundary;
  uint32_t attempts = 0, filtered = 0, misses = 0, hits = 0;
  unsigned long long cmps = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * kStatsThreads;
  for (unsigned long long pos = first + (unsigned long long)blockIdx.x * kStatsThreads + threadIdx.x; pos < last;
       pos += stride) {
    const uint8_t *h = P.buf + (base + (long long)pos);
    if (wb) { // matcher.c:770-776
      const bool cw = is_word_byte(__ldg(h));
      const bool pw = pos > 0 && base + (long long)pos > 0 ? is_word_byte(__ldg(h - 1)) : false;
      if (cw == pw) continue;
    }
    const unsigned long long rem = n - pos;
    if (rem < 4) continue; // matcher.c:782
    ++attempts;
    const uint32_t g = ((uint32_t)__ldg(h) << 24) | ((uint32_t)__ldg(h + 1) << 16) | ((uint32_t)__ldg(h + 2) << 8) |
                       (uint32_t)__ldg(h + 3);
    if (!bloom_has(S, g)) {
      ++filtered;
      continue;
    }
    const int64_t b = bucket_of(S, g);
    if (b < 0) {
      ++misses;
      continue;
    }
    ++hits;
    // comparisons: patterns of the bucket with len <= remaining (matcher.c:203, :210)
    const uint32_t cnt = __ldg(S.lens + b);
    if (rem >= S.largest) {
      cmps += cnt;
    } else {
      for (uint32_t j = 0; j < cnt; ++j) cmps += __ldg(S.lens + b + 1 + j) <= rem ? 1u : 0u;
    }
  }
  unsigned long long v[5] = {attempts, filtered, misses, hits, cmps};
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v[i] += __shfl_xor_sync(0xFFFFFFFFu, v[i], d);
    if ((threadIdx.x & 31) == 0 && v[i]) atomicAdd(out + i, v[i]);
  }
}

} // namespace

cudaError_t stats_launch(const ScanParams &p, const StatsTables &t, unsigned long long *out, int sms,
                         cudaStream_t stream, uint32_t *launches) {
  if (t.largest < 5 || !t.map) return cudaSuccess; // no long patterns: core_match never enters the long path
  const bool windowed = p.flags & kWindowMode;
  const uint32_t segs = windowed ? p.num_tiles / p.tiles_per_win : 1u;
  if (segs == 0) return cudaSuccess;
  const unsigned long long per_seg = windowed ? kWindowBytes : (p.scan_end - p.scan_begin);
  unsigned long long bx = (per_seg + kStatsThreads * 8ull - 1) / (kStatsThreads * 8ull); // ~8 positions per thread
  const unsigned long long cap = windowed ? 64ull : (unsigned long long)sms * 32ull;
  if (bx > cap) bx = cap;
  if (bx == 0) bx = 1;
  stats_kernel<<<dim3((unsigned)bx, segs), kStatsThreads, 0, stream>>>(p, t, out);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

} // namespace olm
