// stats.cu -- see stats.cuh.  One thread per start position, grid-stride; the five counters are
// kept in registers, reduced per warp at the end and added with one atomic per warp and counter.
#include "stats.cuh"

#include "olm_classes.h"
#include "olm_format.h"
#include "scan_device.cuh"

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

// what core_match() counts for one position of the long path (matcher.c:782-799, :203, :210)
__device__ __forceinline__ void count_position(const StatsTables &S, uint32_t g, unsigned long long rem, uint32_t &attempts,
                                               uint32_t &filtered, uint32_t &misses, uint32_t &hits, unsigned long long &cmps) {
  ++attempts;
  if (!bloom_has(S, g)) {
    ++filtered;
    return;
  }
  const int64_t b = bucket_of(S, g);
  if (b < 0) {
    ++misses;
    return;
  }
  ++hits;
  const uint32_t cnt = __ldg(S.lens + b);
  if (rem >= S.largest) {
    cmps += cnt;
  } else {
    for (uint32_t j = 0; j < cnt; ++j) cmps += __ldg(S.lens + b + 1 + j) <= rem ? 1u : 0u;
  }
}

__device__ __forceinline__ void flush_counters(uint32_t attempts, uint32_t filtered, uint32_t misses, uint32_t hits,
                                               unsigned long long cmps, unsigned long long *out) {
  unsigned long long v[5] = {attempts, filtered, misses, hits, cmps};
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v[i] += __shfl_xor_sync(0xFFFFFFFFu, v[i], d);
    if ((threadIdx.x & 31) == 0 && v[i]) atomicAdd(out + i, v[i]);
  }
}

// Stores with a transform flag: the positions are those of the normalised windows, which exist only
// chunk by chunk -- every warp builds the chunks it takes exactly like the scan does
// (scan_device.cuh) and walks their positions.
constexpr int kStatsWarps = kStatsThreads / 32;
constexpr int kStatsStage = kTilePre + kPrivData; // the chunk's source bytes + 16 in front
__global__ void __launch_bounds__(kStatsThreads) stats_window_kernel(const __grid_constant__ ScanParams P,
                                                                     const __grid_constant__ StatsTables S,
                                                                     unsigned long long *out) {
  using namespace dev;
  __shared__ __align__(16) uint8_t s_stage[kStatsWarps][kStatsStage];
  __shared__ __align__(16) uint8_t s_priv[kStatsWarps][kPrivBytes];
  __shared__ __align__(16) uint8_t s_xf[kStatsWarps][kXfRowBytes];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool wb = P.flags & kWordBoundary;
  const uint32_t xf32 = (P.flags & kIdentityMap) ? 0u : smem_u32(s_xf[warp]);
  const uint32_t priv32 = smem_u32(s_priv[warp]);
  uint8_t *buf = s_stage[warp];
  uint32_t attempts = 0, filtered = 0, misses = 0, hits = 0;
  unsigned long long cmps = 0;
  const uint64_t n_chunks = (uint64_t)P.num_tiles * kTileChunks;
  for (uint64_t ch = (uint64_t)blockIdx.x * kStatsWarps + warp; ch < n_chunks; ch += (uint64_t)gridDim.x * kStatsWarps) {
    StageInfo I;
    fill_tile(P, (uint32_t)(ch / kTileChunks), I);
    const uint32_t cbase = (uint32_t)(ch % kTileChunks) * kChunkBytes;
    if (cbase >= I.nscan) continue;
    {
      const long long first = I.boff + cbase - kTilePre;
      uint32_t have = I.staged - cbase;
      if (have > (uint32_t)kPrivData) have = kPrivData;
      for (uint32_t i = lane; i < have + kTilePre; i += 32) buf[i] = first + (long long)i >= 0 ? P.buf[first + (long long)i] : 0;
    }
    __syncwarp();
    TileCtx T;
    if (xf32) {
      // (bytes in front of the chunk that are in `buf`; what lies further back is read from global memory)
      const uint32_t back = I.boff + (long long)cbase >= (long long)kTilePre ? (uint32_t)kTilePre : 0u;
      build_xf(P, I, smem_u32(buf) + kTilePre, cbase, back, priv32, xf32, lane, false, T);
    } else {
      build_copy<true>(I, smem_u32(buf) + kTilePre, cbase, priv32, lane, T);
    }
    __syncwarp();
    for (uint32_t p = lane; p < T.nscan; p += 32) {
      const uint32_t q = priv32 + kTilePre + p;
      if (wb) { // matcher.c:770-776
        const bool cw = is_word_byte(lds8(q));
        const bool pw = (p > 0 || !T.first) ? is_word_byte(lds8(q - 1)) : false;
        if (cw == pw) continue;
      }
      // bytes left in the normalised window from p on; a chunk that does not reach the window's end
      // knows at least 8 of them, the rest only matters below `largest`
      unsigned long long rem = T.rem0 - p;
      if (T.rem0 == kRemUnknown) {
        rem = T.staged - p;
        while (rem < S.largest && xf_walk(P.buf + T.boff, P.store_flags, xf32, p + (uint32_t)rem, false).byte != kBeyond) ++rem;
        if (rem >= S.largest) rem = S.largest;
      }
      if (rem < 4) continue; // matcher.c:782
      const uint32_t g = __byte_perm(lds_le32(q), 0, 0x0123);
      count_position(S, g, rem, attempts, filtered, misses, hits, cmps);
    }
    __syncwarp();
  }
  flush_counters(attempts, filtered, misses, hits, cmps, out);
}

// plain stores: one thread per start position of [scan_begin, scan_end)
__global__ void __launch_bounds__(kStatsThreads) stats_kernel(const __grid_constant__ ScanParams P,
                                                              const __grid_constant__ StatsTables S,
                                                              unsigned long long *out) {
  // positions [first, last) of a segment of `n` bytes whose position 0 is P.buf[base]
  const unsigned long long first = P.scan_begin, n = P.seg_len;
  const unsigned long long last = P.scan_end < n ? P.scan_end : n;
  const long long base = P.seg_buf_off;
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
    const uint32_t g = ((uint32_t)__ldg(h) << 24) | ((uint32_t)__ldg(h + 1) << 16) | ((uint32_t)__ldg(h + 2) << 8) |
                       (uint32_t)__ldg(h + 3);
    count_position(S, g, rem, attempts, filtered, misses, hits, cmps);
  }
  flush_counters(attempts, filtered, misses, hits, cmps, out);
}

} // namespace

cudaError_t stats_launch(const ScanParams &p, const StatsTables &t, unsigned long long *out, int sms,
                         cudaStream_t stream, uint32_t *launches) {
  if (t.largest < 5 || !t.map) return cudaSuccess; // no long patterns: core_match never enters the long path
  if (p.flags & kWindowMode) {
    if (p.num_tiles == 0) return cudaSuccess;
    const uint64_t n_chunks = (uint64_t)p.num_tiles * kTileChunks;
    uint64_t bx = (n_chunks + kStatsWarps - 1) / kStatsWarps;
    if (bx > (uint64_t)sms * 16) bx = (uint64_t)sms * 16;
    stats_window_kernel<<<(unsigned)bx, kStatsThreads, 0, stream>>>(p, t, out);
  } else {
    const unsigned long long n = p.scan_end - p.scan_begin;
    if (n == 0) return cudaSuccess;
    unsigned long long bx = (n + kStatsThreads * 8ull - 1) / (kStatsThreads * 8ull); // ~8 positions per thread
    if (bx > (unsigned long long)sms * 32ull) bx = (unsigned long long)sms * 32ull;
    stats_kernel<<<(unsigned)bx, kStatsThreads, 0, stream>>>(p, t, out);
  }
  if (launches) *launches += 1;
  return cudaGetLastError();
}

} // namespace olm
