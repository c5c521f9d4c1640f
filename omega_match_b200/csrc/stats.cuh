// stats.cuh -- exact omega_match_stats_t counters of the long path (SURVEY 8a "Stats", row N2).
//
// The scan kernel does not probe the reference's structures (3-probe Bloom, gram -> bucket map),
// so the counters that describe them -- attempts, filtered, misses and hits of the long path,
// comparisons (matcher.c:783-799, :210) -- cannot fall out of it.  When a stats struct is
// attached to the matcher (omega_list_matcher_add_stats), one extra kernel per launch group
// walks the same bytes and evaluates exactly what core_match() counts, against the file's own
// Bloom bits and a device copy of its gram -> bucket map.  Without an attached struct nothing
// of this runs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "scan.cuh"

namespace olm {

struct StatsTables {
  const unsigned long long *bloom = nullptr; // bits of the file's Bloom filter (bloom.c:51-64)
  uint32_t bloom_mask = 0;                   // bit_size - 1
  const uint2 *map = nullptr;                // open addressing: {gram, 1 + index of the bucket in lens[]} or {*, 0}
  uint32_t map_shift = 32, map_mask = 0;     // home = (gram * kHashMul) >> map_shift
  const uint32_t *lens = nullptr;            // per bucket: count, then the pattern lengths, longest first
  uint32_t largest = 0;                      // header: largest pattern length
};

// counters written (atomicAdd) by the kernel, relative to `out`
enum StatsCounter : int { kStatAttempts = 0, kStatFiltered = 1, kStatLongMisses = 2, kStatLongHits = 3, kStatComparisons = 4 };

// Same bytes, same positions and same word_boundary skip as the scan launch described by `p`.
cudaError_t stats_launch(const ScanParams &p, const StatsTables &t, unsigned long long *out, int sms,
                         cudaStream_t stream, uint32_t *launches);

} // namespace olm
