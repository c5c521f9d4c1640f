// store.h -- host view of a compiled store and the builder of its HBM layout.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "device_tables.h"
#include "olm_format.h"

namespace olm {

// Parsed (not copied) view of a mapped .olm file.  Reference reader: matcher.c:329-432.
struct StoreView {
  const uint8_t *base = nullptr;
  size_t size = 0;
  Header hdr;
  const uint8_t *patterns = nullptr; // hdr.store_bytes
  uint32_t bloom_bits = 0;
  const uint8_t *bloom = nullptr; // hdr.bloom_bytes
  const uint8_t *index = nullptr; // hdr.table_size * 4
  const uint8_t *blob = nullptr;  // hdr.blob_bytes
  const uint8_t *bitmap1 = nullptr, *bitmap2 = nullptr;
  uint32_t n1 = 0, n2 = 0, n3 = 0, n4 = 0;
  const uint8_t *arr3 = nullptr, *arr4 = nullptr;
};

// Returns an empty string on success, else what is wrong with the file.
std::string parse_store(const uint8_t *file, size_t size, StoreView *out);

// Host-side image of the device tables (device_tables.h), ready to be uploaded verbatim.
struct StagedStore {
  std::vector<uint4> keys;  // buckets of four grams
  std::vector<Slot> slots;  // 4 per bucket
  std::vector<Rec> recs;
  std::vector<uint8_t> store;
  std::vector<uint32_t> g4, p23, set3, bitmap2, sx;
  DeviceStore params; // pointer members are filled in after upload
  uint32_t n_keys = 0;
};

// Limits for the two shared-memory filters, in log2(bits).
struct FilterBudget {
#ifndef OLM_G4_MAX_LOG2
#define OLM_G4_MAX_LOG2 20
#endif
  uint32_t g4_max_log2 = OLM_G4_MAX_LOG2;  // 128 KiB
  uint32_t p23_max_log2 = 18; // 32 KiB
};

std::string stage_store(const StoreView &v, const FilterBudget &budget, StagedStore *out);

// Host image of the tables behind the exact statistics (stats.cuh): the file's Bloom bits as
// 64-bit words, its gram -> bucket map as an open-addressing table, the pattern lengths per bucket.
struct StagedStats {
  std::vector<unsigned long long> bloom;
  uint32_t bloom_mask = 0;
  std::vector<uint2> map; // {gram, 1 + index into lens} or {0, 0}
  uint32_t map_shift = 32, map_mask = 0;
  std::vector<uint32_t> lens; // per bucket: count, lengths (file order: longest first)
};
std::string stage_stats(const StoreView &v, StagedStats *out);

// Walks every pattern of the file through the staged tables exactly as the scan kernel would
// probe them; returns the number of patterns that are NOT reachable (0 = tables are sound).
// Used by CPU-only tests; it does not match haystacks.
uint64_t check_staged_store(const StoreView &v, const StagedStore &s);

// host mirror of the device hashing (scan.cu uses the same expressions)
inline uint32_t key_home(const DeviceStore &d, uint32_t key) { return key >> d.key_shift; }
inline uint32_t g4_bit(const DeviceStore &d, uint32_t key) { return key >> d.g4_shift; }
constexpr uint32_t kSx3Off = 2048; // words of sx in front of its hashed part
// bit of a 3-byte pattern (key3 = b0<<16|b1<<8|b2) in the hashed part of sx
inline uint32_t sx3_bit(const DeviceStore &d, uint32_t key3) { return (key3 * kHashMul) >> d.sx3_shift; }
inline uint32_t p23_bit(const DeviceStore &d, uint32_t gram) {
  return ((gram & d.p23_and) * d.p23_mul) >> d.p23_shift;
}
// byte-class prefilter as the kernel evaluates it (device_tables.h ByteClass)
inline bool class_has(const ByteClass &c, uint32_t b) {
  if (b >= 0x80) return false;
  const uint32_t t = b & c.and_mask;
  for (uint32_t i = 0; i < c.n_ranges; ++i)
    if (t >= c.lo[i] && t <= c.hi[i]) return true;
  return false;
}
inline uint32_t set3_home(const DeviceStore &d, uint32_t key3) { return (key3 * kHashMul >> 8) & d.set3_mask; }

} // namespace olm
