// olm_format.h -- the compiled pattern store (".olm") as this library reads and writes it.
//
// The format is the reference's (writer: omega_match/src/compiler.c:241-380, reader:
// omega_match/src/matcher.c:329-432, structs: omega_match/include/omega/details/common.h:77-213).
// Sections follow each other WITHOUT alignment padding, so nothing here is ever accessed
// through a typed pointer into the mapping: fields are memcpy'd out (rd32/rd64) and the
// sections are re-staged into aligned device tables by store.cpp.
//
//   [ 0] header, 72 bytes, packed little endian:
//        +0  char[8]  "0MGM4tCH"        +36 u32 bloom bytes
//        +8  u32 version (1)            +40 u32 bucket blob bytes
//        +12 u32 flags                  +44 u32 table_size (power of two)
//        +16 u64 pattern store bytes    +48 u32 occupied buckets
//        +24 u32 stored (long) patterns +52 u32 min bucket   +56 u32 max bucket
//        +28 u32 smallest length        +60 u32 short section bytes
//        +32 u32 largest length         +64 f32 load factor  +68 f32 avg bucket
//   [72] pattern bytes of every long (>= 5 byte) pattern, concatenated
//        "0MG8L0oM"  u32 bit_size   u64 bits[bit_size/64]           (3-probe Bloom over grams)
//        "0MG*H4sH"  u32 index[table_size]                          (offset into the blob; 0 = unused)
//        bucket blob: { u32 gram; u32 count; { u64 store_off; u32 len; u32 0 } x count } ...
//        optional "0MG5HOrT" u8 bitmap1[32] u8 bitmap2[8192] u32 n1 n2 n3 n4 u32 arr3[n3] u32 arr4[n4]
#pragma once
#include <cstdint>
#include <cstring>

namespace olm {

constexpr char kMagicHeader[9] = "0MGM4tCH";
constexpr char kMagicBloom[9] = "0MG8L0oM";
constexpr char kMagicHash[9] = "0MG*H4sH";
constexpr char kMagicShort[9] = "0MG5HOrT";
constexpr uint32_t kFormatVersion = 1;
constexpr size_t kHeaderBytes = 72;
constexpr size_t kBucketRecordBytes = 16;

// header.flags bits (common.h:22-24)
constexpr uint32_t kFlagIgnoreCase = 1u << 1;
constexpr uint32_t kFlagIgnorePunct = 1u << 2;
constexpr uint32_t kFlagElideSpace = 1u << 3;
constexpr uint32_t kFlagAnyTransform = kFlagIgnoreCase | kFlagIgnorePunct | kFlagElideSpace;

// The haystack is normalised in independent source windows of this size when the store
// carries a transform flag (matcher.c:60, :945-1010).
constexpr uint32_t kWindowBytes = 4u * 1024u * 1024u;

struct Header {
  uint32_t version = 0, flags = 0;
  uint64_t store_bytes = 0;
  uint32_t stored_patterns = 0, smallest = 0, largest = 0;
  uint32_t bloom_bytes = 0, blob_bytes = 0, table_size = 0;
  uint32_t occupied = 0, min_bucket = 0, max_bucket = 0, short_bytes = 0;
  float load_factor = 0.f, avg_bucket = 0.f;
};

inline uint32_t rd32(const uint8_t *p) {
  uint32_t v;
  std::memcpy(&v, p, 4);
  return v;
}
inline uint64_t rd64(const uint8_t *p) {
  uint64_t v;
  std::memcpy(&v, p, 8);
  return v;
}
inline void wr32(uint8_t *p, uint32_t v) { std::memcpy(p, &v, 4); }
inline void wr64(uint8_t *p, uint64_t v) { std::memcpy(p, &v, 8); }

// util.h:23-26 -- the 4-byte gram is packed big endian.
inline uint32_t gram_be(const uint8_t *p) {
  return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]);
}

// hash.h:13-25 (needed on the host only to keep files interchangeable with the reference)
inline uint32_t ref_fmix32(uint32_t g) {
  g ^= g >> 16;
  g *= 0x85ebca6bu;
  g ^= g >> 13;
  g *= 0xc2b2ae35u;
  g ^= g >> 16;
  return g;
}
inline uint32_t ref_index_hash(uint32_t x) { return (x ^ 0x9e3779b9u) * 0x01000193u; }

} // namespace olm
