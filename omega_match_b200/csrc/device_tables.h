// device_tables.h -- the pattern store as it lives in HBM (built by store.cpp, read by scan.cu).
//
// The on-disk sections are not aligned and the reference's index array does not mark unused
// slots, so nothing of the file is used in place.  At create() the store is
// re-staged into:
//
//   keys[]    the key set: BUCKETS of four 32-bit keys (16 bytes, one vector load).  A key is
//             key_hash() of the first K = key_bytes bytes of a pattern, K = 4 when the store has
//             4-byte patterns, else min(shortest pattern, 8): at 1 M random patterns a third of
//             the candidates share the first FOUR bytes with some pattern but almost none the
//             first six, so nearly every probe ends in this table.  (Two different prefixes may
//             hash to the same key -- ~n^2 / 2^33 pairs -- they then share a slot and are told
//             apart by the byte compare.)  A key lives in the first bucket, counted from its
//             home bucket, that had a free place when it was inserted; places fill left to
//             right, so a bucket whose last place holds `empty_key` ends a probe.  Load <= 0.125:
//             a probe -- hit or miss -- costs one 16-byte load in 99.8% of the cases (role of
//             probe_bucket, hash_table.c:91-109).  `empty_key` is a value that is no key.
//   slots[]   parallel to keys[] (slot = 4*bucket + place), read only after a key hit: pattern
//             bytes 0..7 and the length of the (single) pattern, so that a pattern of up to 8
//             bytes is verified without touching the pattern store, or a reference to
//   recs[]    for keys shared by several patterns: 16-byte records, longest first (the order
//             compiler.c:271 gives the bucket).  Grams of 4-byte patterns live in the same
//             table (flag bit, K = 4) -> the length-4 short matcher (matcher.c:685-692, a binary
//             search) becomes the same probe.
//   store[]   pattern bytes, padded so 4-byte reads never leave the allocation.
//   g4[]      single-probe hashed bitmap over every gram in keys[] -- a superset filter with
//             no false negatives, copied into shared memory by every CTA.  It stands in for
//             the reference's 3-probe Bloom (bloom.c:51-64): same role, one probe, sized to
//             fit on chip; which positions reach the exact table changes, the match set not.
//   p23[]     one hashed bitmap for the 1..3 byte patterns (exact membership is then checked
//             against bitmap1/bitmap2 of the file and a hash set of the 3-byte keys).
//   cls       optional byte-class prefilter for stores without 1..3 byte patterns: every
//             pattern starts with `cls_run` bytes that all fall into at most two byte ranges
//             (after an optional fold of bit 5) -- e.g. letters only.  A position whose next
//             cls_run bytes are not all in the class cannot start a match and never reaches
//             the filters.  Evaluated with SWAR arithmetic on the registers that hold the
//             haystack words, no table.
#pragma once
#include <cstdint>
#include <vector_types.h>

namespace olm {

struct alignas(16) Slot {
  uint32_t w0;   // single-pattern slots: pattern bytes 0..3 as a little-endian word; multi: bit (b & 31) for every byte b that follows the key in one of the patterns
  uint32_t w1;   // single: pattern bytes 4..7, zero padded; multi: the same for the byte after that (all ones: no constraint)
  uint32_t meta; // 0 = empty; see kSlot* below
  uint32_t ref;  // single: offset of the pattern in store[]; multi: first index in recs[]
};
constexpr uint32_t kSlotShort4 = 1u << 31;   // a 4-byte pattern equals this gram
constexpr uint32_t kSlotMulti = 1u << 30;    // low bits = number of recs, else = pattern length (0: none)
constexpr uint32_t kSlotValueMask = (1u << 30) - 1;

struct alignas(16) Rec {
  uint32_t w0; // bytes 0..3
  uint32_t len;
  uint32_t store_off;
  uint32_t w1; // bytes 4..7, zero padded
};

constexpr uint32_t kHashMul = 0x9E3779B1u; // one hash feeds both the g4 filter and the bucket index
constexpr uint32_t kHashMul2 = 0x85EBCA6Bu;

// Key of a pattern prefix / haystack position: `gram` = bytes 0..3 packed big-endian
// (util.h:23-26), `tail` = bytes 4..7 as a little-endian word, masked to the key_bytes - 4
// bytes that belong to the key (0 for 4-byte keys: then the key is a bijection of the gram).
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t key_hash(uint32_t gram, uint32_t tail) { return (gram * kHashMul) ^ (tail * kHashMul2); }

// Byte-class prefilter: byte b is in the class iff, with t = b & and_mask (and_mask clears bit 7
// and optionally bit 5), lo[i] <= t <= hi[i] for one of the n_ranges ranges, and b < 0x80.
struct ByteClass {
  uint32_t run = 0;      // 0 = disabled; else 4, 5, 6 or 8: that many leading pattern bytes are in the class
  uint32_t and_mask = 0; // 0x7f or 0x5f
  uint32_t n_ranges = 0; // 1 or 2
  uint32_t lo[2] = {0, 0}, hi[2] = {0, 0};
  // the same, replicated into the four bytes of a word for the kernel's SWAR test (store.cpp):
  // and4 = and_mask * 0x01010101, addlo[i] = (0x80 - lo[i]) * 0x01010101, addhi[i] = (0x7f - hi[i]) * 0x01010101
  uint32_t and4 = 0, addlo[2] = {0, 0}, addhi[2] = {0, 0};
};

// Everything the scan kernel needs to know about the store; passed by value.
struct DeviceStore {
  const uint4 *keys = nullptr;
  const Slot *slots = nullptr;
  const Rec *recs = nullptr;
  const uint8_t *store = nullptr;
  const uint32_t *g4 = nullptr;      // g4_words 32-bit words
  const uint32_t *p23 = nullptr;     // p23_words 32-bit words
  const uint32_t *set3 = nullptr;    // open addressing, value = key3 + 1, 0 = empty
  const uint32_t *bitmap2 = nullptr; // 2048 words, bit (b0<<8|b1) as in short_matcher_t
  uint32_t bitmap1[8] = {0};         // bit b as in short_matcher_t
  uint32_t key_shift = 32;           // home bucket = key >> key_shift
  uint32_t key_mask = 0;             // number of buckets - 1
  uint32_t empty_key = 0;
  uint32_t key_bytes = 4;            // K: pattern bytes a key covers (4..8)
  uint32_t tail_mask = 0;            // the K - 4 low bytes of the little-endian word of bytes 4..7
  uint32_t g4_shift = 32, g4_words = 0;
  uint32_t p23_and = 0, p23_mul = 1, p23_shift = 32, p23_words = 0;
  // sx: what a p23 candidate is checked against before it costs a verify (shared memory, scan.cu):
  // words [0, 2048) bit (b0<<8|b1) = "b0 b1 is a 2-byte pattern or b0 is a 1-byte pattern" (exact),
  // [2048, 2048 + 2^(32-sx3_shift)/32) a hashed bitmap of the 3-byte patterns (>= 64 bits per
  // pattern; absent when there are none).  A superset test like p23, but close to exact: p23 only says
  // "some short pattern starts with these two bytes / hashes like these three".  sx_words == 0: not used.
  const uint32_t *sx = nullptr;
  uint32_t sx_words = 0, sx3_shift = 32;
  uint32_t set3_mask = 0;
  uint32_t n_long = 0, n1 = 0, n2 = 0, n3 = 0, n4 = 0;
  uint32_t smallest = 0, largest = 0;
  uint32_t flags = 0;
  uint32_t max_recs = 0; // most patterns behind one key (a slot's record count)
  uint32_t coop = 0;     // 1: the scan compares the records of a key as a warp (store.cpp decides)
  ByteClass cls;
};
#ifndef OLM_COOP_MIN
#define OLM_COOP_MIN 4
#endif
constexpr uint32_t kCoopMinRecs = OLM_COOP_MIN; // from this many on the scan compares a key's records as a warp (scan.cu)

} // namespace olm
