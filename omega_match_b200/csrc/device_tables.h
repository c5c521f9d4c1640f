// device_tables.h -- the pattern store as it lives in HBM (built by store.cpp, read by scan.cu).
//
// The on-disk sections are not aligned and the reference's index array does not mark unused
// slots, so nothing of the file is used in place.  At create() the store is
// re-staged into:
//
//   slots[]   one 16-byte record per distinct gram in BUCKETS of two (one 32-byte sector):
//             a gram lives in the first bucket, counted from its home bucket, that had a free
//             slot when it was inserted; a probe stops at the first bucket with a free slot.
//             Load <= 0.5, so most probes -- hit or miss -- cost one sector.  A slot holds
//             everything the scan needs to REJECT a candidate with one 16-byte load: the
//             gram, pattern bytes 4..7 and the length of the (longest) pattern.  Grams of
//             4-byte patterns live in the same table (flag bit) -> the length-4 short matcher
//             (matcher.c:685-692, a binary search) becomes the same probe.
//   recs[]    for buckets with more than one pattern: 16-byte records, longest first
//             (the order compiler.c:271 gives the bucket).
//   store[]   pattern bytes, padded so 4-byte reads never leave the allocation.
//   g4[]      single-probe hashed bitmap over every gram in slots[] -- a superset filter with
//             no false negatives, copied into shared memory by every CTA.  It stands in for
//             the reference's 3-probe Bloom (bloom.c:51-64): same role, one probe, sized to
//             fit on chip; which positions reach the exact table changes, the match set not.
//   p23[]     one hashed bitmap for the 1..3 byte patterns (exact membership is then checked
//             against bitmap1/bitmap2 of the file and a hash set of the 3-byte keys).
#pragma once
#include <cstdint>

namespace olm {

struct alignas(16) Slot {
  uint32_t key;   // big-endian gram (util.h:23-26)
  uint32_t next4; // pattern bytes 4..7 as a little-endian word, zero padded (single-pattern slots)
  uint32_t meta;  // 0 = empty; see kSlot* below
  uint32_t ref;   // single: offset of the pattern in store[]; multi: first index in recs[]
};
constexpr uint32_t kSlotShort4 = 1u << 31;   // a 4-byte pattern equals this gram
constexpr uint32_t kSlotMulti = 1u << 30;    // low bits = number of recs, else = pattern length (0: none)
constexpr uint32_t kSlotValueMask = (1u << 30) - 1;

struct alignas(16) Rec {
  uint32_t next4;
  uint32_t len;
  uint32_t store_off;
  uint32_t _pad;
};

constexpr uint32_t kHashMul = 0x9E3779B1u; // one multiply feeds both the g4 filter and the slot index

// Everything the scan kernel needs to know about the store; passed by value.
struct DeviceStore {
  const Slot *slots = nullptr;
  const Rec *recs = nullptr;
  const uint8_t *store = nullptr;
  const uint32_t *g4 = nullptr;      // g4_words 32-bit words
  const uint32_t *p23 = nullptr;     // p23_words 32-bit words
  const uint32_t *set3 = nullptr;    // open addressing, value = key3 + 1, 0 = empty
  const uint32_t *bitmap2 = nullptr; // 2048 words, bit (b0<<8|b1) as in short_matcher_t
  uint32_t bitmap1[8] = {0};         // bit b as in short_matcher_t
  uint32_t slot_shift = 32;          // home BUCKET = (gram*kHashMul) >> slot_shift; slots 2b, 2b+1
  uint32_t slot_mask = 0;            // number of buckets - 1
  uint32_t g4_shift = 32, g4_words = 0;
  uint32_t p23_and = 0, p23_mul = 1, p23_shift = 32, p23_words = 0;
  uint32_t set3_mask = 0;
  uint32_t n_long = 0, n1 = 0, n2 = 0, n3 = 0, n4 = 0;
  uint32_t smallest = 0, largest = 0;
  uint32_t flags = 0;
};

} // namespace olm
