// store.cpp -- parse a compiled store and re-stage it for HBM (kernel K1 of SURVEY 2.1).
//
// parse_store() follows the section walk of the reference loader (omega_match/src/matcher.c:
// 329-432) with explicit bounds checks.  stage_store() then builds the device layout that
// device_tables.h describes.  The index array of the file is ignored: the bucket blob is
// walked record by record, which yields every (gram -> patterns) pair without depending on
// how unused index slots were written.
#include "store.h"

#include <algorithm>

#include "host_util.h"

namespace olm {

std::string parse_store(const uint8_t *f, size_t size, StoreView *v) {
  *v = StoreView{};
  if (!f || size < kHeaderBytes) return "file shorter than the 72-byte header";
  if (std::memcmp(f, kMagicHeader, 8) != 0) return "header magic mismatch";
  Header &h = v->hdr;
  h.version = rd32(f + 8);
  h.flags = rd32(f + 12);
  h.store_bytes = rd64(f + 16);
  h.stored_patterns = rd32(f + 24);
  h.smallest = rd32(f + 28);
  h.largest = rd32(f + 32);
  h.bloom_bytes = rd32(f + 36);
  h.blob_bytes = rd32(f + 40);
  h.table_size = rd32(f + 44);
  h.occupied = rd32(f + 48);
  h.min_bucket = rd32(f + 52);
  h.max_bucket = rd32(f + 56);
  h.short_bytes = rd32(f + 60);
  std::memcpy(&h.load_factor, f + 64, 4);
  std::memcpy(&h.avg_bucket, f + 68, 4);
  v->base = f;
  v->size = size;

  uint64_t off = kHeaderBytes;
  auto need = [&](uint64_t n) { return off + n <= size; };
  if (!need(h.store_bytes)) return "pattern store runs past the end of the file";
  v->patterns = f + off;
  off += h.store_bytes;
  if (!need(12) || std::memcmp(f + off, kMagicBloom, 8) != 0) return "bloom magic mismatch";
  off += 8;
  v->bloom_bits = rd32(f + off);
  off += 4;
  if (!need(h.bloom_bytes)) return "bloom filter runs past the end of the file";
  v->bloom = f + off;
  off += h.bloom_bytes;
  if (!need(8) || std::memcmp(f + off, kMagicHash, 8) != 0) return "hash table magic mismatch";
  off += 8;
  if (!need(uint64_t(h.table_size) * 4)) return "index array runs past the end of the file";
  v->index = f + off;
  off += uint64_t(h.table_size) * 4;
  if (!need(h.blob_bytes)) return "bucket data runs past the end of the file";
  v->blob = f + off;
  off += h.blob_bytes;
  if (off + h.short_bytes != size) return "short matcher size mismatch"; // matcher.c:425
  if (h.short_bytes) {
    if (h.short_bytes < 8 + 32 + 8192 + 16 || std::memcmp(f + off, kMagicShort, 8) != 0)
      return "short matcher magic mismatch";
    const uint8_t *p = f + off + 8;
    v->bitmap1 = p;
    v->bitmap2 = p + 32;
    p += 32 + 8192;
    v->n1 = rd32(p);
    v->n2 = rd32(p + 4);
    v->n3 = rd32(p + 8);
    v->n4 = rd32(p + 12);
    p += 16;
    if (uint64_t(8 + 32 + 8192 + 16) + 4ull * v->n3 + 4ull * v->n4 != h.short_bytes)
      return "short matcher arrays do not fit their section";
    v->arr3 = p;
    v->arr4 = p + 4ull * v->n3;
  }
  return "";
}

namespace {

uint32_t ceil_log2(uint64_t v) {
  uint32_t l = 0;
  while ((1ull << l) < v) ++l;
  return l;
}

struct BucketRef {
  uint32_t gram;
  uint32_t count;
  const uint8_t *recs; // count x 16 bytes in the blob
};

} // namespace

std::string stage_store(const StoreView &v, const FilterBudget &budget, StagedStore *s) {
  *s = StagedStore{};
  const Header &h = v.hdr;
  if (h.store_bytes >= 0xFFFFFFF0ull) return "pattern store of 4 GiB or more is not supported";

  // ---- walk the bucket blob: [u32 gram][u32 count][{u64 off,u32 len,u32 0} x count]
  std::vector<BucketRef> buckets;
  buckets.reserve(h.occupied);
  uint64_t n_recs_multi = 0, n_long = 0;
  for (uint64_t p = 0; p < h.blob_bytes;) {
    if (p + 8 > h.blob_bytes) return "truncated bucket header";
    BucketRef b{rd32(v.blob + p), rd32(v.blob + p + 4), v.blob + p + 8};
    if (b.count == 0 || p + 8 + uint64_t(b.count) * kBucketRecordBytes > h.blob_bytes)
      return "bucket runs past the bucket data";
    for (uint32_t j = 0; j < b.count; ++j) {
      const uint64_t po = rd64(b.recs + 16ull * j);
      const uint32_t pl = rd32(b.recs + 16ull * j + 8);
      if (pl < 5 || pl > kSlotValueMask || po + pl > h.store_bytes) return "pattern record out of range";
      if (rd32(v.patterns + po) != __builtin_bswap32(b.gram)) return "pattern does not start with its bucket gram";
    }
    n_long += b.count;
    if (b.count > 1) n_recs_multi += b.count;
    buckets.push_back(b);
    p += 8 + uint64_t(b.count) * kBucketRecordBytes;
  }
  if (n_recs_multi > kSlotValueMask) return "too many patterns";

  // ---- pattern bytes, padded for unaligned 4-byte reads
  s->store.assign(size_t(h.store_bytes) + 16, 0);
  if (h.store_bytes) std::memcpy(s->store.data(), v.patterns, size_t(h.store_bytes));

  // ---- slots: bucket grams first, then the 4-byte patterns are merged in
  const uint64_t n_keys_upper = buckets.size() + uint64_t(v.n4);
  // buckets of two slots; capacity >= 2x the keys (4x while the table stays small)
  uint32_t lg_buckets = std::max<uint32_t>(3, ceil_log2(std::max<uint64_t>(1, n_keys_upper)));
  if (n_keys_upper <= (1u << 18)) ++lg_buckets;
  if (lg_buckets > 29) return "too many distinct grams";
  const uint32_t n_slots = 2u << lg_buckets;
  DeviceStore &d = s->params;
  d.slot_shift = 32 - lg_buckets;
  d.slot_mask = (1u << lg_buckets) - 1;
  s->slots.assign(n_slots, Slot{0, 0, 0, 0});
  s->recs.reserve(n_recs_multi);

  auto next4_of = [&](uint64_t off, uint32_t len) {
    uint32_t w = 0;
    const uint32_t m = std::min<uint32_t>(len - 4, 4);
    for (uint32_t i = 0; i < m; ++i) w |= uint32_t(s->store[off + 4 + i]) << (8 * i);
    return w;
  };
  auto find_or_claim = [&](uint32_t gram) -> Slot & {
    for (uint32_t b = slot_home(d, gram);; b = (b + 1) & d.slot_mask) {
      for (uint32_t k = 0; k < 2; ++k) {
        Slot &sl = s->slots[2 * size_t(b) + k];
        if (sl.meta == 0 || sl.key == gram) {
          sl.key = gram;
          return sl;
        }
      }
    }
  };
  uint32_t n_keys = 0;
  for (const BucketRef &b : buckets) {
    Slot &sl = find_or_claim(b.gram);
    if (sl.meta != 0) return "gram appears in two buckets";
    ++n_keys;
    if (b.count == 1) {
      const uint64_t po = rd64(b.recs);
      const uint32_t pl = rd32(b.recs + 8);
      sl.meta = pl;
      sl.ref = uint32_t(po);
      sl.next4 = next4_of(po, pl);
    } else {
      sl.meta = kSlotMulti | b.count;
      sl.ref = uint32_t(s->recs.size());
      sl.next4 = 0;
      const size_t first = s->recs.size();
      for (uint32_t j = 0; j < b.count; ++j) {
        const uint64_t po = rd64(b.recs + 16ull * j);
        const uint32_t pl = rd32(b.recs + 16ull * j + 8);
        s->recs.push_back(Rec{next4_of(po, pl), pl, uint32_t(po), 0});
      }
      // the scan emits a bucket's matches in record order and relies on "longest first"
      // (compiler.c:271 sorts the bucket that way; enforce it for foreign writers)
      std::stable_sort(s->recs.begin() + first, s->recs.end(), [](const Rec &x, const Rec &y) { return x.len > y.len; });
    }
  }
  for (uint32_t i = 0; i < v.n4; ++i) {
    const uint32_t gram = rd32(v.arr4 + 4ull * i);
    Slot &sl = find_or_claim(gram);
    if (sl.meta == 0) ++n_keys;
    sl.meta |= kSlotShort4;
  }
  s->n_keys = n_keys;
  if (s->recs.empty()) s->recs.push_back(Rec{0, 0, 0, 0}); // never dereferenced; keeps the upload non-empty

  // ---- g4: one bit per gram, >= 16 bits per key while it fits the shared-memory budget
  if (n_keys) {
    const uint32_t lg = std::min(budget.g4_max_log2, std::max<uint32_t>(10, ceil_log2(uint64_t(n_keys) * 16)));
    d.g4_shift = 32 - lg;
    d.g4_words = (1u << lg) / 32;
    s->g4.assign(d.g4_words, 0);
    for (const Slot &sl : s->slots)
      if (sl.meta) {
        const uint32_t b = g4_bit(d, sl.key);
        s->g4[b >> 5] |= 1u << (b & 31);
      }
  }

  // ---- p23: candidates for the 1..3 byte patterns
  d.n1 = v.n1;
  d.n2 = v.n2;
  d.n3 = v.n3;
  d.n4 = v.n4;
  if (v.n1 || v.n2 || v.n3) {
    const uint64_t hashed_entries = uint64_t(v.n3) + uint64_t(v.n2) * 256;
    const bool hashed = v.n1 == 0 && hashed_entries * 16 <= (1ull << budget.p23_max_log2);
    if (hashed) { // index = hash of the first three bytes; 2-byte patterns cover all third bytes
      const uint32_t lg = std::min(budget.p23_max_log2, std::max<uint32_t>(10, ceil_log2(hashed_entries * 16)));
      d.p23_and = 0xFFFFFF00u;
      d.p23_mul = kHashMul;
      d.p23_shift = 32 - lg;
      d.p23_words = (1u << lg) / 32;
    } else { // index = the first two bytes; exact for 2-byte patterns, a prefix test for the rest
      d.p23_and = 0xFFFF0000u;
      d.p23_mul = 1;
      d.p23_shift = 16;
      d.p23_words = 65536 / 32;
    }
    s->p23.assign(d.p23_words, 0);
    auto set_bit = [&](uint32_t gram) {
      const uint32_t b = p23_bit(d, gram);
      s->p23[b >> 5] |= 1u << (b & 31);
    };
    for (uint32_t i = 0; i < v.n3; ++i) set_bit(rd32(v.arr3 + 4ull * i) << 8);
    for (uint32_t w = 0; w < 65536; ++w) {
      if (!(v.bitmap2[w >> 3] & (1u << (w & 7)))) continue;
      if (hashed)
        for (uint32_t c = 0; c < 256; ++c) set_bit((w << 16) | (c << 8));
      else
        set_bit(w << 16);
    }
    if (!hashed)
      for (uint32_t b0 = 0; b0 < 256; ++b0)
        if (v.bitmap1[b0 >> 3] & (1u << (b0 & 7)))
          for (uint32_t b1 = 0; b1 < 256; ++b1) set_bit((b0 << 24) | (b1 << 16));
  }

  // ---- exact sets for the short lengths
  {
    const uint32_t lg = std::max<uint32_t>(4, ceil_log2(std::max<uint64_t>(1, uint64_t(v.n3) * 2)));
    d.set3_mask = (1u << lg) - 1;
    s->set3.assign(size_t(1) << lg, 0);
    for (uint32_t i = 0; i < v.n3; ++i) {
      const uint32_t k = rd32(v.arr3 + 4ull * i);
      uint32_t j = set3_home(d, k);
      while (s->set3[j] != 0 && s->set3[j] != k + 1) j = (j + 1) & d.set3_mask;
      s->set3[j] = k + 1;
    }
    s->bitmap2.assign(2048, 0);
    if (v.bitmap2) std::memcpy(s->bitmap2.data(), v.bitmap2, 8192);
    if (v.bitmap1) std::memcpy(d.bitmap1, v.bitmap1, 32);
  }
  d.n_long = uint32_t(n_long);
  d.smallest = h.smallest;
  d.largest = h.largest;
  d.flags = h.flags;
  return "";
}

uint64_t check_staged_store(const StoreView &v, const StagedStore &s) {
  const DeviceStore &d = s.params;
  uint64_t bad = 0;
  auto probe = [&](uint32_t gram) -> const Slot * {
    if (!s.g4.empty()) {
      const uint32_t b = g4_bit(d, gram);
      if (!(s.g4[b >> 5] >> (b & 31) & 1)) return nullptr;
    }
    for (uint32_t b = slot_home(d, gram);; b = (b + 1) & d.slot_mask) {
      const Slot &x = s.slots[2 * size_t(b)], &y = s.slots[2 * size_t(b) + 1];
      if (x.meta != 0 && x.key == gram) return &x;
      if (y.meta != 0 && y.key == gram) return &y;
      if (x.meta == 0 || y.meta == 0) return nullptr;
    }
  };
  for (uint64_t p = 0; p < v.hdr.blob_bytes;) {
    const uint32_t gram = rd32(v.blob + p), count = rd32(v.blob + p + 4);
    const Slot *sl = probe(gram);
    for (uint32_t j = 0; j < count; ++j) {
      const uint64_t po = rd64(v.blob + p + 8 + 16ull * j);
      const uint32_t pl = rd32(v.blob + p + 8 + 16ull * j + 8);
      bool ok = false;
      if (sl && !(sl->meta & kSlotMulti))
        ok = count == 1 && (sl->meta & kSlotValueMask) == pl && sl->ref == po;
      uint32_t got = sl ? sl->next4 : 0;
      if (sl && (sl->meta & kSlotMulti) && (sl->meta & kSlotValueMask) == count) {
        for (uint32_t q = 0; q < count && !ok; ++q) {
          const Rec &rc = s.recs[sl->ref + q];
          if (rc.len == pl && rc.store_off == po) {
            ok = (q == 0 || s.recs[sl->ref + q - 1].len >= rc.len);
            got = rc.next4;
          }
        }
      }
      if (ok) {
        uint32_t w = 0;
        for (uint32_t i = 0; i < std::min<uint32_t>(pl - 4, 4); ++i) w |= uint32_t(v.patterns[po + 4 + i]) << (8 * i);
        ok = w == got && std::memcmp(s.store.data() + po, v.patterns + po, pl) == 0;
      }
      bad += !ok;
    }
    p += 8 + 16ull * count;
  }
  for (uint32_t i = 0; i < v.n4; ++i) {
    const Slot *sl = probe(rd32(v.arr4 + 4ull * i));
    bad += !(sl && (sl->meta & kSlotShort4));
  }
  auto p23_ok = [&](uint32_t gram) {
    const uint32_t b = p23_bit(d, gram);
    return !s.p23.empty() && (s.p23[b >> 5] >> (b & 31) & 1);
  };
  for (uint32_t i = 0; i < v.n3; ++i) {
    const uint32_t k = rd32(v.arr3 + 4ull * i);
    bool found = false;
    for (uint32_t j = set3_home(d, k); s.set3[j] != 0; j = (j + 1) & d.set3_mask)
      if (s.set3[j] == k + 1) {
        found = true;
        break;
      }
    for (uint32_t c = 0; c < 256; c += 51) bad += !p23_ok((k << 8) | c);
    bad += !found;
  }
  for (uint32_t w = 0; w < 65536 && v.bitmap2; ++w)
    if (v.bitmap2[w >> 3] & (1u << (w & 7))) {
      bad += !(s.bitmap2[w >> 5] >> (w & 31) & 1);
      for (uint32_t c = 0; c < 65536; c += 4099) bad += !p23_ok((w << 16) | c);
    }
  for (uint32_t b = 0; b < 256 && v.bitmap1; ++b)
    if (v.bitmap1[b >> 3] & (1u << (b & 7))) {
      bad += !(d.bitmap1[b >> 5] >> (b & 31) & 1);
      for (uint32_t c = 0; c < (1u << 24); c += 65521) bad += !p23_ok((b << 24) | c);
    }
  return bad;
}

} // namespace olm
