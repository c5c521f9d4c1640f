// compiler.cpp -- pattern compiler: builds the ".olm" store on the host.
//
// SURVEY 8f/N1.  Behavioural contract = the reference compiler
// (omega_match/src/compiler.c, pattern_store_append.c, dedupe_set.c, hash_table.c:112-189,
// bloom.c:12-49): same normalisation, same routing (<= 4 bytes -> short matcher, else
// store + gram bucket), same statistics (including their quirks), same file layout, so a
// store written here loads in the reference and vice versa.  The implementation is not the
// reference's: patterns are collected in memory and the sections are produced in one pass
// at destroy() (the reference streams pattern bytes to the file and keeps a robin-hood
// table of realloc'ed pattern arrays).
#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/olm_b200.h"
#include "host_util.h"
#include "olm_classes.h"
#include "olm_format.h"

namespace {

struct LongPattern {
  uint32_t gram;
  uint32_t len;
  uint64_t off; // into `store`
};

// Open addressing set of byte strings that live in `store` (FNV-1a, as hash.h:28-37; any
// hash would do, the set is exact).
class StoredStringSet {
public:
  explicit StoredStringSet(const std::vector<uint8_t> *store) : store_(store) { slots_.assign(1u << 13, 0); }
  // `id`s are 1-based indices into `pats`; returns false if an equal string is present.
  bool insert(const std::vector<LongPattern> &pats, const uint8_t *p, uint32_t len, uint32_t new_id) {
    if ((used_ + 1) * 10 > slots_.size() * 7) grow(pats);
    size_t h = hash(p, len) & (slots_.size() - 1);
    while (slots_[h]) {
      const LongPattern &q = pats[slots_[h] - 1];
      if (q.len == len && std::memcmp(store_->data() + q.off, p, len) == 0) return false;
      h = (h + 1) & (slots_.size() - 1);
    }
    slots_[h] = new_id;
    ++used_;
    return true;
  }

private:
  static uint32_t hash(const uint8_t *p, uint32_t len) {
    uint32_t h = 2166136261u;
    for (uint32_t i = 0; i < len; ++i) h = (h ^ p[i]) * 16777619u;
    return h;
  }
  void grow(const std::vector<LongPattern> &pats) {
    std::vector<uint32_t> old;
    old.swap(slots_);
    slots_.assign(old.size() * 2, 0);
    for (uint32_t id : old) {
      if (!id) continue;
      const LongPattern &q = pats[id - 1];
      size_t h = hash(store_->data() + q.off, q.len) & (slots_.size() - 1);
      while (slots_[h]) h = (h + 1) & (slots_.size() - 1);
      slots_[h] = id;
    }
  }
  const std::vector<uint8_t> *store_;
  std::vector<uint32_t> slots_;
  size_t used_ = 0;
};

class U32Set {
public:
  U32Set() { slots_.assign(1u << 10, 0); occ_.assign(1u << 10, 0); }
  bool insert(uint32_t v) {
    if ((used_ + 1) * 10 > slots_.size() * 7) grow();
    size_t h = (v * 0x9E3779B1u) & (slots_.size() - 1);
    while (occ_[h]) {
      if (slots_[h] == v) return false;
      h = (h + 1) & (slots_.size() - 1);
    }
    occ_[h] = 1;
    slots_[h] = v;
    ++used_;
    return true;
  }
  size_t size() const { return used_; }

private:
  void grow() {
    std::vector<uint32_t> os;
    std::vector<uint8_t> oo;
    os.swap(slots_);
    oo.swap(occ_);
    slots_.assign(os.size() * 2, 0);
    occ_.assign(os.size() * 2, 0);
    used_ = 0;
    for (size_t i = 0; i < os.size(); ++i)
      if (oo[i]) insert(os[i]);
  }
  std::vector<uint32_t> slots_;
  std::vector<uint8_t> occ_;
  size_t used_ = 0;
};

} // namespace

struct omega_list_matcher_compiler_struct {
  std::string path;
  FILE *fp = nullptr;
  bool ci = false, ip = false, ew = false;
  omega_match_pattern_store_stats_t stats{};
  std::vector<uint8_t> store;
  std::vector<LongPattern> longs;
  StoredStringSet long_set{&store};
  // short matcher (common.h:204-213)
  uint8_t bitmap1[32] = {0};
  uint8_t bitmap2[8192] = {0};
  uint32_t n1 = 0, n2 = 0;
  std::vector<uint32_t> arr3, arr4;
  U32Set set3, set4;
  // the reference's table growth: starts at 8192 slots, doubles before an insert whenever
  // (used+1)/size > 0.9 (hash_table.c:14-17, :112-116).  Tracked so table_size -- and with
  // it the Bloom size, table_size*16 bits (compiler.c:18, :257) -- equals the reference's.
  uint32_t table_size = 8192, table_used = 0;
  U32Set gram_set;
  std::vector<uint8_t> scratch;
};

using Compiler = omega_list_matcher_compiler_struct;

namespace olm {

// transform_apply() semantics (transform_table.c:36-88) for one buffer: drop skipped bytes,
// collapse whitespace runs (looking through skipped bytes) to one ' ', fold case, and drop a
// single trailing ' '.
uint32_t normalize_bytes(bool ci, bool ip, bool ew, const uint8_t *src, uint32_t len, uint8_t *out) {
  uint32_t j = 0;
  bool in_space = false;
  for (uint32_t i = 0; i < len; ++i) {
    uint32_t m;
    switch (classify_byte(src[i], ci, ip, ew, &m)) {
    case kSkip:
      break;
    case kSpace:
      if (!in_space) out[j++] = ' ';
      in_space = true;
      break;
    default:
      out[j++] = uint8_t(m);
      in_space = false;
    }
  }
  if (j > 0 && out[j - 1] == ' ') --j;
  return j;
}

} // namespace olm

extern "C" {

omega_list_matcher_compiler_t *omega_list_matcher_compiler_create(const char *compiled_file,
                                                                  int case_insensitive,
                                                                  int ignore_punctuation,
                                                                  int elide_whitespace) {
  if (!compiled_file) return nullptr;
  FILE *fp = std::fopen(compiled_file, "wb");
  if (!fp) {
    std::perror("omega_list_matcher_compiler_create: fopen");
    return nullptr;
  }
  auto *c = new Compiler();
  c->path = compiled_file;
  c->fp = fp;
  c->ci = case_insensitive != 0;
  c->ip = ignore_punctuation != 0;
  c->ew = elide_whitespace != 0;
  c->stats.smallest_pattern_length = UINT32_MAX; // pattern_store_append.c:105-108
  return c;
}

int omega_list_matcher_compiler_add_pattern(omega_list_matcher_compiler_t *c, const uint8_t *pattern,
                                            uint32_t len) {
  if (!c || !pattern || len == 0) return -1;
  const uint8_t *p = pattern;
  if (c->ci || c->ip || c->ew) { // compiler.c:203-206
    c->scratch.resize(len);
    len = olm::normalize_bytes(c->ci, c->ip, c->ew, pattern, len, c->scratch.data());
    p = c->scratch.data();
    if (len == 0) return -1; // nothing left to match on (reference: abort)
  }
  auto &st = c->stats;
  if (len <= 4) { // compiler.c:207-218 + short_matcher_add :78-130
    bool fresh;
    switch (len) {
    case 1:
      fresh = !(c->bitmap1[p[0] >> 3] & (1u << (p[0] & 7)));
      if (fresh) {
        c->bitmap1[p[0] >> 3] |= uint8_t(1u << (p[0] & 7));
        ++c->n1;
      }
      break;
    case 2: {
      const uint32_t v = (uint32_t(p[0]) << 8) | p[1];
      fresh = !(c->bitmap2[v >> 3] & (1u << (v & 7)));
      if (fresh) {
        c->bitmap2[v >> 3] |= uint8_t(1u << (v & 7));
        ++c->n2;
      }
      break;
    }
    case 3: {
      const uint32_t v = (uint32_t(p[0]) << 16) | (uint32_t(p[1]) << 8) | p[2];
      fresh = c->set3.insert(v);
      if (fresh) c->arr3.push_back(v);
      break;
    }
    default: {
      const uint32_t v = olm::gram_be(p);
      fresh = c->set4.insert(v);
      if (fresh) c->arr4.push_back(v);
    }
    }
    if (fresh) ++st.short_pattern_count;
    else ++st.duplicate_patterns;
    // the reference updates these for duplicates too (compiler.c:211-218)
    if (len < st.smallest_pattern_length) st.smallest_pattern_length = len;
    if (len > st.largest_pattern_length) st.largest_pattern_length = len;
    st.total_input_bytes += len;
    return 0;
  }

  // long pattern: pattern_store_append.c:19-63, then hash_table_insert (compiler.c:219-226)
  const uint32_t id = uint32_t(c->longs.size()) + 1;
  if (!c->long_set.insert(c->longs, p, len, id)) {
    ++st.duplicate_patterns;
    return 0;
  }
  const uint64_t off = c->store.size();
  c->store.insert(c->store.end(), p, p + len);
  const uint32_t gram = olm::gram_be(p);
  c->longs.push_back(LongPattern{gram, len, off});
  if (len < st.smallest_pattern_length) st.smallest_pattern_length = len;
  if (len > st.largest_pattern_length) st.largest_pattern_length = len;
  ++st.stored_pattern_count;
  st.total_input_bytes += len;
  st.total_stored_bytes = off + len;
  if ((float)(c->table_used + 1) / (float)c->table_size > 0.9) c->table_size <<= 1;
  if (c->gram_set.insert(gram)) ++c->table_used;
  return 0;
}

const omega_match_pattern_store_stats_t *omega_list_matcher_compiler_get_pattern_store_stats(
    const omega_list_matcher_compiler_t *c) {
  return c ? &c->stats : nullptr;
}

int omega_list_matcher_compiler_destroy(omega_list_matcher_compiler_t *c) {
  if (!c) return -1;
  using namespace olm;
  int rc = 0;
  // ---- buckets: group long patterns by gram, longest first (compiler.c:39-45, :271)
  std::vector<uint32_t> order(c->longs.size());
  for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    const LongPattern &x = c->longs[a], &y = c->longs[b];
    if (x.gram != y.gram) return x.gram < y.gram;
    return x.len > y.len;
  });
  struct Bucket {
    uint32_t gram, first, count;
  };
  std::vector<Bucket> buckets;
  for (uint32_t i = 0; i < order.size(); ++i) {
    const uint32_t g = c->longs[order[i]].gram;
    if (buckets.empty() || buckets.back().gram != g) buckets.push_back(Bucket{g, i, 0});
    ++buckets.back().count;
  }
  // ---- index: open addressing on the reference's hash, so that probe_bucket()
  // (hash_table.c:91-109) reaches every key from its home slot.
  const uint32_t tsize = c->table_size, mask = tsize - 1;
  std::vector<int32_t> slot_bucket(tsize, -1);
  for (uint32_t b = 0; b < buckets.size(); ++b) {
    uint32_t h = ref_index_hash(buckets[b].gram) & mask;
    while (slot_bucket[h] >= 0) h = (h + 1) & mask;
    slot_bucket[h] = int32_t(b);
  }
  std::vector<uint32_t> index(tsize, 0); // unused slots are written as 0 (compiler.c:301)
  std::vector<uint8_t> blob;
  blob.reserve(buckets.size() * 8 + c->longs.size() * kBucketRecordBytes);
  uint32_t min_b = UINT32_MAX, max_b = 0;
  for (uint32_t s = 0; s < tsize; ++s) {
    if (slot_bucket[s] < 0) continue;
    const Bucket &b = buckets[slot_bucket[s]];
    index[s] = uint32_t(blob.size());
    min_b = std::min(min_b, b.count);
    max_b = std::max(max_b, b.count);
    uint8_t rec[16];
    wr32(rec, b.gram);
    wr32(rec + 4, b.count);
    blob.insert(blob.end(), rec, rec + 8);
    for (uint32_t j = 0; j < b.count; ++j) {
      const LongPattern &p = c->longs[order[b.first + j]];
      wr64(rec, p.off);
      wr32(rec + 8, p.len);
      wr32(rec + 12, 0);
      blob.insert(blob.end(), rec, rec + 16);
    }
  }
  // ---- Bloom over the bucket grams: table_size*16 bits, three probes h1 + i*h2
  // (bloom.c:12-17, :37-49)
  const uint32_t bloom_bits = tsize * 16u;
  std::vector<uint64_t> bloom(bloom_bits >> 6, 0);
  for (const Bucket &b : buckets) {
    const uint32_t h1 = ref_fmix32(b.gram), h2 = b.gram * 0x9e3779b1u;
    for (uint32_t i = 0; i < 3; ++i) {
      const uint32_t bp = (h1 + i * h2) & (bloom_bits - 1);
      bloom[bp >> 6] |= 1ull << (bp & 63);
    }
  }
  // ---- short section (compiler.c:333-357)
  std::vector<uint8_t> shorts;
  if (c->n1 || c->n2 || !c->arr3.empty() || !c->arr4.empty()) {
    std::sort(c->arr3.begin(), c->arr3.end());
    std::sort(c->arr4.begin(), c->arr4.end());
    shorts.insert(shorts.end(), kMagicShort, kMagicShort + 8);
    shorts.insert(shorts.end(), c->bitmap1, c->bitmap1 + 32);
    shorts.insert(shorts.end(), c->bitmap2, c->bitmap2 + 8192);
    uint8_t n[16];
    wr32(n, c->n1);
    wr32(n + 4, c->n2);
    wr32(n + 8, uint32_t(c->arr3.size()));
    wr32(n + 12, uint32_t(c->arr4.size()));
    shorts.insert(shorts.end(), n, n + 16);
    const uint8_t *a3 = reinterpret_cast<const uint8_t *>(c->arr3.data());
    const uint8_t *a4 = reinterpret_cast<const uint8_t *>(c->arr4.data());
    shorts.insert(shorts.end(), a3, a3 + c->arr3.size() * 4);
    shorts.insert(shorts.end(), a4, a4 + c->arr4.size() * 4);
  }
  // ---- header (compiler.c:245-295, :359-368)
  uint8_t hdr[kHeaderBytes] = {0};
  std::memcpy(hdr, kMagicHeader, 8);
  wr32(hdr + 8, kFormatVersion);
  uint32_t flags = 0; // compiler.c:170-178
  if (c->ci) flags |= kFlagIgnoreCase;
  if (c->ip) flags |= kFlagIgnorePunct;
  if (c->ew) flags |= kFlagElideSpace;
  wr32(hdr + 12, flags);
  wr64(hdr + 16, c->store.size());
  wr32(hdr + 24, c->stats.stored_pattern_count);
  wr32(hdr + 28, c->stats.smallest_pattern_length);
  wr32(hdr + 32, c->stats.largest_pattern_length);
  wr32(hdr + 36, bloom_bits >> 3);
  wr32(hdr + 40, uint32_t(blob.size()));
  wr32(hdr + 44, tsize);
  wr32(hdr + 48, uint32_t(buckets.size()));
  wr32(hdr + 52, min_b == UINT32_MAX ? 0 : min_b);
  wr32(hdr + 56, max_b);
  wr32(hdr + 60, uint32_t(shorts.size()));
  const float lf = tsize ? float(buckets.size()) / float(tsize) : 0.f;
  const float avg = buckets.empty() ? 0.f : float(c->stats.stored_pattern_count) / float(buckets.size());
  std::memcpy(hdr + 64, &lf, 4);
  std::memcpy(hdr + 68, &avg, 4);

  auto put = [&](const void *p, size_t n) {
    if (n && std::fwrite(p, 1, n, c->fp) != n) rc = -1;
  };
  put(hdr, sizeof hdr);
  put(c->store.data(), c->store.size());
  put(kMagicBloom, 8);
  put(&bloom_bits, 4);
  put(bloom.data(), bloom.size() * 8);
  put(kMagicHash, 8);
  put(index.data(), index.size() * 4);
  put(blob.data(), blob.size());
  put(shorts.data(), shorts.size());
  if (std::fclose(c->fp) != 0) rc = -1;
  delete c;
  return rc;
}

int omega_list_matcher_compile_patterns(const char *compiled_file, const uint8_t *buf, uint64_t size,
                                        int case_insensitive, int ignore_punctuation,
                                        int elide_whitespace,
                                        omega_match_pattern_store_stats_t *out_stats) {
  if (!compiled_file || !buf || size == 0) return -1;
  Compiler *c = omega_list_matcher_compiler_create(compiled_file, case_insensitive,
                                                   ignore_punctuation, elide_whitespace);
  if (!c) return -1;
  int rc = 0;
  // compiler.c:401-415: split on '\n', drop one trailing '\r', skip empty lines
  const uint8_t *p = buf, *end = buf + size;
  while (p < end) {
    const uint8_t *nl = static_cast<const uint8_t *>(std::memchr(p, '\n', size_t(end - p)));
    if (!nl) nl = end;
    uint32_t len = uint32_t(nl - p);
    if (len > 0 && p[len - 1] == '\r') --len;
    if (len > 0 && omega_list_matcher_compiler_add_pattern(c, p, len) != 0) rc = -1;
    p = nl + 1;
  }
  if (out_stats) *out_stats = c->stats;
  if (omega_list_matcher_compiler_destroy(c) != 0) rc = -1;
  return rc;
}

int omega_list_matcher_compile_patterns_filename(const char *compiled_file, const char *patterns_file,
                                                 int case_insensitive, int ignore_punctuation,
                                                 int elide_whitespace,
                                                 omega_match_pattern_store_stats_t *out_stats) {
  if (!compiled_file || !patterns_file) return -1;
  size_t n = 0;
  uint8_t *map = olm::map_whole_file(patterns_file, &n, true);
  if (!map) return -1;
  const int rc = omega_list_matcher_compile_patterns(compiled_file, map, n, case_insensitive,
                                                     ignore_punctuation, elide_whitespace, out_stats);
  olm::unmap(map, n);
  return rc;
}

int omega_list_matcher_is_compiled(const char *compiled_file) {
  if (!compiled_file) return 0;
  FILE *fp = std::fopen(compiled_file, "rb");
  if (!fp) return 0;
  char magic[8];
  const size_t n = std::fread(magic, 1, 8, fp);
  std::fclose(fp);
  return n == 8 && std::memcmp(magic, olm::kMagicHeader, 8) == 0;
}

} // extern "C"
