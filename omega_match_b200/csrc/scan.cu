// scan.cu -- the per-offset scan (kernels K3 + K4 of SURVEY 2.1) for sm_100a.
//
// Replaces the hot loop of core_match() (omega_match/src/matcher.c:767-881) together with
// bloom_filter_query (bloom.c:51-64), probe_bucket (hash_table.c:91-109),
// scan_bucket_and_append (matcher.c:182-255), the short matcher (matcher.c:665-692,
// :804-880) and -- because matches leave the kernel already in final order -- the
// concatenate + radix sort of finalize_match_results (matcher.c:587-623, :258-325).
//
// Shape of scan_kernel
//   * persistent CTAs (grid = #SMs) of 32 warps: 31 scanning warps and one producer warp.  Tiles
//     of 4 KiB positions are handed out by an atomic ticket, so tile k is always started before
//     k+1; inside a CTA the 8 chunks (512 positions) of a tile are grabbed dynamically by the
//     scanning warps -- no warp waits for a slower one;
//   * producer: each tile (+16 bytes in front, +112 behind) is brought into shared memory by
//     ONE cp.async.bulk (TMA, 1-D, L2 evict-first) that completes on a `full` mbarrier; ring of
//     2..16 stages (as many as fit beside the filters; any depth -- stage and phase parity travel
//     with the tile description); a stage is refilled as soon as its 8 chunks have arrived on
//     `scanned` -- no CTA-wide barrier anywhere;
//   * stage 1, every position, in registers:
//       - stores whose patterns all start with a run of bytes from a small class (letters ...):
//         SWAR range tests on the haystack words -> 1 bit per position (no memory access);
//       - else: big-endian gram by PRMT from two registers, one multiply, one probe of the
//         hashed gram bitmap in shared memory (and one of the short-pattern bitmap when the
//         store has 1..3 byte patterns);
//     survivors are compacted, in position order, into the warp's queue Q1;
//   * stage 2a, Q1 entries, 32 x kProbeUnroll at a time: position predicates, (class mode: the
//     gram bitmap probe,) ONE 16-byte load (ld.global.cg) of the key's bucket -- almost every
//     false candidate ends here; key hits (and short-pattern candidates) are compacted into Q2;
//   * stage 2b, Q2 entries, 32 at a time, all lanes busy: the slot (pattern bytes 4..11 +
//     length), remaining bytes against the pattern store, end predicates, the 4/3/2/1-byte
//     sets; accepted matches are appended in candidate order (ballot + popc) to the chunk's
//     staging area;
//   * hand-over: the warp that scanned a chunk takes a run of temp[] from its own block (one
//     global atomic per 1024 entries), copies the staged matches there as packed 4-byte
//     entries (position order) and writes the chunk descriptor {count, temp_index}.  No warp
//     waits for another warp, no CTA for another CTA;
//   * prefix_sum_kernel / prefix_scan_kernel: first result index of every span of 4096 chunks;
//     place_kernel: per-chunk prefix inside the span, packed entries -> final 24-byte records at
//     their final index (offset ascending, length descending: the order radix_sort_matches
//     produces, without a sort);
//   * a chunk whose matches do not fit its staging area is only counted (counts are always
//     exact); redo_kernel (exits at once when there is none) evaluates such chunks again and
//     writes their records directly at the index place_kernel computed.
#include "scan.cuh"

#include <cstddef>
#include <cstdio>

#include "olm_classes.h"
#include "olm_format.h"
#include "scan_device.cuh"

namespace olm {

namespace {

using namespace dev;

#ifndef OLM_FAST_UNROLL
#define OLM_FAST_UNROLL 2
#endif
// chunks a scanning warp takes per grab.  2 was slower with the 4-stage ring of an earlier version
// (nothing left to prefetch into); kept as a build knob for the deeper rings.
#ifndef OLM_GRAB
#define OLM_GRAB 1
#endif
constexpr uint32_t kGrab = OLM_GRAB;
static_assert(kGrab >= 1 && kTileChunks % kGrab == 0, "a grab stays inside one tile");
// Tiles per ticket.  The producer lane takes the tickets of the NEXT group while it hands out the
// current one, so the latency of the global atomic (~1 us under load, once per tile before) is off
// its critical path.
#ifndef OLM_TICKET_BATCH
#define OLM_TICKET_BATCH 4
#endif
// 1: a scanning warp that is ahead of the producer sleeps on an mbarrier of the tile description
// (hardware wait) instead of polling `seq` in shared memory.
#ifndef OLM_DESC_BAR
#define OLM_DESC_BAR 1
#endif
#ifndef OLM_CLS_SKIP_BITMAP
#define OLM_CLS_SKIP_BITMAP 0
#endif

// shared memory header (kSmemHeader bytes)
struct SmemHeader {
  uint64_t full[kMaxStages];
  uint64_t scanned[kMaxStages];
  uint32_t chunk_ctr; // next chunk of the CTA's tile sequence
  uint32_t end_k;     // first tile iteration without a tile (kNoTile until the tickets run out)
  StageInfo info[kInfoRing];
  uint64_t described[kInfoRing]; // OLM_DESC_BAR: phase (k / kInfoRing) of entry k % kInfoRing completes when iteration k is described
};
static_assert(sizeof(SmemHeader) <= kSmemHeader, "header does not fit");

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk copy global -> shared through the TMA unit; completes `bytes` on `bar`.
// The haystack is read once: L2 evict-first keeps the tables resident instead.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// One bucket of the key table.  The buckets never hit in L1 (a 32 MiB table probed at random), so
// they are read with ld.global.cg -- cached in L2 only: 620 vs 608 GB/s at 1 M patterns with
// __ldg; ld.global.nc.L1::no_allocate: 253.  OLM_KEY_LOAD=0 restores __ldg.
#ifndef OLM_KEY_LOAD
#define OLM_KEY_LOAD 1
#endif
__device__ __forceinline__ uint4 ld_keys(const uint4 *p) {
#if OLM_KEY_LOAD == 0
  return __ldg(p);
#else
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#endif
}
__device__ __forceinline__ void mbar_arrive32(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// One final record (omega_match_result_t, list_matcher.h:19-23) in source coordinates.
__device__ __forceinline__ void put_record(const ScanParams &P, unsigned long long r, unsigned long long off, uint32_t len) {
  Record *o = P.out + r;
  o->offset = off;
  *reinterpret_cast<unsigned long long *>(&o->len) = (unsigned long long)len;
  o->ptr = P.match_ptr_base + off;
}

enum ChunkMode { kStageMode = 0, kCountMode = 1, kDirectMode = 2 };

// XF: the store has a transform flag (window mode): chunks are case-folded or normalised into the
// warp's private buffer; plain stores never see that code.
// COOP: keys shared by many patterns are compared by the whole warp (verify_batch); chosen for stores
// that have such keys (DeviceStore::max_recs), compiled out elsewhere -- the kernels are register bound.
// SX: short candidates get their second look (short_look) before they cost a verify; kernels of their
// own -- plain stores whose p23 is the weak two-byte filter -- so that the others do not carry the code.
template <bool HAS_G4, bool HAS_P23, bool HAS_CLS, bool XF, bool COOP, bool SX = false>
struct Scanner {
  const ScanParams &P;
  const uint32_t *g4s;
  const uint32_t *p23s;
  const uint32_t fl;
  // statistics of omega_match_stats_t that cost nothing extra (list_matcher.h:43-49):
  // hits = buckets found + short matches accepted, misses = short candidates rejected by a
  // predicate, comparisons = bucket patterns that fit (matcher.c:783-799, :818-877, :210)
  mutable uint32_t n_hits = 0, n_miss = 0, n_cmp = 0, n_long_hits = 0;
  mutable uint32_t stat_inc = 1; // 0 while a position is evaluated a second time

  uint32_t xf32 = 0; // normalising stores: the warp's rows / walk state (scan_device.cuh), else 0
  uint32_t sx32 = 0; // shared-space address of the short candidates' second look (DeviceStore::sx), 0 = not used

  // A p23 candidate's second look: is one of the 1 / 2 / 3 byte patterns really a prefix of the
  // position's bytes (gram: bytes 0..3, big endian)?  Exact for 1 and 2 bytes (one bitmap over the first
  // two bytes), a hashed bitmap with >= 64 bits per pattern for 3.  Only what passes costs a Q2 entry
  // and a verify (the p23 bitmap alone lets 7-11 % of all positions through on the BASELINE stores:
  // 34-56 entries per chunk for 4-8 matches).
  __device__ __forceinline__ bool short_look(uint32_t gram) const {
    const uint32_t b = gram >> 16;
    bool hit = (lds32(sx32 + ((b >> 5) << 2)) >> (b & 31)) & 1u;
    if (P.st.n3) {
      const uint32_t b3 = ((gram >> 8) * kHashMul) >> P.st.sx3_shift;
      hit = hit || ((lds32(sx32 + 8192u + ((b3 >> 5) << 2)) >> (b3 & 31)) & 1u);
    }
    return hit;
  }

  __device__ __forceinline__ Scanner(const ScanParams &p, const uint32_t *g4, const uint32_t *p23)
      : P(p), g4s(g4), p23s(p23), fl(p.flags) {}

  // Byte `rel` of the segment T describes; kBeyond when a normalised window ends before it.
  __device__ __forceinline__ uint32_t hay_byte(const TileCtx &T, uint32_t rel) const {
    if (rel < T.staged) return lds8(T.sb32 + kTilePre + rel);
    if (XF && xf32) return xf_walk(P.buf + T.boff, P.store_flags, xf32, rel, false).byte;
    const uint32_t c = P.buf[T.boff + (long long)rel];
    return XF ? upper_byte(c) : c; // (a window-mode store without rows only folds case)
  }
  // A match at position tpos of len bytes, as the matcher saw it -> in source coordinates relative
  // to T's first source byte (matcher.c:986-997: the first and the last byte are mapped back).
  __device__ __forceinline__ uint32_t src_of(const TileCtx &T, uint32_t j) const {
    const uint32_t tab_n = lds32(xf32 + kXfCurIdx); // bytes the per-byte table covers
    if (j < tab_n) return xf_src(P.buf + T.boff, P.store_flags, xf32, j);
    if (j < T.staged) return lds32(xf32 + kXfExt + 4u * (j - tab_n));
    return xf_walk(P.buf + T.boff, P.store_flags, xf32, j, false).src;
  }
  __device__ __forceinline__ uint32_t locate_start(const TileCtx &T, uint32_t tpos) const {
    return (XF && xf32) ? src_of(T, tpos) : tpos;
  }
  __device__ __forceinline__ uint32_t locate_len(const TileCtx &T, uint32_t tpos, uint32_t len, uint32_t spos) const {
    return (XF && xf32) ? src_of(T, tpos + len - 1) - spos + 1 : len;
  }

  // Multi-pattern slots: are bytes K and K + 1 of the position (hay = its first 8 bytes, little endian)
  // among the bytes that follow the key in the slot's patterns (Slot::w0 / w1, store.cpp)?
  __device__ __forceinline__ bool next_bytes_ok(unsigned long long hay, uint32_t m0, uint32_t m1) const {
    const uint32_t k8 = P.st.key_bytes * 8u;
    const uint32_t b0 = (uint32_t)(hay >> (k8 < 56u ? k8 : 56u)), b1 = (uint32_t)(hay >> (k8 < 48u ? k8 + 8u : 56u));
    return ((m0 >> (b0 & 31u)) & (m1 >> (b1 & 31u)) & 1u) != 0;
  }
  // bytes [8, len) of a candidate against the pattern store (bytes 0..7 are already equal)
  __device__ __forceinline__ bool tail_equal(const TileCtx &T, uint32_t tpos, uint32_t len, uint32_t store_off) const {
    const uint8_t *pat = P.st.store + store_off;
    for (uint32_t i = 8; i < len; ++i)
      if (hay_byte(T, tpos + i) != __ldg(pat + i)) return false;
    return true;
  }

  // end-side predicates for a long match (matcher.c:233, :239, :247): all guarded by e < n
  __device__ __forceinline__ bool end_ok_long(const TileCtx &T, uint32_t tpos, uint32_t len) const {
    if (!(fl & (kWordBoundary | kWordSuffix | kLineEnd))) return true;
    if (tpos + len >= T.rem0) return true;
    const uint32_t c = hay_byte(T, tpos + len);
    if (c == kBeyond) return true; // the normalised window ends with the match
    if ((fl & (kWordBoundary | kWordSuffix)) && is_word_byte(c)) return false;
    if ((fl & kLineEnd) && !is_line_end_byte(c)) return false;
    return true;
  }
  // short matcher, matcher.c:804-880: for lengths 2..4 the word-boundary test reads
  // haystack[pos+L] without a bound (:812,:830,:848) -> `tail` stands in at pos+L == n.
  __device__ __forceinline__ bool end_ok_short(const TileCtx &T, uint32_t tpos, uint32_t L) const {
    if (!(fl & (kWordBoundary | kWordSuffix | kLineEnd))) return true;
    bool inside = tpos + L < T.rem0;
    uint32_t c = inside ? hay_byte(T, tpos + L) : T.tail;
    if (c == kBeyond) { // the normalised window ends with the match
      inside = false;
      c = T.tail;
    }
    if (fl & kWordBoundary) {
      if (L == 1) {
        if (inside && is_word_byte(c)) return false;
      } else if (is_word_byte(c)) {
        return false;
      }
    }
    if ((fl & kWordSuffix) && inside && is_word_byte(c)) return false;
    if ((fl & kLineEnd) && inside && !is_line_end_byte(c)) return false;
    return true;
  }

  // ---- one candidate position in three steps, so that many loads are in flight ----
  // probe_issue : position predicates (matcher.c:770-776, :195-196, :806-807), the gram,
  //               (class mode: the gram bitmap,) and the load of the gram's key bucket;
  // probe_key   : the key compare; yields the slot index of a hit (or kNoSlot) and whether
  //               the position goes on to
  // verify      : everything else the reference does for the position (matcher.c:782-880);
  //               `emit(len)` is called once per accepted match, longest first.
  static constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
  static constexpr uint32_t kCoopMin = kCoopMinRecs; // records behind one key from which on the warp compares them together
  struct Probe {
    uint32_t tpos, gram, bucket, flags; // gram: the position's key; flags: 1 = alive, 2 = key candidate (enough bytes left), 4 = short candidate, 8 = a start predicate failed (exact statistics only)
    uint4 kb;
  };

  __device__ __forceinline__ void probe_issue(const TileCtx &T, bool valid, uint32_t tpos, bool cand_g, bool cand_p,
                                              Probe &pr) const {
    pr.tpos = tpos;
    pr.flags = 0;
    pr.gram = 0;
    pr.bucket = 0;
    pr.kb = make_uint4(0, 0, 0, 0);
    uint32_t bad_start = 0;
    if (!valid) return;
    const uint32_t q = T.sb32 + kTilePre + tpos;
    if (fl & (kWordBoundary | kWordPrefix | kLineStart)) {
      const bool at0 = T.first && tpos == 0;
      const uint32_t prev = lds8(q - 1);
      if (fl & kWordBoundary) { // matcher.c:770-776
        const bool cw = is_word_byte(lds8(q));
        const bool pw = at0 ? false : is_word_byte(prev);
        if (cw == pw) return;
      }
      bool bad = (fl & kWordPrefix) && !at0 && is_word_byte(prev);       // :195, :806
      bad = bad || ((fl & kLineStart) && !at0 && !is_line_end_byte(prev)); // :196, :807
      if (bad) {
        // nothing can match here.  Exact statistics: a short-matcher candidate still has to be looked
        // up, because the reference counts it as a miss (matcher.c:818-877)
        if (!((fl & kCountAll) && (cand_p || (cand_g && P.st.n4)))) return;
        bad_start = 8u;
      }
    }
    const uint32_t gram = __byte_perm(lds_le32(q), 0, 0x0123);
    if (SX && cand_p) cand_p = short_look(gram);
    pr.flags = 1u | (cand_p ? 4u : 0u) | bad_start;
    if (HAS_G4 && cand_g && T.rem0 - tpos >= P.st.key_bytes) {
      const uint32_t h = key_hash(gram, P.st.tail_mask ? lds_le32(q + 4) & P.st.tail_mask : 0u);
      pr.gram = h; // the key of the position
      if (HAS_CLS) { // Q1 holds class survivors: the gram bitmap is probed here
        const uint32_t b = h >> P.st.g4_shift;
        if (!((g4s[b >> 5] >> (b & 31)) & 1u)) return;
      }
      pr.flags |= 2u;
      pr.bucket = h >> P.st.key_shift;
      pr.kb = __ldg(P.st.keys + pr.bucket);
    }
  }

  // returns true when the position needs verify(); *slot = slot index of the gram or kNoSlot
  __device__ __forceinline__ bool probe_key(const Probe &pr, uint32_t *slot) const {
    *slot = kNoSlot;
    if (!(pr.flags & 1u)) return false;
    if (HAS_G4 && (pr.flags & 2u)) {
      // the gram is in the first bucket (from its home) that has it; a bucket whose last
      // place is unused ends the probe
      const uint32_t gram = pr.gram;
      uint4 kb = pr.kb;
      uint32_t bucket = pr.bucket;
      int place;
      while (true) {
        place = kb.x == gram ? 0 : kb.y == gram ? 1 : kb.z == gram ? 2 : kb.w == gram ? 3 : -1;
        if (place >= 0 || kb.w == P.st.empty_key) break;
        bucket = (bucket + 1) & P.st.key_mask;
        kb = __ldg(P.st.keys + bucket);
      }
      if (place >= 0) *slot = 4 * bucket + (uint32_t)place;
    }
    return *slot != kNoSlot || (HAS_P23 && (pr.flags & 4u));
  }

  // `s` = the position's slot (when slot != kNoSlot); coop_done: the slot's records have been
  // compared by the whole warp already (verify_batch), `emitted` tells whether one of them matched.
  template <typename Emit>
  __device__ __forceinline__ void verify(const TileCtx &T, uint32_t tpos, uint32_t slot, const uint4 &s_in, bool cand_p,
                                         bool bad_start, bool coop_done, bool emitted, Emit &&emit) const {
    const uint32_t rem = T.rem0 - tpos;
    const uint32_t q = T.sb32 + kTilePre + tpos;
    const bool longest = fl & kLongestOnly;
    // exact statistics: the reference counts the short matcher's hits and misses before its
    // longest filter runs (matcher.c:818-877 vs :611), so what `longest` lets this kernel skip
    // is still evaluated -- and counted -- but not emitted
    const bool count_all = fl & kCountAll;

    if (HAS_G4 && slot != kNoSlot) {
      const uint4 s = COOP ? s_in : __ldg(reinterpret_cast<const uint4 *>(P.st.slots + slot));
      const uint32_t meta = s.z; // 0 when the gram equals empty_key and matched an unused place
      if (meta != 0 && bad_start) { // exact statistics: only the 4-byte set is looked at
        if (meta & kSlotShort4) n_miss += stat_inc;
      } else if (meta != 0) {
        const unsigned long long hay = ((unsigned long long)lds_le32(q + 4) << 32) | lds_le32(q);
        if (meta & kSlotValueMask) {
          n_hits += stat_inc;
          n_long_hits += stat_inc;
        }
        // bytes 0..7 of a pattern of length len (>= 5) against the haystack, one 64-bit compare
        auto head_equal = [&](uint32_t len, uint32_t w0, uint32_t w1) {
          const uint32_t drop = len >= 8 ? 0u : (8u - len) * 8u; // bits of the window beyond the pattern
          return ((hay ^ (((unsigned long long)w1 << 32) | w0)) << drop) == 0;
        };
        if (meta & kSlotMulti) {
          // (none of the slot's patterns goes on with the position's next bytes: nothing to compare)
          const uint32_t cnt = (coop_done || !next_bytes_ok(hay, s.x, s.y)) ? 0u : meta & kSlotValueMask;
          for (uint32_t j = 0; j < cnt; ++j) {
            const uint4 r = __ldg(reinterpret_cast<const uint4 *>(P.st.recs + s.w + j));
            const uint32_t len = r.y;
            if (len > rem) continue; // matcher.c:203
            n_cmp += stat_inc;
            if (!head_equal(len, r.x, r.w)) continue;
            if (len > 8 && !tail_equal(T, tpos, len, r.z)) continue;
            if (!end_ok_long(T, tpos, len)) continue;
            emit(len);
            emitted = true;
            if (longest) break;
          }
        } else {
          const uint32_t len = meta & kSlotValueMask;
          if (len != 0 && len <= rem) {
            n_cmp += stat_inc;
            if (head_equal(len, s.x, s.y) && (len <= 8 || tail_equal(T, tpos, len, s.w)) &&
                end_ok_long(T, tpos, len)) {
              emit(len);
              emitted = true;
            }
          }
        }
        if ((meta & kSlotShort4) && (count_all || !(longest && emitted))) {
          if (end_ok_short(T, tpos, 4)) {
            if (!(longest && emitted)) emit(4u);
            emitted = true;
            n_hits += stat_inc;
          } else {
            n_miss += stat_inc;
          }
        }
      }
    }
    if (HAS_P23 && cand_p && (count_all || !(longest && emitted))) {
      const uint32_t gram = __byte_perm(lds_le32(q), 0, 0x0123);
      if (P.st.n3 && rem >= 3) {
        const uint32_t k3 = gram >> 8;
        bool hit = false;
        for (uint32_t j = (k3 * kHashMul >> 8) & P.st.set3_mask;; j = (j + 1) & P.st.set3_mask) {
          const uint32_t v = __ldg(P.st.set3 + j);
          if (v == 0) break;
          if (v == k3 + 1) {
            hit = true;
            break;
          }
        }
        if (hit) {
          if (!bad_start && end_ok_short(T, tpos, 3)) {
            if (!(longest && emitted)) emit(3u);
            emitted = true;
            n_hits += stat_inc;
          } else {
            n_miss += stat_inc;
          }
        }
      }
      if (P.st.n2 && rem >= 2 && (count_all || !(longest && emitted))) {
        const uint32_t k2 = gram >> 16;
        if ((__ldg(P.st.bitmap2 + (k2 >> 5)) >> (k2 & 31)) & 1u) {
          if (!bad_start && end_ok_short(T, tpos, 2)) {
            if (!(longest && emitted)) emit(2u);
            emitted = true;
            n_hits += stat_inc;
          } else {
            n_miss += stat_inc;
          }
        }
      }
      if (P.st.n1 && (count_all || !(longest && emitted))) {
        const uint32_t k1 = gram >> 24;
        if ((P.st.bitmap1[k1 >> 5] >> (k1 & 31)) & 1u) {
          if (!bad_start && end_ok_short(T, tpos, 1)) {
            if (!(longest && emitted)) emit(1u);
            n_hits += stat_inc;
          } else {
            n_miss += stat_inc;
          }
        }
      }
    }
  }

  // bits 7,15,23,31 -> bits 0..3
  static __device__ __forceinline__ uint32_t gather4(uint32_t m) { return __umulhi(m, 0x02040810u) & 0xFu; }

  // class bits of the 24 haystack bytes a lane holds (bit i = byte i is in the class), then
  // bit i = bytes i .. i+run-1 are all in the class (run = 4, 5, 6 or 8; i < 16)
  __device__ __forceinline__ uint32_t class_runs16(const uint4 &v, const uint2 &nx) const {
    const ByteClass &c = P.st.cls;
    const uint32_t w[6] = {v.x, v.y, v.z, v.w, nx.x, nx.y};
    const uint32_t and4 = c.and4, lo0 = c.addlo[0], hi0 = c.addhi[0];
    uint32_t a = 0;
    if (c.n_ranges > 1) { // (uniform branch: one block of straight-line code per case)
      const uint32_t lo1 = c.addlo[1], hi1 = c.addhi[1];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const uint32_t t = w[i] & and4;
        const uint32_t in = ((t + lo0) & ~(t + hi0)) | ((t + lo1) & ~(t + hi1));
        a |= gather4(in & ~w[i] & 0x80808080u) << (4 * i);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const uint32_t t = w[i] & and4;
        a |= gather4((t + lo0) & ~(t + hi0) & ~w[i] & 0x80808080u) << (4 * i);
      }
    }
    a &= a >> 1;
    a &= a >> 2;
    a &= a >> (c.run - 4);
    return a & 0xFFFFu;
  }
  // stage 1 for one lane: 16 positions -> candidate masks (bit k = position lpos + k)
  __device__ __forceinline__ void stage1(const TileCtx &T, uint32_t lpos, uint32_t &cg, uint32_t &cp) const {
    const uint32_t src = T.sb32 + kTilePre + lpos;
    const uint4 v = lds128(src);
    cg = 0;
    cp = 0;
    if (HAS_CLS) {
      cg = class_runs16(v, lds64(src + 16));
      return;
    }
    const uint2 w45 = lds64(src + 16);
    const uint32_t w[6] = {v.x, v.y, v.z, v.w, w45.x, w45.y};
    const uint32_t tmask = P.st.tail_mask;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t gram = __byte_perm(w[k >> 2], w[(k >> 2) + 1], 0x0123u + 0x1111u * (k & 3));
      if (HAS_G4) {
        // (keys longer than 4 bytes: the little-endian word of bytes k+4..k+7 joins the hash)
        const uint32_t tail = __byte_perm(w[(k >> 2) + 1], w[(k >> 2) + 2 > 5 ? 5 : (k >> 2) + 2], 0x3210u + 0x1111u * (k & 3));
        const uint32_t b = key_hash(gram, tail & tmask) >> P.st.g4_shift;
        cg |= ((g4s[b >> 5] >> (b & 31)) & 1u) << k;
      }
      if (HAS_P23) {
        const uint32_t b = ((gram & P.st.p23_and) * P.st.p23_mul) >> P.st.p23_shift;
        cp |= ((p23s[b >> 5] >> (b & 31)) & 1u) << k;
      }
    }
  }

  // Stage 2b for up to 32 entries of Q2 (lane i takes entry i): verify, then append the
  // accepted matches in entry order (ballot + popc; a shuffle scan only when a position has
  // several matches; a position with more than four matches is evaluated a second time for
  // the rest).  Returns the number of matches of the batch.
  //   kStageMode : matches go to `stage` (packed, shared memory, `cap` entries); when they
  //                do not fit, *overflow is set (the tile goes on the redo list);
  //   kCountMode : nothing is written;
  //   kDirectMode: matches go to P.out[out_base ...] as final records.
  //   `cbase`: position of the chunk inside T (0 when T is a chunk in a private buffer); staged
  //   entries and records are in SOURCE coordinates (locate()).
  template <int MODE>
  __device__ __forceinline__ uint32_t verify_batch(const TileCtx &T, uint32_t q2, uint32_t n, uint32_t lane,
                                                   uint32_t stage, uint32_t used, uint32_t cap, uint32_t cbase,
                                                   unsigned long long out_base, uint32_t *overflow) const {
    bool mine = lane < n;
    const unsigned long long ent = mine ? lds64u(q2 + 8u * lane) : 0ull; // q2, stage: shared-space addresses
    const uint32_t slot = (uint32_t)ent, hi = (uint32_t)(ent >> 32);
    const uint32_t tpos = hi & 0xFFFFu;
    const bool cand_p = (hi >> 16) & 1u, bad_start = (hi >> 17) & 1u;
    // One copy of verify() in the instruction stream: positions with more than four matches are
    // evaluated again (skip = 4, 8, ...) by the same code -- rare, and the kernel is I-cache bound.
    uint32_t at = 0, tot = 0, skip = 0;
    const uint32_t keep = stat_inc;
    bool again;
    // the position's slot: pattern bytes 0..7 + length, or where the records of its key are
    uint4 s = make_uint4(0, 0, 0, 0);
    if (COOP && HAS_G4 && mine && slot != kNoSlot) s = __ldg(reinterpret_cast<const uint4 *>(P.st.slots + slot));
    do {
      uint32_t cnt = 0, m0 = 0, m1 = 0, m2 = 0, m3 = 0;
      auto emit = [&](uint32_t len) {
        const uint32_t j = cnt - skip; // wraps to a huge value while cnt < skip
        if (j == 0) m0 = len;
        else if (j == 1) m1 = len;
        else if (j == 2) m2 = len;
        else if (j == 3) m3 = len;
        ++cnt;
      };
      // Keys shared by many patterns (name lists: hundreds of surnames behind one 4-byte prefix): the
      // WARP compares such a key's records, 32 at a time with coalesced loads, instead of one lane
      // walking them one dependent load after the other (probe_bucket + memcmp of matcher.c:182-255
      // as a warp-cooperative compare).  Matches reach the owning lane in record order (longest first).
      bool coop = false, emitted0 = false;
      if (COOP && HAS_G4) {
        coop = mine && s.z != 0 && !bad_start && (s.z & kSlotMulti) && (s.z & kSlotValueMask) >= kCoopMin;
        if (coop) { // (a key none of whose patterns goes on with the position's next bytes is not compared at all)
          const uint32_t mq = T.sb32 + kTilePre + tpos;
          coop = next_bytes_ok(((unsigned long long)lds_le32(mq + 4) << 32) | lds_le32(mq), s.x, s.y);
        }
        uint32_t todo = __ballot_sync(kFull, coop);
        while (todo) {
          const uint32_t src = __ffs(todo) - 1;
          todo &= todo - 1;
          const uint32_t o_tpos = __shfl_sync(kFull, tpos, src), o_first = __shfl_sync(kFull, s.w, src);
          const uint32_t o_cnt = __shfl_sync(kFull, s.z, src) & kSlotValueMask;
          const uint32_t oq = T.sb32 + kTilePre + o_tpos, o_rem = T.rem0 - o_tpos;
          const unsigned long long hay = ((unsigned long long)lds_le32(oq + 4) << 32) | lds_le32(oq);
          uint32_t ncmp = 0;
          for (uint32_t base = 0; base < o_cnt; base += 32) {
            const bool valid = base + lane < o_cnt;
            uint4 r = make_uint4(0, 0, 0, 0);
            if (valid) r = __ldg(reinterpret_cast<const uint4 *>(P.st.recs + o_first + base + lane));
            const bool fits = valid && r.y <= o_rem; // matcher.c:203
            const uint32_t drop = r.y >= 8 ? 0u : (8u - r.y) * 8u;
            bool ok = fits && ((hay ^ (((unsigned long long)r.w << 32) | r.x)) << drop) == 0;
            if (ok && r.y > 8) ok = tail_equal(T, o_tpos, r.y, r.z);
            if (ok) ok = end_ok_long(T, o_tpos, r.y);
            const uint32_t fits_bal = __ballot_sync(kFull, fits);
            uint32_t bal = __ballot_sync(kFull, ok);
            const bool stop = (fl & kLongestOnly) && bal; // only the first (longest) match counts
            if (stop) bal &= 0u - bal;
            ncmp += __popc(stop ? fits_bal & ((bal << 1) - 1u) : fits_bal);
            while (bal) {
              const uint32_t b = __ffs(bal) - 1;
              bal &= bal - 1;
              const uint32_t len = __shfl_sync(kFull, r.y, b);
              if (lane == src) {
                emit(len);
                emitted0 = true;
              }
            }
            if (stop) break;
          }
          if (lane == src) n_cmp += ncmp * stat_inc;
        }
      }
      if (mine) verify(T, tpos, slot, s, cand_p, bad_start, coop, emitted0, emit);
      if (skip == 0) { // where this lane's matches go: ballot + popc, a shuffle scan only when needed
        const uint32_t bal = __ballot_sync(kFull, cnt > 0);
        if (!bal) return 0;
        uint32_t pre;
        if (!__any_sync(kFull, cnt > 1)) {
          pre = __popc(bal & ((1u << lane) - 1u));
          tot = __popc(bal);
        } else {
          uint32_t in2 = cnt;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, in2, d);
            if (lane >= (uint32_t)d) in2 += t;
          }
          pre = in2 - cnt;
          tot = __shfl_sync(kFull, in2, 31);
        }
        at = used + pre;
      }
      // (the start of all matches of a position is translated once)
      const uint32_t spos = (MODE != kCountMode && cnt > skip) ? locate_start(T, tpos) : 0u;
      auto put = [&](uint32_t i, uint32_t len) {
        if (MODE == kCountMode) return;
        const uint32_t slen = locate_len(T, tpos, len, spos);
        if (MODE == kDirectMode) {
          const unsigned long long r = out_base + at + i;
          if (r < P.out_cap) put_record(P, r, T.gbase + spos, slen);
        } else if (MODE == kStageMode) {
          if (at + i < cap && !(slen >> kPackLenBits)) {
            sts32(stage + 4u * (at + i), ((spos - cbase) << kPackLenBits) | slen);
          } else {
            *overflow = 1;
          }
        }
      };
      if (cnt > skip) put(skip, m0);
      if (cnt > skip + 1) put(skip + 1, m1);
      if (cnt > skip + 2) put(skip + 2, m2);
      if (cnt > skip + 3) put(skip + 3, m3);
      skip += 4;
      stat_inc = 0;
      mine = mine && cnt > skip;
      again = MODE != kCountMode && __any_sync(kFull, mine);
    } while (again);
    stat_inc = keep;
    return tot;
  }

  // One 512-byte chunk: stage 1 -> Q1 -> key probes (kProbeUnroll x 32 in flight) -> Q2 ->
  // verify batches.  Returns the exact number of matches of the chunk in all modes.
  static constexpr int kProbeUnroll = 2;
  template <int MODE>
  __device__ __forceinline__ uint32_t scan_chunk(const TileCtx &T, uint32_t cbase, uint32_t lane, uint32_t stage,
                                                 uint32_t cap, uint32_t q1, uint32_t q2,
                                                 unsigned long long out_base, uint32_t *overflow) const {
    uint32_t cg, cp;
    const uint32_t lpos = cbase + lane * 16;
    stage1(T, lpos, cg, cp);
    const uint32_t valid = lpos >= T.nscan ? 0u : (T.nscan - lpos >= 16 ? 0xFFFFu : ((1u << (T.nscan - lpos)) - 1u));
    uint32_t cand = (cg | cp) & valid;
    const uint32_t cnt = __popc(cand);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    if (total == 0) return 0;
    {
      uint32_t o = incl - cnt;
      while (cand) {
        const uint32_t k = __ffs(cand) - 1;
        cand &= cand - 1;
        sts16(q1 + 2u * o++, (lane * 16 + k) | (((cg >> k) & 1u) << 9) | (((cp >> k) & 1u) << 10));
      }
    }
    __syncwarp();
    uint32_t found = 0, q2n = 0;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t base = 0; base < total; base += 32 * kProbeUnroll) {
      Probe pr[kProbeUnroll];
#pragma unroll
      for (int u = 0; u < kProbeUnroll; ++u) {
        const uint32_t idx = base + u * 32 + lane;
        const bool ok = idx < total;
        const uint32_t e = ok ? lds16(q1 + 2u * idx) : 0u;
        probe_issue(T, ok, cbase + (e & 511u), (e >> 9) & 1u, (e >> 10) & 1u, pr[u]);
      }
#pragma unroll
      for (int u = 0; u < kProbeUnroll; ++u) {
        if (base + u * 32 >= total) break;
        uint32_t slot;
        const bool want = probe_key(pr[u], &slot);
        const uint32_t bal = __ballot_sync(kFull, want);
        if (want)
          sts64u(q2 + 8u * (q2n + __popc(bal & lt)),
                 ((unsigned long long)(pr[u].tpos | ((pr[u].flags & 12u) << 14)) << 32) | slot);
        q2n += __popc(bal);
      }
      __syncwarp();
      // ONE verify_batch in the instruction stream (the kernel is I-cache bound): it runs when 32
      // entries are waiting, and drains the queue after the last candidates of the chunk
      const bool last = base + 32 * kProbeUnroll >= total;
      while (q2n >= 32 || (last && q2n)) {
        const uint32_t nb = q2n < 32 ? q2n : 32;
        found += verify_batch<MODE>(T, q2, nb, lane, stage, found, cap, cbase, out_base, overflow);
        const uint32_t rest = q2n - nb; // <= 63
        const unsigned long long mv0 = lane < rest ? lds64u(q2 + 8u * (32 + lane)) : 0ull;
        const unsigned long long mv1 = 32 + lane < rest ? lds64u(q2 + 8u * (64 + lane)) : 0ull;
        __syncwarp();
        if (lane < rest) sts64u(q2 + 8u * lane, mv0);
        if (32 + lane < rest) sts64u(q2 + 8u * (32 + lane), mv1);
        __syncwarp();
        q2n = rest;
      }
    }
    __syncwarp();
    return found;
  }

  // The same chunk for the common case -- no position predicate requested, no 1..3 byte
  // patterns -- written against raw shared-memory offsets with nothing but the essentials in
  // the per-candidate rounds (the generic scan_chunk above spends ~3x the instructions there).
  //   sb_off/g4_off/p23_off/q1_off/q2_off/stage: shared-space byte addresses.
  //   PRED: a start predicate is requested (word_boundary / word_prefix / line_start): it is evaluated
  //   per candidate in front of the key probe, exactly as probe_issue does (matcher.c:770-776, :195-196,
  //   :806-807); the end predicates live in verify() for both paths.  Exact statistics (kCountAll) need
  //   the candidates a start predicate rejects and stay on the generic path.
  template <int MODE, bool PRED>
  __device__ __forceinline__ uint32_t scan_chunk_fast(const TileCtx &T, uint32_t sb_off, uint32_t g4_off,
                                                      uint32_t p23_off, uint32_t q1_off, uint32_t q2_off,
                                                      uint32_t cbase, uint32_t lane, uint32_t stage, uint32_t cap,
                                                      unsigned long long out_base, uint32_t *overflow) const {
    const uint32_t lpos = cbase + lane * 16;
    const uint32_t src = sb_off + kTilePre + lpos;
    const uint4 v = lds128(src);
    uint32_t cand = 0, cp = 0; // cp: candidates of the 1..3 byte patterns (HAS_P23)
    if (HAS_CLS) {
      cand = class_runs16(v, lds64(src + 16));
    } else {
      const uint2 w45 = lds64(src + 16);
      const uint32_t w[6] = {v.x, v.y, v.z, v.w, w45.x, w45.y};
      const uint32_t sh = P.st.g4_shift, tmask = P.st.tail_mask;
      const uint32_t pand = P.st.p23_and, pmul = P.st.p23_mul, psh = P.st.p23_shift;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t gram = __byte_perm(w[k >> 2], w[(k >> 2) + 1], 0x0123u + 0x1111u * (k & 3));
        if (HAS_G4) {
          // (keys longer than 4 bytes: the little-endian word of bytes k+4..k+7 joins the hash;
          // stores with 1..3 byte patterns always have 4-byte keys)
          uint32_t hk = gram * kHashMul;
          if (!HAS_P23) {
            const uint32_t tail = __byte_perm(w[(k >> 2) + 1], w[(k >> 2) + 2 > 5 ? 5 : (k >> 2) + 2], 0x3210u + 0x1111u * (k & 3));
            hk ^= (tail & tmask) * kHashMul2;
          }
          const uint32_t b = hk >> sh;
          const uint32_t word = lds32(g4_off + ((b >> 5) << 2));
          cand |= ((word >> (b & 31)) & 1u) << k;
        }
        if (HAS_P23) {
          const uint32_t b = ((gram & pand) * pmul) >> psh;
          const uint32_t word = lds32(p23_off + ((b >> 5) << 2));
          cp |= ((word >> (b & 31)) & 1u) << k;
        }
      }
    }
    const uint32_t rem_c = T.rem0 - cbase; // >= 1
    const uint32_t kbytes = P.st.key_bytes;
    if (rem_c < (uint32_t)kChunkBytes + 8u) {
      // only the segment's last chunks: a position with fewer than key_bytes bytes left is no key
      // candidate (no pattern behind a key is shorter than the key: matcher.c:203 would reject them all)
      const uint32_t fit = rem_c >= kbytes ? rem_c - kbytes + 1u : 0u, lrel = lane * 16;
      cand &= lrel >= fit ? 0u : (fit - lrel >= 16u ? 0xFFFFu : ((1u << (fit - lrel)) - 1u));
    }
    const uint32_t cg = cand; // gram candidates
    cand |= cp;
    if (lpos + 16 > T.nscan) cand &= lpos >= T.nscan ? 0u : ((1u << (T.nscan - lpos)) - 1u);
    const uint32_t cnt = __popc(cand);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    if (total == 0) return 0;
    {
      uint32_t qa = q1_off + 2u * (incl - cnt);
      if (__any_sync(kFull, cnt > 4)) { // dense: 16 predicated steps, no branches
        const uint32_t eb = lane * 16;
#pragma unroll
        for (int k = 0; k < 16; ++k)
          if ((cand >> k) & 1u) { // entry: position | gram candidate << 9 | short candidate << 10
            uint32_t ent = eb + k;
            if (HAS_P23) ent |= (((cg >> k) & 1u) << 9) | (((cp >> k) & 1u) << 10);
            sts16(qa, ent);
            qa += 2;
          }
      } else {
        const uint32_t eb = lane * 16;
        while (cand) {
          const uint32_t k = __ffs(cand) - 1;
          uint32_t ent = eb + k;
          if (HAS_P23) ent |= (((cg >> k) & 1u) << 9) | (((cp >> k) & 1u) << 10);
          sts16(qa, ent);
          cand &= cand - 1;
          qa += 2;
        }
      }
    }
    __syncwarp();
    const uint32_t q2 = q2_off;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t tile_off = sb_off + kTilePre + cbase;
    const uint32_t key_shift = P.st.key_shift, g4_shift = P.st.g4_shift, empty = P.st.empty_key;
    const uint4 *keys = P.st.keys;
    const uint32_t tmask = P.st.tail_mask;
    uint32_t found = 0, q2n = 0;
    constexpr int U = OLM_FAST_UNROLL;
    for (uint32_t base = 0; base < total; base += 32 * U) {
      uint32_t e[U], gram[U], bucket[U];
      bool pass[U], shortc[U];
      uint4 kb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t idx = base + u * 32 + lane;
        pass[u] = idx < total;
        e[u] = pass[u] ? lds16(q1_off + 2u * idx) : 0u;
        shortc[u] = false;
        if (HAS_P23) { // unpack the flags; without a gram candidate there is no key probe
          shortc[u] = (e[u] >> 10) & 1u;
          pass[u] = pass[u] && ((e[u] >> 9) & 1u);
          e[u] &= 511u;
        }
        const uint32_t a = tile_off + e[u], a4 = a & ~3u, sh8 = a << 3;
        const uint32_t x0 = lds32(a4), x1 = lds32(a4 + 4); // bytes 0..7 of the position from three aligned words
        const uint32_t gbe = __byte_perm(__funnelshift_r(x0, x1, sh8), 0, 0x0123);
        if (PRED) {
          const bool at0 = T.first && cbase + e[u] == 0;
          const uint32_t prev = lds8(a - 1), cur = gbe >> 24;
          bool ok = true;
          if (fl & kWordBoundary) ok = is_word_byte(cur) != (at0 ? false : is_word_byte(prev)); // matcher.c:770-776
          if ((fl & kWordPrefix) && !at0 && is_word_byte(prev)) ok = false;                       // :195, :806
          if ((fl & kLineStart) && !at0 && !is_line_end_byte(prev)) ok = false;                   // :196, :807
          pass[u] = pass[u] && ok;
          shortc[u] = shortc[u] && ok;
        }
        if (SX && __any_sync(kFull, shortc[u])) shortc[u] = shortc[u] && short_look(gbe);
        uint32_t h = gbe * kHashMul;
        if (tmask) h ^= (__funnelshift_r(x1, lds32(a4 + 8), sh8) & tmask) * kHashMul2; // keys longer than 4 bytes
        gram[u] = h; // the key of the position
        if (HAS_CLS && !OLM_CLS_SKIP_BITMAP) {
          const uint32_t b = h >> g4_shift;
          const uint32_t word = lds32(g4_off + ((b >> 5) << 2));
          pass[u] = pass[u] && ((word >> (b & 31)) & 1u);
        }
        bucket[u] = h >> key_shift;
        if (HAS_G4 && pass[u]) kb[u] = ld_keys(keys + bucket[u]); // (not read when !pass[u])
      }
      static_assert(U == 2, "Q2 holds 31 + 2 x 32 entries");
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u > 0 && base + u * 32 >= total) break;
        const uint32_t g = gram[u];
        bool hit = HAS_G4 && pass[u] && (kb[u].x == g || kb[u].y == g || kb[u].z == g || kb[u].w == g);
        // rare: the home bucket is full and does not hold the key -> next bucket(s)
        const bool more = HAS_G4 && pass[u] && !hit && kb[u].w != empty;
        // at 1 M patterns nearly every round ends here: no key hit, no full bucket -- one vote
        if (!__any_sync(kFull, hit || more || (HAS_P23 && shortc[u]))) continue;
        if (HAS_G4 && __any_sync(kFull, more)) {
          while (pass[u] && !hit && kb[u].w != empty) {
            bucket[u] = (bucket[u] + 1) & P.st.key_mask;
            kb[u] = ld_keys(keys + bucket[u]);
            hit = kb[u].x == g || kb[u].y == g || kb[u].z == g || kb[u].w == g;
          }
        }
        const bool want = hit || (HAS_P23 && shortc[u]);
        const uint32_t bal = __ballot_sync(kFull, want);
        if (!bal) continue;
        // (branch-free: the slot of a miss is kNoSlot, the store is predicated)
        const uint32_t place = kb[u].y == g ? 1u : kb[u].z == g ? 2u : kb[u].w == g ? 3u : 0u;
        const uint32_t slot = hit ? 4u * bucket[u] + place : kNoSlot;
        const uint32_t hi32 = (cbase + e[u]) | (HAS_P23 && shortc[u] ? 0x10000u : 0u);
        if (want) sts64u(q2 + 8u * (q2n + __popc(bal & lt)), ((unsigned long long)hi32 << 32) | slot);
        q2n += __popc(bal);
      }
      __syncwarp();
      // ONE verify_batch in the instruction stream: it runs when 32 entries are waiting, and
      // drains the queue after the last candidates of the chunk
      const bool last = base + 32 * U >= total;
      while (q2n >= 32 || (last && q2n)) {
        const uint32_t nb = q2n < 32 ? q2n : 32;
        found += verify_batch<MODE>(T, q2, nb, lane, stage, found, cap, cbase, out_base, overflow);
        const uint32_t rest = q2n - nb; // <= 63
        const unsigned long long mv0 = lane < rest ? lds64u(q2 + 8u * (32 + lane)) : 0ull;
        const unsigned long long mv1 = 32 + lane < rest ? lds64u(q2 + 8u * (64 + lane)) : 0ull;
        __syncwarp();
        if (lane < rest) sts64u(q2 + 8u * lane, mv0);
        if (32 + lane < rest) sts64u(q2 + 8u * (32 + lane), mv1);
        __syncwarp();
        q2n = rest;
      }
    }
    __syncwarp();
    return found;
  }

  __device__ __forceinline__ void flush_stats(uint32_t lane) const {
    unsigned long long h = n_hits, mi = n_miss, cm = n_cmp, lh = n_long_hits;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      h += __shfl_xor_sync(kFull, h, d);
      mi += __shfl_xor_sync(kFull, mi, d);
      cm += __shfl_xor_sync(kFull, cm, d);
      lh += __shfl_xor_sync(kFull, lh, d);
    }
    if (lane == 0 && P.counters) {
      if (h) atomicAdd(P.counters + 0, h);
      if (mi) atomicAdd(P.counters + 1, mi);
      if (cm) atomicAdd(P.counters + 2, cm);
      if (lh) atomicAdd(P.counters + 3, lh);
    }
  }
};

// Starts the bulk copy of tile `I` into `dst` (a stage buffer), completing on `bar`.
__device__ __forceinline__ void copy_tile(const ScanParams &P, const StageInfo &I, uint8_t *dst, uint64_t *bar) {
  if (I.nscan) {
    const long long pre = I.boff >= kTilePre ? kTilePre : 0;
    const uint32_t bytes = (uint32_t)(I.staged + pre);
    mbar_expect_tx(bar, bytes);
    tma_load_1d(dst + (kTilePre - pre), P.buf + (I.boff - pre), bytes, bar);
  } else {
    mbar_expect_tx(bar, 0); // empty tile: the phase completes at once
  }
}
struct SmemLayout {
  SmemHeader *H;
  uint8_t *ring;
  uint32_t *g4s, *p23s, *sxs, *staging;
  uint16_t *q1;
  unsigned long long *q2;
  uint8_t *priv, *xf; // per warp: private chunk buffer (kPrivBytes), rows of a normalised chunk (kXfRowBytes)
};
// header | ring | g4 | p23 | sx | Q2 (8-byte entries) | staging | Q1 | private chunk buffers | rows
template <bool HAS_G4, bool HAS_P23>
__device__ __forceinline__ SmemLayout carve(uint8_t *smem, const ScanParams &P, size_t ring_bytes, uint32_t staging_words) {
  SmemLayout L;
  L.H = reinterpret_cast<SmemHeader *>(smem);
  L.ring = smem + kSmemHeader;
  L.g4s = reinterpret_cast<uint32_t *>(L.ring + ring_bytes);
  L.p23s = L.g4s + (HAS_G4 ? P.st.g4_words : 0);
  L.sxs = L.p23s + (HAS_P23 ? P.st.p23_words : 0);
  L.q2 = reinterpret_cast<unsigned long long *>(L.sxs + (HAS_P23 ? P.st.sx_words : 0));
  L.staging = reinterpret_cast<uint32_t *>(L.q2 + kScanWarps * kQ2Entries);
  L.q1 = reinterpret_cast<uint16_t *>(L.staging + staging_words);
  L.priv = reinterpret_cast<uint8_t *>(L.q1) + kQ1Bytes;
  L.xf = L.priv + (size_t)kScanWarps * kPrivBytes;
  return L;
}
// the fields of a tile description the scanning warps need, from shared memory
__device__ __forceinline__ void load_info(uint32_t I32, StageInfo &I) {
  const uint2 gb = lds64(I32 + (uint32_t)offsetof(StageInfo, gbase));
  const uint2 p0 = lds64(I32 + (uint32_t)offsetof(StageInfo, p0));
  const uint2 bo = lds64(I32 + (uint32_t)offsetof(StageInfo, boff));
  const uint2 rn = lds64(I32 + (uint32_t)offsetof(StageInfo, rem0));
  I.gbase = ((unsigned long long)gb.y << 32) | gb.x;
  I.p0 = ((unsigned long long)p0.y << 32) | p0.x;
  I.boff = (long long)(((unsigned long long)bo.y << 32) | bo.x);
  I.rem0 = rn.x;
  I.nscan = rn.y;
  const uint2 tw = lds64(I32 + (uint32_t)offsetof(StageInfo, tail));
  I.tail = tw.x;
  I.win = tw.y;
  I.staged = lds32(I32 + (uint32_t)offsetof(StageInfo, staged));
}
// T = the whole tile in its stage buffer (plain stores scanned in place)
__device__ __forceinline__ void tile_ctx(const StageInfo &I, uint32_t sb32, TileCtx &T) {
  T.sb32 = sb32;
  T.gbase = I.gbase;
  T.boff = I.boff;
  T.rem0 = I.rem0;
  T.nscan = I.nscan;
  T.staged = I.staged;
  T.tail = I.tail;
  T.first = I.p0 == 0;
}
// T = chunk `cbase` of tile I in the warp's private buffer, as the matcher has to see it
// `back` = buffer bytes in front of the chunk's first byte that are in shared memory in front of src32
template <bool XF>
__device__ __forceinline__ void build_chunk(const ScanParams &P, const StageInfo &I, uint32_t src32, uint32_t cbase,
                                            uint32_t back, uint32_t priv32, uint32_t xf32, uint32_t lane, bool first_pass,
                                            TileCtx &T) {
  if (XF && xf32) {
    build_xf(P, I, src32, cbase, back, priv32, xf32, lane, first_pass, T);
  } else {
    build_copy<XF>(I, src32, cbase, priv32, lane, T);
  }
  __syncwarp();
}
template <bool HAS_G4, bool HAS_P23>
__device__ __forceinline__ void load_filters(const SmemLayout &L, const ScanParams &P, uint32_t tid, uint32_t nthreads) {
  if (HAS_G4) {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.st.g4);
    uint4 *dst = reinterpret_cast<uint4 *>(L.g4s);
    for (uint32_t i = tid; i < P.st.g4_words / 4; i += nthreads) dst[i] = __ldg(src + i);
  }
  if (HAS_P23) {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.st.p23);
    uint4 *dst = reinterpret_cast<uint4 *>(L.p23s);
    for (uint32_t i = tid; i < P.st.p23_words / 4; i += nthreads) dst[i] = __ldg(src + i);
    const uint4 *src2 = reinterpret_cast<const uint4 *>(P.st.sx);
    uint4 *dst2 = reinterpret_cast<uint4 *>(L.sxs);
    for (uint32_t i = tid; i < P.st.sx_words / 4; i += nthreads) dst2[i] = __ldg(src2 + i);
  }
}

__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t *p) {
  return *reinterpret_cast<const volatile uint32_t *>(p);
}

// FAST: 0 = the generic per-candidate path, 1 = the lean path, 2 = the lean path with start predicates
template <bool HAS_G4, bool HAS_P23, bool HAS_CLS, int FAST, bool XF, bool COOP, bool SX = false>
__global__ void __launch_bounds__(kScanThreads, 1) scan_kernel(const __grid_constant__ ScanParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t S = P.stages, cap = P.chunk_cap;
  const SmemLayout L = carve<HAS_G4, HAS_P23>(smem, P, (size_t)S * kStageBytes, kScanWarps * cap);
  SmemHeader &H = *L.H;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  load_filters<HAS_G4, HAS_P23>(L, P, tid, kScanThreads);
  if (tid == 0) {
    for (uint32_t s = 0; s < kMaxStages; ++s) {
      mbar_init(&H.full[s], 1);
      mbar_init(&H.scanned[s], kTileChunks);
    }
    H.chunk_ctr = 0;
    H.end_k = kNoTile;
    for (uint32_t i = 0; i < kInfoRing; ++i) {
      H.info[i].seq = kNoTile;
      mbar_init(&H.described[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kScanWarps) {
    // ============================ producer warp (one lane) ============================
    if (lane != 0) return;
    // Tile iteration k of this CTA -> stage k % S.  `seq` is stored first so that a scanning
    // warp can tell that the stage's mbarrier is in ITS generation before it waits on it.
    // (the ring depth S is any number >= 2: stage and phase parity of an iteration are counted
    // here and handed to the scanning warps through the tile description)
    uint32_t ps = 0, ppar = 0;
    // tickets: groups of `batch` consecutive tiles; the next group's atomic is in flight while this one is handed out
    // (small launches: single tiles, so that no SM is left without work)
    uint32_t batch = P.num_tiles / (gridDim.x * 16u);
    batch = batch < 1u ? 1u : batch > (uint32_t)OLM_TICKET_BATCH ? (uint32_t)OLM_TICKET_BATCH : batch;
    uint32_t t_next = 0, t_left = 0, t_pref = atomicAdd(P.ticket, batch);
    auto produce = [&](uint32_t k) {
      const uint32_t s = ps, par = ppar;
      if (++ps == S) {
        ps = 0;
        ppar ^= 1u;
      }
      if (t_left == 0) {
        t_next = t_pref;
        t_left = batch;
        if (t_next < P.num_tiles) t_pref = atomicAdd(P.ticket, batch);
      }
      const uint32_t t = t_next++;
      --t_left;
      StageInfo &I = H.info[k % kInfoRing];
      I.stage_par = s | (par << 16);
      if (t >= P.num_tiles) {
        I.tile = kNoTile;
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t *>(&I.seq) = k;
        if (k < H.end_k) *reinterpret_cast<volatile uint32_t *>(&H.end_k) = k;
        mbar_expect_tx(&H.full[s], 0);
        if (OLM_DESC_BAR) mbar_arrive(&H.described[k % kInfoRing]);
        return;
      }
      fill_tile(P, t, I);
      __threadfence_block();
      *reinterpret_cast<volatile uint32_t *>(&I.seq) = k;
      if (OLM_DESC_BAR) mbar_arrive(&H.described[k % kInfoRing]);
      copy_tile(P, I, L.ring + (size_t)s * kStageBytes, &H.full[s]);
    };
    // tiles claimed ahead: the ring depth, but not more than this CTA's fair share (small inputs)
    const uint32_t share = (P.num_tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t D = share < S ? (share ? share : 1u) : S;
    for (uint32_t k = 0; k < D; ++k) produce(k);
    uint32_t s = 0, ph = 0, k = 0;
    for (;; ++k) {
      if (H.info[k % kInfoRing].tile == kNoTile) break;
      mbar_wait(&H.scanned[s], ph);
      produce(k + D); // stage (k + D) % S last held tile k + D - S <= k: free; so is the info entry
      if (++s == S) {
        s = 0;
        ph ^= 1;
      }
    }
    if (OLM_DESC_BAR) {
      // Iteration k is the first without a tile, and every tile before it has been scanned.  The
      // scanning warps each take one more chunk before they see that: describe the iterations those
      // chunks fall into as empty too (k .. k + D - 1 already are), so that nobody sleeps on a
      // description that never comes.
      constexpr uint32_t kBeyondEnd = (kScanWarps + kTileChunks - 1) / kTileChunks + 2;
      for (uint32_t j = k + D; j < k + D + kBeyondEnd; ++j) {
        H.info[j % kInfoRing].tile = kNoTile;
        mbar_arrive(&H.described[j % kInfoRing]);
      }
    }
    return;
  }

  // ================================ scanning warps ================================
  // (everything in shared memory is addressed by 32-bit shared-space addresses from here on)
  Scanner<HAS_G4, HAS_P23, HAS_CLS, XF, COOP, SX> sc(P, L.g4s, L.p23s);
  if (SX) sc.sx32 = smem_u32(L.sxs);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t ring32 = sbase + (uint32_t)(L.ring - smem);
  const uint32_t g4_32 = sbase + (uint32_t)(reinterpret_cast<uint8_t *>(L.g4s) - smem);
  const uint32_t p23_32 = sbase + (uint32_t)(reinterpret_cast<uint8_t *>(L.p23s) - smem);
  const uint32_t q1_32 = sbase + (uint32_t)(reinterpret_cast<uint8_t *>(L.q1 + warp * kChunkBytes) - smem);
  const uint32_t q2_32 = sbase + (uint32_t)(reinterpret_cast<uint8_t *>(L.q2 + warp * kQ2Entries) - smem);
  const uint32_t stage32 = sbase + (uint32_t)(reinterpret_cast<uint8_t *>(L.staging + (size_t)warp * cap) - smem);
  const uint32_t hdr32 = sbase; // SmemHeader
  const uint32_t full32 = hdr32 + (uint32_t)offsetof(SmemHeader, full);
  const uint32_t scanned32 = hdr32 + (uint32_t)offsetof(SmemHeader, scanned);
  const uint32_t ctr32 = hdr32 + (uint32_t)offsetof(SmemHeader, chunk_ctr);
  const uint32_t endk32 = hdr32 + (uint32_t)offsetof(SmemHeader, end_k);
  const uint32_t desc32 = hdr32 + (uint32_t)offsetof(SmemHeader, described);
  const uint32_t info32 = hdr32 + (uint32_t)offsetof(SmemHeader, info);
  // plain stores are scanned in the stage buffer; stores with a transform flag in the warp's private
  // buffer, which is where case folding / normalisation happen (scan_device.cuh) -- the stage is
  // handed back before the scan then.  (Private copies for plain stores were measured too: 555 vs
  // 581 GB/s at 1 M patterns, 316 vs 336 on the short-pattern store, 338 vs 322 on names.txt.)
  constexpr bool priv_mode = XF;
  const uint32_t priv32 = sbase + (uint32_t)(L.priv - smem) + warp * (uint32_t)kPrivBytes;
  if (XF && !(P.flags & kIdentityMap)) sc.xf32 = sbase + (uint32_t)(L.xf - smem) + warp * (uint32_t)kXfRowBytes;
  unsigned long long blk_next = 0; // this warp's block of temp[]: next free entry ...
  uint32_t blk_left = 0;           // ... and how many are left
  for (;;) {
    uint32_t c = 0;
    if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(c) : "r"(ctr32), "n"(kGrab) : "memory");
    c = __shfl_sync(kFull, c, 0);
    const uint32_t k = c / kTileChunks, ci0 = c % kTileChunks;
    const uint32_t I32 = info32 + (k % kInfoRing) * (uint32_t)sizeof(StageInfo);
#if OLM_DESC_BAR
    // Wait for the description of iteration k.  A warp sleeps here (at the first iteration that is not
    // described yet), so no warp is ever ahead of the producer, and the producer is at most S < kInfoRing
    // iterations ahead of the oldest chunk in flight: the barrier of entry k % kInfoRing is in the phase of
    // iteration k or of k - kInfoRing, which the parity tells apart.
    mbar_wait32(desc32 + 8u * (k % kInfoRing), (k / kInfoRing) & 1u);
    const uint32_t tile = lds32(I32 + (uint32_t)offsetof(StageInfo, tile));
    if (tile == kNoTile) break;
    const uint32_t sp = lds32(I32 + (uint32_t)offsetof(StageInfo, stage_par));
    const uint32_t s = sp & 0xFFFFu;
    mbar_wait32(full32 + 8u * s, sp >> 16);
#else
    // The mbarrier only tells two phases apart: make sure the stage is in OUR generation first.
    // (Chunks past the CTA's last tile may belong to an iteration that is never produced.)
    bool over = false;
    while (lds32(I32 + (uint32_t)offsetof(StageInfo, seq)) != k) {
      if (k >= lds32(endk32)) {
        over = true;
        break;
      }
      __nanosleep(32);
    }
    if (over) break;
    const uint32_t sp = lds32(I32 + (uint32_t)offsetof(StageInfo, stage_par));
    const uint32_t s = sp & 0xFFFFu;
    mbar_wait32(full32 + 8u * s, sp >> 16);
    const uint32_t tile = lds32(I32 + (uint32_t)offsetof(StageInfo, tile));
    if (tile == kNoTile) break;
#endif
    StageInfo I;
    load_info(I32, I);
    const uint32_t stage_sb = ring32 + s * (uint32_t)kStageBytes;
    // (experiment knob OLM_GRAB: the kGrab chunks of a grab share the wait and the tile description)
#if OLM_GRAB > 1
#pragma unroll 1
    for (uint32_t ci = ci0; ci < ci0 + kGrab; ++ci) {
#else
    {
    const uint32_t ci = ci0;
#endif
    const uint32_t cbase = ci * kChunkBytes;
    uint32_t n = 0, ovf = 0;
    TileCtx T;
    uint32_t cb = cbase; // position of the chunk inside T
    bool work = cbase < I.nscan;
    if (priv_mode) {
      // (the whole tile is in the stage: everything of it in front of the chunk, and the 16 bytes in front of the tile)
      if (work)
        build_chunk<XF>(P, I, stage_sb + kTilePre + cbase, cbase, cbase + (I.boff >= (long long)kTilePre ? (uint32_t)kTilePre : 0u),
                        priv32, sc.xf32, lane, true, T);
      __syncwarp();
      if (lane == 0) mbar_arrive32(scanned32 + 8u * s); // the stage buffer is not read any more
      cb = 0;
      work = work && T.nscan != 0;
    } else {
      tile_ctx(I, stage_sb, T);
    }
    if (work) {
      if (FAST)
        n = sc.template scan_chunk_fast<kStageMode, FAST == 2>(T, T.sb32, g4_32, p23_32, q1_32, q2_32, cb, lane, stage32, cap, 0, &ovf);
      else
        n = sc.template scan_chunk<kStageMode>(T, cb, lane, stage32, cap, q1_32, q2_32, 0, &ovf);
    }
    __syncwarp();
    if (!priv_mode && lane == 0) mbar_arrive32(scanned32 + 8u * s); // the stage buffer is not read any more
    // ---- hand the chunk over: descriptor, and the staged matches into this warp's run of temp[]
    ovf = __any_sync(kFull, ovf != 0);
    ChunkDesc d;
    d.count = n;
    d.temp_index = 0;
    if (ovf) {
      d.count = n | kChunkOverflow;
      if (lane == 0) *reinterpret_cast<volatile unsigned int *>(P.redo_flag) = 1u;
    } else if (n) {
      if (n > blk_left) { // a fresh block (what is left of the old one is lost)
        const uint32_t want = n > kWarpTempBlock / 2 ? 2 * n : kWarpTempBlock;
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(P.temp_count, (unsigned long long)want);
        blk_next = __shfl_sync(kFull, b, 0);
        blk_left = want;
      }
      // (when temp[] is too small the entries are dropped; the host sees it and repeats the call)
      if (blk_next + n <= P.temp_cap && blk_next + n <= 0xFFFFFFFFull) {
        uint32_t *dst = P.temp + blk_next;
        for (uint32_t i = lane; i < n; i += 32) dst[i] = lds32(stage32 + 4u * i);
      }
      d.temp_index = (uint32_t)blk_next;
      blk_next += n;
      blk_left -= n;
    }
    if (lane == 0)
      *reinterpret_cast<uint2 *>(P.chunk_desc + ((size_t)tile * kTileChunks + ci)) = make_uint2(d.count, d.temp_index);
    __syncwarp(); // the staging area is rewritten by the next chunk
    }
  }
  sc.flush_stats(lane);
}

// Chunks flagged kChunkOverflow (their matches did not fit the staging area): every warp takes
// such chunks, reads the chunk's bytes (16 in front, kPrivData from its first byte on) from global
// memory into a small stage of its own, builds its private buffer from that exactly like the main
// pass, evaluates the chunk again and writes final records at the index the prefix pass computed.
constexpr int kRedoStage = kTilePre + kPrivData;  // 656 bytes per warp
constexpr size_t kRedoRingBytes = ((size_t)kScanWarps * kRedoStage + 127) & ~size_t(127);
template <bool HAS_G4, bool HAS_P23, bool HAS_CLS, bool XF, bool COOP>
__global__ void __launch_bounds__(kScanThreads, 1) redo_kernel(const __grid_constant__ ScanParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (*P.redo_flag == 0) return;
  // header | one small stage per warp | g4 | p23 | Q2 | Q1 | private chunk buffers | rows
  const SmemLayout L = carve<HAS_G4, HAS_P23>(smem, P, kRedoRingBytes, 0);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  load_filters<HAS_G4, HAS_P23>(L, P, tid, kScanThreads);
  __syncthreads();
  if (warp >= kScanWarps) return;
  Scanner<HAS_G4, HAS_P23, HAS_CLS, XF, COOP> sc(P, L.g4s, L.p23s); // (without the second look: it only spares verifies)
  sc.stat_inc = 0; // the main pass has counted these chunks already
  const uint32_t q1_32 = smem_u32(L.q1 + warp * kChunkBytes), q2_32 = smem_u32(L.q2 + warp * kQ2Entries);
  uint8_t *buf = L.ring + (size_t)warp * kRedoStage;
  const uint32_t priv32 = smem_u32(L.priv) + warp * (uint32_t)kPrivBytes;
  if (XF && !(P.flags & kIdentityMap)) sc.xf32 = smem_u32(L.xf) + warp * (uint32_t)kXfRowBytes;
  const uint64_t n_chunks = (uint64_t)P.num_tiles * kTileChunks;
  // a warp looks at 32 descriptors at a time and takes the flagged chunks one after the other
  for (uint64_t c0 = ((uint64_t)blockIdx.x * kScanWarps + warp) * 32; c0 < n_chunks;
       c0 += (uint64_t)gridDim.x * kScanWarps * 32) {
    uint32_t flagged = 0;
    if (c0 + lane < n_chunks) flagged = P.chunk_desc[c0 + lane].count & kChunkOverflow;
    uint32_t todo = __ballot_sync(kFull, flagged != 0);
    while (todo) {
      const uint64_t ch = c0 + (__ffs(todo) - 1);
      todo &= todo - 1;
      const uint32_t tile = (uint32_t)(ch / kTileChunks), ci = (uint32_t)(ch % kTileChunks);
      StageInfo I;
      fill_tile(P, tile, I);
      const uint32_t cbase = ci * kChunkBytes;
      { // what the tile's stage buffer would hold around the chunk
        const long long first = I.boff + cbase - kTilePre;
        uint32_t have = I.staged - cbase;
        if (have > (uint32_t)kPrivData) have = kPrivData;
        for (uint32_t i = lane; i < have + kTilePre; i += 32) buf[i] = first + (long long)i >= 0 ? P.buf[first + (long long)i] : 0;
      }
      __syncwarp();
      TileCtx T;
      // (the small stage holds 16 bytes in front of the chunk; what lies further back is read from global memory)
      build_chunk<XF>(P, I, smem_u32(buf) + kTilePre, cbase, I.boff + (long long)cbase >= (long long)kTilePre ? (uint32_t)kTilePre : 0u,
                      priv32, sc.xf32, lane, false, T);
      // first result index: the span's base + the chunks before this one in the span
      unsigned long long base = P.span_base[ch / kPrefixSpan];
      {
        const uint64_t s0 = ch - ch % kPrefixSpan;
        unsigned long long part = 0;
        for (uint64_t j = s0 + lane; j < ch; j += 32) part += P.chunk_desc[j].count & ~kChunkOverflow;
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) part += __shfl_xor_sync(kFull, part, dd);
        base += part;
      }
      if (T.nscan) sc.template scan_chunk<kDirectMode>(T, 0, lane, 0u, 0, q1_32, q2_32, base, nullptr);
      __syncwarp();
    }
  }
}

// ---- prefix over the chunk counts of a launch -------------------------------------------------
// span_base[b] <- matches in span b (kPrefixSpan chunks)
constexpr int kPrefixThreads = 256;
__global__ void __launch_bounds__(kPrefixThreads) prefix_sum_kernel(const __grid_constant__ ScanParams P) {
  __shared__ unsigned long long s_warp[kPrefixThreads / 32];
  const uint64_t n_chunks = (uint64_t)P.num_tiles * kTileChunks;
  const uint64_t c0 = (uint64_t)blockIdx.x * kPrefixSpan;
  unsigned long long sum = 0;
  for (uint32_t i = threadIdx.x; i < kPrefixSpan; i += kPrefixThreads)
    if (c0 + i < n_chunks) sum += P.chunk_desc[c0 + i].count & ~kChunkOverflow;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(kFull, sum, d);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kPrefixThreads / 32; ++w) t += s_warp[w];
    P.span_base[blockIdx.x] = t;
  }
}
// span_base[b] <- *P.total + sum of the spans before b; *P.total += everything (one CTA)
__global__ void __launch_bounds__(1024, 1) prefix_scan_kernel(const __grid_constant__ ScanParams P, uint32_t n_spans) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = *P.total;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < n_spans; b0 += 1024) {
    const unsigned long long mine = b0 + tid < n_spans ? P.span_base[b0 + tid] : 0ull;
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
    if (b0 + tid < n_spans) P.span_base[b0 + tid] = before + incl - mine;
    __syncthreads();
    if (tid == 1023) s_carry = before + incl;
    __syncthreads();
  }
  if (tid == 0) *P.total = s_carry;
}

// Packed entries of temp[] -> final records.  One CTA per span of kPrefixSpan chunks: exclusive
// prefix of the counts inside the span (16 chunks per thread), then the warps copy the chunks
// that have matches, 32 chunks per warp at a time.
__global__ void __launch_bounds__(kPrefixThreads) place_kernel(const __grid_constant__ ScanParams P) {
  __shared__ unsigned long long s_warp[kPrefixThreads / 32];
  __shared__ uint32_t s_pre[kPrefixSpan]; // exclusive prefix inside the span
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t n_chunks = (uint64_t)P.num_tiles * kTileChunks;
  const uint64_t c0 = (uint64_t)blockIdx.x * kPrefixSpan;
  constexpr uint32_t kPer = kPrefixSpan / kPrefixThreads; // 16 consecutive chunks per thread
  uint32_t cnt[kPer];
  uint32_t mine = 0;
#pragma unroll
  for (uint32_t i = 0; i < kPer; ++i) {
    const uint64_t ch = c0 + (uint64_t)tid * kPer + i;
    cnt[i] = ch < n_chunks ? (P.chunk_desc[ch].count & ~kChunkOverflow) : 0u;
    mine += cnt[i];
  }
  uint32_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t run = incl - mine;
  for (uint32_t w = 0; w < warp; ++w) run += (uint32_t)s_warp[w];
#pragma unroll
  for (uint32_t i = 0; i < kPer; ++i) {
    s_pre[tid * kPer + i] = run;
    run += cnt[i];
  }
  __syncthreads();
  const unsigned long long span0 = P.span_base[blockIdx.x];
  for (uint32_t g0 = warp * 32; g0 < kPrefixSpan; g0 += (kPrefixThreads / 32) * 32) {
    const uint64_t ch = c0 + g0 + lane;
    uint2 d = make_uint2(0, 0);
    if (ch < n_chunks) d = *reinterpret_cast<const uint2 *>(P.chunk_desc + ch);
    uint32_t todo = __ballot_sync(kFull, d.x != 0 && !(d.x & kChunkOverflow));
    while (todo) {
      const uint32_t src = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t n = __shfl_sync(kFull, d.x, src);
      const unsigned long long tb = __shfl_sync(kFull, d.y, src);
      if (tb + n > P.temp_cap) continue; // dropped entries: the call is repeated with a larger buffer
      const uint64_t chs = c0 + g0 + src;
      const unsigned long long base = span0 + s_pre[g0 + src];
      const unsigned long long cb = tile_gbase(P, (uint32_t)(chs / kTileChunks)) + (chs % kTileChunks) * kChunkBytes; // global offset of the chunk's first byte
      for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t e = __ldg(P.temp + tb + i);
        const unsigned long long r = base + i;
        if (r < P.out_cap) put_record(P, r, cb + (e >> kPackLenBits), e & ((1u << kPackLenBits) - 1));
      }
    }
  }
}

bool store_normalises(const DeviceStore &st) { return st.flags & (kFlagIgnorePunct | kFlagElideSpace); }

size_t redo_smem_bytes(const DeviceStore &st) {
  return kSmemHeader + kRedoRingBytes + size_t(st.g4_words) * 4 + size_t(st.p23_words) * 4 + size_t(st.sx_words) * 4 + kQ2Bytes + kQ1Bytes +
         size_t(kScanWarps) * kPrivBytes + (store_normalises(st) ? size_t(kScanWarps) * kXfRowBytes : 0);
}

template <bool G, bool Q, bool C, bool XF, bool COOP>
cudaError_t launch_variant(const ScanParams &p, int sms, size_t smem, cudaStream_t stream) {
  const int grid = (int)(p.num_tiles < (uint32_t)sms ? p.num_tiles : (uint32_t)sms);
  // the lean per-candidate path covers every store and every flag set; only exact statistics need the
  // generic one (they count the candidates a start predicate rejects)
  constexpr bool can_fast = G || Q;
  const int mode = !can_fast || (p.flags & kCountAll) ? 0 : (p.flags & (kWordBoundary | kWordPrefix | kLineStart)) ? 2 : 1;
  bool launched = false;
  if constexpr (Q && !XF && !COOP) {
    if (p.st.sx_words) { // (the engine switches the tables on for the stores they pay for)
      if (mode == 2)
        scan_kernel<G, Q, C, 2, XF, COOP, true><<<grid, kScanThreads, smem, stream>>>(p);
      else if (mode == 1)
        scan_kernel<G, Q, C, 1, XF, COOP, true><<<grid, kScanThreads, smem, stream>>>(p);
      else
        scan_kernel<G, Q, C, 0, XF, COOP, true><<<grid, kScanThreads, smem, stream>>>(p);
      launched = true;
    }
  }
  if (launched) {
  } else if (mode == 0) {
    scan_kernel<G, Q, C, 0, XF, COOP><<<grid, kScanThreads, smem, stream>>>(p);
  } else if constexpr (can_fast) {
    if (mode == 2)
      scan_kernel<G, Q, C, 2, XF, COOP><<<grid, kScanThreads, smem, stream>>>(p);
    else
      scan_kernel<G, Q, C, 1, XF, COOP><<<grid, kScanThreads, smem, stream>>>(p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const uint64_t n_chunks = (uint64_t)p.num_tiles * kTileChunks;
  const uint32_t n_spans = (uint32_t)((n_chunks + kPrefixSpan - 1) / kPrefixSpan);
  prefix_sum_kernel<<<n_spans, kPrefixThreads, 0, stream>>>(p);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  prefix_scan_kernel<<<1, 1024, 0, stream>>>(p, n_spans);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  place_kernel<<<n_spans, kPrefixThreads, 0, stream>>>(p);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  redo_kernel<G, Q, C, XF, COOP><<<grid, kScanThreads, redo_smem_bytes(p.st), stream>>>(p);
  return cudaGetLastError();
}
// COOP kernels exist for the stores that can have many patterns behind one key: a 4-byte key (G
// without the class prefilter's long keys)
template <bool G, bool Q, bool C, bool XF>
cudaError_t launch_variant(const ScanParams &p, int sms, size_t smem, cudaStream_t stream) {
  if constexpr (G && !C) {
    if (p.st.coop) return launch_variant<G, Q, C, XF, true>(p, sms, smem, stream);
  }
  return launch_variant<G, Q, C, XF, false>(p, sms, smem, stream);
}
template <bool G, bool Q, bool C>
cudaError_t launch_variant(const ScanParams &p, int sms, size_t smem, cudaStream_t stream) {
  if (p.flags & kWindowMode) return launch_variant<G, Q, C, true>(p, sms, smem, stream);
  return launch_variant<G, Q, C, false>(p, sms, smem, stream);
}

template <bool G, bool Q, bool C, bool XF, bool COOP>
cudaError_t configure_variant(size_t smem_limit) {
  const int lim = (int)smem_limit;
  cudaError_t e = cudaFuncSetAttribute(scan_kernel<G, Q, C, 0, XF, COOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
  if (e != cudaSuccess) return e;
  if constexpr (G || Q) {
    e = cudaFuncSetAttribute(scan_kernel<G, Q, C, 1, XF, COOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(scan_kernel<G, Q, C, 2, XF, COOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return e;
  }
  if constexpr (Q && !XF && !COOP) {
    e = cudaFuncSetAttribute(scan_kernel<G, Q, C, 0, XF, COOP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(scan_kernel<G, Q, C, 1, XF, COOP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(scan_kernel<G, Q, C, 2, XF, COOP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return e;
  }
  return cudaFuncSetAttribute(redo_kernel<G, Q, C, XF, COOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
}
template <bool G, bool Q, bool C>
cudaError_t configure_variant(size_t smem_limit) {
  cudaError_t e = configure_variant<G, Q, C, false, false>(smem_limit);
  if (e == cudaSuccess) e = configure_variant<G, Q, C, true, false>(smem_limit);
  if constexpr (G && !C) {
    if (e == cudaSuccess) e = configure_variant<G, Q, C, false, true>(smem_limit);
    if (e == cudaSuccess) e = configure_variant<G, Q, C, true, true>(smem_limit);
  }
  return e;
}

} // namespace

size_t scan_smem_bytes(const DeviceStore &st, uint32_t stages, uint32_t chunk_cap, bool priv) {
  return kSmemHeader + size_t(stages) * kStageBytes + size_t(st.g4_words) * 4 + size_t(st.p23_words) * 4 + size_t(st.sx_words) * 4 + kQ2Bytes +
         size_t(kScanWarps) * chunk_cap * 4 + kQ1Bytes + (priv ? size_t(kScanWarps) * kPrivBytes : 0) +
         (priv && store_normalises(st) ? size_t(kScanWarps) * kXfRowBytes : 0);
}

ScanGeometry scan_pick_geometry(const DeviceStore &st, size_t smem_limit) {
  ScanGeometry g;
  // Stores with a transform flag are scanned in private chunk buffers (that is where they are folded
  // / normalised): a stage is handed back as soon as its chunks are copied out, so a short ring is
  // enough.  Plain stores are scanned in place: the deepest ring that fits (refills track the warps
  // more closely the more stages there are).  Whatever shared memory is left goes to the warps'
  // staging areas (denser matches before a chunk has to be redone).
  const bool priv = st.flags & kFlagAnyTransform;
  for (uint32_t s = priv ? kPrivStagesMax : (uint32_t)kMaxStages; s >= 2; --s) {
    if (scan_smem_bytes(st, s, kChunkCapMin, priv) > smem_limit) continue;
    const size_t spare = smem_limit - scan_smem_bytes(st, s, 0, priv);
    uint32_t cap = uint32_t(spare / (size_t(kScanWarps) * 4)) & ~7u;
    if (cap > kChunkCapMax) cap = kChunkCapMax;
    g.stages = s;
    g.chunk_cap = cap;
    return g;
  }
  return g;
}

cudaError_t scan_configure(size_t smem_limit) {
  cudaError_t e;
  if ((e = configure_variant<true, true, false>(smem_limit)) != cudaSuccess) return e;
  if ((e = configure_variant<true, false, false>(smem_limit)) != cudaSuccess) return e;
  if ((e = configure_variant<true, false, true>(smem_limit)) != cudaSuccess) return e;
  if ((e = configure_variant<false, true, false>(smem_limit)) != cudaSuccess) return e;
  return configure_variant<false, false, false>(smem_limit);
}

cudaError_t scan_launch(const ScanParams &p, int sms, cudaStream_t stream, uint32_t *launches) {
  const size_t smem = scan_smem_bytes(p.st, p.stages, p.chunk_cap, (p.flags & kWindowMode) != 0);
  const bool g = p.st.g4_words != 0, q = p.st.p23_words != 0, c = g && !q && p.st.cls.run != 0;
  if (launches) *launches += 5;
  if (g && q) return launch_variant<true, true, false>(p, sms, smem, stream);
  if (c) return launch_variant<true, false, true>(p, sms, smem, stream);
  if (g) return launch_variant<true, false, false>(p, sms, smem, stream);
  if (q) return launch_variant<false, true, false>(p, sms, smem, stream);
  return launch_variant<false, false, false>(p, sms, smem, stream);
}

} // namespace olm
