// scan.cu -- the per-offset scan (kernels K3 + K4 of SURVEY 2.1) for sm_100a.
//
// Replaces the hot loop of core_match() (omega_match/src/matcher.c:767-881) together with
// bloom_filter_query (bloom.c:51-64), probe_bucket (hash_table.c:91-109),
// scan_bucket_and_append (matcher.c:182-255), the short matcher (matcher.c:665-692,
// :804-880) and -- because matches leave the kernel already in final order -- the
// concatenate + radix sort of finalize_match_results (matcher.c:587-623, :258-325).
//
// Shape of the kernel
//   * persistent CTAs (grid = #SMs), 512 threads; tiles of 32 KiB positions are handed out
//     by an atomic ticket, so tile k is always started before tile k+1;
//   * each tile (+16 bytes in front, +112 behind) is brought into shared memory by ONE
//     cp.async.bulk (TMA, 1-D) that completes on an mbarrier; a ring of 2-3 stages keeps
//     60-100 KiB per SM in flight;
//   * stage 1, every position: big-endian gram by PRMT from two registers, one multiply,
//     one probe of the hashed gram bitmap in shared memory (and one of the short-pattern
//     bitmap when the store has 1..3 byte patterns) -> 16-bit candidate masks per lane;
//   * stage 2, candidates only: position predicates, ONE 16-byte slot load that carries
//     gram + bytes 4..7 + length (rejects almost every false candidate), remaining bytes
//     against the pattern store, end predicates, then the 4/3/2/1-byte sets;
//   * emission: a lane keeps up to four matches in registers, the warp prefix-sums the
//     counts, matches go to the warp's private staging area in position order; at the end
//     of the tile a decoupled look-back over the tile descriptors yields the tile's global
//     base and every warp copies its staged matches to their FINAL place in the result
//     array: offset ascending, length descending, no sort pass.
//   * a tile whose matches do not fit the staging area is re-evaluated once, writing to
//     HBM directly (counts are always exact, so the result is the same).
#include "scan.cuh"

#include <cstdio>

#include "olm_classes.h"
#include "olm_format.h"

namespace olm {

namespace {

constexpr unsigned long long kStateAggregate = 1ull << 62;
constexpr unsigned long long kStatePrefix = 2ull << 62;
constexpr unsigned long long kStateValueMask = (1ull << 62) - 1;
constexpr uint32_t kFull = 0xFFFFFFFFu;

struct StageInfo { // written by the producer thread, read by everyone after the mbarrier wait
  unsigned long long p0;   // segment-relative position of the tile's first byte
  unsigned long long len;  // segment length (N, or M_w)
  unsigned long long end;  // one past the last start position to evaluate
  long long boff;          // buffer offset of position p0
  unsigned long long emit_base;
  uint32_t tile;           // launch-local tile index
  uint32_t tail;
  uint32_t win;
  uint32_t staged;         // bytes valid behind p0 in the stage buffer
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
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
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.global.acquire.gpu.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned long long *p, unsigned long long v) {
  asm volatile("st.global.release.gpu.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// 16-byte slot load through the read-only path.  (ld.global.nc.L1::no_allocate was tried: the
// 32 MiB table of the 1M-pattern store then stopped being retained by L2 -- hit rate 95% -> 25%,
// 61 GB of DRAM reads per GiB scanned, profiles/r1_notes.md -- so the default policy stays.)
__device__ __forceinline__ uint4 ldg_slot(const Slot *p) {
  return __ldg(reinterpret_cast<const uint4 *>(p));
}

// little-endian 32-bit word at an arbitrary shared-memory byte address
__device__ __forceinline__ uint32_t lds_le32(const uint8_t *q) {
  const uint32_t a = smem_u32(q);
  uint32_t lo, hi;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(a & ~3u));
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"((a & ~3u) + 4));
  return __funnelshift_r(lo, hi, (a & 3u) * 8u);
}

struct TileCtx {
  const uint8_t *sb; // stage buffer; sb[kTilePre + i] is the byte at tile position i
  unsigned long long p0, len, end;
  long long boff;
  uint32_t staged, tail;
};

template <bool HAS_G4, bool HAS_P23>
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

  __device__ __forceinline__ uint32_t hay_byte(const TileCtx &T, unsigned long long pos) const {
    const unsigned long long rel = pos - T.p0;
    if (rel < T.staged) return T.sb[kTilePre + rel];
    return P.buf[T.boff + (long long)rel];
  }

  // bytes [8, len) of a candidate against the pattern store (bytes 0..7 are already equal)
  __device__ __forceinline__ bool tail_equal(const TileCtx &T, unsigned long long pos, uint32_t len,
                                             uint32_t store_off) const {
    const uint8_t *pat = P.st.store + store_off;
    for (uint32_t i = 8; i < len; ++i)
      if (hay_byte(T, pos + i) != __ldg(pat + i)) return false;
    return true;
  }

  // end-side predicates for a long match (matcher.c:233, :239, :247): all guarded by e < n
  __device__ __forceinline__ bool end_ok_long(const TileCtx &T, unsigned long long e) const {
    if (!(fl & (kWordBoundary | kWordSuffix | kLineEnd))) return true;
    if (e >= T.len) return true;
    const uint32_t c = hay_byte(T, e);
    if ((fl & (kWordBoundary | kWordSuffix)) && is_word_byte(c)) return false;
    if ((fl & kLineEnd) && !is_line_end_byte(c)) return false;
    return true;
  }
  // short matcher, matcher.c:804-880: for lengths 2..4 the word-boundary test reads
  // haystack[pos+L] without a bound (:812,:830,:848) -> `tail` stands in at pos+L == n.
  __device__ __forceinline__ bool end_ok_short(const TileCtx &T, unsigned long long e, uint32_t L) const {
    if (!(fl & (kWordBoundary | kWordSuffix | kLineEnd))) return true;
    const bool inside = e < T.len;
    const uint32_t c = inside ? hay_byte(T, e) : T.tail;
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

  // ---- one candidate position, split in two so that several probes can be in flight ----
  // probe_issue : position predicates (matcher.c:770-776, :195-196, :806-807), the gram, and
  //               the loads of the candidate's home bucket (two 16-byte slots = one sector);
  // probe_finish: everything else the reference does for the position (matcher.c:782-880);
  //               `emit(len)` is called once per accepted match, longest first.
  struct Probe {
    uint32_t tpos, gram, bucket, flags; // flags: 1 = alive, 2 = gram candidate (and >= 4 bytes left), 4 = short candidate
    uint4 a, b;
  };

  __device__ __forceinline__ void probe_issue(const TileCtx &T, bool valid, uint32_t tpos, bool cand_g, bool cand_p,
                                              Probe &pr) const {
    pr.tpos = tpos;
    pr.flags = 0;
    pr.gram = 0;
    pr.bucket = 0;
    pr.a = make_uint4(0, 0, 0, 0);
    pr.b = make_uint4(0, 0, 0, 0);
    if (!valid) return;
    const unsigned long long pos = T.p0 + tpos;
    const uint8_t *q = T.sb + kTilePre + tpos;
    if (fl & (kWordBoundary | kWordPrefix | kLineStart)) {
      const uint32_t prev = q[-1];
      if (fl & kWordBoundary) { // matcher.c:770-776
        const bool cw = is_word_byte(q[0]);
        const bool pw = pos > 0 ? is_word_byte(prev) : false;
        if (cw == pw) return;
      }
      if ((fl & kWordPrefix) && pos > 0 && is_word_byte(prev)) return;     // :195, :806
      if ((fl & kLineStart) && pos > 0 && !is_line_end_byte(prev)) return; // :196, :807
    }
    pr.gram = __byte_perm(lds_le32(q), 0, 0x0123);
    pr.flags = 1u | (cand_p ? 4u : 0u);
    if (HAS_G4 && cand_g && T.len - pos >= 4) {
      pr.flags |= 2u;
      pr.bucket = (pr.gram * kHashMul) >> P.st.slot_shift;
      pr.a = ldg_slot(P.st.slots + 2 * (size_t)pr.bucket);
      pr.b = ldg_slot(P.st.slots + 2 * (size_t)pr.bucket + 1);
    }
  }

  template <typename Emit>
  __device__ __forceinline__ void probe_finish(const TileCtx &T, const Probe &pr, Emit &&emit) const {
    if (!(pr.flags & 1u)) return;
    const uint32_t tpos = pr.tpos, gram = pr.gram;
    const unsigned long long pos = T.p0 + tpos;
    const unsigned long long rem = T.len - pos;
    const uint8_t *q = T.sb + kTilePre + tpos;
    bool emitted = false;
    const bool longest = fl & kLongestOnly;

    if (HAS_G4 && (pr.flags & 2u)) {
      // bucketized linear probing: the key is in the first bucket (from its home) that has it;
      // a bucket with a free slot ends the probe
      uint4 a = pr.a, b = pr.b, s = make_uint4(0, 0, 0, 0);
      uint32_t bucket = pr.bucket;
      while (true) {
        if (a.z != 0 && a.x == gram) {
          s = a;
          break;
        }
        if (b.z != 0 && b.x == gram) {
          s = b;
          break;
        }
        if (a.z == 0 || b.z == 0) break;
        bucket = (bucket + 1) & P.st.slot_mask;
        a = ldg_slot(P.st.slots + 2 * (size_t)bucket);
        b = ldg_slot(P.st.slots + 2 * (size_t)bucket + 1);
      }
      if (s.z != 0) {
        const uint32_t meta = s.z;
        const uint32_t hay4 = lds_le32(q + 4);
        if (meta & kSlotValueMask) {
          n_hits += stat_inc;
          n_long_hits += stat_inc;
        }
        if (meta & kSlotMulti) {
          const uint32_t cnt = meta & kSlotValueMask;
          for (uint32_t j = 0; j < cnt; ++j) {
            const uint4 r = __ldg(reinterpret_cast<const uint4 *>(P.st.recs + s.w + j));
            const uint32_t len = r.y;
            if (len > rem) continue; // matcher.c:203
            n_cmp += stat_inc;
            const uint32_t m = len >= 8 ? kFull : ((1u << ((len - 4) * 8)) - 1u);
            if ((hay4 ^ r.x) & m) continue;
            if (len > 8 && !tail_equal(T, pos, len, r.z)) continue;
            if (!end_ok_long(T, pos + len)) continue;
            emit(len);
            emitted = true;
            if (longest) break;
          }
        } else {
          const uint32_t len = meta & kSlotValueMask;
          if (len != 0 && len <= rem) {
            n_cmp += stat_inc;
            const uint32_t m = len >= 8 ? kFull : ((1u << ((len - 4) * 8)) - 1u);
            if (((hay4 ^ s.y) & m) == 0 && (len <= 8 || tail_equal(T, pos, len, s.w)) &&
                end_ok_long(T, pos + len)) {
              emit(len);
              emitted = true;
            }
          }
        }
        if ((meta & kSlotShort4) && !(longest && emitted)) {
          if (end_ok_short(T, pos + 4, 4)) {
            emit(4u);
            emitted = true;
            n_hits += stat_inc;
          } else {
            n_miss += stat_inc;
          }
        }
      }
    }
    if (HAS_P23 && (pr.flags & 4u) && !(longest && emitted)) {
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
          if (end_ok_short(T, pos + 3, 3)) {
            emit(3u);
            emitted = true;
            n_hits += stat_inc;
          } else {
            n_miss += stat_inc;
          }
        }
      }
      if (P.st.n2 && rem >= 2 && !(longest && emitted)) {
        const uint32_t k2 = gram >> 16;
        if ((__ldg(P.st.bitmap2 + (k2 >> 5)) >> (k2 & 31)) & 1u) {
          if (end_ok_short(T, pos + 2, 2)) {
            emit(2u);
            emitted = true;
            n_hits += stat_inc;
          } else {
            n_miss += stat_inc;
          }
        }
      }
      if (P.st.n1 && !(longest && emitted)) {
        const uint32_t k1 = gram >> 24;
        if ((P.st.bitmap1[k1 >> 5] >> (k1 & 31)) & 1u) {
          if (end_ok_short(T, pos + 1, 1)) {
            emit(1u);
            n_hits += stat_inc;
          } else {
            n_miss += stat_inc;
          }
        }
      }
    }
  }

  // stage 1 for one lane: 16 positions -> candidate masks (bit k = position lpos + k)
  __device__ __forceinline__ void stage1(const TileCtx &T, uint32_t lpos, uint32_t &cg, uint32_t &cp,
                                         uint32_t &valid) const {
    const uint8_t *src = T.sb + kTilePre + lpos;
    const uint4 v = *reinterpret_cast<const uint4 *>(src);
    const uint32_t w4 = *reinterpret_cast<const uint32_t *>(src + 16);
    const uint32_t w[5] = {v.x, v.y, v.z, v.w, w4};
    cg = 0;
    cp = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t gram = __byte_perm(w[k >> 2], w[(k >> 2) + 1], 0x0123u + 0x1111u * (k & 3));
      if (HAS_G4) {
        const uint32_t b = (gram * kHashMul) >> P.st.g4_shift;
        cg |= ((g4s[b >> 5] >> (b & 31)) & 1u) << k;
      }
      if (HAS_P23) {
        const uint32_t b = ((gram & P.st.p23_and) * P.st.p23_mul) >> P.st.p23_shift;
        cp |= ((p23s[b >> 5] >> (b & 31)) & 1u) << k;
      }
    }
    const unsigned long long lp = T.p0 + lpos;
    valid = 0;
    if (lp < T.end) valid = (T.end - lp >= 16) ? 0xFFFFu : ((1u << (uint32_t)(T.end - lp)) - 1u);
  }

  // One 512-byte chunk of a warp.  The warp's candidates are compacted into a queue in
  // position order; then every lane takes one candidate per sub-step, kProbeUnroll sub-steps
  // are issued together (their slot loads overlap), and accepted matches are appended in
  // candidate order (ballot + popc; a shuffle scan only when a position has several matches;
  // a position with more than four matches is evaluated a second time for the rest).
  //   direct == false: matches go to `stage` (packed, shared memory, `cap` entries); when they
  //                    do not fit, *overflow is set and the tile is redone with
  //   direct == true : matches go to P.out[out_base ...] as final records.
  // Returns the exact number of matches of the chunk in both modes.
  static constexpr int kProbeUnroll = 4;
  __device__ __forceinline__ uint32_t scan_chunk(const TileCtx &T, uint32_t cbase, uint32_t lane, uint32_t *stage,
                                                 uint32_t used, uint32_t cap, uint16_t *queue, bool direct,
                                                 unsigned long long out_base, unsigned long long emit_base,
                                                 const uint32_t *map, uint32_t *overflow) const {
    uint32_t cg, cp, valid;
    stage1(T, cbase + lane * 16, cg, cp, valid);
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
        queue[o++] = (uint16_t)((lane * 16 + k) | (((cg >> k) & 1u) << 9) | (((cp >> k) & 1u) << 10));
      }
    }
    __syncwarp();
    uint32_t found = 0;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t base = 0; base < total; base += 32 * kProbeUnroll) {
      Probe pr[kProbeUnroll];
#pragma unroll
      for (int u = 0; u < kProbeUnroll; ++u) {
        const uint32_t idx = base + u * 32 + lane;
        const bool ok = idx < total;
        const uint32_t e = ok ? queue[idx] : 0u;
        probe_issue(T, ok, cbase + (e & 511u), (e >> 9) & 1u, (e >> 10) & 1u, pr[u]);
      }
#pragma unroll
      for (int u = 0; u < kProbeUnroll; ++u) {
        if (base + u * 32 >= total) break;
        uint32_t n = 0, m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        const uint32_t tpos = pr[u].tpos;
        probe_finish(T, pr[u], [&](uint32_t len) {
          if (n == 0) m0 = len;
          else if (n == 1) m1 = len;
          else if (n == 2) m2 = len;
          else if (n == 3) m3 = len;
          ++n;
        });
        const uint32_t bal = __ballot_sync(kFull, n > 0);
        if (!bal) continue;
        uint32_t pre, tot;
        const bool many = __any_sync(kFull, n > 1);
        if (!many) {
          pre = __popc(bal & lt);
          tot = __popc(bal);
        } else {
          uint32_t in2 = n;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, in2, d);
            if (lane >= (uint32_t)d) in2 += t;
          }
          pre = in2 - n;
          tot = __shfl_sync(kFull, in2, 31);
        }
        const uint32_t at = used + found + pre;
        auto put = [&](uint32_t i, uint32_t len) {
          if (direct) {
            const unsigned long long r = out_base + at + i;
            if (r < P.out_cap) write_record(r, emit_base, T.p0 + tpos, len, map);
          } else if (at + i < cap && !(len >> kPackLenBits)) {
            stage[at + i] = (tpos << kPackLenBits) | len;
          } else {
            *overflow = 1;
          }
        };
        if (n > 0) put(0, m0);
        if (n > 1) put(1, m1);
        if (n > 2) put(2, m2);
        if (n > 3) put(3, m3);
        if (many && __any_sync(kFull, n > 4)) {
          if (n > 4) { // evaluate the position again for matches 5, 6, ...
            uint32_t i = 0;
            stat_inc = 0;
            probe_finish(T, pr[u], [&](uint32_t len) {
              if (i >= 4) put(i, len);
              ++i;
            });
            stat_inc = 1;
          }
        }
        found += tot;
      }
    }
    __syncwarp();
    return found;
  }

  __device__ __forceinline__ void write_record(unsigned long long r, unsigned long long emit_base,
                                               unsigned long long pos, uint32_t len, const uint32_t *map) const {
    unsigned long long off = emit_base + pos;
    if (map) { // matcher.c:986-997: back to source coordinates
      const uint32_t a = __ldg(map + pos), b = __ldg(map + pos + len - 1);
      off = emit_base + a;
      len = b - a + 1;
    }
    Record *o = P.out + r;
    o->offset = off;
    *reinterpret_cast<unsigned long long *>(&o->len) = (unsigned long long)len;
    o->ptr = P.match_ptr_base + off;
  }
};

template <bool HAS_G4, bool HAS_P23>
__global__ void __launch_bounds__(kScanThreads, 1) scan_kernel(const __grid_constant__ ScanParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t S = P.stages;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);                        // [3]
  uint32_t *s_ovf = reinterpret_cast<uint32_t *>(smem + 24);                  // [2]
  unsigned long long *s_excl = reinterpret_cast<unsigned long long *>(smem + 32);
  uint32_t *s_wcnt = reinterpret_cast<uint32_t *>(smem + 40);                 // [16]
  uint32_t *s_wpre = reinterpret_cast<uint32_t *>(smem + 104);                // [16]
  StageInfo *s_info = reinterpret_cast<StageInfo *>(smem + 256);              // [3] x 64 B
  uint8_t *ring = smem + 256 + 3 * 64 + 64;                                   // 512: 128-aligned
  uint32_t *g4s = reinterpret_cast<uint32_t *>(ring + S * kStageBytes);
  uint32_t *p23s = g4s + (HAS_G4 ? P.st.g4_words : 0);
  uint32_t *staging = p23s + (HAS_P23 ? P.st.p23_words : 0);
  uint16_t *queues = reinterpret_cast<uint16_t *>(staging + kScanWarps * P.stage_cap);

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t fl = P.flags;
  const bool window_mode = fl & kWindowMode;

  // ---- filters into shared memory
  if (HAS_G4) {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.st.g4);
    uint4 *dst = reinterpret_cast<uint4 *>(g4s);
    for (uint32_t i = tid; i < P.st.g4_words / 4; i += kScanThreads) dst[i] = __ldg(src + i);
  }
  if (HAS_P23) {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.st.p23);
    uint4 *dst = reinterpret_cast<uint4 *>(p23s);
    for (uint32_t i = tid; i < P.st.p23_words / 4; i += kScanThreads) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    for (uint32_t s = 0; s < S; ++s) mbar_init(&full[s], 1);
    s_ovf[0] = s_ovf[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // The producer thread takes the next ticket and starts the bulk copy of that tile.
  auto produce = [&](uint32_t s) {
    const uint32_t t = atomicAdd(P.ticket, 1u);
    StageInfo &I = s_info[s];
    I.tile = t;
    if (t >= P.num_tiles) return;
    uint32_t win = 0;
    if (window_mode) {
      win = t / P.tiles_per_win;
      const WindowDesc wd = P.windows[win];
      I.p0 = (unsigned long long)(t % P.tiles_per_win) * kTileBytes;
      I.len = wd.norm_len;
      I.end = wd.norm_len;
      I.boff = (long long)(P.win_buf_off + (unsigned long long)win * P.win_stride + I.p0);
      I.tail = wd.tail;
      I.emit_base = P.win_src_base + (unsigned long long)win * kWindowBytes;
    } else {
      I.p0 = P.scan_begin + (unsigned long long)t * kTileBytes;
      I.len = P.seg_len;
      I.end = P.scan_end < P.seg_len ? P.scan_end : P.seg_len;
      I.boff = P.seg_buf_off + (long long)I.p0;
      I.tail = P.tail_byte;
      I.emit_base = 0;
    }
    I.win = win;
    uint32_t bytes = 0;
    I.staged = 0;
    if (I.p0 < I.end) {
      const long long pre = I.boff >= kTilePre ? kTilePre : 0;
      long long e = I.boff + kTileBytes + kTileHalo;
      if (e > (long long)P.buf_len) e = (long long)P.buf_len;
      bytes = (uint32_t)(e - (I.boff - pre));
      I.staged = (uint32_t)(e - I.boff);
      mbar_expect_tx(&full[s], bytes);
      tma_load_1d(ring + (size_t)s * kStageBytes + (kTilePre - pre), P.buf + (I.boff - pre), bytes, &full[s]);
    } else {
      mbar_expect_tx(&full[s], 0); // empty tile: the phase completes at once
    }
  };
  if (tid == 32)
    for (uint32_t s = 0; s < S; ++s) produce(s);
  __syncthreads();

  Scanner<HAS_G4, HAS_P23> sc{P, g4s, p23s, fl};
  const uint32_t cap = P.stage_cap;
  uint32_t *my_stage = staging + warp * cap;
  uint16_t *my_queue = queues + warp * kChunkBytes;

  for (uint32_t k = 0;; ++k) {
    const uint32_t s = k % S;
    const StageInfo &I = s_info[s];
    const uint32_t tile = *reinterpret_cast<const volatile uint32_t *>(&I.tile);
    if (tile >= P.num_tiles) break;
    mbar_wait(&full[s], (k / S) & 1u);

    TileCtx T;
    T.sb = ring + (size_t)s * kStageBytes;
    T.p0 = I.p0;
    T.len = I.len;
    T.end = I.end;
    T.boff = I.boff;
    T.staged = I.staged;
    T.tail = I.tail;
    const unsigned long long emit_base = I.emit_base;
    const uint32_t *map =
        (window_mode && !(fl & kIdentityMap)) ? P.map + (size_t)I.win * kWindowBytes : nullptr;
    uint32_t *ovf = &s_ovf[k & 1];

    // ---- scan this warp's 2 KiB of the tile
    uint32_t wc = 0;
    for (uint32_t it = 0; it < kWarpSpan / kChunkBytes; ++it) {
      const uint32_t cbase = warp * kWarpSpan + it * kChunkBytes;
      if (T.p0 + cbase >= T.end) break;
      wc += sc.scan_chunk(T, cbase, lane, my_stage, wc, cap, my_queue, false, 0, emit_base, map, ovf);
    }
    if (lane == 0) s_wcnt[warp] = wc;
    __syncthreads(); // (A) tile evaluated, counts visible

    const bool overflow = *reinterpret_cast<volatile uint32_t *>(ovf) != 0;
    if (tid == 32) {
      s_ovf[(k + 1) & 1] = 0;
      if (!overflow) produce(s); // stage buffer is free again
    }
    if (warp == 0) {
      // warp totals -> exclusive prefixes; then the decoupled look-back for the tile base
      uint32_t c = lane < kScanWarps ? s_wcnt[lane] : 0, incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, d);
        if (lane >= (uint32_t)d) incl += t;
      }
      if (lane < kScanWarps) s_wpre[lane] = incl - c;
      const unsigned long long tile_total = __shfl_sync(kFull, incl, 31);
      const long long g = (long long)P.tile_base + tile;
      unsigned long long excl = 0;
      if (g > 0) {
        if (lane == 0) st_release(P.tile_state + g, kStateAggregate | tile_total);
        long long j = g - 1;
        while (true) {
          const long long mine = j - lane;
          unsigned long long v = kStatePrefix; // virtual predecessor of tile 0: prefix 0
          if (mine >= 0) {
            do {
              v = ld_acquire(P.tile_state + mine);
            } while ((v >> 62) == 0);
          }
          const uint32_t is_prefix = __ballot_sync(kFull, (v >> 62) == 2);
          const uint32_t first = is_prefix ? (__ffs(is_prefix) - 1) : 32;
          unsigned long long add = lane <= first ? (v & kStateValueMask) : 0;
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) add += __shfl_xor_sync(kFull, add, d);
          excl += add;
          if (is_prefix) break;
          j -= 32;
        }
      }
      if (lane == 0) {
        st_release(P.tile_state + g, kStatePrefix | (excl + tile_total));
        *s_excl = excl;
        if (tile == P.num_tiles - 1) *P.total = excl + tile_total;
      }
    }
    __syncthreads(); // (B) tile base known

    const unsigned long long base = *s_excl + s_wpre[warp];
    if (!overflow) {
      for (uint32_t i = lane; i < wc; i += 32) {
        const uint32_t e = my_stage[i];
        const unsigned long long r = base + i;
        if (r < P.out_cap) sc.write_record(r, emit_base, T.p0 + (e >> kPackLenBits), e & ((1u << kPackLenBits) - 1), map);
      }
    } else {
      // rare: staging overflowed somewhere in this tile -> evaluate again, straight to HBM
      uint32_t dummy = 0;
      uint32_t done = 0;
      for (uint32_t it = 0; it < kWarpSpan / kChunkBytes; ++it) {
        const uint32_t cbase = warp * kWarpSpan + it * kChunkBytes;
        if (T.p0 + cbase >= T.end) break;
        done += sc.scan_chunk(T, cbase, lane, my_stage, done, cap, my_queue, true, base, emit_base, map, &dummy);
      }
      __syncthreads();
      if (tid == 32) produce(s);
    }
    __syncwarp();
  }

  // statistics: one atomic per warp and counter
  unsigned long long h = sc.n_hits, mi = sc.n_miss, cm = sc.n_cmp, lh = sc.n_long_hits;
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

template <bool G, bool Q>
cudaError_t launch_variant(const ScanParams &p, int grid, size_t smem, cudaStream_t stream) {
  scan_kernel<G, Q><<<grid, kScanThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

} // namespace

size_t scan_smem_bytes(const DeviceStore &st, uint32_t stages, uint32_t stage_cap) {
  return 512 + size_t(stages) * kStageBytes + size_t(st.g4_words) * 4 + size_t(st.p23_words) * 4 +
         size_t(kScanWarps) * stage_cap * 4 + kQueueBytes;
}

uint32_t scan_pick_stages(const DeviceStore &st, size_t smem_limit, uint32_t *stage_cap) {
  for (uint32_t s = 3; s >= 2; --s) {
    if (scan_smem_bytes(st, s, kStageCapMin) > smem_limit) continue;
    // whatever shared memory is left goes to the staging areas (denser matches before a tile
    // has to be redone)
    const size_t spare = smem_limit - scan_smem_bytes(st, s, 0);
    uint32_t cap = uint32_t(spare / (size_t(kScanWarps) * 4)) & ~31u;
    if (cap > kStageCapMax) cap = kStageCapMax;
    *stage_cap = cap;
    return s;
  }
  return 0;
}

cudaError_t scan_configure(size_t smem_limit) {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(scan_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(scan_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(scan_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(scan_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit)) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t scan_launch(const ScanParams &p, int grid, cudaStream_t stream) {
  const size_t smem = scan_smem_bytes(p.st, p.stages, p.stage_cap);
  const bool g = p.st.g4_words != 0, q = p.st.p23_words != 0;
  if (g && q) return launch_variant<true, true>(p, grid, smem, stream);
  if (g) return launch_variant<true, false>(p, grid, smem, stream);
  if (q) return launch_variant<false, true>(p, grid, smem, stream);
  return launch_variant<false, false>(p, grid, smem, stream);
}

} // namespace olm
