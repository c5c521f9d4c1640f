// scan_device.cuh -- device-side building blocks shared by scan.cu and stats.cu: shared-memory
// access by 32-bit shared-space address, the description of a tile / of a chunk, and the chunk
// builders that put the bytes of one 512-position chunk -- as the matcher has to see them -- into a
// warp's private buffer:
//
//   build_copy   plain stores: the bytes themselves; case-folding-only stores: a-z -> A-Z
//                (transform_table.c:25) while copying;
//   build_xf     stores with ignore-punctuation / elide-whitespace: transform_apply()
//                (transform_table.c:36-88) of the chunk's 512 source bytes + the 128 behind them,
//                by the warp, in registers: SWAR byte classes, the "inside a whitespace run" state
//                carried from lane to lane by ballots, a prefix scan of the kept-byte counts, the
//                kept bytes stored at their normalised index.  No normalised copy of the haystack
//                and no offset map ever reach HBM: a match is translated back to source
//                coordinates (matcher.c:986-997) from the keep masks the warp leaves in shared
//                memory ("rows": first normalised index + keep bits of every lane).
//
// Exactness at the edges: windows are independent -- the run state restarts at
// a window's first byte and nothing of the next window is looked at; the byte in front of the
// chunk and the run state come from the nearest non-skipped source byte in front of it (looked up
// in the staged tile, in global memory when a run of skipped bytes is longer than that); what lies
// behind the 640 source bytes is produced, when a comparison really needs it, by a per-thread
// walk over the source (xf_walk) that also knows about the window's trailing-space trim
// (transform_table.c:82-84).
#pragma once
#include <cstddef>
#include <cstdint>

#include "olm_classes.h"
#include "olm_format.h"
#include "scan.cuh"

namespace olm {
namespace dev {

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kNoTile = 0xFFFFFFFFu;
constexpr uint32_t kBeyond = 0x100u; // "byte" behind the end of a normalised window

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// Shared memory by 32-bit shared-space address.  The scanning warps never dereference a generic
// pointer into shared memory: every generic access makes nvcc recompute the shared window base
// (S2R SR_CgaCtaId + LEA, on the slow XU pipe) -- ~35 of them per chunk saturated that pipe.
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long lds64u(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64u(uint32_t a, unsigned long long v) {
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4 &v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// little-endian 32-bit word at an arbitrary shared-memory byte address
__device__ __forceinline__ uint32_t lds_le32(uint32_t a) {
  const uint32_t lo = lds32(a & ~3u), hi = lds32((a & ~3u) + 4);
  return __funnelshift_r(lo, hi, a << 3);
}

// One tile of a launch: written by the producer lane of scan_kernel (or computed on the spot by the
// other kernels), read by the warps that scan its chunks.
struct StageInfo {
  unsigned long long gbase; // global haystack offset of the tile's first byte (what a record's offset counts from)
  uint32_t rem0;            // plain / case-folded: matchable bytes from the tile's first position on (capped at 2^31);
                            // normalising stores: SOURCE bytes from there to the end of the window
  uint32_t nscan;           // start positions (source bytes) of this tile to evaluate (<= kTileBytes)
  unsigned long long p0;    // position of the tile's first byte inside its segment (the haystack / its window)
  long long boff;           // buffer offset of that byte
  uint32_t tile;            // launch-local tile index, kNoTile = no more work
  uint32_t staged;          // bytes valid behind p0 in the stage buffer
  uint32_t tail;            // the byte assumed behind the end of the segment   (tail, win: one 8-byte load)
  uint32_t win;             // window of the launch the tile belongs to
  uint32_t seq;             // tile iteration of the CTA this entry describes
  uint32_t stage_par;       // stage of the ring that holds the tile | parity of its mbarrier phase << 16
  uint32_t _pad[2];
};
static_assert(sizeof(StageInfo) == 64 && offsetof(StageInfo, tail) % 8 == 0 && offsetof(StageInfo, rem0) % 8 == 0, "StageInfo layout");

// What the matching code knows about the bytes it works on: a whole tile in its stage buffer
// (positions relative to the tile) or one chunk in a warp's private buffer (relative to the chunk).
struct TileCtx {
  uint32_t sb32;            // shared-space address of the buffer; byte sb32 + kTilePre + i is position i
  unsigned long long gbase; // global offset of position 0 (normalising stores: of the chunk's first SOURCE byte)
  long long boff;           // buffer offset of position 0 (normalising stores: of the chunk's first source byte)
  uint32_t rem0;            // bytes of the segment from position 0 on (kRemUnknown: see build_xf)
  uint32_t nscan;           // positions to evaluate
  uint32_t staged, tail;    // bytes valid behind position 0 in the buffer; byte assumed at the segment's end
  bool first;               // position 0 is the first byte of the segment
};

// Global haystack offset of the first byte of launch-local tile t (StageInfo::gbase without the rest).
__device__ __forceinline__ unsigned long long tile_gbase(const ScanParams &P, uint32_t t) {
  if (P.flags & kWindowMode) {
    constexpr uint32_t kTilesPerWin = kWindowBytes / kTileBytes;
    return P.win_src_base + (unsigned long long)(t / kTilesPerWin) * kWindowBytes + (unsigned long long)(t % kTilesPerWin) * kTileBytes;
  }
  return P.scan_begin + (unsigned long long)t * kTileBytes;
}

// Fills `I` for launch-local tile t.
__device__ __forceinline__ void fill_tile(const ScanParams &P, uint32_t t, StageInfo &I) {
  I.tile = t;
  uint32_t win = 0;
  unsigned long long len, end;
  if (P.flags & kWindowMode) {
    constexpr uint32_t kTilesPerWin = kWindowBytes / kTileBytes; // (P.tiles_per_win; a constant spares the division)
    win = t / kTilesPerWin;
    I.p0 = (unsigned long long)(t % kTilesPerWin) * kTileBytes;
    const unsigned long long wbase = (unsigned long long)win * kWindowBytes;
    unsigned long long wlen = P.win_src_len - wbase;
    if (wlen > kWindowBytes) wlen = kWindowBytes;
    I.tail = ' ';
    if (P.windows) {
      const WindowDesc wd = P.windows[win];
      I.tail = wd.tail;
      if (P.flags & kIdentityMap) wlen = wd.norm_len; // (case folding only: the trim is known up front)
    }
    len = wlen;
    end = wlen;
    I.boff = (long long)(P.win_buf_off + wbase + I.p0);
    I.gbase = P.win_src_base + wbase + I.p0;
  } else {
    I.p0 = P.scan_begin + (unsigned long long)t * kTileBytes;
    len = P.seg_len;
    end = P.scan_end < P.seg_len ? P.scan_end : P.seg_len;
    I.boff = P.seg_buf_off + (long long)I.p0;
    I.tail = P.tail_byte;
    I.gbase = I.p0;
  }
  I.win = win;
  I.staged = 0;
  I.rem0 = 0;
  I.nscan = 0;
  if (I.p0 < end) {
    long long e = I.boff + kTileBytes + kTileHalo;
    if (e > (long long)P.buf_len) e = (long long)P.buf_len;
    I.staged = (uint32_t)(e - I.boff);
    const unsigned long long left = len - I.p0, ns = end - I.p0;
    I.rem0 = left > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)left;
    I.nscan = ns > (unsigned long long)kTileBytes ? (uint32_t)kTileBytes : (uint32_t)ns;
  }
}

// ---- byte classes of the four bytes of a word: bit 7 of every byte that is in the class ----------
// lo <= t <= hi for bytes t < 0x80: bit 7 of t + (0x80 - lo) is set iff t >= lo, of t + (0x7f - hi) iff t > hi
__device__ __forceinline__ uint32_t swar_in(uint32_t t, uint32_t lo, uint32_t hi) {
  return (t + (0x80u - lo) * 0x01010101u) & ~(t + (0x7Fu - hi) * 0x01010101u);
}
// bits 7,15,23,31 -> bits 0..3
__device__ __forceinline__ uint32_t gather4(uint32_t m) { return __umulhi(m, 0x02040810u) & 0xFu; }

// transform_init()'s table (transform_table.c:13-34) for four bytes at once: returns the mapped
// bytes; *sp / *sk = nibbles of the bytes that are whitespace (emitted as ' ', runs collapse) /
// skipped.  The whitespace test comes first, then punctuation, then case folding.
__device__ __forceinline__ uint32_t xf_classify4(uint32_t x, bool ci, bool ip, bool ew, uint32_t *sp, uint32_t *sk) {
  const uint32_t t = x & 0x7F7F7F7Fu, ascii = ~x & 0x80808080u;
  uint32_t out = x;
  *sp = 0;
  *sk = 0;
  if (ew) { // IS_SPACE, common.h:54-57: \a \b \t \n \v \f \r and ' '
    const uint32_t s = (swar_in(t, 7, 13) | swar_in(t, 32, 32)) & ascii;
    const uint32_t sb = (s >> 7) * 0xFFu;
    out = (out & ~sb) | (0x20202020u & sb);
    *sp = gather4(s);
  }
  if (ip) { // IS_PUNCT, common.h:45-52: ASCII punctuation without '_'
    const uint32_t p = (swar_in(t, 33, 47) | swar_in(t, 58, 64) | swar_in(t, 91, 94) | swar_in(t, 96, 96) |
                        swar_in(t, 123, 126)) & ascii;
    *sk = gather4(p);
  }
  if (ci) { // toupper in the C locale: only a-z change (punctuation and whitespace are not letters)
    const uint32_t l = swar_in(t, 97, 122) & ascii;
    out ^= l >> 2;
  }
  return out;
}
__device__ __forceinline__ uint32_t fold4(uint32_t x) {
  const uint32_t t = x & 0x7F7F7F7Fu;
  return x ^ ((swar_in(t, 97, 122) & ~x & 0x80808080u) >> 2);
}

// ---- per-warp state of a normalised chunk (shared memory, kXfRowBytes per warp) ------------------
//   [0, 64)     source offsets of normalised bytes appended by the walk (build_xf, rare)
//   [64, 96)    where the walk starts: normalised index, source offset, run state, source bytes to the
//               window's end; the run state in front of the chunk; bytes the per-byte table covers
//   [128, 768)  per normalised byte: source offset - normalised index (the bytes dropped in front of
//               it), 255 = more than 254: found by a walk from the chunk's first byte
constexpr uint32_t kXfExt = 0, kXfCurIdx = 64, kXfCurSrc = 68, kXfCurSpace = 72, kXfSrcLim = 76, kXfCin0 = 80, kXfDelta = 128;
constexpr uint32_t kXfExtSlots = 16;
static_assert(kXfDelta + kPrivData <= (uint32_t)kXfRowBytes && kXfExt + 4 * kXfExtSlots <= kXfCurIdx, "layout of the chunk state");

struct Walk {
  uint32_t byte; // normalised byte, kBeyond when the index lies behind the window's (trimmed) end
  uint32_t src;  // chunk-relative source offset of that byte
};
// Normalised index j -> its byte, by transforming the source from where the chunk builder stopped
// (from0: from the chunk's first byte).  One thread, global memory; rare: patterns longer than
// what 128 source bytes behind a chunk give, long runs of skipped bytes.  A ' ' only exists if
// another kept byte follows it inside the window: the window's last byte is dropped when it is a
// space (transform_table.c:82-84).
static __device__ __noinline__ Walk xf_walk(const uint8_t *src0, uint32_t flags, uint32_t xf32, uint32_t j, bool from0) {
  const bool ci = flags & kFlagIgnoreCase, ip = flags & kFlagIgnorePunct, ew = flags & kFlagElideSpace;
  uint32_t idx = from0 ? 0u : lds32(xf32 + kXfCurIdx), src = from0 ? 0u : lds32(xf32 + kXfCurSrc);
  uint32_t in_space = lds32(xf32 + (from0 ? kXfCin0 : kXfCurSpace));
  const uint32_t lim = lds32(xf32 + kXfSrcLim);
  Walk found{kBeyond, 0};
  for (; src < lim; ++src) {
    uint32_t m;
    const ByteAction a = classify_byte(src0[src], ci, ip, ew, &m);
    if (a == kSkip) continue;
    if (a == kSpace) {
      if (in_space) continue;
      in_space = 1;
    } else {
      in_space = 0;
    }
    if (found.byte != kBeyond) return found; // a kept byte follows the space at index j
    if (idx == j) {
      if (m != ' ') return Walk{m, src};
      found = Walk{m, src};
    }
    ++idx;
  }
  return Walk{kBeyond, 0};
}

// Source offset (chunk relative) of normalised index j < (number of bytes the per-byte table covers)
__device__ __forceinline__ uint32_t xf_src(const uint8_t *src0, uint32_t flags, uint32_t xf32, uint32_t j) {
  const uint32_t d = lds8(xf32 + kXfDelta + j);
  return d != 255u ? j + d : xf_walk(src0, flags, xf32, j, true).src;
}

// ---- chunk builders -----------------------------------------------------------------------------
// All take: src32 = shared-space address of the chunk's first byte in the stage buffer, I = the
// tile, cbase = the chunk's offset in the tile, priv32 = the warp's private buffer.

// plain / case-folded: 16 bytes in front + 512 + kChunkHalo behind, 16 bytes per lane
template <bool FOLD>
__device__ __forceinline__ void build_copy(const StageInfo &I, uint32_t src32, uint32_t cbase, uint32_t priv32,
                                           uint32_t lane, TileCtx &T) {
  uint4 v = lds128(src32 + 16u * lane);
  if (FOLD) v = make_uint4(fold4(v.x), fold4(v.y), fold4(v.z), fold4(v.w));
  sts128(priv32 + kTilePre + 16u * lane, v);
  if (lane < (kChunkHalo / 16) + 1) { // lanes 0..6: the bytes behind the chunk, lane 7: the 16 in front
    const bool front = lane == kChunkHalo / 16;
    const uint32_t a = front ? src32 - kTilePre : src32 + kChunkBytes + 16u * lane;
    uint4 w = lds128(a);
    if (FOLD) w = make_uint4(fold4(w.x), fold4(w.y), fold4(w.z), fold4(w.w));
    sts128(front ? priv32 : priv32 + kTilePre + kChunkBytes + 16u * lane, w);
  }
  T.sb32 = priv32;
  T.gbase = I.gbase + cbase;
  T.boff = I.boff + cbase;
  T.rem0 = I.rem0 - cbase;
  const uint32_t ns = I.nscan - cbase, st = I.staged - cbase;
  T.nscan = ns < (uint32_t)kChunkBytes ? ns : (uint32_t)kChunkBytes;
  T.staged = st < (uint32_t)(kChunkBytes + kChunkHalo) ? st : (uint32_t)(kChunkBytes + kChunkHalo);
  T.tail = I.tail;
  T.first = (I.p0 + cbase) == 0;
}

// normalising stores.  `back` = bytes in front of the chunk's first byte that are present in shared
// memory (in front of src32); P.buf + I.boff + cbase is the chunk's first source byte in global memory.  The 640
// source bytes are taken in five rounds of 128 (lane l: bytes 4l .. 4l+3 of the round), so that the
// kept bytes of a round follow those of the round before.
// `first_pass`: the chunk is built for the first time (the scan itself, not a second evaluation): its
// kept bytes are added to the window's extent when the launch asks for that.
__device__ __forceinline__ void build_xf(const ScanParams &P, const StageInfo &I, uint32_t src32, uint32_t cbase,
                                         uint32_t back, uint32_t priv32, uint32_t xf32, uint32_t lane, bool first_pass,
                                         TileCtx &T) {
  const uint32_t sf = P.store_flags;
  const bool ci = sf & kFlagIgnoreCase, ip = sf & kFlagIgnorePunct, ew = sf & kFlagElideSpace;
  const uint32_t s0 = (uint32_t)I.p0 + cbase;     // window-relative source offset of the chunk
  const uint32_t src_left = I.rem0 - cbase;       // source bytes from there to the end of the window (> 0)
  const uint8_t *g0 = P.buf + (I.boff + cbase);   // the same byte in global memory
  const uint32_t lt = (1u << lane) - 1u;

  // -- what precedes the chunk in its window: the last non-skipped source byte decides the byte in
  //    front of position 0 and whether a whitespace run is open (transform_table.c:54-78)
  uint32_t cin0 = 0, prevb = 0;
  bool have_prev = false;
  // (the `back` bytes in front of the chunk that are in shared memory first -- nearly always enough --
  // then global memory, 32 bytes per round)
  for (uint32_t done = 0; done < s0;) {
    const bool near = done < back;
    uint32_t upto = near ? back : done + 32u;
    if (upto > done + 32u) upto = done + 32u;
    if (upto > s0) upto = s0;
    const uint32_t d = done + lane;
    uint32_t c = kBeyond;
    if (d < upto) c = near ? lds8(src32 - 1u - d) : (uint32_t)g0[-1 - (long long)d];
    const uint32_t bal = __ballot_sync(kFull, c != kBeyond && !(ip && is_punct_byte(c)));
    if (bal) {
      const uint32_t cc = __shfl_sync(kFull, c, __ffs(bal) - 1);
      if (ew && is_space_byte(cc)) {
        prevb = ' ';
        cin0 = 1;
      } else {
        prevb = ci ? upper_byte(cc) : cc;
      }
      have_prev = true;
      break;
    }
    done = upto;
  }

  uint32_t avail = I.staged - cbase; // source bytes in shared memory from the chunk's first byte on
  if (avail > (uint32_t)kPrivData) avail = kPrivData;
  if (avail > src_left) avail = src_left;
  const bool reached = avail == src_left;                                  // the window ends inside what is normalised here
  const uint32_t own = src_left < (uint32_t)kChunkBytes ? src_left : (uint32_t)kChunkBytes; // source bytes of this chunk

  uint32_t cin = cin0;      // run state in front of the round (warp uniform)
  uint32_t total = 0;       // kept bytes so far (warp uniform)
  uint32_t k0 = 0;          // kept bytes among the chunk's own source bytes = its start positions
  uint32_t lastb = 0, last_src = 0;
#pragma unroll 1
  for (uint32_t rb = 0; rb < (uint32_t)kPrivData; rb += 128) {
    const uint32_t base = rb + 4u * lane;
    const uint32_t nv = avail > base ? (avail - base < 4u ? avail - base : 4u) : 0u;
    uint32_t S, K;
    const uint32_t m = xf_classify4(lds32(src32 + base), ci, ip, ew, &S, &K);
    const uint32_t valid = (1u << nv) - 1u;
    S &= valid;
    const uint32_t N = ~K & valid; // non-skipped bytes
    // "inside a whitespace run" in front of the lane's bytes: the state behind the nearest earlier
    // lane that has a non-skipped byte, else behind the round before
    const uint32_t t_has = N != 0;
    const uint32_t t_last = t_has ? (S >> (31 - __clz(N))) & 1u : 0u;
    const uint32_t bal_has = __ballot_sync(kFull, t_has), bal_last = __ballot_sync(kFull, t_last);
    uint32_t st = cin;
    if (bal_has & lt) st = (bal_last >> (31 - __clz(bal_has & lt))) & 1u;
    uint32_t keep = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t bit = 1u << k;
      if (N & bit) {
        const uint32_t sp = (S >> k) & 1u;
        if (!(sp & st)) keep |= bit;
        st = sp;
      }
    }
    // where the lane's kept bytes go: kept bytes of the earlier lanes (one ballot per byte column)
    uint32_t o = total, n_round = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t bk = __ballot_sync(kFull, (keep >> k) & 1u);
      o += __popc(bk & lt);
      n_round += __popc(bk);
    }
    {
      uint32_t dst = priv32 + kTilePre + o, ddst = xf32 + kXfDelta + o;
      uint32_t dl = base - o; // bytes dropped in front of the lane's next kept byte (source offset - index)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((keep >> k) & 1u) {
          sts8(dst, (m >> (8 * k)) & 0xFFu);
          sts8(ddst, dl < 255u ? dl : 255u);
          ++dst;
          ++ddst;
        } else {
          ++dl;
        }
      }
    }
    if (own >= rb && own < rb + 128u) { // the chunk's own bytes end in this round (or exactly in front of it)
      const uint32_t ol = (own - rb) >> 2, orest = (own - rb) & 3u;
      k0 = __shfl_sync(kFull, o + __popc(keep & ((1u << orest) - 1u)), ol);
    }
    const uint32_t bal_keep = __ballot_sync(kFull, keep != 0);
    if (bal_keep) {
      const uint32_t ll = 31 - __clz(bal_keep), hk = keep ? 31 - __clz(keep) : 0;
      lastb = __shfl_sync(kFull, (m >> (8 * hk)) & 0xFFu, ll);
      last_src = __shfl_sync(kFull, base + hk, ll);
    }
    if (bal_has) cin = (bal_last >> (31 - __clz(bal_has))) & 1u;
    total += n_round;
  }
  if (lane == 0) {
    sts8(priv32 + kTilePre - 1, prevb);
    if (first_pass && P.win_extent && k0) atomicAdd(P.win_extent + I.win, k0); // bytes the reference writes into its scratch buffer
  }

  T.sb32 = priv32;
  T.gbase = I.gbase + cbase;
  T.boff = I.boff + cbase;
  T.tail = I.tail;
  T.first = !have_prev;
  uint32_t staged, rem0 = kRemUnknown;
  if (reached) { // the window's (trimmed) end is known exactly
    staged = total - ((total && lastb == ' ') ? 1u : 0u);
    rem0 = staged;
    if (lane == 0) {
      sts32(xf32 + kXfCurIdx, staged);
      sts32(xf32 + kXfCurSrc, src_left);
      sts32(xf32 + kXfCurSpace, 0);
    }
  } else { // a last ' ' only exists if a kept byte follows: the walk decides, from that byte on
    const bool unsure = total && lastb == ' ';
    staged = total - (unsure ? 1u : 0u);
    if (lane == 0) {
      sts32(xf32 + kXfCurIdx, staged);
      sts32(xf32 + kXfCurSrc, unsure ? last_src : avail);
      sts32(xf32 + kXfCurSpace, unsure ? 0u : cin);
    }
  }
  if (lane == 0) {
    sts32(xf32 + kXfSrcLim, src_left);
    sts32(xf32 + kXfCin0, cin0);
  }
  __syncwarp();
  // every start position needs its first 8 bytes in the buffer (or the window's end known): append
  // what a long run of skipped bytes left missing, byte by byte (rare)
  if (!reached && k0 && staged < k0 + 8u) {
    const uint32_t upto = k0 + 8u;
    uint32_t end = kRemUnknown;
    if (lane == 0) {
      const uint32_t first = staged;
      for (uint32_t j = first; j < upto; ++j) {
        const Walk w = xf_walk(g0, sf, xf32, j, false);
        if (w.byte == kBeyond) {
          end = j;
          break;
        }
        sts8(priv32 + kTilePre + j, w.byte);
        sts32(xf32 + kXfExt + 4u * (j - first), w.src);
      }
    }
    end = __shfl_sync(kFull, end, 0);
    if (end != kRemUnknown) {
      staged = end;
      rem0 = end;
    } else {
      staged = upto;
    }
    __syncwarp();
  }
  T.staged = staged;
  T.rem0 = rem0;
  T.nscan = k0 < staged ? k0 : staged;
}

} // namespace dev
} // namespace olm
