// transform.cu -- haystack normalisation for stores compiled with a transform flag
// (kernel K2 of SURVEY 2.1).
//
// Reference: the serial window loop of omega_list_matcher_match (omega_match/src/matcher.c:
// 945-1010) calling transform_apply (transform_table.c:36-88) once per 4 MiB SOURCE window.
// Semantics reproduced exactly:
//   * windows are independent: the "inside a whitespace run" state restarts at every window;
//   * with elide-whitespace a whitespace byte is emitted (as ' ') iff the previous NON-SKIPPED
//     byte of the window is not whitespace -- the run looks through removed punctuation;
//   * after the window one trailing ' ' is dropped from the length (the byte stays in the
//     buffer: `extent` = length + 1) -- also when the space is a literal one of a
//     case-folding-only store;
//   * `tail` = the byte the reference's unguarded short-matcher test reads at index M_w of
//     its re-used scratch buffer: ' ' after a trim, else whatever an earlier window (of this
//     call or a previous one) left there -- kept in `ghost`, a 4 MiB + 1 image of that buffer.
//
// A window is normalised by 256 CTAs in two passes over its 16 KiB blocks (an earlier version
// gave one CTA a whole window: 0.9 GB/s per window, 55 GB/s for a batch of 64 windows):
//   count   : per block, with the "in a whitespace run" carry assumed 0: kept bytes, whether the
//             block has a non-skipped byte, whether its first one is whitespace (it is dropped
//             when the real carry is 1), whether its last one is whitespace, its last kept byte;
//   resolve : per window, one warp walks the 256 block summaries: real carries, output offsets,
//             window length, trailing-space trim -> window descriptor;
//   write   : per block again, now with carry and offset known: 16 bytes per thread,
//             classification, carries inside the block by ballots, exclusive block scan of the
//             kept-byte counts, scattered stores of the kept bytes and of their source indices
//             (the transformed->original offset map).
#include "transform.cuh"

#include "olm_classes.h"

namespace olm {

namespace {

constexpr int kTfThreads = 1024;
constexpr int kTfWarps = kTfThreads / 32;
constexpr uint32_t kFull = 0xFFFFFFFFu;

constexpr uint32_t kTfBlockBytes = kTfThreads * 16;                 // 16 KiB
constexpr uint32_t kTfBlocksPerWin = kWindowBytes / kTfBlockBytes;   // 256
constexpr size_t kTfWriteSmem = (kTfBlockBytes + 32) + (kTfBlockBytes + 8) * 4; // staged bytes + staged map

// One 16 KiB block of a window.  WRITE = false: summary with carry 0 -> P.blocks[].
// WRITE = true: carry and offset from P.blocks[] (resolved), bytes and map written.
template <bool WRITE>
__global__ void __launch_bounds__(kTfThreads, 1) transform_block_kernel(TransformParams P) {
  __shared__ uint32_t s_cnt[kTfWarps];   // kept bytes per warp
  __shared__ uint32_t s_has[kTfWarps];   // warp saw a non-skipped byte
  __shared__ uint32_t s_last[kTfWarps];  // ... and the last one was whitespace
  __shared__ uint32_t s_first[kTfWarps]; // ... and the first one was whitespace
  __shared__ uint32_t s_lastb[kTfWarps]; // mapped value of the warp's last non-skipped byte

  const uint32_t win = blockIdx.x / kTfBlocksPerWin, bidx = blockIdx.x % kTfBlocksPerWin;
  const uint64_t src_base = P.src_off + (uint64_t)win * kWindowBytes;
  const uint64_t remain = P.src_len - (uint64_t)win * kWindowBytes;
  const uint32_t wlen = remain < kWindowBytes ? (uint32_t)remain : kWindowBytes;
  const uint32_t blk = bidx * kTfBlockBytes;
  TfBlock &B = P.blocks[blockIdx.x];
  if (blk >= wlen) { // past the end of a short last window
    if (!WRITE && threadIdx.x == 0) B = TfBlock{0, 0, 0, 0};
    return;
  }
  const uint8_t *src = P.src + src_base;
  uint8_t *out = P.norm + P.norm_off + (uint64_t)win * P.win_stride;
  uint32_t *map = P.map ? P.map + (uint64_t)win * kWindowBytes : nullptr;
  const bool ci = P.flags & kFlagIgnoreCase, ip = P.flags & kFlagIgnorePunct, ew = P.flags & kFlagElideSpace;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t carry_space = WRITE ? (B.flags >> 8) & 1u : 0u; // transform_table.c:54: 0 at the start of a window
  const uint32_t out_base = WRITE ? B.out_base : 0u;

  const uint32_t i0 = blk + tid * 16;
  uint32_t bytes[4] = {0, 0, 0, 0};
  uint32_t nvalid = 0;
  if (i0 < wlen) {
    nvalid = wlen - i0 < 16 ? wlen - i0 : 16;
    if (nvalid == 16 && ((reinterpret_cast<uintptr_t>(src + i0) & 15) == 0)) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + i0));
      bytes[0] = v.x; bytes[1] = v.y; bytes[2] = v.z; bytes[3] = v.w;
    } else {
      for (uint32_t k = 0; k < nvalid; ++k) bytes[k >> 2] |= (uint32_t)src[i0 + k] << (8 * (k & 3));
    }
  }
  // classify; per thread: does it contain a non-skipped byte, and is the last one a space
  uint32_t act_space = 0, act_skip = 0; // bit k set: byte k is whitespace-class / skipped
  uint32_t mapped[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t c = (bytes[k >> 2] >> (8 * (k & 3))) & 0xFF;
    uint32_t m;
    const ByteAction a = classify_byte(c, ci, ip, ew, &m);
    if ((uint32_t)k < nvalid) {
      if (a == kSpace) act_space |= 1u << k;
      if (a == kSkip) act_skip |= 1u << k;
    } else {
      act_skip |= 1u << k;
    }
    mapped[k >> 2] |= m << (8 * (k & 3));
  }
  const uint32_t nonskip = ~act_skip & 0xFFFFu;
  const uint32_t t_has = nonskip != 0;
  const uint32_t t_last = t_has ? ((act_space >> (31 - __clz(nonskip))) & 1u) : 0u;

  // carry-in of this thread: state after the nearest earlier thread that has a non-skipped byte
  const uint32_t bal_has = __ballot_sync(kFull, t_has);
  const uint32_t bal_last = __ballot_sync(kFull, t_last);
  if (lane == 0) {
    s_has[warp] = bal_has != 0;
    s_last[warp] = bal_has ? ((bal_last >> (31 - __clz(bal_has))) & 1u) : 0u;
  }
  if (!WRITE) { // what the resolve pass needs about the two ends of the block
    const uint32_t t_first = t_has ? ((act_space >> (__ffs(nonskip) - 1)) & 1u) : 0u;
    const uint32_t kl = t_has ? 31 - __clz(nonskip) : 0;
    const uint32_t t_lastb = (mapped[kl >> 2] >> (8 * (kl & 3))) & 0xFFu;
    const uint32_t bal_first = __ballot_sync(kFull, t_first);
    const uint32_t src_lane = bal_has ? 31 - __clz(bal_has) : 0;
    const uint32_t w_lastb = __shfl_sync(kFull, t_lastb, src_lane);
    if (lane == 0) {
      s_first[warp] = bal_has ? ((bal_first >> (__ffs(bal_has) - 1)) & 1u) : 0u;
      s_lastb[warp] = w_lastb;
    }
  }
  __syncthreads();
  uint32_t warp_in = carry_space;
  const uint32_t wh = __ballot_sync(kFull, s_has[lane]);
  const uint32_t wl = __ballot_sync(kFull, s_last[lane]);
  {
    const uint32_t before = wh & ((1u << warp) - 1u);
    if (before) warp_in = (wl >> (31 - __clz(before))) & 1u;
  }
  uint32_t in_space = warp_in;
  {
    const uint32_t before = bal_has & ((1u << lane) - 1u);
    if (before) in_space = (bal_last >> (31 - __clz(before))) & 1u;
  }
  // keep mask (transform_table.c:56-78)
  uint32_t keep = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t bit = 1u << k;
    if (act_skip & bit) continue;
    if (act_space & bit) {
      if (!in_space) keep |= bit;
      in_space = 1;
    } else {
      keep |= bit;
      in_space = 0;
    }
  }
  const uint32_t cnt = __popc(keep);
  // block exclusive scan of cnt
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) s_cnt[warp] = incl;
  __syncthreads();
  uint32_t wsum = s_cnt[lane], wincl = wsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, wincl, d);
    if (lane >= (uint32_t)d) wincl += t;
  }
  const uint32_t warp_excl = __shfl_sync(kFull, wincl - wsum, warp);
  const uint32_t block_total = __shfl_sync(kFull, wincl, 31);
  if (!WRITE) {
    if (tid == 0) {
      TfBlock b;
      b.count = block_total; // with carry 0
      b.out_base = 0;
      b.flags = (wh != 0 ? 1u : 0u);
      if (wh) {
        b.flags |= ((wl >> (31 - __clz(wh))) & 1u) << 1;           // last non-skipped byte is whitespace
        b.flags |= (s_first[__ffs(wh) - 1] & 1u) << 2;             // first non-skipped byte is whitespace
        b.flags |= (s_lastb[31 - __clz(wh)] & 0xFFu) << 16;        // mapped value of the last non-skipped byte
      }
      b._pad = 0;
      B = b;
    }
    return;
  }
  // Kept bytes and their source indices go through shared memory so that the global stores are
  // 16-byte vectors: the staging offset is chosen congruent to the global offset mod 16 bytes.
  extern __shared__ __align__(16) uint8_t tf_smem[];
  uint8_t *s_bytes = tf_smem;                                                   // kTfBlockBytes + 32
  uint32_t *s_map = reinterpret_cast<uint32_t *>(tf_smem + kTfBlockBytes + 32); // kTfBlockBytes + 8 entries
  const uint32_t ab = out_base & 15u, am = out_base & 3u; // alignment of the block's first byte / map entry
  {
    uint32_t o = warp_excl + (incl - cnt); // block-relative
    uint32_t kk = keep;
    while (kk) {
      const uint32_t k = __ffs(kk) - 1;
      kk &= kk - 1;
      s_bytes[ab + o] = (uint8_t)(mapped[k >> 2] >> (8 * (k & 3)));
      if (map) s_map[am + o] = i0 + k;
      ++o;
    }
  }
  __syncthreads();
  {
    // bytes: global range [out_base, out_base + block_total) = staging range [ab, ab + block_total)
    uint8_t *gdst = out + (out_base - ab); // 16-byte aligned
    const uint32_t end = ab + block_total;
    for (uint32_t v = tid * 16; v < end; v += kTfThreads * 16) {
      if (v >= ab && v + 16 <= end) {
        *reinterpret_cast<uint4 *>(gdst + v) = *reinterpret_cast<const uint4 *>(s_bytes + v);
      } else {
        for (uint32_t j = v < ab ? ab : v; j < v + 16 && j < end; ++j) gdst[j] = s_bytes[j];
      }
    }
    if (map) {
      uint32_t *mdst = map + (out_base - am); // 16-byte aligned
      const uint32_t mend = am + block_total;
      for (uint32_t v = tid * 4; v < mend; v += kTfThreads * 4) {
        if (v >= am && v + 4 <= mend) {
          *reinterpret_cast<uint4 *>(mdst + v) = *reinterpret_cast<const uint4 *>(s_map + v);
        } else {
          for (uint32_t j = v < am ? am : v; j < v + 4 && j < mend; ++j) mdst[j] = s_map[j];
        }
      }
    }
  }
}

// One warp per window walks the block summaries: carry into every block, its output offset,
// the window's length, the trailing-space trim (transform_table.c:82-84), the descriptor.
__global__ void transform_resolve_kernel(TransformParams P, uint32_t n_windows) {
  const uint32_t win = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (win >= n_windows) return;
  TfBlock *blocks = P.blocks + (size_t)win * kTfBlocksPerWin;
  uint32_t carry = 0, base = 0, lastb = 0, any = 0;
  for (uint32_t b0 = 0; b0 < kTfBlocksPerWin; b0 += 32) {
    TfBlock b = blocks[b0 + lane];
    const uint32_t has = b.flags & 1u, last_sp = (b.flags >> 1) & 1u, first_sp = (b.flags >> 2) & 1u;
    // carry into lane's block: state after the nearest earlier block with a non-skipped byte
    const uint32_t bal_has = __ballot_sync(kFull, has), bal_last = __ballot_sync(kFull, last_sp);
    uint32_t cin = carry;
    const uint32_t before = bal_has & ((1u << lane) - 1u);
    if (before) cin = (bal_last >> (31 - __clz(before))) & 1u;
    // a whitespace run that continues across the block edge: its first byte here is not kept
    const uint32_t cnt = b.count - ((has && first_sp && cin) ? 1u : 0u);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    b.out_base = base + incl - cnt;
    b.flags = (b.flags & ~0x100u) | (cin << 8);
    blocks[b0 + lane] = b;
    base += __shfl_sync(kFull, incl, 31);
    if (bal_has) {
      const uint32_t l = 31 - __clz(bal_has);
      carry = (bal_last >> l) & 1u;
      lastb = __shfl_sync(kFull, (b.flags >> 16) & 0xFFu, l);
      any = 1;
    }
  }
  if (lane == 0) {
    const bool ew = P.flags & kFlagElideSpace;
    // the last byte of the normalised window: ' ' when the window ends in an elided run,
    // else the mapped value of its last non-skipped byte
    const uint32_t out_last = any ? ((ew && carry) ? (uint32_t)' ' : lastb) : 0u;
    uint32_t m = base;
    if (m > 0 && out_last == ' ') --m;
    WindowDesc d;
    d.norm_len = m;
    d.extent = base;
    d.tail = (m != base) ? (uint32_t)' ' : 0xFFFFFFFFu; // resolved by window_tails_kernel
    d._pad = 0;
    P.windows[win] = d;
  }
}

// tail(w) for windows that were not trimmed: the byte at index M_w left behind by the most
// recent earlier window whose written extent exceeds M_w, else the ghost image.
__global__ void window_tails_kernel(TransformParams P, uint32_t n_windows) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (uint32_t w = 0; w < n_windows; ++w) {
    WindowDesc d = P.windows[w];
    if (d.tail != 0xFFFFFFFFu) continue;
    const uint32_t m = d.norm_len;
    uint32_t t = P.ghost[m];
    for (uint32_t v = w; v-- > 0;) {
      if (P.windows[v].extent > m) {
        t = P.norm[P.norm_off + (uint64_t)v * P.win_stride + m];
        break;
      }
    }
    P.windows[w].tail = t;
  }
}

// ghost[i] <- byte i of the last window of this batch whose extent exceeds i (older content
// survives elsewhere), i.e. the state of the reference's scratch buffer after these windows.
__global__ void ghost_update_kernel(TransformParams P, uint32_t n_windows) {
  const uint32_t span = (kWindowBytes + gridDim.x) / gridDim.x;
  const uint32_t lo = blockIdx.x * span;
  uint32_t hi = lo + span;
  if (hi > kWindowBytes + 1) hi = kWindowBytes + 1;
  // walk the windows from last to first; `done_to` = indices below it are final
  uint32_t done_to = lo;
  for (uint32_t v = n_windows; v-- > 0 && done_to < hi;) {
    uint32_t ext = P.windows[v].extent;
    if (ext > hi) ext = hi;
    if (ext > done_to) {
      const uint8_t *srcw = P.norm + P.norm_off + (uint64_t)v * P.win_stride;
      for (uint32_t i = done_to + threadIdx.x; i < ext; i += blockDim.x) P.ghost[i] = srcw[i];
      done_to = ext;
    }
  }
}

// Case folding only: no compaction, identity offset map.  Streams the whole batch at once.
__global__ void fold_case_kernel(TransformParams P, uint32_t n_windows) {
  const uint64_t total = P.src_len;
  const uint64_t n16 = (total + 15) / 16;
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n16;
       v += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = v * 16;
    const uint32_t win = (uint32_t)(i / kWindowBytes);
    const uint32_t wi = (uint32_t)(i % kWindowBytes);
    const uint8_t *s = P.src + P.src_off + i;
    uint8_t *o = P.norm + P.norm_off + (uint64_t)win * P.win_stride + wi;
    uint32_t w[4];
    const uint32_t nvalid = total - i < 16 ? (uint32_t)(total - i) : 16;
    if (nvalid == 16 && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) {
      const uint4 q = __ldg(reinterpret_cast<const uint4 *>(s));
      w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
    } else {
      w[0] = w[1] = w[2] = w[3] = 0;
      for (uint32_t k = 0; k < nvalid; ++k) w[k >> 2] |= (uint32_t)s[k] << (8 * (k & 3));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // SWAR: bytes in 'a'..'z' get bit 5 cleared
      const uint32_t x = w[j];
      const uint32_t hi7 = x & 0x7F7F7F7Fu;
      const uint32_t ge_a = hi7 + 0x1F1F1F1Fu;          // bit7 set iff (x&0x7f) >= 'a' (0x61)
      const uint32_t gt_z = hi7 + 0x05050505u;          // bit7 set iff (x&0x7f) >  'z' (0x7a)
      const uint32_t is_lower = ge_a & ~gt_z & ~x & 0x80808080u;
      w[j] = x ^ (is_lower >> 2);
    }
    *reinterpret_cast<uint4 *>(o) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  // descriptors: one thread per window
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gtid < n_windows) {
    const uint64_t remain = total - (uint64_t)gtid * kWindowBytes;
    const uint32_t wlen = remain < kWindowBytes ? (uint32_t)remain : kWindowBytes;
    const uint8_t last = P.src[P.src_off + (uint64_t)gtid * kWindowBytes + wlen - 1];
    WindowDesc d;
    d.extent = wlen;
    d.norm_len = (last == ' ') ? wlen - 1 : wlen;
    d.tail = (last == ' ') ? (uint32_t)' ' : 0xFFFFFFFFu;
    d._pad = 0;
    P.windows[gtid] = d;
  }
}

} // namespace

cudaError_t transform_launch(const TransformParams &p, uint32_t n_windows, bool need_tails, int sms,
                             cudaStream_t stream, uint32_t *launches) {
  if (n_windows == 0) return cudaSuccess;
  static bool configured = false; // the write pass stages 80 KB in dynamic shared memory
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(transform_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kTfWriteSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const bool fold_only = (p.flags & kFlagAnyTransform) == kFlagIgnoreCase;
  if (fold_only) {
    fold_case_kernel<<<sms * 4, 512, 0, stream>>>(p, n_windows);
  } else {
    transform_block_kernel<false><<<n_windows * kTfBlocksPerWin, kTfThreads, 0, stream>>>(p);
    transform_resolve_kernel<<<(n_windows + 3) / 4, 128, 0, stream>>>(p, n_windows);
    transform_block_kernel<true><<<n_windows * kTfBlocksPerWin, kTfThreads, kTfWriteSmem, stream>>>(p);
    *launches += 2;
  }
  ++*launches;
  window_tails_kernel<<<1, 32, 0, stream>>>(p, n_windows);
  ++*launches;
  if (need_tails) {
    ghost_update_kernel<<<sms, 256, 0, stream>>>(p, n_windows);
    ++*launches;
  }
  return cudaGetLastError();
}

} // namespace olm
