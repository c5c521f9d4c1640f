// transform.cu -- per-window facts of a normalising store that the scan cannot know on its own.
//
// The normalisation itself (transform_apply, transform_table.c:36-88) happens inside the scan,
// chunk by chunk (scan_device.cuh build_xf / build_copy); no normalised copy of the haystack and no
// offset map exist.  What is left for this file is what the reference derives from a WHOLE 4 MiB
// window before it scans it (matcher.c:945-1010), and only the launches that need it pay for it:
//
//   * case-folding-only stores: M_w = window length minus the trailing-space trim
//     (transform_table.c:82-84) -- one thread per window, one byte read;
//   * stores with 2..4 byte patterns: `tail` = the byte the reference's unguarded short-matcher
//     test reads at index M_w of its re-used scratch buffer: ' ' after a trim, else
//     whatever an earlier window (of this call or of a previous one) left there.  `ghost` is a
//     4 MiB + 1 image of that buffer, brought up to date after every launch.  For stores that drop
//     bytes this needs the kept-byte count of every window:
//       count   : per 4 KiB block of a window, with the "in a whitespace run" carry assumed 0: kept
//                 bytes, whether the block has a non-skipped byte, whether its first / last one is
//                 whitespace, its last kept byte;
//       resolve : per window, one warp walks the block summaries: real carries, first normalised
//                 index of every block, window length, trailing-space trim -> window descriptor;
//       tails   : per window that was not trimmed: the normalised byte at index M_w of the most
//                 recent earlier window that is longer, found by walking that window's block;
//       ghost   : the parts of the launch's windows that are visible in the scratch buffer
//                 afterwards (the last window, then every earlier one that is longer than all later
//                 ones) are normalised once more, block by block, into the image.
#include "transform.cuh"

#include "olm_classes.h"

namespace olm {

namespace {

constexpr int kTfThreads = 256;
constexpr int kTfWarps = kTfThreads / 32;
constexpr uint32_t kFull = 0xFFFFFFFFu;

constexpr uint32_t kTfBlockBytes = kTfThreads * 16;                 // 4 KiB
constexpr uint32_t kTfBlocksPerWin = kWindowBytes / kTfBlockBytes;   // 1024
static_assert(kTfBlocksPerWin == kTfBlocksPerWindow, "transform.cuh");
constexpr uint32_t kTfUnresolved = 0xFFFFFFFFu;

// One 4 KiB block of a window, 16 source bytes per thread.
//   kCount: summary with carry 0 -> P.blocks[] (all blocks, or -- `visible` given -- only those of
//           windows that have a visible part);
//   kWrite: carry and first index from P.blocks[] (resolved); the block's kept bytes whose
//           normalised index lies in the visible part [lo, hi) of its window go to P.ghost;
//   kPick : CTA w evaluates the block plan[w] names and stores the normalised byte at index
//           plan[w].z as windows[w].tail.
enum TfMode : int { kCount = 0, kWrite = 1, kPick = 2 };
template <int MODE>
__global__ void __launch_bounds__(kTfThreads) transform_block_kernel(TransformParams P, const uint2 *visible,
                                                                     const uint4 *plan, uint32_t n_windows) {
  __shared__ uint32_t s_cnt[kTfWarps];   // kept bytes per warp
  __shared__ uint32_t s_has[kTfWarps];   // warp saw a non-skipped byte
  __shared__ uint32_t s_last[kTfWarps];  // ... and the last one was whitespace
  __shared__ uint32_t s_first[kTfWarps]; // ... and the first one was whitespace
  __shared__ uint32_t s_lastb[kTfWarps]; // mapped value of the warp's last non-skipped byte

  const uint64_t n_items = MODE == kPick ? n_windows : (uint64_t)n_windows * kTfBlocksPerWin;
  for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    uint64_t gb = item;
    uint32_t pick_idx = 0;
    if (MODE == kPick) {
      const uint4 pl = plan[item];
      if (!pl.w) continue;
      gb = (uint64_t)pl.x * kTfBlocksPerWin + pl.y;
      pick_idx = pl.z;
    }
    const uint32_t win = (uint32_t)(gb / kTfBlocksPerWin), bidx = (uint32_t)(gb % kTfBlocksPerWin);
    const uint64_t src_base = P.src_off + (uint64_t)win * kWindowBytes;
    const uint64_t remain = P.src_len - (uint64_t)win * kWindowBytes;
    const uint32_t wlen = remain < kWindowBytes ? (uint32_t)remain : kWindowBytes;
    const uint32_t blk = bidx * kTfBlockBytes;
    TfBlock &B = P.blocks[gb];
    uint32_t vis_lo = 0, vis_hi = 0;
    if (MODE != kPick && visible) {
      const uint2 v = visible[win];
      vis_lo = v.x;
      vis_hi = v.y;
      if (vis_hi <= vis_lo) continue; // nothing of this window is visible
    }
    if (blk >= wlen) { // past the end of a short last window
      if (MODE == kCount && threadIdx.x == 0) B = TfBlock{0, 0, 0, 0};
      continue;
    }
    if (MODE == kWrite) { // does the block hold a visible byte at all?
      // (the resolve pass gave every block of the window its first index, empty blocks included)
      const uint32_t b0 = B.out_base, b1 = bidx + 1 < kTfBlocksPerWin ? P.blocks[gb + 1].out_base : 0xFFFFFFFFu;
      if (b0 >= vis_hi || b1 <= vis_lo) continue;
    }
    const uint8_t *src = P.src + src_base;
    const bool ci = P.flags & kFlagIgnoreCase, ip = P.flags & kFlagIgnorePunct, ew = P.flags & kFlagElideSpace;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t carry_space = MODE != kCount ? (B.flags >> 8) & 1u : 0u; // transform_table.c:54: 0 at the start of a window
    const uint32_t out_base = MODE != kCount ? B.out_base : 0u;

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
    __syncthreads(); // (the shared arrays of the previous block of this CTA have been read)
    if (lane == 0) {
      s_has[warp] = bal_has != 0;
      s_last[warp] = bal_has ? ((bal_last >> (31 - __clz(bal_has))) & 1u) : 0u;
    }
    if (MODE == kCount) { // what the resolve pass needs about the two ends of the block
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
    const uint32_t wh = __ballot_sync(kFull, lane < kTfWarps && s_has[lane]);
    const uint32_t wl = __ballot_sync(kFull, lane < kTfWarps && s_last[lane]);
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
    uint32_t wsum = lane < kTfWarps ? s_cnt[lane] : 0u, wincl = wsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, wincl, d);
      if (lane >= (uint32_t)d) wincl += t;
    }
    const uint32_t warp_excl = __shfl_sync(kFull, wincl - wsum, warp);
    const uint32_t block_total = __shfl_sync(kFull, wincl, 31);
    if (MODE == kCount) {
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
      continue;
    }
    // this thread's kept bytes by normalised index
    uint32_t o = out_base + warp_excl + (incl - cnt);
    uint32_t kk = keep;
    while (kk) {
      const uint32_t k = __ffs(kk) - 1;
      kk &= kk - 1;
      const uint32_t byte = (mapped[k >> 2] >> (8 * (k & 3))) & 0xFFu;
      if (MODE == kWrite) {
        if (o >= vis_lo && o < vis_hi) P.ghost[o] = (uint8_t)byte;
      } else if (o == pick_idx) {
        P.windows[item].tail = byte;
      }
      ++o;
    }
  }
}

// One warp per window walks the block summaries: carry into every block, its first normalised
// index, the window's length, the trailing-space trim (transform_table.c:82-84), the descriptor.
__global__ void transform_resolve_kernel(TransformParams P, uint32_t n_windows, const uint2 *visible) {
  const uint32_t win = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (win >= n_windows) return;
  if (visible && visible[win].y <= visible[win].x) return; // (only the windows the image needs)
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
  if (lane == 0 && !visible) {
    const bool ew = P.flags & kFlagElideSpace;
    // the last byte of the normalised window: ' ' when the window ends in an elided run,
    // else the mapped value of its last non-skipped byte
    const uint32_t out_last = any ? ((ew && carry) ? (uint32_t)' ' : lastb) : 0u;
    uint32_t m = base;
    if (m > 0 && out_last == ' ') --m;
    WindowDesc d;
    d.norm_len = m;
    d.extent = base;
    d.tail = (m != base) ? (uint32_t)' ' : kTfUnresolved; // resolved by window_tails_kernel
    d._pad = 0;
    P.windows[win] = d;
  }
}

// tail(w) for windows that were not trimmed: the byte at index M_w left behind by the most
// recent earlier window whose written extent exceeds M_w, else the ghost image.  One
// thread per window finds WHERE that byte is -- plan[w] = {window, block, index, 1} -- and the pick
// pass of transform_block_kernel normalises that one block to read it.
__global__ void window_tails_plan_kernel(TransformParams P, uint32_t n_windows, uint4 *plan) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_windows) return;
  plan[w] = make_uint4(0, 0, 0, 0);
  const WindowDesc d = P.windows[w];
  if (d.tail != kTfUnresolved) return;
  const uint32_t m = d.norm_len;
  for (uint32_t v = w; v-- > 0;) {
    if (P.windows[v].extent > m) {
      // the last block whose first index is <= m (first indices never decrease; blocks without a
      // kept byte share theirs with the next one, so the block found holds index m)
      const TfBlock *blocks = P.blocks + (size_t)v * kTfBlocksPerWin;
      uint32_t lo = 0, hi = kTfBlocksPerWin;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) / 2;
        if (blocks[mid].out_base <= m) lo = mid; else hi = mid;
      }
      plan[w] = make_uint4(v, lo, m, 1);
      return;
    }
  }
  P.windows[w].tail = P.ghost[m];
}

// visible[v] = the range of normalised indices of window v that are still in the scratch buffer
// after the launch's last window: [extent of the longest later window, own extent).  The extents
// come from the window descriptors or -- launches without them -- from what the scan counted.
__global__ void visible_ranges_kernel(TransformParams P, uint32_t n_windows, const uint32_t *extents, uint2 *visible) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t mx = 0;
  for (uint32_t v = n_windows; v-- > 0;) {
    const uint32_t e = extents ? extents[v] : P.windows[v].extent;
    visible[v] = make_uint2(mx, e > mx ? e : mx);
    if (e > mx) mx = e;
  }
}

// ---- case folding only: nothing is dropped, so everything follows from the window lengths ----------
__global__ void fold_windows_kernel(TransformParams P, uint32_t n_windows, bool need_tails) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_windows) return;
  const uint64_t remain = P.src_len - (uint64_t)w * kWindowBytes;
  const uint32_t wlen = remain < kWindowBytes ? (uint32_t)remain : kWindowBytes;
  const uint8_t *src = P.src + P.src_off;
  const uint8_t last = src[(uint64_t)w * kWindowBytes + wlen - 1];
  WindowDesc d;
  d.extent = wlen;
  d.norm_len = (last == ' ') ? wlen - 1 : wlen;
  d.tail = ' ';
  d._pad = 0;
  if (last != ' ' && need_tails) {
    // index wlen of the scratch buffer: written by the previous window when that one is longer
    // (every window but the last is 4 MiB long), else by an earlier call
    if (w > 0 && wlen < kWindowBytes)
      d.tail = upper_byte(src[(uint64_t)(w - 1) * kWindowBytes + wlen]);
    else
      d.tail = P.ghost[wlen];
  }
  P.windows[w] = d;
}
// the image after the launch: the last window, and behind its end what the window before it wrote
__global__ void fold_ghost_kernel(TransformParams P, uint32_t n_windows) {
  const uint32_t last = n_windows - 1;
  const uint64_t remain = P.src_len - (uint64_t)last * kWindowBytes;
  const uint32_t wlen = remain < kWindowBytes ? (uint32_t)remain : kWindowBytes;
  const uint8_t *src = P.src + P.src_off;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < kWindowBytes; i += gridDim.x * blockDim.x) {
    if (i < wlen)
      P.ghost[i] = (uint8_t)upper_byte(src[(uint64_t)last * kWindowBytes + i]);
    else if (last > 0)
      P.ghost[i] = (uint8_t)upper_byte(src[(uint64_t)(last - 1) * kWindowBytes + i]);
  }
}

} // namespace

cudaError_t window_descs_launch(const TransformParams &p, uint32_t n_windows, bool need_tails, int sms,
                                cudaStream_t stream, uint32_t *launches) {
  if (n_windows == 0) return cudaSuccess;
  const bool fold_only = (p.flags & kFlagAnyTransform) == kFlagIgnoreCase;
  if (fold_only) {
    fold_windows_kernel<<<(n_windows + 127) / 128, 128, 0, stream>>>(p, n_windows, need_tails);
    ++*launches;
    if (need_tails) {
      fold_ghost_kernel<<<sms * 2, 512, 0, stream>>>(p, n_windows);
      ++*launches;
    }
    return cudaGetLastError();
  }
  if (!need_tails) return cudaSuccess; // the scan needs nothing from here
  const uint64_t n_blocks = (uint64_t)n_windows * kTfBlocksPerWin;
  const unsigned grid = (unsigned)(n_blocks < (uint64_t)sms * 64 ? n_blocks : (uint64_t)sms * 64);
  transform_block_kernel<kCount><<<grid, kTfThreads, 0, stream>>>(p, nullptr, nullptr, n_windows);
  transform_resolve_kernel<<<(n_windows + 3) / 4, 128, 0, stream>>>(p, n_windows, nullptr);
  window_tails_plan_kernel<<<(n_windows + 63) / 64, 64, 0, stream>>>(p, n_windows, p.plan);
  transform_block_kernel<kPick><<<n_windows, kTfThreads, 0, stream>>>(p, nullptr, p.plan, n_windows);
  visible_ranges_kernel<<<1, 32, 0, stream>>>(p, n_windows, nullptr, p.visible);
  transform_block_kernel<kWrite><<<grid, kTfThreads, 0, stream>>>(p, p.visible, nullptr, n_windows);
  *launches += 6;
  return cudaGetLastError();
}

cudaError_t ghost_update_launch(const TransformParams &p, uint32_t n_windows, const uint32_t *extents, int sms,
                                cudaStream_t stream, uint32_t *launches) {
  if (n_windows == 0) return cudaSuccess;
  const uint64_t n_blocks = (uint64_t)n_windows * kTfBlocksPerWin;
  const unsigned grid = (unsigned)(n_blocks < (uint64_t)sms * 64 ? n_blocks : (uint64_t)sms * 64);
  visible_ranges_kernel<<<1, 32, 0, stream>>>(p, n_windows, extents, p.visible);
  transform_block_kernel<kCount><<<grid, kTfThreads, 0, stream>>>(p, p.visible, nullptr, n_windows);
  transform_resolve_kernel<<<(n_windows + 3) / 4, 128, 0, stream>>>(p, n_windows, p.visible);
  transform_block_kernel<kWrite><<<grid, kTfThreads, 0, stream>>>(p, p.visible, nullptr, n_windows);
  *launches += 4;
  return cudaGetLastError();
}

} // namespace olm
