// radix_sort.cu -- stable LSD radix sort of match records (kernel K5 of SURVEY 2.1).
//
// Reference: radix_sort_matches (omega_match/src/matcher.c:258-325): twelve 8-bit passes,
// bytes of ~len first and bytes of offset second, i.e. offset ascending then length
// descending.  The single-GPU scan emits records already in that order, so this sort is not
// on its path; it serves olm_cuda_sort_records() (records gathered out of order, tests).
//
// GPU form: every warp owns a contiguous run of kWarpItems records.  Per pass
//   1. histogram: per-warp 256-bin counts -> hist[digit][warp]            (digit-major)
//   2. exclusive scan of hist (three-phase block scan over 256 * n_warps counters)
//   3. scatter: the warp walks its run in order; __match_any_sync gives each lane its rank
//      among equal digits of the same 32-record step, the warp-private running bases give
//      the rest -> stable.
// Passes whose byte is identical in all keys are skipped (an OR/AND reduction of the keys
// finds them up front).
#include "filters.cuh"

namespace olm {

namespace {

constexpr int kSThreads = 256;
constexpr int kSWarps = kSThreads / 32;
constexpr int kWarpItems = 2048;
constexpr uint32_t kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t digit_of(const Record &r, int pass) {
  return pass < 4 ? ((~r.len) >> (8 * pass)) & 0xFFu : (uint32_t)(r.offset >> (8 * (pass - 4))) & 0xFFu;
}

// OR / AND of offset and ~len over all records: red[0]=or(offset) red[1]=and(offset) red[2]=or(~len) red[3]=and(~len)
__global__ void __launch_bounds__(kSThreads) key_bits_kernel(const Record *r, uint64_t n, unsigned long long *red) {
  unsigned long long o_or = 0, o_and = ~0ull, l_or = 0, l_and = ~0ull;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long off = r[i].offset, nl = (uint32_t)~r[i].len;
    o_or |= off;
    o_and &= off;
    l_or |= nl;
    l_and &= nl;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    o_or |= __shfl_xor_sync(kFull, o_or, d);
    o_and &= __shfl_xor_sync(kFull, o_and, d);
    l_or |= __shfl_xor_sync(kFull, l_or, d);
    l_and &= __shfl_xor_sync(kFull, l_and, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicOr(red + 0, o_or);
    atomicAnd(red + 1, o_and);
    atomicOr(red + 2, l_or);
    atomicAnd(red + 3, l_and);
  }
}

__global__ void __launch_bounds__(kSThreads) sort_hist_kernel(const Record *in, uint64_t n, uint32_t n_warps, int pass,
                                                              uint32_t *hist) {
  __shared__ uint32_t bins[kSWarps][256];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * kSWarps + warp;
  for (int d = lane; d < 256; d += 32) bins[warp][d] = 0;
  __syncwarp();
  if (gw < n_warps) {
    const uint64_t base = (uint64_t)gw * kWarpItems;
    for (int s = 0; s < kWarpItems; s += 32) {
      const uint64_t i = base + s + lane;
      if (i < n) atomicAdd(&bins[warp][digit_of(in[i], pass)], 1u);
    }
    __syncwarp();
    for (int d = lane; d < 256; d += 32) hist[(uint64_t)d * n_warps + gw] = bins[warp][d];
  }
}

// ---- exclusive sum scan over m uint32 counters, three phases
constexpr int kScanItems = 16;
__device__ uint32_t block_excl_sum(uint32_t v, uint32_t *total) {
  __shared__ uint32_t s_w[kSWarps];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  __syncthreads();
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  uint32_t pre = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kSWarps; ++w) {
    if ((uint32_t)w < warp) pre += s_w[w];
    all += s_w[w];
  }
  *total = all;
  return pre + incl - v;
}
__global__ void __launch_bounds__(kSThreads) u32_block_sums_kernel(const uint32_t *a, uint64_t m, uint32_t *sums) {
  const uint64_t base = ((uint64_t)blockIdx.x * kSThreads + threadIdx.x) * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < m) s += a[base + k];
  uint32_t total;
  block_excl_sum(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kSThreads) u32_scan_sums_kernel(uint32_t *sums, uint32_t nb) {
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nb; base += kSThreads) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nb ? sums[i] : 0;
    uint32_t total;
    const uint32_t ex = block_excl_sum(v, &total);
    if (i < nb) sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kSThreads) u32_apply_kernel(uint32_t *a, uint64_t m, const uint32_t *sums) {
  const uint64_t base = ((uint64_t)blockIdx.x * kSThreads + threadIdx.x) * kScanItems;
  uint32_t v[kScanItems], s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < m ? a[base + k] : 0;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = sums[blockIdx.x] + block_excl_sum(s, &total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < m) {
      a[base + k] = run;
      run += v[k];
    }
}

__global__ void __launch_bounds__(kSThreads) sort_scatter_kernel(const Record *in, Record *out, uint64_t n,
                                                                 uint32_t n_warps, int pass, const uint32_t *hist) {
  __shared__ uint32_t bases[kSWarps][256];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * kSWarps + warp;
  if (gw >= n_warps) return;
  for (int d = lane; d < 256; d += 32) bases[warp][d] = hist[(uint64_t)d * n_warps + gw];
  __syncwarp();
  const uint64_t base = (uint64_t)gw * kWarpItems;
  for (int s = 0; s < kWarpItems; s += 32) {
    const uint64_t i = base + s + lane;
    const bool valid = i < n;
    const uint32_t act = __ballot_sync(kFull, valid);
    if (!act) break;
    if (valid) {
      const Record r = in[i];
      const uint32_t d = digit_of(r, pass);
      const uint32_t peers = __match_any_sync(act, d);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      const uint32_t b = bases[warp][d];
      __syncwarp(act);
      if (rank == 0) bases[warp][d] = b + __popc(peers);
      __syncwarp(act);
      out[b + rank] = r;
    }
  }
}

} // namespace

size_t sort_scratch_bytes(uint64_t n) {
  const uint64_t n_warps = (n + kWarpItems - 1) / kWarpItems;
  const uint64_t m = n_warps * 256;
  const uint64_t nb = (m + kSThreads * kScanItems - 1) / (kSThreads * kScanItems);
  return size_t(m * 4 + (nb + 1) * 4 + 256);
}

cudaError_t sort_records_launch(Record *data, Record *tmp, uint64_t n, void *scratch, cudaStream_t st,
                                uint32_t *launches) {
  if (n < 2) return cudaSuccess;
  if (n >= (1ull << 32)) return cudaErrorInvalidValue; // counters are 32-bit, as in the reference (uint32 count)
  const uint32_t n_warps = uint32_t((n + kWarpItems - 1) / kWarpItems);
  const uint64_t m = uint64_t(n_warps) * 256;
  const uint32_t nb = uint32_t((m + kSThreads * kScanItems - 1) / (kSThreads * kScanItems));
  uint32_t *hist = static_cast<uint32_t *>(scratch);
  uint32_t *sums = hist + m;
  unsigned long long *red = reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(sums + nb + 1) + 63) & ~uintptr_t(63));

  // which key bytes vary at all?
  const unsigned long long init[4] = {0ull, ~0ull, 0ull, ~0ull};
  cudaError_t e = cudaMemcpyAsync(red, init, sizeof init, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  key_bits_kernel<<<256, kSThreads, 0, st>>>(data, n, red);
  ++*launches;
  unsigned long long bits[4];
  if ((e = cudaMemcpyAsync(bits, red, sizeof bits, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
  const unsigned long long off_var = bits[0] ^ bits[1], len_var = (bits[2] ^ bits[3]) & 0xFFFFFFFFull;

  Record *src = data, *dst = tmp;
  const uint32_t warp_blocks = (n_warps + kSWarps - 1) / kSWarps;
  for (int pass = 0; pass < 12; ++pass) { // matcher.c:275-316: ~len bytes 0..3, then offset bytes 0..7
    const unsigned long long var = pass < 4 ? (len_var >> (8 * pass)) & 0xFF : (off_var >> (8 * (pass - 4))) & 0xFF;
    if (!var) continue;
    sort_hist_kernel<<<warp_blocks, kSThreads, 0, st>>>(src, n, n_warps, pass, hist);
    u32_block_sums_kernel<<<nb, kSThreads, 0, st>>>(hist, m, sums);
    u32_scan_sums_kernel<<<1, kSThreads, 0, st>>>(sums, nb);
    u32_apply_kernel<<<nb, kSThreads, 0, st>>>(hist, m, sums);
    sort_scatter_kernel<<<warp_blocks, kSThreads, 0, st>>>(src, dst, n, n_warps, pass, hist);
    *launches += 5;
    std::swap(src, dst);
  }
  if (src != data) {
    if ((e = cudaMemcpyAsync(data, src, n * sizeof(Record), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

} // namespace olm
