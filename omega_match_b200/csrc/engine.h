// engine.h -- one matcher on one GPU: the uploaded store, work buffers, and the sequence of
// kernels that make up a match call.  Plain C++ interface; api.cpp puts the C ABI on top.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "../../include/olm_b200.h"
#include "store.h"

namespace olm {

struct MatchFlags {
  bool no_overlap = false, longest_only = false, word_boundary = false, word_prefix = false,
       word_suffix = false, line_start = false, line_end = false;
};

// What part of which haystack a call scans.
struct ScanRange {
  const void *dev = nullptr;   // device bytes of [slice_begin, slice_begin + slice_len)
  uint64_t slice_begin = 0, slice_len = 0;
  uint64_t own_begin = 0, own_end = 0; // start positions reported by this call
  uint64_t global_size = 0;
  uint64_t match_ptr_base = 0;
};

// Large result arrays of the host API (cuda_util.cu): pinned, recycled across calls.
void *pinned_result_alloc(size_t bytes);
bool pinned_result_release(void *p); // false: not one of ours (plain malloc)
constexpr size_t kPinnedResultMin = size_t(1) << 20;

struct EngineImpl;

class Engine {
public:
  // Parses + re-stages the mapped store and uploads it.  On failure returns nullptr and
  // fills *err.
  static Engine *create(const uint8_t *file, size_t size, int device, std::string *err);
  ~Engine();

  int device() const;
  const Header &header() const;

  // Device-resident input, device-resident output (records stay in the engine's buffer).
  int match_device(const ScanRange &r, const MatchFlags &f, olm_cuda_results_t *out);
  // Host input/output: H2D, match_device, D2H into a malloc'ed array.
  omega_match_results_t *match_host(const uint8_t *haystack, size_t n, const MatchFlags &f);
  // A byte range (range.dev is ignored) whose slice lies in HOST memory: segmented H2D overlapped
  // with the scan, device-resident records.
  int match_shard_host(const uint8_t *host_slice, const ScanRange &range, const MatchFlags &f, olm_cuda_results_t *out);

  // "offset:bytes\n" for every record (the CLI's listing, main.c:89-133) as one device text buffer;
  // dev_haystack = device address of the haystack byte with offset `offset0`
  int format_records(const void *dev_records, uint64_t count, const void *dev_haystack, uint64_t offset0,
                     void **dev_text, uint64_t *text_bytes);
  int64_t no_overlap_inplace(void *dev_records, uint64_t count);
  int sort_records(void *dev_records, uint64_t count);

  // for the multi-GPU layer (multi.cpp)
  bool needs_window_tails() const;           // transforming store with 2..4 byte patterns (transform.cu)
  void *ghost_image() const;                 // device image of the reference's scratch buffer (kWindowBytes + 1 bytes) or nullptr
  void *stream() const;                      // the matcher's cudaStream_t
  void *gather_buffer(size_t bytes, size_t *cap = nullptr); // device memory for gathered records (kept until a larger request); *cap = its size
  int records_to_host(void *host_dst, const void *dev_records, uint64_t count); // asynchronous, on stream()
  int sync();                                // waits for stream()

  const olm_cuda_timing_t &timing() const;
  void collect_stats(omega_match_stats_t *accum); // adds the counters of the last call
  // While on, every call also runs stats_kernel (stats.cuh) so that collect_stats() reports the
  // reference's counters exactly; off (default): the scan's own counters, no extra kernel.
  void set_exact_stats(bool on);
  // host threads that stage PAGEABLE haystacks through pinned memory (omega_matcher_set_num_threads)
  void set_host_threads(int n);

private:
  Engine() = default;
  int stage_host(const uint8_t *src, size_t n);
  int stage_pageable(const uint8_t *src, size_t n, uint64_t nseg);
  int finish_staging();
  omega_match_results_t *match_host_spans(const uint8_t *haystack, size_t n, const MatchFlags &f, uint64_t span);
  EngineImpl *impl_ = nullptr;
};

} // namespace olm
