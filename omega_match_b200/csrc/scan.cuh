// scan.cuh -- launch interface of the scan kernels (scan.cu) and the record layout in HBM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "device_tables.h"

namespace olm {

// One match in HBM.  Byte-for-byte the reference's omega_match_result_t
// (omega_match/include/omega/list_matcher.h:19-23: size_t offset; uint32_t len; pointer),
// so the host API can copy results out without a conversion pass.
struct alignas(8) Record {
  uint64_t offset;
  uint32_t len;
  uint32_t _pad;
  uint64_t ptr;
};
static_assert(sizeof(Record) == 24, "Record must match omega_match_result_t");

// Per normalised window (stores with a transform flag).  Written by transform.cu on the device.
struct WindowDesc {
  uint32_t norm_len; // M_w: bytes of the normalised window (after the trailing-space trim)
  uint32_t extent;   // bytes written into the window's buffer (M_w, +1 if a space was trimmed)
  uint32_t tail;     // the byte the reference would read at index M_w
  uint32_t _pad;
};

enum ScanFlags : uint32_t {
  kWordBoundary = 1u << 0,
  kWordPrefix = 1u << 1,
  kWordSuffix = 1u << 2,
  kLineStart = 1u << 3,
  kLineEnd = 1u << 4,
  kLongestOnly = 1u << 5,
  kWindowMode = 1u << 6,  // segments are normalised 4 MiB windows described by `windows`
  kIdentityMap = 1u << 7, // window mode without an offset map (case folding only)
};

// Geometry of the kernel (compile-time; DESIGN.md "scan kernel").
constexpr int kScanWarps = 16;                      // warps that scan
constexpr int kScanThreads = (kScanWarps + 1) * 32; // + one control warp (TMA producer, look-back)
constexpr int kTileBytes = 16384;                   // positions per tile
constexpr int kTilePre = 16;                        // bytes staged in front of a tile (previous byte)
constexpr int kTileHalo = 112;                      // bytes staged behind a tile
constexpr int kStageBytes = kTilePre + kTileBytes + kTileHalo; // 16512 = 129*128
constexpr int kWarpSpan = kTileBytes / kScanWarps;  // 1024 positions per warp and tile
constexpr int kChunkBytes = 512;                    // 32 lanes x 16 bytes
constexpr int kMaxStages = 4;
constexpr uint32_t kStageCapMin = 64;               // staged matches per warp and tile: at least ...
constexpr uint32_t kStageCapMax = 1024;             // ... at most
constexpr int kQueueBytes = kScanWarps * kChunkBytes * 2; // candidate queue: u16 per position of a chunk
constexpr int kSmemHeader = 1024;                   // barriers, per-tile bookkeeping, stage infos
constexpr uint32_t kPackLenBits = 18;               // staged entry = pos_in_tile << 18 | len

struct ScanParams {
  DeviceStore st;
  // input bytes
  const uint8_t *buf;  // 16-byte aligned device buffer
  uint64_t buf_len;    // readable bytes (a multiple of 16)
  int64_t seg_buf_off; // plain/shard mode: buffer offset of segment position 0 (may be negative)
  uint64_t seg_len;    // plain/shard mode: length of the whole haystack (global N)
  uint64_t scan_begin; // first owned start position (multiple of 16)
  uint64_t scan_end;   // one past the last owned start position
  uint32_t tail_byte;  // value assumed at position seg_len (plain mode)
  // window mode
  const WindowDesc *windows;
  const uint32_t *map;   // per window: kWindowBytes entries, normalised index -> source index
  uint64_t win_stride;   // bytes between normalised windows in buf
  uint64_t win_buf_off;  // buffer offset of window 0
  uint64_t win_src_base; // source offset of window 0 (global)
  uint32_t tiles_per_win;
  // tiles
  uint32_t num_tiles;             // tiles of this launch
  uint32_t tile_base;             // global index of this launch's first tile in tile_state[]
  unsigned long long *tile_state; // decoupled look-back descriptors
  unsigned int *ticket;           // dynamic tile counter of this launch
  // tiles whose matches did not fit the staging area; rewritten by redo_kernel
  uint32_t *redo_list;     // launch-local tile indices
  unsigned int *redo_count;
  // output
  Record *out;
  uint64_t out_cap;
  uint64_t match_ptr_base;
  unsigned long long *total;    // inclusive count after the last tile of this launch
  unsigned long long *counters; // hits, misses, comparisons, long hits (omega_match_stats_t)
  uint32_t flags;
  uint32_t stages;    // ring depth
  uint32_t stage_cap; // staged matches per warp and tile (two buffers of this size per warp)
};

size_t scan_smem_bytes(const DeviceStore &st, uint32_t stages, uint32_t stage_cap);
// chooses the deepest ring that fits (and the staging capacity); returns 0 if the filters do not fit at all
uint32_t scan_pick_stages(const DeviceStore &st, size_t smem_limit, uint32_t *stage_cap);
// main pass followed by the redo pass (which exits at once when no tile overflowed)
cudaError_t scan_launch(const ScanParams &p, int sms, cudaStream_t stream, uint32_t *launches);
cudaError_t scan_configure(size_t smem_limit);

} // namespace olm
