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

// Per 4 MiB source window (stores with a transform flag).  Written by transform.cu on the device;
// only launches that need it get one (case-folding-only stores: always, for norm_len; launches
// with word_boundary over a store with 2..4 byte patterns: for the tail byte).
struct WindowDesc {
  uint32_t norm_len; // M_w: bytes of the normalised window (after the trailing-space trim)
  uint32_t extent;   // bytes the reference writes into its scratch buffer (M_w, +1 if a space was trimmed)
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
  kWindowMode = 1u << 6,  // the store has a transform flag: 4 MiB source windows, normalised chunk by chunk inside the scan
  kIdentityMap = 1u << 7, // window mode without compaction (case folding only)
  kCountAll = 1u << 8,    // exact statistics: count the short candidates kLongestOnly skips (stats.cuh)
};

// Geometry of the kernel (compile-time; DESIGN.md "scan kernel").
constexpr int kScanWarps = 31;                      // warps that scan
constexpr int kScanThreads = 1024;                  // + producer warp (31: tickets, TMA)
// (experiment knobs; only the defaults are covered by the GPU tests -- 2 KiB / 1 KiB tiles were measured on
// the small stores, profiles/r2_tile_size_variants.log, and their 1 M pattern launch failed)
#ifndef OLM_TILE_BYTES
#define OLM_TILE_BYTES 4096
#endif
#ifndef OLM_MAX_STAGES
#define OLM_MAX_STAGES 16
#endif
constexpr int kTileBytes = OLM_TILE_BYTES;          // positions per tile
constexpr int kTilePre = 16;                        // bytes staged in front of a tile (previous byte)
constexpr int kTileHalo = 128;                      // bytes staged behind a tile
constexpr int kStageBytes = kTilePre + kTileBytes + kTileHalo; // 4240
constexpr int kChunkBytes = 512;                    // 32 lanes x 16 bytes: the unit a warp grabs
constexpr int kTileChunks = kTileBytes / kChunkBytes; // 16
constexpr int kMaxStages = OLM_MAX_STAGES;          // ring of tile buffers (a power of two >= 2)
constexpr int kInfoRing = 2 * kMaxStages;           // tile descriptions in shared memory (>= 2 * kMaxStages)
constexpr uint32_t kChunkCapMin = 32;               // staged matches per chunk: at least ...
constexpr uint32_t kChunkCapMax = 1024;             // ... at most
constexpr int kQ1Bytes = kScanWarps * kChunkBytes * 2; // candidate queues: u16 per position of a chunk
constexpr int kQ2Entries = 96;                      // hit queue per warp (u64 entries): 31 left over + 2 x 32 new
constexpr int kQ2Bytes = kScanWarps * kQ2Entries * 8;
constexpr int kSmemHeader = 16 * kMaxStages + 72 * kInfoRing + 128; // barriers, counters, stage infos (a multiple of 128)
constexpr uint32_t kPackLenBits = 23;               // staged entry = pos_in_chunk << 23 | len (source coordinates)
// Private chunk buffer of a scanning warp: the chunk's bytes as the matcher sees them (copied,
// case-folded or normalised), 16 bytes in front of position 0 and up to kPrivData bytes from it on.
constexpr int kChunkHalo = 112;                     // plain / case-folded chunks: bytes copied behind the 512 positions
constexpr int kPrivData = 640;                      // normalised chunks: at most 512 + 128 source bytes
constexpr int kPrivBytes = kTilePre + kPrivData + 48; // 704 (16-byte units; the hash of the last positions reads past the data)
constexpr int kXfRowBytes = 128 + kPrivData;        // per warp: state of a normalised chunk (scan_device.cuh): walk state, source offset - index of every byte
constexpr uint32_t kRemUnknown = 0x40000000u;       // normalised chunk that does not reach its window's end
constexpr uint32_t kChunkOverflow = 0x80000000u;    // ChunkDesc::count flag: records are written by redo_kernel
constexpr uint32_t kPrefixSpan = 4096;              // chunks per block of the prefix / place kernels

struct ChunkDesc;
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
  // window mode (the store has a transform flag): buf holds SOURCE bytes, tiles are 4 KiB of one window
  const WindowDesc *windows; // per window of the launch, or nullptr (see WindowDesc)
  uint32_t *win_extent;      // per window of the launch, or nullptr: the scan adds every chunk's kept bytes (transform.cu ghost_update_launch)
  uint64_t win_buf_off;  // buffer offset of the first source byte of window 0
  uint64_t win_src_base; // global source offset of window 0
  uint64_t win_src_len;  // source bytes of the launch's windows (the last one may be short)
  uint32_t tiles_per_win;
  uint32_t store_flags;  // kFlagIgnoreCase | kFlagIgnorePunct | kFlagElideSpace of the store
  // tiles
  uint32_t num_tiles;             // tiles of this launch
  ChunkDesc *chunk_desc;          // [num_tiles * kTileChunks] written by the scan
  unsigned long long *span_base;  // [ceil(chunks / kPrefixSpan)] sums, then first result index, of every span of chunks
  uint32_t *temp;                 // packed matches (pos_in_tile << 18 | len), one run per chunk
  unsigned long long temp_cap;    // entries (result capacity + slack: the warps' blocks are not filled to the end)
  unsigned long long *temp_count; // bump allocator of temp[] (may run past temp_cap: entries are then dropped)
  unsigned int *ticket;           // dynamic tile counter of this launch
  unsigned int *redo_flag;        // set when a chunk's matches did not fit its staging area (redo_kernel)
  // output
  Record *out;
  uint64_t out_cap;
  uint64_t match_ptr_base;
  unsigned long long *total;    // matches so far: start value of this launch's prefix, updated by it
  unsigned long long *counters; // hits, misses, comparisons, long hits (omega_match_stats_t)
  uint32_t flags;
  uint32_t stages;    // ring depth
  uint32_t chunk_cap; // staged matches per chunk
};

// One per 512-byte chunk, written by the warp that scanned it.
struct alignas(8) ChunkDesc {
  uint32_t count;      // exact number of matches of the chunk (| kChunkOverflow: not in temp[])
  uint32_t temp_index; // first entry of the chunk's run in temp[]
};

// temp[] is handed to the scanning warps in blocks; what a warp leaves unused is lost
constexpr uint32_t kWarpTempBlock = 1024;
inline uint64_t temp_slack_entries(int sms) { return uint64_t(sms) * kScanWarps * 2 * kWarpTempBlock; }

struct ScanGeometry {
  uint32_t stages = 0, chunk_cap = 0;
};
constexpr uint32_t kPrivStagesMax = (6u * 4096u / kTileBytes) < (uint32_t)kMaxStages ? (6u * 4096u / kTileBytes) : (uint32_t)kMaxStages; // ring depth with private chunk buffers (24 KiB of tiles)

size_t scan_smem_bytes(const DeviceStore &st, uint32_t stages, uint32_t chunk_cap, bool priv);
// chooses ring depth and staging capacity for the shared memory there is; stages == 0 if the
// filters do not fit at all
ScanGeometry scan_pick_geometry(const DeviceStore &st, size_t smem_limit);
// scan -> prefix over the chunk counts (two kernels) -> placement of the records in final order
// -> redo pass (exits at once when no chunk overflowed)
cudaError_t scan_launch(const ScanParams &p, int sms, cudaStream_t stream, uint32_t *launches);
cudaError_t scan_configure(size_t smem_limit);

} // namespace olm
