// host_util.h -- small host helpers shared by the compiler, the store loader and the C ABI.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>

namespace olm {

// Read-only private mapping of a whole regular file (the reference does the same for the
// store, the pattern list and the haystack: omega_match/src/util.c:147-241).  Returns nullptr
// for missing, unreadable or empty files; never aborts.
uint8_t *map_whole_file(const char *path, size_t *size, bool sequential_hint);
uint8_t *map_fd(int fd, size_t *size, bool sequential_hint);
void unmap(const uint8_t *addr, size_t size);

uint32_t next_pow2_u32(uint32_t v);

// compiler.cpp
uint32_t normalize_bytes(bool ci, bool ip, bool ew, const uint8_t *src, uint32_t len, uint8_t *out);

} // namespace olm
