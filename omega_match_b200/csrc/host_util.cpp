// host_util.cpp -- file mapping helpers + the three mapping entry points of the C ABI.
#include "host_util.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/olm_b200.h"

namespace olm {

uint8_t *map_fd(int fd, size_t *size, bool sequential_hint) {
  struct stat st;
  if (fd < 0 || fstat(fd, &st) != 0) return nullptr;
  if (size) *size = size_t(st.st_size);
  if (!S_ISREG(st.st_mode) || st.st_size == 0) return nullptr;
  void *p = mmap(nullptr, size_t(st.st_size), PROT_READ, MAP_PRIVATE, fd, 0);
  if (p == MAP_FAILED) return nullptr;
  if (sequential_hint) posix_madvise(p, size_t(st.st_size), POSIX_MADV_SEQUENTIAL);
  return static_cast<uint8_t *>(p);
}

uint8_t *map_whole_file(const char *path, size_t *size, bool sequential_hint) {
  if (!path) return nullptr;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return nullptr;
  uint8_t *p = map_fd(fd, size, sequential_hint);
  close(fd);
  return p;
}

void unmap(const uint8_t *addr, size_t size) {
  if (addr && size) munmap(const_cast<uint8_t *>(addr), size);
}

uint32_t next_pow2_u32(uint32_t v) {
  uint32_t p = 1;
  while (p < v && p < 0x80000000u) p <<= 1;
  return p;
}

} // namespace olm

extern "C" {

// [ref list_matcher.h:221-222 -> util.c:207-220]
uint8_t *omega_matcher_map_file(FILE *file, size_t *size, int prefetch_sequential) {
  if (!file) return nullptr;
  return olm::map_fd(fileno(file), size, prefetch_sequential != 0);
}

// [ref list_matcher.h:231-233 -> util.c:223-241]
uint8_t *omega_matcher_map_filename(const char *filename, size_t *size, int prefetch_sequential) {
  return olm::map_whole_file(filename, size, prefetch_sequential != 0);
}

// [ref list_matcher.h:241 -> util.c:244-250]
int omega_matcher_unmap_file(const uint8_t *addr, size_t size) {
  if (!addr || size == 0) return -1;
  olm::unmap(addr, size);
  return 0;
}

} // extern "C"
