// olm_classes.h -- byte classes used by the matcher, usable from host and device code.
//
// Reference definitions: IS_WORD omega_match/src/matcher.c:90-104, IS_PUNCT
// include/omega/details/common.h:45-52 (ASCII punctuation WITHOUT '_'), IS_SPACE
// common.h:54-57 (\t \n \v \f \r ' ' PLUS \a \b), line ends matcher.c:107-109, upper-casing
// transform_table.c:9,25 (libc toupper in the C locale: only a-z change).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define OLM_HD __host__ __device__ __forceinline__
#else
#define OLM_HD inline
#endif

namespace olm {

OLM_HD bool is_word_byte(uint32_t c) {
  const uint32_t l = c | 0x20u; // folds A-Z onto a-z; digits and '_' are tested on c itself
  return (l - 'a') < 26u || (c - '0') < 10u || c == '_';
}
OLM_HD bool is_punct_byte(uint32_t c) {
  return ((c - 33u) < 15u) || ((c - 58u) < 7u) || ((c - 91u) < 6u && c != '_') || ((c - 123u) < 4u);
}
OLM_HD bool is_space_byte(uint32_t c) { return (c - 7u) < 7u || c == ' '; }
OLM_HD bool is_line_end_byte(uint32_t c) { return c == '\n' || c == '\r'; }
OLM_HD uint32_t upper_byte(uint32_t c) { return (c - 'a') < 26u ? c - 32u : c; }

// What transform_init() (transform_table.c:13-34) decides for one byte.  The whitespace test
// comes first, then punctuation, then case folding.
enum ByteAction : int { kEmit = 0, kSkip = 1, kSpace = 2 };
OLM_HD ByteAction classify_byte(uint32_t c, bool ci, bool ip, bool ew, uint32_t *mapped) {
  if (ew && is_space_byte(c)) {
    *mapped = ' ';
    return kSpace;
  }
  if (ip && is_punct_byte(c)) {
    *mapped = c;
    return kSkip;
  }
  *mapped = ci ? upper_byte(c) : c;
  return kEmit;
}

} // namespace olm
