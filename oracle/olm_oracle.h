/* olm_oracle.h -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference algorithm behind omega_list_matcher_match()
 * (reference: omega_match/src/matcher.c:934-1019 and everything it calls).  Only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may
 * load this.  The shipped library (omega_match_b200/csrc) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
 *   - the reference's own golden files (data/matcher_found.txt, data/grep_found.txt,
 *     data/expected_*.txt) and the known-answer vectors of bindings/python/tests,
 *   - the unmodified reference compiled by `make -C oracle ref` (oracle/_ref/), on seeded
 *     random inputs over the whole flag matrix of perf_test.py:69-91.
 */
#ifndef OLM_ORACLE_H
#define OLM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct olm_oracle olm_oracle_t;

typedef struct {
  uint64_t offset;
  uint32_t len;
  uint32_t _pad;
} olm_oracle_match_t;

typedef struct {
  uint64_t hits, misses, filtered, attempts, comparisons;
} olm_oracle_stats_t;

/* Build from a newline separated pattern buffer, i.e. what compile_patterns()
 * (compiler.c:382-425) would put into a store.  Returns NULL if a pattern normalises to
 * nothing (the reference aborts there, compiler.c:126-127). */
olm_oracle_t *olm_oracle_from_patterns(const uint8_t *buf, size_t size, int case_insensitive,
                                       int ignore_punctuation, int elide_whitespace);

/* Build from the bytes of a compiled .olm store (layout: compiler.c:241-380). */
olm_oracle_t *olm_oracle_from_olm(const uint8_t *file, size_t size);

void olm_oracle_free(olm_oracle_t *o);

/* Pattern-set facts (header fields of the store, common.h:77-98). */
uint32_t olm_oracle_flags(const olm_oracle_t *o);
uint32_t olm_oracle_smallest(const olm_oracle_t *o);
uint32_t olm_oracle_largest(const olm_oracle_t *o);
uint32_t olm_oracle_long_count(const olm_oracle_t *o);
uint32_t olm_oracle_short_count(const olm_oracle_t *o, int len /*1..4*/);
uint32_t olm_oracle_table_size(const olm_oracle_t *o);
/* order-independent digest of the normalised pattern set */
uint64_t olm_oracle_pattern_digest(const olm_oracle_t *o);

/* The hot path.  Returns the number of matches and a malloc'ed array in *out (free with
 * olm_oracle_free_matches).  `tail_byte` is what the reference would read one past the end
 * of a non-transformed haystack (matcher.c:812,830,848 read haystack[pos+L] unguarded);
 * pass 0 for "caller's buffer is followed by a NUL", which is what cffi/mmap give. */
int64_t olm_oracle_match(olm_oracle_t *o, const uint8_t *haystack, size_t size, int no_overlap,
                         int longest_only, int word_boundary, int word_prefix, int word_suffix,
                         int line_start, int line_end, uint8_t tail_byte,
                         olm_oracle_match_t **out, olm_oracle_stats_t *stats_accum);

void olm_oracle_free_matches(olm_oracle_match_t *m);

/* transform_apply() restated (transform_table.c:36-88), exposed for unit tests.
 * out must hold len bytes, map (optional) len uint32.  Returns normalised length. */
uint32_t olm_oracle_transform(int case_insensitive, int ignore_punctuation, int elide_whitespace,
                              const uint8_t *src, uint32_t len, uint8_t *out, uint32_t *map);

/* Order-sensitive 64-bit digest of a match stream (used for large-scale parity). */
uint64_t olm_oracle_stream_digest(const olm_oracle_match_t *m, size_t n);

#ifdef __cplusplus
}
#endif
#endif
