/* olm_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY (see olm_oracle.h).
 *
 * Plain-C restatement of the reference matching path.  Written for clarity, not speed:
 * one loop iteration per haystack byte, qsort instead of the 12-pass radix sort, sequential
 * filters.  Every function names the reference lines it follows (paths relative to the
 * reference checkout, omega_match/...).
 */
#include "olm_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ byte classes */

/* matcher.c:90-104  IS_WORD = [A-Za-z0-9_] */
static int is_word(uint8_t c) {
  return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_';
}
/* common.h:45-52  IS_PUNCT = ASCII punctuation except '_' */
static int is_punct(uint8_t c) {
  if (c == '_') return 0;
  return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) ||
         (c >= 123 && c <= 126);
}
/* common.h:54-57  IS_SPACE = \t \n \v \f \r ' ' plus \a \b */
static int is_space(uint8_t c) { return (c >= 7 && c <= 13) || c == ' '; }
/* matcher.c:107-109 */
static int is_line_end(uint8_t c) { return c == '\n' || c == '\r'; }
/* transform_table.c:9,25 toupper() in the C locale: only a-z change */
static uint8_t ascii_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

/* ------------------------------------------------------------------ hashes */

/* hash.h:13-20 */
static uint32_t fmix32(uint32_t g) {
  g ^= g >> 16;
  g *= 0x85ebca6bu;
  g ^= g >> 13;
  g *= 0xc2b2ae35u;
  g ^= g >> 16;
  return g;
}
/* util.h:23-26 */
static uint32_t be_gram(const uint8_t *p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

/* ------------------------------------------------------------------ the pattern set */

typedef struct {
  uint32_t key;   /* big-endian first four bytes */
  uint32_t first; /* index into lp_* arrays (sorted by key, then len desc) */
  uint32_t count;
} bucket_t;

struct olm_oracle {
  uint32_t flags; /* common.h:22-24: bit1 CI, bit2 IP, bit3 EW */
  uint32_t smallest, largest;
  /* long patterns (len >= 5), grouped per bucket, len descending inside a bucket */
  uint32_t n_long;
  uint64_t *lp_off;
  uint32_t *lp_len;
  uint8_t *store;
  uint64_t store_size;
  /* gram -> bucket, open addressing */
  uint32_t n_buckets, map_cap;
  bucket_t *buckets;
  int32_t *map; /* index into buckets or -1 */
  /* short matcher (common.h:204-213) */
  uint8_t bitmap1[32];
  uint8_t bitmap2[8192];
  uint32_t len1, len2, len3, len4;
  uint32_t *arr3, *arr4;
  /* literal 3-probe bloom, only for the statistics counters */
  uint32_t table_size;
  uint32_t bloom_bits;
  uint64_t *bloom;
  /* transform scratch that persists across windows and calls (transform_table.c:40-51) */
  uint8_t *scratch;
  uint32_t scratch_cap;
};

static void *xcalloc(size_t n, size_t sz) {
  void *p = calloc(n ? n : 1, sz ? sz : 1);
  if (!p) abort();
  return p;
}

static uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

static void map_build(olm_oracle_t *o) {
  o->map_cap = next_pow2(o->n_buckets * 2 + 16);
  o->map = (int32_t *)xcalloc(o->map_cap, sizeof(int32_t));
  for (uint32_t i = 0; i < o->map_cap; ++i) o->map[i] = -1;
  for (uint32_t b = 0; b < o->n_buckets; ++b) {
    uint32_t h = fmix32(o->buckets[b].key) & (o->map_cap - 1);
    while (o->map[h] >= 0) h = (h + 1) & (o->map_cap - 1);
    o->map[h] = (int32_t)b;
  }
}

/* hash_table.c:91-109 + compiler.c:299-312: semantically an exact map gram -> bucket. */
static const bucket_t *map_find(const olm_oracle_t *o, uint32_t key) {
  if (!o->n_buckets) return NULL;
  uint32_t h = fmix32(key) & (o->map_cap - 1);
  while (o->map[h] >= 0) {
    const bucket_t *b = &o->buckets[o->map[h]];
    if (b->key == key) return b;
    h = (h + 1) & (o->map_cap - 1);
  }
  return NULL;
}

/* bloom.c:37-49 */
static void bloom_add(olm_oracle_t *o, uint32_t key) {
  const uint32_t h1 = fmix32(key), h2 = key * 0x9e3779b1u, mask = o->bloom_bits - 1;
  for (uint32_t i = 0; i < 3; ++i) {
    const uint32_t bp = (h1 + i * h2) & mask;
    o->bloom[bp >> 6] |= 1ull << (bp & 63);
  }
}
/* bloom.c:51-64 */
static int bloom_query(const olm_oracle_t *o, uint32_t key) {
  const uint32_t h1 = fmix32(key), h2 = key * 0x9e3779b1u, mask = o->bloom_bits - 1;
  for (uint32_t i = 0; i < 3; ++i) {
    const uint32_t bp = (h1 + i * h2) & mask;
    if (!((o->bloom[bp >> 6] >> (bp & 63)) & 1)) return 0;
  }
  return 1;
}

void olm_oracle_free(olm_oracle_t *o) {
  if (!o) return;
  free(o->lp_off);
  free(o->lp_len);
  free(o->store);
  free(o->buckets);
  free(o->map);
  free(o->arr3);
  free(o->arr4);
  free(o->bloom);
  free(o->scratch);
  free(o);
}

/* ------------------------------------------------------------------ transform */

/* transform_table.c:13-34 (table) and :36-88 (apply), restated without the table:
 * whitespace test first (if EW), then punctuation (if IP), then upper-casing (if CI).
 * Whitespace runs collapse to one ' ' and look through skipped punctuation because
 * in_space is only cleared by an emitted non-space byte.  One trailing ' ' is dropped --
 * whatever produced it (:82-84). */
static uint32_t transform_core(int ci, int ip, int ew, const uint8_t *src, uint32_t len,
                               uint8_t *out, uint32_t *map, int *trimmed) {
  uint32_t j = 0;
  int in_space = 0;
  for (uint32_t i = 0; i < len; ++i) {
    const uint8_t c = src[i];
    if (ew && is_space(c)) {
      if (!in_space) {
        out[j] = ' ';
        if (map) map[j] = i;
        ++j;
        in_space = 1;
      }
      continue;
    }
    if (ip && is_punct(c)) continue;
    out[j] = ci ? ascii_upper(c) : c;
    if (map) map[j] = i;
    ++j;
    in_space = 0;
  }
  int t = 0;
  if (j > 0 && out[j - 1] == ' ') {
    --j;
    t = 1;
  }
  if (trimmed) *trimmed = t;
  return j;
}

uint32_t olm_oracle_transform(int ci, int ip, int ew, const uint8_t *src, uint32_t len,
                              uint8_t *out, uint32_t *map) {
  return transform_core(ci, ip, ew, src, len, out, map, NULL);
}

/* ------------------------------------------------------------------ building the set */

typedef struct {
  const uint8_t *p;
  uint32_t len;
  uint32_t order; /* input order */
} pat_ref_t;

static int cmp_pat_bytes(const void *a, const void *b) {
  const pat_ref_t *x = (const pat_ref_t *)a, *y = (const pat_ref_t *)b;
  if (x->len != y->len) return x->len < y->len ? -1 : 1;
  const int c = memcmp(x->p, y->p, x->len);
  if (c) return c;
  return x->order < y->order ? -1 : (x->order > y->order);
}
static int cmp_pat_order(const void *a, const void *b) {
  const pat_ref_t *x = (const pat_ref_t *)a, *y = (const pat_ref_t *)b;
  return x->order < y->order ? -1 : (x->order > y->order);
}
static int cmp_u32(const void *a, const void *b) {
  const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return x < y ? -1 : (x > y);
}
/* bucket grouping: key ascending, then length descending (compiler.c:39-45, :271) */
static int cmp_pat_bucket(const void *a, const void *b) {
  const pat_ref_t *x = (const pat_ref_t *)a, *y = (const pat_ref_t *)b;
  const uint32_t kx = be_gram(x->p), ky = be_gram(y->p);
  if (kx != ky) return kx < ky ? -1 : 1;
  if (x->len != y->len) return x->len > y->len ? -1 : 1;
  return memcmp(x->p, y->p, x->len);
}

/* Turn a de-duplicated list of normalised long patterns (input order) into buckets. */
static void install_long(olm_oracle_t *o, pat_ref_t *lp, uint32_t n) {
  qsort(lp, n, sizeof(*lp), cmp_pat_bucket);
  o->n_long = n;
  o->lp_off = (uint64_t *)xcalloc(n, sizeof(uint64_t));
  o->lp_len = (uint32_t *)xcalloc(n, sizeof(uint32_t));
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; ++i) total += lp[i].len;
  o->store = (uint8_t *)xcalloc(total + 8, 1);
  o->store_size = total;
  o->buckets = (bucket_t *)xcalloc(n, sizeof(bucket_t));
  uint64_t w = 0;
  for (uint32_t i = 0; i < n; ++i) {
    memcpy(o->store + w, lp[i].p, lp[i].len);
    o->lp_off[i] = w;
    o->lp_len[i] = lp[i].len;
    w += lp[i].len;
    const uint32_t key = be_gram(lp[i].p);
    if (o->n_buckets == 0 || o->buckets[o->n_buckets - 1].key != key) {
      o->buckets[o->n_buckets].key = key;
      o->buckets[o->n_buckets].first = i;
      o->buckets[o->n_buckets].count = 0;
      ++o->n_buckets;
    }
    ++o->buckets[o->n_buckets - 1].count;
  }
  map_build(o);
}

olm_oracle_t *olm_oracle_from_patterns(const uint8_t *buf, size_t size, int ci, int ip, int ew) {
  if (!buf || !size) return NULL;
  olm_oracle_t *o = (olm_oracle_t *)xcalloc(1, sizeof(*o));
  const int any = ci || ip || ew;
  /* compiler.c:170-178: flag bits are only recorded when a transform exists */
  if (ci) o->flags |= 1u << 1;
  if (ip) o->flags |= 1u << 2;
  if (ew) o->flags |= 1u << 3;

  /* split like compile_patterns(), compiler.c:401-415; normalise like add_pattern :203-206 */
  uint8_t *arena = (uint8_t *)xcalloc(size + 1, 1);
  size_t arena_used = 0, n_pat = 0, cap_pat = 1024;
  pat_ref_t *pats = (pat_ref_t *)xcalloc(cap_pat, sizeof(*pats));
  const uint8_t *ptr = buf, *end = buf + size;
  int bad = 0;
  while (ptr < end) {
    const uint8_t *nl = (const uint8_t *)memchr(ptr, '\n', (size_t)(end - ptr));
    if (!nl) nl = end;
    uint32_t len = (uint32_t)(nl - ptr);
    if (len > 0 && ptr[len - 1] == '\r') --len;
    if (len > 0) {
      uint8_t *dst = arena + arena_used;
      uint32_t nlen;
      if (any) {
        nlen = transform_core(ci, ip, ew, ptr, len, dst, NULL, NULL);
      } else {
        memcpy(dst, ptr, len);
        nlen = len;
      }
      if (nlen == 0) {
        bad = 1; /* reference: ABORT("short_matcher_add: invalid pattern length") */
        break;
      }
      if (n_pat == cap_pat) {
        cap_pat *= 2;
        pats = (pat_ref_t *)realloc(pats, cap_pat * sizeof(*pats));
        if (!pats) abort();
      }
      pats[n_pat].p = dst;
      pats[n_pat].len = nlen;
      pats[n_pat].order = (uint32_t)n_pat;
      ++n_pat;
      arena_used += nlen;
    }
    ptr = nl + 1;
  }
  if (bad || n_pat == 0) {
    free(arena);
    free(pats);
    olm_oracle_free(o);
    return NULL;
  }

  /* de-duplicate (dedupe_set.c:91-140; separate sets for short and long patterns, which
   * cannot collide because their lengths differ): keep the first occurrence. */
  qsort(pats, n_pat, sizeof(*pats), cmp_pat_bytes);
  size_t u = 0;
  for (size_t i = 0; i < n_pat; ++i) {
    if (u > 0 && pats[u - 1].len == pats[i].len && memcmp(pats[u - 1].p, pats[i].p, pats[i].len) == 0)
      continue;
    pats[u++] = pats[i];
  }
  n_pat = u;
  qsort(pats, n_pat, sizeof(*pats), cmp_pat_order);

  o->smallest = 0xFFFFFFFFu;
  o->largest = 0;
  o->arr3 = (uint32_t *)xcalloc(n_pat, sizeof(uint32_t));
  o->arr4 = (uint32_t *)xcalloc(n_pat, sizeof(uint32_t));
  pat_ref_t *lp = (pat_ref_t *)xcalloc(n_pat, sizeof(*lp));
  uint32_t n_long = 0;

  /* hash table growth as hash_table_insert() does it (hash_table.c:112-116): the load
   * check runs before every insert of a non-duplicate long pattern; table starts at 8192 */
  uint32_t tsize = 8192, tused = 0;
  uint32_t kcap = next_pow2((uint32_t)n_pat * 2 + 16);
  uint32_t *kset = (uint32_t *)xcalloc(kcap, sizeof(uint32_t));
  uint8_t *kocc = (uint8_t *)xcalloc(kcap, 1);

  for (size_t i = 0; i < n_pat; ++i) {
    const uint8_t *p = pats[i].p;
    const uint32_t len = pats[i].len;
    if (len < o->smallest) o->smallest = len;
    if (len > o->largest) o->largest = len;
    switch (len) { /* compiler.c:78-130 */
    case 1:
      o->bitmap1[p[0] >> 3] |= (uint8_t)(1u << (p[0] & 7));
      ++o->len1;
      break;
    case 2: {
      const uint32_t v = ((uint32_t)p[0] << 8) | p[1];
      o->bitmap2[v >> 3] |= (uint8_t)(1u << (v & 7));
      ++o->len2;
      break;
    }
    case 3:
      o->arr3[o->len3++] = ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2];
      break;
    case 4:
      o->arr4[o->len4++] = be_gram(p);
      break;
    default: {
      lp[n_long++] = pats[i];
      if ((float)(tused + 1) / (float)tsize > 0.9) tsize <<= 1;
      const uint32_t key = be_gram(p);
      uint32_t h = fmix32(key) & (kcap - 1);
      while (kocc[h] && kset[h] != key) h = (h + 1) & (kcap - 1);
      if (!kocc[h]) {
        kocc[h] = 1;
        kset[h] = key;
        ++tused;
      }
    }
    }
  }
  free(kset);
  free(kocc);
  qsort(o->arr3, o->len3, sizeof(uint32_t), cmp_u32); /* compiler.c:336-339 */
  qsort(o->arr4, o->len4, sizeof(uint32_t), cmp_u32);
  install_long(o, lp, n_long);
  free(lp);
  free(pats);
  free(arena);

  /* compiler.c:257-275: bloom sized table.size*16 bits, filled with the bucket keys */
  o->table_size = tsize;
  o->bloom_bits = tsize * 16u;
  o->bloom = (uint64_t *)xcalloc(o->bloom_bits >> 6, sizeof(uint64_t));
  for (uint32_t b = 0; b < o->n_buckets; ++b) bloom_add(o, o->buckets[b].key);
  return o;
}

/* ---- compiled store reader: matcher.c:329-432 (layout written by compiler.c:241-380) */

static uint32_t rd32(const uint8_t *p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}
static uint64_t rd64(const uint8_t *p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;
}

olm_oracle_t *olm_oracle_from_olm(const uint8_t *f, size_t size) {
  if (!f || size < 72 || memcmp(f, "0MGM4tCH", 8) != 0) return NULL;
  /* common.h:77-98 (packed, 72 bytes) */
  const uint32_t flags = rd32(f + 12);
  const uint64_t store_size = rd64(f + 16);
  const uint32_t smallest = rd32(f + 28), largest = rd32(f + 32);
  const uint32_t bloom_bytes = rd32(f + 36), buckets_bytes = rd32(f + 40);
  const uint32_t table_size = rd32(f + 44), short_bytes = rd32(f + 60);
  size_t off = 72;
  if (off + store_size + 8 + 4 > size) return NULL;
  const uint8_t *store = f + off;
  off += store_size;
  if (memcmp(f + off, "0MG8L0oM", 8) != 0) return NULL;
  off += 8;
  const uint32_t bloom_bits = rd32(f + off);
  off += 4;
  const uint8_t *bloom = f + off;
  off += bloom_bytes;
  if (off + 8 > size || memcmp(f + off, "0MG*H4sH", 8) != 0) return NULL;
  off += 8;
  off += (size_t)table_size * 4; /* the index array is not needed: buckets are walked */
  const uint8_t *blob = f + off;
  off += buckets_bytes;
  if (off + short_bytes != size) return NULL; /* matcher.c:425 */

  olm_oracle_t *o = (olm_oracle_t *)xcalloc(1, sizeof(*o));
  o->flags = flags;
  o->smallest = smallest;
  o->largest = largest;
  o->table_size = table_size;
  o->bloom_bits = bloom_bits;
  o->bloom = (uint64_t *)xcalloc((bloom_bytes + 7) / 8, 8);
  memcpy(o->bloom, bloom, bloom_bytes);

  /* walk [u32 key][u32 count][pattern_t x count] records, compiler.c:313-320 */
  uint32_t n = 0;
  for (size_t p = 0; p + 8 <= buckets_bytes;) {
    const uint32_t cnt = rd32(blob + p + 4);
    n += cnt;
    p += 8 + (size_t)cnt * 16;
  }
  pat_ref_t *lp = (pat_ref_t *)xcalloc(n, sizeof(*lp));
  uint32_t k = 0;
  for (size_t p = 0; p + 8 <= buckets_bytes;) {
    const uint32_t cnt = rd32(blob + p + 4);
    for (uint32_t j = 0; j < cnt; ++j) {
      const uint8_t *rec = blob + p + 8 + (size_t)j * 16;
      lp[k].p = store + rd64(rec);
      lp[k].len = rd32(rec + 8);
      lp[k].order = k;
      ++k;
    }
    p += 8 + (size_t)cnt * 16;
  }
  install_long(o, lp, n);
  free(lp);

  if (short_bytes) {
    const uint8_t *s = f + off;
    if (memcmp(s, "0MG5HOrT", 8) != 0) {
      olm_oracle_free(o);
      return NULL;
    }
    s += 8;
    memcpy(o->bitmap1, s, 32);
    s += 32;
    memcpy(o->bitmap2, s, 8192);
    s += 8192;
    o->len1 = rd32(s);
    o->len2 = rd32(s + 4);
    o->len3 = rd32(s + 8);
    o->len4 = rd32(s + 12);
    s += 16;
    o->arr3 = (uint32_t *)xcalloc(o->len3, 4);
    o->arr4 = (uint32_t *)xcalloc(o->len4, 4);
    memcpy(o->arr3, s, (size_t)o->len3 * 4);
    s += (size_t)o->len3 * 4;
    memcpy(o->arr4, s, (size_t)o->len4 * 4);
  }
  return o;
}

uint32_t olm_oracle_flags(const olm_oracle_t *o) { return o->flags; }
uint32_t olm_oracle_smallest(const olm_oracle_t *o) { return o->smallest; }
uint32_t olm_oracle_largest(const olm_oracle_t *o) { return o->largest; }
uint32_t olm_oracle_long_count(const olm_oracle_t *o) { return o->n_long; }
uint32_t olm_oracle_table_size(const olm_oracle_t *o) { return o->table_size; }
uint32_t olm_oracle_short_count(const olm_oracle_t *o, int len) {
  switch (len) {
  case 1: return o->len1;
  case 2: return o->len2;
  case 3: return o->len3;
  case 4: return o->len4;
  default: return 0;
  }
}

static uint64_t mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

uint64_t olm_oracle_pattern_digest(const olm_oracle_t *o) {
  uint64_t acc = 0;
  for (uint32_t i = 0; i < o->n_long; ++i) {
    uint64_t h = 0xcbf29ce484222325ull ^ o->lp_len[i];
    for (uint32_t j = 0; j < o->lp_len[i]; ++j) h = (h ^ o->store[o->lp_off[i] + j]) * 0x100000001b3ull;
    acc += mix64(h);
  }
  for (uint32_t b = 0; b < 256; ++b)
    if (o->bitmap1[b >> 3] & (1u << (b & 7))) acc += mix64(0x1000000ull + b);
  for (uint32_t v = 0; v < 65536; ++v)
    if (o->bitmap2[v >> 3] & (1u << (v & 7))) acc += mix64(0x2000000ull + v);
  for (uint32_t i = 0; i < o->len3; ++i) acc += mix64(0x300000000ull + o->arr3[i]);
  for (uint32_t i = 0; i < o->len4; ++i) acc += mix64(0x400000000ull + o->arr4[i]);
  return acc;
}

/* ------------------------------------------------------------------ matching */

typedef struct {
  olm_oracle_match_t *v;
  size_t n, cap;
} mvec_t;

static void mv_push(mvec_t *m, uint64_t off, uint32_t len) {
  if (m->n == m->cap) {
    m->cap = m->cap ? m->cap * 2 : 1024;
    m->v = (olm_oracle_match_t *)realloc(m->v, m->cap * sizeof(*m->v));
    if (!m->v) abort();
  }
  m->v[m->n].offset = off;
  m->v[m->n].len = len;
  m->v[m->n]._pad = 0;
  ++m->n;
}

/* matcher.c:258-325: LSD radix, ~len bytes first then offset bytes => offset ascending,
 * length descending.  (offset,len) pairs are unique, so any comparison sort agrees. */
static int cmp_match(const void *a, const void *b) {
  const olm_oracle_match_t *x = (const olm_oracle_match_t *)a, *y = (const olm_oracle_match_t *)b;
  if (x->offset != y->offset) return x->offset < y->offset ? -1 : 1;
  if (x->len != y->len) return x->len > y->len ? -1 : 1;
  return 0;
}

/* matcher.c:552-584 */
static void filter_longest(mvec_t *m) {
  size_t w = 0;
  for (size_t i = 0; i < m->n; ++i)
    if (w == 0 || m->v[i].offset != m->v[w - 1].offset) m->v[w++] = m->v[i];
  m->n = w;
}
static void filter_no_overlap(mvec_t *m) {
  size_t w = 0;
  for (size_t i = 0; i < m->n; ++i)
    if (w == 0 || m->v[i].offset >= m->v[w - 1].offset + m->v[w - 1].len) m->v[w++] = m->v[i];
  m->n = w;
}
/* matcher.c:587-623 */
static void finalize(mvec_t *m, int no_overlap, int longest_only) {
  qsort(m->v, m->n, sizeof(*m->v), cmp_match);
  if (longest_only) filter_longest(m);
  if (no_overlap) filter_no_overlap(m);
}

static int bsearch_u32(const uint32_t *a, uint32_t n, uint32_t key) { /* matcher.c:625-662 */
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] == key) return 1;
    if (key < a[mid]) hi = mid; else lo = mid + 1;
  }
  return 0;
}

typedef struct {
  int wb, wp, ws, ls, le;
} pred_t;

/* core_match(), matcher.c:697-895, on one buffer h[0..n).  `tail` is the byte the reference
 * reads at h[n] through the unguarded short-matcher word-boundary test (:812,:830,:848). */
static void core(const olm_oracle_t *o, const uint8_t *h, size_t n, pred_t f, uint8_t tail,
                 mvec_t *out, olm_oracle_stats_t *st) {
  const int use_sm = o->smallest <= 4;
  for (size_t pos = 0; pos < n; ++pos) {
    if (f.wb) { /* :770-776 */
      const int cw = is_word(h[pos]);
      const int pw = pos > 0 ? is_word(h[pos - 1]) : 0;
      if (cw == pw) continue;
    }
    const size_t rem = n - pos;
    const int prefix_ok = !f.wp || pos == 0 || !is_word(h[pos - 1]);      /* :195,:806 */
    const int lstart_ok = !f.ls || pos == 0 || is_line_end(h[pos - 1]);   /* :196,:807 */

    if (o->largest >= 5 && rem >= 4) { /* :782-801 */
      ++st->attempts;
      const uint32_t g = be_gram(h + pos);
      if (!bloom_query(o, g)) {
        ++st->filtered;
      } else {
        const bucket_t *b = map_find(o, g);
        if (!b) {
          ++st->misses;
        } else {
          ++st->hits;
          for (uint32_t j = 0; j < b->count; ++j) { /* scan_bucket_and_append :182-255 */
            const uint32_t len = o->lp_len[b->first + j];
            if (len > rem) continue;
            ++st->comparisons;
            if (memcmp(h + pos, o->store + o->lp_off[b->first + j], len) != 0) continue;
            const size_t e = pos + len;
            if (f.wb && e < n && is_word(h[e])) continue; /* :233 */
            if (!prefix_ok) continue;                     /* :236 */
            if (f.ws && e < n && is_word(h[e])) continue; /* :239 */
            if (!lstart_ok) continue;                     /* :244 */
            if (f.le && !(e >= n || is_line_end(h[e]))) continue; /* :247 */
            mv_push(out, pos, len);
          }
        }
      }
    }

    if (use_sm) { /* :804-880, lengths 4,3,2,1 in that order */
      for (uint32_t L = 4; L >= 1; --L) {
        const uint32_t cnt = L == 4 ? o->len4 : L == 3 ? o->len3 : L == 2 ? o->len2 : o->len1;
        if (!cnt) continue;
        if (L > 1 && rem < L) continue; /* length 1 always fits */
        int hit;
        if (L == 4) hit = bsearch_u32(o->arr4, o->len4, be_gram(h + pos));
        else if (L == 3)
          hit = bsearch_u32(o->arr3, o->len3,
                            ((uint32_t)h[pos] << 16) | ((uint32_t)h[pos + 1] << 8) | h[pos + 2]);
        else if (L == 2) {
          const uint32_t v = ((uint32_t)h[pos] << 8) | h[pos + 1];
          hit = o->bitmap2[v >> 3] & (1u << (v & 7));
        } else
          hit = o->bitmap1[h[pos] >> 3] & (1u << (h[pos] & 7));
        if (!hit) continue;
        const size_t e = pos + L;
        const uint8_t at_e = e < n ? h[e] : tail;
        int wb_ok;
        if (L == 1) wb_ok = !f.wb || e >= n || !is_word(h[e]);  /* :866 guarded */
        else wb_ok = !f.wb || !is_word(at_e);                    /* :812,:830,:848 unguarded */
        const int ws_ok = !f.ws || e >= n || !is_word(h[e]);
        const int le_ok = !f.le || e >= n || is_line_end(h[e]);
        if (wb_ok && prefix_ok && ws_ok && lstart_ok && le_ok) {
          ++st->hits;
          mv_push(out, pos, L);
        } else {
          ++st->misses;
        }
      }
    }
  }
}

#define OLM_WINDOW (4u * 1024u * 1024u) /* matcher.c:60 */

int64_t olm_oracle_match(olm_oracle_t *o, const uint8_t *hay, size_t size, int no_overlap,
                         int longest_only, int wb, int wp, int ws, int ls, int le,
                         uint8_t tail_byte, olm_oracle_match_t **out,
                         olm_oracle_stats_t *stats_accum) {
  mvec_t all = {0, 0, 0};
  olm_oracle_stats_t st = {0, 0, 0, 0, 0};
  const pred_t f = {wb, wp, ws, ls, le};
  const uint32_t tflags = o->flags & ((1u << 1) | (1u << 2) | (1u << 3));

  if (!tflags) { /* matcher.c:939-943 */
    core(o, hay, size, f, tail_byte, &all, &st);
    finalize(&all, no_overlap, longest_only);
  } else { /* matcher.c:945-1018: independent 4 MiB source windows */
    const int ci = !!(o->flags & (1u << 1)), ip = !!(o->flags & (1u << 2)),
              ew = !!(o->flags & (1u << 3));
    uint32_t *map = (uint32_t *)xcalloc(OLM_WINDOW, sizeof(uint32_t));
    for (size_t base = 0; base < size; base += OLM_WINDOW) {
      const uint32_t win = (uint32_t)((size - base) < OLM_WINDOW ? (size - base) : OLM_WINDOW);
      /* transform_table.c:40-51: scratch grows by doubling from 8192, keeps old bytes, and
       * is never cleared: byte [M] after a window is whatever an earlier window left */
      if (o->scratch_cap < win) {
        uint32_t nc = o->scratch_cap ? o->scratch_cap : 8192;
        while (nc < win) nc <<= 1;
        uint8_t *nb = (uint8_t *)xcalloc((size_t)nc + 1, 1);
        if (o->scratch) memcpy(nb, o->scratch, o->scratch_cap);
        free(o->scratch);
        o->scratch = nb;
        o->scratch_cap = nc;
      }
      const uint32_t m = transform_core(ci, ip, ew, hay + base, win, o->scratch, map, NULL);
      mvec_t w = {0, 0, 0};
      core(o, o->scratch, m, f, o->scratch[m], &w, &st);
      finalize(&w, no_overlap, longest_only); /* per window, then again globally */
      for (size_t i = 0; i < w.n; ++i) {      /* matcher.c:986-1006 */
        const uint64_t s = base + map[w.v[i].offset];
        const uint64_t e = base + map[w.v[i].offset + w.v[i].len - 1];
        mv_push(&all, s, (uint32_t)(e - s + 1));
      }
      free(w.v);
    }
    free(map);
    finalize(&all, no_overlap, longest_only);
  }
  if (stats_accum) { /* matcher.c:887-893 */
    stats_accum->hits += st.hits;
    stats_accum->misses += st.misses;
    stats_accum->filtered += st.filtered;
    stats_accum->attempts += st.attempts;
    stats_accum->comparisons += st.comparisons;
  }
  *out = all.v;
  return (int64_t)all.n;
}

void olm_oracle_free_matches(olm_oracle_match_t *m) { free(m); }

uint64_t olm_oracle_stream_digest(const olm_oracle_match_t *m, size_t n) {
  uint64_t h = 0x9e3779b97f4a7c15ull;
  for (size_t i = 0; i < n; ++i) {
    h = mix64(h ^ m[i].offset);
    h = mix64(h ^ (uint64_t)m[i].len);
  }
  return h ^ n;
}
