"""Seeded, version-independent input generators shared by tests, golden generation and bench.

Everything is derived from splitmix64 counters (no numpy RNG), so the same bytes come out in
the build container, on the GPU box and in any later round.  The synthetic haystack and
pattern formulas are the ones SURVEY.md 8d fixes for BASELINE configs 4 and 5.
"""
from __future__ import annotations

import lzma
from functools import lru_cache
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser of (x + golden gamma), elementwise on uint64."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rand_u64(seed: int, n: int, start: int = 0) -> np.ndarray:
    with np.errstate(over="ignore"):
        return splitmix64(np.uint64(seed) + np.arange(start, start + n, dtype=np.uint64))


@lru_cache(maxsize=None)
def golden_data(name: str) -> bytes:
    """A file of the reference's data/ directory (committed xz copy)."""
    return lzma.decompress((GOLDEN / "data" / (name + ".xz")).read_bytes())


# ------------------------------------------------------------------ SURVEY 8d generators

ALPHABET64 = np.frombuffer(
    b"abcdefghijklmnopqrstuvwxyz" b"ABCDEFGHIJKLMNOPQRSTUVWXYZ" b"        " b"\n" b".,-", dtype=np.uint8)
assert ALPHABET64.size == 64
SEED_H4, SEED_H5 = 0x4F4C4D48, 0x4F4C4D16
SEED_P4, SEED_P5 = 0x4F4C4D04, 0x4F4C4D05


def synth_haystack(n: int, seed: int, start: int = 0) -> np.ndarray:
    """byte[i] = ALPHABET[(splitmix64(seed + (i>>3)) >> (8*(i&7))) & 63], i in [start, start+n)."""
    first, last = start >> 3, (start + n + 7) >> 3
    words = rand_u64(seed, last - first, first)
    b = words.view(np.uint8).reshape(-1, 8)  # little endian: byte k = bits 8k..8k+7
    out = ALPHABET64[(b & 63).reshape(-1)]
    off = start - (first << 3)
    return np.ascontiguousarray(out[off:off + n])


def synth_long_patterns(count: int, seed: int = SEED_P5, min_len: int = 6, max_len: int = 24) -> list[bytes]:
    """`count` distinct strings over a-zA-Z, length uniform in [min_len, max_len] (config 5)."""
    letters = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ", dtype=np.uint8)
    out, seen, k = [], set(), 0
    span = max_len - min_len + 1
    while len(out) < count:
        batch = max(1024, (count - len(out)) * 11 // 10)
        r = rand_u64(seed, batch * 4, k * 4).reshape(batch, 4)
        k += batch
        lens = (r[:, 0] % np.uint64(span)).astype(np.int64) + min_len
        # 24 letters from three 64-bit words, 8 letters each
        idx = np.stack([(r[:, 1 + j // 8] >> np.uint64(8 * (j % 8))) & np.uint64(0xFF) for j in range(24)], axis=1)
        chars = letters[(idx % np.uint64(52)).astype(np.int64)]
        for row, ln in zip(chars, lens):
            p = row[:ln].tobytes()
            if p not in seen:
                seen.add(p)
                out.append(p)
                if len(out) == count:
                    break
    return out


def synth_short_patterns(seed: int = SEED_P4) -> list[bytes]:
    """Config 4: tlds + 4 one-byte, 64 two-byte, 1024 three-byte, 8192 four-byte printable strings."""
    pats = [b"~", b"^", b"`", b"|"]
    r = rand_u64(seed, 64 + 1024 + 8192 + 4096)
    printable = np.arange(33, 127, dtype=np.uint8)
    k = 0
    for ln, cnt in ((2, 64), (3, 1024), (4, 8192)):
        seen = set()
        while len(seen) < cnt:
            v = int(r[k % r.size]) ^ (k // r.size)
            k += 1
            p = bytes(printable[(v >> (8 * j)) % 94] for j in range(ln))
            seen.add(p)
        pats.extend(sorted(seen))
    return golden_data("tlds.txt").split() + pats


def plant(hay: np.ndarray, patterns: list[bytes], seed: int, start: int = 0, block: int = 4096) -> np.ndarray:
    """One pattern per `block`-byte block at splitmix64(seed+block) % (block-8-len), flanked by spaces.
    `hay` holds the global bytes [start, start+len(hay)); only blocks fully inside are planted."""
    n = hay.size
    b0 = (start + block - 1) // block
    b1 = (start + n) // block
    if b1 <= b0:
        return hay
    r = rand_u64(seed, b1 - b0, b0)
    r2 = rand_u64(seed ^ 0x5555, b1 - b0, b0)
    for i in range(b1 - b0):
        p = patterns[int(r2[i] % np.uint64(len(patterns)))]
        room = block - 2 - len(p)
        at = (b0 + i) * block - start + 1 + int(r[i] % np.uint64(max(room, 1)))
        hay[at - 1] = 32
        hay[at:at + len(p)] = np.frombuffer(p, dtype=np.uint8)
        hay[at + len(p)] = 32
    return hay


# ------------------------------------------------------------------ text-like haystacks

_GLUE = [b" ", b" ", b" ", b" ", b" ", b" ", b"  ", b", ", b". ", b".\n", b"\n", b"\r\n", b"\t", b"; ", b" - ", b"-",
         b"'s ", b"! ", b"? ", b" (", b") ", b"\n\n", b" \t\n ", b"_", b"1 ", b"/"]
_WORDS = (b"the and of to in that he shall unto for his a they be is him not them it with all thou thy was god "
          b"which my me said but ye their have will thee from as are when this out were upon man by you up there "
          b"hath then people came had house into on her come one we children s before your also day land men "
          b"against shalt king James Mary John Peter Paul Martin O'Malley Maryland mayor leadership known").split()


def text_haystack(n: int, seed: int, vocab: list[bytes] | None = None) -> np.ndarray:
    """Word salad with names, mixed case, punctuation and irregular whitespace, exactly n bytes."""
    words = list(_WORDS) + (vocab or [])
    est = n // 5 + 16
    r = rand_u64(seed, est * 3).reshape(est, 3)
    parts, size = [], 0
    for i in range(est):
        w = words[int(r[i, 0] % np.uint64(len(words)))]
        style = int(r[i, 1] % np.uint64(8))
        if style == 0:
            w = w.upper()
        elif style == 1:
            w = w.capitalize()
        elif style == 2:
            w = w.lower()
        g = _GLUE[int(r[i, 2] % np.uint64(len(_GLUE)))]
        parts.append(w)
        parts.append(g)
        size += len(w) + len(g)
        if size >= n:
            break
    buf = b"".join(parts)
    while len(buf) < n:
        buf += buf
    return np.frombuffer(buf[:n], dtype=np.uint8).copy()


def sentence_haystack(n: int) -> np.ndarray:
    """perf_test.py:109-120: one English sentence repeated."""
    base = b"There was a Maryland mayor named Martin O'Malley who was known for leadership. "
    reps = n // len(base) + 1
    return np.frombuffer((base * reps)[:n], dtype=np.uint8).copy()


@lru_cache(maxsize=1)
def pseudo_kjv() -> bytes:
    """4 606 955 bytes of 0x01 with every `OFF:TEXT` line of data/matcher_found.txt
    written at OFF.  The reference reproduces matcher_found.txt / grep_found.txt on it."""
    hay = np.full(4606955, 1, dtype=np.uint8)
    for line in golden_data("matcher_found.txt").split(b"\n"):
        if line:
            off, txt = line.split(b":", 1)
            hay[int(off):int(off) + len(txt)] = np.frombuffer(txt, dtype=np.uint8)
    return hay.tobytes()


def parse_expected(name: str) -> np.ndarray:
    """`OFF:TEXT` lines of a reference expected-output file -> (offset, len) structured array."""
    rows = []
    for line in golden_data(name).split(b"\n"):
        if line:
            off, txt = line.split(b":", 1)
            rows.append((int(off), len(txt.rstrip(b"\r")) if False else len(txt)))
    out = np.zeros(len(rows), dtype=[("offset", "<u8"), ("len", "<u4"), ("_pad", "<u4")])
    if rows:
        out["offset"] = [r[0] for r in rows]
        out["len"] = [r[1] for r in rows]
    return out


# ------------------------------------------------------------------ the parity matrix

# perf_test.py:69-91: (variant name, (ignore_case, ignore_punct, elide_ws), match flags)
PERF_VARIANTS = [
    ("baseline", (0, 0, 0), ()),
    ("ignore-case", (1, 0, 0), ()),
    ("ignore-case+ignore-punct", (1, 1, 0), ()),
    ("ignore-case+ignore-punct+word-boundary", (1, 1, 0), ("word_boundary",)),
    ("ignore-case+ignore-punct+word-boundary+elide-whitespace", (1, 1, 1), ("word_boundary",)),
    ("ignore-case+no-overlap+longest", (1, 0, 0), ("no_overlap", "longest_only")),
    ("ignore-case+word-boundary", (1, 0, 0), ("word_boundary",)),
    ("ignore-punct", (0, 1, 0), ()),
    ("line-end", (0, 0, 0), ("line_end", "longest_only", "no_overlap")),
    ("line-end+ignore-case", (1, 0, 0), ("line_end", "longest_only", "no_overlap")),
    ("line-end+word-boundary", (0, 0, 0), ("line_end", "word_boundary", "longest_only", "no_overlap")),
    ("line-start", (0, 0, 0), ("line_start", "longest_only", "no_overlap")),
    ("line-start+ignore-case", (1, 0, 0), ("line_start", "longest_only", "no_overlap")),
    ("line-start+line-end", (0, 0, 0), ("line_start", "line_end", "longest_only", "no_overlap")),
    ("line-start+line-end+word-boundary", (0, 0, 0),
     ("line_start", "line_end", "word_boundary", "longest_only", "no_overlap")),
    ("longest+no-overlap", (0, 0, 0), ("longest_only", "no_overlap")),
    ("longest+no-overlap+word-boundary", (0, 0, 0), ("longest_only", "no_overlap", "word_boundary")),
    ("no-overlap+word-boundary", (0, 0, 0), ("no_overlap", "word_boundary")),
    ("word-boundary", (0, 0, 0), ("word_boundary",)),
    ("word-prefix", (0, 0, 0), ("word_prefix",)),
    ("word-suffix", (0, 0, 0), ("word_suffix",)),
]

# every match-flag set of the matrix, plus the single filters and two mixed sets
MATCH_FLAG_SETS = sorted({tuple(sorted(v[2])) for v in PERF_VARIANTS} |
                         {("longest_only",), ("no_overlap",), ("word_prefix", "word_suffix"),
                          ("line_end", "word_suffix"), ("line_start", "word_prefix")})

STORE_FLAG_SETS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1), (0, 1, 0), (0, 0, 1)]

WINDOW = 4 * 1024 * 1024


def vector_cases() -> list[dict]:
    cases = []
    for sf in STORE_FLAG_SETS:
        tag = "".join("cpw"[i] for i in range(3) if sf[i]) or "plain"
        # two and a bit windows of text, names list: window edges, trims, punctuation
        cases.append(dict(name=f"names-text-{tag}", patterns="names", store_flags=sf, hay="text",
                          size=2 * WINDOW + 70001, seed=0xA11CE + sum(sf)))
        cases.append(dict(name=f"synthshort-synth-{tag}", patterns="synth_short", store_flags=sf, hay="synth",
                          size=WINDOW + 12345, seed=SEED_H4))
    cases.append(dict(name="census-text-c", patterns="census", store_flags=(1, 0, 0), hay="text",
                      size=WINDOW + 999, seed=0xC3A5))
    cases.append(dict(name="census-text-cpw", patterns="census", store_flags=(1, 1, 1), hay="text",
                      size=WINDOW + 999, seed=0xC3A6))
    cases.append(dict(name="synthlong-synth-plain", patterns="synth_long_20k", store_flags=(0, 0, 0), hay="synth_planted",
                      size=2 * WINDOW + 777, seed=SEED_H5))
    cases.append(dict(name="names-sentence-plain", patterns="names", store_flags=(0, 0, 0), hay="sentence",
                      size=1 << 20, seed=0))
    cases.append(dict(name="tlds-email-plain", patterns="tlds", store_flags=(0, 0, 0), hay="email",
                      size=0, seed=0))
    return cases


@lru_cache(maxsize=None)
def _pattern_set(kind: str) -> tuple:
    if kind == "names":
        return tuple(l for l in golden_data("names.txt").split(b"\n") if l)
    if kind == "census":
        return tuple(l for l in golden_data("surnames_us_census.txt").split(b"\n") if l)
    if kind == "tlds":
        return tuple(golden_data("tlds.txt").split())
    if kind == "synth_short":
        return tuple(synth_short_patterns())
    if kind == "synth_long_20k":
        return tuple(synth_long_patterns(20000))
    raise KeyError(kind)


_PUNCT = set(b"!\"#$%&'()*+,-./:;<=>?@[\\]^`{|}~")
_SPACE = set(b"\t\n\v\f\r \a\b")


def py_normalize(p: bytes, ci: int, ip: int, ew: int) -> bytes:
    """transform_apply (transform_table.c:36-88) in pure Python, for small inputs."""
    out, in_space = bytearray(), False
    for c in p:
        if ew and c in _SPACE:
            if not in_space:
                out.append(32)
            in_space = True
            continue
        if ip and c in _PUNCT:
            continue
        out.append(c - 32 if ci and 97 <= c <= 122 else c)
        in_space = False
    if out and out[-1] == 32:
        out.pop()
    return bytes(out)


def case_patterns(case: dict) -> bytes:
    """Pattern file of a case.  Patterns that normalise to nothing are left out: the reference
    aborts the process on them (compiler.c:126-127)."""
    sf = case["store_flags"]
    pats = _pattern_set(case["patterns"])
    if any(sf):
        pats = [p for p in pats if py_normalize(p, *sf)]
    return b"\n".join(pats) + b"\n"


def case_haystack(case: dict) -> np.ndarray:
    kind, n, seed = case["hay"], case["size"], case["seed"]
    if kind == "text":
        vocab = [p for p in _pattern_set("names")[::37]]
        return text_haystack(n, seed, vocab)
    if kind == "synth":
        return synth_haystack(n, seed)
    if kind == "synth_planted":
        return plant(synth_haystack(n, seed), list(_pattern_set(case["patterns"])), seed ^ 0x77)
    if kind == "sentence":
        return sentence_haystack(n)
    if kind == "email":
        return np.frombuffer(golden_data("haystack_email.txt") * 64, dtype=np.uint8).copy()
    raise KeyError(kind)
