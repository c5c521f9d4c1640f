"""CPU model of the experimental stride-2 sampled scan (scan.cu `scan_chunk_s2` / `verify_s2`,
device_tables.h `S2Store`) against the oracle.

The CUDA path of that mode is off by default and has only a quick parity run on a GPU behind it
(DESIGN.md 7b item 4); this test pins the ALGORITHM the kernel implements -- which positions are
probed, which entries a key lists and in which order, the ownership and end-of-buffer rules -- so
that what the GPU checks validate is the transcription, not the idea:

  * tiles of 4096 start positions, chunks of 512; only the ODD tile-relative positions p <= nscan
    are probed, and only when K bytes are left from p on;
  * a key (the K bytes at p, K = min(8, shortest pattern - 1)) lists, for every pattern, an entry
    with shift 0 (the pattern's bytes [0, K)) and one with shift 1 (bytes [1, K + 1)), ordered
    shift 1 first, then longest first;
  * an entry (shift, len) at p is a match starting at p - shift iff that start is one of the
    tile's positions (< nscan), len - shift bytes are left from p on, the bytes from p on equal the
    pattern's from `shift` on, and (shift 1) the byte before p equals the pattern's first byte;
    `longest_only` keeps the first match per start.
The stream that falls out -- with no sort -- must be the oracle's (offset ascending, length
descending), for every haystack length around the chunk and tile edges.
"""
import numpy as np
import pytest

import inputs
from oracle.oracle import Oracle

TILE, CHUNK = 4096, 512


def s2_model(patterns, hay: bytes, longest_only=False):
    smallest = min(map(len, patterns))
    assert smallest >= 6
    K = min(8, smallest - 1)
    cls_run = min(8, smallest)
    if cls_run == 7:
        cls_run = 6
    cls = set()
    for p in patterns:
        cls.update(p[:cls_run])
    run = cls_run - 1
    table = {}
    for pat in patterns:
        for sh in (0, 1):
            table.setdefault(pat[sh:sh + K], []).append((sh, len(pat), pat))
    for ents in table.values():
        ents.sort(key=lambda e: (-e[0], -e[1]))
    n = len(hay)
    out = []
    for p0 in range(0, n, TILE):
        nscan, rem0 = min(TILE, n - p0), n - p0
        for cbase in range(0, nscan, CHUNK):
            for p in range(cbase + 1, cbase + CHUNK, 2):
                if p > nscan or p + K > rem0:
                    continue
                g = p0 + p
                if any(b not in cls for b in hay[g:g + run]):
                    continue
                done = [False, False]
                for sh, ln, pat in table.get(hay[g:g + K], ()):
                    body = ln - sh
                    if p - sh >= nscan or body > rem0 - p:
                        continue
                    if longest_only and done[sh]:
                        continue
                    if hay[g:g + body] != pat[sh:] or (sh and hay[g - 1] != pat[0]):
                        continue
                    out.append((g - sh, ln))
                    done[sh] = True
    return out


ADVERSARIAL = [b"aaaaaaa", b"aaaaaaaa", b"aaaaaab", b"baaaaaa", b"abababab", b"bababababa", b"abcdefgh", b"bcdefghi",
               b"xabcdefgh", b"abcdefghijklmnopqrstuvwxyz", b"bcdefghijklmnopqrstuvwxyza", b"a" * 33]


def _check(patterns, hay: np.ndarray):
    o = Oracle.from_patterns(b"\n".join(patterns))
    hb = hay.tobytes()
    for longest in (False, True):
        want = o.match(hay, longest_only=longest)
        got = s2_model(patterns, hb, longest)
        assert len(got) == want.size, (len(hb), longest, len(got), want.size)
        assert [g[0] for g in got] == want["offset"].tolist() and [g[1] for g in got] == want["len"].tolist()


@pytest.mark.parametrize("n", [5, 6, 7, 8, 15, 16, 17, 511, 512, 513, 4095, 4096, 4097, 8191, 8193, 20_001])
def test_model_equals_oracle_at_every_edge(n):
    pats = inputs.synth_long_patterns(300)
    hay = inputs.plant(inputs.synth_haystack(n, inputs.SEED_H5 + n), pats, 0x51 + n, block=128)
    if n >= 64:  # a pattern flush with the end, one at 0, one at an odd offset right after it
        hay[n - len(pats[0]):] = np.frombuffer(pats[0], dtype=np.uint8)
        hay[:len(pats[1])] = np.frombuffer(pats[1], dtype=np.uint8)
        at = len(pats[1]) + (0 if len(pats[1]) % 2 else 1)
        hay[at:at + len(pats[2])] = np.frombuffer(pats[2], dtype=np.uint8)
    _check(pats, hay)
    # every pattern ends exactly at the end of the buffer once
    for pat in pats[:12]:
        for cut in (0, 1):
            if n - cut < len(pat):
                continue
            h = hay[: n - cut].copy()
            h[len(h) - len(pat):] = np.frombuffer(pat, dtype=np.uint8)
            _check(pats, h)


@pytest.mark.parametrize("n", [33, 1000, 9001])
def test_model_with_coinciding_keys_and_shifted_patterns(n):
    rng = np.random.default_rng(n)
    _check(ADVERSARIAL, rng.choice(np.frombuffer(b"ab", dtype=np.uint8), size=n).astype(np.uint8))
    text = (b"xabcdefghijklmnopqrstuvwxyza aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaab " * (n // 60 + 1))[:n]
    _check(ADVERSARIAL, np.frombuffer(text, dtype=np.uint8).copy())
    _check(ADVERSARIAL, np.full(n, ord("a"), dtype=np.uint8))


def test_model_long_keys():
    pats = [p for p in inputs.synth_long_patterns(600) if len(p) >= 9]  # K = 8
    _check(pats, inputs.plant(inputs.synth_haystack(30_011, 97), pats, 7, block=100))
