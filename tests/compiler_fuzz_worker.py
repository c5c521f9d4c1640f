"""Differential fuzz of the product's `.olm` writer (compiler.cpp, no GPU needed) against the
compiled reference (oracle/_ref): SURVEY 8f N1.

For random pattern lists -- duplicates, CR/LF line ends, empty lines, leading/trailing blanks,
bytes >= 0x80, very long lines -- and every store-flag combination:
  * both compilers report the same `omega_match_pattern_store_stats_t` and write files of the
    same size (`compiler.c:197-425`);
  * the oracle reads the same header facts from both files;
  * the REFERENCE matcher gives the same (offset, len) stream with either file, with and
    without match flags;
  * the device tables staged from the product's file (store.cpp: key buckets, slots, records,
    bitmaps, class prefilter) pass their self check -- every pattern is found again through the
    probe sequence the scan kernel uses (`olm_store_inspect`).
Run by tests/test_host_logic.py in a subprocess (MALLOC_PERTURB_ as in ref_fuzz_worker.py).
Exit code 0 = all trials agree.
"""
import random
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import ctypes as C  # noqa: E402
import os  # noqa: E402

from omega_match_b200 import Compiler, _lib  # noqa: E402
from omega_match_b200._lib import StoreInfoC  # noqa: E402
from oracle.oracle import Oracle, RefLib  # noqa: E402

FLAGS = ("no_overlap", "longest_only", "word_boundary", "word_prefix", "word_suffix", "line_start", "line_end")
ALPHABETS = (
    b"abcABC xyz.-'_09",
    b"ab",
    bytes(range(0x20, 0x7F)),
    bytes(range(1, 256)).replace(b"\n", b"").replace(b"\r", b""),
    b"aA  \t.,;:!?-_'\"()[]\x07\x08\x0b\x0c",
)


def random_list(rng):
    alph = rng.choice(ALPHABETS)
    n = rng.choice([1, 2, 5, 20, 100, 400])
    lens = rng.choice([[1, 2, 3, 4], [1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 9, 12], [5, 6, 7, 8, 9, 13, 33], [4, 5], [6, 7, 9, 14, 24], [9, 10, 30], [40, 200, 700]])
    pats = []
    for _ in range(n):
        p = bytes(rng.choice(alph) for _ in range(rng.choice(lens)))
        if rng.random() < 0.1:
            p = b" " + p
        if rng.random() < 0.1:
            p = p + rng.choice([b" ", b"  ", b"\t", b"."])
        pats.append(p)
        if rng.random() < 0.15:
            pats.append(rng.choice(pats))                        # exact duplicate
        if rng.random() < 0.1:
            pats.append(rng.choice(pats).swapcase())             # duplicate after case folding
    sep = rng.choice([b"\n", b"\r\n", b"\n", b"\n\n"])
    buf = sep.join(pats)
    if rng.random() < 0.5:
        buf += rng.choice([b"\n", b"\r\n", b"\n\n"])
    return buf, alph


def main(seed: int, trials: int) -> int:
    rng = random.Random(seed)
    d = Path(tempfile.mkdtemp())
    mine, theirs = d / "mine.olm", d / "ref.olm"
    done = skipped = 0
    for _ in range(trials):
        buf, alph = random_list(rng)
        sf = (rng.random() < 0.5, rng.random() < 0.4, rng.random() < 0.4)
        try:
            Oracle.from_patterns(buf, *sf)
        except ValueError:
            # a pattern normalises to nothing: the reference abort()s; the product must refuse too
            try:
                Compiler.compile_from_buffer(str(mine), buf, *sf)
            except Exception:
                skipped += 1
                continue
            print("PRODUCT ACCEPTED what the reference aborts on", buf, sf)
            return 1
        st = Compiler.compile_from_buffer(str(mine), buf, *sf)
        rs = RefLib.compile(theirs, buf, *sf)
        if st.__dict__ != rs:
            print("STATS MISMATCH", buf, sf, st.__dict__, rs)
            return 1
        if mine.stat().st_size != theirs.stat().st_size:
            print("SIZE MISMATCH", buf, sf, mine.stat().st_size, theirs.stat().st_size)
            return 1
        if Oracle.from_olm(mine).info() != Oracle.from_olm(theirs).info():
            print("INFO MISMATCH", buf, sf, Oracle.from_olm(mine).info(), Oracle.from_olm(theirs).info())
            return 1
        for f in (mine, theirs):  # staging + self check of the device tables, from either writer's file
            if _lib.load().olm_store_inspect(os.fsencode(f), C.byref(StoreInfoC())) != 0:
                print("STAGING SELF CHECK FAILED", f.name, buf, sf)
                return 1
        a, b = RefLib(mine), RefLib(theirs)
        for _ in range(3):
            hay = bytes(rng.choice(alph + b" \n") for _ in range(rng.choice([0, 3, 40, 1500, 6000])))
            if rng.random() < 0.5 and len(buf) < 4000:
                hay += buf  # every pattern occurs at least once
            kw = {f: rng.random() < 0.25 for f in FLAGS}
            x, y = a.match(hay, **kw), b.match(hay, **kw)
            if not (x.size == y.size and (x == y).all()):
                print("MATCH MISMATCH", buf, sf, kw, hay, x, y)
                return 1
            done += 1
        a.close()
        b.close()
    print(f"{done} comparisons agree, {skipped} lists refused by both")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]), int(sys.argv[2])))
