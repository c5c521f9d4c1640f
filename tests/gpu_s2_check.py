#!/usr/bin/env python3
"""GPU-box checker for the EXPERIMENTAL stride-2 sampled mode (OLM_SAMPLE2=1; DESIGN.md 7b,
device_tables.h S2Store, scan.cu scan_chunk_s2).  Not collected by pytest: the mode is off by
default; `--quick` ran bit-exact on a B200 (profiles/r1_s2_quick_parity.log), the full set and the
throughput run are for the next GPU session.

  OLM_SAMPLE2=1 python tests/gpu_s2_check.py            # product (sampled mode) vs oracle
  OLM_SAMPLE2=1 python tools/profile_scan.py --size-gib 4 --workload cfg5 --iters 3   # throughput

Every case is a store that qualifies for the mode (all patterns >= 6 bytes, class prefilter with
run >= 5); flag sets without a position predicate go through scan_chunk_s2, the others through the
regular tables (they must keep working with the second table loaded).
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("OLM_SAMPLE2", "1")

import numpy as np  # noqa: E402

import inputs  # noqa: E402
from gpu_quick import run  # noqa: E402  (imports the oracle: this file lives under tests/)

FS = [(), ("longest_only",), ("no_overlap",), ("longest_only", "no_overlap"), ("word_boundary",), ("line_end", "longest_only")]


def main() -> int:
    bad = 0
    pats = inputs.synth_long_patterns(3000)
    if "--quick" in sys.argv:  # a few seconds: edges, coinciding keys, overflow -> redo, one windowed store
        for n in (7, 513, 4097, 100_003, (1 << 20) + 1):
            hay = inputs.plant(inputs.synth_haystack(n, inputs.SEED_H5 + n), pats, 0x51 + n, block=256)
            bad += run(f"synth3000-n{n}", b"\n".join(pats), (0, 0, 0), hay, FS[:4])
        adv = [b"aaaaaaa", b"aaaaaaaa", b"aaaaaab", b"baaaaaa", b"abababab", b"bababababa", b"abcdefgh", b"bcdefghi", b"xabcdefgh"]
        text = (b"xabcdefghijklmnopqrstuvwxyza aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaab " * 200)[:9001]
        bad += run("adversarial-alpha", b"\n".join(adv), (0, 0, 0), np.frombuffer(text, dtype=np.uint8).copy(), FS[:4])
        bad += run("dense-a", b"\n".join(adv), (0, 0, 0), np.full(50_000, ord("a"), dtype=np.uint8), FS[:2])
        bad += run("synth3000-ci", b"\n".join(pats), (1, 0, 0), inputs.plant(inputs.synth_haystack(5_000_001, 99), pats, 5, block=512), FS[:2])
        print("TOTAL BAD", bad)
        return 1 if bad else 0
    # planted patterns at odd and even offsets, chunk / tile / end-of-buffer edges
    for n in (5, 6, 7, 511, 512, 513, 4095, 4096, 4097, 8193, 100_003, 1 << 20, (1 << 22) + 77):
        hay = inputs.plant(inputs.synth_haystack(n, inputs.SEED_H5 + n), pats, 0x51 + n, block=256)
        if n >= 64:  # a pattern flush with the end, one starting at 0 and at 1
            hay[n - len(pats[0]):] = np.frombuffer(pats[0], dtype=np.uint8)
            hay[:len(pats[1])] = np.frombuffer(pats[1], dtype=np.uint8)
            hay[1 + len(pats[1]) + 1:1 + len(pats[1]) + 1 + len(pats[2])] = np.frombuffer(pats[2], dtype=np.uint8)
        bad += run(f"synth3000-n{n}", b"\n".join(pats), (0, 0, 0), hay, FS[:4] if n < 4096 else FS)
    # keys that coincide for both shifts, patterns that are shifts / prefixes / suffixes of each other
    adv = [b"aaaaaaa", b"aaaaaaaa", b"aaaaaab", b"baaaaaa", b"abababab", b"bababababa", b"abcdefgh", b"bcdefghi",
           b"xabcdefgh", b"abcdefghijklmnopqrstuvwxyz", b"bcdefghijklmnopqrstuvwxyza", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa"]
    rng = np.random.default_rng(7)
    for n in (33, 1000, 70_001):
        hay = rng.choice(np.frombuffer(b"ab", dtype=np.uint8), size=n).astype(np.uint8)
        hay2 = np.frombuffer((b"xabcdefghijklmnopqrstuvwxyza aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaab " * (n // 60 + 1))[:n], dtype=np.uint8).copy()
        bad += run(f"adversarial-ab-n{n}", b"\n".join(adv), (0, 0, 0), hay, FS[:4])
        bad += run(f"adversarial-alpha-n{n}", b"\n".join(adv), (0, 0, 0), hay2, FS)
    # dense: every position of a run of 'a' starts several matches (staging overflow -> redo_kernel)
    bad += run("dense-a", b"\n".join(adv), (0, 0, 0), np.full(300_000, ord("a"), dtype=np.uint8), FS[:4])
    # a transforming store (4 MiB windows): case-folded long patterns
    bad += run("synth3000-ci", b"\n".join(pats), (1, 0, 0), inputs.plant(inputs.synth_haystack(9_000_001, 99), pats, 5, block=512), FS[:4])
    bad += run("synth3000-cpw", b"\n".join(pats), (1, 1, 1), inputs.plant(inputs.synth_haystack(5_000_001, 98), pats, 6, block=512), FS[:4])
    # long patterns only: K = 8
    long9 = [p for p in inputs.synth_long_patterns(4000) if len(p) >= 9]
    bad += run("synth-min9", b"\n".join(long9), (0, 0, 0), inputs.plant(inputs.synth_haystack(2_000_003, 97), long9, 7, block=300), FS[:4])
    print("TOTAL BAD", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
