#!/usr/bin/env python3
"""Development helper (GPU box): product vs oracle on a handful of cases, prints first diffs."""
import os, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import inputs
from omega_match_b200 import Compiler, Matcher
from oracle.oracle import Oracle

def run(name, pats, sf, hay, flagsets):
    bad = 0
    with tempfile.TemporaryDirectory() as d:
        olm = os.path.join(d, "t.olm")
        Compiler.compile_from_buffer(olm, pats, *map(bool, sf))
        o = Oracle.from_olm(olm)
        with Matcher(olm) as m:
            for mf in flagsets:
                kw = {k: True for k in mf}
                t0 = time.time(); got = m.match_arrays(hay, **kw); t1 = time.time()
                want = o.match(hay, **kw)
                ok = got.size == want.size and (got["offset"] == want["offset"]).all() and (got["len"] == want["len"]).all()
                tm = m.last_timing()
                print(f"{name:28s} sf={sf} {','.join(mf) or '-':50s} n={want.size:9d} got={got.size:9d} "
                      f"{'OK ' if ok else 'BAD'} scan={tm['scan_ms']:.3f}ms filt={tm['filter_ms']:.3f}ms wall={t1-t0:.3f}s", flush=True)
                if not ok:
                    bad += 1
                    a = set(zip(got["offset"].tolist(), got["len"].tolist())); b = set(zip(want["offset"].tolist(), want["len"].tolist()))
                    print("   only product:", sorted(a - b)[:8]); print("   only oracle :", sorted(b - a)[:8])
                    if a == b: 
                        idx = np.nonzero((got["offset"] != want["offset"]) | (got["len"] != want["len"]))[0][:5]
                        print("   same set, order differs at", idx, got[idx], want[idx])
    return bad

def main():
    bad = 0
    names = inputs.case_patterns(dict(patterns="names", store_flags=(0,0,0)))
    fs = [(), ("word_boundary",), ("longest_only","no_overlap"), ("line_end","longest_only","no_overlap"), ("word_prefix",), ("word_suffix",), ("no_overlap",)]
    small = inputs.text_haystack(100000, 7)
    bad += run("names-small-text", names, (0,0,0), small, fs)
    bad += run("names-kjv", names, (0,0,0), np.frombuffer(inputs.pseudo_kjv(), dtype=np.uint8), [(), ("longest_only","no_overlap")])
    for c in inputs.vector_cases():
        if c["name"] in ("names-text-c", "names-text-cpw", "synthshort-synth-plain", "synthlong-synth-plain", "names-sentence-plain", "synthshort-synth-cp"):
            bad += run(c["name"], inputs.case_patterns(c), c["store_flags"], inputs.case_haystack(c), fs)
    print("TOTAL BAD", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
