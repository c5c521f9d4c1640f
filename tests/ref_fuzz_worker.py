"""Differential fuzz of the oracle against the compiled reference (oracle/_ref).

Run by tests/test_oracle_golden.py in a subprocess with MALLOC_PERTURB_=255 so that the
reference's reads of never-written scratch bytes (see tests/golden/make_golden.py) are
deterministic.  Exit code 0 = all trials agree.
"""
import random
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle.oracle import Oracle, RefLib  # noqa: E402

FLAGS = ("no_overlap", "longest_only", "word_boundary", "word_prefix", "word_suffix", "line_start", "line_end")


def main(seed: int, trials: int) -> int:
    rng = random.Random(seed)
    alph = b"abcABC xyz.,-'\n\r\t_09"
    tmp = Path(tempfile.mkdtemp()) / "f.olm"
    done = 0
    for _ in range(trials):
        pats = set()
        while len(pats) < rng.choice([1, 3, 10, 60]):
            pats.add(bytes(rng.choice(b"abcABC xyz.-'_09") for _ in range(rng.choice([1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 9, 12]))))
        sf = (rng.random() < 0.5, rng.random() < 0.4, rng.random() < 0.4)
        buf = b"\n".join(sorted(pats))
        try:
            o = Oracle.from_patterns(buf, *sf)
        except ValueError:
            continue  # a pattern normalises to nothing: the reference would abort()
        pst = RefLib.compile(tmp, buf, *sf)
        assert Oracle.from_olm(tmp).info() == o.info()
        assert pst["stored_pattern_count"] == o.info()["long"]
        ref = RefLib(tmp)
        for _ in range(5):
            hay = bytes(rng.choice(alph) for _ in range(rng.choice([0, 1, 3, 4, 5, 17, 300, 4000, 9000])))
            kw = {f: rng.random() < 0.3 for f in FLAGS}
            a, b = o.match(hay, **kw), ref.match(hay, **kw)
            if not (a.size == b.size and (a == b).all()):
                print("MISMATCH", sorted(pats), sf, kw, hay, a, b)
                return 1
            if o.stats.as_dict() != ref.stats.as_dict():
                print("STATS MISMATCH", o.stats.as_dict(), ref.stats.as_dict())
                return 1
            done += 1
        ref.close()
    print(f"{done} comparisons agree")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]), int(sys.argv[2])))
