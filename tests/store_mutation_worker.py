"""Mutation fuzz of the store loader (store.cpp parse_store / stage_store / stage_stats / the
stride-2 table): valid `.olm` files with random bytes overwritten, header fields set to extreme
values, or the file truncated must be REJECTED OR LOADED -- never crash, never exhaust memory
(`omega_list_matcher_create` returns NULL on a bad file, matcher.c:492-495; a header that asks
for absurd allocations is a bad file).  Run by tests/test_host_logic.py in a subprocess so that a
crash shows up as a non-zero exit code.  No GPU needed: `olm_store_inspect` runs the same
parse + staging + self checks as `create()`.
"""
import ctypes as C
import os
import random
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))
import inputs
from omega_match_b200 import Compiler, _lib
from omega_match_b200._lib import StoreInfoC
seed=int(sys.argv[1]); n=int(sys.argv[2])
rng=random.Random(seed)
d=tempfile.mkdtemp(); base=os.path.join(d,"b.olm"); mut=os.path.join(d,"m.olm")
lists=[b"\n".join(inputs.synth_long_patterns(300)), inputs.golden_data("tlds.txt"), b"ab\nabc\nabcd\nabcde\nx\nhello world\nHELLO\n", b"\n".join(inputs.synth_short_patterns()[:500])]
L=_lib.load()
ok=bad=0
for i in range(n):
    try:
        Compiler.compile_from_buffer(base, rng.choice(lists), rng.random()<.5, rng.random()<.3, rng.random()<.3)
    except RuntimeError:
        continue
    b=bytearray(open(base,'rb').read())
    k=rng.choice([1,1,2,4,16])
    mode=rng.random()
    if mode<0.15: b=b[:rng.randrange(len(b))]
    else:
        for _ in range(k):
            pos=rng.randrange(len(b)) if rng.random()<0.5 else rng.randrange(min(len(b),72+64))
            if rng.random()<0.5: b[pos]=rng.randrange(256)
            else: b[pos:pos+4]=rng.choice([b"\xff\xff\xff\xff", b"\x00\x00\x00\x00", b"\xff\xff\xff\x7f", b"\x01\x00\x00\x00"])
    open(mut,'wb').write(bytes(b))
    rc=L.olm_store_inspect(os.fsencode(mut), C.byref(StoreInfoC()))
    if rc==0: ok+=1
    else: bad+=1
print("done ok",ok,"rejected",bad)
