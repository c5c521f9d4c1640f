#!/usr/bin/env python3
"""Small driver for ncu: cfg5-style workload at a reduced size, a few device-resident matches.

  python tools/profile_scan.py --size-gib 2 --patterns 1000000 --iters 3
"""
import argparse, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import inputs, synth_torch
from omega_match_b200 import Compiler, Matcher

ap = argparse.ArgumentParser()
ap.add_argument("--size-gib", type=float, default=2.0)
ap.add_argument("--patterns", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--workload", default="cfg5")
ap.add_argument("--flags", default="")
a = ap.parse_args()

if a.workload == "cfg5":
    pats, sf, seed = inputs.synth_long_patterns(a.patterns), (0, 0, 0), inputs.SEED_H5
elif a.workload == "cfg4":
    pats, sf, seed = inputs.synth_short_patterns(), (0, 0, 0), inputs.SEED_H4
elif a.workload == "names":
    pats, sf, seed = [p for p in inputs.golden_data("names.txt").split(b"\n") if p], (0, 0, 0), inputs.SEED_H5
elif a.workload == "names-cpw":
    pats, sf, seed = [p for p in inputs.golden_data("names.txt").split(b"\n") if p], (1, 1, 1), inputs.SEED_H5
else:
    raise SystemExit("workload?")
olm = f"/tmp/prof_{a.workload}_{len(pats)}.olm"
Compiler.compile_from_buffer(olm, b"\n".join(pats) + b"\n", *map(bool, sf))
n = int(a.size_gib * (1 << 30))
hay = synth_torch.synth_haystack_torch(n, seed, device="cuda")
pb, pl = synth_torch.pack_patterns(pats, "cuda")
synth_torch.plant_torch(hay, pb, pl, seed ^ 0x77)
torch.cuda.synchronize()
kw = {k: True for k in a.flags.split(",") if k}
with Matcher(olm) as m:
    for i in range(a.iters):
        cnt, ptr = m.match_device(hay.data_ptr(), n, **kw)
        t = m.last_timing()
        print(f"iter {i}: {cnt} matches scan {t['scan_ms']:.3f} ms -> {n / t['scan_ms'] / 1e6:.1f} GB/s "
              f"total {t['total_ms']:.3f} ms launches {t['kernel_launches']}", flush=True)
