#!/usr/bin/env python3
"""Aggregate an .ncu-rep by CUDA source line: shared-memory wavefronts (ideal / excessive) and global L1 tag requests.
   python tools/ncu_wavefronts.py gpurun_out/x.ncu-rep [--top 30]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname, hdr, lines = "", None, []
def f(r, k):
    try:
        return float(r[hdr[k]] or 0)
    except (ValueError, IndexError, KeyError):
        return 0.0
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
    elif hdr and r[0].strip().isdigit():
        lines.append((fname, int(r[0]), r[1].strip(), f(r, "L1 Wavefronts Shared"), f(r, "L1 Wavefronts Shared Ideal"),
                      f(r, "L1 Tag Requests Global"), f(r, "Instructions Executed"), f(r, "L2 Theoretical Sectors Global")))
tw = sum(l[3] for l in lines) or 1
tg = sum(l[5] for l in lines) or 1
print(f"shared wavefronts {tw:.4g} (ideal {sum(l[4] for l in lines):.4g}), global L1 tag requests {tg:.4g}, "
      f"L2 theoretical sectors {sum(l[7] for l in lines):.4g}, warp-instructions {sum(l[6] for l in lines):.4g}")
print("-- by shared wavefronts")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"  wf={100*l[3]/tw:5.2f}% ({l[3]:.3g}, ideal {l[4]:.3g}) {l[0]}:{l[1]:<4d} {l[2][:90]}")
print("-- by global tag requests")
for l in sorted(lines, key=lambda l: -l[5])[:12]:
    print(f"  tag={100*l[5]/tg:5.2f}% ({l[5]:.3g}, sectors {l[7]:.3g}) {l[0]}:{l[1]:<4d} {l[2][:90]}")
