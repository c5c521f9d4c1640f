#!/usr/bin/env python3
"""Aggregate an .ncu-rep by CUDA source line: executed warp-instructions and stall samples.
   python tools/ncu_lines.py gpurun_out/x.ncu-rep [--top 40]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname, hdr, lines = "", None, []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
    elif hdr and r[0].strip().isdigit():
        try:
            lines.append((fname, int(r[0]), r[1].strip(), float(r[hdr["Instructions Executed"]] or 0),
                          float(r[hdr["# Samples"]] or 0), r[hdr["Avg. Threads Executed"]]))
        except (ValueError, IndexError):
            pass
ti = sum(l[3] for l in lines) or 1
ts = sum(l[4] for l in lines) or 1
print(f"total warp-instructions {ti:.4g}, samples {ts:.0f}")
print("-- by executed instructions")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"  exec={100*l[3]/ti:5.2f}% samp={100*l[4]/ts:5.2f}% thr={l[5]:>5s} {l[0]}:{l[1]:<4d} {l[2][:100]}")
print("-- by stall samples")
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"  samp={100*l[4]/ts:5.2f}% exec={100*l[3]/ti:5.2f}% thr={l[5]:>5s} {l[0]}:{l[1]:<4d} {l[2][:100]}")
if "--ranges" in sys.argv:
    import re
    spec = sys.argv[sys.argv.index("--ranges") + 1]  # "name:lo-hi,name:lo-hi"
    print("-- by line range (scan.cu)")
    for part in spec.split(","):
        name, rng = part.split(":")
        lo, hi = map(int, rng.split("-"))
        e = sum(l[3] for l in lines if l[0] == "scan.cu" and lo <= l[1] <= hi)
        sm = sum(l[4] for l in lines if l[0] == "scan.cu" and lo <= l[1] <= hi)
        print(f"  {name:16s} exec={100*e/ti:5.2f}% samp={100*sm/ts:5.2f}%")
    e = sum(l[3] for l in lines if l[0] != "scan.cu")
    sm = sum(l[4] for l in lines if l[0] != "scan.cu")
    print(f"  {'other files':16s} exec={100*e/ti:5.2f}% samp={100*sm/ts:5.2f}%")
