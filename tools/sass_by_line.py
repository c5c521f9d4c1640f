#!/usr/bin/env python3
"""Static code size by source region: nvdisasm -g -c of an object's cubin, SASS instructions of ONE kernel
counted per (file, line) and summed per region of scan.cu / scan_device.cuh.
   python tools/sass_by_line.py omega_match_b200/build/scan.o 'scan_kernelILb1ELb1ELb0ELb1ELb1E' [--lines]"""
import os, re, subprocess, sys, tempfile, collections
obj, pat = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
inside, cur, counts = False, ("?", 0), collections.Counter()
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        inside = pat in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        counts[cur] += 1
total = sum(counts.values())
print("total SASS instructions", total)
byfile = collections.Counter()
for (f, l), n in counts.items():
    byfile[f] += n
for f, n in byfile.most_common():
    print(f"  {n:6d}  {f}")
if "--lines" in sys.argv:
    for (f, l), n in sorted(counts.items(), key=lambda x: -x[1])[:60]:
        print(f"  {n:5d}  {f}:{l}")
# regions by function: find function starts in the two sources
def regions(path):
    out = []
    for i, ln in enumerate(open(path), 1):
        m = re.match(r"\s*(?:template.*>\s*)?(?:static\s+)?__(?:device|global)__.*?\b(\w+)\s*\(", ln)
        if m and "=" not in ln.split("(")[0]:
            out.append((i, m.group(1)))
    return out
root = os.path.dirname(os.path.abspath(__file__)) + "/../omega_match_b200/csrc/"
for fn in ("scan.cu", "scan_device.cuh"):
    regs = regions(root + fn)
    tot = collections.Counter()
    for (f, l), n in counts.items():
        if f != fn:
            continue
        name = "?"
        for s, nm in regs:
            if s <= l:
                name = nm
        tot[name] += n
    print("--", fn)
    for nm, n in tot.most_common():
        print(f"  {n:6d}  {nm}")
