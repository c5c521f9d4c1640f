#!/usr/bin/env python3
"""cuobjdump -sass of an object / library -> SASS instruction count, local loads/stores and calls per kernel."""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
name, rows = None, []
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        rows.append([name, 0, 0, 0])
        continue
    if name and re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", ln):
        rows[-1][1] += 1
        if re.search(r"\b(STL|LDL)\b", ln): rows[-1][2] += 1
        if "CALL" in ln: rows[-1][3] += 1
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
for r, n in zip(rows, names):
    n = re.sub(r"olm::\(anonymous namespace\)::", "", n); n = re.sub(r"\(olm::ScanParams.*", "", n)
    if len(sys.argv) < 3 or sys.argv[2] in n:
        print(f"{r[1]:6d} instr  {r[2]:3d} local  {r[3]:2d} calls  {n}")
