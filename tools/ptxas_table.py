#!/usr/bin/env python3
"""nvcc -Xptxas -v output (stdin or file) -> one line per kernel: registers, spills, demangled name."""
import re, subprocess, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
cur = None
rows = []
for ln in txt.splitlines():
    m = re.search(r"Compiling entry function '([^']+)'", ln)
    if m:
        cur = {"name": m.group(1), "spill": "", "regs": ""}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m and not cur["spill"]:
        cur["spill"] = f"stack {m.group(1):>4} st {m.group(2):>4} ld {m.group(3):>4}"
    m = re.search(r"Used (\d+) registers", ln)
    if m:
        cur["regs"] = m.group(1)
names = subprocess.run(["c++filt"] + [r["name"] for r in rows], capture_output=True, text=True).stdout.splitlines()
for r, n in zip(rows, names):
    n = re.sub(r"olm::\(anonymous namespace\)::", "", n)
    n = re.sub(r"\(olm::ScanParams.*", "", n)
    print(f"{r['regs']:>3} regs  {r['spill']}  {n}")
