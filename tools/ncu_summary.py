#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): key counters + the hottest source lines.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep [--lines 25]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            print(f"  {k:70s} {d[k]:>18s} {units[hdr.index(k)]}")
    st = [(k, float(d[k])) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and d[k] not in ("", "-nan", "nan")]
    for k, v in sorted(st, key=lambda x: -x[1])[:8]:
        print(f"  stall {k.split('stalled_')[1].split('_per_issue')[0]:28s} {v:8.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
if hi is not None:
    h = rows[hi]
    cols = {n: i for i, n in enumerate(h)}
    i_src, samp, inst, thr = cols["Source"], cols["# Samples"], cols["Instructions Executed"], cols["Avg. Threads Executed"]
    body = [r for r in rows[hi + 1:] if len(r) > samp and r[0].startswith("0x")]
    tot = sum(float(r[samp] or 0) for r in body) or 1
    tin = sum(float(r[inst] or 0) for r in body) or 1
    print(f"-- SASS: {len(body)} instructions, {tin:.3g} warp-instructions executed, {tot:.0f} samples")
    print("-- hottest SASS by stall samples")
    for r in sorted(body, key=lambda r: -float(r[samp] or 0))[:nlines]:
        print(f"  {100*float(r[samp] or 0)/tot:5.1f}%  exec={float(r[inst] or 0)/tin*100:5.2f}%  thr={r[thr]:>5s}  {r[0][-5:]}  {r[i_src].strip()[:110]}")
    print("-- most executed SASS")
    for r in sorted(body, key=lambda r: -float(r[inst] or 0))[:nlines]:
        print(f"  exec={float(r[inst] or 0)/tin*100:5.2f}%  samples={100*float(r[samp] or 0)/tot:5.1f}%  thr={r[thr]:>5s}  {r[0][-5:]}  {r[i_src].strip()[:110]}")
