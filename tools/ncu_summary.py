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
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h = rows[0]
    try:
        i_src = h.index("Source"); 
    except ValueError:
        i_src = 1
    cols = {n: i for i, n in enumerate(h)}
    samp = cols.get("# Samples") or cols.get("Warp Stall Sampling (All Samples)") or cols.get("Sampling Data (All)")
    inst = cols.get("Instructions Executed")
    print("columns:", [c for c in h][:40])
    if samp is not None:
        agg = []
        for r in rows[1:]:
            try:
                agg.append((float(r[samp] or 0), r))
            except Exception:
                pass
        tot = sum(a for a, _ in agg) or 1
        print(f"-- hottest lines by samples (total {tot:.0f})")
        for a, r in sorted(agg, key=lambda x: -x[0])[:nlines]:
            print(f"  {100*a/tot:5.1f}%  inst={r[inst] if inst is not None else '?':>12s}  {r[i_src][:150]}")
