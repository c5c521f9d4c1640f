#!/usr/bin/env python3
"""bench.py -- haystack GB/s of the matching hot path (BASELINE.json metric).

Workload (default `cfg5`, BASELINE.json configs[4]): a 16 GiB synthetic haystack (SURVEY 8d
generator, one planted pattern per 4 KiB) x 1,000,000 compiled patterns (length 6-24 over
a-zA-Z), byte-range sharded over N GPUs of one node.  Weak scaling: every GPU owns a 16 GiB
byte range of an N x 16 GiB haystack (the path shards by byte range with no exchange on the
data path; the only collective is the gather of the per-rank sorted records).

A step = one pass of the hot path over the whole haystack:
  value  -- inputs already resident in HBM: every rank scans the start positions it owns
            (olm_cuda_match_shard), the per-rank sorted records are gathered to rank 0 over NCCL.
            K steps bracketed by barrier + cuda synchronize, max over ranks.
  e2e    -- the same through the host-pointer call (N=1: omega_list_matcher_match, the
            reference's own entry point): haystack in pinned HOST memory, H2D copy, kernels,
            D2H of the result records, every step.
  roofline -- algorithmic bytes (1 byte per haystack byte) / CUDA-event duration of
            the scan kernel, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline -- the unmodified reference library (oracle/_ref) with all host threads on a
            bounded prefix of the same haystack (the reference needs ~table_size probes per
            Bloom false positive, SURVEY F8, so the full size is out of reach).

  other_workloads -- (N=1, cfg5 runs only; skipped by `--no-experimental`) the default path on names.txt
            and on names.txt compiled with all three transform flags (4 MiB normalisation windows),
            4 GiB synthetic text each, device-timed, in processes of their own.

`--impl reference` times only that CPU reference on the same workload definition.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import inputs  # noqa: E402

GIB = 1 << 30
METRIC = "haystack_throughput"
UNIT = "GB/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there)
# must not get in between: fd 1 is pointed at stderr for the whole run and the line is
# written to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_traffic():
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return None
    return None


def workload_patterns(args):
    if args.workload == "cfg5":
        return inputs.synth_long_patterns(args.patterns), (0, 0, 0), inputs.SEED_H5, {}
    if args.workload == "cfg4":
        return inputs.synth_short_patterns(), (0, 0, 0), inputs.SEED_H4, {}
    if args.workload == "names":
        pats = [p for p in inputs.golden_data("names.txt").split(b"\n") if p]
        return pats, (0, 0, 0), inputs.SEED_H5, {}
    raise SystemExit(f"unknown workload {args.workload}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


class DevArray:
    """Zero-copy torch view of library-owned device memory (24-byte records as [n,3] int64)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count, 3), "typestr": "<i8", "data": (ptr, False), "version": 2}


def cpu_reference_setup(pattern_buf: bytes, store_flags):
    from oracle.oracle import Oracle, RefLib, ref_available
    tmp = tempfile.NamedTemporaryFile(suffix=".olm", delete=False)
    tmp.close()
    if ref_available():
        RefLib.compile(tmp.name, pattern_buf, *store_flags)
        ref = RefLib(tmp.name)
        return "reference", ref, ref.threads(), tmp.name
    o = Oracle.from_patterns(pattern_buf, *store_flags)
    return "port", o, 1, tmp.name


def cpu_match_timed(kind, obj, sample: np.ndarray, n: int, mflags):
    if kind == "reference":
        cnt, dt = obj.match_timed(sample, n, **mflags)
        return cnt, dt
    t0 = time.perf_counter()
    m = obj.match(sample[:n], **mflags)
    return m.size, time.perf_counter() - t0


def calibrate_sample(kind, obj, make_sample, mflags, target_s: float, max_bytes: int):
    """Grow the prefix until one match call takes about target_s seconds."""
    n = 1 << 20
    while True:
        s = make_sample(n)
        _, dt = cpu_match_timed(kind, obj, s, n, mflags)
        if dt >= target_s / 4 or n >= max_bytes:
            break
        n = min(max_bytes, int(n * min(8.0, max(2.0, target_s / max(dt, 1e-3) / 2))))
        n &= ~4095
    return n, s


def run_reference(args):
    """--impl reference: the reference's CPU implementation on a bounded prefix per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pats, sflags, seed_h, mflags = workload_patterns(args)
    pbuf = b"\n".join(pats) + b"\n"
    kind, obj, cores, tmpname = cpu_reference_setup(pbuf, sflags)

    def make_sample(n):
        h = inputs.plant(inputs.synth_haystack(n, seed_h), pats, seed_h ^ 0x77)
        buf = np.zeros(n + 64, dtype=np.uint8)
        buf[:n] = h
        return buf

    budget = 150.0 / max(1, args.steps + args.warmup)
    n, sample = calibrate_sample(kind, obj, make_sample, mflags, min(10.0, budget), 256 << 20)
    for _ in range(args.warmup):
        cpu_match_timed(kind, obj, sample, n, mflags)
    t = 0.0
    cnt = 0
    for _ in range(args.steps):
        cnt, dt = cpu_match_timed(kind, obj, sample, n, mflags)
        t += dt
    val = n * args.steps / t / 1e9
    os.unlink(tmpname)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, n_bytes=n, note="bounded prefix of the workload per step"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"first {n} bytes of the haystack, {cnt} matches"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(args, n_bytes=None, note=None):
    c = {"workload": {"cfg5": "BASELINE configs[4]: synthetic haystack x 1M compiled patterns, byte-range sharded",
                      "cfg4": "BASELINE configs[3]: short-matcher-heavy synthetic",
                      "names": "names.txt x synthetic haystack"}[args.workload],
         "haystack_bytes": int(n_bytes if n_bytes is not None else args.size_gib * GIB * args.gpus),
         "haystack_bytes_per_gpu": int(n_bytes if n_bytes is not None else args.size_gib * GIB),
         "patterns": args.patterns if args.workload == "cfg5" else None,
         "match_flags": [], "l2": "haystack is far larger than the 126 MB L2, no flush needed",
         "parallelism": f"byte-range shards x{args.gpus}"}
    if note:
        c["note"] = note
    return c


def run_ours(args):
    import torch
    import torch.distributed as dist
    import synth_torch
    from omega_match_b200 import Compiler, Matcher, _lib
    from omega_match_b200.sharding import gather_records, shard_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    pats, sflags, seed_h, mflags = workload_patterns(args)
    pbuf = b"\n".join(pats) + b"\n"
    olm = f"/tmp/olm_bench_{os.getuid()}_{args.workload}_{len(pats)}.olm"
    t0 = time.time()
    if local == 0:
        st = Compiler.compile_from_buffer(olm + ".tmp", pbuf, *map(bool, sflags))
        os.replace(olm + ".tmp", olm)
        log(f"[bench] compiled {len(pats)} patterns in {time.time() - t0:.1f}s: {st}")
    barrier()
    m = Matcher(olm, device=local)
    largest = max(len(p) for p in pats)

    total = int(args.size_gib * GIB) * world  # weak scaling: size_gib per GPU
    plan = shard_plan(total, world, largest, windowed=any(sflags))
    sh = plan[rank]
    # generate this rank's slice on its GPU; planting works on whole 4 KiB blocks, so generate block aligned
    gen_b = (sh.slice_begin // 4096) * 4096
    gen_e = min(total, ((sh.slice_end + 4095) // 4096) * 4096)
    t0 = time.time()
    gen = synth_torch.synth_haystack_torch(gen_e - gen_b, seed_h, start=gen_b, device=dev)
    pb, pl = synth_torch.pack_patterns(pats, dev)
    planted = synth_torch.plant_torch(gen, pb, pl, seed_h ^ 0x77, start=gen_b)
    del pb, pl
    # the slice handed to the library must start 16-byte aligned
    hay = torch.empty(((sh.slice_end - sh.slice_begin + 15) // 16) * 16 + 256, dtype=torch.uint8, device=dev)
    hay[:sh.slice_end - sh.slice_begin] = gen[sh.slice_begin - gen_b:sh.slice_end - gen_b]
    del gen
    torch.cuda.synchronize()
    log(f"[bench] rank {rank}: slice [{sh.slice_begin},{sh.slice_end}) own [{sh.own_begin},{sh.own_end}) "
        f"generated in {time.time() - t0:.1f}s, {planted} planted")

    def step():
        cnt, ptr = m.match_shard(hay.data_ptr(), sh.slice_begin, sh.slice_end - sh.slice_begin, sh.own_begin,
                                 sh.own_end, total, 0, **mflags)
        t = m.last_timing()
        merged = None
        if world > 1:
            loc = torch.as_tensor(DevArray(ptr, cnt), device=dev) if cnt else torch.empty((0, 3), dtype=torch.int64, device=dev)
            merged = gather_records(loc, dist, 0)
        return cnt, t, merged

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scan_ms, launches, cnt = 0.0, 0, 0
    for _ in range(args.steps):
        cnt, t, merged = step()
        scan_ms += t["scan_ms"]
        launches += int(t["kernel_launches"])
    torch.cuda.synchronize()
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    stats = torch.tensor([dt, scan_ms / args.steps, float(cnt), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dt, scan_ms_step = float(mx[0]), float(mx[1])
        total_matches, launches = int(sm[2]), int(sm[3])
    else:
        scan_ms_step, total_matches = scan_ms / args.steps, cnt
    value = total * args.steps / dt / 1e9

    # ---- e2e: host buffers, H2D + kernels + D2H inside the timed region
    own_len = sh.slice_end - sh.slice_begin
    host = torch.empty(own_len + 64, dtype=torch.uint8, pin_memory=True)
    host[:own_len].copy_(hay[:own_len])
    torch.cuda.synchronize()
    lib = _lib.load()
    e2e_steps = max(1, min(args.steps, 3))
    d2h = 0

    rec_host = [None]

    def e2e_step():
        nonlocal d2h
        if world == 1:
            res = lib.omega_list_matcher_match(m._matcher, host.data_ptr(), total, 0, 0, 0, 0, 0, 0, 0)
            if not res:
                raise SystemExit("omega_list_matcher_match failed")
            d2h = int(res.contents.count) * 24
            lib.omega_match_results_destroy(res)
        else:
            # this rank's slice from pinned host memory: segmented H2D overlapped with the scan
            c, ptr = m.match_shard_host(host.data_ptr(), sh.slice_begin, own_len, sh.own_begin, sh.own_end, total, 0,
                                        **mflags)
            if c:
                rec = torch.as_tensor(DevArray(ptr, c), device=dev)
                if rec_host[0] is None or rec_host[0].numel() < rec.numel():  # pinned, recycled (as the host API does)
                    rec_host[0] = torch.empty(rec.numel() + rec.numel() // 8, dtype=rec.dtype, pin_memory=True)
                rec_host[0][:rec.numel()].view_as(rec).copy_(rec, non_blocking=True)
                d2h = rec.numel() * 8
            torch.cuda.synchronize()

    e2e_step()
    if rank == 0:
        log(f"[bench] e2e call breakdown (device events, ms): {m.last_timing()}")
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    barrier()
    e2e_dt = time.perf_counter() - t0
    e2e_stats = torch.tensor([e2e_dt, float(own_len), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = e2e_stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = e2e_stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_dt, h2d_b, d2h_b = float(mx[0]), int(sm[1]), int(sm[2])
    else:
        h2d_b, d2h_b = own_len, d2h
    e2e_value = total * e2e_steps / e2e_dt / 1e9
    del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    own_bytes = sh.own_end - sh.own_begin
    achieved = own_bytes / (scan_ms_step * 1e-3) / 1e9 if scan_ms_step > 0 else 0.0
    traffic = kernel_traffic()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args),
            "matches_per_step": total_matches,
            "roofline": {"bound": "hbm", "kernel": "scan_kernel (+ prefix/place/redo, 4 launches)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": own_bytes, "kernel_ms": scan_ms_step,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
                    "steps": e2e_steps, "api": "omega_list_matcher_match(host ptr)" if world == 1
                    else "pinned host slice -> olm_cuda_match_shard_host -> host records, per rank"},
            "gpu_launches": launches, "clocks": clocks}

    # ---- CPU baseline: the reference on a bounded prefix, same bytes (N=1 only)
    if world == 1 and not args.no_cpu:
        try:
            kind, obj, cores, tmpname = cpu_reference_setup(pbuf, sflags)

            def make_sample(n):
                buf = np.zeros(n + 64, dtype=np.uint8)
                buf[:n] = hay[:n].cpu().numpy()
                return buf

            n, sample = calibrate_sample(kind, obj, make_sample, mflags, 10.0, 256 << 20)
            cnt_cpu, t_cpu = cpu_match_timed(kind, obj, sample, n, mflags)
            # parity of the very same prefix through the CUDA path
            from oracle.oracle import Oracle
            got = m.match_arrays(sample[:n], **mflags)
            ref_m = obj.match(sample[:n], **mflags) if kind == "port" else obj.match(sample[:n].tobytes(), **mflags)
            same = got.size == ref_m.size and bool((got["offset"] == ref_m["offset"]).all()) and bool(
                (got["len"] == ref_m["len"]).all())
            line["cpu_baseline"] = {"value": n / t_cpu / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"first {n} bytes of the same haystack, {cnt_cpu} matches, "
                                              f"{t_cpu:.2f}s; CUDA result on the same prefix identical: {same}"}
            os.unlink(tmpname)
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
    if world == 1 and args.workload == "cfg5" and not args.no_experimental:
        del hay
        torch.cuda.empty_cache()
        # the default path on the store shapes of BASELINE configs[0..2] (names.txt: 29 k patterns with
        # 1..4 byte ones; the same compiled with ignore-case + ignore-punctuation + elide-whitespace, i.e.
        # through the 4 MiB normalisation windows), 4 GiB synthetic text each, device-timed
        line["other_workloads"] = {w: profile_scan_leg(["--size-gib", "4", "--workload", w, "--iters", "3"], {})
                                   for w in ("names", "names-cpw")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_EXTRA_DEADLINE = [None]  # all extra legs together get 200 s; a leg that would not fit is skipped


def _leg_timeout():
    if _EXTRA_DEADLINE[0] is None:
        _EXTRA_DEADLINE[0] = time.time() + 200.0
    return min(75.0, _EXTRA_DEADLINE[0] - time.time())


def profile_scan_leg(argv, env):
    """One run of tools/profile_scan.py in a process of its own -> {"achieved": best GB/s after the
    first call, "matches_per_step": n} or {"error": ...}.  Reported, never required."""
    try:
        tmo = _leg_timeout()
        if tmo < 15.0:
            return {"error": "skipped: the extra legs' time budget is used up"}
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "profile_scan.py"), *argv], env=dict(os.environ, **env),
                           capture_output=True, text=True, timeout=tmo)
        iters = [ln for ln in r.stdout.splitlines() if ln.startswith("iter ")]
        if r.returncode != 0 or not iters:
            return {"error": (r.stderr or r.stdout)[-300:]}
        best = max(float(ln.split("->")[1].split("GB/s")[0]) for ln in (iters[1:] or iters))
        return {"achieved": best, "unit": UNIT, "matches_per_step": int(iters[-1].split(":")[1].split("matches")[0])}
    except Exception as e:
        return {"error": repr(e)[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg5", "cfg4", "names"])
    ap.add_argument("--size-gib", type=float, default=16.0)
    ap.add_argument("--patterns", type=int, default=1_000_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-experimental", action="store_true", help="skip the other_workloads legs (N=1, cfg5)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: fewer than 3 warm-up steps requested")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
