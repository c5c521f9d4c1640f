"""Multi-GPU host logic on CPU: shard plan, ownership rule, ordered gather (gloo, world_size 2).

The per-shard scan is done here by the ORACLE on the bytes a rank would hold (its slice), so
what is tested is exactly what the N>1 path adds on top of a single-GPU scan: which bytes a
rank needs, which starts it owns, and that concatenating per-rank results in rank order (plus
one global no-overlap pass) equals the unsharded result."""
import os
import socket

import numpy as np
import pytest

import inputs
from conftest import same_matches
from omega_match_b200.sharding import WINDOW, Shard, gather_records, shard_plan
from oracle.oracle import MATCH_DTYPE, Oracle


def test_shard_plan_covers_everything():
    for size in (0, 1, 4095, 4096, 100_000, 9 * WINDOW + 17):
        for world in (1, 2, 3, 8):
            for windowed in (False, True):
                plan = shard_plan(size, world, 24, windowed)
                assert [s.rank for s in plan] == list(range(world))
                assert plan[0].own_begin == 0 and plan[-1].own_end == size
                for a, b in zip(plan, plan[1:]):
                    assert a.own_end == b.own_begin
                for s in plan:
                    assert s.slice_begin <= s.own_begin <= s.own_end <= s.slice_end <= size
                    assert (s.own_begin - s.slice_begin) % 16 == 0
                    if windowed:
                        assert s.own_begin % WINDOW == 0 and (s.slice_begin, s.slice_end) == (s.own_begin, s.own_end)
                    elif s.own_len:
                        assert s.slice_end == min(size, s.own_end + 25)
                        assert s.own_begin == 0 or s.slice_begin <= s.own_begin - 1


def scan_shard_with_oracle(o: Oracle, hay: np.ndarray, s: Shard, windowed: bool, **flags) -> np.ndarray:
    """What a rank reports: matches of ITS bytes whose start it owns, with global offsets."""
    if s.own_len == 0:
        return np.zeros(0, dtype=MATCH_DTYPE)
    local = o.match(hay[s.slice_begin:s.slice_end], **flags)
    local = local.copy()
    local["offset"] += np.uint64(s.slice_begin)
    keep = (local["offset"] >= s.own_begin) & (local["offset"] < s.own_end)
    return local[keep]


def no_overlap_host(m: np.ndarray) -> np.ndarray:
    keep, last_end = [], -1
    for i in range(m.size):
        if int(m["offset"][i]) >= last_end:
            keep.append(i)
            last_end = int(m["offset"][i]) + int(m["len"][i])
    return m[keep]


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("flags", [{}, {"word_boundary": True}, {"longest_only": True},
                                   {"line_start": True, "word_suffix": True}, {"no_overlap": True},
                                   {"longest_only": True, "no_overlap": True}])
def test_sharded_equals_unsharded_plain(world, flags):
    pats = inputs.synth_long_patterns(3000) + [b"ab", b"the", b"o", b"King", b"zzzz", b"a b"]
    hay = inputs.plant(inputs.synth_haystack(300_000, 77), pats, 5, block=512)
    o = Oracle.from_patterns(pats)
    want = o.match(hay, **flags)
    shard_flags = {k: v for k, v in flags.items() if k != "no_overlap"}
    parts = [scan_shard_with_oracle(o, hay, s, False, **shard_flags) for s in shard_plan(hay.size, world, 24, False)]
    got = np.concatenate(parts)
    if flags.get("no_overlap"):
        got = no_overlap_host(got)
    assert same_matches(got, want)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded_windowed(world):
    pats = [p for p in inputs.golden_data("names.txt").split(b"\n") if p][::9]
    hay = inputs.text_haystack(3 * WINDOW + 5000, 11)
    o = Oracle.from_patterns(pats, 1, 1, 1)
    want = Oracle.from_patterns(pats, 1, 1, 1).match(hay, longest_only=True)
    parts = [scan_shard_with_oracle(Oracle.from_patterns(pats, 1, 1, 1), hay, s, True, longest_only=True)
             for s in shard_plan(hay.size, world, 24, True)]
    assert same_matches(np.concatenate(parts), want)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank: int, world: int, port: int, out_path: str):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        pats = inputs.synth_long_patterns(2000) + [b"ab", b"the", b"King"]
        hay = inputs.plant(inputs.synth_haystack(200_000, 99), pats, 6, block=1024)
        if world > 2:  # one rank without a single match: its gather leg is skipped on both sides
            hay[shard_plan(hay.size, world, 24, False)[1].slice_begin:shard_plan(hay.size, world, 24, False)[1].slice_end] = 0x23
        o = Oracle.from_patterns(pats)
        s = shard_plan(hay.size, world, 24, False)[rank]
        mine = scan_shard_with_oracle(o, hay, s, False)
        rec = np.zeros((mine.size, 3), dtype=np.int64)  # the 24-byte record layout the library uses
        rec[:, 0] = mine["offset"].astype(np.int64)
        rec[:, 1] = mine["len"].astype(np.int64)
        merged = gather_records(torch.from_numpy(rec), dist, dst=0)
        if rank == 0:
            want = o.match(hay)
            got = merged.numpy()
            ok = got.shape[0] == want.size and (got[:, 0] == want["offset"].astype(np.int64)).all() and (
                got[:, 1] == want["len"].astype(np.int64)).all()
            with open(out_path, "w") as f:
                f.write("OK" if ok else f"MISMATCH {got.shape[0]} {want.size}")
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_gather_over_gloo_world2(tmp_path):
    """One process per shard, torch.distributed (gloo) gather to rank 0, concatenation in rank order."""
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    out = tmp_path / "result.txt"
    mp.spawn(_gloo_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert out.read_text() == "OK"


def test_gather_over_gloo_world4_with_an_empty_rank(tmp_path):
    """Four ranks, one of them with no match at all (its send and the matching receive are both
    skipped); the result on rank 0 is still the unsharded stream."""
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    out = tmp_path / "result4.txt"
    mp.spawn(_gloo_worker, args=(4, _free_port(), str(out)), nprocs=4, join=True)
    assert out.read_text() == "OK"
