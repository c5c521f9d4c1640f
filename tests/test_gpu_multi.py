"""Several GPUs behind the C ABI (include/olm_b200.h "several GPUs", csrc/multi.cpp) -- needs a B200.

A matcher that owns several engines shards every host haystack by byte range and must deliver
exactly the single-GPU / reference result: same records, same order, for plain and transforming
stores, with every filter (no_overlap is the one that crosses shards).  On a box with one GPU the
engines all live on GPU 0 -- the sharding, ownership and merge logic is the same; with more GPUs
the test uses them (peer copies over NVLink for the no_overlap gather).
"""
import numpy as np
import pytest

import inputs
from conftest import describe_diff, same_matches
from omega_match_b200 import Matcher, _lib
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

MIB = 1 << 20
FLAGSETS = [{}, {"no_overlap": True}, {"longest_only": True, "no_overlap": True}, {"word_boundary": True},
            {"line_end": True, "longest_only": True, "no_overlap": True}, {"word_prefix": True, "word_suffix": True}]


def _devices(n):
    have = _lib.load().olm_cuda_device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("sf", [(0, 0, 0), (1, 0, 0), (1, 1, 1)])
@pytest.mark.parametrize("ngpu", [2, 3])
def test_multi_matcher_equals_the_oracle(store_cache, sf, ngpu):
    names = inputs.case_patterns(dict(patterns="names", store_flags=sf))
    path = store_cache("multi-names", names, sf)
    o = Oracle.from_olm(path)
    n = 13 * MIB + 4321
    hay = inputs.text_haystack(n, 31 + sum(sf))
    # matches across every shard edge the plans below can produce (4 KiB / 4 MiB units)
    for edge in range(MIB, n - 8, MIB):
        hay[edge - 5:edge + 6] = np.frombuffer(b"Christopher", dtype=np.uint8)
    with Matcher(path, devices=_devices(ngpu)) as m:
        assert m.device_count == ngpu
        for kw in FLAGSETS:
            got = m.match_arrays(hay, **kw)
            want = o.match(hay, **kw)
            assert same_matches(got, want), f"{sf} x{ngpu} {kw}: " + describe_diff(got, want)
        # small inputs: fewer units than GPUs, empty shards
        for small in (b"", b"x", hay[:5000].tobytes(), hay[:4 * MIB + 5].tobytes()):
            got = m.match_arrays(np.frombuffer(small, dtype=np.uint8), no_overlap=True)
            want = o.match(small, no_overlap=True)
            assert same_matches(got, want), describe_diff(got, want)


def test_multi_matcher_statistics_add_up(store_cache):
    """Exact statistics of a sharded call = the single call's (every position belongs to one shard)."""
    pats = inputs.case_patterns(dict(patterns="names", store_flags=(0, 0, 0)))
    path = store_cache("multi-names", pats, (0, 0, 0))
    hay = inputs.text_haystack(6 * MIB + 17, 5)
    with Matcher(path) as one, Matcher(path, devices=_devices(3)) as many:
        for m in (one, many):
            m.set_exact_stats(True)
            m.match_arrays(hay, word_boundary=True)
        assert one.get_match_stats() == many.get_match_stats()


def test_comm_gather_world_of_one(store_cache):
    """The NCCL gather with a single rank: counts, copy into the gather buffer, no_overlap on the whole."""
    torch = pytest.importorskip("torch")
    pats = inputs.case_patterns(dict(patterns="names", store_flags=(0, 0, 0)))
    path = store_cache("multi-names", pats, (0, 0, 0))
    o = Oracle.from_olm(path)
    hay = inputs.text_haystack(3 * MIB + 99, 77)
    dev = torch.from_numpy(np.concatenate([hay, np.zeros(64, dtype=np.uint8)])).cuda()
    torch.cuda.synchronize()
    with Matcher(path) as m:
        m.comm_init(Matcher.comm_unique_id(), 0, 1)
        cnt, ptr = m.match_shard(dev.data_ptr(), 0, hay.size, 0, hay.size, hay.size, 0)
        total, gptr = m.gather_records(ptr, cnt, 0, no_overlap=True)
        rec = np.zeros(total, dtype=[("offset", "<u8"), ("len", "<u4"), ("pad", "<u4"), ("match", "<u8")])
        if total:
            src = torch.as_tensor(_DevArray(gptr, total), device="cuda")  # plain device memory, zero-copy view
            rec_np = src.cpu().numpy()
            rec["offset"], rec["len"] = rec_np[:, 0].astype(np.uint64), (rec_np[:, 1] & 0xFFFFFFFF).astype(np.uint32)
        want = o.match(hay, no_overlap=True)
        got = np.zeros(total, dtype=want.dtype)
        got["offset"], got["len"] = rec["offset"], rec["len"]
        assert same_matches(got, want), describe_diff(got, want)


class _DevArray:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count, 3), "typestr": "<i8", "data": (ptr, False), "version": 2}
