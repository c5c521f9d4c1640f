"""Parity of the CUDA path against the oracle / the reference's goldens -- needs a B200.

Every comparison goes through the C ABI of libomega_match.so (omega_list_matcher_match with a
host buffer, or the olm_cuda_* device entry points) and is bit-exact: same (offset, len)
records in the same order.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import inputs
from conftest import describe_diff, same_matches
from omega_match_b200 import Compiler, Matcher, MatchStats, RECORD_DTYPE
from omega_match_b200.sharding import WINDOW, shard_plan
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

VECTORS = json.loads((inputs.GOLDEN / "vectors.json").read_text())["vectors"]
_CASES = {c["name"]: c for c in inputs.vector_cases()}


def check(m: Matcher, o: Oracle, hay, **flags):
    got = m.match_arrays(hay, **flags)
    want = o.match(hay, **flags)
    assert same_matches(got, want), f"{flags}: " + describe_diff(got, want)
    return got


# ---- reference goldens through the product -------------------------------------------------

def test_kjv_goldens(store_cache):
    """names.txt x pseudo-KJV: data/matcher_found.txt and data/grep_found.txt (BASELINE configs[0])."""
    with Matcher(store_cache("names", inputs.golden_data("names.txt"))) as m:
        hay = inputs.pseudo_kjv()
        got = m.match_arrays(np.frombuffer(hay, dtype=np.uint8))
        want = inputs.parse_expected("matcher_found.txt")
        assert same_matches(got, want), describe_diff(got, want)
        got = m.match_arrays(np.frombuffer(hay, dtype=np.uint8), longest_only=True, no_overlap=True)
        want = inputs.parse_expected("grep_found.txt")
        assert same_matches(got, want), describe_diff(got, want)


@pytest.mark.parametrize("pats,hay,flags,expected", [
    ("usernames.txt", "haystack_email.txt", dict(word_prefix=True), "expected_word_prefix.txt"),
    ("tlds.txt", "haystack_email.txt", dict(word_suffix=True), "expected_word_suffix.txt"),
    ("line_anchor_patterns.txt", "line_anchor_haystack.txt",
     dict(line_start=True, longest_only=True, no_overlap=True), "expected_line_start.txt"),
    ("line_anchor_patterns.txt", "line_anchor_haystack.txt",
     dict(line_end=True, longest_only=True, no_overlap=True), "expected_line_end.txt"),
    ("line_anchor_patterns.txt", "line_anchor_haystack.txt",
     dict(line_start=True, word_boundary=True, longest_only=True, no_overlap=True),
     "expected_line_start_word_boundary.txt"),
    ("line_exact_match_patterns.txt", "line_exact_match_haystack.txt",
     dict(line_start=True, line_end=True, longest_only=True, no_overlap=True), "expected_line_exact_match.txt"),
])
def test_reference_cli_goldens(store_cache, pats, hay, flags, expected):
    with Matcher(store_cache(pats, inputs.golden_data(pats))) as m:
        got = m.match_arrays(inputs.golden_data(hay), **flags)
        want = inputs.parse_expected(expected)
        assert same_matches(got, want), describe_diff(got, want)


# ---- the reference's own CLI, relinked against the product ----

CLI = inputs.GOLDEN.parent.parent / "oracle" / "_ref" / "olm_b200"


@pytest.mark.skipif(not CLI.exists(), reason="oracle/_ref/olm_b200 not built (needs the reference checkout at build time)")
@pytest.mark.parametrize("pats,hay,cflags,mflags,expected", [
    ("names.txt", "pseudo-kjv", [], [], "matcher_found.txt"),
    ("names.txt", "pseudo-kjv", [], ["--longest", "--no-overlap"], "grep_found.txt"),
    ("usernames.txt", "haystack_email.txt", [], ["--word-prefix"], "expected_word_prefix.txt"),
    ("tlds.txt", "haystack_email.txt", [], ["--word-suffix"], "expected_word_suffix.txt"),
    ("line_anchor_patterns.txt", "line_anchor_haystack.txt", [], ["--line-start", "--longest", "--no-overlap"],
     "expected_line_start.txt"),
    ("line_exact_match_patterns.txt", "line_exact_match_haystack.txt", [],
     ["--line-start", "--line-end", "--longest", "--no-overlap"], "expected_line_exact_match.txt"),
])
def test_reference_cli_relinked(tmp_path, pats, hay, cflags, mflags, expected):
    """`olm compile` + `olm match` of the UNMODIFIED omega_match/main.c (main.c:400-464 calls the
    library, :89-133 prints `offset:bytes`) linked against libomega_match.so of this repository:
    the command lines of the reference's ctest goldens (CMakeLists.txt smoke / aio_* tests)
    produce the golden files byte for byte."""
    import subprocess
    (tmp_path / "p.txt").write_bytes(inputs.golden_data(pats))
    (tmp_path / "h.txt").write_bytes(inputs.pseudo_kjv() if hay == "pseudo-kjv" else inputs.golden_data(hay))
    olm, out = str(tmp_path / "p.olm"), str(tmp_path / "out.txt")
    subprocess.run([str(CLI), "compile", *cflags, olm, str(tmp_path / "p.txt")], check=True, capture_output=True)
    subprocess.run([str(CLI), "match", *mflags, "-o", out, olm, str(tmp_path / "h.txt")], check=True, capture_output=True)
    got, want = open(out, "rb").read(), inputs.golden_data(expected)
    assert got.replace(b"\r\n", b"\n") == want.replace(b"\r\n", b"\n"), (len(got), len(want), got[:200], want[:200])


# ---- outputs of the reference library on seeded inputs (tests/golden/vectors.json) -----------

@pytest.mark.parametrize("name", sorted(_CASES))
def test_reference_vectors(store_cache, name):
    """All 17 match-flag sets x 17 cases = the perf_test.py matrix (+ word-prefix/suffix), at
    sizes that cross 4 MiB windows.  The calls are replayed in generation order on ONE matcher:
    with word_boundary the reference's result depends on what earlier calls left in its scratch
    buffer, and so must ours."""
    case = _CASES[name]
    hay = inputs.case_haystack(case)
    path = store_cache(name + "-store", inputs.case_patterns(case), case["store_flags"])
    with Matcher(path) as m:
        for v in [v for v in VECTORS if v["case"] == name]:
            got = m.match_arrays(hay, **{k: True for k in v["flags"]})
            assert got.size == v["count"], (name, v["flags"], got.size, v["count"])
            assert f"{Oracle.stream_digest(got):016x}" == v["digest"], (name, v["flags"])


# ---- product vs oracle, edge cases -------------------------------------------------------

EDGE_PATTERNS = [b"a", b"ab", b"abc", b"abcd", b"abcde", b"abcdefghij", b"bcd", b"cd", b"d", b"zz", b"hello world",
                 b"x" * 40, b"x" * 200, b"the", b"King", b" ", b"\n", b"e\n", b"line"]


@pytest.mark.parametrize("sf", inputs.STORE_FLAG_SETS)
def test_edge_haystacks(store_cache, sf):
    """Empty and tiny inputs, matches at both ends, patterns longer than the staged halo, every
    prefix length around a tile edge."""
    buf = b"\n".join(p for p in EDGE_PATTERNS if inputs.py_normalize(p, *sf)) + b"\n"
    path = store_cache("edge", buf, sf)
    o = Oracle.from_olm(path)
    hays = [b"", b"a", b"ab", b"abc", b"abcd", b"abcde", b"d", b"xabcdefghij", b"abcdefghij" * 3,
            b"x" * 39, b"x" * 40, b"x" * 41, b"x" * 450, b"hello world", b"Hello,  World!\n", b"the King\nline\n",
            b"abcd" * 1000, (b"abcde " * 7000)[:32768 + 5], b"q" * 32766 + b"abcdefghij", b"q" * 32760 + b"x" * 220]
    with Matcher(path) as m:
        for hay in hays:
            for flags in ({}, {"word_boundary": True}, {"longest_only": True, "no_overlap": True},
                          {"line_start": True}, {"line_end": True, "word_suffix": True}, {"word_prefix": True}):
                check(m, o, hay, **flags)


def test_every_length_near_the_end(store_cache):
    """remaining-bytes conditions (matcher.c:782, :810, :828, :846) for every tail length."""
    pats = [b"a", b"aa", b"aaa", b"aaaa", b"aaaaa", b"aaaaaa", b"aaaaaaaaa"]
    path = store_cache("alla", b"\n".join(pats))
    o = Oracle.from_olm(path)
    with Matcher(path) as m:
        for n in list(range(0, 40)) + [511, 512, 513, 32767, 32768, 32769, 32768 + 111, 32768 + 112, 32768 + 113]:
            hay = b"a" * n
            check(m, o, hay)
            check(m, o, hay, longest_only=True)
            check(m, o, hay, no_overlap=True)


def test_dense_matches_overflow_the_staging_area(store_cache):
    """Many matches per position: the tile is redone writing to HBM directly; also the worst
    case for the no-overlap chain (every record overlaps the next)."""
    pats = [b"a" * k for k in range(1, 13)] + [b"ab", b"ba"]
    path = store_cache("dense", b"\n".join(pats))
    o = Oracle.from_olm(path)
    hay = (b"a" * 70000) + b"ab" * 5000 + b"a" * 1000
    with Matcher(path) as m:
        for flags in ({}, {"longest_only": True}, {"no_overlap": True}, {"longest_only": True, "no_overlap": True}):
            check(m, o, hay, **flags)


def test_long_patterns_cross_tiles_and_halo(store_cache):
    rng = np.random.default_rng(5)
    pats = [bytes(rng.integers(97, 123, size=n, dtype=np.uint8)) for n in (5, 8, 9, 16, 100, 113, 114, 500, 5000)]
    path = store_cache("longpats", b"\n".join(pats))
    o = Oracle.from_olm(path)
    hay = bytearray(rng.integers(97, 123, size=200_000, dtype=np.uint8).tobytes())
    for i, p in enumerate(pats):
        for at in (32768 - len(p) // 2, 65536 - 3, 100_000 + 37 * i, 200_000 - len(p)):
            if at >= 0:
                hay[at:at + len(p)] = p
    with Matcher(path) as m:
        check(m, o, bytes(hay))
        check(m, o, bytes(hay), word_boundary=True)


def test_window_edges_and_stale_tail(store_cache):
    """SURVEY F4/F5/H6 through the CUDA path."""
    W = WINDOW
    path = store_cache("hw", b"HELLOWORLD\nhello\nzq\nab\n", (1, 0, 0))
    o = Oracle.from_olm(path)
    hay = bytearray(b"." * (W + 100))
    hay[W - 5:W + 5] = b"helloworld"
    with Matcher(path) as m:
        for flags in ({}, {"line_start": True}, {"word_prefix": True}, {"line_end": True}, {"word_suffix": True}):
            check(m, o, bytes(hay), **flags)
        hay2 = bytearray(b"." * W + b"y" * 96 + b" ab")
        got = check(m, o, bytes(hay2), word_boundary=True)
        assert [(int(a), int(b)) for a, b in zip(got["offset"], got["len"])] == [(W + 97, 2)]
        hay2[99] = ord("Q")  # what window 0 leaves at scratch index 99 decides the match in window 1
        got = check(m, o, bytes(hay2), word_boundary=True)
        assert got.size == 0
        # a window that ends in a literal space is trimmed even without elide-whitespace
        hay3 = bytearray(b"x" * (W - 3) + b"ab " + b"zq rest")
        check(m, o, bytes(hay3), word_boundary=True)
        check(m, o, bytes(hay3))


@pytest.mark.parametrize("sf", [(0, 1, 0), (1, 1, 1), (0, 0, 1)])
def test_stale_tail_across_calls_of_a_normalising_store(store_cache, sf):
    """SURVEY H6 for stores that DROP bytes: what an earlier call -- with or without word_boundary, i.e.
    through either way the scratch-buffer image is kept (transform.cu) -- left at index M_w of the
    reference's scratch buffer decides a 2..4 byte match at the very end of a later, shorter haystack."""
    path = store_cache("stale-xf", b"ab\nzq\nhello\nabc\n", sf)
    o = Oracle.from_olm(path)
    rng = np.random.default_rng(7)
    with Matcher(path) as m:
        for rnd in range(6):
            # a long haystack: words and dropped bytes; then shorter ones that END in a short pattern
            long_hay = bytes(rng.choice(np.frombuffer(b"abQz  ..,-q\n", dtype=np.uint8), size=int(rng.integers(3000, 9000))))
            for flags in ({}, {"word_boundary": True}, {"longest_only": True, "no_overlap": True}):
                check(m, o, long_hay, **flags)
                for cut in rng.integers(5, 2500, size=4):
                    short = long_hay[:int(cut)].rstrip(b" .,-\n") + b".ab"
                    check(m, o, short, word_boundary=True)
                    check(m, o, short + b"..", word_boundary=True)
        # across 4 MiB windows: the second window is shorter than the first and ends in "zq"
        W = WINDOW
        hay = bytearray(rng.choice(np.frombuffer(b"abQz ,.", dtype=np.uint8), size=W + 5000).tobytes())
        hay[-2:] = b"zq"
        check(m, o, bytes(hay))
        check(m, o, bytes(hay), word_boundary=True)
        hay[W - 40:W] = b"." * 40
        check(m, o, bytes(hay), word_boundary=True)


@pytest.mark.parametrize("seed", range(6))
def test_random_stores_and_haystacks(store_cache, seed):
    """Differential fuzz against the oracle: random pattern sets (all lengths), random flags."""
    import random
    rng = random.Random(1000 + seed)
    alph = b"abcABC xyz.,-'\n\r\t_09"
    pats = set()
    while len(pats) < rng.choice([3, 30, 300]):
        pats.add(bytes(rng.choice(b"abcABC xyz.-'_09") for _ in range(rng.choice([1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 9, 12]))))
    sf = (int(rng.random() < 0.5), int(rng.random() < 0.4), int(rng.random() < 0.4))
    buf = b"\n".join(sorted(p for p in pats if inputs.py_normalize(p, *sf)))
    path = store_cache(f"fuzz{seed}", buf, sf)
    o = Oracle.from_olm(path)
    with Matcher(path) as m:
        for _ in range(12):
            n = rng.choice([0, 1, 3, 4, 5, 17, 300, 4000, 40000, 70001])
            hay = bytes(rng.choice(alph) for _ in range(n))
            kw = {f: rng.random() < 0.3 for f in ("no_overlap", "longest_only", "word_boundary", "word_prefix",
                                                 "word_suffix", "line_start", "line_end")}
            check(m, o, hay, **kw)


# ---- the reference binding's own known-answer tests, on our mirror ---------------------------

def write_file(path, lines):
    path.write_text("\n".join(lines), encoding="utf-8")


def test_binding_compile_and_match(tmp_path):
    """bindings/python/tests/test_omega_match.py:41-77."""
    pat_file = tmp_path / "patterns.txt"
    write_file(pat_file, ["foo", "bar", "bazinga"])
    compiled = str(tmp_path / "matcher.olm")
    Compiler.compile_from_filename(compiled, str(pat_file))
    with Matcher(compiled) as m2:
        results = m2.match(b"xx foobar yy foo zz bar")
        assert [r.offset for r in results] == [3, 6, 13, 20]
        assert [r.length for r in results] == [3, 3, 3, 3]
        assert [r.match for r in results] == [b"foo", b"bar", b"foo", b"bar"]
        st = m2.get_match_stats()
        assert isinstance(st, MatchStats) and st.total_hits == len(results)
        m2.reset_match_stats()
        assert m2.get_match_stats() == MatchStats(0, 0, 0, 0, 0)


@pytest.mark.parametrize("sf", [(0, 0, 0), (1, 0, 0), (1, 1, 1)])
def test_exact_stats_equal_the_reference_counters(store_cache, sf):
    """omega_match_stats_t (list_matcher.h:43-49; matcher.c:783-799, :818-877, :210, accumulated
    :887-893): with olm_cuda_set_exact_stats() all five counters equal the oracle's -- which
    tests/ref_fuzz_worker.py pins on the reference's -- for long + short pattern sets, with and
    without word_boundary, across 4 MiB windows, and accumulate over calls."""
    names = [p for p in inputs.golden_data("names.txt").split(b"\n") if p][:6000]
    pats = names + inputs.synth_long_patterns(3000) + [b"ab", b"the", b"King", b"x", b"of ", b"e\n"]
    path = store_cache("stats-mixed", b"\n".join(pats), sf)
    long_only = store_cache("stats-long", b"\n".join(inputs.synth_long_patterns(20000)), sf)
    hay = inputs.plant(inputs.text_haystack((9 << 20) + 12345, 11), pats, 0x51)
    for store in (path, long_only):
        o = Oracle.from_olm(store)
        with Matcher(store) as m:
            m.set_exact_stats(True)
            for flags in ({}, {"word_boundary": True}, {"longest_only": True, "no_overlap": True},
                          {"word_prefix": True, "line_end": True}):
                for h in (hay, hay[:70001], hay[:3], b""):
                    got = m.match_arrays(h, **flags)
                    want = o.match(h, **flags)
                    assert same_matches(got, want), describe_diff(got, want)
                    st = m.get_match_stats()
                    ref = o.stats.as_dict()
                    assert (st.total_hits, st.total_misses, st.total_filtered, st.total_attempts,
                            st.total_comparisons) == (ref["hits"], ref["misses"], ref["filtered"], ref["attempts"],
                                                      ref["comparisons"]), (store == path, flags, len(h), st, ref)
            # default mode again: no extra kernel, the scan's own counters
            m.match_arrays(hay[:70001])
            on = m.last_timing()["kernel_launches"]
            m.set_exact_stats(False)
            m.match_arrays(hay[:70001])
            assert m.last_timing()["kernel_launches"] == on - 1


def test_binding_flags(tmp_path):
    """test_omega_match.py:107-196, :242-325 (case, punctuation, overlap, word and line options)."""
    pat_file = tmp_path / "p.txt"
    write_file(pat_file, ["Foo", "BaR"])
    with Matcher(str(pat_file), case_insensitive=True) as m:
        r = m.match(b"foo BAR Baz fooBar")
        assert [x.offset for x in r] == [0, 4, 12, 15] and [x.match for x in r] == [b"foo", b"BAR", b"foo", b"Bar"]
    compiled = str(tmp_path / "m.olm")
    Compiler.compile_from_buffer(compiled, b"f'oo\nbar\n", ignore_punctuation=True, case_insensitive=True)
    with Matcher(compiled) as m:
        r = m.match(b"f'oo BAR Baz fooBar")
        assert [x.offset for x in r] == [0, 5, 13, 16] and [x.match for x in r] == [b"f'oo", b"BAR", b"foo", b"Bar"]
    write_file(pat_file, ["abc", "abcd"])
    with Matcher(str(pat_file)) as m:
        assert {x.match for x in m.match(b"xxabcdyy")} == {b"abc", b"abcd"}
        assert [x.match for x in m.match(b"xxabcdyy", longest_only=True)] == [b"abcd"]
        assert [x.match for x in m.match(b"xxabcdyy", no_overlap=True)] == [b"abcd"]
    write_file(pat_file, ["in", "and"])
    with Matcher(str(pat_file)) as m:
        r = m.match(b"land and inland", word_boundary=True)
        assert [(x.offset, x.match) for x in r] == [(5, b"and")]
    write_file(pat_file, ["foo", "bar"])
    with Matcher(str(pat_file)) as m:
        assert [x.offset for x in m.match(b"foobar foo barbar", word_prefix=True)] == [0, 7, 11]
        assert [x.offset for x in m.match(b"foofoo toolbar bar", word_suffix=True)] == [3, 11, 15]
    write_file(pat_file, ["start", "end", "middle"])
    with Matcher(str(pat_file)) as m:
        hay = b"start of line\nmiddle start here\nsome middle text\nline end"
        assert [x.offset for x in m.match(hay, line_start=True)] == [0, 14]
        assert [x.offset for x in m.match(hay, line_end=True)] == [54]
        assert m.match(hay, line_start=True, line_end=True) == []
    write_file(pat_file, ["exactline"])
    with Matcher(str(pat_file)) as m:
        r = m.match(b"before\nexactline\nafter", line_start=True, line_end=True)
        assert [(x.offset, x.match) for x in r] == [(7, b"exactline")]


def test_binding_threads_and_chunk(tmp_path):
    """test_omega_match.py:198-239: the setters keep the reference's validation and defaults."""
    pat_file = tmp_path / "p.txt"
    write_file(pat_file, ["foo", "bar"])
    with Matcher(str(pat_file)) as m:
        m.set_threads(1)
        assert m.get_threads() == 1
        m.set_chunk_size(1024)
        assert m.get_chunk_size() == 1024
        m.set_chunk_size(1000)
        assert m.get_chunk_size() == 1024
        assert [x.offset for x in m.match(b"xx foobar yy foo zz bar")] == [3, 6, 13, 20]
        m.set_threads(0)
        assert m.get_threads() > 0
        m.set_chunk_size(0)
        assert m.get_chunk_size() == 4096
        with pytest.raises(ValueError):
            m.set_threads(-1)
        with pytest.raises(ValueError):
            m.set_chunk_size(-1)
        with pytest.raises(TypeError):
            m.match("not bytes")


def test_create_errors(tmp_path):
    with pytest.raises(RuntimeError):
        Matcher(str(tmp_path / "missing.olm"))
    bad = tmp_path / "bad.olm"
    good = tmp_path / "good.olm"
    Compiler.compile_from_buffer(str(good), b"alpha\nbeta\n")
    bad.write_bytes(good.read_bytes()[:-3])
    with pytest.raises(RuntimeError):
        Matcher(str(bad))


def test_result_pointers_alias_the_haystack(store_cache):
    """list_matcher.h:19-23: result.match points into the caller's buffer (checked in _match_records)."""
    with Matcher(store_cache("names", inputs.golden_data("names.txt"))) as m:
        hay = inputs.text_haystack(50000, 3)
        rec = m._match_records(hay)
        assert rec.dtype == RECORD_DTYPE and rec.size > 0


# ---- device-resident entry points, shards, filters, sort -------------------------------------

def _device_records(torch, ptr, count):
    class Dev:
        __cuda_array_interface__ = {"shape": (count, 3), "typestr": "<i8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(Dev(), device="cuda").clone() if count else torch.empty((0, 3), dtype=torch.int64, device="cuda")


def _as_matches(t):
    a = t.cpu().numpy()
    out = np.zeros(a.shape[0], dtype=[("offset", "<u8"), ("len", "<u4"), ("_pad", "<u4")])
    out["offset"] = a[:, 0].astype(np.uint64)
    out["len"] = (a[:, 1] & 0xFFFFFFFF).astype(np.uint32)
    return out


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("kind", ["plain", "windowed"])
def test_shards_concatenate_to_the_whole(store_cache, world, kind):
    """SURVEY 8e on one GPU: the 2/4/8-way partitions, scanned shard by shard with
    olm_cuda_match_shard from per-shard device slices, concatenate (in rank order, then one
    no-overlap pass) to exactly the single-call result."""
    torch = pytest.importorskip("torch")
    if kind == "plain":
        pats = inputs.synth_long_patterns(5000) + [b"ab", b"the", b"o", b"King"]
        hay = inputs.plant(inputs.synth_haystack(3_000_000, 123), pats, 9, block=2048)
        sf, largest = (0, 0, 0), 24
    else:
        pats = [p for p in inputs.golden_data("names.txt").split(b"\n") if p][::5]
        hay = inputs.text_haystack(9 * WINDOW + 4321, 21)
        sf, largest = (1, 1, 1), 30
    path = store_cache(f"shard-{kind}", b"\n".join(pats), sf)
    o = Oracle.from_olm(path)
    dev = torch.from_numpy(hay).cuda()
    with Matcher(path) as m:
        for flags in ({}, {"longest_only": True, "word_boundary": True}, {"no_overlap": True}):
            want = o.match(hay, **flags)
            parts = []
            for s in shard_plan(hay.size, world, largest, windowed=any(sf)):
                sl = torch.zeros(((s.slice_end - s.slice_begin + 15) // 16) * 16 + 16, dtype=torch.uint8, device="cuda")
                sl[:s.slice_end - s.slice_begin] = dev[s.slice_begin:s.slice_end]
                fl = {k: v for k, v in flags.items() if k != "no_overlap"}
                cnt, ptr = m.match_shard(sl.data_ptr(), s.slice_begin, s.slice_end - s.slice_begin, s.own_begin,
                                         s.own_end, hay.size, 0, **fl)
                parts.append(_device_records(torch, ptr, cnt))
            allrec = torch.cat(parts).contiguous()
            n = allrec.shape[0]
            if flags.get("no_overlap"):
                n = m.no_overlap_device(allrec.data_ptr(), n)
            got = _as_matches(allrec[:n])
            assert same_matches(got, want), f"{kind} x{world} {flags}: " + describe_diff(got, want)


@pytest.mark.parametrize("kind", ["plain", "windowed"])
def test_shard_from_host_memory_streams_like_the_device_path(store_cache, kind):
    """olm_cuda_match_shard_host: a rank's slice in HOST memory, long enough (> 512 MiB) for the
    segmented copy to overlap the scan, gives the records olm_cuda_match_shard gives for the same
    slice on the device; both halves of a 2-way plan concatenate to the whole."""
    torch = pytest.importorskip("torch")
    import synth_torch
    sf = (0, 0, 0) if kind == "plain" else (1, 0, 1)
    pats = inputs.synth_long_patterns(3000) + ([b"the", b"ab"] if kind == "plain" else [])
    path = store_cache("host-shard-" + kind, b"\n".join(pats), sf)
    n = (1200 << 20) + 4096 * 3 + 77
    hay = synth_torch.synth_haystack_torch(n, inputs.SEED_H5, device="cuda")
    pb, pl = synth_torch.pack_patterns(pats, "cuda")
    synth_torch.plant_torch(hay, pb, pl, 0x99)
    torch.cuda.synchronize()
    with Matcher(path) as m:
        cnt, ptr = m.match_device(hay.data_ptr(), n)
        whole = _as_matches(_device_records(torch, ptr, cnt))
        parts = []
        for s in shard_plan(n, 2, 24, windowed=any(sf)):
            ln = s.slice_end - s.slice_begin
            host = torch.empty(ln + 64, dtype=torch.uint8, pin_memory=True)
            host[:ln].copy_(hay[s.slice_begin:s.slice_end])
            torch.cuda.synchronize()
            c1, p1 = m.match_shard_host(host.data_ptr(), s.slice_begin, ln, s.own_begin, s.own_end, n, 0)
            a = _device_records(torch, p1, c1).clone()
            assert m.last_timing()["h2d_ms"] > 0 and m.last_timing()["scan_launches"] >= (2 if kind == "plain" else 1)
            dev = torch.zeros(((ln + 15) // 16) * 16 + 16, dtype=torch.uint8, device="cuda")
            dev[:ln] = hay[s.slice_begin:s.slice_end]
            torch.cuda.synchronize()
            c2, p2 = m.match_shard(dev.data_ptr(), s.slice_begin, ln, s.own_begin, s.own_end, n, 0)
            b = _device_records(torch, p2, c2)
            assert c1 == c2 and bool((a[:, :2] == b[:, :2]).all())
            parts.append(a)
        assert Oracle.stream_digest(_as_matches(torch.cat(parts))) == Oracle.stream_digest(whole)


def test_match_device_and_sort(store_cache):
    torch = pytest.importorskip("torch")
    pats = inputs.synth_long_patterns(2000) + [b"ab", b"the", b"abc", b"abcd"]
    hay = inputs.plant(inputs.synth_haystack(1_000_000, 5), pats, 3, block=1024)
    path = store_cache("devapi", b"\n".join(pats))
    o = Oracle.from_olm(path)
    dev = torch.zeros(hay.size + 64, dtype=torch.uint8, device="cuda")
    dev[:hay.size] = torch.from_numpy(hay).cuda()
    with Matcher(path) as m:
        for flags in ({}, {"no_overlap": True}, {"longest_only": True, "line_end": True}):
            cnt, ptr = m.match_device(dev.data_ptr(), hay.size, **flags)
            rec = _device_records(torch, ptr, cnt)
            want = o.match(hay, **flags)
            assert same_matches(_as_matches(rec), want)
            assert bool((rec[:, 2] == rec[:, 0] + dev.data_ptr()).all())  # match = base + offset
        # the LSD radix sort (matcher.c:258-325 order): shuffle the records, sort, compare
        cnt, ptr = m.match_device(dev.data_ptr(), hay.size)
        rec = _device_records(torch, ptr, cnt)
        perm = torch.randperm(cnt, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        shuffled = rec[perm].contiguous()
        m.sort_records_device(shuffled.data_ptr(), cnt)
        assert bool((shuffled == rec).all())
        t = m.last_timing()
        assert t["scan_ms"] > 0 and t["kernel_launches"] >= 1


def test_shard_without_its_halo_is_rejected_and_empty_calls_add_no_statistics(store_cache):
    """A byte-range shard has to bring 16 bytes in front of its first owned position and the longest
    pattern + 1 behind the last one: the call fails instead of reading what is not there.
    An empty haystack adds nothing to an attached statistics struct (matcher.c:887-893)."""
    torch = pytest.importorskip("torch")
    path = store_cache("names", inputs.golden_data("names.txt"))
    hay = inputs.text_haystack(1 << 20, 3)
    dev = torch.from_numpy(np.concatenate([hay, np.zeros(64, dtype=np.uint8)])).cuda()
    torch.cuda.synchronize()
    n, half = hay.size, 1 << 19
    with Matcher(path) as m:
        ok = m.match_shard(dev.data_ptr() + half - 16, half - 16, n - half + 16, half, n, n, 0)
        assert ok[0] > 0
        with pytest.raises(RuntimeError):  # no front halo
            m.match_shard(dev.data_ptr() + half, half, n - half, half, n, n, 0)
        with pytest.raises(RuntimeError):  # no back halo: the slice ends with the owned range
            m.match_shard(dev.data_ptr(), 0, half, 0, half, n, 0)
        m.match_arrays(hay)
        before = m.get_match_stats()
        assert m.match_arrays(b"").size == 0
        assert m.get_match_stats() == before


def test_gpu_listing_equals_the_cli_format(store_cache):
    """`olm match` prints "offset:bytes\\n" per match with snprintf("%zu:%.*s\\n")
    (omega_match/main.c:89-133), so a line's bytes stop at a NUL inside the match.  The same text,
    formatted on the GPU from the device records (olm_cuda_format_records)."""
    torch = pytest.importorskip("torch")
    names = [p for p in inputs.golden_data("names.txt").split(b"\n") if p]
    pats = names[:4000] + [b"ab\x00cd", b"\x00x", b"zz\x00", b"q"]
    path = store_cache("listing", b"\n".join(pats))
    hay = inputs.text_haystack((3 << 20) + 17, 123)
    hay[1000:1005] = np.frombuffer(b"ab\x00cd", dtype=np.uint8)
    hay[5000:5003] = np.frombuffer(b"zz\x00", dtype=np.uint8)
    hay[9000:9002] = np.frombuffer(b"\x00x", dtype=np.uint8)
    dev = torch.from_numpy(np.concatenate([hay, np.zeros(64, dtype=np.uint8)])).cuda()
    torch.cuda.synchronize()
    o = Oracle.from_olm(path)
    with Matcher(path) as m:
        for flags in ({}, {"longest_only": True, "no_overlap": True}):
            cnt, ptr = m.match_device(dev.data_ptr(), hay.size, **flags)
            tptr, tlen = m.format_records_device(ptr, cnt, dev.data_ptr(), 0)
            got = bytes(torch.as_tensor(_DevBytes(tptr, tlen), device="cuda").cpu().numpy()) if tlen else b""
            want = o.match(hay, **flags)
            hb = hay.tobytes()
            lines = []
            for off, ln in zip(want["offset"].tolist(), want["len"].tolist()):
                body = hb[off:off + ln]
                lines.append(str(off).encode() + b":" + body.split(b"\x00")[0] + b"\n")
            assert got == b"".join(lines)
            assert cnt == want.size and cnt > 1000
        assert m.format_records_device(0, 0, dev.data_ptr(), 0) == (0, 0)


class _DevBytes:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def test_large_synthetic_properties(store_cache):
    """A slice of BASELINE config 5 (256 MiB, 100k patterns) generated on the device: every
    planted pattern is reported, offsets ascend, shard digest equals the single-call digest,
    and a 4 MiB prefix agrees with the oracle record by record."""
    torch = pytest.importorskip("torch")
    import synth_torch
    pats = inputs.synth_long_patterns(100_000)
    path = store_cache("cfg5-100k", b"\n".join(pats))
    n = 256 << 20
    hay = synth_torch.synth_haystack_torch(n, inputs.SEED_H5, device="cuda")
    pb, pl = synth_torch.pack_patterns(pats, "cuda")
    planted = synth_torch.plant_torch(hay, pb, pl, inputs.SEED_H5 ^ 0x77)
    with Matcher(path) as m:
        cnt, ptr = m.match_device(hay.data_ptr(), n)
        rec = _device_records(torch, ptr, cnt)
        off, ln = rec[:, 0], rec[:, 1] & 0xFFFFFFFF
        assert cnt >= planted
        assert bool((off[1:] >= off[:-1]).all())
        same_off = off[1:] == off[:-1]
        assert bool((ln[1:][same_off] < ln[:-1][same_off]).all())  # length strictly descending within an offset
        # every 4 KiB block reports its planted pattern: check via block histogram
        blocks = torch.unique(off // 4096)
        assert blocks.numel() == planted
        # the reported bytes really are patterns: re-read a sample on the host
        sample = torch.randint(0, cnt, (2000,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
        hs, pset = hay.cpu().numpy(), set(pats)
        for o_, l_ in zip(off[sample].tolist(), ln[sample].tolist()):
            assert hs[o_:o_ + l_].tobytes() in pset
        whole = _as_matches(rec)
        # shards (4-way) give the identical stream
        parts = []
        for s in shard_plan(n, 4, 24, False):
            sl = torch.zeros(((s.slice_end - s.slice_begin + 15) // 16) * 16 + 16, dtype=torch.uint8, device="cuda")
            sl[:s.slice_end - s.slice_begin] = hay[s.slice_begin:s.slice_end]
            torch.cuda.synchronize()  # the library scans on its own stream: the bytes must be there
            c2, p2 = m.match_shard(sl.data_ptr(), s.slice_begin, s.slice_end - s.slice_begin, s.own_begin, s.own_end,
                                   n, 0)
            parts.append(_device_records(torch, p2, c2))
        sharded = _as_matches(torch.cat(parts))
        assert Oracle.stream_digest(sharded) == Oracle.stream_digest(whole)
        # prefix against the oracle
        pre = hs[:4 << 20]
        o = Oracle.from_olm(path)
        want = o.match(pre)
        got = m.match_arrays(pre)
        assert same_matches(got, want), describe_diff(got, want)


# ---- byte-class prefilter, lean path, streaming host path ------------------------------------

def _class_haystack(rng, n, alphabet, breakers, run_mean):
    """Runs of class bytes of random length separated by single non-class bytes, so that runs
    shorter than / equal to / longer than the prefilter's run length all occur."""
    out = np.empty(n, dtype=np.uint8)
    i = 0
    while i < n:
        r = int(rng.integers(1, 2 * run_mean))
        r = min(r, n - i)
        out[i:i + r] = rng.choice(alphabet, size=r)
        i += r
        if i < n:
            out[i] = rng.choice(breakers)
            i += 1
    return out


@pytest.mark.parametrize("kind", ["letters", "digits", "hex-two-ranges", "with-len4", "min8", "mixed-case-fold"])
def test_class_prefilter_stores(store_cache, kind):
    """Stores whose patterns start with bytes from a small class take the SWAR prefilter
    (device_tables.h ByteClass): runs just below / at / above the run length, class bytes at
    tile and chunk edges, every flag that the lean and the generic path handle."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(kind.encode()))
    lower = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz", dtype=np.uint8)
    upper = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ", dtype=np.uint8)
    digits = np.frombuffer(b"0123456789", dtype=np.uint8)
    if kind == "letters":
        alpha, lens = np.concatenate([lower, upper]), (6, 7, 9, 13, 24)
    elif kind == "digits":
        alpha, lens = digits, (5, 6, 11)
    elif kind == "hex-two-ranges":
        alpha, lens = np.concatenate([digits, lower[:6]]), (6, 8, 12, 17)
    elif kind == "with-len4":
        alpha, lens = lower, (4, 4, 5, 9)
    elif kind == "min8":
        alpha, lens = upper, (8, 9, 15, 30)
    else:
        alpha, lens = np.concatenate([lower[:13], upper[13:]]), (7, 10)
    few = alpha[::max(1, alpha.size // 6)][:6]  # a small alphabet (spread over the class) so that matches happen
    pats = sorted({bytes(rng.choice(few, size=int(rng.choice(lens)))) for _ in range(400)})
    breakers = np.frombuffer(b" \n.-_@\x00\xff\x80", dtype=np.uint8)
    hay = _class_haystack(rng, 200_000, few, breakers, run_mean=int(min(lens)))
    # plant patterns flush against run edges, chunk (512) and tile (16384) edges
    for k, pos in enumerate([0, 505, 512 - 3, 16384 - 5, 16384, 16384 * 3 - 1, 100_000, hay.size - 9, hay.size - 4]):
        p = np.frombuffer(pats[k % len(pats)], dtype=np.uint8)
        p = p[:max(0, hay.size - pos)]
        hay[pos:pos + p.size] = p
    path = store_cache(f"class-{kind}", b"\n".join(pats))
    o = Oracle.from_olm(path)
    with Matcher(path) as m:
        for flags in ({}, {"longest_only": True}, {"no_overlap": True, "longest_only": True}, {"word_boundary": True},
                      {"word_prefix": True}, {"line_start": True}, {"line_end": True, "word_suffix": True}):
            check(m, o, hay, **flags)
        check(m, o, hay[:5])
        check(m, o, hay[:517])


def test_lean_path_without_class_and_dense_redo(store_cache):
    """Gram-only stores that do NOT qualify for the class prefilter (bytes >= 0x80, wide byte
    spread) take the lean path with the bitmap probe in stage 1; a periodic haystack makes
    every position match several patterns so that chunks overflow their staging area."""
    rng = np.random.default_rng(11)
    pats = {bytes(rng.integers(0, 256, size=int(n), dtype=np.uint8)).replace(b"\n", b"\x01") for n in rng.integers(5, 40, size=300)}
    unit = b"\xc3\xa9t\xc3\xa9 "
    pats |= {(unit * 8)[i:i + n] for i in range(len(unit)) for n in (5, 6, 7, 12, 13, 20)}
    pats = sorted(pats)
    hay = np.frombuffer(unit * 20_000 + bytes(rng.integers(0, 256, size=50_000, dtype=np.uint8)) + unit * 3000, dtype=np.uint8).copy()
    for k in range(0, 40):
        p = np.frombuffer(pats[(7 * k) % len(pats)], dtype=np.uint8)
        pos = 120_000 + 1000 * k
        hay[pos:pos + p.size] = p
    path = store_cache("lean-dense", b"\n".join(pats))
    o = Oracle.from_olm(path)
    with Matcher(path) as m:
        for flags in ({}, {"longest_only": True}, {"no_overlap": True}, {"word_boundary": True}):
            check(m, o, hay, **flags)


def test_streaming_host_path_matches_device_path(store_cache):
    """Host buffers of 512 MiB and more are copied in 256 MiB segments that overlap with the
    scan (engine.cu match_host).  The record stream must equal the one of the device-resident
    call on the same bytes, for a plain and for a transforming store."""
    torch = pytest.importorskip("torch")
    import synth_torch
    n = (512 << 20) + 12345
    pats = inputs.synth_long_patterns(20_000) + [b"zq", b"The", b"of"]
    for sf in ((0, 0, 0), (1, 0, 1)):
        path = store_cache(f"stream-{sf}", b"\n".join(pats), sf)
        hay = synth_torch.synth_haystack_torch(n + 64, inputs.SEED_H5, device="cuda")
        pb, pl = synth_torch.pack_patterns(pats[:20_000], "cuda")
        synth_torch.plant_torch(hay[:n & ~4095], pb, pl, 99)
        host = hay[:n].cpu().numpy()
        with Matcher(path) as m:
            cnt, ptr = m.match_device(hay.data_ptr(), n, longest_only=True)
            dev = _as_matches(_device_records(torch, ptr, cnt)).copy()
            got = m.match_arrays(host, longest_only=True)
            assert same_matches(got, dev), describe_diff(got, dev)
            assert cnt > 100_000


def test_config5_at_full_size(store_cache):
    """BASELINE configs[4] at its full size -- 16 GiB synthetic haystack x 1,000,000 patterns --
    through size-independent properties: every planted pattern is reported (one per 4 KiB
    block), offsets ascend with lengths strictly descending inside an offset, reported bytes are
    patterns, two byte-range shards give the identical record stream, and a 2 MiB slice in the
    middle agrees with the oracle record by record."""
    torch = pytest.importorskip("torch")
    import synth_torch
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~45 GB of free device memory")
    pats = inputs.synth_long_patterns(1_000_000)
    path = store_cache("cfg5-1m", b"\n".join(pats))
    n = 16 << 30
    hay = synth_torch.synth_haystack_torch(n, inputs.SEED_H5, device="cuda")
    pb, pl = synth_torch.pack_patterns(pats, "cuda")
    planted = synth_torch.plant_torch(hay, pb, pl, inputs.SEED_H5 ^ 0x77)
    del pb, pl
    with Matcher(path) as m:
        cnt, ptr = m.match_device(hay.data_ptr(), n)
        rec = _device_records(torch, ptr, cnt).clone()
        off, ln = rec[:, 0], rec[:, 1] & 0xFFFFFFFF
        assert cnt >= planted == n // 4096
        assert bool((off[1:] >= off[:-1]).all())
        same_off = off[1:] == off[:-1]
        assert bool((ln[1:][same_off] < ln[:-1][same_off]).all())
        assert torch.unique(off // 4096).numel() == planted
        sample = torch.randint(0, cnt, (500,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
        pset = set(pats)
        for o_, l_ in zip(off[sample].tolist(), ln[sample].tolist()):
            assert hay[o_:o_ + l_].cpu().numpy().tobytes() in pset
        whole_digest = Oracle.stream_digest(_as_matches(rec))
        # two shards, each from its own copy of its slice (+ halo), as two GPUs would hold them
        parts = []
        for s in shard_plan(n, 2, 24, False):
            sl = torch.zeros(((s.slice_end - s.slice_begin + 15) // 16) * 16 + 16, dtype=torch.uint8, device="cuda")
            sl[:s.slice_end - s.slice_begin] = hay[s.slice_begin:s.slice_end]
            torch.cuda.synchronize()  # the library scans on its own stream: the bytes must be there
            c2, p2 = m.match_shard(sl.data_ptr(), s.slice_begin, s.slice_end - s.slice_begin, s.own_begin, s.own_end,
                                   n, 0)
            parts.append(_device_records(torch, p2, c2).clone())
            del sl
        assert Oracle.stream_digest(_as_matches(torch.cat(parts))) == whole_digest
        # a slice across the shard boundary against the oracle (matches that start inside it)
        lo = (n // 2) - (1 << 20)
        piece = hay[lo:lo + (2 << 20)].cpu().numpy()
        want = Oracle.from_olm(path).match(piece)
        got = m.match_arrays(piece)
        assert same_matches(got, want), describe_diff(got, want)
        inside = (off >= lo) & (off + ln <= lo + (2 << 20))
        assert int(inside.sum()) == want.size


def _colliding_patterns(k: int):
    """Two patterns whose first k bytes differ but hash to the same 32-bit key
    (device_tables.h key_hash: gram * M1 ^ tail * M2, M1 odd hence invertible)."""
    M1, M2, mask32 = 0x9E3779B1, 0x85EBCA6B, 0xFFFFFFFF
    inv = pow(M1, -1, 1 << 32)
    a = b"abcdefgh"[:k] + b"-first"
    gram_a = int.from_bytes(a[:4], "big")
    tail_mask = (1 << (8 * (k - 4))) - 1
    tail_a = int.from_bytes(a[4:8], "little") & tail_mask
    key = ((gram_a * M1) ^ (tail_a * M2)) & mask32
    for t in range(1, 1 << 16):
        tail_b = int.from_bytes(bytes([0x41 + t % 26, 0x61 + (t // 26) % 26, 0x30 + (t // 676) % 10, 0x42][:k - 4]).ljust(4, b"\0"), "little")
        gram_b = (((key ^ (tail_b * M2)) & mask32) * inv) & mask32
        head = gram_b.to_bytes(4, "big") + tail_b.to_bytes(4, "little")[:k - 4]
        if b"\n" in head or b"\r" in head or head == a[:k]:
            continue
        b = head + b"-second"
        assert (((int.from_bytes(b[:4], "big") * M1) ^ ((int.from_bytes(b[4:8], "little") & tail_mask) * M2)) & mask32) == key
        return a, b
    raise AssertionError("no collision found")


@pytest.mark.parametrize("k", [5, 6, 7, 8])
def test_key_collisions_between_different_prefixes(store_cache, k):
    """Keys are 32-bit hashes of the first K pattern bytes: two different prefixes with the same
    key share a slot and must be told apart by the byte compare."""
    a, b = _colliding_patterns(k)
    filler = [bytes([0x61 + i % 26]) * k + b"pad%03d" % i for i in range(40)]  # keeps the shortest pattern at k + ...
    pats = [a[:k + 1], a, b, b[:k + 2]] + filler
    pats = [p for p in pats if len(p) >= k]
    shortest = min(len(p) for p in pats)
    pats.append(b"q" * k)  # pins K = k
    path = store_cache(f"collide-{k}", b"\n".join(pats))
    o = Oracle.from_olm(path)
    hay = b"..".join([a, b, a[:k] + b"x", b[:k] + b"y", b + a, a[:k + 1], b"q" * (k + 3)] * 50)
    with Matcher(path) as m:
        got = check(m, o, hay)
        check(m, o, hay, longest_only=True)
        check(m, o, hay, word_boundary=True)
        assert got.size >= 50 * 8 and shortest >= k

