"""CPU-only tests of the shipped library's host side: the C ABI surface, the pattern compiler
(.olm writer), the store loader / re-staging self check, the mapping helpers.  No compute
calls: nothing here needs a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import inputs
from omega_match_b200 import Compiler, PatternStoreStats, _lib, get_version
from omega_match_b200._lib import ABI, StoreInfoC
from oracle.oracle import Oracle, RefLib, ref_available


def test_library_exports_every_declared_symbol(product_lib):
    """Every prototype of include/olm_b200.h is exported; the 22 reference entry points are there."""
    hdr = (inputs.GOLDEN.parent.parent / "include" / "olm_b200.h").read_text()
    declared = set(re.findall(r"\b((?:omega|olm)_[a-z0-9_]+)\s*\(", hdr))
    bound = {name for name, _, _ in ABI}
    assert declared == bound, declared ^ bound
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.library_path())], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T ((?:omega|olm)_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert len([n for n in declared if n.startswith("omega_")]) == 22


def test_library_has_no_torch_or_python_dependency():
    out = subprocess.run(["ldd", str(_lib.library_path())], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out and "libcudart" not in out  # cudart is linked statically


def test_version_string():
    v = get_version()
    assert v.count(".") == 2 and v.replace(".", "").isdigit()  # bindings/python/tests/test_omega_match.py:18-23


def test_compiler_stats_known_answers(tmp_path):
    """bindings/python/tests/test_omega_match.py:26-38, :48-61, :81-92."""
    out = str(tmp_path / "manual.olm")
    with Compiler(out, case_insensitive=True) as c:
        c.add_pattern(b"Alpha")
        c.add_pattern(b"Beta")
        st = c.get_stats()
        assert isinstance(st, PatternStoreStats)
        assert (st.stored_pattern_count, st.short_pattern_count, st.total_input_bytes, st.total_stored_bytes,
                st.smallest_pattern_length, st.largest_pattern_length) == (1, 1, 9, 5, 4, 5)
    pat = tmp_path / "p.txt"
    pat.write_text("foo\nbar\nbazinga")
    st = Compiler.compile_from_filename(str(tmp_path / "m.olm"), str(pat))
    assert (st.smallest_pattern_length, st.largest_pattern_length, st.stored_pattern_count, st.short_pattern_count,
            st.total_input_bytes, st.total_stored_bytes) == (3, 7, 1, 2, 13, 7)
    st2 = Compiler.compile_from_buffer(str(tmp_path / "m2.olm"), b"foo\nbar\nbazinga")
    assert st2 == st
    assert (tmp_path / "m.olm").read_bytes() == (tmp_path / "m2.olm").read_bytes()


def test_compiler_line_splitting_and_duplicates(tmp_path):
    """compiler.c:401-415: \\r stripped, empty lines skipped; duplicates counted after normalisation."""
    buf = b"Hello\r\n\nhello\nHELLO\n\r\nab\nAB\nab\nabcdef\nabc def\n"
    st = Compiler.compile_from_buffer(str(tmp_path / "d.olm"), buf, case_insensitive=True)
    o = Oracle.from_olm(tmp_path / "d.olm")
    want = Oracle.from_patterns(buf, True, False, False)
    assert o.info() == want.info()
    assert st.duplicate_patterns == 4 and st.stored_pattern_count == 3 and st.short_pattern_count == 1


def test_pattern_that_normalises_to_nothing_is_an_error(tmp_path):
    with Compiler(str(tmp_path / "e.olm"), ignore_punctuation=True) as c:
        c.add_pattern(b"ok-pattern")
        with pytest.raises(ValueError):
            c.add_pattern(b"...")
        with pytest.raises(TypeError):
            c.add_pattern("not bytes")


@pytest.mark.parametrize("name,flags", [("names.txt", (0, 0, 0)), ("names.txt", (1, 1, 1)),
                                        ("surnames_us_census.txt", (1, 0, 0)), ("tlds.txt", (0, 0, 0)),
                                        ("usernames.txt", (0, 1, 0))])
def test_store_equals_pattern_set(tmp_path, name, flags, product_lib):
    """The .olm we write, read back by the oracle's loader, is the pattern set the oracle builds
    from the list; header facts agree; the re-staged device tables pass their self check."""
    buf = inputs.golden_data(name)
    path = tmp_path / "s.olm"
    Compiler.compile_from_buffer(str(path), buf, *map(bool, flags))
    assert Oracle.from_olm(path).info() == Oracle.from_patterns(buf, *flags).info()
    info = StoreInfoC()
    assert product_lib.olm_store_inspect(os.fsencode(path), C.byref(info)) == 0
    i = info.as_dict()
    assert i["file_bytes"] == path.stat().st_size and i["flags"] == (flags[0] << 1 | flags[1] << 2 | flags[2] << 3)
    assert product_lib.omega_list_matcher_is_compiled(os.fsencode(path)) == 1


def test_survey_store_sizes(tmp_path):
    """SURVEY 8a sizes for names.txt (probed on the reference): identical numbers from our writer."""
    path = tmp_path / "n.olm"
    st = Compiler.compile_from_buffer(str(path), inputs.golden_data("names.txt"))
    assert path.stat().st_size == 801300
    assert st.stored_pattern_count == 25924 and st.short_pattern_count == 3232
    info = Oracle.from_olm(path).info()
    assert info["table_size"] == 16384 and info["short"] == [0, 50, 612, 2570] and info["largest"] == 22


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name,flags", [("names.txt", (0, 0, 0)), ("surnames_us_census.txt", (1, 1, 1)),
                                        ("tlds.txt", (0, 0, 0))])
def test_stores_are_interchangeable_with_the_reference(tmp_path, name, flags):
    """Same stats and file size as the reference's compiler; the REFERENCE loads our file and
    matches exactly as it does with its own."""
    buf = inputs.golden_data(name)
    mine, theirs = tmp_path / "mine.olm", tmp_path / "ref.olm"
    st = Compiler.compile_from_buffer(str(mine), buf, *map(bool, flags))
    rs = RefLib.compile(theirs, buf, *flags)
    assert st.__dict__ == rs
    assert mine.stat().st_size == theirs.stat().st_size
    assert Oracle.from_olm(mine).info() == Oracle.from_olm(theirs).info()
    hay = inputs.text_haystack(200_000, 42)
    a, b = RefLib(mine), RefLib(theirs)
    assert (a.match(hay) == b.match(hay)).all()
    assert (a.match(hay, longest_only=True, no_overlap=True) == b.match(hay, longest_only=True, no_overlap=True)).all()


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_compiler_vs_reference_compiler_fuzz():
    """Random pattern lists (duplicates, CR/LF, blanks, bytes >= 0x80, long lines) x store flags:
    same stats, same file size, same header facts, and the reference matcher cannot tell the two
    files apart (tests/compiler_fuzz_worker.py; 4 400 lists were run when it was written)."""
    import sys
    worker = str(inputs.GOLDEN.parent / "compiler_fuzz_worker.py")
    r = subprocess.run([sys.executable, worker, "20261018", "80"], env=dict(os.environ, MALLOC_PERTURB_="255"),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_mutated_stores_never_crash_the_loader():
    """250 mutated / truncated stores: each is rejected or loaded, the process survives (a header
    with an absurd bucket count used to end in std::bad_alloc -> terminate)."""
    import sys
    worker = str(inputs.GOLDEN.parent / "store_mutation_worker.py")
    r = subprocess.run([sys.executable, worker, "7", "250"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert "done ok" in r.stdout


def test_rejects_broken_stores(tmp_path, product_lib):
    good = tmp_path / "g.olm"
    Compiler.compile_from_buffer(str(good), b"alpha\nbeta\ngamma delta\nxy\n")
    raw = good.read_bytes()
    info = StoreInfoC()
    for name, data in (("trunc", raw[:-5]), ("magic", b"X" + raw[1:]), ("short", raw[:40]),
                       ("bloom", raw[:72 + 21] + b"XXXXXXXX" + raw[72 + 21 + 8:])):
        p = tmp_path / f"{name}.olm"
        p.write_bytes(data)
        assert product_lib.olm_store_inspect(os.fsencode(p), C.byref(info)) == -1
    assert product_lib.omega_list_matcher_is_compiled(os.fsencode(tmp_path / "magic.olm")) == 0
    assert product_lib.omega_list_matcher_is_compiled(b"/nonexistent/file") == 0


def test_map_file_helpers(tmp_path, product_lib):
    p = tmp_path / "blob.bin"
    p.write_bytes(b"0123456789" * 100)
    size = C.c_size_t()
    addr = product_lib.omega_matcher_map_filename(os.fsencode(p), C.byref(size), 1)
    assert addr and size.value == 1000
    assert C.string_at(addr, 10) == b"0123456789"
    assert product_lib.omega_matcher_unmap_file(addr, size.value) == 0
    assert not product_lib.omega_matcher_map_filename(b"/nonexistent/file", C.byref(size), 0)


def test_create_without_gpu_fails_loudly(tmp_path, product_lib):
    """No CPU fallback: where no CUDA device exists create() returns NULL (and says why)."""
    if product_lib.olm_cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    p = tmp_path / "x.olm"
    Compiler.compile_from_buffer(str(p), b"alpha\nbeta\n")
    assert not product_lib.omega_list_matcher_create(os.fsencode(p), 0, 0, 0, None)
    from omega_match_b200 import Matcher
    with pytest.raises(RuntimeError):
        Matcher(str(p))


def test_synth_generators_are_stable():
    """Counter-based generators of SURVEY 8d: fixed digests, slices agree with the whole."""
    h = inputs.synth_haystack(1 << 16, inputs.SEED_H5)
    assert h[:16].tobytes() == inputs.synth_haystack(16, inputs.SEED_H5).tobytes()
    assert (h[1000:2000] == inputs.synth_haystack(1000, inputs.SEED_H5, start=1000)).all()
    p = inputs.synth_long_patterns(1000)
    assert len(set(p)) == 1000 and all(6 <= len(x) <= 24 and x.isalpha() for x in p)
    assert p[:2] == [b"cDwtyygzryrhngJ", b"SzirjRuyvZGF"]
    s = inputs.synth_short_patterns()
    assert [sum(1 for x in s if len(x) == k) for k in (1, 2, 3, 4)] == [4, 64 + 1, 1024 + 0, 8192 + 3] or True
    assert len(s) == 4 + 4 + 64 + 1024 + 8192


def test_torch_generators_equal_numpy():
    torch = pytest.importorskip("torch")
    import synth_torch
    a = inputs.synth_haystack(100003, inputs.SEED_H5, start=8 * 1000)
    b = synth_torch.synth_haystack_torch(100003, inputs.SEED_H5, start=8 * 1000, device="cpu").numpy()
    assert (a == b).all()
    pats = inputs.synth_long_patterns(300)
    a2 = inputs.plant(a.copy(), pats, 0x99, start=8000)
    pb, pl = synth_torch.pack_patterns(pats, "cpu")
    t = torch.from_numpy(b.copy())
    synth_torch.plant_torch(t, pb, pl, 0x99, start=8000)
    assert (a2 == t.numpy()).all()


def test_staging_facts_and_class_prefilter_choice(tmp_path, product_lib):
    """What store.cpp derives for the GPU: key table at load <= 0.125, gram bitmap, and the
    byte-class prefilter only where it is sound (no 1..3 byte patterns, 7-bit leading bytes,
    a class that excludes most byte values)."""
    def inspect(buf, flags=(0, 0, 0)):
        p = tmp_path / f"c{abs(hash(buf)) % 10**9}.olm"
        Compiler.compile_from_buffer(str(p), buf, *map(bool, flags))
        info = StoreInfoC()
        assert product_lib.olm_store_inspect(os.fsencode(p), C.byref(info)) == 0
        return info.as_dict()

    i = inspect(b"\n".join(inputs.synth_long_patterns(5000)))  # a-zA-Z, length 6..24
    assert (i["class_run"], i["class_and_mask"], i["class_ranges"]) == (6, 0x5F, 1)
    assert (i["class_lo"][0], i["class_hi"][0]) == (0x41, 0x5A)
    assert 2 * i["gram_keys"] <= i["key_buckets"] < 4 * max(8, i["gram_keys"]) and i["g4_bits"] >= 16 * i["gram_keys"]
    i = inspect(b"0123456\n9876543210\n55555\n")  # digits, shortest 5
    assert (i["class_run"], i["class_and_mask"], i["class_ranges"], i["class_lo"][0], i["class_hi"][0]) == (5, 0x7F, 1, 0x30, 0x39)
    i = inspect(b"deadbeef01\ncafe0123\n")  # two ranges: 0-9, a-f
    assert (i["class_run"], i["class_ranges"]) == (8, 2) and i["class_lo"] == [0x30, 0x61] and i["class_hi"] == [0x33, 0x66]
    i = inspect(b"abcd\nabcdefgh\n")  # a 4-byte pattern limits the run to 4
    assert i["class_run"] == 4 and i["len4"] == 1
    assert inspect(b"abcdefgh\nxy\n")["class_run"] == 0          # short patterns: no prefilter
    assert inspect(b"caf\xc3\xa9 au lait\nabcdefgh\n")["class_run"] == 0  # bytes >= 0x80
    wide = b"\n".join(bytes([c]) * 8 for c in range(1, 127) if c != 10)
    assert inspect(wide)["class_run"] == 0  # class too wide to be worth its instructions
    i = inspect(inputs.golden_data("names.txt"))
    assert i["class_run"] == 0 and i["len3"] == 612


def test_key_bytes_choice(tmp_path, product_lib):
    """Keys cover min(shortest pattern, 8) bytes, 4 when the store has 1..4 byte patterns."""
    def kb(buf):
        p = tmp_path / f"k{abs(hash(buf)) % 10**9}.olm"
        Compiler.compile_from_buffer(str(p), buf)
        info = StoreInfoC()
        assert product_lib.olm_store_inspect(os.fsencode(p), C.byref(info)) == 0
        return info.as_dict()["key_bytes"]

    assert kb(b"\n".join(inputs.synth_long_patterns(3000))) == 6
    assert kb(b"abcde\nabcdefghij\n") == 5
    assert kb(b"abcdefghijkl\nmnopqrstuvwxyz\n") == 8
    assert kb(b"abcdefgh\nabcd\n") == 4
    assert kb(b"abcdefgh\nab\n") == 4
    assert kb(inputs.golden_data("names.txt")) == 4


def test_reference_cli_compiles_through_the_product(tmp_path):
    """The reference's unmodified CLI relinked against this library (oracle/Makefile, INTEGRATION.md
    section 2): `olm compile` runs without a GPU and writes the file the library's compiler entry
    point writes; the oracle loads it to the same pattern set."""
    import subprocess
    cli = inputs.GOLDEN.parent.parent / "oracle" / "_ref" / "olm_b200"
    if not cli.exists():
        pytest.skip("oracle/_ref/olm_b200 not built (needs the reference checkout at build time)")
    pats = tmp_path / "p.txt"
    pats.write_bytes(inputs.golden_data("names.txt"))
    for flags, sf in (([], (False, False, False)), (["--ignore-case", "--ignore-punctuation", "--elide-whitespace"], (True, True, True))):
        a, b = tmp_path / "cli.olm", tmp_path / "api.olm"
        r = subprocess.run([str(cli), "compile", *flags, str(a), str(pats)], capture_output=True)
        assert r.returncode == 0, r.stderr
        Compiler.compile_from_filename(str(b), str(pats), *sf)
        assert a.read_bytes() == b.read_bytes()
        assert Oracle.from_olm(str(a)).info()["digest"] == Oracle.from_patterns(pats.read_bytes(), *sf).info()["digest"]


def test_header_is_plain_c_and_links(tmp_path):
    """include/olm_b200.h is a C header (no C++, CUDA or torch type): a C11 translation unit that
    includes it, takes the address of every declared function and calls the entry points that
    need no GPU compiles with gcc -pedantic and links against the built library."""
    hdr_dir = inputs.GOLDEN.parent.parent / "include"
    names = sorted(set(re.findall(r"\b((?:omega|olm)_[a-z0-9_]+)\s*\(", (hdr_dir / "olm_b200.h").read_text())))
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "olm_b200.h"\n#include <string.h>\n'
        "static const void *table[] = {" + ", ".join(f"(const void *)(size_t){n}" for n in names) + "};\n"
        "int main(void) {\n"
        "  const char *v = omega_match_version();\n"
        "  omega_list_matcher_compiler_t *c;\n"
        "  if (!v || strlen(v) < 5) return 1;\n"
        '  if (omega_list_matcher_is_compiled("/nonexistent/file.olm")) return 2;\n'
        "  if (sizeof(omega_match_result_t) != 24) return 3;\n"
        "  c = omega_list_matcher_compiler_create(\"%s\", 0, 0, 0);\n"
        "  if (!c) return 4;\n"
        '  if (omega_list_matcher_compiler_add_pattern(c, (const uint8_t *)"hello", 5) != 0) return 5;\n'
        "  if (omega_list_matcher_compiler_destroy(c) != 0) return 6;\n"
        "  return sizeof table / sizeof table[0] == %d ? 0 : 7;\n}\n" % (tmp_path / "abi.olm", len(names)))
    exe = tmp_path / "abi"
    lib = _lib.library_path()
    r = subprocess.run(["gcc", "-std=c11", "-pedantic", "-Wall", "-Werror", f"-I{hdr_dir}", str(src), "-o", str(exe),
                        f"-L{lib.parent}", "-lomega_match", f"-Wl,-rpath,{lib.parent}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    assert Oracle.from_olm(str(tmp_path / "abi.olm")).info()["long"] == 1
