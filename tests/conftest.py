"""pytest configuration: the `gpu` marker and shared helpers.

  python -m pytest tests -x -q -m "not gpu"   # build container, no GPU: oracle, host logic, ABI
  python -m pytest tests -x -q -m gpu         # B200: parity of the CUDA path through the C ABI
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (runs the CUDA path through the C ABI)")


def same_matches(a: np.ndarray, b: np.ndarray) -> bool:
    return a.size == b.size and bool((a["offset"] == b["offset"]).all()) and bool((a["len"] == b["len"]).all())


def describe_diff(got: np.ndarray, want: np.ndarray) -> str:
    a = set(zip(got["offset"].tolist(), got["len"].tolist()))
    b = set(zip(want["offset"].tolist(), want["len"].tolist()))
    return (f"got {got.size} want {want.size}; only in product {sorted(a - b)[:6]}; "
            f"only in oracle {sorted(b - a)[:6]}; same set, different order: {a == b}")


@pytest.fixture(scope="session")
def product_lib():
    """The shipped shared library; building is the job of __graft_entry__.build()."""
    from omega_match_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def store_cache(tmp_path_factory, product_lib):
    """Compiles pattern sets with the PRODUCT compiler once per (set, flags)."""
    from omega_match_b200 import Compiler
    d = tmp_path_factory.mktemp("stores")
    cache = {}

    def get(key: str, pattern_buf: bytes, store_flags=(0, 0, 0)) -> str:
        k = (key, tuple(store_flags))
        if k not in cache:
            path = str(d / f"{len(cache)}.olm")
            Compiler.compile_from_buffer(path, pattern_buf, *map(bool, store_flags))
            cache[k] = path
        return cache[k]

    return get
