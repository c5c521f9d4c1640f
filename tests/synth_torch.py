"""Device-side (torch) twins of tests/inputs.py generators, for multi-GB synthetic haystacks.

Bench/test plumbing only: the formulas are SURVEY.md 8d's (counter based, so any slice can be
generated on any device); tests/test_gpu_parity.py checks these against the numpy versions.
"""
from __future__ import annotations

import numpy as np
import torch

_GAMMA = 0x9E3779B97F4A7C15 - (1 << 64)
_C1 = 0xBF58476D1CE4E5B9 - (1 << 64)
_C2 = 0x94D049BB133111EB - (1 << 64)


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    return (z >> k) & ((1 << (64 - k)) - 1)


def splitmix64_t(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 on int64 tensors holding uint64 bit patterns."""
    z = x + _GAMMA
    z = (z ^ _lsr(z, 30)) * _C1
    z = (z ^ _lsr(z, 27)) * _C2
    return z ^ _lsr(z, 31)


def _umod(u: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
    """(uint64 u) % m for int64 bit patterns, m < 2**31."""
    hi = _lsr(u, 1) % m
    return (hi * 2 + (u & 1)) % m


def _to_i64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def synth_haystack_torch(n: int, seed: int, start: int = 0, device="cuda", out: torch.Tensor | None = None,
                         chunk_words: int = 1 << 24) -> torch.Tensor:
    """inputs.synth_haystack on the device: uint8 tensor of the global bytes [start, start+n).
    `start` must be a multiple of 8."""
    assert start % 8 == 0
    if out is None:
        out = torch.empty(((n + 7) // 8) * 8, dtype=torch.uint8, device=device)
    first = start >> 3
    total_words = (n + 7) >> 3
    for w0 in range(0, total_words, chunk_words):
        w1 = min(total_words, w0 + chunk_words)
        idx = torch.arange(first + w0, first + w1, dtype=torch.int64, device=device) + _to_i64(seed)
        v = splitmix64_t(idx).view(torch.uint8) & 63  # little endian: byte k = bits 8k..8k+7
        # ALPHABET64: a-z, A-Z, 8 x ' ', '\n', '.', ',', '-'
        b = torch.where(v < 26, v + 97, torch.where(v < 52, v + (65 - 26), torch.full_like(v, 32)))
        b = torch.where(v == 60, torch.full_like(v, 10), b)
        b = torch.where(v == 61, torch.full_like(v, 46), b)
        b = torch.where(v == 62, torch.full_like(v, 44), b)
        b = torch.where(v == 63, torch.full_like(v, 45), b)
        out[w0 * 8:w1 * 8] = b
    return out[:n]


def pack_patterns(patterns: list[bytes], device="cuda"):
    width = max(len(p) for p in patterns)
    arr = np.zeros((len(patterns), width), dtype=np.uint8)
    lens = np.zeros(len(patterns), dtype=np.int64)
    for i, p in enumerate(patterns):
        arr[i, :len(p)] = np.frombuffer(p, dtype=np.uint8)
        lens[i] = len(p)
    return torch.from_numpy(arr).to(device), torch.from_numpy(lens).to(device)


def plant_torch(hay: torch.Tensor, pat_bytes: torch.Tensor, pat_lens: torch.Tensor, seed: int, start: int = 0,
                block: int = 4096, chunk_blocks: int = 1 << 18) -> int:
    """inputs.plant on the device (in place).  Returns the number of planted patterns."""
    n = hay.numel()
    b0 = (start + block - 1) // block
    b1 = (start + n) // block
    if b1 <= b0:
        return 0
    dev = hay.device
    npat, width = pat_bytes.shape
    cols = torch.arange(-1, width + 1, dtype=torch.int64, device=dev)  # -1 = leading space
    for c0 in range(b0, b1, chunk_blocks):
        c1 = min(b1, c0 + chunk_blocks)
        blocks = torch.arange(c0, c1, dtype=torch.int64, device=dev)
        r = splitmix64_t(blocks + _to_i64(seed))
        r2 = splitmix64_t(blocks + _to_i64(seed ^ 0x5555))
        pi = _umod(r2, torch.tensor(npat, dtype=torch.int64, device=dev))
        ln = pat_lens[pi]
        room = torch.clamp(block - 2 - ln, min=1)
        at = blocks * block - start + 1 + _umod(r, room)
        pos = at[:, None] + cols[None, :]
        vals = torch.full((c1 - c0, width + 2), 32, dtype=torch.uint8, device=dev)
        vals[:, 1:width + 1] = pat_bytes[pi]
        j = cols[None, :]
        is_pat = (j >= 0) & (j < ln[:, None])
        vals = torch.where(is_pat, vals, torch.full_like(vals, 32))
        mask = j <= ln[:, None]
        hay[pos[mask]] = vals[mask]
    return b1 - b0
