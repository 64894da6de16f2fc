"""Deterministic, integer-exact synthetic vectors.

Every value is a function of (seed, stream, index) through a splitmix64 finaliser, so any
row can be regenerated anywhere (host numpy, or torch on the GPU) with identical bits and
no shared file.  Two datasets:

  lowrank  x[i][j] = sum_{l<L} z[i][l] * A[l][j]   "embedding-like" (L = 16 by default)
  iid      x[i][j] = centred integer draw          (parity only; 1-bit recall is poor)

z, A and the iid draws are centred sums of four hash bytes (range -510..510); the low-rank
product is exact in int64 / float64 and |x| < 2^24, so float32(x) is exact as well.
"""
from __future__ import annotations

import numpy as np

_M1 = 0x9E3779B97F4A7C15
_M2 = 0xBF58476D1CE4E5B9
_M3 = 0x94D049BB133111EB
STREAM_CORPUS, STREAM_MIX, STREAM_QUERY, STREAM_IID, STREAM_IID_QUERY = 1, 2, 3, 4, 5


# ---- numpy ----------------------------------------------------------------------------
def _mix_np(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(_M1)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M2)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M3)
        return z ^ (z >> np.uint64(31))


def _draw_np(seed: int, stream: int, index: np.ndarray) -> np.ndarray:
    """centred integer draw in [-510, 510] as int64"""
    with np.errstate(over="ignore"):
        base = np.uint64((seed * _M1 + stream * _M2) & 0xFFFFFFFFFFFFFFFF)
        h = _mix_np(index.astype(np.uint64) + base)
    b = (h & np.uint64(0xFF)) + ((h >> np.uint64(8)) & np.uint64(0xFF)) + \
        ((h >> np.uint64(16)) & np.uint64(0xFF)) + ((h >> np.uint64(24)) & np.uint64(0xFF))
    return b.astype(np.int64) - 510


def mixing_matrix(dim: int, latent: int = 16, seed: int = 42) -> np.ndarray:
    idx = np.arange(latent * dim, dtype=np.uint64)
    return _draw_np(seed, STREAM_MIX, idx).reshape(latent, dim)


def lowrank_rows(first: int, n: int, dim: int, latent: int = 16, seed: int = 42,
                 stream: int = STREAM_CORPUS) -> np.ndarray:
    """rows [first, first+n) of the low-rank dataset as float32 [n, dim]"""
    idx = (np.arange(n * latent, dtype=np.uint64) + np.uint64(first * latent))
    z = _draw_np(seed, stream, idx).reshape(n, latent).astype(np.float64)
    a = mixing_matrix(dim, latent, seed).astype(np.float64)
    return (z @ a).astype(np.float32)      # exact: |x| <= 16*510^2 < 2^24


def lowrank_queries(first: int, n: int, dim: int, latent: int = 16, seed: int = 42) -> np.ndarray:
    return lowrank_rows(first, n, dim, latent, seed, STREAM_QUERY)


def iid_rows(first: int, n: int, dim: int, seed: int = 42, stream: int = STREAM_IID) -> np.ndarray:
    idx = np.arange(n * dim, dtype=np.uint64) + np.uint64(first * dim)
    return _draw_np(seed, stream, idx).reshape(n, dim).astype(np.float32)


def iid_queries(first: int, n: int, dim: int, seed: int = 42) -> np.ndarray:
    return iid_rows(first, n, dim, seed, STREAM_IID_QUERY)


# ---- torch (same bits; runs on the GPU for corpora that do not fit on the host) ----------
def _s64(v: int) -> int:
    v &= 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr_t(x, s: int):
    return (x >> s) & ((1 << (64 - s)) - 1)


def _mix_t(x):
    z = x + _s64(_M1)
    z = (z ^ _lsr_t(z, 30)) * _s64(_M2)
    z = (z ^ _lsr_t(z, 27)) * _s64(_M3)
    return z ^ _lsr_t(z, 31)


def _draw_t(seed: int, stream: int, index):
    base = _s64(seed * _M1 + stream * _M2)
    h = _mix_t(index + base)
    b = (h & 0xFF) + (_lsr_t(h, 8) & 0xFF) + (_lsr_t(h, 16) & 0xFF) + (_lsr_t(h, 24) & 0xFF)
    return b - 510


def lowrank_rows_torch(first: int, n: int, dim: int, device, latent: int = 16, seed: int = 42,
                       stream: int = STREAM_CORPUS):
    import torch
    idx = torch.arange(n * latent, dtype=torch.int64, device=device) + first * latent
    z = _draw_t(seed, stream, idx).reshape(n, latent).to(torch.float64)
    aidx = torch.arange(latent * dim, dtype=torch.int64, device=device)
    a = _draw_t(seed, STREAM_MIX, aidx).reshape(latent, dim).to(torch.float64)
    return (z @ a).to(torch.float32).contiguous()


def lowrank_queries_torch(first: int, n: int, dim: int, device, latent: int = 16, seed: int = 42):
    return lowrank_rows_torch(first, n, dim, device, latent, seed, STREAM_QUERY)


def iid_rows_torch(first: int, n: int, dim: int, device, seed: int = 42, stream: int = STREAM_IID):
    import torch
    idx = torch.arange(n * dim, dtype=torch.int64, device=device) + first * dim
    return _draw_t(seed, stream, idx).reshape(n, dim).to(torch.float32).contiguous()
