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


# ---- sparse side of the hybrid config ------------------------------
STREAM_SPARSE_DOC, STREAM_SPARSE_QUERY = 6, 7


def _zipf_terms(seed: int, stream: int, first: int, n: int, vocab: int) -> np.ndarray:
    """n term ids drawn Zipf(1.0) over [0, vocab): inverse CDF on a 32-bit hash uniform."""
    w = 1.0 / np.arange(1, vocab + 1, dtype=np.float64)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    with np.errstate(over="ignore"):
        base = np.uint64((seed * _M1 + stream * _M2) & 0xFFFFFFFFFFFFFFFF)
        h = _mix_np(np.arange(n, dtype=np.uint64) + np.uint64(first) + base)
    u = ((h >> np.uint64(32)).astype(np.float64) + 0.5) / 4294967296.0
    return np.minimum(np.searchsorted(cdf, u, side="left"), vocab - 1).astype(np.uint32)


def sparse_corpus(n_docs: int, vocab: int = 100_000, tokens_per_doc: int = 32, seed: int = 42):
    """Each document: `tokens_per_doc` Zipf term draws, tf = count / tokens (SimpleTokenizer::tokenize,
    src/sparse.rs:288-315), document_length = sum of the tfs in ascending term order (:341).
    Returns CSR postings by term (post_off u64 [vocab+1], post_doc u32, post_tf f32) with documents
    ascending inside each term, and doc_len f32 [n_docs]."""
    t = _zipf_terms(seed, STREAM_SPARSE_DOC, 0, n_docs * tokens_per_doc, vocab).reshape(n_docs, tokens_per_doc)
    t = np.sort(t, axis=1)
    doc = np.repeat(np.arange(n_docs, dtype=np.uint32), tokens_per_doc).reshape(n_docs, tokens_per_doc)
    first = np.ones_like(t, dtype=bool)
    first[:, 1:] = t[:, 1:] != t[:, :-1]
    flat_first = np.flatnonzero(first.ravel())
    counts = np.diff(np.append(flat_first, t.size))
    e_term, e_doc = t.ravel()[flat_first], doc.ravel()[flat_first]
    e_tf = counts.astype(np.float32) / np.float32(tokens_per_doc)
    # document_length = sum of the document's tfs: multiples of 1/tokens_per_doc with every partial sum
    # <= 1, exact in f32 in any order when tokens_per_doc is a power of two
    assert tokens_per_doc & (tokens_per_doc - 1) == 0, "tokens_per_doc must be a power of two (exact f32 sums)"
    start = np.flatnonzero(np.r_[True, e_doc[1:] != e_doc[:-1]])
    doc_len = np.zeros(n_docs, dtype=np.float32)
    doc_len[e_doc[start]] = np.add.reduceat(e_tf, start)
    order = np.argsort(e_term.astype(np.uint64) << np.uint64(32) | e_doc.astype(np.uint64), kind="stable")
    post_doc, post_tf = e_doc[order], e_tf[order]
    post_off = np.zeros(vocab + 1, dtype=np.uint64)
    np.cumsum(np.bincount(e_term, minlength=vocab), out=post_off[1:])
    return post_off, np.ascontiguousarray(post_doc), np.ascontiguousarray(post_tf), doc_len


def sparse_queries(nq: int, vocab: int = 100_000, terms_per_query: int = 4, seed: int = 42):
    """nq queries of `terms_per_query` Zipf term draws, tokenised like a document: distinct terms in
    ascending order with tf = count / terms_per_query.  Returns a list of (term ids u32, tfs f32)."""
    t = _zipf_terms(seed, STREAM_SPARSE_QUERY, 0, nq * terms_per_query, vocab).reshape(nq, terms_per_query)
    out = []
    for row in t:
        ids, cnt = np.unique(row, return_counts=True)
        out.append((ids.astype(np.uint32), cnt.astype(np.float32) / np.float32(terms_per_query)))
    return out
