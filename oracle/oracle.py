"""ctypes binding of the CPU oracle (oracle/gvdb_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.

PARITY STATUS: parity unpinned by the reference's own suite (see the header of
gvdb_oracle.cpp); pinned to hand-derived known answers in tests/test_oracle_kat.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgvdb_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gvdb_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libgvdb_oracle.so"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, f32, i64, u64 = C.c_void_p, C.c_size_t, C.c_float, C.c_int64, C.c_uint64
        L.gvo_quantize.argtypes = [vp, sz, f32, vp]
        L.gvo_quantize.restype = None
        L.gvo_quantize_batch.argtypes = [vp, sz, sz, f32, vp]
        L.gvo_quantize_batch.restype = None
        L.gvo_hamming.argtypes = [vp, vp, sz]
        L.gvo_hamming.restype = u64
        L.gvo_similarity.argtypes = [vp, vp, sz]
        L.gvo_similarity.restype = f32
        L.gvo_cosine_similarity.argtypes = [vp, vp, sz]
        L.gvo_cosine_similarity.restype = f32
        L.gvo_cosine_distance.argtypes = [vp, vp, sz]
        L.gvo_cosine_distance.restype = f32
        L.gvo_rescore_count.argtypes = [sz, f32]
        L.gvo_rescore_count.restype = sz
        L.gvo_multi_stage_search.argtypes = [vp, vp, vp, vp, sz, sz, sz, vp, vp, vp, vp]
        L.gvo_multi_stage_search.restype = i64
        L.gvo_multi_stage_search_select.argtypes = [vp, vp, vp, vp, sz, sz, sz, vp, vp]
        L.gvo_multi_stage_search_select.restype = i64
        L.gvo_flat_search.argtypes = [vp, vp, vp, sz, sz, sz, vp, vp]
        L.gvo_flat_search.restype = i64
        L.gvo_similarity_search.argtypes = [vp, vp, vp, sz, sz, sz, f32, C.c_int, vp, vp]
        L.gvo_similarity_search.restype = i64
        L.gvo_multi_stage_search_batch.argtypes = [vp, sz, vp, vp, sz, sz, f32, sz, sz, vp, vp,
                                                   C.c_int, C.c_int]
        L.gvo_multi_stage_search_batch.restype = i64
        L.gvo_flat_search_batch.argtypes = [vp, sz, vp, vp, sz, sz, sz, vp, vp, C.c_int]
        L.gvo_flat_search_batch.restype = i64
        L.gvo_shard_merge.argtypes = [vp, vp, vp, sz, sz, sz, vp, vp]
        L.gvo_shard_merge.restype = i64
        L.gvo_concat_sort_truncate.argtypes = [vp, vp, sz, sz, vp, vp]
        L.gvo_concat_sort_truncate.restype = i64
        L.gvo_rrf_fusion.argtypes = [vp, sz, vp, sz, vp, sz, f32, vp, vp, sz]
        L.gvo_rrf_fusion.restype = i64
        L.gvo_weighted_fusion.argtypes = [vp, vp, sz, vp, vp, sz, vp, vp, sz, f32, f32, f32, C.c_int, vp, vp, sz]
        L.gvo_weighted_fusion.restype = i64
        L.gvo_bm25_avg_len.argtypes = [vp, vp, vp, sz, sz]
        L.gvo_bm25_avg_len.restype = f32
        L.gvo_bm25_search.argtypes = [vp, vp, sz, vp, vp, vp, vp, sz, sz, f32, f32, f32, sz, vp, vp]
        L.gvo_bm25_search.restype = i64
        L.gvo_hardware_threads.argtypes = []
        L.gvo_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def hardware_threads() -> int:
    return int(lib().gvo_hardware_threads())


def quantize(x, threshold: float = 0.0) -> np.ndarray:
    """BinaryQuantizer::quantize -> the bytes of BinaryVector::to_bytes()."""
    x = _f32(x)
    assert x.ndim == 1
    code = np.zeros((x.shape[0] + 7) // 8, dtype=np.uint8)
    lib().gvo_quantize(_p(x), x.shape[0], threshold, _p(code))
    return code


def quantize_batch(x, threshold: float = 0.0) -> np.ndarray:
    x = _f32(x)
    assert x.ndim == 2
    n, dim = x.shape
    codes = np.zeros((n, (dim + 7) // 8), dtype=np.uint8)
    if n:
        lib().gvo_quantize_batch(_p(x), n, dim, threshold, _p(codes))
    return codes


def hamming(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    assert a.shape == b.shape
    return int(lib().gvo_hamming(_p(a), _p(b), a.size))


def similarity(a: np.ndarray, b: np.ndarray, dim: int) -> np.float32:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    return np.float32(lib().gvo_similarity(_p(a), _p(b), dim))


def cosine_similarity(a, b) -> np.float32:
    a, b = _f32(a), _f32(b)
    return np.float32(lib().gvo_cosine_similarity(_p(a), _p(b), a.shape[0]))


def cosine_distance(a, b) -> np.float32:
    a, b = _f32(a), _f32(b)
    return np.float32(lib().gvo_cosine_distance(_p(a), _p(b), a.shape[0]))


def rescore_count(n: int, ratio: float) -> int:
    return int(lib().gvo_rescore_count(n, ratio))


def hamming_all(qcode: np.ndarray, codes: np.ndarray) -> np.ndarray:
    """All N Hamming distances (numpy popcount; integer, exact)."""
    x = np.bitwise_xor(codes, qcode[None, :])
    return np.unpackbits(x, axis=1).sum(axis=1, dtype=np.uint32)


def multi_stage_search(q, rows, rescore_count_: int, threshold: float = 0.0, codes=None,
                       select: bool = False, want_candidates: bool = False):
    """BinaryQuantizer::multi_stage_search; returns (idx[R], score[R][, cand_idx, cand_ham])."""
    q, rows = _f32(q), _f32(rows)
    n, dim = rows.shape
    if codes is None:
        codes = quantize_batch(rows, threshold)
    qcode = quantize(q, threshold)
    r = min(rescore_count_, n)
    idx = np.full(max(r, 1), np.iinfo(np.uint64).max, dtype=np.uint64)
    sc = np.zeros(max(r, 1), dtype=np.float32)
    ci = np.zeros(max(r, 1), dtype=np.uint64)
    ch = np.zeros(max(r, 1), dtype=np.uint32)
    if select:
        got = lib().gvo_multi_stage_search_select(_p(qcode), _p(codes), _p(q), _p(rows), n, dim, r,
                                                  _p(idx), _p(sc))
    else:
        got = lib().gvo_multi_stage_search(_p(qcode), _p(codes), _p(q), _p(rows), n, dim, r,
                                           _p(idx), _p(sc), _p(ci), _p(ch))
    if got < 0:
        raise FloatingPointError("NaN reached partial_cmp().unwrap() (reference panics)")
    if want_candidates:
        return idx[:got], sc[:got], ci[:got], ch[:got]
    return idx[:got], sc[:got]


def multi_stage_search_batch(queries, rows, rescore_count_: int, k: int, threshold: float = 0.0,
                             codes=None, nthreads: int = 1, select: bool = False):
    queries, rows = _f32(queries), _f32(rows)
    nq, dim = queries.shape
    n = rows.shape[0]
    if codes is None:
        codes = quantize_batch(rows, threshold)
    idx = np.full((nq, k), np.iinfo(np.uint64).max, dtype=np.uint64)
    sc = np.full((nq, k), -np.inf, dtype=np.float32)
    got = lib().gvo_multi_stage_search_batch(_p(queries), nq, _p(codes), _p(rows), n, dim,
                                             threshold, rescore_count_, k, _p(idx), _p(sc),
                                             nthreads, 1 if select else 0)
    if got < 0:
        raise FloatingPointError("NaN in scores")
    return idx, sc


def flat_search(q, rows, k: int, live=None):
    """FaissVectorIndex::search: (row idx[k'], cosine DISTANCE[k']) ascending."""
    q, rows = _f32(q), _f32(rows)
    n, dim = rows.shape
    lv = None if live is None else np.ascontiguousarray(live, dtype=np.uint8)
    idx = np.zeros(max(k, 1), dtype=np.uint64)
    ds = np.zeros(max(k, 1), dtype=np.float32)
    got = lib().gvo_flat_search(_p(q), _p(rows), _p(lv), n, dim, k, _p(idx), _p(ds))
    if got < 0:
        raise FloatingPointError("NaN distance")
    return idx[:got], ds[:got]


def similarity_search(q, rows, limit: int, threshold=None, live=None):
    """BasicVectorStore::vector_search: (row idx[k'], cosine SIMILARITY[k']) descending."""
    q, rows = _f32(q), _f32(rows)
    n, dim = rows.shape
    lv = None if live is None else np.ascontiguousarray(live, dtype=np.uint8)
    idx = np.zeros(max(limit, 1), dtype=np.uint64)
    sm = np.zeros(max(limit, 1), dtype=np.float32)
    got = lib().gvo_similarity_search(_p(q), _p(rows), _p(lv), n, dim, limit,
                                      0.0 if threshold is None else float(threshold),
                                      0 if threshold is None else 1, _p(idx), _p(sm))
    if got < 0:
        raise FloatingPointError("NaN similarity")
    return idx[:got], sm[:got]


def flat_search_batch(queries, rows, k: int, live=None, nthreads: int = 1):
    queries, rows = _f32(queries), _f32(rows)
    nq, dim = queries.shape
    n = rows.shape[0]
    lv = None if live is None else np.ascontiguousarray(live, dtype=np.uint8)
    idx = np.full((nq, k), np.iinfo(np.uint64).max, dtype=np.uint64)
    ds = np.full((nq, k), np.inf, dtype=np.float32)
    got = lib().gvo_flat_search_batch(_p(queries), nq, _p(rows), _p(lv), n, dim, k, _p(idx),
                                      _p(ds), nthreads)
    if got < 0:
        raise FloatingPointError("NaN distance")
    return idx, ds


def shard_merge(ham, idx, score, rescore_count_: int, k: int):
    ham = np.ascontiguousarray(ham, dtype=np.uint32).ravel()
    idx = np.ascontiguousarray(idx, dtype=np.uint64).ravel()
    score = _f32(score).ravel()
    oi = np.zeros(max(k, 1), dtype=np.uint64)
    os_ = np.zeros(max(k, 1), dtype=np.float32)
    got = lib().gvo_shard_merge(_p(ham), _p(idx), _p(score), ham.size, rescore_count_, k, _p(oi),
                                _p(os_))
    return oi[:got], os_[:got]


def concat_sort_truncate(idx, score, limit: int):
    idx = np.ascontiguousarray(idx, dtype=np.uint64).ravel()
    score = _f32(score).ravel()
    oi = np.zeros(max(limit, 1), dtype=np.uint64)
    os_ = np.zeros(max(limit, 1), dtype=np.float32)
    got = lib().gvo_concat_sort_truncate(_p(idx), _p(score), idx.size, limit, _p(oi), _p(os_))
    return oi[:got], os_[:got]


def rrf_fusion(dense, sparse, text, k: float = 60.0):
    d = np.ascontiguousarray(dense, dtype=np.uint64)
    s = np.ascontiguousarray(sparse, dtype=np.uint64)
    t = np.ascontiguousarray(text, dtype=np.uint64)
    cap = d.size + s.size + t.size
    oi = np.zeros(max(cap, 1), dtype=np.uint64)
    os_ = np.zeros(max(cap, 1), dtype=np.float32)
    got = lib().gvo_rrf_fusion(_p(d), d.size, _p(s), s.size, _p(t), t.size, k, _p(oi), _p(os_), cap)
    return oi[:got], os_[:got]


def weighted_fusion(dense, dense_sc, sparse, sparse_sc, text, text_sc, weights=(0.7, 0.2, 0.1), normalize: bool = False):
    """linear_fusion (normalize=False) / normalized_fusion (normalize=True), src/hybrid.rs:491-616."""
    d = np.ascontiguousarray(dense, dtype=np.uint64); ds = np.ascontiguousarray(dense_sc, dtype=np.float32)
    s = np.ascontiguousarray(sparse, dtype=np.uint64); ss = np.ascontiguousarray(sparse_sc, dtype=np.float32)
    t = np.ascontiguousarray(text, dtype=np.uint64); ts = np.ascontiguousarray(text_sc, dtype=np.float32)
    assert d.size == ds.size and s.size == ss.size and t.size == ts.size
    cap = d.size + s.size + t.size
    oi = np.zeros(max(cap, 1), dtype=np.uint64)
    os_ = np.zeros(max(cap, 1), dtype=np.float32)
    got = lib().gvo_weighted_fusion(_p(d), _p(ds), d.size, _p(s), _p(ss), s.size, _p(t), _p(ts), t.size,
                                    weights[0], weights[1], weights[2], int(normalize), _p(oi), _p(os_), cap)
    return oi[:got], os_[:got]


def bm25_avg_len(post_off, post_doc, doc_len) -> np.float32:
    post_off = np.ascontiguousarray(post_off, dtype=np.uint64)
    post_doc = np.ascontiguousarray(post_doc, dtype=np.uint32)
    doc_len = _f32(doc_len)
    return np.float32(lib().gvo_bm25_avg_len(_p(post_off), _p(post_doc), _p(doc_len),
                                             post_off.size - 1, doc_len.size))


def bm25_search(q_terms, q_tf, post_off, post_doc, post_tf, doc_len, limit: int, avg_len=None,
                k1: float = 1.2, b: float = 0.75):
    q_terms = np.ascontiguousarray(q_terms, dtype=np.uint32)
    q_tf = _f32(q_tf)
    post_off = np.ascontiguousarray(post_off, dtype=np.uint64)
    post_doc = np.ascontiguousarray(post_doc, dtype=np.uint32)
    post_tf = _f32(post_tf)
    doc_len = _f32(doc_len)
    if avg_len is None:
        avg_len = bm25_avg_len(post_off, post_doc, doc_len)
    od = np.zeros(max(limit, 1), dtype=np.uint64)
    os_ = np.zeros(max(limit, 1), dtype=np.float32)
    got = lib().gvo_bm25_search(_p(q_terms), _p(q_tf), q_terms.size, _p(post_off), _p(post_doc),
                                _p(post_tf), _p(doc_len), post_off.size - 1, doc_len.size,
                                float(avg_len), k1, b, limit, _p(od), _p(os_))
    return od[:got], os_[:got]
