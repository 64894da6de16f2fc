"""Hybrid dense + sparse search on one B200 (BASELINE configs[3]): the GPU mirror of
HybridSearchEngine::search (/root/reference/src/hybrid.rs:286-356) for a batch of requests.

    dense list   2 * limit documents from the vector index (:295-298) — two-stage (1-bit scan +
                 exact rescoring, R = oversample * 2 * limit) or the exact flat scan
    sparse list  2 * limit documents from BM25 over the postings (:305-308)
    fusion       rrf_fusion, FusionStrategy::RRF { k } (:331-333, :422-488), first `limit` (:336)

All three stages run on the GPU through the C ABI (gvdb_search_batch_device /
gvdb_flat_search_batch_device, gvdb_sparse_search_bm25_batch_device, gvdb_rrf_fusion_batch_device);
only the fused top-`limit` lists travel back to the host.  Documents are dense numbers shared by
the two indexes (row i of the vector index == document i of the postings)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .errors import raise_for_status
from .index import NO_ID, _np, _ptr


def rrf_fusion_batch(dense, sparse, text=None, k: float = 60.0, limit: int | None = None, device: int = 0):
    """rrf_fusion for nq queries with HOST lists (nq x n uint64, NO_ID ends a list early)."""
    lib = _ffi.lib()
    lists = []
    nq = None
    for a in (dense, sparse, text):
        if a is None:
            lists.append(None)
            continue
        a = _np(a, np.uint64)
        a = a.reshape(1, -1) if a.ndim == 1 else a
        nq = a.shape[0] if nq is None else nq
        assert a.shape[0] == nq
        lists.append(a if a.shape[1] else None)
    n = [0 if a is None else a.shape[1] for a in lists]
    limit = sum(n) if limit is None else limit
    ids = np.full((nq, limit), NO_ID, dtype=np.uint64)
    sc = np.full((nq, limit), -np.inf, dtype=np.float32)
    raise_for_status(lib.gvdb_rrf_fusion_batch(device, _ptr(lists[0]), n[0], _ptr(lists[1]), n[1], _ptr(lists[2]), n[2],
                                               nq, k, limit, _ptr(ids), _ptr(sc)), lib)
    return ids, sc


def weighted_fusion_batch(dense, dense_scores, sparse, sparse_scores, text=None, text_scores=None,
                          weights=(0.7, 0.2, 0.1), normalize: bool = False, limit: int | None = None, device: int = 0):
    """linear_fusion / normalized_fusion (src/hybrid.rs:491-616) for nq queries with HOST lists: document numbers
    (nq x n uint64, NO_ID ends a list early) and their scores (nq x n float32)."""
    lib = _ffi.lib()
    lists, scores = [], []
    nq = None
    for a, sc in ((dense, dense_scores), (sparse, sparse_scores), (text, text_scores)):
        if a is None:
            lists.append(None); scores.append(None)
            continue
        a = _np(a, np.uint64)
        a = a.reshape(1, -1) if a.ndim == 1 else a
        sc = _np(sc, np.float32).reshape(a.shape)
        nq = a.shape[0] if nq is None else nq
        assert a.shape[0] == nq
        lists.append(a if a.shape[1] else None); scores.append(sc if a.shape[1] else None)
    n = [0 if a is None else a.shape[1] for a in lists]
    limit = sum(n) if limit is None else limit
    ids = np.full((nq, limit), NO_ID, dtype=np.uint64)
    out = np.full((nq, limit), -np.inf, dtype=np.float32)
    raise_for_status(lib.gvdb_weighted_fusion_batch(
        device, _ptr(lists[0]), _ptr(scores[0]), n[0], _ptr(lists[1]), _ptr(scores[1]), n[1], _ptr(lists[2]), _ptr(scores[2]), n[2],
        nq, weights[0], weights[1], weights[2], int(normalize), limit, _ptr(ids), _ptr(out)), lib)
    return ids, out


class HybridSearcher:
    """dense GpuIndex + GpuSparseIndex + fusion, all on one GPU.
    fusion: "rrf" (FusionStrategy::RRF { k }, the default), "linear" or "normalized" (FusionStrategy::Linear /
    Normalized with `weights` = (dense, sparse, text), src/hybrid.rs:370-394): the weighted strategies fuse the
    scores the two searches return (the dense list's second field as the index returns it — cosine for the
    two-stage search, 1 - cosine for the exact flat search — and the BM25 scores)."""

    def __init__(self, dense_index, sparse_index, rrf_k: float = 60.0, oversample: int = 4, exact_dense: bool = False,
                 fusion: str = "rrf", weights=(0.7, 0.2, 0.1)):
        assert fusion in ("rrf", "linear", "normalized")
        self.dense, self.sparse = dense_index, sparse_index
        self.k, self.oversample, self.exact_dense = rrf_k, oversample, exact_dense
        self.fusion, self.weights = fusion, tuple(float(w) for w in weights)
        self._lib = _ffi.lib()

    def search_batch_device(self, dense_queries_t, sparse_queries, limit: int):
        """dense_queries_t [nq, dim] f32 CUDA tensor (or None), sparse_queries list of (term ids, tfs)
        (or None) -> fused ids [nq, limit] int64 (-1 unfilled) and RRF scores [nq, limit] f32, on the GPU."""
        import torch
        dev = torch.device("cuda", self.dense.device)
        want = 2 * limit
        nq = dense_queries_t.shape[0] if dense_queries_t is not None else len(sparse_queries)
        d_ids = s_ids = d_sc = s_sc = None
        if dense_queries_t is not None:
            if self.exact_dense:
                d_ids, d_sc = self.dense.flat_search_batch_device(dense_queries_t, want)
            else:
                d_ids, d_sc = self.dense.search_batch_device(dense_queries_t, want, want * self.oversample)
        if sparse_queries is not None:
            s_ids, s_sc = self.sparse.search_bm25_batch_device(sparse_queries, want)
        ids = torch.empty((nq, limit), dtype=torch.int64, device=dev)
        sc = torch.empty((nq, limit), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        if self.fusion != "rrf":
            raise_for_status(self._lib.gvdb_weighted_fusion_batch_device(
                self.dense.device, C.c_void_p(st), p(d_ids), p(d_sc), want if d_ids is not None else 0,
                p(s_ids), p(s_sc), want if s_ids is not None else 0, None, None, 0, nq,
                self.weights[0], self.weights[1], self.weights[2], 1 if self.fusion == "normalized" else 0, limit,
                C.c_void_p(ids.data_ptr()), C.c_void_p(sc.data_ptr())), self._lib)
            return ids, sc
        raise_for_status(self._lib.gvdb_rrf_fusion_batch_device(
            self.dense.device, C.c_void_p(st),
            C.c_void_p(d_ids.data_ptr()) if d_ids is not None else None, want if d_ids is not None else 0,
            C.c_void_p(s_ids.data_ptr()) if s_ids is not None else None, want if s_ids is not None else 0,
            None, 0, nq, self.k, limit, C.c_void_p(ids.data_ptr()), C.c_void_p(sc.data_ptr())), self._lib)
        return ids, sc

    def search_batch(self, dense_queries, sparse_queries, limit: int):
        """Host arrays in, host arrays out (ids uint64 with NO_ID unfilled)."""
        import torch
        dev = torch.device("cuda", self.dense.device)
        qd = None if dense_queries is None else torch.from_numpy(_np(dense_queries, np.float32)).to(dev)
        ids, sc = self.search_batch_device(qd, sparse_queries, limit)
        return ids.cpu().numpy().astype(np.uint64), sc.cpu().numpy()
