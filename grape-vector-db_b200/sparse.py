"""GpuSparseIndex: the sparse (BM25) side of the hybrid search on one B200, through the C ABI.

Mirrors SparseIndex of the reference (src/sparse.rs:31-222): `add_document` collects
(document, term, tf) postings on the host, `build()` freezes them into CSR and uploads the
snapshot, `search_bm25(_batch)` scores on the GPU (gvdb_sparse_search_bm25_batch).  Documents are
dense numbers; the String ids stay with the caller, like the dense index's row numbers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .errors import raise_for_status
from .index import NO_ID, _np, _ptr


class GpuSparseIndex:
    def __init__(self, k1: float = 1.2, b: float = 0.75, device: int = 0):
        self._lib = _ffi.lib()
        h = C.c_void_p()
        raise_for_status(self._lib.gvdb_sparse_create(device, k1, b, C.byref(h)), self._lib)
        self._h = h
        self.k1, self.b = k1, b
        self.device = device
        self.n_docs = 0
        self.n_terms = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gvdb_sparse_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def build(self, post_off, post_doc, post_tf, doc_len):
        """CSR postings by term id (documents strictly ascending inside a term) + document lengths."""
        post_off = _np(post_off, np.uint64)
        post_doc = _np(post_doc, np.uint32)
        post_tf = _np(post_tf, np.float32)
        doc_len = _np(doc_len, np.float32)
        assert post_off.ndim == 1 and post_off.size >= 1 and post_doc.size == post_tf.size == int(post_off[-1])
        raise_for_status(self._lib.gvdb_sparse_build(self._h, doc_len.size, post_off.size - 1, _ptr(post_off),
                                                     _ptr(post_doc), _ptr(post_tf), _ptr(doc_len)), self._lib)
        self.n_docs, self.n_terms = int(doc_len.size), int(post_off.size - 1)

    @property
    def average_document_length(self) -> np.float32:
        return np.float32(self._lib.gvdb_sparse_average_document_length(self._h))

    @staticmethod
    def _csr(queries):
        nq = len(queries)
        q_off = np.zeros(nq + 1, dtype=np.uint64)
        for i, (t, _) in enumerate(queries):
            q_off[i + 1] = q_off[i] + np.uint64(len(t))
        q_terms = _np(np.concatenate([np.asarray(t, dtype=np.uint32) for t, _ in queries]) if nq else [], np.uint32)
        q_tfs = _np(np.concatenate([np.asarray(v, dtype=np.float32) for _, v in queries]) if nq else [], np.float32)
        return q_off, q_terms, q_tfs

    def search_bm25_batch(self, queries, limit: int):
        """queries: list of (term ids, query tfs).  Returns docs (nq, limit) u64 and scores (nq, limit) f32;
        unfilled slots are NO_ID / -inf."""
        nq = len(queries)
        q_off, q_terms, q_tfs = self._csr(queries)
        docs = np.full((nq, limit), NO_ID, dtype=np.uint64)
        sc = np.full((nq, limit), -np.inf, dtype=np.float32)
        raise_for_status(self._lib.gvdb_sparse_search_bm25_batch(self._h, nq, _ptr(q_off), _ptr(q_terms), _ptr(q_tfs),
                                                                 limit, _ptr(docs), _ptr(sc)), self._lib)
        return docs, sc

    def search_bm25_batch_device(self, queries, limit: int, docs_out=None, scores_out=None):
        """The same with the answers left on the GPU (torch tensors, current stream): docs int64
        (nq, limit), -1 unfilled; scores f32 (nq, limit), -inf unfilled."""
        import torch
        nq = len(queries)
        q_off, q_terms, q_tfs = self._csr(queries)
        dev = torch.device("cuda", self.device)
        if docs_out is None:
            docs_out = torch.empty((nq, limit), dtype=torch.int64, device=dev)
        if scores_out is None:
            scores_out = torch.empty((nq, limit), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        raise_for_status(self._lib.gvdb_sparse_search_bm25_batch_device(
            self._h, C.c_void_p(st), nq, _ptr(q_off), _ptr(q_terms), _ptr(q_tfs), limit,
            C.c_void_p(docs_out.data_ptr()), C.c_void_p(scores_out.data_ptr())), self._lib)
        return docs_out, scores_out

    @property
    def launches(self) -> int:
        return int(self._lib.gvdb_sparse_launches(self._h))

    def search_bm25(self, terms, tfs, limit: int):
        d, s = self.search_bm25_batch([(terms, tfs)], limit)
        n = int((d[0] != NO_ID).sum())
        return d[0, :n], s[0, :n]
