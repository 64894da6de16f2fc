"""GPU parity of the sparse side of the hybrid search:
gvdb_sparse_search_bm25_batch through the C ABI against the oracle's restatement of
SparseIndex::search_bm25 (src/sparse.rs:153-222).  Documents and score BITS must be identical;
ties are ordered by document number on both sides (unspecified in the reference)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle  # noqa: E402


@pytest.fixture(scope="module")
def gv(built):
    import grape_vector_db_b200 as g
    return g


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _check(gv, post_off, post_doc, post_tf, doc_len, queries, limit, k1=1.2, b=0.75):
    with gv.GpuSparseIndex(k1, b) as sp:
        sp.build(post_off, post_doc, post_tf, doc_len)
        avg = oracle.bm25_avg_len(post_off, post_doc, doc_len)
        assert _bits([sp.average_document_length])[0] == _bits([avg])[0]
        docs, sc = sp.search_bm25_batch(queries, limit)
    for qi, (t, v) in enumerate(queries):
        od, os_ = oracle.bm25_search(t, v, post_off, post_doc, post_tf, doc_len, limit, avg, k1, b)
        r = len(od)
        assert np.array_equal(docs[qi, :r], od), f"BM25 documents differ (query {qi})"
        assert np.array_equal(_bits(sc[qi, :r]), _bits(os_)), f"BM25 score bits differ (query {qi})"
        assert np.all(docs[qi, r:] == gv.NO_ID) and np.all(np.isneginf(sc[qi, r:]))


def test_bm25_known_values(gv):
    # the oracle KAT (tests/test_oracle_kat.py::test_bm25_known_values): 3 docs over 2 terms
    post_off = np.array([0, 2, 3], dtype=np.uint64)
    post_doc = np.array([0, 1, 2], dtype=np.uint32)
    post_tf = np.array([0.5, 1.0, 1.0], dtype=np.float32)
    doc_len = np.array([0.5, 1.0, 1.0], dtype=np.float32)
    _check(gv, post_off, post_doc, post_tf, doc_len, [([0], [1.0]), ([1], [1.0]), ([0, 1], [0.5, 0.5]),
                                                      ([5], [1.0]), ([], [])], 10)


@pytest.mark.parametrize("n_docs,vocab,limit", [(20_000, 5_000, 200), (3_000, 50, 100), (200_000, 100_000, 200)])
def test_bm25_zipf_corpus(gv, n_docs, vocab, limit):
    from grape_vector_db_b200 import synth
    post = synth.sparse_corpus(n_docs, vocab=vocab)
    queries = synth.sparse_queries(48, vocab=vocab)
    _check(gv, *post, queries, limit)


def test_bm25_massive_ties_and_small_limits(gv):
    # every document has the same single term with the same tf: all scores tie, order = document number
    n = 70_000
    post_off = np.array([0, n], dtype=np.uint64)
    post_doc = np.arange(n, dtype=np.uint32)
    post_tf = np.full(n, 0.25, dtype=np.float32)
    doc_len = np.ones(n, dtype=np.float32)
    for limit in (1, 7, 4096):
        _check(gv, post_off, post_doc, post_tf, doc_len, [([0], [1.0]), ([0, 0], [0.5, 0.25])], limit)


def test_bm25_negative_idf_and_repeated_terms(gv):
    # df > N/2 gives a negative idf (the reference keeps it); a term repeated in the query adds twice
    rng = np.random.default_rng(5)
    n, vocab = 5_000, 8
    rows = [np.sort(rng.choice(n, size=int(n * f), replace=False)).astype(np.uint32)
            for f in (0.9, 0.7, 0.5, 0.3, 0.1, 0.05, 0.01, 0.0)]
    post_off = np.zeros(vocab + 1, dtype=np.uint64)
    post_off[1:] = np.cumsum([len(r) for r in rows])
    post_doc = np.concatenate(rows)
    post_tf = (rng.integers(1, 9, size=post_doc.size) / 8).astype(np.float32)
    doc_len = (rng.integers(1, 40, size=n) / 8).astype(np.float32)
    queries = [([0, 4], [0.5, 0.5]), ([0], [1.0]), ([1, 1, 5], [0.25, 0.5, 0.25]), ([7], [1.0]), ([6, 3, 0], [0.3, 0.3, 0.4])]
    _check(gv, post_off, post_doc, post_tf, doc_len, queries, 300, k1=1.5, b=0.6)


def test_bm25_empty_index_and_bad_postings(gv):
    with gv.GpuSparseIndex() as sp:
        docs, sc = sp.search_bm25_batch([([0], [1.0])], 5)          # src/sparse.rs:161-163
        assert np.all(docs == gv.NO_ID)
        with pytest.raises(gv.ConfigError):
            sp.build(np.array([0, 2], dtype=np.uint64), np.array([1, 1], dtype=np.uint32),
                     np.ones(2, dtype=np.float32), np.ones(3, dtype=np.float32))
