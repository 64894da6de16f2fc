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
    for limit in (1, 7, 4096, 6000):          # 6000: above the block sort's capacity (device radix sort per query)
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


def test_bm25_both_paths_and_their_limits(gv, monkeypatch):
    """The blocked path (shared-memory accumulators: limit <= 1024, <= 64 query terms) and the dense-accumulator
    path (everything else, or GVDB_BM25_DENSE=1) answer alike: a query of exactly 64 terms, one of 70, limits on
    both sides of 1024, a single-query batch (every document block its own segment), ties with a large limit."""
    from grape_vector_db_b200 import synth
    n_docs, vocab = 60_000, 3_000
    post = synth.sparse_corpus(n_docs, vocab=vocab)
    rng = np.random.default_rng(11)
    def many(nt):
        t = rng.choice(vocab, size=nt, replace=False).astype(np.uint32)
        return t, np.full(nt, 1.0 / nt, dtype=np.float32)
    queries = synth.sparse_queries(6, vocab=vocab) + [many(64), many(70)]
    _check(gv, *post, queries[:7], 1000)          # blocked, LP = 1024
    _check(gv, *post, queries, 50)                # a 70-term query sends the batch down the dense path
    _check(gv, *post, queries[:7], 1500)          # limit > 1024: dense
    _check(gv, *post, queries[:3], 5000)          # limit > 4096: dense, ordered by a device radix sort
    _check(gv, *post, queries[:1], 200)           # one query: many segments of one block each
    monkeypatch.setenv("GVDB_BM25_DENSE", "1")
    _check(gv, *post, queries[:7], 200)           # the dense path on what the blocked path normally answers
    monkeypatch.delenv("GVDB_BM25_DENSE")
    n = 70_000                                    # all scores tie: the block select walks ties in document order
    _check(gv, np.array([0, n], dtype=np.uint64), np.arange(n, dtype=np.uint32), np.full(n, 0.25, dtype=np.float32),
           np.ones(n, dtype=np.float32), [([0], [1.0])], 1000)


def test_bm25_empty_index_and_bad_postings(gv):
    with gv.GpuSparseIndex() as sp:
        docs, sc = sp.search_bm25_batch([([0], [1.0])], 5)          # src/sparse.rs:161-163
        assert np.all(docs == gv.NO_ID)
        with pytest.raises(gv.ConfigError):
            sp.build(np.array([0, 2], dtype=np.uint64), np.array([1, 1], dtype=np.uint32),
                     np.ones(2, dtype=np.float32), np.ones(3, dtype=np.float32))


# ---- rrf_fusion on the GPU (src/hybrid.rs:422-488) ---------------------------------------------
def test_rrf_known_values(gv):
    # the lists of the reference's test_rrf_fusion (src/hybrid.rs:991-1025) as document numbers:
    # dense doc1, doc2, doc3; sparse doc2, doc1, doc4
    ids, sc = gv.rrf_fusion_batch([[1, 2, 3]], [[2, 1, 4]], None, 60.0)
    one = np.float32(1.0)
    d1 = one / np.float32(61) + one / np.float32(62)
    d2 = one / np.float32(62) + one / np.float32(61)
    assert ids[0, :2].tolist() == [1, 2] and sc[0, 0] == d1 and sc[0, 1] == d2
    assert ids[0, 2:4].tolist() == [3, 4] and sc[0, 2] == sc[0, 3] == one / np.float32(63)   # tie: first appearance
    assert np.all(ids[0, 4:] == gv.NO_ID)


def test_rrf_matches_oracle(gv):
    rng = np.random.default_rng(11)
    for n_d, n_s, n_t, pool in ((200, 200, 0, 300), (200, 200, 200, 350), (7, 0, 5, 9), (0, 64, 0, 100),
                                (1500, 1500, 1000, 2500)):
        nq = 9
        def lists(n):
            if n == 0:
                return None
            out = np.full((nq, n), gv.NO_ID, dtype=np.uint64)
            for q in range(nq):
                m = n if q % 3 else int(rng.integers(0, n + 1))          # some lists end early
                out[q, :m] = rng.choice(pool, size=min(m, pool), replace=False)[:m].astype(np.uint64) if m <= pool else 0
            return out
        d, s, t = lists(n_d), lists(n_s), lists(n_t)
        limit = 100
        ids, sc = gv.rrf_fusion_batch(d, s, t, 60.0, limit)
        for q in range(nq):
            trim = lambda a: [] if a is None else a[q][a[q] != gv.NO_ID]
            oi, os_ = oracle.rrf_fusion(trim(d), trim(s), trim(t), 60.0)
            r = min(limit, len(oi))
            assert np.array_equal(ids[q, :r], oi[:r]), (n_d, n_s, n_t, q)
            assert np.array_equal(_bits(sc[q, :r]), _bits(os_[:r]))
            assert np.all(ids[q, r:] == gv.NO_ID) and np.all(np.isneginf(sc[q, r:]))


def test_weighted_fusions_match_oracle(gv):
    """linear_fusion and normalized_fusion (src/hybrid.rs:491-616) on the GPU: documents and score BITS equal the
    oracle's, incl. duplicates inside the dense list (insert overwrites), lists cut short by NO_ID, exact score
    ties (first appearance), constant lists (normalize -> 1.0), an empty text list."""
    rng = np.random.default_rng(21)
    nq, nd, ns, nt = 9, 40, 30, 12
    d = rng.integers(0, 60, size=(nq, nd)).astype(np.uint64)          # duplicates on purpose
    s = np.stack([rng.choice(60, size=ns, replace=False) for _ in range(nq)]).astype(np.uint64)
    t = np.stack([rng.choice(60, size=nt, replace=False) for _ in range(nq)]).astype(np.uint64)
    dsc = (rng.integers(-16, 17, size=(nq, nd)) / 8).astype(np.float32)  # coarse grid: plenty of exact ties
    ssc = (rng.integers(0, 40, size=(nq, ns)) / 4).astype(np.float32)
    tsc = (rng.integers(0, 6, size=(nq, nt))).astype(np.float32)
    dsc[3] = 0.25                                                       # a constant list
    d[1, 25:] = gv.NO_ID; s[2, 7:] = gv.NO_ID; t[4, :] = gv.NO_ID       # lists that end early / an empty list
    trim = lambda a, sc: (a[a != gv.NO_ID][:np.argmax(a == gv.NO_ID) if (a == gv.NO_ID).any() else a.size],
                          sc[:np.argmax(a == gv.NO_ID) if (a == gv.NO_ID).any() else a.size])
    for normalize in (False, True):
        for limit in (15, nd + ns + nt):
            ids, sc = gv.weighted_fusion_batch(d, dsc, s, ssc, t, tsc, (0.7, 0.2, 0.1), normalize, limit)
            for q in range(nq):
                (a, asc), (b, bsc), (c, csc) = trim(d[q], dsc[q]), trim(s[q], ssc[q]), trim(t[q], tsc[q])
                oi, os_ = oracle.weighted_fusion(a, asc, b, bsc, c, csc, (0.7, 0.2, 0.1), normalize)
                r = min(limit, len(oi))
                assert np.array_equal(ids[q, :r], oi[:r]), (normalize, limit, q)
                assert np.array_equal(_bits(sc[q, :r]), _bits(os_[:r])), (normalize, limit, q)
                assert np.all(ids[q, r:] == gv.NO_ID) and np.all(np.isneginf(sc[q, r:]))
    ids, sc = gv.weighted_fusion_batch([[1, 2]], [[0.9, 0.5]], [[2, 3]], [[4.0, 1.0]], None, None, (0.7, 0.2, 0.1), False)
    assert ids[0].tolist()[:3] == [2, 1, 3]


def test_hybrid_search_matches_oracle_composition(gv):
    """HybridSearchEngine::search (src/hybrid.rs:286-356) on the GPU == oracle dense list + oracle BM25
    list + oracle RRF, for two-stage and exact dense lists."""
    from grape_vector_db_b200 import synth
    n, dim, vocab, nq, limit = 20_000, 256, 3_000, 70, 25
    rows = synth.lowrank_rows(0, n, dim)
    qs = synth.lowrank_queries(0, nq, dim)
    post = synth.sparse_corpus(n, vocab=vocab)
    sq = synth.sparse_queries(nq, vocab=vocab)
    want = 2 * limit
    with gv.GpuIndex(dim) as dense, gv.GpuSparseIndex() as sparse:
        dense.add(rows)
        sparse.build(*post)
        for exact in (False, True):
            hy = gv.HybridSearcher(dense, sparse, rrf_k=60.0, oversample=4, exact_dense=exact)
            ids, sc = hy.search_batch(qs, sq, limit)
            if exact:
                od, _ = oracle.flat_search_batch(qs, rows, want)
            else:
                od, _ = oracle.multi_stage_search_batch(qs, rows, want * 4, want, nthreads=8)
            for q in range(nq):
                ob, _ = oracle.bm25_search(sq[q][0], sq[q][1], *post, want)
                oi, os_ = oracle.rrf_fusion(od[q], ob, [], 60.0)
                assert np.array_equal(ids[q], oi[:limit]), (exact, q)
                assert np.array_equal(_bits(sc[q]), _bits(os_[:limit]))
        # the weighted strategies fuse the scores the two searches return
        od, osc = oracle.multi_stage_search_batch(qs, rows, want * 4, want, nthreads=8)
        for fusion in ("linear", "normalized"):
            hy = gv.HybridSearcher(dense, sparse, oversample=4, fusion=fusion, weights=(0.6, 0.3, 0.1))
            ids, sc = hy.search_batch(qs, sq, limit)
            for q in range(nq):
                ob, obs = oracle.bm25_search(sq[q][0], sq[q][1], *post, want)
                oi, os_ = oracle.weighted_fusion(od[q], osc[q], ob, obs, [], [], (0.6, 0.3, 0.1), fusion == "normalized")
                assert np.array_equal(ids[q], oi[:limit]), (fusion, q)
                assert np.array_equal(_bits(sc[q]), _bits(os_[:limit])), (fusion, q)
        # sparse-only and dense-only requests (the Option fields of HybridSearchRequest)
        hy = gv.HybridSearcher(dense, sparse)
        ids, _ = hy.search_batch(None, sq[:8], limit)
        for q in range(8):
            ob, _ = oracle.bm25_search(sq[q][0], sq[q][1], *post, want)
            assert np.array_equal(ids[q, :min(limit, len(ob))], ob[:limit])
        ids, _ = hy.search_batch(qs[:8], None, limit)
        od, _ = oracle.multi_stage_search_batch(qs[:8], rows, want * 4, want, nthreads=8)
        assert np.array_equal(ids, od[:, :limit])


def test_golden_hybrid_fixture(gv):
    """The CUDA path against the committed fixture tests/golden/kat_hybrid.json (BM25 lists, fused lists,
    a filtered two-stage search)."""
    import json
    import os
    from grape_vector_db_b200 import synth
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat_hybrid.json")))["hybrid"]
    n, dim, nq, limit = g["n"], g["dim"], g["nq"], g["limit"]
    rows, qs = synth.lowrank_rows(0, n, dim), synth.lowrank_queries(0, nq, dim)
    post = synth.sparse_corpus(n, vocab=g["vocab"])
    sq = synth.sparse_queries(nq, vocab=g["vocab"])
    with gv.GpuIndex(dim) as dense, gv.GpuSparseIndex() as sparse:
        dense.add(rows)
        sparse.build(*post)
        assert int(sparse.average_document_length.view(np.uint32)) == g["avg_len_bits"]
        docs, sc = sparse.search_bm25_batch(sq, 2 * limit)
        for q in range(nq):
            r = len(g["bm25"][q]["docs"])
            assert docs[q, :r].tolist() == g["bm25"][q]["docs"] and _bits(sc[q, :r]).tolist() == g["bm25"][q]["score_bits"]
        ids, fs = gv.HybridSearcher(dense, sparse, rrf_k=g["rrf_k"], oversample=g["oversample"]).search_batch(qs, sq, limit)
        for q in range(nq):
            assert ids[q].tolist() == g["fused"][q]["ids"] and _bits(fs[q]).tolist() == g["fused"][q]["score_bits"]
        f = g["filtered"]
        fi, fsc = dense.search_batch_filtered(qs, (np.arange(n) % 7) < 2, f["k"], f["R"])
        assert fi.tolist() == f["ids"] and _bits(fsc).tolist() == f["score_bits"]
