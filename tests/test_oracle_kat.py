"""Pins the CPU oracle to known answers derived BY HAND from the inputs of the reference's own
tests (the reference asserts only shapes there):
  src/quantization.rs:361-400 (test_binary_quantization, test_hamming_distance,
  test_binary_vector_store), src/hybrid.rs:991-1025 (test_rrf_fusion),
  src/sparse.rs:383-390 (test_sparse_vector_operations).
The reference cannot be executed in this image (no cargo/rustc): parity is otherwise unpinned.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_quantize_reference_test_vector():
    # quantization.rs:365-370: [0.5,-0.3,0.8,-0.1,0.2] -> bits 1,0,1,0,1 (comment at :369)
    code = oracle.quantize([0.5, -0.3, 0.8, -0.1, 0.2])
    assert code.tolist() == [0xA8]           # 10101000, Msb0, pad bits 0
    assert code.size == 1                    # byte_size() = ceil(5/8)


def test_hamming_reference_test_vectors():
    # quantization.rs:378-385
    a = oracle.quantize([1.0, -1.0, 1.0, -1.0])
    b = oracle.quantize([1.0, 1.0, -1.0, -1.0])
    assert a.tolist() == [0xA0] and b.tolist() == [0xC0]
    assert oracle.hamming(a, b) == 2
    assert oracle.similarity(a, b, 4) == np.float32(0.5)     # 1 - 2/4


def test_store_reference_vector():
    # quantization.rs:394-395: [0.1,0.2,0.3] -> 111
    assert oracle.quantize([0.1, 0.2, 0.3]).tolist() == [0xE0]


def test_threshold_is_strict_and_nan_zero():
    x = np.array([0.0, -0.0, np.nan, 1e-45, -1e-45, np.inf, -np.inf, 0.5], dtype=np.float32)
    assert oracle.quantize(x).tolist() == [0b00010101]
    assert oracle.quantize(np.array([0.5, 0.5000001], dtype=np.float32), 0.5).tolist() == [0x40]


def test_dim_not_multiple_of_8_pads_zero():
    x = np.ones(13, dtype=np.float32)
    assert oracle.quantize(x).tolist() == [0xFF, 0xF8]


def test_rescore_count_f32_semantics():
    # quantization.rs:178-179: (len as f32 * ratio) as usize, min(len)
    assert oracle.rescore_count(10_000, 0.1) == 1000
    assert oracle.rescore_count(7, 0.1) == 0
    assert oracle.rescore_count(1_000_000, 0.1) == 100_000
    assert oracle.rescore_count(10, 2.0) == 10
    assert oracle.rescore_count(10, -1.0) == 0
    assert oracle.rescore_count(10, float("nan")) == 0
    # computed in f32: 16_777_217 is not representable, 0.1f32 is 0.100000001490116...
    n = 16_777_217
    assert oracle.rescore_count(n, 0.1) == int(np.float32(np.float32(n) * np.float32(0.1)))


def test_cosine_zero_norm_conventions():
    z = np.zeros(4, dtype=np.float32)
    v = np.array([1, 2, 3, 4], dtype=np.float32)
    assert oracle.cosine_similarity(z, v) == 0.0          # quantization.rs:211-212
    assert oracle.cosine_distance(z, v) == np.inf         # index.rs:695-697
    w = np.array([3, 4, 0, 0], dtype=np.float32)          # ||w|| = 5 exactly
    assert oracle.cosine_similarity(w, w) == np.float32(1.0)
    assert oracle.cosine_distance(w, w) == np.float32(0.0)
    # sqrt(30)*sqrt(30) rounds above 30 in f32: the faithful fold gives 1 - 2^-24, not 1.0
    assert oracle.cosine_similarity(v, v) == np.float32(0.99999994)


def test_cosine_is_sequential_f32_fold():
    rng = np.random.default_rng(0)
    a = rng.standard_normal(768).astype(np.float32)
    b = rng.standard_normal(768).astype(np.float32)
    dot = np.float32(0)
    sa = np.float32(0)
    sb = np.float32(0)
    for x, y in zip(a, b):
        dot = np.float32(dot + np.float32(x * y))
        sa = np.float32(sa + np.float32(x * x))
        sb = np.float32(sb + np.float32(y * y))
    want = np.float32(dot / np.float32(np.sqrt(sa) * np.sqrt(sb)))
    assert oracle.cosine_similarity(a, b) == want


def test_rrf_reference_test_lists():
    # hybrid.rs:1001-1016 lists; doc ids 1..4; k = 60
    dense = [1, 2, 3]
    sparse = [2, 1, 4]
    ids, sc = oracle.rrf_fusion(dense, sparse, [], 60.0)
    d = dict(zip(ids.tolist(), sc.tolist()))
    f = np.float32
    assert d[1] == f(f(1) / f(61)) + f(f(1) / f(62))
    assert np.isclose(d[1], 0.032522473, rtol=1e-6)
    assert d[2] == d[1]                                   # ranks (2,1) and (1,2): same sum
    assert d[3] == f(f(1) / f(63)) and d[4] == f(f(1) / f(63))
    assert ids.tolist()[:2] == [1, 2] and set(ids.tolist()[2:]) == {3, 4}
    assert d[1] > d[3]                                    # the reference's assertion (:1022-1024)


def test_multi_stage_select_equals_full_sort():
    rng = np.random.default_rng(1)
    rows = rng.integers(-3, 4, size=(500, 40)).astype(np.float32)   # many ties
    q = rng.integers(-3, 4, size=40).astype(np.float32)
    a = oracle.multi_stage_search(q, rows, 50)
    b = oracle.multi_stage_search(q, rows, 50, select=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_multi_stage_order_is_ham_then_index_then_cos():
    rows = np.array([[1, 1, 1, 1], [1, 1, 1, 1], [1, -1, 1, 1], [2, 2, 2, 2], [-1, -1, -1, -1]],
                    dtype=np.float32)
    q = np.array([1, 1, 1, 1], dtype=np.float32)
    idx, sc, ci, ch = oracle.multi_stage_search(q, rows, 4, want_candidates=True)
    assert ci.tolist() == [0, 1, 3, 2] and ch.tolist() == [0, 0, 0, 1]
    assert idx.tolist() == [0, 1, 3, 2]                   # cos ties keep stage-1 order
    assert sc[0] == np.float32(1.0)


def test_flat_search_skips_tombstones_and_sorts_ascending():
    rows = np.array([[1, 0], [0, 1], [1, 1], [0, 0], [-1, 0]], dtype=np.float32)
    q = np.array([1, 0], dtype=np.float32)
    idx, d = oracle.flat_search(q, rows, 5)
    assert idx.tolist() == [0, 2, 1, 4, 3] and d[-1] == np.inf
    idx, d = oracle.flat_search(q, rows, 2, live=[0, 1, 1, 1, 1])
    assert idx.tolist() == [2, 1]


def test_shard_merge_equals_single_index():
    from grape_vector_db_b200 import synth
    rows = synth.lowrank_rows(0, 3000, 64)
    q = synth.lowrank_queries(0, 1, 64)[0]
    R, k = 16, 5
    want_idx, want_sc = oracle.multi_stage_search(q, rows, R)
    hams, idxs, scs = [], [], []
    for lo, hi in ((0, 1000), (1000, 2000), (2000, 3000)):
        i, s, ci, ch = oracle.multi_stage_search(q, rows[lo:hi], R, want_candidates=True)
        order = {int(c): t for t, c in enumerate(ci)}
        sc_by_cand = np.zeros(R, dtype=np.float32)
        for ii, ss in zip(i, s):
            sc_by_cand[order[int(ii)]] = ss
        hams.append(ch); idxs.append(ci + np.uint64(lo)); scs.append(sc_by_cand)
    gi, gs = oracle.shard_merge(np.concatenate(hams), np.concatenate(idxs), np.concatenate(scs), R, k)
    assert np.array_equal(gi, want_idx[:k]) and np.array_equal(gs, want_sc[:k])


def test_bm25_known_values():
    # 3 docs over 2 terms; doc_len = sum of tf (sparse.rs:341)
    post_off = [0, 2, 3]
    post_doc = [0, 1, 2]
    post_tf = [1.0, 2.0, 1.0]
    doc_len = [1.0, 2.0, 1.0]
    avg = oracle.bm25_avg_len(post_off, post_doc, doc_len)
    assert avg == np.float32(4.0 / 3.0)                   # (1+2+1)/3, one entry per (term,doc)
    docs, sc = oracle.bm25_search([0], [1.0], post_off, post_doc, post_tf, doc_len, 10)
    f = np.float32
    idf = f(np.log(f((f(3) - f(2) + f(0.5)) / (f(2) + f(0.5)))))
    def tfc(tf, ln):
        return f(f(tf) * f(2.2)) / f(f(tf) + f(f(1.2) * f(f(f(1.0) - f(0.75)) + f(f(0.75) * f(f(ln) / avg)))))
    want = {0: f(f(f(1.0) * tfc(1.0, 1.0)) * idf), 1: f(f(f(1.0) * tfc(2.0, 2.0)) * idf)}
    got = dict(zip(docs.tolist(), sc.tolist()))
    assert set(got) == {0, 1}
    for d in (0, 1):
        assert np.isclose(got[d], want[d], rtol=2e-6)


def test_golden_fixture_roundtrip():
    """tests/golden/kat_small.json is produced by tests/golden/make_golden.py from the oracle on
    seeded inputs; it freezes the oracle's behaviour so later edits cannot drift silently."""
    path = os.path.join(GOLDEN, "kat_small.json")
    if not os.path.exists(path):
        pytest.skip("golden fixture not generated")
    g = json.load(open(path))
    from grape_vector_db_b200 import synth
    for case in g["cases"]:
        gen = synth.lowrank_rows if case["dataset"] == "lowrank" else synth.iid_rows
        genq = synth.lowrank_queries if case["dataset"] == "lowrank" else synth.iid_queries
        rows = gen(0, case["n"], case["dim"])
        qs = genq(0, case["nq"], case["dim"])
        codes = oracle.quantize_batch(rows)
        assert [int(x) for x in codes[:4].ravel()[:32]] == case["codes_head"]
        for qi in range(case["nq"]):
            idx, sc = oracle.multi_stage_search(qs[qi], rows, case["R"], codes=codes)
            assert idx[:case["k"]].tolist() == case["ids"][qi]
            assert sc[:case["k"]].view(np.uint32).tolist() == case["score_bits"][qi]


def test_storage_vector_search_kat():
    """BasicVectorStore::vector_search (src/storage.rs:296-339): cosine similarity, zero norm -> 0.0,
    `similarity < threshold` skipped, stable descending sort, truncate."""
    rows = np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 0], [-1, 0, 0], [2, 0, 0]], dtype=np.float32)
    q = np.array([1, 0, 0], dtype=np.float32)
    idx, sim = oracle.similarity_search(q, rows, 10)
    # rows 0 and 5 tie at 1.0 (row order kept), then 1/sqrt(2), then the two zeros (row 1 orthogonal,
    # row 3 zero norm -> 0.0), then -1
    assert idx.tolist() == [0, 5, 2, 1, 3, 4]
    assert sim[0] == 1.0 and sim[1] == 1.0 and sim[3] == 0.0 and sim[4] == 0.0 and sim[5] == -1.0
    assert sim[2] == np.float32(1.0) / (np.float32(1.0) * np.sqrt(np.float32(2.0)))
    idx, sim = oracle.similarity_search(q, rows, 10, threshold=0.0)      # -1 < 0 is dropped, 0.0 stays
    assert idx.tolist() == [0, 5, 2, 1, 3]
    idx, sim = oracle.similarity_search(q, rows, 2, threshold=0.5)
    assert idx.tolist() == [0, 5]


def test_sparse_synthetic_corpus_shape():
    """The C4 sparse generator: tf = count/tokens, document_length = sum of tfs = 1,
    CSR postings ascending by document inside a term; the oracle's BM25 runs on it."""
    from grape_vector_db_b200 import synth
    post_off, post_doc, post_tf, doc_len = synth.sparse_corpus(2000, vocab=500)
    assert post_off[-1] == post_doc.size == post_tf.size and doc_len.shape == (2000,)
    assert np.all(doc_len == np.float32(1.0))
    for t in range(500):
        d = post_doc[int(post_off[t]):int(post_off[t + 1])]
        assert np.all(d[1:] > d[:-1])
    per_doc = np.zeros(2000)
    np.add.at(per_doc, post_doc, post_tf.astype(np.float64))
    assert np.allclose(per_doc, 1.0)
    again = synth.sparse_corpus(2000, vocab=500)
    assert all(np.array_equal(a, b) for a, b in zip((post_off, post_doc, post_tf, doc_len), again))
    q = synth.sparse_queries(8, vocab=500)
    docs, sc = oracle.bm25_search(q[0][0], q[0][1], post_off, post_doc, post_tf, doc_len, 20)
    assert len(docs) == 20 and np.all(sc[:-1] >= sc[1:])


def test_golden_hybrid_fixture_roundtrip():
    """tests/golden/kat_hybrid.json freezes the oracle's BM25 lists, fused lists and a filtered search on
    seeded inputs (same generator script); the GPU tests check the CUDA path against the same file."""
    from grape_vector_db_b200 import synth
    g = json.load(open(os.path.join(GOLDEN, "kat_hybrid.json")))["hybrid"]
    n, dim, nq, limit = g["n"], g["dim"], g["nq"], g["limit"]
    rows, qs = synth.lowrank_rows(0, n, dim), synth.lowrank_queries(0, nq, dim)
    post = synth.sparse_corpus(n, vocab=g["vocab"])
    sq = synth.sparse_queries(nq, vocab=g["vocab"])
    assert int(post[0][-1]) == g["postings"]
    assert int(np.float32(oracle.bm25_avg_len(post[0], post[1], post[3])).view(np.uint32)) == g["avg_len_bits"]
    want = 2 * limit
    dense_ids, _ = oracle.multi_stage_search_batch(qs, rows, want * g["oversample"], want)
    for q in range(nq):
        d, s = oracle.bm25_search(sq[q][0], sq[q][1], *post, want)
        assert d.tolist() == g["bm25"][q]["docs"] and s.view(np.uint32).tolist() == g["bm25"][q]["score_bits"]
        fi, fs = oracle.rrf_fusion(dense_ids[q], d, [], g["rrf_k"])
        assert fi[:limit].tolist() == g["fused"][q]["ids"]
        assert fs[:limit].view(np.uint32).tolist() == g["fused"][q]["score_bits"]


def test_msb0_packing_and_hamming_against_numpy_packbits():
    """An independent pin of the two conventions everything else rests on: BitVec<u8, Msb0>
    packing (src/quantization.rs:98-109: bit i of the vector is bit 7 - i % 8 of byte i / 8) is numpy's
    packbits(bitorder="big"), and hamming::distance (crate hamming 0.1.3, :137-139) is the popcount of the XOR,
    i.e. the number of differing unpacked bits — on the reference's own test inputs (:361-386) and on random
    vectors of ragged dimensions, without going through the oracle's own packing code."""
    cases = [np.array([0.5, -0.3, 0.8, -0.1, 0.2], dtype=np.float32),          # test_binary_quantization
             np.array([1.0, -1.0, 1.0, -1.0], dtype=np.float32),                # test_hamming_distance, vec1
             np.array([1.0, 1.0, -1.0, -1.0], dtype=np.float32),                # test_hamming_distance, vec2
             np.array([0.1, 0.2, 0.3], dtype=np.float32)]                       # test_binary_vector_store
    rng = np.random.default_rng(7)
    for dim in (1, 7, 8, 9, 63, 64, 65, 100, 768, 1000, 1536):
        cases.append(rng.standard_normal(dim).astype(np.float32))
        cases.append(rng.standard_normal(dim).astype(np.float32))
    for x in cases:
        want = np.packbits((x > np.float32(0.0)).astype(np.uint8), bitorder="big")
        got = oracle.quantize(x)
        assert got.dtype == np.uint8 and np.array_equal(got[:want.size], want) and not got[want.size:].any()
    assert oracle.quantize(cases[0])[0] == 0b10101000                           # [1, 0, 1, 0, 1] -> 0xA8
    for a, b in zip(cases[1::2], cases[2::2]):
        if a.size != b.size:
            continue
        ba, bb = (a > 0).astype(np.uint8), (b > 0).astype(np.uint8)
        d = int(np.count_nonzero(ba != bb))
        ca, cb = oracle.quantize(a), oracle.quantize(b)
        assert oracle.hamming(ca, cb) == d
        assert d == int(np.unpackbits(np.bitwise_xor(ca, cb)).sum())
        assert oracle.similarity(ca, cb, a.size) == np.float32(1.0) - np.float32(d) / np.float32(a.size)
    # the reference's test_hamming_distance pair: two of four signs differ
    assert oracle.hamming(oracle.quantize(cases[1]), oracle.quantize(cases[2])) == 2
