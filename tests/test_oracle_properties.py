"""Property tests (hypothesis) of the C++ oracle against the independent Python restatement in
tests/test_oracle_independent.py and against numpy, over inputs a seeded generator does not reach:
arbitrary f32 bit patterns (NaN, +-inf, +-0, subnormals) for the quantizer, arbitrary thresholds,
dimensions that are not multiples of 8, heavy ties.  CPU only; the oracle is the checker here."""
import os
import sys

import numpy as np
from hypothesis import given, settings, strategies as st
from hypothesis.extra import numpy as hnp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))       # the sibling module, whatever the import mode

from oracle import oracle                                            # noqa: E402
from test_oracle_independent import F, _bits, py_cosine, py_flat, py_hamming, py_quantize, py_two_stage   # noqa: E402

# the same examples on every run (the driver's CPU pass must not depend on a random seed), no example database on disk
settings.register_profile("gvdb", derandomize=True, deadline=None, database=None)
settings.load_profile("gvdb")

ANY_F32 = st.floats(width=32, allow_nan=True, allow_infinity=True)
FINITE = st.floats(min_value=-1e6, max_value=1e6, width=32, allow_nan=False)     # sums of 64 squares stay finite in f32


@settings(max_examples=150, deadline=None)
@given(x=st.lists(ANY_F32, min_size=1, max_size=70), thr=st.floats(width=32, allow_nan=False, allow_infinity=False))
def test_quantize_any_bit_pattern(x, thr):
    """bit = value > threshold, strictly (src/quantization.rs:99): NaN, the threshold itself and -0.0 at
    threshold 0 give 0; Msb0 bytes; pad bits zero — also against numpy.packbits."""
    xs = np.asarray(x, dtype=np.float32)
    code = oracle.quantize(xs, thr)
    assert bytes(code) == py_quantize(xs, F(thr))
    with np.errstate(invalid="ignore"):
        assert np.array_equal(code, np.packbits(xs > F(thr), bitorder="big"))


@settings(max_examples=100, deadline=None)
@given(data=st.data(), nbytes=st.integers(1, 40))
def test_hamming_is_the_number_of_differing_bits(data, nbytes):
    a = np.frombuffer(data.draw(st.binary(min_size=nbytes, max_size=nbytes)), dtype=np.uint8)
    b = np.frombuffer(data.draw(st.binary(min_size=nbytes, max_size=nbytes)), dtype=np.uint8)
    d = oracle.hamming(a, b)
    assert d == py_hamming(bytes(a), bytes(b)) == int(np.unpackbits(a ^ b).sum())
    assert oracle.hamming(b, a) == d and oracle.hamming(a, a) == 0
    dim = nbytes * 8
    assert _bits(oracle.similarity(a, b, dim)) == _bits(F(F(1.0) - F(F(d) / F(dim))))       # :145-147


@settings(max_examples=150, deadline=None)
@given(data=st.data(), dim=st.integers(1, 64))
def test_cosine_is_three_sequential_f32_folds(data, dim):
    q = np.asarray(data.draw(st.lists(FINITE, min_size=dim, max_size=dim)), dtype=np.float32)
    c = np.asarray(data.draw(st.lists(FINITE, min_size=dim, max_size=dim)), dtype=np.float32)
    with np.errstate(all="ignore"):
        want = py_cosine(q, c)
    assert _bits(oracle.cosine_similarity(q, c)) == _bits(want)
    # the flat index's distance: 1 - cos, +inf when a norm is zero (src/index.rs:686-700)
    with np.errstate(all="ignore"):
        flat = py_flat(q, c[None, :], np.ones(1, dtype=bool), 1)[0][1]
    assert _bits(oracle.cosine_distance(q, c)) == _bits(flat)


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 1 << 40), ratio=st.floats(min_value=-1.0, max_value=3.0, width=32, allow_nan=False))
def test_rescore_count_is_the_f32_product_truncated(n, ratio):
    """(candidates.len() as f32 * rescore_ratio) as usize, then .min(len) — src/quantization.rs:178-179;
    Rust's float -> usize cast saturates (negative -> 0)."""
    p = F(F(n) * F(ratio))
    want = min(int(p) if p > 0 else 0, n)
    assert oracle.rescore_count(n, float(F(ratio))) == want


@settings(max_examples=40, deadline=None)
@given(data=st.data(), n=st.integers(1, 40), dim=st.integers(1, 20),
       ratio=st.sampled_from([0.0, 0.1, 0.34, 0.5, 1.0, 1.5]),
       thr=st.sampled_from([0.0, -0.0, 0.5, -1.0]))
def test_two_stage_on_tie_heavy_rows(data, n, dim, ratio, thr):
    """Values from a three-letter alphabet: many equal Hamming distances and equal cosines — both stable sorts
    (src/quantization.rs:175,190) decide the order."""
    vals = st.sampled_from([-1.0, 0.0, 2.0])
    rows = np.asarray(data.draw(hnp.arrays(np.float32, (n, dim), elements=vals)))
    q = np.asarray(data.draw(hnp.arrays(np.float32, (dim,), elements=vals)))
    with np.errstate(all="ignore"):
        want = py_two_stage(q, rows, ratio, thr)
    r = len(want)
    assert r == oracle.rescore_count(n, ratio)
    if r == 0:
        return
    ids, sc = oracle.multi_stage_search(q, rows, r, thr)
    assert [int(i) for i in ids] == [i for i, _ in want]
    assert np.array_equal(_bits(sc), _bits([s for _, s in want]))


@settings(max_examples=40, deadline=None)
@given(data=st.data(), n=st.integers(1, 30), dim=st.integers(1, 12), k=st.integers(1, 35))
def test_flat_search_with_tombstones_and_zero_rows(data, n, dim, k):
    vals = st.sampled_from([-2.0, 0.0, 0.0, 1.0, 3.0])
    rows = np.asarray(data.draw(hnp.arrays(np.float32, (n, dim), elements=vals)))
    q = np.asarray(data.draw(hnp.arrays(np.float32, (dim,), elements=vals)))
    live = np.asarray(data.draw(hnp.arrays(np.bool_, (n,))))
    with np.errstate(all="ignore"):
        want = py_flat(q, rows, live, k)
    ids, d = oracle.flat_search(q, rows, k, live=live)
    assert [int(i) for i in ids] == [i for i, _ in want]          # min(k, live rows) answers, (distance, row) ascending
    assert np.array_equal(_bits(d), _bits([x for _, x in want]))


# ---- the hybrid side: RRF, weighted fusions, BM25, and the cross-shard merge ----------------------------------------
from test_oracle_independent import py_bm25, py_rrf, py_weighted      # noqa: E402

DOC_LIST = st.lists(st.integers(0, 25), max_size=14)        # duplicates allowed: an id may sit twice in one list


@settings(max_examples=120, deadline=None)
@given(dense=DOC_LIST, sparse=DOC_LIST, text=DOC_LIST, k=st.sampled_from([60.0, 1.0, 0.0, 1000.0]))
def test_rrf_any_lists(dense, sparse, text, k):
    """src/hybrid.rs:422-488: the dense list INSERTS (a repeated id keeps its last rank), the other lists add;
    ties in first-appearance order (the oracle's statement of the reference's unspecified HashMap order)."""
    want = py_rrf(dense, sparse, text, k)
    ids, sc = oracle.rrf_fusion(dense, sparse, text, k)
    assert [int(i) for i in ids] == [i for i, _ in want]
    assert np.array_equal(_bits(sc), _bits([x for _, x in want]))


SCORED = st.lists(st.tuples(st.integers(0, 25), st.sampled_from([-2.0, -0.25, 0.0, 0.5, 0.5, 1.0, 3.75])), max_size=12)


@settings(max_examples=120, deadline=None)
@given(dense=SCORED, sparse=SCORED, text=SCORED, normalize=st.booleans(),
       weights=st.sampled_from([(0.7, 0.2, 0.1), (1.0, 0.0, 0.0), (0.5, 0.5, 0.0), (0.2, 0.3, 0.5)]))
def test_weighted_fusion_any_lists(dense, sparse, text, normalize, weights):
    """linear_fusion / normalized_fusion (src/hybrid.rs:491-616): constant lists normalise to 1.0, empty lists vanish."""
    with np.errstate(all="ignore"):
        want = py_weighted(dense, sparse, text, weights, normalize)
    ids, sc = oracle.weighted_fusion([d for d, _ in dense], [s for _, s in dense], [d for d, _ in sparse],
                                     [s for _, s in sparse], [d for d, _ in text], [s for _, s in text], weights, normalize)
    assert [int(i) for i in ids] == [i for i, _ in want]
    assert np.array_equal(_bits(sc), _bits([x for _, x in want]))


@settings(max_examples=60, deadline=None)
@given(data=st.data(), n_docs=st.integers(1, 30), vocab=st.integers(1, 8), limit=st.integers(1, 40))
def test_bm25_any_small_index(data, n_docs, vocab, limit):
    """src/sparse.rs:153-222 incl. negative idf (df > N/2), terms outside the vocabulary, repeated query terms and the
    average length taken per POSTING (:96-104)."""
    eighths = st.integers(1, 8).map(lambda v: v / 8)
    doc_len = np.asarray(data.draw(st.lists(eighths, min_size=n_docs, max_size=n_docs)), dtype=np.float32)
    postings, post_off, post_doc, post_tf = {}, [0], [], []
    for t in range(vocab):
        docs = sorted(data.draw(st.sets(st.integers(0, n_docs - 1), max_size=n_docs)))
        tfs = data.draw(st.lists(eighths, min_size=len(docs), max_size=len(docs)))
        postings[t] = list(zip(docs, tfs))
        post_doc += docs
        post_tf += tfs
        post_off.append(len(post_doc))
    if not post_doc:
        return
    total = F(0)
    for t in range(vocab):
        for doc, _ in postings[t]:
            total = F(total + doc_len[doc])
    avg = F(total / F(n_docs))
    assert _bits([avg])[0] == _bits([oracle.bm25_avg_len(post_off, post_doc, doc_len)])[0]
    qt = data.draw(st.lists(st.integers(0, vocab + 1), min_size=1, max_size=5))
    qv = np.asarray(data.draw(st.lists(st.integers(1, 4).map(lambda v: v / 4), min_size=len(qt), max_size=len(qt))),
                    dtype=np.float32)
    with np.errstate(all="ignore"):
        want = py_bm25(qt, qv, postings, doc_len, n_docs, avg, limit)
    docs, sc = oracle.bm25_search(qt, qv, post_off, post_doc, post_tf, doc_len, limit)
    assert [int(d) for d in docs] == [d for d, _ in want]
    assert np.array_equal(_bits(sc), _bits([s for _, s in want]))


@settings(max_examples=40, deadline=None)
@given(data=st.data(), n=st.integers(2, 48), dim=st.integers(1, 16), shards=st.integers(1, 5),
       R=st.integers(1, 12), k=st.integers(1, 12))
def test_shard_merge_equals_the_single_index(data, n, dim, shards, R, k):
    """per-shard top R by (hamming, row) + cosine, merged by the global stage-1 cut and the final order
    (cos desc, ham asc, row asc), equals multi_stage_search over the whole corpus — for any contiguous split."""
    vals = st.sampled_from([-1.0, 0.0, 2.0, 2.0])
    rows = np.asarray(data.draw(hnp.arrays(np.float32, (n, dim), elements=vals)))
    q = np.asarray(data.draw(hnp.arrays(np.float32, (dim,), elements=vals)))
    R = min(R, n)
    ids, sc = oracle.multi_stage_search(q, rows, R)
    want_ids, want_sc = np.asarray(ids)[:k], np.asarray(sc)[:k]
    qc = oracle.quantize(q)
    per = (n + shards - 1) // shards
    ham, idx, score = [], [], []
    for s_ in range(shards):
        lo, hi = min(n, s_ * per), min(n, (s_ + 1) * per)
        if lo == hi:
            continue
        li, ls = oracle.multi_stage_search(q, rows[lo:hi], min(R, hi - lo))
        for i, c in zip(np.asarray(li), np.asarray(ls)):
            g = lo + int(i)
            ham.append(oracle.hamming(qc, oracle.quantize(rows[g])))
            idx.append(g)
            score.append(c)
    mi, ms = oracle.shard_merge(ham, idx, score, R, k)
    assert [int(i) for i in mi] == [int(i) for i in want_ids]
    assert np.array_equal(_bits(ms), _bits(want_sc))
