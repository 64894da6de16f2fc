"""A second, independent restatement of the reference's arithmetic — plain Python loops over
numpy.float32 scalars, written from the cited lines without looking at
oracle/gvdb_oracle.cpp — cross-checked against the C++ oracle on small random cases.  Two
restatements that agree bit for bit make a transcription error in either one unlikely; the
reference itself cannot run here (Rust, no toolchain)."""
import math

import numpy as np

from oracle import oracle

F = np.float32


def py_quantize(x, thr):                       # src/quantization.rs:97-101, BitVec<u8, Msb0>
    out = bytearray((len(x) + 7) // 8)
    for j, v in enumerate(x):
        if v > thr:                            # strict; NaN compares false
            out[j // 8] |= 1 << (7 - j % 8)
    return bytes(out)


def py_hamming(a, b):                          # :130-141 (hamming::distance = differing bits)
    return sum(bin(p ^ q).count("1") for p, q in zip(a, b))


def py_cosine(q, c):                           # :206-216: three separate left-to-right f32 folds
    dot, nq, nc = F(0), F(0), F(0)
    for a, b in zip(q, c):
        dot = F(dot + F(a * b))
    for a in q:
        nq = F(nq + F(a * a))
    for b in c:
        nc = F(nc + F(b * b))
    nq, nc = F(math.sqrt(nq)), F(math.sqrt(nc))          # sqrt of an f32 is correctly rounded either way
    if nq == 0 or nc == 0:
        return F(0)
    return F(dot / F(nq * nc))


def py_two_stage(q, rows, ratio, thr=0.0):     # :151-193
    n, dim = rows.shape
    qc = py_quantize(q, thr)
    s1 = []
    for i in range(n):
        d = py_hamming(qc, py_quantize(rows[i], thr))
        s1.append((i, F(F(1.0) - F(F(d) / F(dim)))))                    # :145-147
    s1.sort(key=lambda t: -t[1])                                         # stable, descending (:175)
    r = min(int(F(F(n) * F(ratio))), n)                                  # :178-179, f32 product, truncation
    s2 = [(i, py_cosine(q, rows[i])) for i, _ in s1[:r]]
    s2.sort(key=lambda t: -t[1])                                         # stable (:190)
    return s2


def py_flat(q, rows, live, k):                 # src/index.rs:620-640, 686-700
    out = []
    for i in range(rows.shape[0]):
        if not live[i]:
            continue
        dot, nq, nc = F(0), F(0), F(0)
        for a, b in zip(q, rows[i]):
            dot = F(dot + F(a * b))
        for a in q:
            nq = F(nq + F(a * a))
        for b in rows[i]:
            nc = F(nc + F(b * b))
        nq, nc = F(math.sqrt(nq)), F(math.sqrt(nc))
        d = F(np.inf) if (nq == 0 or nc == 0) else F(F(1.0) - F(dot / F(nq * nc)))
        out.append((i, d))
    out.sort(key=lambda t: t[1])               # stable, ascending
    return out[:k]


def py_bm25(q_terms, q_tfs, postings, doc_len, n_docs, avg, limit, k1=F(1.2), b=F(0.75)):   # src/sparse.rs:153-222
    acc, order = {}, []
    for t, qtf in zip(q_terms, q_tfs):
        plist = postings.get(int(t))
        if not plist:
            continue
        df = len(plist)
        idf = F(math.log(F(F(F(n_docs) - F(df)) + F(0.5)) / F(F(df) + F(0.5))))   # f32 division, then ln
        for doc, tf in plist:
            tf = F(tf)
            denom = F(tf + F(k1 * F(F(F(1.0) - b) + F(b * F(doc_len[doc] / avg)))))
            tfc = F(F(tf * F(k1 + F(1.0))) / denom)
            s = F(F(F(qtf) * tfc) * idf)
            if doc not in acc:
                acc[doc] = F(F(0) + s)
                order.append(doc)
            else:
                acc[doc] = F(acc[doc] + s)
    docs = sorted(acc, key=lambda d: (-acc[d], d))       # ties (unspecified in the reference): document number
    return [(d, acc[d]) for d in docs[:limit]]


def py_rrf(dense, sparse, text, k):            # src/hybrid.rs:422-488
    score, first = {}, []
    for rank, d in enumerate(dense):
        if d not in score:
            first.append(d)
        score[d] = F(F(1.0) / F(F(k) + F(rank + 1)))      # insert: overwrites
    for lst in (sparse, text):
        for rank, d in enumerate(lst):
            r = F(F(1.0) / F(F(k) + F(rank + 1)))
            if d in score:
                score[d] = F(score[d] + r)
            else:
                score[d] = r
                first.append(d)
    order = sorted(range(len(first)), key=lambda i: -score[first[i]])    # stable: ties by first appearance
    return [(first[i], score[first[i]]) for i in order]


def _bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def test_two_stage_matches_independent_restatement():
    rng = np.random.default_rng(7)
    for n, dim, ratio, thr in ((60, 24, 0.5, 0.0), (37, 13, 0.3, 0.25), (80, 8, 1.0, 0.0), (25, 40, 0.1, -0.5)):
        rows = rng.integers(-3, 4, size=(n, dim)).astype(np.float32)     # small integers: many Hamming and cosine ties
        rows[3] = 0
        qs = rng.integers(-3, 4, size=(4, dim)).astype(np.float32)
        for q in qs:
            want = py_two_stage(q, rows, ratio, thr)
            r = len(want)
            assert r == oracle.rescore_count(n, ratio)
            ids, sc = oracle.multi_stage_search(q, rows, r, thr)
            assert [int(i) for i in ids] == [i for i, _ in want]
            assert np.array_equal(_bits(sc), _bits([s for _, s in want]))
        codes = oracle.quantize_batch(rows, thr)
        for i in range(n):
            assert bytes(codes[i]) == py_quantize(rows[i], thr)


def test_flat_search_matches_independent_restatement():
    rng = np.random.default_rng(8)
    rows = (rng.standard_normal((50, 16)) * 3).astype(np.float32)
    rows[7] = 0
    live = np.ones(50, dtype=bool)
    live[[2, 9]] = False
    for q in (rng.standard_normal((3, 16)) * 2).astype(np.float32):
        want = py_flat(q, rows, live, 12)
        ids, d = oracle.flat_search(q, rows, 12, live=live)
        assert [int(i) for i in ids] == [i for i, _ in want]
        assert np.array_equal(_bits(d), _bits([x for _, x in want]))


def test_bm25_and_rrf_match_independent_restatement():
    rng = np.random.default_rng(9)
    n_docs, vocab = 40, 12
    postings, post_off, post_doc, post_tf = {}, [0], [], []
    doc_len = (rng.integers(1, 9, size=n_docs) / 8).astype(np.float32)
    for t in range(vocab):
        docs = np.sort(rng.choice(n_docs, size=int(rng.integers(0, 30)), replace=False))
        tfs = (rng.integers(1, 9, size=docs.size) / 8).astype(np.float32)
        postings[t] = list(zip(docs.tolist(), tfs.tolist()))
        post_doc += docs.tolist()
        post_tf += tfs.tolist()
        post_off.append(len(post_doc))
    total = F(0)
    for t in range(vocab):                                   # src/sparse.rs:96-104, in term order
        for doc, _ in postings[t]:
            total = F(total + doc_len[doc])
    avg = F(total / F(n_docs))
    assert _bits([avg])[0] == _bits([oracle.bm25_avg_len(post_off, post_doc, doc_len)])[0]
    for _ in range(6):
        qt = rng.integers(0, vocab + 2, size=int(rng.integers(1, 5)))
        qv = (rng.integers(1, 5, size=qt.size) / 4).astype(np.float32)
        want = py_bm25(qt, qv, postings, doc_len, n_docs, avg, 15)
        docs, sc = oracle.bm25_search(qt, qv, post_off, post_doc, post_tf, doc_len, 15)
        assert [int(d) for d in docs] == [d for d, _ in want]
        assert np.array_equal(_bits(sc), _bits([s for _, s in want]))
    for _ in range(5):
        d = rng.choice(30, size=12, replace=False).tolist()
        s = rng.choice(30, size=10, replace=False).tolist()
        t = rng.choice(30, size=int(rng.integers(0, 8)), replace=False).tolist()
        want = py_rrf(d, s, t, 60.0)
        ids, sc = oracle.rrf_fusion(d, s, t, 60.0)
        assert [int(i) for i in ids] == [i for i, _ in want]
        assert np.array_equal(_bits(sc), _bits([x for _, x in want]))


def py_normalize(scores):                       # src/hybrid.rs:589-616
    if not scores:
        return []
    mx, mn = F(-np.inf), F(np.inf)
    for s in scores:
        mx = s if s > mx else mx                 # f32::max / f32::min
        mn = s if s < mn else mn
    rng_ = F(mx - mn)
    return [F(F(s - mn) / rng_) if rng_ > 0 else F(1.0) for s in scores]


def py_weighted(dense, sparse, text, weights, normalize):    # src/hybrid.rs:491-587; lists of (doc, score)
    lists = []
    for lst in (dense, sparse, text):
        sc = [F(s) for _, s in lst]
        if normalize:
            sc = py_normalize(sc)
        lists.append([(d, s) for (d, _), s in zip(lst, sc)])
    score, first = {}, []
    for d, s in lists[0]:
        if d not in score:
            first.append(d)
        score[d] = F(s * F(weights[0]))          # insert: overwrites
    for li in (1, 2):
        for d, s in lists[li]:
            w = F(s * F(weights[li]))
            if d in score:
                score[d] = F(score[d] + w)
            else:
                score[d] = w
                first.append(d)
    order = sorted(range(len(first)), key=lambda i: -score[first[i]])    # stable: ties by first appearance
    return [(first[i], score[first[i]]) for i in order]


def test_weighted_fusions_match_independent_restatement_and_known_values():
    # known answers: dense [(1, .9), (2, .5)], sparse [(2, 4.0), (3, 1.0)], weights .7 / .2 / .1
    ids, sc = oracle.weighted_fusion([1, 2], [0.9, 0.5], [2, 3], [4.0, 1.0], [], [], (0.7, 0.2, 0.1), False)
    want = {1: F(F(0.9) * F(0.7)), 2: F(F(F(0.5) * F(0.7)) + F(F(4.0) * F(0.2))), 3: F(F(1.0) * F(0.2))}
    assert [int(i) for i in ids] == [2, 1, 3]
    assert np.array_equal(_bits(sc), _bits([want[2], want[1], want[3]]))
    # normalized: every list mapped to [0, 1] first — dense (1.0, 0.0), sparse (1.0, 0.0); a one-entry list becomes 1.0
    ids, sc = oracle.weighted_fusion([1, 2], [0.9, 0.5], [2, 3], [4.0, 1.0], [7], [123.0], (0.7, 0.2, 0.1), True)
    assert [int(i) for i in ids] == [1, 2, 7, 3]
    assert np.array_equal(_bits(sc), _bits([F(0.7), F(F(0.0) + F(0.2)), F(0.1), F(0.0)]))
    rng = np.random.default_rng(13)
    for trial in range(8):
        mk = lambda n: list(zip(rng.choice(30, size=n, replace=trial % 2 == 1).tolist(),
                                (rng.integers(-8, 9, size=n) / 4).astype(np.float32).tolist()))
        d, s, t = mk(12), mk(10), mk(int(rng.integers(0, 8)))
        for normalize in (False, True):
            want = py_weighted(d, s, t, (0.7, 0.2, 0.1), normalize)
            ids, sc = oracle.weighted_fusion([x for x, _ in d], [y for _, y in d], [x for x, _ in s], [y for _, y in s],
                                             [x for x, _ in t], [y for _, y in t], (0.7, 0.2, 0.1), normalize)
            assert [int(i) for i in ids] == [i for i, _ in want]
            assert np.array_equal(_bits(sc), _bits([x for _, x in want]))
