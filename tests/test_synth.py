"""The synthetic corpora of bench.py exist twice: a torch generator (rows made on the GPU, nothing travels) and a
numpy one (the oracle check regenerates single rows on the host).  The parity verdicts of bench.py rest on the two
agreeing bit for bit at any offset; checked here on the CPU device.  Also: the generators are functions of the row
number alone (any split of a corpus into chunks or shards gives the same rows), and integer-exact (every value is a
small integer, so no summation order can change it — SURVEY.md §8d)."""
import numpy as np
import pytest

from grape_vector_db_b200 import synth

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("first,n,dim", [(0, 257, 768), (999_983, 64, 1536), (12_499_990, 33, 768), (3, 5, 13)])
def test_torch_and_numpy_generators_agree_bit_for_bit(first, n, dim):
    dev = torch.device("cpu")
    a = synth.lowrank_rows(first, n, dim)
    b = synth.lowrank_rows_torch(first, n, dim, dev).numpy()
    assert a.dtype == b.dtype == np.float32 and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    qa = synth.lowrank_queries(first, n, dim)
    qb = synth.lowrank_queries_torch(first, n, dim, dev).numpy()
    assert np.array_equal(qa.view(np.uint32), qb.view(np.uint32))
    ia = synth.iid_rows(first, n, dim)
    ib = synth.iid_rows_torch(first, n, dim, dev).numpy()
    assert np.array_equal(ia.view(np.uint32), ib.view(np.uint32))
    assert not np.array_equal(a[:min(n, 5)], qa[:min(n, 5)])          # rows and queries come from separate streams


def test_rows_depend_on_the_row_number_only():
    whole = synth.lowrank_rows(1000, 300, 96)
    parts = np.concatenate([synth.lowrank_rows(1000, 7, 96), synth.lowrank_rows(1007, 193, 96), synth.lowrank_rows(1200, 100, 96)])
    assert np.array_equal(whole, parts)
    assert np.array_equal(synth.iid_rows(50, 20, 40), np.concatenate([synth.iid_rows(50, 1, 40), synth.iid_rows(51, 19, 40)]))


def test_values_are_small_integers():
    for x in (synth.lowrank_rows(0, 400, 768), synth.iid_rows(0, 400, 768), synth.lowrank_queries(0, 50, 1536)):
        assert np.array_equal(x, np.rint(x)) and float(np.abs(x).max()) < 2.0 ** 24
        assert 0.35 < float((x > 0).mean()) < 0.65                      # centred: the 0.0 threshold splits the bits evenly


def test_sparse_corpus_is_csr_by_term_with_sorted_documents():
    post_off, post_doc, post_tf, doc_len = synth.sparse_corpus(500, vocab=300)
    assert post_off[0] == 0 and post_off[-1] == post_doc.size == post_tf.size and np.all(np.diff(post_off.astype(np.int64)) >= 0)
    for t in range(300):
        d = post_doc[int(post_off[t]):int(post_off[t + 1])].astype(np.int64)
        assert np.all(np.diff(d) > 0)                                   # one posting per (term, document), ascending
    # document_length = sum of the document's tfs (src/sparse.rs:341)
    acc = np.zeros(500, dtype=np.float64)
    np.add.at(acc, post_doc.astype(np.int64), post_tf.astype(np.float64))
    assert np.allclose(acc, doc_len, rtol=1e-5)
    qs = synth.sparse_queries(9, vocab=300)
    assert len(qs) == 9 and all(len(t) == len(w) and len(t) >= 1 for t, w in qs)
