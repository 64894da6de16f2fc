"""Generates tests/golden/kat_small.json from the CPU oracle on seeded synthetic inputs.

The reference (Rust) cannot run in this image, so these are ORACLE outputs, frozen so the
oracle and the CUDA path are both checked against the same committed numbers.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from grape_vector_db_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

cases = []
for dataset, n, dim, nq, R, k in (("lowrank", 2000, 768, 4, 40, 10), ("iid", 1500, 100, 3, 150, 20),
                                  ("lowrank", 700, 1536, 2, 70, 70)):
    gen = synth.lowrank_rows if dataset == "lowrank" else synth.iid_rows
    genq = synth.lowrank_queries if dataset == "lowrank" else synth.iid_queries
    rows, qs = gen(0, n, dim), genq(0, nq, dim)
    codes = oracle.quantize_batch(rows)
    ids, bits = [], []
    for qi in range(nq):
        idx, sc = oracle.multi_stage_search(qs[qi], rows, R, codes=codes)
        ids.append(idx[:k].tolist())
        bits.append(sc[:k].view(np.uint32).tolist())
    cases.append(dict(dataset=dataset, n=n, dim=dim, nq=nq, R=R, k=k,
                      codes_head=[int(x) for x in codes[:4].ravel()[:32]], ids=ids, score_bits=bits))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_small.json")
json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/gvdb_oracle.cpp",
           "cases": cases}, open(out, "w"))
print("wrote", out)

# ---- sparse side, fusion, filtered search (kat_hybrid.json) ----------------------------------------
n, dim, vocab, nq, limit = 3000, 256, 400, 6, 12
rows, qs = synth.lowrank_rows(0, n, dim), synth.lowrank_queries(0, nq, dim)
post = synth.sparse_corpus(n, vocab=vocab)
sq = synth.sparse_queries(nq, vocab=vocab)
want = 2 * limit
dense_ids, _ = oracle.multi_stage_search_batch(qs, rows, want * 4, want)
bm25, fused = [], []
for q in range(nq):
    d, s = oracle.bm25_search(sq[q][0], sq[q][1], *post, want)
    bm25.append(dict(docs=d.tolist(), score_bits=s.view(np.uint32).tolist()))
    fi, fs = oracle.rrf_fusion(dense_ids[q], d, [], 60.0)
    fused.append(dict(ids=fi[:limit].tolist(), score_bits=fs[:limit].view(np.uint32).tolist()))
allow = (np.arange(n) % 7) < 2                                      # rows 0, 1 of every 7
sub = np.flatnonzero(allow)
fi, fs = oracle.multi_stage_search_batch(qs, rows[sub], 40, 10)
hy = dict(n=n, dim=dim, vocab=vocab, nq=nq, limit=limit, oversample=4, rrf_k=60.0,
          avg_len_bits=int(np.float32(oracle.bm25_avg_len(post[0], post[1], post[3])).view(np.uint32)),
          postings=int(post[0][-1]), bm25=bm25, fused=fused,
          filtered=dict(rule="row % 7 < 2", R=40, k=10, ids=sub[fi.astype(np.int64)].tolist(),
                        score_bits=fs.view(np.uint32).tolist()))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_hybrid.json")
json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/gvdb_oracle.cpp", "hybrid": hy}, open(out, "w"))
print("wrote", out)
