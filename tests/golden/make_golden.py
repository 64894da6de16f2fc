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
