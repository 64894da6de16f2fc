#!/usr/bin/env python
"""QPS and recall@10 against the exact f32 flat search as the oversampling factor grows (SURVEY.md §8d:
"the harness sweeps oversample in {4, 8, 16, 32, 64} and reports QPS at the smallest value reaching
>= 0.95 and at the config's stated 4x").  One GPU, one shard, batches of 1024 queries, k = 10.
usage: python tools/recall_sweep.py [rows] [dim] [recall_queries]      -> one JSON line"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import grape_vector_db_b200 as gv
from grape_vector_db_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
nr = int(sys.argv[3]) if len(sys.argv) > 3 else 64
nq, k = 1024, 10
dev = torch.device("cuda", 0)
t0 = time.time()
idx = gv.GpuIndex(dim, device=0, capacity_rows=n)
for i in range(0, n, 262144):
    idx.add_device(synth.lowrank_rows_torch(i, min(262144, n - i), dim, dev))
build_s = time.time() - t0
q_t = synth.lowrank_queries_torch(0, nq, dim, dev)
exact, _ = idx.flat_search_batch_device(q_t[:nr].contiguous(), k)
exact = exact.cpu().numpy()
out = {}
for ov in (4, 8, 16, 32, 64):
    R = k * ov
    ids, _ = idx.search_batch_device(q_t, k, R)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ids, _ = idx.search_batch_device(q_t, k, R)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    got = ids[:nr].cpu().numpy()
    recall = float(np.mean([len(set(got[i]) & set(exact[i])) / k for i in range(nr)]))
    out[f"x{ov}"] = {"rescore_count": R, "recall_at_10": recall, "ms_per_batch": ms, "qps": nq / ms * 1e3}
ok = [v for v in out.values() if v["recall_at_10"] >= 0.95]
print(json.dumps({"rows": n, "dim": dim, "batch": nq, "k": k, "recall_queries": nr, "build_s": build_s,
                  "dataset": "lowrank L=16 integer-exact, seed 42", "by_oversample": out,
                  "qps_at_recall_0.95": max(ok, key=lambda v: v["qps"]) if ok else None,
                  "optimistic_reruns": idx.profile_read()["optimistic_reruns"]}), flush=True)
