#!/usr/bin/env python
"""Full-size check of one C5 shard (BASELINE configs[4]: 100M x 768 over 8 GPUs = 12.5M rows per GPU):
builds the shard on cuda:0, runs a 1024-query batch (R = 40, k = 10) and verifies a sample of the
queries against the CPU oracle WITHOUT holding the 38 GB of f32 rows on the host:
  * stage 1: the GPU's candidate list == the R smallest (hamming, row) keys of oracle.hamming_all
    over the shard's codes (read back with gvdb_get_codes);
  * stage 2: each candidate's score == oracle.cosine_similarity on that row (regenerated on the host);
  * order: ids/scores == candidates ordered by (cosine desc, stage-1 position).
usage: python tools/fullsize_check.py [rows] [dim] [nq] [checked]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import grape_vector_db_b200 as gv
from grape_vector_db_b200 import synth
from oracle import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
checked = int(sys.argv[4]) if len(sys.argv) > 4 else 6
k, R = 10, 40
dev = torch.device("cuda", 0)
t0 = time.time()
idx = gv.GpuIndex(dim, device=0, capacity_rows=n)
for i in range(0, n, 262144):
    idx.add_device(synth.lowrank_rows_torch(i, min(262144, n - i), dim, dev))
print(f"built {n} x {dim} in {time.time() - t0:.1f} s; HBM {idx.stats()['hbm_bytes'] / 2**30:.1f} GiB", flush=True)
qs = synth.lowrank_queries(0, nq, dim)
q_t = torch.from_numpy(qs).to(dev)
ids_t, sc_t = idx.search_batch_device(q_t, k, R)       # warm-up
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ids_t, sc_t = idx.search_batch_device(q_t, k, R)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"batch of {nq}: {ms:.2f} ms -> {nq / ms * 1e3:.0f} QPS on this shard; reruns {idx.profile_read()['optimistic_reruns']}", flush=True)
ids, sc, ci, ch = idx.search_batch(qs[:checked], k, R, want_candidates=True)
assert np.array_equal(ids, ids_t[:checked].cpu().numpy().astype(np.uint64)) and np.array_equal(sc.view(np.uint32), sc_t[:checked].cpu().numpy().view(np.uint32))
codes = idx.get_codes()
bad = 0
for qi in range(checked):
    qc = oracle.quantize(qs[qi])
    if codes.shape[1] % 8 == 0:      # popcount of the xor, 64 bits at a time (same integers as oracle.hamming_all)
        ham = np.bitwise_count(np.bitwise_xor(codes.view(np.uint64), qc.view(np.uint64)[None, :])).sum(axis=1, dtype=np.int64)
        assert np.array_equal(ham[:1000], oracle.hamming_all(qc, codes[:1000]).astype(np.int64))
    else:
        ham = oracle.hamming_all(qc, codes).astype(np.int64)
    key = (ham << 32) | np.arange(n, dtype=np.int64)
    order = np.sort(np.partition(key, R)[:R]) & 0xFFFFFFFF
    ok1 = np.array_equal(ci[qi], order.astype(np.uint64)) and np.array_equal(ch[qi], ham[order].astype(np.uint32))
    cos = np.array([oracle.cosine_similarity(qs[qi], synth.lowrank_rows(int(r), 1, dim)[0]) for r in order], dtype=np.float32)
    fin = np.argsort(-cos, kind="stable")[:k]
    ok2 = np.array_equal(ids[qi], order[fin].astype(np.uint64)) and np.array_equal(sc[qi].view(np.uint32), cos[fin].view(np.uint32))
    print(f"query {qi}: stage-1 list {'exact' if ok1 else 'DIFFERS'}, top-{k} ids+scores {'exact' if ok2 else 'DIFFER'}", flush=True)
    bad += (not ok1) + (not ok2)
print("FULL-SIZE CHECK", "PASSED" if bad == 0 else "FAILED")
sys.exit(1 if bad else 0)
