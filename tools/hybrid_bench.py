"""BASELINE configs[3] on one B200: hybrid dense + sparse search, 5M x 768 dense vectors + Zipf
postings (vocabulary 100k, 32 tokens per document, 4-term queries), RRF fusion, top-100
(SURVEY.md §8d C4).  Times the three GPU stages of HybridSearcher (dense two-stage list of 200,
BM25 list of 200, rrf_fusion) with CUDA events, checks a sample of queries against the oracle
composition, and times the oracle's BM25 + RRF on the host cores beside it.  One JSON line.

    python tools/hybrid_bench.py [--docs 5000000] [--batch 1024] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=5_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--vocab", type=int, default=100_000)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--limit", type=int, default=100)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", type=int, default=16)
    args = ap.parse_args()
    import torch
    import grape_vector_db_b200 as gv
    from grape_vector_db_b200 import synth
    dev = torch.device("cuda:0")
    n, dim, limit, want = args.docs, args.dim, args.limit, 2 * args.limit

    t0 = time.perf_counter()
    post = synth.sparse_corpus(n, vocab=args.vocab)
    t_gen = time.perf_counter() - t0
    sparse = gv.GpuSparseIndex()
    sparse.build(*post)
    dense = gv.GpuIndex(dim, capacity_rows=n)
    for i in range(0, n, 131072):
        dense.add_device(synth.lowrank_rows_torch(i, min(131072, n - i), dim, dev))
    hy = gv.HybridSearcher(dense, sparse, rrf_k=60.0, oversample=4)
    NB = 2
    q_dense = [synth.lowrank_queries_torch(b * args.batch, args.batch, dim, dev) for b in range(NB)]
    q_sparse_all = synth.sparse_queries(NB * args.batch, vocab=args.vocab)
    q_sparse = [q_sparse_all[b * args.batch:(b + 1) * args.batch] for b in range(NB)]

    def timed(fn, reps):
        fn(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            fn(r % NB)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_dense = timed(lambda b: dense.search_batch_device(q_dense[b], want, want * 4), args.steps)
    l0 = sparse.launches
    ms_bm25 = timed(lambda b: sparse.search_bm25_batch_device(q_sparse[b], want), args.steps)
    bm25_launches = (sparse.launches - l0) // (args.steps + 1)
    ms_hybrid = timed(lambda b: hy.search_batch_device(q_dense[b], q_sparse[b], limit), args.steps)
    t0 = time.perf_counter()
    for r in range(args.steps):
        hy.search_batch(q_dense[r % NB].cpu().numpy(), q_sparse[r % NB], limit)
    ms_e2e = 1e3 * (time.perf_counter() - t0) / args.steps

    # parity on a sample + the host-side BM25/RRF timed beside it (the dense oracle is timed by bench.py)
    from oracle import oracle
    ids, sc = hy.search_batch(q_dense[0][:args.check].cpu().numpy(), q_sparse[0][:args.check], limit)
    d_ids, _ = dense.search_batch_device(q_dense[0][:args.check].contiguous(), want, want * 4)
    d_ids = d_ids.cpu().numpy().astype(np.uint64)
    t0 = time.perf_counter()
    ok = True
    for q in range(args.check):
        ob, _ = oracle.bm25_search(q_sparse[0][q][0], q_sparse[0][q][1], *post, want)
        oi, os_ = oracle.rrf_fusion(d_ids[q], ob, [], 60.0)
        ok &= bool(np.array_equal(ids[q], oi[:limit]) and np.array_equal(sc[q].view(np.uint32), os_[:limit].view(np.uint32)))
    cpu_ms_per_query = 1e3 * (time.perf_counter() - t0) / args.check
    postings_per_query = float(np.mean([sum(int(post[0][t + 1] - post[0][t]) for t in q[0]) for q in q_sparse[0]]))
    print(json.dumps({
        "workload": f"configs[3]: hybrid dense+sparse, {n}x{dim} dense + {int(post[0][-1])} postings (vocab {args.vocab}), "
                    f"batch {args.batch}, RRF k=60, top-{limit} (lists of {want})",
        "ms_per_batch": {"dense_two_stage_top200_R800": ms_dense, "bm25_top200": ms_bm25, "hybrid_all_three_stages": ms_hybrid,
                         "hybrid_host_arrays_in_out": ms_e2e},
        "hybrid_qps_device": args.batch / (ms_hybrid * 1e-3), "hybrid_qps_e2e": args.batch / (ms_e2e * 1e-3),
        "bm25_qps": args.batch / (ms_bm25 * 1e-3), "bm25_kernel_launches_per_batch": int(bm25_launches),
        "bm25_postings_per_query": postings_per_query,
        "bm25_effective_GBps": postings_per_query * 8 * args.batch / (ms_bm25 * 1e-3) / 1e9,
        "fused_lists_bit_exact_vs_oracle_bm25_rrf": ok, "checked_queries": args.check,
        "cpu_oracle_bm25_rrf_ms_per_query_1_thread": cpu_ms_per_query,
        "sparse_corpus_generation_s": t_gen,
    }), flush=True)


if __name__ == "__main__":
    main()
