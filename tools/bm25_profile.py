"""One chunk of BM25 queries on the configs[3] postings, for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bm25_launches.csv \
        python tools/bm25_profile.py --docs 5000000 --queries 25"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=5_000_000)
    ap.add_argument("--vocab", type=int, default=100_000)
    ap.add_argument("--queries", type=int, default=25)
    ap.add_argument("--limit", type=int, default=200)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--bps-list", default="", help="comma list of GVDB_BM25_BPS values to time one after another")
    args = ap.parse_args()
    import grape_vector_db_b200 as gv
    from grape_vector_db_b200 import synth
    post = synth.sparse_corpus(args.docs, vocab=args.vocab)
    qs = synth.sparse_queries(args.queries, vocab=args.vocab)
    with gv.GpuSparseIndex() as sp:
        sp.build(*post)
        import time
        for bps in [b for b in args.bps_list.split(",") if b]:
            os.environ["GVDB_BM25_BPS"] = bps
            for _ in range(args.reps):
                t0 = time.perf_counter()
                docs, sc = sp.search_bm25_batch(qs, args.limit)
                print(f"bps {bps}: bm25 {args.queries} queries: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
        os.environ.pop("GVDB_BM25_BPS", None)
        for _ in range(args.reps):
            t0 = time.perf_counter()
            docs, sc = sp.search_bm25_batch(qs, args.limit)
            print(f"bm25 {args.queries} queries: {1e3 * (time.perf_counter() - t0):.2f} ms (host arrays in and out)", flush=True)
    print("ok", int((docs[:, 0] != gv.NO_ID).sum()), "queries answered")


if __name__ == "__main__":
    main()
