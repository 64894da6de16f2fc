#!/usr/bin/env python
"""Throughput with T host threads, each issuing its own batches on its own CUDA stream
(gvdb_search_batch_device is re-entrant: per-call workspace + stream).  Prints QPS per T."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import grape_vector_db_b200 as gv
from grape_vector_db_b200 import synth

dev = torch.device("cuda", 0)
n, dim, B, k, R = 1_000_000, 768, 1024, 10, 40
idx = gv.GpuIndex(dim, device=0, capacity_rows=n)
for i in range(0, n, 131072):
    idx.add_device(synth.lowrank_rows_torch(i, min(131072, n - i), dim, dev))
qs = [synth.lowrank_queries_torch(b * B, B, dim, dev) for b in range(8)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
K = 40
for T in (1, 2, 3, 4):
    streams = [torch.cuda.Stream(dev) for _ in range(T)]
    outs = [(torch.empty((B, k), dtype=torch.int64, device=dev), torch.empty((B, k), dtype=torch.float32, device=dev)) for _ in range(T)]
    def work(t, steps):
        with torch.cuda.stream(streams[t]):
            for s in range(steps):
                idx.search_batch_device(qs[(s * T + t) % 8], k, R, outs[t][0], outs[t][1])
    for t in range(T):
        work(t, 2)
    torch.cuda.synchronize()
    for with_flush in (False,):
        th = [threading.Thread(target=work, args=(t, K // T)) for t in range(T)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for x in th: x.start()
        for x in th: x.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        steps = (K // T) * T
        print(f"threads={T}: {steps} steps in {dt*1e3:.2f} ms -> {dt/steps*1e3:.3f} ms/step, {B*steps/dt/1e6:.3f} M QPS", flush=True)
