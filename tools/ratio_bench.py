#!/usr/bin/env python
"""Ratio mode (the reference's default rescore_ratio = 0.1) on BASELINE configs[1]'s corpus: 1M x 768, batches of
1024 queries, R = (N as f32 * 0.1) as usize = 100 000 candidates per query, top-10.  The candidates are rescored
through the dense tcgen05 bf16 filter (gvdb_ratio.cuh) + exact rescoring of its survivors.  Prints one JSON line:
QPS, per-stage device times, the bf16 tensor rate of tc_dot_kernel against MEASURED_PEAKS.json's bf16 figure,
parity of a few queries against the CPU oracle, and the oracle's own rate.
usage: python tools/ratio_bench.py [rows] [dim] [nq] [steps] [checked]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import grape_vector_db_b200 as gv
from grape_vector_db_b200 import synth
from oracle import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
checked = int(sys.argv[5]) if len(sys.argv) > 5 else 8
k, ratio = 10, 0.1
R = oracle.rescore_count(n, ratio)
dev = torch.device("cuda", 0)
idx = gv.GpuIndex(dim, device=0, capacity_rows=n)
for i in range(0, n, 131072):
    idx.add_device(synth.lowrank_rows_torch(i, min(131072, n - i), dim, dev))
qd = [synth.lowrank_queries_torch(b * nq, nq, dim, dev) for b in range(2)]
ids_t, sc_t = idx.search_batch_device(qd[0], k, R)        # builds the bf16 copy of the rows, warms up
torch.cuda.synchronize()
idx.profile_read(reset=True)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
for s in range(steps):
    ev[s][0].record()
    idx.search_batch_device(qd[s % 2], k, R)
    ev[s][1].record()
torch.cuda.synchronize()
ms = sum(a.elapsed_time(b) for a, b in ev) / steps
idx.profile_enable(True)
for s in range(steps):
    idx.search_batch_device(qd[s % 2], k, R)
torch.cuda.synchronize()
p = idx.profile_read(reset=True)
idx.profile_enable(False)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {}
bf16_peak = peaks.get("bf16_tflops", 1590.0)
dot_tflops = 2.0 * p["dot_macs"] / (p["dot_ms"] * 1e-3) / 1e12 if p["dot_ms"] > 0 else 0.0
ids_t, sc_t = idx.search_batch_device(qd[0], k, R)
torch.cuda.synchronize()
ids, sc = ids_t[:checked].cpu().numpy().astype(np.uint64), sc_t[:checked].cpu().numpy()
rows = np.concatenate([synth.lowrank_rows_torch(i, min(131072, n - i), dim, dev).cpu().numpy() for i in range(0, n, 131072)])
qs = qd[0][:checked].cpu().numpy()
t0 = time.perf_counter()
oi, os_ = oracle.multi_stage_search_batch(qs, rows, R, k, nthreads=oracle.hardware_threads())
cpu_s = time.perf_counter() - t0
line = {
    "workload": f"ratio mode: {n}x{dim}, batch {nq}, rescore_ratio {ratio} -> R = {R} candidates per query, top-{k}",
    "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms,
    "stage_ms_per_batch": {x: p[x] / steps for x in ("prep_ms", "sample_ms", "tc_ms", "scatter_ms", "select_ms", "rescore_ms", "topk_ms", "dot_ms")},
    "tc_dot_kernel": {"bound": "tensor", "achieved": dot_tflops, "peak": bf16_peak, "unit": "TFLOP/s", "frac": dot_tflops / bf16_peak,
                      "peak_source": "bf16_tflops (burst) of MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                      "ms_per_launch": p["dot_ms"] / max(1, p["dot_launches"]),
                      "note": "2 x rows x padded queries x dims over the kernel's CUDA-event time; kind::f16 (bf16 operands, f32 accumulate)"},
    "fp4_scan_launches_per_batch": p["tc_launches"] / steps,
    "ratio_fallback_queries_per_batch": p["ratio_fallback_queries"] / steps,
    "parity": {"checked_queries": checked, "topk_ids_bit_exact": bool(np.array_equal(ids, oi)),
               "scores_bit_exact": bool(np.array_equal(sc.view(np.uint32), os_.view(np.uint32)))},
    "cpu_baseline": {"value": checked / cpu_s, "unit": "queries/s", "cores": oracle.hardware_threads(), "kind": "port",
                     "sample": f"{checked} queries of the batch, oracle port of multi_stage_search (full stable sort of {n} pairs, {R} cosines per query)"},
}
print(json.dumps(line))
