#!/usr/bin/env python
"""The HBM-bound operating point of the CUDA-core scan: 1 query per corpus pass over 8M x 768
(codes 768 MB > L2).  Used under ncu: -k regex:^scan_kernel -s 9 -c 3 captures one pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import grape_vector_db_b200 as gv
from grape_vector_db_b200 import synth
dev = torch.device("cuda", 0)
n, dim = 8_000_000, 768
idx = gv.GpuIndex(dim, device=0, capacity_rows=n)
for i in range(0, n, 262144):
    idx.add_device(synth.lowrank_rows_torch(i, min(262144, n - i), dim, dev))
q = synth.lowrank_queries_torch(0, 1, dim, dev)
idx.profile_enable(True)
for _ in range(5):
    idx.search_batch_device(q, 10, 40)
torch.cuda.synchronize()
p = idx.profile_read()
print("scan GB/s", p["scan_bytes"] / (p["scan_ms"] * 1e-3) / 1e9, "launches", p["scan_launches"])
