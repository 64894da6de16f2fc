#!/usr/bin/env python
"""peer-rows layout: rescoring time when every rank searches at once vs. when only rank 0 does.
torchrun --nproc-per-node N tools/peer_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import grape_vector_db_b200 as gv
from grape_vector_db_b200 import dist as gdist, synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n, dim, B, k, R = 1_000_000, 768, 1024, 10, 40
lo, hi = gdist.shard_bounds(n, world, rank)
idx = gv.GpuIndex(dim, device=lr, capacity_rows=n, row_window=(lo, hi - lo))
for i in range(0, n, 131072):
    idx.add_device(synth.lowrank_rows_torch(i, min(131072, n - i), dim, dev))
gdist.attach_peer_rows(idx, n)
q = synth.lowrank_queries_torch(rank * B, B, dim, dev)
def run(steps):
    idx.profile_read(reset=True); idx.profile_enable(True)
    for _ in range(steps):
        idx.search_batch_device(q, k, R)
    torch.cuda.synchronize()
    p = idx.profile_read(reset=True); idx.profile_enable(False)
    return p["rescore_ms"] / steps
run(3)
dist.barrier()
all_ms = run(10)
dist.barrier()
solo = None
if rank == 0:
    solo = run(10)
dist.barrier()
if rank == 0:
    print(f"world {world}: rescore_ms per 1024-query batch: all ranks at once {all_ms:.3f}, rank 0 alone {solo:.3f}", flush=True)
dist.destroy_process_group()
