"""Row-sharded search across the GPUs of one box: one process per GPU, torch.distributed
(NCCL over NVLink/NVSwitch) for the single exchange step.

Partitioning: rank g owns the contiguous global rows
[g*ceil(N/G), min(N, (g+1)*ceil(N/G))), both codes and f32 originals; queries are replicated.
Each rank answers with one packed record buffer (its local top-R per query, already
rescored); ONE all-gather moves the buffers; every rank then runs the merge kernel, which
re-applies the stage-1 cut globally and orders by (cosine desc, hamming asc, row asc) —
the rule that makes the sharded answer equal to the single-index answer, where the
reference's own scatter/gather is concat + sort + truncate
(/root/reference/src/distributed/shard.rs:760-786).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_total: int, n_shards: int, shard: int) -> tuple[int, int]:
    """Global row range [lo, hi) of `shard`."""
    per = (n_total + n_shards - 1) // n_shards
    lo = min(n_total, shard * per)
    return lo, min(n_total, lo + per)


# ---- packed record layout helpers (host side; layout documented in include/gvdb.h) --------
def record_bytes(nq: int, r: int) -> int:
    return nq * r * 16


def pack_records(ids: np.ndarray, ham: np.ndarray, score: np.ndarray) -> np.ndarray:
    """[ids u64 | ham u32 | score f32], each nq x R, into one uint8 buffer."""
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    ham = np.ascontiguousarray(ham, dtype=np.uint32)
    score = np.ascontiguousarray(score, dtype=np.float32)
    assert ids.shape == ham.shape == score.shape and ids.ndim == 2
    return np.concatenate([ids.view(np.uint8).ravel(), ham.view(np.uint8).ravel(),
                           score.view(np.uint8).ravel()])


def unpack_records(buf: np.ndarray, nq: int, r: int):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    n = nq * r
    assert buf.size == n * 16
    ids = buf[:n * 8].view(np.uint64).reshape(nq, r)
    ham = buf[n * 8:n * 12].view(np.uint32).reshape(nq, r)
    score = buf[n * 12:].view(np.float32).reshape(nq, r)
    return ids, ham, score


def all_gather_records(local, group=None):
    """One collective: every rank's packed buffer, concatenated in rank order.
    `local` is a 1-D uint8 torch tensor (CUDA under NCCL, CPU under gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local, group=group)
    else:  # gloo
        parts = list(out.chunk(world))
        dist.all_gather(parts, local, group=group)
    return out


class ShardedSearcher:
    """Drives one rank's GpuIndex shard and the exchange + merge."""

    def __init__(self, index, group=None):
        import torch.distributed as dist
        self.index = index
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self._local = None
        self._all = None

    def search_batch_device(self, queries_t, k: int, rescore_count: int, ids_out=None,
                            scores_out=None):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return self.index.search_batch_device(queries_t, k, rescore_count, ids_out, scores_out)
        nq = queries_t.shape[0]
        nbytes = self.index.shard_record_bytes(nq, rescore_count)
        if self._local is None or self._local.numel() != nbytes:
            self._local = torch.empty(nbytes, dtype=torch.uint8, device=queries_t.device)
            self._all = torch.empty(nbytes * self.world, dtype=torch.uint8, device=queries_t.device)
        self.index.search_shard_device(queries_t, rescore_count, records_out=self._local)
        dist.all_gather_into_tensor(self._all, self._local, group=self.group)
        return self.index.merge_shards_device(self._all, self.world, nq, rescore_count, k,
                                              ids_out, scores_out)
