"""The N>1 host logic on CPU: two gloo ranks run the product's sharding arithmetic, packed-record
layout and the one-collective exchange (grape_vector_db_b200.dist); per-shard searches and the
merge are done by the CPU oracle here (test infrastructure) because the product's shard search and
merge are CUDA kernels.  The merged answer must equal the single-index oracle answer."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, dim, nq, R, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from grape_vector_db_b200 import dist as gdist
    from grape_vector_db_b200 import synth
    from oracle import oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = gdist.shard_bounds(n, world, rank)
    rows = synth.lowrank_rows(lo, hi - lo, dim)          # each rank regenerates ITS rows only
    qs = synth.lowrank_queries(0, nq, dim)
    ids = np.full((nq, R), np.iinfo(np.uint64).max, dtype=np.uint64)
    ham = np.full((nq, R), 0xFFFFFFFF, dtype=np.uint32)
    sc = np.full((nq, R), -np.inf, dtype=np.float32)
    for qi in range(nq):
        i, s, ci, ch = oracle.multi_stage_search(qs[qi], rows, R, want_candidates=True)
        by = {int(c): t for t, c in enumerate(ci)}
        r = len(ci)
        ids[qi, :r] = ci + np.uint64(lo)
        ham[qi, :r] = ch
        for ii, ss in zip(i, s):
            sc[qi, by[int(ii)]] = ss
    local = torch.from_numpy(gdist.pack_records(ids, ham, sc))
    assert local.numel() == gdist.record_bytes(nq, R)
    allrec = gdist.all_gather_records(local).numpy()
    # the query-sliced exchange the product uses (one slice per rank): every rank merges ITS slice
    per_q = nq // world
    send = torch.from_numpy(np.concatenate([
        gdist.pack_records(ids[s * per_q:(s + 1) * per_q], ham[s * per_q:(s + 1) * per_q],
                           sc[s * per_q:(s + 1) * per_q]) for s in range(world)]))
    recv = gdist.all_to_all_records(send).numpy()
    per_b = gdist.record_bytes(per_q, R)
    mine = [gdist.unpack_records(recv[s * per_b:(s + 1) * per_b], per_q, R) for s in range(world)]
    sl_i = np.zeros((per_q, k), np.uint64)
    sl_s = np.zeros((per_q, k), np.float32)
    for qi in range(per_q):
        gi, gs = oracle.shard_merge(np.concatenate([p[1][qi] for p in mine]),
                                    np.concatenate([p[0][qi] for p in mine]),
                                    np.concatenate([p[2][qi] for p in mine]), R, k)
        sl_i[qi, :len(gi)] = gi
        sl_s[qi, :len(gs)] = gs
    lo_q, hi_q = gdist.slice_bounds(nq, world, rank)
    assert hi_q - lo_q == per_q
    np.save(os.path.join(out_dir, f"slice_ids_{rank}.npy"), sl_i)
    np.save(os.path.join(out_dir, f"slice_sc_{rank}.npy"), sl_s)
    if rank == 0:
        per = gdist.record_bytes(nq, R)
        parts = [gdist.unpack_records(allrec[s * per:(s + 1) * per], nq, R) for s in range(world)]
        got_i = np.zeros((nq, k), np.uint64)
        got_s = np.zeros((nq, k), np.float32)
        for qi in range(nq):
            gi, gs = oracle.shard_merge(np.concatenate([p[1][qi] for p in parts]),
                                        np.concatenate([p[0][qi] for p in parts]),
                                        np.concatenate([p[2][qi] for p in parts]), R, k)
            got_i[qi, :len(gi)] = gi
            got_s[qi, :len(gs)] = gs
        np.save(os.path.join(out_dir, "ids.npy"), got_i)
        np.save(os.path.join(out_dir, "sc.npy"), got_s)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_and_partition():
    from grape_vector_db_b200 import dist as gdist
    for n in (0, 1, 7, 1000, 1_000_000, 100_000_001):
        for g in (1, 2, 3, 4, 8):
            b = [gdist.shard_bounds(n, g, s) for s in range(g)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(g - 1))
            assert all(lo <= hi for lo, hi in b)


def test_pack_unpack_roundtrip():
    from grape_vector_db_b200 import dist as gdist
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 2**40, size=(5, 7)).astype(np.uint64)
    ham = rng.integers(0, 768, size=(5, 7)).astype(np.uint32)
    sc = rng.standard_normal((5, 7)).astype(np.float32)
    buf = gdist.pack_records(ids, ham, sc)
    assert buf.size == gdist.record_bytes(5, 7)
    a, b, c = gdist.unpack_records(buf, 5, 7)
    assert np.array_equal(a, ids) and np.array_equal(b, ham) and np.array_equal(c, sc)


@pytest.mark.timeout(300)
def test_two_rank_gloo_exchange_and_merge(tmp_path):
    import torch.multiprocessing as mp
    from grape_vector_db_b200 import synth
    from oracle import oracle
    n, dim, nq, R, k, world = 5000, 128, 6, 40, 10, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, dim, nq, R, k, str(tmp_path)), nprocs=world, join=True)
    rows = synth.lowrank_rows(0, n, dim)
    qs = synth.lowrank_queries(0, nq, dim)
    want_i, want_s = oracle.multi_stage_search_batch(qs, rows, R, k)
    got_i = np.load(tmp_path / "ids.npy")
    got_s = np.load(tmp_path / "sc.npy")
    assert np.array_equal(got_i, want_i)
    assert np.array_equal(got_s.view(np.uint32), want_s.view(np.uint32))
    # query-sliced all-to-all: rank r's merged slice == rows [r*nq/world, ...) of the answer
    sl_i = np.concatenate([np.load(tmp_path / f"slice_ids_{r}.npy") for r in range(world)])
    sl_s = np.concatenate([np.load(tmp_path / f"slice_sc_{r}.npy") for r in range(world)])
    assert np.array_equal(sl_i, want_i)
    assert np.array_equal(sl_s.view(np.uint32), want_s.view(np.uint32))


class _CpuShard:
    """Stands in for a windowed GpuIndex on CPU (the product's are CUDA kernels): same three calls
    QueryParallelSearcher makes, answered by the oracle.  Holds ALL rows' codes, f32 rows of
    [lo, hi) only."""

    def __init__(self, rows_all, lo, hi):
        from oracle import oracle
        self.o = oracle
        self.codes = oracle.quantize_batch(rows_all)
        self.lo, self.hi = lo, hi
        self.rows = rows_all[lo:hi].copy()
        self.n = rows_all.shape[0]

    def stage1_device(self, q_t, R):
        import torch
        q = q_t.numpy()
        keys = np.full((q.shape[0], R), -1, dtype=np.int64)
        for i in range(q.shape[0]):
            ham = self.o.hamming_all(self.o.quantize(q[i]), self.codes).astype(np.int64)
            order = np.lexsort((np.arange(self.n), ham))[:R]
            keys[i, :len(order)] = (ham[order] << 40) | order
        return torch.from_numpy(keys)

    def rescore_keys_device(self, q_t, keys_t):
        import torch
        q, keys = q_t.numpy(), keys_t.numpy()
        out = np.zeros(keys.shape, dtype=np.float32)
        for i in range(keys.shape[0]):
            for j in range(keys.shape[1]):
                if keys[i, j] < 0:
                    continue
                row = int(keys[i, j] & ((1 << 40) - 1))
                if self.lo <= row < self.hi:
                    out[i, j] = self.o.cosine_similarity(q[i], self.rows[row - self.lo])
        return torch.from_numpy(out)

    def finish_owned_device(self, keys_t, by_owner_t, rows_per_owner, k, ids_out=None, scores_out=None):
        import torch
        keys, by_owner = keys_t.numpy(), by_owner_t.numpy()
        nq, R = keys.shape
        ids = np.full((nq, k), np.iinfo(np.uint64).max, dtype=np.uint64)
        sc = np.full((nq, k), -np.inf, dtype=np.float32)
        for i in range(nq):
            valid = keys[i] >= 0
            rows = (keys[i] & ((1 << 40) - 1))[valid]
            owner = np.minimum(by_owner.shape[0] - 1, rows // rows_per_owner)
            score = by_owner[owner, i, np.flatnonzero(valid)]
            order = np.argsort(-score, kind="stable")[:k]          # stage-1 order breaks ties
            ids[i, :len(order)] = rows[order].astype(np.uint64)
            sc[i, :len(order)] = score[order]
        return torch.from_numpy(ids.view(np.int64)), torch.from_numpy(sc)


def _qp_worker(rank, world, port, n, dim, nq_per, R, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from grape_vector_db_b200 import dist as gdist
    from grape_vector_db_b200 import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rows = synth.lowrank_rows(0, n, dim)
    lo, hi = gdist.shard_bounds(n, world, rank)
    searcher = gdist.QueryParallelSearcher(_CpuShard(rows, lo, hi), n)
    assert searcher.rows_per_owner == (n + world - 1) // world
    my_q = synth.lowrank_queries(rank * nq_per, nq_per, dim)
    ids, sc = searcher.search_batch_device(torch.from_numpy(my_q), k, R)
    np.save(os.path.join(out_dir, f"qp_ids_{rank}.npy"), ids.numpy().view(np.uint64))
    np.save(os.path.join(out_dir, f"qp_sc_{rank}.npy"), sc.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_query_parallel(tmp_path):
    """dist.QueryParallelSearcher's three collectives (all-gather queries, all-gather keys,
    all-to-all scores) with two gloo ranks: each rank's answers for ITS queries equal the
    single-index oracle."""
    import torch.multiprocessing as mp
    from grape_vector_db_b200 import synth
    from oracle import oracle
    n, dim, nq_per, R, k, world = 3000, 64, 5, 40, 10, 2
    port = _free_port()
    mp.spawn(_qp_worker, args=(world, port, n, dim, nq_per, R, k, str(tmp_path)), nprocs=world, join=True)
    rows = synth.lowrank_rows(0, n, dim)
    for r in range(world):
        qs = synth.lowrank_queries(r * nq_per, nq_per, dim)
        want_i, want_s = oracle.multi_stage_search_batch(qs, rows, R, k)
        assert np.array_equal(np.load(tmp_path / f"qp_ids_{r}.npy"), want_i)
        assert np.array_equal(np.load(tmp_path / f"qp_sc_{r}.npy").view(np.uint32), want_s.view(np.uint32))


class _CpuRowShard:
    """Stands in for a row-sharded GpuIndex on CPU (the product's shard search and merge are CUDA kernels): the
    calls ShardedSearcher makes, answered by the oracle.  `refuse_first` makes the first enqueued pass leave garbage
    records and a nonzero verdict, as a refused single pass would."""

    def __init__(self, rows, lo, refuse_first):
        self.rows, self.lo, self.refuse = rows, lo, refuse_first
        self.enqueued = self.synchronous = 0

    def shard_record_bytes(self, nq, R):
        return nq * R * 16

    def _records(self, q_t, R, W, out):
        from grape_vector_db_b200 import dist as gdist
        from oracle import oracle
        qs = q_t.numpy()
        nq = qs.shape[0]
        ids = np.full((nq, R), np.iinfo(np.uint64).max, dtype=np.uint64)
        ham = np.full((nq, R), 0xFFFFFFFF, dtype=np.uint32)
        sc = np.full((nq, R), -np.inf, dtype=np.float32)
        for qi in range(nq):
            i, s, ci, ch = oracle.multi_stage_search(qs[qi], self.rows, R, want_candidates=True)
            by = {int(c): t for t, c in enumerate(ci)}
            ids[qi, :len(ci)] = ci + np.uint64(self.lo)
            ham[qi, :len(ci)] = ch
            for ii, ss in zip(i, s):
                sc[qi, by[int(ii)]] = ss
        per = nq // W
        import torch
        out.copy_(torch.from_numpy(np.concatenate([
            gdist.pack_records(ids[s * per:(s + 1) * per], ham[s * per:(s + 1) * per], sc[s * per:(s + 1) * per])
            for s in range(W)])))

    def search_shard_sliced_enqueue_device(self, q_t, R, W, send, verdict_out=None):
        self.enqueued += 1
        if self.refuse and self.enqueued == 1:          # (a worker may re-arm this with enqueued = -1: the second enqueue)
            send.fill_(0x5A)
            verdict_out[4] = 1
        else:
            self._records(q_t, R, W, send)
            verdict_out.zero_()
        return True

    def search_shard_sliced_device(self, q_t, R, W, records_out=None):
        self.synchronous += 1
        self._records(q_t, R, W, records_out)
        return records_out

    def search_shard_verify(self, device):
        return False

    def merge_shards_device(self, recv, W, per, R, k, ids_out, sc_out):
        import torch
        from grape_vector_db_b200 import dist as gdist
        from oracle import oracle
        buf = recv.numpy()
        pb = self.shard_record_bytes(per, R)
        parts = [gdist.unpack_records(buf[s * pb:(s + 1) * pb], per, R) for s in range(W)]
        for qi in range(per):
            gi, gs = oracle.shard_merge(np.concatenate([p[1][qi] for p in parts]), np.concatenate([p[0][qi] for p in parts]),
                                        np.concatenate([p[2][qi] for p in parts]), R, k)
            ids_out[qi, :len(gi)] = torch.from_numpy(gi.astype(np.int64))
            sc_out[qi, :len(gs)] = torch.from_numpy(gs)
        return ids_out, sc_out


class _CpuRatioShard(_CpuRowShard):
    """The ratio-mode calls of a row shard restated with numpy (an independent statement of the protocol in
    include/gvdb.h: histograms -> gathered -> the same global cut on every shard -> each shard's members of the
    global top R rescored where they live -> best-k records merged)."""

    def __init__(self, rows, lo):
        super().__init__(rows, lo, refuse_first=False)
        from oracle import oracle
        self.codes = oracle.quantize_batch(rows)
        self.bins = rows.shape[1] + 1

    def shard_hist_bins(self):
        return self.bins

    def _ham(self, q):
        from oracle import oracle
        return oracle.hamming_all(oracle.quantize(q), self.codes).astype(np.int64)

    def shard_hist_device(self, q_t):
        import torch
        qs = q_t.numpy()
        return torch.from_numpy(np.stack([np.bincount(self._ham(q), minlength=self.bins) for q in qs]).astype(np.int32))

    def search_shard_ratio_device(self, q_t, R, k, hists_all, W, my):
        import torch
        from grape_vector_db_b200 import dist as gdist
        from oracle import oracle
        qs, H = q_t.numpy(), hists_all.numpy().astype(np.int64)
        nq = qs.shape[0]
        ids = np.full((nq, k), np.iinfo(np.uint64).max, dtype=np.uint64)
        ham_o = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        sc = np.full((nq, k), -np.inf, dtype=np.float32)
        for qi in range(nq):
            g = H[:, qi, :].sum(axis=0)
            cum = np.cumsum(g)
            r_eff = min(R, int(cum[-1]))
            if r_eff == 0:
                continue
            b = int(np.searchsorted(cum, r_eff))                   # first bin whose cumulative count reaches r_eff
            need = r_eff - int(cum[b] - g[b])
            before = int(H[:my, qi, b].sum())
            keep = max(0, min(int(H[my, qi, b]), need - before))
            ham = self._ham(qs[qi])
            sel = np.concatenate([np.flatnonzero(ham < b), np.flatnonzero(ham == b)[:keep]])
            sel = sel[np.lexsort((sel, ham[sel]))]                 # stage-1 order: (hamming, row)
            cos = np.array([oracle.cosine_similarity(qs[qi], self.rows[r]) for r in sel], dtype=np.float32)
            best = np.sort(np.argsort(-cos, kind="stable")[:k])    # best k by cosine, listed in stage-1 order again
            ids[qi, :best.size] = sel[best].astype(np.uint64) + np.uint64(self.lo)
            ham_o[qi, :best.size] = ham[sel[best]]
            sc[qi, :best.size] = cos[best]
        return torch.from_numpy(gdist.pack_records(ids, ham_o, sc))

    def merge_shards_ratio_device(self, rec_all, W, nq, k, ids_out=None, scores_out=None):
        import torch
        from grape_vector_db_b200 import dist as gdist
        buf = rec_all.numpy()
        pb = self.shard_record_bytes(nq, k)
        parts = [gdist.unpack_records(buf[s * pb:(s + 1) * pb], nq, k) for s in range(W)]
        out_i = np.full((nq, k), -1, dtype=np.int64)
        out_s = np.full((nq, k), -np.inf, dtype=np.float32)
        for qi in range(nq):
            i_ = np.concatenate([p[0][qi] for p in parts]); h_ = np.concatenate([p[1][qi] for p in parts])
            s_ = np.concatenate([p[2][qi] for p in parts])
            ok = i_ != np.iinfo(np.uint64).max
            i_, h_, s_ = i_[ok], h_[ok], s_[ok]
            o1 = np.lexsort((i_, h_))
            o2 = np.argsort(-s_[o1], kind="stable")[:k]
            out_i[qi, :o2.size] = i_[o1][o2].astype(np.int64)
            out_s[qi, :o2.size] = s_[o1][o2]
        return torch.from_numpy(out_i), torch.from_numpy(out_s)


def _ratio_worker(rank, world, port, n, dim, nq, R, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from grape_vector_db_b200 import dist as gdist
    from grape_vector_db_b200 import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = gdist.shard_bounds(n, world, rank)
    rng = np.random.default_rng(5)
    rows_all = rng.integers(-1, 2, size=(n, dim)).astype(np.float32)      # small alphabet: ties straddle the shards
    searcher = gdist.ShardedSearcher(_CpuRatioShard(rows_all[lo:hi], lo))
    qs = torch.from_numpy(rng.integers(-1, 2, size=(nq, dim)).astype(np.float32))
    ids, sc = searcher.search_batch_device(qs, k, R)                      # R > 2048: the ratio-mode exchange
    np.save(os.path.join(out_dir, f"r_ids_{rank}.npy"), ids.numpy())
    np.save(os.path.join(out_dir, f"r_sc_{rank}.npy"), sc.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_ratio_mode_across_shards(tmp_path):
    """ShardedSearcher's ratio-mode exchange (rescore_count > 2048) on two gloo ranks with numpy shards: histograms
    gathered, the same global cut on both ranks, ties inside the cut bin shared out in rank order, best-k records
    merged — equal to the oracle's single-index answer."""
    import torch.multiprocessing as mp
    from oracle import oracle
    n, dim, nq, R, k, world = 6000, 24, 4, 2500, 12, 2
    port = _free_port()
    mp.spawn(_ratio_worker, args=(world, port, n, dim, nq, R, k, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    rows = rng.integers(-1, 2, size=(n, dim)).astype(np.float32)
    qs = rng.integers(-1, 2, size=(nq, dim)).astype(np.float32)
    for qi in range(nq):
        oi, os_ = oracle.multi_stage_search(qs[qi], rows, R)
        for r in range(world):
            assert np.array_equal(np.load(tmp_path / f"r_ids_{r}.npy")[qi].astype(np.uint64), oi[:k])
            assert np.array_equal(np.load(tmp_path / f"r_sc_{r}.npy")[qi].view(np.uint32), os_[:k].view(np.uint32))


def _sharded_worker(rank, world, port, n, dim, nq, R, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from grape_vector_db_b200 import dist as gdist
    from grape_vector_db_b200 import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = gdist.shard_bounds(n, world, rank)
    shard = _CpuRowShard(synth.lowrank_rows(lo, hi - lo, dim), lo, refuse_first=(rank == 1))
    searcher = gdist.ShardedSearcher(shard)
    qs = torch.from_numpy(synth.lowrank_queries(0, nq, dim))
    ids1, sc1 = searcher.search_batch_device(qs, k, R)            # rank 1 refuses its pass: every rank repeats the step
    assert searcher.reruns == 1 and shard.synchronous == 1, (searcher.reruns, shard.synchronous)
    ids2, sc2 = searcher.search_batch_device(qs[:nq - 1], k, R)   # nq not a multiple of the world size; no rerun
    assert searcher.reruns == 1 and shard.synchronous == 1
    # several tiles per batch (all enqueued before the one host wait), ragged tile sizes, a refusal in the middle
    searcher.QUERY_TILE = 3
    shard.refuse, shard.enqueued = (rank == 0), -1            # rank 0 refuses its SECOND enqueue of this batch (tile 1)
    ids3, sc3 = searcher.search_batch_device(qs, k, R)
    assert searcher.reruns == 2 and shard.synchronous == 2, (searcher.reruns, shard.synchronous)
    assert torch.equal(ids3, ids1) and torch.equal(sc3.view(torch.int32), sc1.view(torch.int32))
    np.save(os.path.join(out_dir, f"a_ids_{rank}.npy"), ids1.numpy())
    np.save(os.path.join(out_dir, f"a_sc_{rank}.npy"), sc1.numpy())
    np.save(os.path.join(out_dir, f"b_ids_{rank}.npy"), ids2.numpy())
    np.save(os.path.join(out_dir, f"b_sc_{rank}.npy"), sc2.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_searcher_one_gather_and_verdicts(tmp_path):
    """ShardedSearcher itself on two gloo ranks: packed k-lists + verdict words in ONE all-gather, a refused pass on
    one rank repeats the step on every rank, ragged batch sizes; every rank ends with the single-index answer."""
    import torch.multiprocessing as mp
    from grape_vector_db_b200 import synth
    from oracle import oracle
    n, dim, nq, R, k, world = 3000, 128, 8, 20, 5, 2
    port = _free_port()
    mp.spawn(_sharded_worker, args=(world, port, n, dim, nq, R, k, str(tmp_path)), nprocs=world, join=True)
    rows = synth.lowrank_rows(0, n, dim)
    qs = synth.lowrank_queries(0, nq, dim)
    want_i, want_s = oracle.multi_stage_search_batch(qs, rows, R, k)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"a_ids_{r}.npy").astype(np.uint64), want_i)
        assert np.array_equal(np.load(tmp_path / f"a_sc_{r}.npy").view(np.uint32), want_s.view(np.uint32))
        assert np.array_equal(np.load(tmp_path / f"b_ids_{r}.npy").astype(np.uint64), want_i[:nq - 1])
        assert np.array_equal(np.load(tmp_path / f"b_sc_{r}.npy").view(np.uint32), want_s[:nq - 1].view(np.uint32))
