"""GpuIndex: one shard of the corpus in one B200's HBM, driven through the C ABI.

Thin, typed wrapper over include/gvdb.h.  Host-array entry points take numpy arrays;
`*_device` entry points take torch CUDA tensors and run on torch's current stream
(torch is plumbing here: device memory, streams, torch.distributed).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .errors import raise_for_status

NO_ID = np.uint64(_ffi.GVDB_NO_ID)


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_allow_bits(allowed, n_rows: int) -> np.ndarray:
    """Allow-list of a filtered search as the uint32 bitmap the C ABI takes: bit (r % 32) of word r / 32 = 1
    when local row r may be returned (the tombstone bitmap's layout).  `allowed` is a boolean mask over the
    stored rows or an array of allowed local row numbers."""
    a = np.asarray(allowed)
    if a.dtype != np.bool_:
        m = np.zeros(n_rows, dtype=bool)
        m[a.astype(np.int64)] = True
        a = m
    assert a.shape == (n_rows,)
    pad = np.zeros(((n_rows + 31) // 32) * 32, dtype=bool)
    pad[:n_rows] = a
    return np.packbits(pad.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).ravel().copy()


class GpuIndex:
    """Rows [row_base, row_base + len) of the corpus: 1-bit codes + f32 originals in HBM."""

    def __init__(self, dim: int, threshold: float = 0.0, rescore_ratio: float = 0.1,
                 device: int = 0, capacity_rows: int = 0, row_base: int = 0,
                 row_window: tuple[int, int] | None = None):
        """row_window = (first, count): keep the f32 originals only for those local rows (codes and
        norms are kept for every row) — the "codes replicated, rows sharded" multi-GPU layout."""
        self._lib = _ffi.lib()
        cfg = _ffi.GvdbConfig()
        cfg.struct_size = C.sizeof(_ffi.GvdbConfig)
        cfg.dim = dim
        cfg.threshold = threshold
        cfg.rescore_ratio = rescore_ratio
        cfg.device = device
        cfg.flags = 0
        cfg.capacity_rows = capacity_rows
        cfg.row_base = row_base
        if row_window is not None:
            cfg.flags = _ffi.GVDB_FLAG_ROW_WINDOW
            cfg.window_first, cfg.window_count = int(row_window[0]), int(row_window[1])
        h = C.c_void_p()
        raise_for_status(self._lib.gvdb_create(C.byref(cfg), C.byref(h)), self._lib)
        self._h = h
        self.dim = dim
        self.threshold = threshold
        self.rescore_ratio = rescore_ratio
        self.device = device
        self.row_base = row_base
        self.nbytes = (dim + 7) // 8

    # -- persistence ---------------------------------------------------------------------
    def save(self, path: str):
        """Write the shard (codes, norms, tombstones, f32 rows) to one flat file (gvdb_save)."""
        self._ok(self._lib.gvdb_save(self._h, str(path).encode()))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "GpuIndex":
        """Restore a shard written by save() without re-quantising (gvdb_load)."""
        lib = _ffi.lib()
        h = C.c_void_p()
        raise_for_status(lib.gvdb_load(str(path).encode(), device, C.byref(h)), lib)
        self = cls.__new__(cls)
        self._lib = lib
        self._h = h
        s = _ffi.GvdbStats()
        raise_for_status(lib.gvdb_get_stats(h, C.byref(s)), lib)
        self.dim = int(s.dimension)
        self.device = device
        self.nbytes = (self.dim + 7) // 8
        self.threshold = None      # stored in the file; not needed on the Python side
        self.rescore_ratio = 0.1
        self.row_base = None
        return self

    # -- lifecycle ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gvdb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ok(self, st):
        raise_for_status(st, self._lib)

    # -- ingest ------------------------------------------------------------------------
    def add(self, rows) -> int:
        rows = _np(rows, np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            from .errors import DimensionMismatch
            raise DimensionMismatch(self.dim, rows.shape[-1] if rows.ndim else 0)
        first = C.c_uint64()
        self._ok(self._lib.gvdb_add(self._h, _ptr(rows), rows.shape[0], C.byref(first)))
        return int(first.value)

    def add_device(self, rows_t) -> int:
        import torch
        assert rows_t.is_cuda and rows_t.dtype == torch.float32 and rows_t.is_contiguous()
        assert rows_t.dim() == 2 and rows_t.shape[1] == self.dim
        first = C.c_uint64()
        st = torch.cuda.current_stream(rows_t.device).cuda_stream
        self._ok(self._lib.gvdb_add_device(self._h, C.c_void_p(st), C.c_void_p(rows_t.data_ptr()),
                                           rows_t.shape[0], C.byref(first)))
        return int(first.value)

    def reserve(self, capacity_rows: int):
        self._ok(self._lib.gvdb_reserve(self._h, capacity_rows))

    def remove(self, local_row: int) -> bool:
        was = C.c_int32()
        self._ok(self._lib.gvdb_remove(self._h, local_row, C.byref(was)))
        return bool(was.value)

    def clear(self):
        self._ok(self._lib.gvdb_clear(self._h))

    def __len__(self):
        return int(self._lib.gvdb_len(self._h))

    def stats(self) -> dict:
        s = _ffi.GvdbStats()
        self._ok(self._lib.gvdb_get_stats(self._h, C.byref(s)))
        return {f: int(getattr(s, f)) for f, _ in s._fields_}

    @property
    def rows(self) -> int:
        return self.stats()["rows"]

    # -- quantizer pieces --------------------------------------------------------------
    def quantize(self, x) -> np.ndarray:
        x = _np(x, np.float32)
        single = x.ndim == 1
        x2 = x.reshape(1, -1) if single else x
        if x2.shape[1] != self.dim:
            from .errors import DimensionMismatch
            raise DimensionMismatch(self.dim, x2.shape[1])
        out = np.zeros((x2.shape[0], self.nbytes), dtype=np.uint8)
        self._ok(self._lib.gvdb_quantize(self._h, _ptr(x2), x2.shape[0], _ptr(out)))
        return out[0] if single else out

    def get_codes(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.rows - first if n is None else n
        out = np.zeros((n, self.nbytes), dtype=np.uint8)
        self._ok(self._lib.gvdb_get_codes(self._h, first, n, _ptr(out)))
        return out

    def hamming(self, q_codes) -> np.ndarray:
        q = _np(q_codes, np.uint8)
        q2 = q.reshape(1, -1) if q.ndim == 1 else q
        assert q2.shape[1] == self.nbytes
        out = np.zeros((q2.shape[0], self.rows), dtype=np.uint32)
        self._ok(self._lib.gvdb_hamming(self._h, _ptr(q2), q2.shape[0], _ptr(out)))
        return out

    def rescore_count(self, n: int | None = None, ratio: float | None = None) -> int:
        n = len(self) if n is None else n
        ratio = self.rescore_ratio if ratio is None else ratio
        return int(self._lib.gvdb_rescore_count(n, ratio))

    # -- two-stage search ----------------------------------------------------------------
    def search_batch(self, queries, k: int, rescore_count: int, want_candidates: bool = False):
        q = _np(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            from .errors import DimensionMismatch
            raise DimensionMismatch(self.dim, q.shape[-1] if q.ndim else 0)
        nq = q.shape[0]
        ids = np.full((nq, k), NO_ID, dtype=np.uint64)
        sc = np.full((nq, k), -np.inf, dtype=np.float32)
        ci = ch = None
        if want_candidates:
            ci = np.full((nq, rescore_count), NO_ID, dtype=np.uint64)
            ch = np.full((nq, rescore_count), 0xFFFFFFFF, dtype=np.uint32)
        self._ok(self._lib.gvdb_search_batch(self._h, _ptr(q), nq, k, rescore_count, _ptr(ids),
                                             _ptr(sc), _ptr(ci), _ptr(ch)))
        return (ids, sc, ci, ch) if want_candidates else (ids, sc)

    def search_batch_device(self, queries_t, k: int, rescore_count: int, ids_out=None,
                            scores_out=None):
        import torch
        assert queries_t.is_cuda and queries_t.dtype == torch.float32 and queries_t.is_contiguous()
        nq = queries_t.shape[0]
        dev = queries_t.device
        if ids_out is None:
            ids_out = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if scores_out is None:
            scores_out = torch.empty((nq, k), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._ok(self._lib.gvdb_search_batch_device(
            self._h, C.c_void_p(st), C.c_void_p(queries_t.data_ptr()), nq, k, rescore_count,
            C.c_void_p(ids_out.data_ptr()), C.c_void_p(scores_out.data_ptr()), None, None))
        return ids_out, scores_out

    # -- exact flat search ----------------------------------------------------------------
    def flat_search_batch(self, queries, k: int):
        q = _np(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            from .errors import DimensionMismatch
            raise DimensionMismatch(self.dim, q.shape[-1] if q.ndim else 0)
        nq = q.shape[0]
        ids = np.full((nq, k), NO_ID, dtype=np.uint64)
        ds = np.full((nq, k), np.inf, dtype=np.float32)
        self._ok(self._lib.gvdb_flat_search_batch(self._h, _ptr(q), nq, k, _ptr(ids), _ptr(ds)))
        return ids, ds

    def similarity_search_batch(self, queries, k: int, threshold: float | None = None):
        """BasicVectorStore::vector_search: (ids, cosine similarities) descending, optional threshold."""
        q = _np(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            from .errors import DimensionMismatch
            raise DimensionMismatch(self.dim, q.shape[-1] if q.ndim else 0)
        nq = q.shape[0]
        ids = np.full((nq, k), NO_ID, dtype=np.uint64)
        sims = np.full((nq, k), -np.inf, dtype=np.float32)
        self._ok(self._lib.gvdb_similarity_search_batch(self._h, _ptr(q), nq, k,
                                                        0.0 if threshold is None else float(threshold),
                                                        0 if threshold is None else 1, _ptr(ids), _ptr(sims)))
        return ids, sims

    def flat_search_batch_device(self, queries_t, k: int):
        import torch
        assert queries_t.is_cuda and queries_t.dtype == torch.float32 and queries_t.is_contiguous()
        nq = queries_t.shape[0]
        dev = queries_t.device
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        ds = torch.empty((nq, k), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._ok(self._lib.gvdb_flat_search_batch_device(
            self._h, C.c_void_p(st), C.c_void_p(queries_t.data_ptr()), nq, k,
            C.c_void_p(ids.data_ptr()), C.c_void_p(ds.data_ptr())))
        return ids, ds

    # -- sharded search: the two halves around the exchange ---------------------------------
    def shard_record_bytes(self, nq: int, rescore_count: int) -> int:
        return int(self._lib.gvdb_shard_record_bytes(nq, rescore_count))

    def search_shard_device(self, queries_t, rescore_count: int, records_out=None):
        """This shard's packed records (uint8 tensor, [ids u64 | ham u32 | score f32] x nq x R)."""
        import torch
        nq = queries_t.shape[0]
        dev = queries_t.device
        nbytes = self.shard_record_bytes(nq, rescore_count)
        if records_out is None:
            records_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        assert records_out.numel() == nbytes and records_out.is_contiguous()
        st = torch.cuda.current_stream(dev).cuda_stream
        self._ok(self._lib.gvdb_search_shard_device(
            self._h, C.c_void_p(st), C.c_void_p(queries_t.data_ptr()), nq, rescore_count,
            C.c_void_p(records_out.data_ptr())))
        return records_out

    def search_shard_sliced_device(self, queries_t, rescore_count: int, n_slices: int, records_out=None):
        """n_slices packed record buffers back to back, buffer s for queries
        [s*nq/n_slices, (s+1)*nq/n_slices) — the send buffer of a query-sliced all-to-all."""
        import torch
        nq = queries_t.shape[0]
        assert nq % n_slices == 0
        dev = queries_t.device
        nbytes = n_slices * self.shard_record_bytes(nq // n_slices, rescore_count)
        if records_out is None:
            records_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        assert records_out.numel() == nbytes and records_out.is_contiguous()
        st = torch.cuda.current_stream(dev).cuda_stream
        self._ok(self._lib.gvdb_search_shard_sliced_device(
            self._h, C.c_void_p(st), C.c_void_p(queries_t.data_ptr()), nq, rescore_count, n_slices,
            C.c_void_p(records_out.data_ptr())))
        return records_out

    def search_shard_sliced_enqueue_device(self, queries_t, rescore_count: int, n_slices: int, records_out,
                                           verdict_out=None) -> bool:
        """The sliced shard search enqueued without a host synchronisation; False when this index / rescore count
        has no such form (use search_shard_sliced_device).  Follow with search_shard_verify().  verdict_out: optional
        device tensor (>= 8 bytes) that receives the pass's verdict words on the stream (nonzero = repeat)."""
        import torch
        nq = queries_t.shape[0]
        if (self.dim & 3) or rescore_count > 256 or nq % n_slices or 1024 % (nq // n_slices):
            return False
        st = torch.cuda.current_stream(queries_t.device).cuda_stream
        self._ok(self._lib.gvdb_search_shard_sliced_enqueue_device(
            self._h, C.c_void_p(st), C.c_void_p(queries_t.data_ptr()), nq, rescore_count, n_slices,
            C.c_void_p(records_out.data_ptr()), C.c_void_p(verdict_out.data_ptr() if verdict_out is not None else None)))
        return True

    def search_shard_verify(self, device) -> bool:
        """Waits for the stream; True when the enqueued pass must be repeated synchronously."""
        import torch
        rerun = C.c_int32(0)
        st = torch.cuda.current_stream(device).cuda_stream
        self._ok(self._lib.gvdb_search_shard_verify(self._h, C.c_void_p(st), C.byref(rerun)))
        return bool(rerun.value)

    def merge_shards_device(self, records_all, n_shards: int, nq: int, rescore_count: int, k: int,
                            ids_out=None, scores_out=None):
        """n_shards packed record buffers back to back -> (ids [nq,k] int64, scores [nq,k] f32)."""
        import torch
        assert records_all.is_cuda and records_all.is_contiguous()
        assert records_all.numel() == n_shards * self.shard_record_bytes(nq, rescore_count)
        dev = records_all.device
        if ids_out is None:
            ids_out = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if scores_out is None:
            scores_out = torch.empty((nq, k), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._ok(self._lib.gvdb_merge_shards_device(
            self._h, C.c_void_p(st), n_shards, C.c_void_p(records_all.data_ptr()), nq,
            rescore_count, k, C.c_void_p(ids_out.data_ptr()), C.c_void_p(scores_out.data_ptr())))
        return ids_out, scores_out

    # -- ratio mode across row shards (rescore_count > 2048) ------------------------------------
    def shard_hist_bins(self) -> int:
        return int(self._lib.gvdb_shard_hist_bins(self._h))

    def shard_hist_device(self, queries_t, hist_out=None):
        """[nq, bins] int32 tensor: this shard's live rows per Hamming distance, per query."""
        import torch
        nq = queries_t.shape[0]
        if hist_out is None:
            hist_out = torch.empty((nq, self.shard_hist_bins()), dtype=torch.int32, device=queries_t.device)
        st = torch.cuda.current_stream(queries_t.device).cuda_stream
        self._ok(self._lib.gvdb_shard_hist_device(self._h, C.c_void_p(st), C.c_void_p(queries_t.data_ptr()), nq,
                                                  C.c_void_p(hist_out.data_ptr())))
        return hist_out

    def search_shard_ratio_device(self, queries_t, rescore_count: int, k: int, hists_all_t, n_shards: int, my_shard: int,
                                  records_out=None):
        """hists_all_t: [n_shards, nq, bins] int32 (every shard's shard_hist_device output, rank order) ->
        this shard's packed records (its best k members of the global top rescore_count)."""
        import torch
        nq = queries_t.shape[0]
        assert hists_all_t.is_contiguous() and tuple(hists_all_t.shape) == (n_shards, nq, self.shard_hist_bins())
        nbytes = self.shard_record_bytes(nq, k)
        if records_out is None:
            records_out = torch.empty(nbytes, dtype=torch.uint8, device=queries_t.device)
        st = torch.cuda.current_stream(queries_t.device).cuda_stream
        self._ok(self._lib.gvdb_search_shard_ratio_device(
            self._h, C.c_void_p(st), C.c_void_p(queries_t.data_ptr()), nq, int(rescore_count), k,
            C.c_void_p(hists_all_t.data_ptr()), n_shards, my_shard, C.c_void_p(records_out.data_ptr())))
        return records_out

    def merge_shards_ratio_device(self, records_all, n_shards: int, nq: int, k: int, ids_out=None, scores_out=None):
        import torch
        assert records_all.is_contiguous() and records_all.numel() == n_shards * self.shard_record_bytes(nq, k)
        dev = records_all.device
        if ids_out is None:
            ids_out = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if scores_out is None:
            scores_out = torch.empty((nq, k), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._ok(self._lib.gvdb_merge_shards_ratio_device(
            self._h, C.c_void_p(st), n_shards, C.c_void_p(records_all.data_ptr()), nq, k,
            C.c_void_p(ids_out.data_ptr()), C.c_void_p(scores_out.data_ptr())))
        return ids_out, scores_out

    # -- query-parallel search: replicated codes, row-sharded originals ------------------------
    def stage1_device(self, queries_t, rescore_count: int, keys_out=None):
        """Stage 1 only: [nq, R] int64 keys hamming << 40 | global row, ascending, -1 (GVDB_NO_ID) unfilled."""
        import torch
        assert queries_t.is_cuda and queries_t.dtype == torch.float32 and queries_t.is_contiguous()
        nq = queries_t.shape[0]
        if keys_out is None:
            keys_out = torch.empty((nq, rescore_count), dtype=torch.int64, device=queries_t.device)
        st = torch.cuda.current_stream(queries_t.device).cuda_stream
        self._ok(self._lib.gvdb_stage1_device(self._h, C.c_void_p(st), C.c_void_p(queries_t.data_ptr()), nq,
                                              rescore_count, C.c_void_p(keys_out.data_ptr())))
        return keys_out

    def rescore_keys_device(self, queries_t, keys_t, scores_out=None):
        """Exact cosine for the keys whose row this index holds in f32 (its window), 0.0 elsewhere."""
        import torch
        assert queries_t.is_cuda and queries_t.is_contiguous() and keys_t.is_contiguous()
        nq, R = keys_t.shape
        assert queries_t.shape[0] == nq
        if scores_out is None:
            scores_out = torch.empty((nq, R), dtype=torch.float32, device=queries_t.device)
        st = torch.cuda.current_stream(queries_t.device).cuda_stream
        self._ok(self._lib.gvdb_rescore_keys_device(self._h, C.c_void_p(st), C.c_void_p(queries_t.data_ptr()), nq, R,
                                                    C.c_void_p(keys_t.data_ptr()), C.c_void_p(scores_out.data_ptr())))
        return scores_out

    def finish_owned_device(self, keys_t, scores_by_owner_t, rows_per_owner: int, k: int, ids_out=None,
                            scores_out=None):
        """keys [nq, R]; scores_by_owner [n_owners, nq, R] -> (ids [nq, k] int64, scores [nq, k] f32)."""
        import torch
        assert keys_t.is_contiguous() and scores_by_owner_t.is_contiguous()
        nq, R = keys_t.shape
        n_owners = scores_by_owner_t.shape[0]
        assert scores_by_owner_t.numel() == n_owners * nq * R
        dev = keys_t.device
        if ids_out is None:
            ids_out = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if scores_out is None:
            scores_out = torch.empty((nq, k), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._ok(self._lib.gvdb_finish_owned_device(
            self._h, C.c_void_p(st), C.c_void_p(keys_t.data_ptr()), C.c_void_p(scores_by_owner_t.data_ptr()),
            n_owners, rows_per_owner, nq, R, k, C.c_void_p(ids_out.data_ptr()), C.c_void_p(scores_out.data_ptr())))
        return ids_out, scores_out

    # -- peer rows: rescoring out of the other GPUs' HBM over NVLink -----------------------------
    def export_rows_ipc(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._ok(self._lib.gvdb_export_rows_ipc(self._h, buf))
        return bytes(buf)

    def attach_peer_rows_ipc(self, handles: list[bytes], rows_per_owner: int, my_owner: int):
        raw = b"".join(handles)
        assert len(raw) == 64 * len(handles)
        buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        self._ok(self._lib.gvdb_attach_peer_rows_ipc(self._h, len(handles), rows_per_owner, my_owner, buf))

    def rows_device_ptr(self) -> int:
        return int(self._lib.gvdb_rows_device_ptr(self._h) or 0)

    def attach_peer_rows_ptr(self, ptrs: list[int], rows_per_owner: int, my_owner: int):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self._ok(self._lib.gvdb_attach_peer_rows_ptr(self._h, len(ptrs), rows_per_owner, my_owner, arr))

    # -- filtered search -------------------------------------------------------------------------
    def _allow_bits(self, allowed) -> np.ndarray:
        return pack_allow_bits(allowed, self.rows)

    def search_batch_filtered(self, queries, allowed, k: int, rescore_count: int):
        """Two-stage search over the allowed live rows only (gvdb_search_batch_filtered)."""
        q = _np(queries, np.float32)
        nq = q.shape[0]
        bits = self._allow_bits(allowed)
        ids = np.full((nq, k), NO_ID, dtype=np.uint64)
        sc = np.full((nq, k), -np.inf, dtype=np.float32)
        self._ok(self._lib.gvdb_search_batch_filtered(self._h, _ptr(q), _ptr(bits), nq, k, rescore_count, _ptr(ids), _ptr(sc)))
        return ids, sc

    def flat_search_batch_filtered(self, queries, allowed, k: int):
        q = _np(queries, np.float32)
        nq = q.shape[0]
        bits = self._allow_bits(allowed)
        ids = np.full((nq, k), NO_ID, dtype=np.uint64)
        ds = np.full((nq, k), np.inf, dtype=np.float32)
        self._ok(self._lib.gvdb_flat_search_batch_filtered(self._h, _ptr(q), _ptr(bits), nq, k, _ptr(ids), _ptr(ds)))
        return ids, ds

    def search_batch_filtered_device(self, queries_t, allow_bits_t, k: int, rescore_count: int):
        import torch
        assert queries_t.is_cuda and queries_t.dtype == torch.float32 and queries_t.is_contiguous()
        assert allow_bits_t.is_cuda and allow_bits_t.dtype == torch.int32 and allow_bits_t.is_contiguous()
        nq = queries_t.shape[0]
        dev = queries_t.device
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._ok(self._lib.gvdb_search_batch_filtered_device(
            self._h, C.c_void_p(st), C.c_void_p(queries_t.data_ptr()), C.c_void_p(allow_bits_t.data_ptr()), nq, k,
            rescore_count, C.c_void_p(ids.data_ptr()), C.c_void_p(sc.data_ptr())))
        return ids, sc

    # -- peer exchange (no collective library in the data path) ---------------------------------
    def exchange_create(self, world: int, rank: int, rows_per_owner: int, nq_max: int, rescore_max: int):
        self._ok(self._lib.gvdb_exchange_create(self._h, world, rank, rows_per_owner, nq_max, rescore_max))

    def exchange_export_ipc(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._ok(self._lib.gvdb_exchange_export_ipc(self._h, buf))
        return bytes(buf)

    def exchange_attach_ipc(self, handles: list[bytes]):
        raw = b"".join(handles)
        assert len(raw) == 64 * len(handles)
        buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        self._ok(self._lib.gvdb_exchange_attach_ipc(self._h, buf))

    def exchange_mailbox_ptr(self) -> int:
        return int(self._lib.gvdb_exchange_mailbox_ptr(self._h) or 0)

    def exchange_attach_ptr(self, ptrs: list[int]):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self._ok(self._lib.gvdb_exchange_attach_ptr(self._h, arr))

    def search_exchange_device(self, my_queries_t, k: int, rescore_count: int, ids_out=None, scores_out=None):
        """One step of the peer exchange: every rank calls it with ITS batch (same nq everywhere)."""
        import torch
        assert my_queries_t.is_cuda and my_queries_t.dtype == torch.float32 and my_queries_t.is_contiguous()
        nq = my_queries_t.shape[0]
        dev = my_queries_t.device
        if ids_out is None:
            ids_out = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if scores_out is None:
            scores_out = torch.empty((nq, k), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._ok(self._lib.gvdb_search_exchange_device(self._h, C.c_void_p(st), C.c_void_p(my_queries_t.data_ptr()),
                                                       nq, k, rescore_count, C.c_void_p(ids_out.data_ptr()),
                                                       C.c_void_p(scores_out.data_ptr())))
        return ids_out, scores_out

    def exchange_status(self):
        """Waits for the steps in flight; raises IndexError_ if one of them timed out on a peer."""
        self._ok(self._lib.gvdb_exchange_status(self._h, None))

    def approx_dot(self, queries) -> np.ndarray:
        """gvdb_approx_dot: the bf16 tensor-core dot products of the ratio-mode filter, [nq, rows] f32."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        out = np.empty((q.shape[0], self.rows), dtype=np.float32)
        self._ok(self._lib.gvdb_approx_dot(self._h, q.ctypes.data_as(C.c_void_p), q.shape[0], out.ctypes.data_as(C.c_void_p)))
        return out

    # -- measurement hooks --------------------------------------------------------------------
    def profile_enable(self, on: bool = True):
        self._ok(self._lib.gvdb_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset: bool = True) -> dict:
        p = _ffi.GvdbProfile()
        self._ok(self._lib.gvdb_profile_read(self._h, C.byref(p), 1 if reset else 0))
        return {f: getattr(p, f) for f, _ in p._fields_}


def measure_fp4_mma_rate(device: int = 0) -> dict:
    """gvdb_measure_fp4_mma_rate: the dense FP4 tensor-core rate of `device`, measured (TMAC/s and
    SM clocks per M128 x N128 x K64 MMA) — the roofline denominator of the tensor-core scan."""
    lib = _ffi.lib()
    t, c = C.c_double(0.0), C.c_double(0.0)
    raise_for_status(lib.gvdb_measure_fp4_mma_rate(device, C.byref(t), C.byref(c)), lib)
    return {"tmacs_per_s": t.value, "clk_per_mma": c.value}
