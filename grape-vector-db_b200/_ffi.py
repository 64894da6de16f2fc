"""ctypes declarations of the C ABI in include/gvdb.h (libgvdb.so).

Loading fails loudly when the CUDA library has not been built: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgvdb.so")

GVDB_OK = 0
GVDB_ERR_INDEX_NOT_BUILT = 1
GVDB_ERR_DIMENSION_MISMATCH = 2
GVDB_ERR_INVALID_VECTOR_DIMENSION = 3
GVDB_ERR_QUANTIZATION = 4
GVDB_ERR_INDEX = 5
GVDB_ERR_INVALID_ARGUMENT = 6
GVDB_ERR_NOT_IMPLEMENTED = 7
GVDB_NO_ID = 0xFFFFFFFFFFFFFFFF


class GvdbConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("dim", C.c_uint32), ("threshold", C.c_float),
                ("rescore_ratio", C.c_float), ("device", C.c_int32), ("flags", C.c_uint32),
                ("capacity_rows", C.c_uint64), ("row_base", C.c_uint64),
                ("window_first", C.c_uint64), ("window_count", C.c_uint64)]


GVDB_FLAG_ROW_WINDOW = 1


class GvdbStats(C.Structure):
    _fields_ = [("vector_count", C.c_uint64), ("rows", C.c_uint64), ("dimension", C.c_uint64),
                ("memory_usage", C.c_uint64), ("hbm_bytes", C.c_uint64),
                ("code_bytes_per_row", C.c_uint64)]


class GvdbProfile(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("scan_launches", C.c_uint64), ("scan_ms", C.c_double),
                ("scan_bytes", C.c_double), ("scan_pairs", C.c_double), ("select_ms", C.c_double),
                ("rescore_ms", C.c_double), ("topk_ms", C.c_double), ("prep_ms", C.c_double),
                ("flat_ms", C.c_double), ("merge_ms", C.c_double), ("tc_launches", C.c_uint64),
                ("tc_ms", C.c_double), ("tc_macs", C.c_double), ("tc_bytes", C.c_double),
                ("scatter_ms", C.c_double), ("optimistic_reruns", C.c_uint64),
                ("overflow_fallbacks", C.c_uint64), ("exchange_ms", C.c_double),
                ("exchange_wait_ms", C.c_double), ("sample_ms", C.c_double), ("dot_ms", C.c_double),
                ("dot_launches", C.c_uint64), ("dot_macs", C.c_double), ("ratio_fallback_queries", C.c_uint64)]


# every symbol include/gvdb.h declares: name -> (restype, argtypes)
_vp, _u32, _u64, _i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
SYMBOLS = {
    "gvdb_abi_version": (_u32, []),
    "gvdb_approx_dot": (_i32, [_vp, _vp, _u32, _vp]),
    "gvdb_measure_fp4_mma_rate": (_i32, [_i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "gvdb_last_error": (C.c_char_p, []),
    "gvdb_create": (_i32, [C.POINTER(GvdbConfig), C.POINTER(_vp)]),
    "gvdb_destroy": (None, [_vp]),
    "gvdb_add": (_i32, [_vp, _vp, _u64, C.POINTER(_u64)]),
    "gvdb_add_device": (_i32, [_vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    "gvdb_reserve": (_i32, [_vp, _u64]),
    "gvdb_remove": (_i32, [_vp, _u64, C.POINTER(_i32)]),
    "gvdb_clear": (_i32, [_vp]),
    "gvdb_len": (_u64, [_vp]),
    "gvdb_get_stats": (_i32, [_vp, C.POINTER(GvdbStats)]),
    "gvdb_save": (_i32, [_vp, C.c_char_p]),
    "gvdb_load": (_i32, [C.c_char_p, _i32, C.POINTER(_vp)]),
    "gvdb_quantize": (_i32, [_vp, _vp, _u64, _vp]),
    "gvdb_get_codes": (_i32, [_vp, _u64, _u64, _vp]),
    "gvdb_hamming": (_i32, [_vp, _vp, _u32, _vp]),
    "gvdb_rescore_count": (_u64, [_u64, C.c_float]),
    "gvdb_search_batch": (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp]),
    "gvdb_search_batch_device": (_i32, [_vp, _vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp]),
    "gvdb_flat_search_batch": (_i32, [_vp, _vp, _u32, _u32, _vp, _vp]),
    "gvdb_flat_search_batch_device": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp, _vp]),
    "gvdb_similarity_search_batch": (_i32, [_vp, _vp, _u32, _u32, C.c_float, _i32, _vp, _vp]),
    "gvdb_similarity_search_batch_device": (_i32, [_vp, _vp, _vp, _u32, _u32, C.c_float, _i32, _vp, _vp]),
    "gvdb_shard_record_bytes": (_u64, [_u32, _u32]),
    "gvdb_search_shard_device": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp]),
    "gvdb_search_shard_sliced_device": (_i32, [_vp, _vp, _vp, _u32, _u32, _u32, _vp]),
    "gvdb_shard_hist_bins": (_u32, [_vp]),
    "gvdb_shard_hist_device": (_i32, [_vp, _vp, _vp, _u32, _vp]),
    "gvdb_search_shard_ratio_device": (_i32, [_vp, _vp, _vp, _u32, _u64, _u32, _vp, _u32, _u32, _vp]),
    "gvdb_merge_shards_ratio_device": (_i32, [_vp, _vp, _u32, _vp, _u32, _u32, _vp, _vp]),
    "gvdb_search_shard_sliced_enqueue_device": (_i32, [_vp, _vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    "gvdb_search_shard_verify": (_i32, [_vp, _vp, C.POINTER(_i32)]),
    "gvdb_merge_shards_device": (_i32, [_vp, _vp, _u32, _vp, _u32, _u32, _u32, _vp, _vp]),
    "gvdb_stage1_device": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp]),
    "gvdb_rescore_keys_device": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp, _vp]),
    "gvdb_finish_owned_device": (_i32, [_vp, _vp, _vp, _vp, _u32, _u64, _u32, _u32, _u32, _vp, _vp]),
    "gvdb_export_rows_ipc": (_i32, [_vp, _vp]),
    "gvdb_attach_peer_rows_ipc": (_i32, [_vp, _u32, _u64, _u32, _vp]),
    "gvdb_rows_device_ptr": (_vp, [_vp]),
    "gvdb_attach_peer_rows_ptr": (_i32, [_vp, _u32, _u64, _u32, _vp]),
    "gvdb_search_batch_filtered": (_i32, [_vp, _vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    "gvdb_search_batch_filtered_device": (_i32, [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    "gvdb_flat_search_batch_filtered": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp, _vp]),
    "gvdb_exchange_create": (_i32, [_vp, _u32, _u32, _u64, _u32, _u32]),
    "gvdb_exchange_export_ipc": (_i32, [_vp, _vp]),
    "gvdb_exchange_attach_ipc": (_i32, [_vp, _vp]),
    "gvdb_exchange_mailbox_ptr": (_vp, [_vp]),
    "gvdb_exchange_attach_ptr": (_i32, [_vp, _vp]),
    "gvdb_search_exchange_device": (_i32, [_vp, _vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    "gvdb_exchange_status": (_i32, [_vp, _vp]),
    "gvdb_sparse_create": (_i32, [_i32, C.c_float, C.c_float, C.POINTER(_vp)]),
    "gvdb_sparse_destroy": (None, [_vp]),
    "gvdb_sparse_build": (_i32, [_vp, _u64, _u32, _vp, _vp, _vp, _vp]),
    "gvdb_sparse_average_document_length": (C.c_float, [_vp]),
    "gvdb_sparse_search_bm25_batch": (_i32, [_vp, _u32, _vp, _vp, _vp, _u32, _vp, _vp]),
    "gvdb_sparse_search_bm25_batch_device": (_i32, [_vp, _vp, _u32, _vp, _vp, _vp, _u32, _vp, _vp]),
    "gvdb_sparse_launches": (_u64, [_vp]),
    "gvdb_rrf_fusion_batch": (_i32, [_i32, _vp, _u32, _vp, _u32, _vp, _u32, _u32, C.c_float, _u32, _vp, _vp]),
    "gvdb_weighted_fusion_batch": (_i32, [_i32, _vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp, _u32, _u32, C.c_float, C.c_float, C.c_float,
                                           _i32, _u32, _vp, _vp]),
    "gvdb_weighted_fusion_batch_device": (_i32, [_i32, _vp, _vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp, _u32, _u32, C.c_float, C.c_float,
                                                  C.c_float, _i32, _u32, _vp, _vp]),
    "gvdb_rrf_fusion_batch_device": (_i32, [_i32, _vp, _vp, _u32, _vp, _u32, _vp, _u32, _u32, C.c_float, _u32, _vp, _vp]),
    "gvdb_profile_enable": (_i32, [_vp, _i32]),
    "gvdb_profile_read": (_i32, [_vp, _vp, _i32]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        if L.gvdb_abi_version() != 5:
            raise RuntimeError("libgvdb.so ABI version mismatch")
        _lib = L
    return _lib
