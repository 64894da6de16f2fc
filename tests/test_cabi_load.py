"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/gvdb.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gvdb.h")).read()
    return sorted(set(re.findall(r"GVDB_API[^;(]*?\b(gvdb_\w+)\s*\(", hdr)))


def test_header_symbols_all_exported(built):
    from grape_vector_db_b200 import _ffi
    names = _declared_symbols()
    assert len(names) >= 20
    lib = C.CDLL(_ffi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"libgvdb.so does not export {n}"
    # and the ctypes table covers exactly the header
    assert sorted(_ffi.SYMBOLS) == names


def test_abi_version_and_rescore_count(built):
    from grape_vector_db_b200 import _ffi
    L = _ffi.lib()
    assert L.gvdb_abi_version() == 5
    assert L.gvdb_rescore_count(10_000, 0.1) == 1000
    assert L.gvdb_rescore_count(7, 0.1) == 0
    assert L.gvdb_rescore_count(10, 2.0) == 10
    from oracle import oracle
    for n in (1, 39, 1000, 16_777_217, 100_000_000):
        for r in (0.1, 0.05, 4e-5, 1.0):
            assert L.gvdb_rescore_count(n, r) == oracle.rescore_count(n, r)


def test_config_validation_without_gpu(built):
    from grape_vector_db_b200 import _ffi
    L = _ffi.lib()
    cfg = _ffi.GvdbConfig()
    h = C.c_void_p()
    cfg.struct_size = 4          # wrong size -> ConfigError, before any CUDA call
    cfg.dim = 8
    assert L.gvdb_create(C.byref(cfg), C.byref(h)) == _ffi.GVDB_ERR_INVALID_ARGUMENT
    cfg.struct_size = C.sizeof(_ffi.GvdbConfig)
    cfg.dim = 0
    assert L.gvdb_create(C.byref(cfg), C.byref(h)) == _ffi.GVDB_ERR_INVALID_VECTOR_DIMENSION
    assert b"dimension" in L.gvdb_last_error()


def test_no_cpu_fallback(built):
    """Without a CUDA device the product must refuse to run, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import grape_vector_db_b200 as gv
    with pytest.raises(gv.IndexError_, match="no usable CUDA device"):
        gv.GpuIndex(768)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "grape-vector-db_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".rs")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower().replace("test oracle", ""), f"{f} mentions the oracle"


def test_sparse_and_fusion_entry_points_refuse_without_gpu(built):
    """The sparse side and the fusion have no CPU path either."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import grape_vector_db_b200 as gv
    with pytest.raises(gv.IndexError_, match="no usable CUDA device"):
        gv.GpuSparseIndex()
    with pytest.raises(gv.IndexError_, match="no usable CUDA device"):
        gv.rrf_fusion_batch(np.array([[1, 2, 3]], dtype=np.uint64), np.array([[2, 1, 4]], dtype=np.uint64))


def test_allow_bits_packing():
    """Row r of an allow-list lands in bit (r % 32) of word r / 32 (the tombstone bitmap's layout)."""
    import numpy as np
    from grape_vector_db_b200.index import pack_allow_bits
    assert pack_allow_bits([0, 5, 31, 32, 69], 70).tolist() == [0x80000021, 0x1, 0x20]
    mask = np.zeros(70, dtype=bool)
    mask[[0, 5, 31, 32, 69]] = True
    assert pack_allow_bits(mask, 70).tolist() == [0x80000021, 0x1, 0x20]
    assert pack_allow_bits([], 0).size == 0
