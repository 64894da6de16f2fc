"""Public names of the package (re-exported by the import shim)."""
from .errors import (ConfigError, DimensionMismatch, IndexError_, IndexNotBuilt,  # noqa: F401
                     InvalidVectorDimension, NotImplementedError_, QuantizationError,
                     VectorDbError)
from .index import NO_ID, GpuIndex, measure_fp4_mma_rate  # noqa: F401
from .sparse import GpuSparseIndex  # noqa: F401
from .hybrid import HybridSearcher, rrf_fusion_batch, weighted_fusion_batch  # noqa: F401

__all__ = ["GpuIndex", "measure_fp4_mma_rate", "GpuSparseIndex", "HybridSearcher", "rrf_fusion_batch", "weighted_fusion_batch", "NO_ID", "VectorDbError", "IndexNotBuilt", "DimensionMismatch",
           "InvalidVectorDimension", "QuantizationError", "IndexError_", "ConfigError",
           "NotImplementedError_"]
