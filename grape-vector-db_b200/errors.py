"""VectorDbError variants this path emits (/root/reference/src/types.rs:859-920), one Python
exception per variant, mapped 1:1 from the C ABI's gvdb_status."""
from __future__ import annotations


class VectorDbError(Exception):
    """Base of the mirrored error enum."""


class IndexNotBuilt(VectorDbError):           # src/index.rs:213,621-623
    pass


class DimensionMismatch(VectorDbError):       # src/index.rs:590-594
    def __init__(self, expected=None, actual=None, msg=None):
        self.expected, self.actual = expected, actual
        super().__init__(msg or f"dimension mismatch: expected {expected}, actual {actual}")


class InvalidVectorDimension(VectorDbError):  # src/quantization.rs:131-133,335
    pass


class QuantizationError(VectorDbError):       # src/quantization.rs:158-162
    pass


class IndexError_(VectorDbError):             # VectorDbError::IndexError(String)
    pass


class ConfigError(VectorDbError):
    pass


class NotImplementedError_(VectorDbError):
    pass


_BY_STATUS = {
    1: IndexNotBuilt,
    2: DimensionMismatch,
    3: InvalidVectorDimension,
    4: QuantizationError,
    5: IndexError_,
    6: ConfigError,
    7: NotImplementedError_,
}


def raise_for_status(status: int, lib) -> None:
    if status == 0:
        return
    msg = lib.gvdb_last_error()
    msg = msg.decode("utf-8", "replace") if msg else ""
    cls = _BY_STATUS.get(status, VectorDbError)
    if cls is DimensionMismatch:
        raise DimensionMismatch(msg=msg)
    raise cls(msg)
