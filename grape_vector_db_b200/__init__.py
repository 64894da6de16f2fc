"""Importable alias of the package directory `grape-vector-db_b200/`.

The product lives in `grape-vector-db_b200/` (the directory name the project contract
uses; a hyphen cannot appear in a Python module name).  This two-line shim points the
package search path at that directory, so `import grape_vector_db_b200.index` loads
`grape-vector-db_b200/index.py`.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          "grape-vector-db_b200")]

from ._api import *  # noqa: E402,F401,F403
from ._api import __all__  # noqa: E402,F401
