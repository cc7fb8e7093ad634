"""qserve_b200 -- Blackwell (sm_100a) W4A8KV4 kernel library behind QServe's operator API.

`qserve_b200.backend` mirrors the reference's `qserve_backend` functions over the C ABI of
`libqserve_b200.so` (include/qserve_b200.h); the top-level `qserve_backend` package re-exports them under
the reference's module names.  Importing this package loads the shared library and fails if it is missing.
"""
from . import _lib  # noqa: F401  (loads libqserve_b200.so; raises ImportError when it has not been built)

__all__ = ["backend"]
__version__ = "0.1.0"
