"""qserve_b200 -- Blackwell (sm_100a) W4A8KV4 kernel library behind QServe's operator API.

`qserve_b200.backend` mirrors the reference's `qserve_backend` functions over the C ABI of
`libqserve_b200.so` (include/qserve_b200.h); the top-level `qserve_backend` package re-exports them under
the reference's module names.  Importing `qserve_b200.backend` (or `qserve_backend`) loads the shared library and
fails loudly if it has not been built (`python -m qserve_b200.build`); there is no CPU fallback.
"""
__all__ = ["backend", "build", "decode", "tp"]
__version__ = "0.1.0"
