"""Loader for the UNMODIFIED reference CUDA extensions built by oracle/build_ref.py (lives in oracle/refmods.py)."""
from oracle.refmods import REF_DIR, load  # noqa: F401
