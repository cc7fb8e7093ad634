"""Loader for the UNMODIFIED reference CUDA extensions built by oracle/build_ref.py into oracle/_ref/ (test infrastructure)."""
import importlib.util
import os

import torch  # noqa: F401  (the extensions link against libtorch)

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def load(name: str):
    path = os.path.join(REF_DIR, f"ref_{name}.so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(f"ref_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
