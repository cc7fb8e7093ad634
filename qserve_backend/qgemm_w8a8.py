"""qserve_backend.qgemm_w8a8 (kernels/csrc/qgemm/w8a8/pybind.cpp:13-17)."""
from qserve_b200.backend import w8a8_gemm_forward_cuda  # noqa: F401
