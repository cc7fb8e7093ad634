"""qserve_backend.qgemm_w4a8_per_chn (kernels/csrc/qgemm/w4a8_per_chn/pybind.cpp:13-16)."""
from qserve_b200.backend import w4a8_per_chn_gemm_forward_cuda as gemm_forward_cuda  # noqa: F401
