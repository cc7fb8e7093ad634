"""qserve_backend.fused_kernels (kernels/csrc/fused.cpp:47-70)."""
from qserve_b200.backend import (  # noqa: F401
    invoke_dequant,
    invoke_dequant_add_residual,
    invoke_quant,
    invoke_quant_fuse_sum,
)
