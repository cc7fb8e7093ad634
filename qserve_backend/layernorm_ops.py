"""qserve_backend.layernorm_ops (kernels/csrc/layernorm.cpp:47-72)."""
from qserve_b200.backend import (  # noqa: F401
    invoke_dequant_add_residual_rms_norm_quant,
    rms_norm,
    rms_norm_general,
    rms_norm_general_fuse_sum,
)
