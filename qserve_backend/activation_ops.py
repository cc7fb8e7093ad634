"""qserve_backend.activation_ops (kernels/csrc/activation.cpp:25-39)."""
from qserve_b200.backend import (  # noqa: F401
    gelu_fast,
    gelu_new,
    invoke_dequant_silu_and_mul_quant,
    silu_and_mul,
)
