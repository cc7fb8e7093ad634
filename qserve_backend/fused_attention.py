"""qserve_backend.fused_attention (kernels/csrc/fused_attention/fused_attention.cpp:243-256)."""
from qserve_b200.backend import (  # noqa: F401
    apply_bias_rope_update_kv_cache,
    compute_padding_offsets,
    single_query_attention,
)
