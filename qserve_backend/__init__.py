"""Drop-in replacement for the reference's `qserve_backend` package (kernels/setup.py:158-245).

The seven sub-modules export exactly the functions the reference's pybind11 extensions export
(SURVEY.md section 8b); all of them forward to `qserve_b200.backend`, the host-side mirror over the sm_100a C ABI.
"""
from . import activation_ops, fused_attention, fused_kernels, layernorm_ops  # noqa: F401
from . import qgemm_w4a8_per_chn, qgemm_w4a8_per_group, qgemm_w8a8  # noqa: F401
