"""CPU oracle for the qserve_backend hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Every function here is a numpy restatement of one piece of the reference
(mit-han-lab/qserve @ de6a2ec) and cites the reference file:line it follows.
Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline legs of `bench.py`
may import this package.  The product path (`qserve_b200`, `qserve_backend`)
never imports it and fails loudly when the CUDA library is missing.

Parity pin status (see DESIGN.md "Oracle"):
  * weight packing (`oracle.w4a8.pack_*`) is PINNED against golden vectors
    produced by importing the reference's own Python packer
    (`W4A8OF16LinearDynamicInputScale.from_linear`, tests/golden/make_golden_pack.py).
  * kernel arithmetic (GEMM epilogues, act-quant, norm, silu, KV quant, decode
    attention) has no golden vector or test in the reference ("parity
    unpinned" by the reference itself).  It is pinned here
    against outputs of the UNMODIFIED reference CUDA kernels compiled for
    sm_100a (`oracle/build_ref.py` -> `oracle/_ref/`) and run on a B200 by
    `tests/golden/make_golden_ref_gpu.py`; the captured vectors are committed
    under tests/golden/ and checked by the CPU test-suite.
"""
