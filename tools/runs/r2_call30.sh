#!/bin/bash
# ncu --set full of the two new kernels (one launch each), after warm-up launches
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -c 1 -s 5 -k regex:gemm_pair_kernel -o $O/r2_ncu_gemm_pair -f python tools/run_prefill_gemm.py 4096 > $O/r2_ncu_gemm_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -c 1 -s 8 -k regex:prefill_attention_kernel -o $O/r2_ncu_prefill_attention -f python tools/prefill_attn_smoke.py > $O/r2_ncu_prefill_attention.log 2>&1
