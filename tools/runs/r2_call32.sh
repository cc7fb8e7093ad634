#!/bin/bash
# silu_and_mul grid-stride + prefill append per (token, kv head) + 8 unpack warps in the pair GEMM: parity tests, then the prompt-step breakdown
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_attention.py tests/test_gpu_vs_reference.py tests/test_gpu_gemm.py tests/test_gpu_refmodel.py -m gpu -q -x 2>&1 | grep -v Warning | tail -8 > $O/r2_tests32.log
timeout 300 python tools/pair_smoke.py > $O/r2_pair_smoke32.txt 2>&1
timeout 300 python tools/prefill_breakdown.py w4a8kv4-g128 > $O/r2_prefill_breakdown32.txt 2>&1
for M in 4096; do
  echo "per-chn pair: $(timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_breakdown32.txt
done
