#!/bin/bash
# new GPU tests (prefill attention, refmodel over it, GEMM with the pair kernel), the flash-attn golden, bench refmodel block
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 900 python -m pytest tests/test_gpu_prefill_attention.py tests/test_gpu_refmodel.py tests/test_gpu_gemm.py tests/test_gpu_vs_reference.py -m gpu -q -x -s 2>&1 | tail -25 > $O/r2_tests27.log
timeout 120 python tests/golden/make_golden_prefill_attn.py $O/prefill_attn_flash.npz > $O/r2_golden27.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-tp --no-cpu-baseline > $O/r2_bench27.json 2> $O/r2_bench27.err
