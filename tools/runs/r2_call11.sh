#!/bin/bash
# sanity after the reverted experiments, unfused attention+quant A/B, then the full default bench line
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -5 > $O/r2_tests11.log
grep -q "passed" $O/r2_tests11.log && ! grep -q "failed" $O/r2_tests11.log || exit 1
FAST="--steps 20 --warmup 5 --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline"
QS_UNFUSED_ATTN_QUANT=1 timeout 300 python bench.py $FAST > $O/r2_bench11_unfused_attnq.json 2> $O/r2_bench11_unf.err
timeout 300 python bench.py $FAST > $O/r2_bench11_fused_attnq.json 2> $O/r2_bench11_fus.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench11_full.json 2> $O/r2_bench11_full.err
