#!/bin/bash
# g128 level-2 producer thread, W8A8 split-K at 128 tokens, unfused attention+quant A/B
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_vs_reference.py tests/test_gpu_fused.py tests/test_gpu_decode_runner.py -m gpu -q -x 2>&1 | tail -8 > $O/r2_tests10.log
FAST="--steps 20 --warmup 5 --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline"
python bench.py $FAST --precision w4a8kv4-g128 > $O/r2_bench10_config3.json 2> $O/r2_bench10_config3.err
python bench.py $FAST --model mistral-7b --precision w8a8kv8 --batch 128 > $O/r2_bench10_config4.json 2> $O/r2_bench10_config4.err
QS_UNFUSED_ATTN_QUANT=1 python bench.py $FAST > $O/r2_bench10_config2_unfused_attnq.json 2> $O/r2_bench10_unf.err
