#!/bin/bash
# verify the attention ring-initialisation race fix: parity, 2 x 150 full-size Qwen steps (multi-wave attention), config 4 (multi-wave), TP record at N = 1
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_fused.py tests/test_gpu_vs_reference.py -m gpu -q -x 2>&1 | tail -5 > $O/r2_tests16.log
grep -q "passed" $O/r2_tests16.log && ! grep -q "failed" $O/r2_tests16.log || exit 1
QS_STEPS=150 timeout 500 python tools/tp_shard_single.py --tp 1 --layers 40 --graph > $O/r2_qwen_l40_fix_a.txt 2>&1; echo "rc=$?" >> $O/r2_qwen_l40_fix_a.txt
QS_STEPS=150 timeout 500 python tools/tp_shard_single.py --tp 1 --layers 40 --graph --no-pdl > $O/r2_qwen_l40_fix_b.txt 2>&1; echo "rc=$?" >> $O/r2_qwen_l40_fix_b.txt
timeout 400 python bench.py --steps 50 --warmup 5 --tp-only > $O/r2_bench16_tp1.json 2> $O/r2_bench16_tp1.err; echo "tp1 rc=$?" >> $O/r2_tests16.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline --model mistral-7b --precision w8a8kv8 --batch 128 > $O/r2_bench16_config4.json 2> $O/r2_bench16_config4.err; echo "config4 rc=$?" >> $O/r2_tests16.log
