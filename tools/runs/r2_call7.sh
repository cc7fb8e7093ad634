#!/bin/bash
# boundary-cost micro-benchmark (graph PDL / plain / flag chains), config 3 / config 4 decode steps of the fused runner
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 120 ./tools/ubench/sync_costs > $O/r2_sync_costs2.txt 2>&1
FAST="--steps 20 --warmup 5 --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline"
python bench.py $FAST --precision w4a8kv4-g128 > $O/r2_bench_config3_g128.json 2> $O/r2_bench_config3.err
python bench.py $FAST --model mistral-7b --precision w8a8kv8 --batch 128 > $O/r2_bench_config4.json 2> $O/r2_bench_config4.err
