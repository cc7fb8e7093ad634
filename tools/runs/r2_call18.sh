#!/bin/bash
# W8A8: even ring depths (NT=64: AS 3 -> 2; NT=128 narrow layers: 2-deep weight ring + split-K): parity first, then config 4, then the regression test + full suite
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -k "w8" 2>&1 | tail -5 > $O/r2_tests18_w8.log
grep -q "passed" $O/r2_tests18_w8.log && ! grep -q "failed" $O/r2_tests18_w8.log || exit 1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2_tests18_all.log
grep -q "passed" $O/r2_tests18_all.log && ! grep -q "failed" $O/r2_tests18_all.log || exit 1
timeout 400 python bench.py --steps 50 --warmup 5 --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline --model mistral-7b --precision w8a8kv8 --batch 128 > $O/r2_bench18_config4.json 2> $O/r2_bench18_config4.err
