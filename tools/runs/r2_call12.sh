#!/bin/bash
# (1) parity of the register-resident norm / silu kernels, (2) which block of the full bench line crashed in call 11
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 600 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_fused.py tests/test_gpu_vs_reference.py tests/test_gpu_decode_runner.py tests/test_gpu_refmodel.py -m gpu -q -x 2>&1 | tail -8 > $O/r2_tests12.log
grep -q "passed" $O/r2_tests12.log && ! grep -q "failed" $O/r2_tests12.log || exit 1
timeout 300 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline > $O/r2_bench12_fast.json 2> $O/r2_bench12_fast.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-tp --no-cpu-baseline > $O/r2_bench12_refmodel.json 2> $O/r2_bench12_refmodel.err
echo "refmodel rc=$?" >> $O/r2_tests12.log
timeout 600 python bench.py --steps 20 --warmup 5 --tp-only > $O/r2_bench12_tp1.json 2> $O/r2_bench12_tp1.err
echo "tp1 rc=$?" >> $O/r2_tests12.log
timeout 600 python bench.py --impl reference-gpu --steps 10 --warmup 3 > $O/r2_bench12_refgpu.json 2> $O/r2_bench12_refgpu.err
echo "refgpu rc=$?" >> $O/r2_tests12.log
