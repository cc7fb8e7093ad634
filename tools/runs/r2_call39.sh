#!/bin/bash
# decode GEMM epilogue with cvt.rn.f32.s32: GEMM parity + quick bench (kernel table, headline)
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | grep -v Warning | tail -3 > $O/r2_tests39.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline > $O/r2_bench39.json 2> $O/r2_bench39.err
