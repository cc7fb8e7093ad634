#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | grep -v Warning | tail -6 > $O/r2_tests43.log
