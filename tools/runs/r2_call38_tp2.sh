#!/bin/bash
# 2 GPUs: the tensor-parallel GPU tests after this round's changes
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -x 2>&1 | grep -v Warning | tail -8 > $O/r2_tests38_tp2.log
