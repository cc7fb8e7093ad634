#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 600 python -m pytest tests/test_gpu_refmodel.py -m gpu -q -x -s -k "prompt_attention" 2>&1 | grep -v Warning | tail -60 > $O/r2_tests28.log
