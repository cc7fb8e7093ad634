#!/bin/bash
# first contact with the cta_group::2 prefill GEMM
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 120 python tools/pair_smoke.py > $O/r2_pair_smoke.txt 2>&1
echo "rc=$?" >> $O/r2_pair_smoke.txt
nvidia-smi --query-gpu=name --format=csv,noheader >> $O/r2_pair_smoke.txt 2>&1
