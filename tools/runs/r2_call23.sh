#!/bin/bash
# persistent cta_group::2 prefill GEMM: smoke (oracle + vs the 128-token tiles), then A/B timing
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 300 python tools/pair_smoke.py > $O/r2_pair_smoke23.txt 2>&1 || { echo "smoke failed/timeout rc=$?" >> $O/r2_pair_smoke23.txt; exit 1; }
for M in 512 1024 2048 4096 8192; do
  echo "pair:    $(timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_pair_ab23.txt
done
