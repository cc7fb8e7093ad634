#!/bin/bash
# where should the CTA-pair prefill kernel take over?  M = 192 .. 512, pair forced vs 128-token tiles
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
rm -f $O/r2_pair_threshold.txt
for M in 192 256 320 384 448 512 640 768; do
  echo "pair:    $(QS_GEMM_PAIR_MIN_M=1 timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_pair_threshold.txt
  echo "NT=128:  $(QS_GEMM_NO_PAIR=1 timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_pair_threshold.txt
done
