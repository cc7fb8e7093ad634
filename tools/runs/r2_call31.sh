#!/bin/bash
# g128 prefill GEMM rate and the breakdown of the config-3 prompt step
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
rm -f $O/r2_prefill_g128.txt
for M in 1024 4096 8192; do
  echo "g128 pair:   $(QS_G128=1 timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_g128.txt
  echo "g128 NT=128: $(QS_G128=1 QS_GEMM_NO_PAIR=1 timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_g128.txt
done
timeout 300 python tools/prefill_breakdown.py w4a8kv4-g128 > $O/r2_prefill_breakdown.txt 2>&1
timeout 300 python tools/prefill_breakdown.py w4a8kv4 >> $O/r2_prefill_breakdown.txt 2>&1
