#!/bin/bash
# pair vs 128-token tiles on the narrow layers (few channel tiles): where does the pair kernel stop paying?
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
rm -f $O/r2_pair_threshold2.txt
for cfg in "256 4096 4096" "512 4096 4096" "1024 4096 4096" "2048 4096 4096" "512 4096 14336" "1024 4096 14336" "512 6144 4096" "1024 6144 4096" "256 6144 4096"; do
  set -- $cfg
  echo "pair:    $(QS_N=$2 QS_K=$3 QS_GEMM_PAIR_MIN_M=1 timeout 120 python tools/run_prefill_gemm.py $1 2>&1 | tail -1)" >> $O/r2_pair_threshold2.txt
  echo "NT=128:  $(QS_N=$2 QS_K=$3 QS_GEMM_NO_PAIR=1 timeout 120 python tools/run_prefill_gemm.py $1 2>&1 | tail -1)" >> $O/r2_pair_threshold2.txt
done
