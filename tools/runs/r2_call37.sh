#!/bin/bash
# pair kernel drain with cvt.rn.f32.s32: exactness + timing
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 300 python tools/pair_smoke.py > $O/r2_pair_smoke37.txt 2>&1
rm -f $O/r2_prefill_pair_ab37.txt
for M in 1024 4096 8192; do
  echo "per-chn pair: $(timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_pair_ab37.txt
  echo "g128 pair:    $(QS_G128=1 timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_pair_ab37.txt
done
