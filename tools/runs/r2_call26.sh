#!/bin/bash
# (1) pair kernel with 8 epilogue warps: smoke + timing + timeline; (2) first run of the tcgen05 prefill attention
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 300 python tools/pair_smoke.py > $O/r2_pair_smoke26.txt 2>&1 || echo "smoke failed/timeout rc=$?" >> $O/r2_pair_smoke26.txt
rm -f $O/r2_prefill_pair_ab26.txt
for M in 512 1024 2048 4096 8192; do
  echo "pair:    $(timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_pair_ab26.txt
done
timeout 200 python tools/pair_timeline.py > $O/r2_pair_timeline26.txt 2>&1
timeout 300 python tools/prefill_attn_smoke.py > $O/r2_prefill_attn26.txt 2>&1 || echo "attn smoke rc=$?" >> $O/r2_prefill_attn26.txt
