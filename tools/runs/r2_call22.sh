#!/bin/bash
# cta_group::2 prefill GEMM: full GEMM parity (per-channel, g128, live reference), then A/B against the 128-token tiles
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_vs_reference.py -m gpu -q -x 2>&1 | tail -6 > $O/r2_tests22.log
grep -q "passed" $O/r2_tests22.log && ! grep -q "failed" $O/r2_tests22.log || exit 1
for M in 512 1024 2048 4096 8192; do
  echo "pair:    $(timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_pair_ab.txt
  echo "NT=128:  $(QS_GEMM_NO_PAIR=1 timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_pair_ab.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-tp --no-cpu-baseline > $O/r2_bench22_refmodel.json 2> $O/r2_bench22_refmodel.err
