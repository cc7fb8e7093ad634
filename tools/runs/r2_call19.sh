#!/bin/bash
# full -m gpu suite with the register-resident quant kernel and the single-pass attention quant tail, then the full default bench line
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2_tests19_all.log
grep -q "passed" $O/r2_tests19_all.log && ! grep -q "failed" $O/r2_tests19_all.log || exit 1
timeout 1200 python bench.py --steps 50 --warmup 5 > $O/r2_bench19_full.json 2> $O/r2_bench19_full.err
echo "full rc=$?" >> $O/r2_tests19_all.log
python __graft_entry__.py smoke > $O/r2_smoke19.txt 2>&1
