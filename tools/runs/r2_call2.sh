#!/bin/bash
# round-2 GPU call: full bench line + live-reference tests + attention/L2 experiments
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/r2_bench1.json 2> $O/r2_bench1.err
python -m pytest tests/test_gpu_vs_reference.py -m gpu -q 2>&1 | tail -30 > $O/r2_tests2.log
FAST="--steps 20 --warmup 5 --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline"
python bench.py $FAST > $O/r2_expA_base.json 2> $O/r2_expA.err
python bench.py $FAST --l2-prefetch > $O/r2_expB_pf.json 2> $O/r2_expB.err
QS_ATTN_SPLIT=2 python bench.py $FAST > $O/r2_expC_split2.json 2> $O/r2_expC.err
QS_ATTN_SPLIT=2 python bench.py $FAST --l2-prefetch > $O/r2_expD_both.json 2> $O/r2_expD.err
