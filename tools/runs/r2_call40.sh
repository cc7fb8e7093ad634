#!/bin/bash
# full GPU suite + smoke + full bench after the prompt-path changes (8 unpack warps, silu grid-stride, RoPE append per kv head, robust TP record)
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -8 > $O/r2_tests40.log
timeout 300 python __graft_entry__.py smoke > $O/r2_smoke40.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench40.json 2> $O/r2_bench40.err
