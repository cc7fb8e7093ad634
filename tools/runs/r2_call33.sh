#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 300 python tools/prefill_breakdown.py w4a8kv4-g128 --no-pdl > $O/r2_prefill_breakdown33.txt 2>&1
