#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -5 > $O/r2_tests44.log
