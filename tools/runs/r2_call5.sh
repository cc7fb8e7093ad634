#!/bin/bash
# launch list of exactly one eager decode step (cudaProfilerStart/Stop range)
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
QS_PROFILE_STEP=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline --kernel-reps 1 > $O/r2_launches_bench.log 2>&1
python tools/launches_summary.py $O/r2_launches.csv > $O/r2_launches_summary.txt 2>&1
