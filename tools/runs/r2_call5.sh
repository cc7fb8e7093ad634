#!/bin/bash
# GEMM parity incl. the multicast-pair prefill path, prefill A/B, launch list of one eager decode step
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_vs_reference.py -m gpu -q -x 2>&1 | tail -15 > $O/r2_tests5.log
for M in 512 1024 4096 8192; do
  echo "MC:    $(timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_ab.txt
  echo "no-MC: $(QS_GEMM_NO_MC=1 timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_ab.txt
done
QS_PROFILE_STEP=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline --kernel-reps 1 > $O/r2_launches_bench.log 2>&1
python tools/launches_summary.py $O/r2_launches.csv > $O/r2_launches_summary.txt 2>&1
