#!/bin/bash
# round-2 GPU call 3: full -m gpu suite, graph-timed bench, in-graph step timeline, ncu launch list + --set full captures of every hot kernel
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r2_tests3.log
FAST="--steps 20 --warmup 5 --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline"
python bench.py $FAST > $O/r2_bench3_fast.json 2> $O/r2_bench3.err
python tools/step_timeline.py > $O/r2_step_timeline.txt 2>&1
# launch list of one step (shares must agree with the bench kernel table)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline --kernel-reps 1 > $O/r2_launches_bench.log 2>&1
for op in gemm_qkv gemm_o gemm_gate_up gemm_down attention; do
  ncu --set full --clock-control none --import-source on -c 1 -s 4 -k regex:"gemm_kernel|decode_attention" -o $O/r2_ncu_$op -f python tools/run_op.py $op --reps 6 > $O/r2_ncu_$op.log 2>&1
  python tools/ncu_summary.py $O/r2_ncu_$op.ncu-rep > $O/r2_ncu_$op.txt 2>&1
done
rm -f $O/r2_ncu_gemm_qkv.ncu-rep $O/r2_ncu_gemm_o.ncu-rep $O/r2_ncu_gemm_down.ncu-rep   # keep gate_up + attention reports (64 MiB cap)
ls -la $O | tail -30 > $O/r2_call3_ls.txt
