#!/bin/bash
# split-epilogue ILP change: parity, graph-timed kernel table, per-CTA phase timeline
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused.py tests/test_gpu_decode_runner.py -m gpu -q -x 2>&1 | tail -8 > $O/r2_tests6.log
python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-refmodel --no-tp --no-cpu-baseline > $O/r2_bench6_fast.json 2> $O/r2_bench6.err
python tools/gemm_timeline.py > $O/r2_gemm_timeline.txt 2>&1
# the TP = 2 exact-mode graph run crashed on rank 1 ("unspecified launch failure"): look at one rank's shard on one GPU
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/tp_shard_single.py --exact --layers 2 > $O/r2_sanitizer_tp_exact.txt 2>&1
timeout 300 python tools/tp_shard_single.py --exact --layers 8 --graph > $O/r2_tp_exact_graph_single.txt 2>&1
