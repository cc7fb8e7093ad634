#!/bin/bash
# 2 GPUs after the race fix and the fast row kernels: TP tests, driver-shaped line (peer all-reduce), NCCL and exact-mode TP records
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 600 python -m pytest tests/test_gpu_tp.py -m gpu -q -s 2>&1 | tail -12 > $O/r2_tests17_tp2.log
grep -q "passed" $O/r2_tests17_tp2.log && ! grep -q "failed" $O/r2_tests17_tp2.log || exit 1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_bench17_tp2_peer.json 2> $O/r2_bench17_tp2_peer.err
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --tp-only --tp-allreduce nccl > $O/r2_bench17_tp2_nccl.json 2> $O/r2_bench17_tp2_nccl.err
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --tp-only --tp-exact > $O/r2_bench17_tp2_exact.json 2> $O/r2_bench17_tp2_exact.err
