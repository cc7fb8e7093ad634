#!/bin/bash
# 4 GPUs: the driver-shaped bench line (DP headline + Qwen1.5-72B TP = 4 with the fused peer all-reduce), then NCCL for comparison
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612"
timeout 900 $TR bench.py --gpus 4 --steps 20 --warmup 5 > $O/r2_bench_tp4_peer.json 2> $O/r2_bench_tp4_peer.err
timeout 600 $TR bench.py --gpus 4 --steps 20 --warmup 5 --tp-only --tp-allreduce nccl > $O/r2_bench_tp4_nccl.json 2> $O/r2_bench_tp4_nccl.err
