#!/bin/bash
# the driver's 8-GPU command (fewer steps): data-parallel headline + Qwen1.5-72B TP = 8 record
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err
echo "rc=$?" >> $O/r2_bench_n8.err
