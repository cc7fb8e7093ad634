#!/bin/bash
# round-2 GPU call (2 GPUs): TP parity tests over NCCL, then the Qwen1.5-72B TP=2 record with the fused peer all-reduce and with NCCL
cd "$(dirname "$0")/../.."
O=gpurun_out
python -c "import qserve_backend" 2>/dev/null || python -m qserve_b200.build > $O/r2_rebuild.log 2>&1
for nt in 128 256; do for M in 1024 4096; do
  echo "NT=$nt no-MC: $(QS_GEMM_NO_MC=1 QS_FORCE_NT=$nt timeout 120 python tools/run_prefill_gemm.py $M 2>&1 | tail -1)" >> $O/r2_prefill_nt.txt
done; done
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -s 2>&1 | tail -40 > $O/r2_tests_tp2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_bench_tp2_peer.json 2> $O/r2_bench_tp2_peer.err
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --tp-only --tp-allreduce nccl > $O/r2_bench_tp2_nccl.json 2> $O/r2_bench_tp2_nccl.err
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --tp-only --tp-exact > $O/r2_bench_tp2_exact.json 2> $O/r2_bench_tp2_exact.err
