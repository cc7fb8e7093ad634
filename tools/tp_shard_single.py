#!/usr/bin/env python
"""One rank's shard of the tensor-parallel Qwen1.5-72B step on ONE GPU with the collectives skipped (no_comm): lets compute-sanitizer /
ncu look at the TP-only kernels (row_absmax, invoke_quant_given_amax, the sharded GEMM shapes) without a second process.
usage: [compute-sanitizer --tool memcheck] python tools/tp_shard_single.py [--exact] [--layers N] [--graph]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qserve_b200.decode import DecodeRunner  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--exact", action="store_true")
ap.add_argument("--layers", type=int, default=2)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--tp", type=int, default=2)
ap.add_argument("--model", default="qwen1.5-72b")
a = ap.parse_args()
run = DecodeRunner(a.model, "w4a8kv4", 64, 1024, torch.device("cuda:0"), tp_rank=min(1, a.tp - 1), tp_size=a.tp, layers=a.layers, fused=True, tp_exact=a.exact, no_comm=True)
with torch.no_grad():
    for _ in range(2):
        tok = run.forward(run.tokens_in)
    torch.cuda.synchronize()
    print("eager ok", tok[:4].tolist())
    if a.graph:
        run.capture()
        for _ in range(3):
            run.step()
        torch.cuda.synchronize()
        print("graph ok", run.tokens_out[:4].tolist())
