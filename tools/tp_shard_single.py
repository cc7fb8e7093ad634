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
ap.add_argument("--no-pdl", action="store_true")
ap.add_argument("--trace", action="store_true", help="record kernel entry / dependency / exit stamps in PINNED host memory (survives a device fault)")
ap.add_argument("--model", default="qwen1.5-72b")
a = ap.parse_args()
from qserve_b200 import backend  # noqa: E402
from qserve_b200._lib import lib  # noqa: E402
backend.set_pdl(not a.no_pdl)
run = DecodeRunner(a.model, "w4a8kv4", 64, 1024, torch.device("cuda:0"), tp_rank=min(1, a.tp - 1), tp_size=a.tp, layers=a.layers, fused=True, tp_exact=a.exact, no_comm=True)
with torch.no_grad():
    for _ in range(2):
        tok = run.forward(run.tokens_in)
    torch.cuda.synchronize()
    print("eager ok", tok[:4].tolist())
    if a.graph:
        run.capture()
        import time
        trace = None
        if a.trace:
            cap = 1 << 16
            trace = torch.zeros(1 + 2 * cap, dtype=torch.int64).pin_memory()
            lib.qs_set_trace_buffer(trace.data_ptr(), cap)
        NAMES = {1: "gemm", 2: "attention", 3: "norm", 4: "quant", 5: "silu", 6: "rms_norm", 7: "add_norm", 8: "silu_quant", 9: "prefill"}

        def dump_tail():
            if trace is None:
                return
            n = int(trace[0])
            k = min(n, cap)
            print(f"trace: {n} records (cap {cap}); last 40:")
            for j in range(max(0, k - 40), k):
                tag = int(trace[1 + 2 * j]); t = int(trace[2 + 2 * j])
                print(f"  {j:6d} {NAMES.get(tag >> 8, tag >> 8):10s} phase {tag & 0xff} t={t}")
        import atexit
        atexit.register(dump_tail)
        for i in range(int(__import__("os").environ.get("QS_STEPS", "3"))):
            t0 = time.perf_counter()
            if trace is not None:
                trace[0] = 0  # one step per window
            run.step()
            try:
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print("FAULT in graph step", i, repr(e)[:200], flush=True)
                dump_tail()
                raise SystemExit(3)
            print(f"graph step {i} ok {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
        print("graph ok", run.tokens_out[:4].tolist())
