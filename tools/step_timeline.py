#!/usr/bin/env python
"""In-graph timeline of one decode step from the kernels' own %globaltimer stamps (qs_set_trace_buffer).
Per kernel: entry of block 0, dependency resolved (= the previous kernel has fully completed), exit of block 0."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qserve_b200 import backend  # noqa: E402
from qserve_b200._lib import lib  # noqa: E402
from qserve_b200.decode import DecodeRunner  # noqa: E402

NAMES = {1: "gemm", 2: "attention", 3: "norm", 4: "quant", 5: "silu", 6: "rms_norm", 7: "add_norm", 8: "silu_quant", 9: "prefill"}
fused = "--no-fused" not in sys.argv
backend.set_pdl("--no-pdl" not in sys.argv)
layers = 32
run = DecodeRunner("llama-3-8b", "w4a8kv4", 64, 1024, torch.device("cuda:0"), layers=layers, fused=fused)
run.capture()
for _ in range(3):
    run.step()
torch.cuda.synchronize()
cap = 4096
buf = torch.zeros(1 + 2 * cap, dtype=torch.int64, device="cuda")
lib.qs_set_trace_buffer(buf.data_ptr(), cap)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run.step(); e1.record()
torch.cuda.synchronize()
lib.qs_set_trace_buffer(None, 0)
b = buf.cpu().numpy()
n = int(b[0])
rec = b[1:1 + 2 * n].reshape(n, 2)
rec = rec[np.argsort(rec[:, 1], kind="stable")]
t0 = rec[0, 1]
print(f"step (events): {e0.elapsed_time(e1) * 1e3:.1f} us; {n} trace records; fused={fused}")
# group per kernel instance in order of entry
ent = [(int(r[0]) >> 8, (r[1] - t0) / 1e3) for r in rec if (int(r[0]) & 0xFF) == 0]
dep = [(int(r[0]) >> 8, (r[1] - t0) / 1e3) for r in rec if (int(r[0]) & 0xFF) == 1]
ext = [(int(r[0]) >> 8, (r[1] - t0) / 1e3) for r in rec if (int(r[0]) & 0xFF) == 2]
print("first two layers (us since first kernel entry):  kernel  entry  dep_resolved  [exit]")
k = min(len(ent), len(dep))
for i in range(min(24, k)):
    print(f"  {NAMES.get(dep[i][0], dep[i][0]):10s} entry {ent[i][1]:8.2f}   dep {dep[i][1]:8.2f}")
# dep-to-dep intervals = effective serialized cost of each kernel type (kernel i runs from dep[i] to dep[i+1])
import collections
agg = collections.defaultdict(list)
for i in range(k - 1):
    agg[NAMES.get(dep[i][0], dep[i][0])].append(dep[i + 1][1] - dep[i][1])
tot = dep[k - 1][1] - dep[0][1]
print("effective in-graph cost per kernel type (dep[i+1] - dep[i]):")
for name, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
    print(f"  {name:10s} n={len(v):4d}  mean {np.mean(v):7.2f} us  total {sum(v):8.1f} us  ({100 * sum(v) / tot:4.1f}%)")
print(f"  traced span {tot:.1f} us")
