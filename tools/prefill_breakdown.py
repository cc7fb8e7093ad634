#!/usr/bin/env python
"""Where the time of a BASELINE config-3 prompt step goes (8 x 1024 tokens, Llama-3-8B g128, reference model code over this backend):
kernel-time table from torch.profiler (CUPTI), 4 layers."""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qserve_b200 import refmodel  # noqa: E402
from qserve_b200.decode import DecodeRunner  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "w4a8kv4-g128"
if "--no-pdl" in sys.argv:  # with programmatic dependent launch a kernel's profiled duration includes its wait for the predecessor
    from qserve_b200 import backend
    backend.set_pdl(False)
L = 4
dev = torch.device("cuda:0")
run = DecodeRunner("llama-3-8b", prec, 64, 1024, dev, layers=L, fused=False, seed=3)
ref = refmodel.RefModel(run)
ref.use_prefill_attention("qserve_b200")
lens = [1024] * 8
toks = torch.randint(0, run.cfg.vocab, (sum(lens),), device=dev)
meta = ref.prefill_metadata(lens)
for _ in range(2):
    ref.prefill_logits(toks, lens, meta)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ref.prefill_logits(toks, lens, meta)
e1.record()
torch.cuda.synchronize()
print(f"{prec}: prompt step of {sum(lens)} tokens, {L} layers: {e0.elapsed_time(e1) / 3:.3f} ms")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    ref.prefill_logits(toks, lens, meta)
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name[:100]][0] += ev.device_time
        agg[ev.name[:100]][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"total kernel time {tot / 1e3:.3f} ms")
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {t / 1e3:8.3f} ms  {100 * t / tot:5.1f} %  x{n:<4d} {name}")
