#!/usr/bin/env python
"""How fast can ONE CTA of the tiled GEMM consume 256-K stages when the data is L2 resident?  (N = 128: a single tile,
split forced to 1, K long.)  Separates the per-CTA unpack -> TMEM -> MMA -> commit chain from HBM effects."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qserve_b200._lib import lib  # noqa: E402
import qserve_backend.qgemm_w4a8_per_chn as op  # noqa: E402
import qserve_backend.qgemm_w8a8 as op8  # noqa: E402

dev = torch.device("cuda:0")
for M in (64, 16):
    for N, K in ((128, 16384), (128, 4096), (148 * 128, 4096), (296 * 128, 2048)):
        g = torch.Generator(device="cpu").manual_seed(1)
        a = torch.randint(-127, 128, (M, K), dtype=torch.int8, generator=g).to(dev)
        w = torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, generator=g).to(dev)
        w8 = torch.randint(-128, 128, (N, K), dtype=torch.int8, generator=g).to(dev)
        s1 = torch.full((N,), 0.01, dtype=torch.half, device=dev)
        sa = torch.full((M,), 0.01, dtype=torch.half, device=dev)
        out = torch.empty((M, N), dtype=torch.half, device=dev)
        lib.qs_gemm_force_split(1)
        res = []
        for name, fn in (("w4", lambda: op.gemm_forward_cuda(a, w, s1, sa, s1, sa, out)), ("w8", lambda: op8.w8a8_gemm_forward_cuda(a, w8, s1, sa, out))):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 50
            res.append(f"{name}: {us:7.2f} us = {us / (K / 256):.3f} us/stage")
        lib.qs_gemm_force_split(0)
        print(f"M={M:3d} N={N:6d} K={K:6d} ({N // 128} CTAs, {K // 256} stages each)  " + "   ".join(res))
