#!/usr/bin/env python
"""Phase timeline of the cta_group::2 prefill GEMM (persistent CTA pairs) from %globaltimer stamps: a lone pair and the full chip."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qserve_backend.qgemm_w4a8_per_chn as op  # noqa: E402
from qserve_b200 import backend  # noqa: E402
from qserve_b200._lib import lib  # noqa: E402

NAMES = ["entry", "setup_done", "zero_seen(mma0)", "afull0_seen", "xfull0_seen", "mma0 at g=8", "tile1 zero_seen", "dfull commit t0", "unpack done t0",
         "dfull seen t0", "drain done t0", "zero refilled", "exit", "mma0 at g=14", "unpack it=8 done"]
backend.set_pdl(False)
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
for (M, N, K) in [(512, 256, 4096), (512, 256, 16384), (4096, 28672, 4096)]:
    a = torch.randint(-127, 128, (M, K), dtype=torch.int8, generator=g).to(dev)
    w = torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, generator=g).to(dev)
    s1 = torch.full((N,), 0.01, dtype=torch.half, device=dev)
    sa = torch.full((M,), 0.01, dtype=torch.half, device=dev)
    out = torch.empty((M, N), dtype=torch.half, device=dev)
    for _ in range(3):
        op.gemm_forward_cuda(a, w, s1, sa, s1, sa, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        op.gemm_forward_cuda(a, w, s1, sa, s1, sa, out)
    e1.record()
    torch.cuda.synchronize()
    prof = torch.zeros(512 * 16, dtype=torch.int64, device=dev)
    lib.qs_gemm_set_profile_buffer(prof.data_ptr())
    op.gemm_forward_cuda(a, w, s1, sa, s1, sa, out)
    torch.cuda.synchronize()
    lib.qs_gemm_set_profile_buffer(None)
    p = prof.cpu().numpy().reshape(-1, 16)
    p = p[p[:, 0] > 0].astype(np.float64)
    t0 = p[:, 0].min()
    print(f"== M={M} N={N} K={K}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per launch (no PDL), {len(p)} CTAs, {K // 256} stages per tile")
    for j, nm in enumerate(NAMES):
        col = p[:, j]
        ok = col > 0
        if ok.any():
            rel = (col[ok] - t0) / 1e3
            print(f"   {j:2d} {nm:18s} n={ok.sum():4d}  min {rel.min():8.2f}  med {np.median(rel):8.2f}  max {rel.max():8.2f} us")
