#!/usr/bin/env python
"""Smallest possible check of the cta_group::2 prefill GEMM (one pair tile, then a few more shapes) against the oracle's INT32 accumulators."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qserve_backend.qgemm_w4a8_per_chn as op  # noqa: E402
from oracle import w4a8  # noqa: E402

dev = torch.device("cuda:0")
for (M, N, K) in [(512, 256, 512), (512, 512, 1024), (1000, 512, 1152), (2048, 768, 512)]:
    rng = np.random.default_rng(M + N + K)
    q, qw, s1, s1z = w4a8.synth_per_channel(rng, N, K)
    aq = rng.integers(-127, 128, size=(M, K), dtype=np.int8)
    sa = rng.uniform(0.01, 0.05, size=M).astype(np.float16)
    asum = rng.uniform(-1, 1, size=M).astype(np.float16)
    out_o, acc_o = w4a8.gemm_w4a8_per_chn(aq, qw, s1, sa, s1z, asum, return_acc=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = torch.full((M, N), float("nan"), dtype=torch.half, device=dev)
    acc = torch.zeros((M, N), dtype=torch.int32, device=dev)
    op.gemm_forward_cuda(t(aq), t(qw), t(s1), t(sa), t(s1z), t(asum), out, _acc_out=acc)
    torch.cuda.synchronize()
    a = acc.cpu().numpy()
    bad = int((a != acc_o).sum())
    print(f"M={M} N={N} K={K}: acc mismatches {bad} / {a.size}; fp16 identical {bool(np.array_equal(out.cpu().numpy().view(np.uint16), out_o.view(np.uint16)))}", flush=True)
    if bad:
        idx = np.argwhere(a != acc_o)
        print("   first mismatches (token, channel):", idx[:6].tolist(), "got", a[tuple(idx[0])], "want", acc_o[tuple(idx[0])])
        print("   mismatch token range", idx[:, 0].min(), idx[:, 0].max(), "channel range", idx[:, 1].min(), idx[:, 1].max())

# ---- large shapes (several tiles per persistent CTA pair, ragged M): the pair kernel against the verified 128-token-tile kernel, whole tensors ----
import qserve_backend.qgemm_w4a8_per_group as opg  # noqa: E402
from qserve_b200._lib import lib  # noqa: E402

g = torch.Generator(device="cpu").manual_seed(11)
for (M, N, K) in [(4096, 4096, 4096), (5000, 2560, 1536), (8192, 28672, 4096), (777, 14336, 640)]:
    a = torch.randint(-127, 128, (M, K), dtype=torch.int8, generator=g).to(dev)
    qw = torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, generator=g).to(dev)
    ws = (torch.rand(N, generator=g) * 0.01 + 0.001).half().to(dev)
    wz = (torch.rand(N, generator=g) - 0.5).half().to(dev)
    sa = (torch.rand(M, generator=g) * 0.05 + 0.01).half().to(dev)
    asum = (torch.rand(M, generator=g) - 0.5).half().to(dev)
    z2 = torch.randint(-100, 0, (K // 128, N), dtype=torch.int8, generator=g).to(dev)
    s2 = torch.randint(1, 9, (K // 128, N), dtype=torch.int8, generator=g).to(dev)
    res = {}
    for nt in (0, 128):
        lib.qs_gemm_force_tile_tokens(nt)
        o1 = torch.full((M, N), float("nan"), dtype=torch.half, device=dev)
        c1 = torch.zeros((M, N), dtype=torch.int32, device=dev)
        op.gemm_forward_cuda(a, qw, ws, sa, wz, asum, o1, _acc_out=c1)
        o2 = torch.full((M, N), float("nan"), dtype=torch.half, device=dev)
        c2 = torch.zeros((M, N), dtype=torch.int32, device=dev)
        opg.gemm_forward_cuda(a, qw, z2, s2, ws, sa, o2, _acc_out=c2)
        o3 = torch.full((M, N), float("nan"), dtype=torch.half, device=dev)
        op.gemm_forward_cuda(a, qw, ws, sa, wz, asum, o3)  # the instantiation without the accumulator output
        torch.cuda.synchronize()
        res[nt] = (o1, c1, o2, c2, o3)
    lib.qs_gemm_force_tile_tokens(0)
    same = [bool(torch.equal(x.view(torch.int16) if x.dtype == torch.half else x, y.view(torch.int16) if y.dtype == torch.half else y))
            for x, y in zip(res[0], res[128])]
    print(f"M={M} N={N} K={K}: pair == NT128  per-chn fp16 {same[0]} acc {same[1]} | g128 fp16 {same[2]} acc {same[3]} | no-acc fp16 {same[4]}", flush=True)
