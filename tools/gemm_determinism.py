#!/usr/bin/env python
"""Run-to-run determinism of the GEMMs under load (two MMA issuers share one accumulator; any read-modify-write hazard in the
tensor pipe would show up as a differing INT32 accumulator).  Every shape is launched many times back to back with other
GEMMs in between; all accumulator tensors must be identical to the first."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qserve_backend.qgemm_w4a8_per_chn as op  # noqa: E402
import qserve_backend.qgemm_w8a8 as op8  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
bad = 0
for (M, N, K) in ((64, 28672, 4096), (64, 4096, 14336), (64, 6144, 4096), (16, 4096, 4096), (128, 4096, 4096), (300, 2048, 1024)):
    a = torch.randint(-127, 128, (M, K), dtype=torch.int8, generator=g).to(dev)
    w = torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, generator=g).to(dev)
    w8 = torch.randint(-128, 128, (N, K), dtype=torch.int8, generator=g).to(dev)
    s1 = torch.full((N,), 0.01, dtype=torch.half, device=dev)
    sa = torch.full((M,), 0.01, dtype=torch.half, device=dev)
    out = torch.empty((M, N), dtype=torch.half, device=dev)
    ref4 = ref8 = None
    for it in range(100):
        acc4 = torch.zeros((M, N), dtype=torch.int32, device=dev)
        acc8 = torch.zeros((M, N), dtype=torch.int32, device=dev)
        op.gemm_forward_cuda(a, w, s1, sa, s1, sa, out, _acc_out=acc4)
        op8.w8a8_gemm_forward_cuda(a, w8, s1, sa, out, _acc_out=acc8)
        op.gemm_forward_cuda(a, w, s1, sa, s1, sa, out)  # unrelated traffic in between
        if ref4 is None:
            ref4, ref8 = acc4.clone(), acc8.clone()
        else:
            bad += int(not torch.equal(acc4, ref4)) + int(not torch.equal(acc8, ref8))
    torch.cuda.synchronize()
    print(f"M={M} N={N} K={K}: 100 repetitions, mismatches so far {bad}")
print("DETERMINISTIC" if bad == 0 else f"NON-DETERMINISTIC: {bad} mismatching launches")
sys.exit(0 if bad == 0 else 1)
