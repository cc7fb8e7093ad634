#!/usr/bin/env python
"""Launch the prefill-sized per-channel GEMM (M tokens through gate_up_proj of Llama-3-8B) a few times -- target for the
tensor-pipe utilisation capture (`ncu --set full -k regex:gemm_kernel`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qserve_backend.qgemm_w4a8_per_chn as op  # noqa: E402
import qserve_backend.qgemm_w4a8_per_group as opg  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
if os.environ.get("QS_FORCE_NT"):  # tile-size A/B: tokens per tile (32 / 64 / 128 / 256)
    from qserve_b200._lib import lib
    lib.qs_gemm_force_tile_tokens(int(os.environ["QS_FORCE_NT"]))
N, K = int(os.environ.get("QS_N", 28672)), int(os.environ.get("QS_K", 4096))
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
a = torch.randint(-127, 128, (M, K), dtype=torch.int8, generator=g).to(dev)
w = torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, generator=g).to(dev)
s1 = torch.full((N,), 0.01, dtype=torch.half, device=dev)
sa = torch.full((M,), 0.01, dtype=torch.half, device=dev)
out = torch.empty((M, N), dtype=torch.half, device=dev)
if os.environ.get("QS_G128"):  # the per-group (g128) kernel on the same shape
    z2 = torch.randint(-100, 0, (K // 128, N), dtype=torch.int8, generator=g).to(dev)
    s2 = torch.randint(1, 9, (K // 128, N), dtype=torch.int8, generator=g).to(dev)
    _chn = op.gemm_forward_cuda

    class op:  # noqa: N801
        @staticmethod
        def gemm_forward_cuda(a, w, s1, sa, _z, _s, out):
            opg.gemm_forward_cuda(a, w, z2, s2, s1, sa, out)
for _ in range(4):
    op.gemm_forward_cuda(a, w, s1, sa, s1, sa, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    op.gemm_forward_cuda(a, w, s1, sa, s1, sa, out)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10 * 1e-3
print(f"M={M} N={N} K={K}: {t * 1e6:.1f} us, {2.0 * M * N * K / t / 1e12:.0f} INT8 TOP/s")
