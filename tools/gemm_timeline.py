#!/usr/bin/env python
"""Per-CTA phase timeline of the tcgen05 GEMM from %globaltimer stamps (qs_gemm_set_profile_buffer)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qserve_b200 import backend  # noqa: E402
from qserve_b200._lib import lib  # noqa: E402
from qserve_b200.decode import DecodeRunner  # noqa: E402

NAMES = ["entry", "setup_done", "pdl_wait_done", "loads_issued", "first_full", "first_afull", "last_commit(seg)", "epi_start", "dfull_seen",
         "partial_stored", "flag_known", "epi_done(seg0)", "exit"]
pdl = "--no-pdl" not in sys.argv
backend.set_pdl(pdl)
run = DecodeRunner("llama-3-8b", "w4a8kv4", 64, 1024, torch.device("cuda:0"), layers=3)
run.q_scale.fill_(0.01); run.q_sum.fill_(0.1)
for name, xq, buf in (("qkv", run.q_hidden, run.qkv_buf), ("o", run.q_attn, run.out_buf), ("gate_up", run.q_hidden, run.gate_up_buf), ("down", run.q_mlp, run.out_buf)):
    prof = torch.zeros(1024 * 16, dtype=torch.int64, device="cuda")
    for i in range(3):  # warm
        run.layers[i % 3][name](xq, run.q_scale, run.q_sum, buf)
    torch.cuda.synchronize()
    lib.qs_gemm_set_profile_buffer(prof.data_ptr())
    run.layers[0][name](xq, run.q_scale, run.q_sum, buf)
    torch.cuda.synchronize()
    lib.qs_gemm_set_profile_buffer(None)
    p = prof.cpu().numpy().reshape(-1, 16)
    used = p[:, 0] > 0
    p = p[used].astype(np.float64)
    t0 = p[:, 0].min()
    print(f"== gemm_{name}: {used.sum()} CTAs, span {(p[:, 12].max() - t0) / 1e3:.2f} us (pdl={pdl})")
    for j, nm in enumerate(NAMES):
        col = p[:, j]
        ok = col > 0
        if ok.any():
            rel = (col[ok] - t0) / 1e3
            print(f"   {nm:18s} n={ok.sum():4d}  min {rel.min():7.2f}  med {np.median(rel):7.2f}  max {rel.max():7.2f} us")
