#!/usr/bin/env python
"""Per-CTA phase timeline of the decode attention kernel from %globaltimer stamps (qs_gemm_set_profile_buffer)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qserve_b200 import backend  # noqa: E402
from qserve_b200._lib import lib  # noqa: E402
from qserve_b200.decode import DecodeRunner  # noqa: E402
import qserve_backend.fused_attention as fa  # noqa: E402

NAMES = ["entry", "dep_resolved", "page_table_staged", "first_page_issued", "consumer_prologue_done", "first_page_landed", "mainloop_done",
         "partials_synced", "merged", "exit"]
precision = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "w4a8kv4"
B, ctx = 64, 1024
run = DecodeRunner("llama-3-8b", precision, B, ctx, torch.device("cuda:0"), layers=3)
D = run.cfg.head_dim
run.qkv_buf.normal_()
q, k, v = run.qkv_buf.split([run.q_size, run.kv_size, run.kv_size], dim=-1)
q, k, v = q.reshape(B, run.Hq, D), k.reshape(B, run.Hkv, D), v.reshape(B, run.Hkv, D)


def call(i):
    return fa.single_query_attention(q, k, v, run.block_tables[i % 3], run.context_lens, None, 8192, 64, run.size_per_token, run.max_seq_len, D,
                                     run.cfg.rope_theta, True, run.kv_bits == 4, True)


for i in range(3):
    call(i)
torch.cuda.synchronize()
prof = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
lib.qs_gemm_set_profile_buffer(prof.data_ptr())
call(0)
torch.cuda.synchronize()
lib.qs_gemm_set_profile_buffer(None)
p = prof.cpu().numpy().reshape(-1, 16)
p = p[p[:, 0] > 0].astype(np.float64)
t0 = p[:, 0].min()
print(f"== decode attention {precision} B={B} ctx={ctx}: {len(p)} CTAs, span {(p[:, 9].max() - t0) / 1e3:.2f} us")
for j, nm in enumerate(NAMES):
    col = p[:, j]
    ok = col > 0
    if ok.any():
        rel = (col[ok] - t0) / 1e3
        print(f"   {nm:24s} n={ok.sum():4d}  min {rel.min():7.2f}  med {np.median(rel):7.2f}  max {rel.max():7.2f} us")
d = p[:, 6] - p[:, 5]
print(f"   mainloop duration per CTA: med {np.median(d) / 1e3:.2f} us, max {d.max() / 1e3:.2f} us")
