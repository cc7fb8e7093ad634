#!/usr/bin/env python
"""First check of the tcgen05 prefill attention against a float32 torch reference (and flash-attn when importable), then timing."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qserve_b200 import backend  # noqa: E402

dev = torch.device("cuda:0")


def reference(q, k, v, cu):
    out = torch.empty_like(q, dtype=torch.float32)
    g = q.size(1) // k.size(1)
    for b in range(len(cu) - 1):
        s, e = cu[b], cu[b + 1]
        qq = q[s:e].float().transpose(0, 1)                      # [Hq, L, D]
        kk = k[s:e].float().repeat_interleave(g, dim=1).transpose(0, 1)
        vv = v[s:e].float().repeat_interleave(g, dim=1).transpose(0, 1)
        sc = qq @ kk.transpose(1, 2) / math.sqrt(128)
        L = e - s
        sc = sc.masked_fill(torch.triu(torch.ones(L, L, dtype=torch.bool, device=q.device), 1), float("-inf"))
        out[s:e] = (torch.softmax(sc, dim=-1) @ vv).transpose(0, 1)
    return out


try:
    from flash_attn import flash_attn_varlen_func as fa
except Exception as e:  # noqa: BLE001
    fa = None
    print("flash_attn not importable:", e)

for (lens, hq, hkv) in [([128], 1, 1), ([256], 2, 1), ([100], 4, 2), ([1024, 333, 1, 129], 32, 8), ([2048, 700], 8, 8)]:
    T = sum(lens)
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    g = torch.Generator(device="cpu").manual_seed(T + hq)
    qkv = (torch.randn(T, (hq + 2 * hkv) * 128, generator=g) * 1.5).half().to(dev)
    q, k, v = qkv.split([hq * 128, hkv * 128, hkv * 128], dim=-1)
    q, k, v = q.reshape(T, hq, 128), k.reshape(T, hkv, 128), v.reshape(T, hkv, 128)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=dev)
    out = backend.flash_attn_varlen_func(q, k, v, cu_t, cu_t, max(lens), max(lens), dropout_p=0.0, causal=True)
    torch.cuda.synchronize()
    ref = reference(q, k, v, cu)
    err = (out.float() - ref).abs().max().item()
    msg = f"lens={lens} Hq={hq} Hkv={hkv}: max |ours - fp32| = {err:.3e}  finite={bool(torch.isfinite(out).all())}"
    if fa is not None:
        o2 = fa(q, k, v, cu_t, cu_t, max(lens), max(lens), dropout_p=0.0, causal=True)
        msg += f"  max |flash_attn - fp32| = {(o2.float() - ref).abs().max().item():.3e}  max |ours - flash_attn| = {(out.float() - o2.float()).abs().max().item():.3e}"
    print(msg, flush=True)

# timing: Llama-3-8B heads, 8 prompts of 1024 / 2 of 4096
for (lens, hq, hkv) in [([1024] * 8, 32, 8), ([4096] * 2, 32, 8)]:
    T = sum(lens)
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    qkv = torch.randn(T, (hq + 2 * hkv) * 128, device=dev).half()
    q, k, v = qkv.split([hq * 128, hkv * 128, hkv * 128], dim=-1)
    q, k, v = q.reshape(T, hq, 128), k.reshape(T, hkv, 128), v.reshape(T, hkv, 128)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=dev)
    flops = sum(4 * 128 * hq * n * (n + 1) / 2 for n in lens)
    for name, fn in (("ours", backend.flash_attn_varlen_func), ("flash_attn", fa)):
        if fn is None:
            continue
        for _ in range(3):
            fn(q, k, v, cu_t, cu_t, max(lens), max(lens), dropout_p=0.0, causal=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn(q, k, v, cu_t, cu_t, max(lens), max(lens), dropout_p=0.0, causal=True)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 10 * 1e-3
        print(f"{name:10s} lens={lens[0]}x{len(lens)}: {t * 1e6:8.1f} us  {flops / t / 1e12:6.1f} TFLOP/s (causal flops)", flush=True)
