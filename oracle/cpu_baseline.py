"""CPU baseline leg (TEST / MEASUREMENT INFRASTRUCTURE, not product).

The reference ships no CPU implementation; BASELINE.json's north_star names "the reference's torch-CPU
dequant-then-FP16-matmul path" as the thing to time next to the GPU kernels.  This module is that restatement
(BASELINE.md section 2a): unpack the checkpoint-layout INT4 weights (inverse of w4a8_linear.py:292-322), de-quantise
W = (q - z) * s1 and X = a_q * sa, matmul with fp32 accumulation -> fp16; attention = de-quantise the KV4 pages and
run fp32 scaled-dot-product attention with GQA expansion.  torch CPU with all host threads.

One "sample" = ONE decoder layer of the decode step (4 GEMMs at M = batch + one attention layer) plus, once, the fp16
lm_head; tokens/s is extrapolated to the full depth:  batch / (layers * t_layer + t_lm_head).
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import kv, w4a8


def _unpack_dequant(qweight: np.ndarray, s1: np.ndarray, s1z: np.ndarray) -> torch.Tensor:
    q = torch.from_numpy(w4a8.unpack_w4(qweight))
    return q.to(torch.float32) * torch.from_numpy(s1.astype(np.float32))[:, None] - torch.from_numpy(s1z.astype(np.float32))[:, None]


class LayerSample:
    """Inputs of one decoder layer's sample, generated ONCE (outside every timed region); `run()` does the timed CPU work.
    Used by `bench.py --impl reference`, whose steps are bounded samples of the workload (one layer each)."""

    def __init__(self, hidden: int, intermediate: int, heads: int, kv_heads: int, batch: int, ctx: int, threads: int, seed: int = 0):
        torch.set_num_threads(threads)
        rng = np.random.default_rng(seed)
        D = 128
        self.batch, self.ctx, self.heads, self.kv_heads, self.D = batch, ctx, heads, kv_heads, D
        self.gemms = []
        for K, N in [(hidden, (heads + 2 * kv_heads) * D), (heads * D, hidden), (hidden, 2 * intermediate), (intermediate, hidden)]:
            qw = rng.integers(-128, 128, size=(N, K // 2), dtype=np.int8)
            s1 = rng.uniform(0.005, 0.02, size=N).astype(np.float16)
            s1z = (s1.astype(np.float32) * 8).astype(np.float16)
            aq = torch.from_numpy(rng.integers(-127, 128, size=(batch, K), dtype=np.int8))
            sa = torch.from_numpy(rng.uniform(0.01, 0.05, size=batch).astype(np.float32))
            self.gemms.append((qw, s1, s1z, aq, sa))
        self.codes_k = torch.from_numpy(rng.integers(0, 256, size=(batch, kv_heads, ctx, D // 2), dtype=np.uint8))
        self.codes_v = torch.from_numpy(rng.integers(0, 256, size=(batch, kv_heads, ctx, D // 2), dtype=np.uint8))
        self.sk = torch.from_numpy(rng.uniform(0.01, 0.1, size=(batch, kv_heads, ctx, 1)).astype(np.float32))
        self.zk = torch.from_numpy(rng.uniform(0, 15, size=(batch, kv_heads, ctx, 1)).astype(np.float32))
        self.q = torch.from_numpy(rng.standard_normal((batch, heads, 1, D)).astype(np.float32))

    def run(self):
        t_deq = t_mm = 0.0
        for qw, s1, s1z, aq, sa in self.gemms:
            t0 = time.perf_counter()
            w = _unpack_dequant(qw, s1, s1z)
            x = aq.to(torch.float32) * sa[:, None]
            t1 = time.perf_counter()
            y = (x @ w.T).to(torch.float16)
            t2 = time.perf_counter()
            t_deq += t1 - t0
            t_mm += t2 - t1
            del w, y
        batch, kv_heads, ctx, D = self.batch, self.kv_heads, self.ctx, self.D
        t0 = time.perf_counter()

        def deq(c):
            lo = (c & 0xF).to(torch.float32)
            hi = (c >> 4).to(torch.float32)
            u = torch.stack([lo, hi], dim=-1).reshape(batch, kv_heads, ctx, D)
            return (u - self.zk) * self.sk

        kf, vf = deq(self.codes_k), deq(self.codes_v)
        g = self.heads // kv_heads
        o = torch.nn.functional.scaled_dot_product_attention(self.q, kf.repeat_interleave(g, dim=1), vf.repeat_interleave(g, dim=1))
        t_attn = time.perf_counter() - t0
        del o, kf, vf
        return {"dequant_s": t_deq, "matmul_s": t_mm, "attention_s": t_attn, "layer_s": t_deq + t_mm + t_attn}


def layer_sample(hidden: int, intermediate: int, heads: int, kv_heads: int, batch: int, ctx: int, threads: int, seed: int = 0):
    """Time one decoder layer's worth of the hot path on the CPU.  Returns dict of seconds."""
    torch.set_num_threads(threads)
    rng = np.random.default_rng(seed)
    D = 128
    shapes = [(hidden, (heads + 2 * kv_heads) * D), (heads * D, hidden), (hidden, 2 * intermediate), (intermediate, hidden)]
    t_deq = t_mm = 0.0
    for K, N in shapes:
        qw = rng.integers(-128, 128, size=(N, K // 2), dtype=np.int8)
        s1 = rng.uniform(0.005, 0.02, size=N).astype(np.float16)
        s1z = (s1.astype(np.float32) * 8).astype(np.float16)
        aq = torch.from_numpy(rng.integers(-127, 128, size=(batch, K), dtype=np.int8))
        sa = torch.from_numpy(rng.uniform(0.01, 0.05, size=batch).astype(np.float32))
        t0 = time.perf_counter()
        w = _unpack_dequant(qw, s1, s1z)
        x = aq.to(torch.float32) * sa[:, None]
        t1 = time.perf_counter()
        y = (x @ w.T).to(torch.float16)
        t2 = time.perf_counter()
        t_deq += t1 - t0
        t_mm += t2 - t1
        del w, y
    # attention: de-quantise one layer's KV4 pages, fp32 SDPA with GQA expansion
    n_tok = batch * ctx
    codes_k = torch.from_numpy(rng.integers(0, 256, size=(batch, kv_heads, ctx, D // 2), dtype=np.uint8))
    codes_v = torch.from_numpy(rng.integers(0, 256, size=(batch, kv_heads, ctx, D // 2), dtype=np.uint8))
    sk = torch.from_numpy(rng.uniform(0.01, 0.1, size=(batch, kv_heads, ctx, 1)).astype(np.float32))
    zk = torch.from_numpy(rng.uniform(0, 15, size=(batch, kv_heads, ctx, 1)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((batch, heads, 1, D)).astype(np.float32))
    t0 = time.perf_counter()

    def deq(c):
        lo = (c & 0xF).to(torch.float32)
        hi = (c >> 4).to(torch.float32)
        u = torch.stack([lo, hi], dim=-1).reshape(batch, kv_heads, ctx, D)
        return (u - zk) * sk

    kf, vf = deq(codes_k), deq(codes_v)
    g = heads // kv_heads
    kf = kf.repeat_interleave(g, dim=1)
    vf = vf.repeat_interleave(g, dim=1)
    o = torch.nn.functional.scaled_dot_product_attention(q, kf, vf)
    t_attn = time.perf_counter() - t0
    del o, kf, vf
    return {"dequant_s": t_deq, "matmul_s": t_mm, "attention_s": t_attn, "layer_s": t_deq + t_mm + t_attn, "tokens_in_kv": n_tok}


def lm_head_sample(hidden: int, vocab: int, batch: int, threads: int) -> float:
    torch.set_num_threads(threads)
    w = torch.randn((vocab, hidden), dtype=torch.float16)
    x = torch.randn((batch, hidden), dtype=torch.float16)
    t0 = time.perf_counter()
    y = torch.nn.functional.linear(x.float(), w.float())
    t = time.perf_counter() - t0
    del y
    return t


def decode_tokens_per_s(hidden, intermediate, heads, kv_heads, layers, vocab, batch, ctx, threads=None, repeats: int = 1):
    threads = threads or os.cpu_count() or 1
    best = None
    for r in range(repeats):
        s = layer_sample(hidden, intermediate, heads, kv_heads, batch, ctx, threads, seed=r)
        if best is None or s["layer_s"] < best["layer_s"]:
            best = s
    t_lm = lm_head_sample(hidden, vocab, batch, threads)
    step_s = layers * best["layer_s"] + t_lm
    return {"tokens_per_s": batch / step_s, "step_s": step_s, "lm_head_s": t_lm, "threads": threads, **best}
