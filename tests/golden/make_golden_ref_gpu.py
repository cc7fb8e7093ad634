#!/usr/bin/env python
"""Capture golden vectors from the UNMODIFIED reference CUDA kernels (oracle/_ref, built by oracle/build_ref.py for
sm_100a) on a B200:  gpurun -- python tests/golden/make_golden_ref_gpu.py
Writes gpurun_out/golden_ref/*.npz; the files are then committed under tests/golden/ref_*.npz and pin the CPU oracle
(tests/test_oracle_vs_reference_golden.py).  Inputs are seeded; every output of the reference kernels is stored."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import kv, ops, w4a8  # noqa: E402
from tests import refmods  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden_ref")
os.makedirs(OUT, exist_ok=True)
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
n = lambda x: x.detach().cpu().numpy()


def gemms():
    rng = np.random.default_rng(11)
    M, N, K = 48, 256, 512
    x = rng.standard_normal((M, K)).astype(np.float16)
    aq, sa, asum = ops.quant_per_token(x)
    q, qw, s1, s1z = w4a8.synth_per_channel(rng, N, K)
    out = torch.zeros((M, N), dtype=torch.half, device=dev)
    refmods.load("qgemm_w4a8_per_chn").gemm_forward_cuda(t(aq), t(qw), t(s1), t(sa), t(s1z), t(asum), out)
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(OUT, "ref_gemm_per_chn.npz"), aq=aq, qw=qw, s1=s1, sa=sa, s1z=s1z, asum=asum, out=n(out))
    q, qw, s1, s2s, s2z = w4a8.synth_per_group(rng, N, K)
    out = torch.zeros((M, N), dtype=torch.half, device=dev)
    refmods.load("qgemm_w4a8_per_group").gemm_forward_cuda(t(aq), t(qw), t(s2z), t(s2s), t(s1), t(sa), out)
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(OUT, "ref_gemm_per_group.npz"), aq=aq, qw=qw, s1=s1, sa=sa, s2s=s2s, s2z=s2z, out=n(out))
    w = rng.integers(-128, 128, size=(N, K), dtype=np.int8)
    sw = rng.uniform(0.001, 0.01, size=N).astype(np.float16)
    out = torch.zeros((M, N), dtype=torch.half, device=dev)
    refmods.load("qgemm_w8a8").w8a8_gemm_forward_cuda(t(aq), t(w), t(sw), t(sa), out)
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(OUT, "ref_gemm_w8a8.npz"), aq=aq, w=w, sw=sw, sa=sa, out=n(out))


def elementwise():
    rng = np.random.default_rng(12)
    M, H = 12, 1024
    x = (rng.standard_normal((M, H)) * 2 + 0.3).astype(np.float16)
    g = (1 + 0.1 * rng.standard_normal(H)).astype(np.float16)
    fk, ln, act = refmods.load("fused_kernels"), refmods.load("layernorm_ops"), refmods.load("activation_ops")
    q = torch.empty((M, H), dtype=torch.int8, device=dev); s = torch.empty(M, dtype=torch.half, device=dev); sm = torch.empty(M, dtype=torch.half, device=dev)
    fk.invoke_quant_fuse_sum(q, t(x), sm, s)
    r = dict(x=x, gamma=g, quant_q=n(q), quant_scale=n(s), quant_sum=n(sm))
    ln.rms_norm_general_fuse_sum(q, t(x), t(g), sm, s, 1e-5, True)
    r.update(ln_q=n(q), ln_scale=n(s), ln_sum=n(sm))
    o = torch.empty((M, H), dtype=torch.half, device=dev)
    ln.rms_norm(o, t(x), t(g), 1e-5, False)
    r.update(rms=n(o))
    o2 = torch.empty((M, H // 2), dtype=torch.half, device=dev)
    act.silu_and_mul(o2, t(x))
    r.update(silu=n(o2))
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(OUT, "ref_elementwise.npz"), **r)


def attention():
    fa = refmods.load("fused_attention")
    for bits in (4, 8):
        rng = np.random.default_rng(13 + bits)
        B, Hq, Hkv, D = 3, 8, 2, 128
        lens = [1, 65, 200]
        kp, vp = kv.PagePool(13, Hkv, D, bits, rng), kv.PagePool(13, Hkv, D, bits, rng)
        bt = (1 + np.arange(B * 4).reshape(B, 4)) % 13
        q, k, v = (rng.standard_normal(s).astype(np.float16) for s in ((B, Hq, D), (B, Hkv, D), (B, Hkv, D)))
        kd, vd = t(kp.data), t(vp.data)
        table = torch.from_numpy(np.stack([kd.data_ptr() + bt * kp.pb, vd.data_ptr() + bt * vp.pb], axis=1)).to(dev)
        # the reference takes the row stride of q, k AND v from q.stride(0) (fused_attention.cpp:215): they must be views of
        # one packed qkv buffer, exactly as llama_w4a8_unpad.py:245-252 passes them
        qkv_d = torch.from_numpy(np.concatenate([q.reshape(B, -1), k.reshape(B, -1), v.reshape(B, -1)], axis=1)).to(dev)
        qd, kd_, vd_ = qkv_d.split([Hq * D, Hkv * D, Hkv * D], dim=-1)
        out = fa.single_query_attention(qd.reshape(B, Hq, D), kd_.reshape(B, Hkv, D), vd_.reshape(B, Hkv, D), table,
                                        torch.tensor(lens, dtype=torch.int32, device=dev), None, 8192, 64,
                                        Hkv * D * bits // 8, max(lens), D, 10000.0, True, bits == 4, True)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(OUT, f"ref_decode_attn_kv{bits}.npz"), q=q, k=k, v=v, kpool=kp.data, vpool=vp.data, bt=bt, lens=np.array(lens),
                            out=n(out), kpool_after=n(kd), vpool_after=n(vd))
        # prefill append
        lens_p = np.array([5, 70, 130], np.int32)
        T, maxlen = int(lens_p.sum()), int(lens_p.max())
        cu = np.concatenate([[0], np.cumsum(lens_p)]).astype(np.int32)
        pad = fa.compute_padding_offsets(t(cu), maxlen, T)
        qkv = rng.standard_normal((T, (Hq + 2 * Hkv) * D)).astype(np.float16)
        kp2, vp2 = kv.PagePool(9, Hkv, D, bits), kv.PagePool(9, Hkv, D, bits)
        bt2 = np.arange(9).reshape(3, 3)
        kd2, vd2 = t(kp2.data), t(vp2.data)
        table2 = torch.from_numpy(np.stack([kd2.data_ptr() + bt2 * kp2.pb, vd2.data_ptr() + bt2 * vp2.pb], axis=1)).to(dev)
        qkv_d = t(qkv)
        fa.apply_bias_rope_update_kv_cache(qkv_d, t(lens_p), pad, table2, Hq, Hkv, maxlen, 64, Hkv * D * bits // 8, D, 10000.0, 8192, True, bits == 4, True)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(OUT, f"ref_prefill_kv{bits}.npz"), qkv=qkv, lens=lens_p, pad=n(pad), bt=bt2, qkv_after=n(qkv_d),
                            kpool_after=n(kd2), vpool_after=n(vd2))


if __name__ == "__main__":
    for fn in (gemms, elementwise, attention):
        try:
            fn()
            print("captured", fn.__name__)
        except Exception as e:  # keep going: each family is independent
            print("FAILED", fn.__name__, repr(e))
