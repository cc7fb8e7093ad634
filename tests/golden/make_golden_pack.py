#!/usr/bin/env python
"""Generate golden weight-packing vectors by importing the REFERENCE's own Python packer.

Runs in the build container only (needs /root/reference); the produced
tests/golden/pack_*.npz files are committed and travel to the GPU box.

The reference module cannot be imported as-is without a GPU
(`device=torch.cuda.current_device()` is evaluated at class-definition time,
w4a8_linear.py:19, and the per-channel branch calls `.cuda()`, :281), and it imports
the compiled `qserve_backend` extensions (:7-8).  We stub exactly those three things and run
`W4A8OF16LinearDynamicInputScale.from_linear` (w4a8_linear.py:136-332) unmodified on CPU.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/qserve/modeling/layers/quantized_linear/w4a8_linear.py"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference_packer():
    for name in ("qserve_backend", "qserve_backend.qgemm_w4a8_per_chn", "qserve_backend.qgemm_w4a8_per_group"):
        sys.modules.setdefault(name, types.ModuleType(name))
    torch.cuda.current_device = lambda: "cpu"  # evaluated as a default argument only
    torch.Tensor.cuda = lambda self, *a, **k: self
    spec = importlib.util.spec_from_file_location("ref_w4a8_linear", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.W4A8OF16LinearDynamicInputScale


def make(seed, N, K, group):
    rng = np.random.default_rng(seed)
    cls = load_reference_packer()
    q = rng.integers(0, 16, size=(N, K)).astype(np.int64)
    s1 = rng.uniform(0.005, 0.02, size=(N,)).astype(np.float16)
    lin = torch.nn.Linear(K, N, bias=False)
    if group == -1:
        z = rng.integers(0, 16, size=(N,)).astype(np.int64)
        w = (q - z[:, None]).astype(np.float32) * s1.astype(np.float32)[:, None]
        w_in = w.copy()  # from_linear quantises linear.weight.data in place
        lin.weight.data = torch.from_numpy(w)
        ql = cls.from_linear(lin, 4, -1, s1_scale=torch.from_numpy(s1), zeros=torch.from_numpy(z).to(torch.int8))
        out = dict(q=q.astype(np.uint8), s1=s1, z=z, w=w_in, qweight=ql.qweight.numpy(), s1_scales=ql.s1_scales.numpy(),
                   s1_szeros=ql.s1_szeros.numpy())
    else:
        G = K // group
        s2 = rng.integers(1, 9, size=(N, G)).astype(np.int64)  # |(q-z)*s2| <= 120 keeps stage-1 int8 in range
        z = rng.integers(0, 16, size=(N, G)).astype(np.int64)
        w8 = (q.reshape(N, G, group) - z[:, :, None]) * s2[:, :, None]
        w = w8.reshape(N, K).astype(np.float32) * s1.astype(np.float32)[:, None]
        w_in = w.copy()  # from_linear quantises linear.weight.data in place
        lin.weight.data = torch.from_numpy(w)
        ql = cls.from_linear(lin, 4, group, s1_scale=torch.from_numpy(s1),
                             s2_scale=torch.from_numpy(s2.astype(np.float16)),
                             zeros=torch.from_numpy(z).to(torch.int8))
        out = dict(q=q.astype(np.uint8), s1=s1, s2=s2, z=z, w=w_in, qweight=ql.qweight.numpy(),
                   s1_scales=ql.s1_scales.numpy(), s2_scales=ql.s2_scales.numpy(), s2_zeros=ql.s2_zeros.numpy())
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(OUT, "pack_per_chn_64x96.npz"), **make(1, 64, 96, -1))
    np.savez_compressed(os.path.join(OUT, "pack_per_chn_128x256.npz"), **make(2, 128, 256, -1))
    np.savez_compressed(os.path.join(OUT, "pack_per_group_64x256.npz"), **make(3, 64, 256, 128))
    np.savez_compressed(os.path.join(OUT, "pack_per_group_96x384.npz"), **make(4, 96, 384, 128))
    print("wrote golden pack vectors to", OUT)
