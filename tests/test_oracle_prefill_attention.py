"""CPU: pins oracle/prefill_attention.py (a) against torch's float64 scaled_dot_product_attention and (b) against the committed output of
flash-attn itself (tests/golden/prefill_attn_flash.npz, generated on a B200 by tests/golden/make_golden_prefill_attn.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import prefill_attention as o

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "prefill_attn_flash.npz")


def test_oracle_matches_torch_sdpa(rng):
    lens, hq, hkv = [7, 1, 33], 4, 2
    T = sum(lens)
    q = rng.standard_normal((T, hq, 128)).astype(np.float16)
    k = rng.standard_normal((T, hkv, 128)).astype(np.float16)
    v = rng.standard_normal((T, hkv, 128)).astype(np.float16)
    cu = np.concatenate([[0], np.cumsum(lens)])
    out = o.causal_varlen_attention(q, k, v, cu)
    for b in range(len(lens)):
        s = slice(cu[b], cu[b + 1])
        ref = torch.nn.functional.scaled_dot_product_attention(
            torch.from_numpy(q[s]).double().transpose(0, 1), torch.from_numpy(k[s]).double().repeat_interleave(2, 1).transpose(0, 1),
            torch.from_numpy(v[s]).double().repeat_interleave(2, 1).transpose(0, 1), is_causal=True).transpose(0, 1).numpy()
        assert np.abs(out[s] - ref).max() < 1e-12


def test_oracle_matches_flash_attn_golden():
    if not os.path.exists(GOLDEN):
        pytest.skip("golden fixture not generated yet")
    z = np.load(GOLDEN)
    exact = o.causal_varlen_attention(z["q"], z["k"], z["v"], z["cu_seqlens"])
    # flash-attn rounds P to fp16 and the output to fp16: 2 fp16 ulps of the value + 1.5e-3 * max|v| (the bound the GPU tests use for our kernel)
    a = np.abs(exact)
    ulp = 2.0 ** (np.floor(np.log2(np.maximum(a, 2.0 ** -14))) - 10)
    assert (np.abs(z["out"].astype(np.float64) - exact) <= 2 * ulp + 1.5e-3 * float(np.abs(z["v"]).max())).all()
