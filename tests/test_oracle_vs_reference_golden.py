"""CPU: the oracle against golden vectors captured from the UNMODIFIED reference CUDA kernels on a B200
(oracle/build_ref.py -> oracle/_ref, tests/golden/make_golden_ref_gpu.py).  This is what pins the oracle.

The reference binaries are built with --use_fast_math (kernels/setup.py:33): approximate division / sin / cos and FMA
contraction move a result by at most one rounding step, so the stated bars are
  GEMM fp16 outputs          <= 1 fp16 ulp, on < 0.1 % of the elements (per-group / W8A8: 0 differences observed)
  per-token INT8 codes       <= 1 LSB, on < 0.05 % of the elements; scales and sums bit-exact
  norm (N1)                  codes, scale and the fp16-accumulated row sum bit-exact
  rms_norm, silu_and_mul     bit-exact
  prefill RoPE               rotated q/k within 4e-3 absolute; KV pages: codes within 1 LSB on < 0.1 % of the bytes
  decode attention           output within 1e-2 * max|out| (fp16 logits / tree reduction are order dependent, 8c)
"""
import os

import numpy as np
import pytest

from oracle import kv, ops, w4a8
from tests.util import bits16, ulp16_diff

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not captured yet")
    return np.load(path)


def test_gemm_per_channel():
    g = _load("ref_gemm_per_chn.npz")
    d = ulp16_diff(w4a8.gemm_w4a8_per_chn(g["aq"], g["qw"], g["s1"], g["sa"], g["s1z"], g["asum"]), g["out"])
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_gemm_per_group():
    g = _load("ref_gemm_per_group.npz")
    d = ulp16_diff(w4a8.gemm_w4a8_per_group(g["aq"], g["qw"], g["s2z"], g["s2s"], g["s1"], g["sa"]), g["out"])
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_gemm_w8a8():
    g = _load("ref_gemm_w8a8.npz")
    d = ulp16_diff(w4a8.gemm_w8a8(g["aq"], g["w"], g["sw"], g["sa"]), g["out"])
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_elementwise():
    g = _load("ref_elementwise.npz")
    q, s, sm = ops.quant_per_token(g["x"])
    dq = np.abs(q.astype(np.int32) - g["quant_q"].astype(np.int32))
    assert dq.max() <= 1 and (dq > 0).mean() < 5e-4
    assert np.array_equal(bits16(s), bits16(g["quant_scale"])) and np.array_equal(bits16(sm), bits16(g["quant_sum"]))
    q, s, sm, _ = ops.layernorm_general_quant(g["x"], g["gamma"], 1e-5)
    assert np.array_equal(q, g["ln_q"]) and np.array_equal(bits16(s), bits16(g["ln_scale"])) and np.array_equal(bits16(sm), bits16(g["ln_sum"]))
    assert np.array_equal(bits16(ops.rms_norm(g["x"], g["gamma"], 1e-5)), bits16(g["rms"]))
    assert np.array_equal(bits16(ops.silu_and_mul(g["x"])), bits16(g["silu"]))


@pytest.mark.parametrize("bits", [4, 8])
def test_prefill_rope_append(bits):
    g = _load(f"ref_prefill_kv{bits}.npz")
    Hq, Hkv, D = 8, 2, 128
    kp = kv.PagePool(g["kpool_after"].shape[0], Hkv, D, bits)
    vp = kv.PagePool(g["vpool_after"].shape[0], Hkv, D, bits)
    qkv = kv.prefill_rope_append(g["qkv"].copy(), g["lens"], g["pad"], kp, vp, g["bt"], Hq, Hkv, int(g["lens"].max()), 10000.0, 8192)
    assert np.abs(qkv.astype(np.float32) - g["qkv_after"].astype(np.float32)).max() <= 4e-3
    assert np.array_equal(kv.compute_padding_offsets(np.concatenate([[0], np.cumsum(g["lens"])]), int(g["lens"].max()), int(g["lens"].sum())), g["pad"])
    for mine, ref in ((kp, g["kpool_after"]), (vp, g["vpool_after"])):
        assert (mine.data != ref).mean() < 1e-3
        d = np.abs(mine.data[:, : mine.cb].astype(np.int32) - ref[:, : mine.cb].astype(np.int32))
        if bits == 8:
            assert d.max() <= 1


@pytest.mark.parametrize("bits", [4, 8])
def test_decode_attention(bits):
    g = _load(f"ref_decode_attn_kv{bits}.npz")
    B, Hq, D = g["q"].shape
    Hkv = g["k"].shape[1]
    kp = kv.PagePool(g["kpool"].shape[0], Hkv, D, bits); kp.data[:] = g["kpool"]
    vp = kv.PagePool(g["vpool"].shape[0], Hkv, D, bits); vp.data[:] = g["vpool"]
    o = kv.decode_attention(g["q"], g["k"], g["v"], kp, vp, g["bt"], g["lens"], 10000.0, faithful=True).astype(np.float32)
    ref = g["out"].astype(np.float32)
    assert np.isfinite(ref).all()
    assert np.abs(o - ref).max() <= 1e-2 * max(1.0, np.abs(ref).max())
    # appended K / V slots: V is pure IEEE -> (near) bit-exact; K goes through fast-math RoPE in the reference
    assert (vp.data != g["vpool_after"]).mean() < 1e-4
    assert (kp.data != g["kpool_after"]).mean() < 1e-3
