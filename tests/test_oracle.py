"""CPU tests: the oracle against the golden vectors and against itself (no GPU)."""
import glob
import os

import numpy as np
import pytest

from oracle import kv, ops, w4a8

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pack_*.npz"))))
def test_pack_matches_reference_packer(path):
    """oracle.w4a8.pack_* == W4A8OF16LinearDynamicInputScale.from_linear run from /root/reference (make_golden_pack.py)."""
    g = np.load(path)
    if "s2" in g:
        qw, s1, s2s, s2z = w4a8.pack_per_group(g["q"], g["s1"], g["s2"], g["z"])
        assert np.array_equal(s2s, g["s2_scales"]) and np.array_equal(s2z, g["s2_zeros"])
    else:
        qw, s1, sz = w4a8.pack_per_channel(g["q"], g["s1"], g["z"])
        assert np.array_equal(sz.view(np.uint16), g["s1_szeros"].view(np.uint16))
    assert np.array_equal(qw, g["qweight"])
    assert np.array_equal(s1.view(np.uint16), g["s1_scales"].view(np.uint16))
    assert np.array_equal(w4a8.unpack_w4(g["qweight"]), g["q"])


def test_pack_unpack_roundtrip_property(rng):
    for N, K in ((32, 32), (64, 96), (160, 256)):
        q = rng.integers(0, 16, size=(N, K), dtype=np.uint8)
        assert np.array_equal(w4a8.unpack_w4(w4a8.pack_w4(q)), q)
        x = rng.integers(-128, 128, size=(3, N)).astype(np.int32)
        assert np.array_equal(w4a8.unshuffle_n32(w4a8.shuffle_n32(x)), x)


def test_pack_layout_formula(rng):
    """lane t = c*4+e, byte j = d*8+b*4+f of a 32x32 tile."""
    q = rng.integers(0, 16, size=(64, 64), dtype=np.uint8)
    P = w4a8.pack_w4(q).view(np.uint8).reshape(2, 2, 32, 16)
    for _ in range(200):
        n32, k32, c, e, d, b, f = (rng.integers(0, m) for m in (2, 2, 8, 4, 2, 2, 4))
        byte = P[n32, k32, c * 4 + e, d * 8 + b * 4 + f]
        n, k = n32 * 32 + b * 8 + c, k32 * 32 + d * 16 + e * 4 + f
        assert byte & 0xF == q[n, k] and byte >> 4 == q[n + 16, k]


def test_level2_dequant_matches_plain_formula_in_protective_range(rng):
    q, qw, s1, s2s, s2z = w4a8.synth_per_group(rng, 64, 256)
    w8 = w4a8.dequant_level2(q, s2s, s2z)
    s2 = w4a8.unshuffle_n32(s2s.view(np.uint8).astype(np.int32))
    z2 = w4a8.unshuffle_n32(s2z.astype(np.int32))
    want = (q.astype(np.int32) * np.repeat(s2.T, 128, axis=1) + np.repeat(z2.T, 128, axis=1))
    assert np.array_equal(w8.astype(np.int32), ((want + 128) % 256) - 128)


def test_config1_plumbing_per_channel_vs_dequant_matmul(rng):
    """BASELINE config 1: M=16, K=4096, N=4096, integer-exact oracle vs torch-CPU style dequant-then-matmul."""
    M, K, N = 16, 4096, 4096
    w = rng.normal(0, 0.02, size=(N, K)).astype(np.float32)
    q, s1, z = w4a8.fake_quant_per_channel(w)
    qw, s1h, s1z = w4a8.pack_per_channel(q, s1, z)
    x = rng.standard_normal((M, K)).astype(np.float16)
    aq, sa, _ = ops.quant_per_token(x)
    asum = (sa.astype(np.float32) * aq.astype(np.int32).sum(axis=1)).astype(np.float16)  # asum = sa * sum(a_q)
    out, acc = w4a8.gemm_w4a8_per_chn(aq, qw, s1h, sa, s1z, asum, return_acc=True)
    ref = w4a8.dequant_then_matmul_per_chn(aq, qw, s1h, s1z, sa)
    assert np.array_equal(acc, w4a8.int_matmul(aq, q))
    err = np.abs(out.astype(np.float32) - ref.astype(np.float32))
    # fp16 output (|y| <~ 4): the two formulas differ only by fp16 rounding of asum and the final rounding
    assert err.max() <= 2e-2, err.max()
    assert np.median(err) <= 1e-3


def test_quant_roundtrip_and_saturation(rng):
    x = (rng.standard_normal((7, 512)) * 3).astype(np.float16)
    q, s, sm = ops.quant_per_token(x)
    assert q.dtype == np.int8 and np.abs(q).max() == 127
    back = q.astype(np.float32) * s.astype(np.float32)[:, None]
    assert np.abs(back - x.astype(np.float32)).max() <= s.astype(np.float32).max() * 0.51
    assert np.allclose(sm.astype(np.float32), x.astype(np.float32).sum(axis=1), rtol=2e-3, atol=2e-2)


def test_layernorm_is_mean_subtracting(rng):
    x = (rng.standard_normal((5, 256)) + 3.0).astype(np.float16)
    q, s, sm, y = ops.layernorm_general_quant(x, np.ones(256, np.float16), 1e-5)
    assert abs(y.mean()) < 1e-3 and abs(y.std() - 1) < 1e-2  # N1 quirk: LayerNorm without beta
    assert np.abs(sm.astype(np.float32)).max() < 0.5


def test_kv_quant_roundtrip(rng):
    for bits in (4, 8):
        x = rng.standard_normal((11, 3, 128)).astype(np.float16)
        s, z = kv.kv_quant_params(x, bits)
        u = kv.kv_quant_codes(x, s, z, bits)
        assert u.max() <= (15 if bits == 4 else 255)
        xd = kv.kv_dequant(u, s, z, bits)
        assert np.abs(xd.astype(np.float32) - x.astype(np.float32)).max() <= 0.51 * float(s.max()) + 2e-2
        assert np.array_equal(kv.unpack_nibbles(kv.pack_nibbles(u & 0xF)), u & 0xF)


def test_rope_is_a_rotation(rng):
    x = rng.standard_normal((6, 4, 128)).astype(np.float16)
    y = kv.rope_neox(x, np.arange(6)[:, None] * 37, 10000.0)
    n0 = np.linalg.norm(x.astype(np.float32), axis=-1)
    n1 = np.linalg.norm(y.astype(np.float32), axis=-1)
    assert np.allclose(n0, n1, rtol=2e-3)
    assert np.array_equal(kv.rope_neox(x, 0, 10000.0), x)  # position 0 is the identity


def test_decode_attention_faithful_close_to_exact(rng):
    B, Hq, Hkv, D = 3, 8, 2, 128
    for bits in (4, 8):
        kp, vp = kv.PagePool(16, Hkv, D, bits, rng), kv.PagePool(16, Hkv, D, bits, rng)
        bt = np.arange(B * 4).reshape(B, 4) % 16
        lens = np.array([1, 65, 200])
        q, k, v = (rng.standard_normal(s).astype(np.float16) for s in ((B, Hq, D), (B, Hkv, D), (B, Hkv, D)))
        kp2, vp2 = kv.PagePool(16, Hkv, D, bits), kv.PagePool(16, Hkv, D, bits)
        kp2.data[:], vp2.data[:] = kp.data, vp.data
        o1 = kv.decode_attention(q, k, v, kp, vp, bt, lens, 10000.0, faithful=True)
        o2 = kv.decode_attention(q, k, v, kp2, vp2, bt, lens, 10000.0, faithful=False)
        assert np.array_equal(kp.data, kp2.data) and np.array_equal(vp.data, vp2.data)
        scale = np.abs(o2.astype(np.float32)).max()
        assert np.abs(o1.astype(np.float32) - o2.astype(np.float32)).max() <= 4e-3 * max(scale, 1.0)
        # sequence of length 1: softmax over the single new token -> out == v (up to the 1e-6 in the normaliser)
        assert np.allclose(o1[0].reshape(Hkv, Hq // Hkv, D).astype(np.float32), v[0][:, None, :].astype(np.float32), atol=2e-3)


def test_prefill_then_decode_consistency(rng):
    """Pages written by the prefill path are exactly what the decode path would have written token by token."""
    Hq, Hkv, D, bits = 4, 2, 128, 4
    lens = np.array([5, 70], np.int32)
    T = int(lens.sum())
    cu = np.concatenate([[0], np.cumsum(lens)])
    pad = kv.compute_padding_offsets(cu, int(lens.max()), T)
    qkv = rng.standard_normal((T, (Hq + 2 * Hkv) * D)).astype(np.float16)
    bt = np.array([[0, 1], [2, 3]])
    kp, vp = kv.PagePool(4, Hkv, D, bits), kv.PagePool(4, Hkv, D, bits)
    orig = qkv.copy()
    kv.prefill_rope_append(qkv, lens, pad, kp, vp, bt, Hq, Hkv, int(lens.max()), 10000.0, 8192)
    # v untouched, q/k rotated
    assert np.array_equal(qkv[:, (Hq + Hkv) * D:], orig[:, (Hq + Hkv) * D:])
    # token 3 of sequence 1 via the decode path
    t = int(cu[1]) + 3
    kp2, vp2 = kv.PagePool(4, Hkv, D, bits), kv.PagePool(4, Hkv, D, bits)
    k_new = orig[t, Hq * D:(Hq + Hkv) * D].reshape(1, Hkv, D)
    v_new = orig[t, (Hq + Hkv) * D:].reshape(1, Hkv, D)
    q_new = orig[t, : Hq * D].reshape(1, Hq, D)
    kv.decode_attention(q_new, k_new, v_new, kp2, vp2, bt[1:2], np.array([4]), 10000.0)
    assert np.array_equal(kp2.codes()[2, :, 3], kp.codes()[2, :, 3])
    assert np.array_equal(vp2.codes()[2, :, 3], vp.codes()[2, :, 3])
    assert np.array_equal(kp2.scales()[2, :, 3].view(np.uint16), kp.scales()[2, :, 3].view(np.uint16))


def test_prefill_attention_reference_properties(rng):
    """The prefill-attention restatement (next scope row): agrees with torch SDPA per sequence, is causal, and its last row
    equals single-query attention over the un-quantised keys of the same sequence."""
    import torch
    Hq, Hkv, D = 4, 2, 128
    lens = [5, 1, 9]
    cu = np.concatenate([[0], np.cumsum(lens)])
    T = int(cu[-1])
    qkv = rng.standard_normal((T, (Hq + 2 * Hkv) * D)).astype(np.float16)
    out = kv.prefill_attention(qkv, cu, Hq, Hkv, D).reshape(T, Hq, D)
    q = torch.from_numpy(qkv[:, : Hq * D].astype(np.float64)).reshape(T, Hq, D)
    k = torch.from_numpy(qkv[:, Hq * D:(Hq + Hkv) * D].astype(np.float64)).reshape(T, Hkv, D)
    v = torch.from_numpy(qkv[:, (Hq + Hkv) * D:].astype(np.float64)).reshape(T, Hkv, D)
    for b, L in enumerate(lens):
        s, e = int(cu[b]), int(cu[b + 1])
        kk = k[s:e].repeat_interleave(Hq // Hkv, dim=1).transpose(0, 1)
        vv = v[s:e].repeat_interleave(Hq // Hkv, dim=1).transpose(0, 1)
        ref = torch.nn.functional.scaled_dot_product_attention(q[s:e].transpose(0, 1), kk, vv, is_causal=True).transpose(0, 1).numpy()
        assert np.abs(out[s:e] - ref).max() < 1e-9
    # causality: perturbing a later token of a sequence leaves the earlier rows untouched
    qkv2 = qkv.copy()
    qkv2[int(cu[3]) - 1] += 1
    out2 = kv.prefill_attention(qkv2, cu, Hq, Hkv, D).reshape(T, Hq, D)
    assert np.array_equal(out2[: int(cu[3]) - 1], out[: int(cu[3]) - 1])
    assert not np.array_equal(out2[int(cu[3]) - 1], out[int(cu[3]) - 1])
