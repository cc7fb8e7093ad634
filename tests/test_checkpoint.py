"""CPU: the checkpoint converter / loader (qserve_b200/checkpoint.py) against golden vectors produced by the
reference's own `from_linear` (tests/golden/make_golden_pack.py) and against the oracle's packer."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import w4a8
from qserve_b200 import checkpoint as ck

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pack_per_chn_*.npz"))))
def test_quantize_per_channel_matches_reference_from_linear(path):
    g = np.load(path)
    out = ck.quantize_w4a8(torch.from_numpy(g["w"]), torch.from_numpy(g["s1"]), torch.from_numpy(g["z"]).to(torch.int8))
    assert np.array_equal(out["qweight"].numpy(), g["qweight"])
    assert np.array_equal(out["s1_scales"].numpy().view(np.uint16), g["s1_scales"].view(np.uint16))
    assert np.array_equal(out["s1_szeros"].numpy().view(np.uint16), g["s1_szeros"].view(np.uint16))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pack_per_group_*.npz"))))
def test_quantize_per_group_matches_reference_from_linear(path):
    g = np.load(path)
    out = ck.quantize_w4a8(torch.from_numpy(g["w"]), torch.from_numpy(g["s1"]), torch.from_numpy(g["z"]).to(torch.int8),
                           torch.from_numpy(g["s2"].astype(np.float16)), group_size=128)
    for k in ("qweight", "s2_scales", "s2_zeros"):
        assert np.array_equal(out[k].numpy(), g[k]), k
    assert np.array_equal(out["s1_scales"].numpy().view(np.uint16), g["s1_scales"].view(np.uint16))


def test_pack_matches_oracle_and_roundtrips():
    rng = np.random.default_rng(0)
    q = rng.integers(0, 16, size=(96, 160)).astype(np.uint8)
    packed = ck.pack_int4(torch.from_numpy(q))
    assert np.array_equal(packed.numpy(), w4a8.pack_w4(q))
    assert np.array_equal(ck.unpack_int4(packed).numpy(), q)
    with pytest.raises(ValueError):
        ck.pack_int4(torch.full((32, 32), 16, dtype=torch.int16))
    with pytest.raises(ValueError):
        ck.pack_int4(torch.zeros((31, 32), dtype=torch.int16))


def test_quantize_w8a8():
    g = torch.Generator().manual_seed(1)
    s1 = torch.rand(64, generator=g) * 0.01 + 0.001
    q = torch.randint(-127, 128, (64, 96), generator=g)
    out = ck.quantize_w8a8(q.float() * s1[:, None], s1)
    assert torch.equal(out["weight"], q.to(torch.int8)) and out["dequant_scale"].dtype == torch.float32
    with pytest.raises(ValueError):
        ck.quantize_w8a8(torch.full((32, 32), 2.0), torch.full((32,), 0.01))


def _fake_checkpoint(layers, H, I, Hq, Hkv, D, group, seed=0):
    g = torch.Generator().manual_seed(seed)
    fake, params = {}, {}
    shapes = {"self_attn.q_proj": (Hq * D, H), "self_attn.k_proj": (Hkv * D, H), "self_attn.v_proj": (Hkv * D, H), "self_attn.o_proj": (H, Hq * D),
              "mlp.gate_proj": (I, H), "mlp.up_proj": (I, H), "mlp.down_proj": (H, I)}
    for i in range(layers):
        for lin, (N, K) in shapes.items():
            name = f"model.layers.{i}.{lin}.weight"
            s1 = torch.rand(N, generator=g) * 0.01 + 0.005
            q = torch.randint(0, 16, (N, K), generator=g)
            if group == -1:
                z = torch.randint(0, 16, (N,), generator=g)
                fake[name] = (q - z[:, None]).float() * s1[:, None]
            else:
                G = K // group
                z = torch.randint(0, 16, (N, G), generator=g)
                s2 = torch.randint(1, 9, (N, G), generator=g)
                fake[name] = ((q.reshape(N, G, group) - z[:, :, None]) * s2[:, :, None]).reshape(N, K).float() * s1[:, None]
                params[f"{name}.scale.1"] = s2.to(torch.float16)
            params[f"{name}.scale.0"] = s1
            params[f"{name}.zero"] = z
        fake[f"model.layers.{i}.input_layernorm.weight"] = torch.ones(H)
        fake[f"model.layers.{i}.post_attention_layernorm.weight"] = torch.ones(H)
    fake["model.embed_tokens.weight"] = torch.randn(64, H, generator=g)
    fake["model.norm.weight"] = torch.ones(H)
    fake["lm_head.weight"] = torch.randn(64, H, generator=g)
    return fake, params


@pytest.mark.parametrize("group", [-1, 128])
def test_convert_and_fuse_tensor_parallel_shards_reassemble(group):
    """convert -> fuse at TP = 2: the per-rank buffers are exact slices of the TP = 1 buffers, and a per-channel GEMM on the
    column-parallel shards / the sum over the row-parallel shards reproduces the unsharded INT32 accumulators (oracle)."""
    H, I, Hq, Hkv, D = 256, 512, 2, 2, 128
    fake, params = _fake_checkpoint(1, H, I, Hq, Hkv, D, group)
    sd = ck.convert_fake_quant_checkpoint(fake, params, num_layers=1, w_bit=4, group_size=group)
    full = ck.fuse_llama_state_dict(sd, 1)
    r0, r1 = ck.fuse_llama_state_dict(sd, 1, 0, 2), ck.fuse_llama_state_dict(sd, 1, 1, 2)
    pre = "model.layers.0."
    qkv = full[pre + "self_attn.qkv_proj.qweight"]
    assert qkv.shape == ((Hq + 2 * Hkv) * D, H // 2)
    # q | k | v blocks of the fused weight are the individual projections
    assert torch.equal(qkv[: Hq * D], sd[pre + "self_attn.q_proj.qweight"])
    assert torch.equal(qkv[Hq * D + Hkv * D:], sd[pre + "self_attn.v_proj.qweight"])
    # column parallel: rank r owns half of q, half of k, half of v (heads)
    half = Hq * D // 2
    assert torch.equal(r0[pre + "self_attn.qkv_proj.qweight"][:half], sd[pre + "self_attn.q_proj.qweight"][:half])
    assert torch.equal(r1[pre + "self_attn.qkv_proj.qweight"][:half], sd[pre + "self_attn.q_proj.qweight"][half:])
    # row parallel: summing the two K shards' integer products gives the unsharded product
    rng = np.random.default_rng(0)
    a = rng.integers(-127, 128, size=(4, I)).astype(np.int8)
    wq = w4a8.unpack_w4(full[pre + "mlp.down_proj.qweight"].numpy())
    parts = [w4a8.unpack_w4(r[pre + "mlp.down_proj.qweight"].numpy()) for r in (r0, r1)]
    assert np.array_equal(np.concatenate(parts, axis=1), wq)
    acc = sum(w4a8.int_matmul(a[:, i * I // 2:(i + 1) * I // 2], p) for i, p in enumerate(parts))
    assert np.array_equal(acc, w4a8.int_matmul(a, wq))
    if group == 128:
        assert r0[pre + "mlp.down_proj.s2_scales"].shape == (I // 128 // 2, H)
        assert full[pre + "mlp.gate_up_proj.s2_scales"].shape == (H // 128, 2 * I)
        assert torch.equal(full[pre + "mlp.gate_up_proj.s2_scales"][:, :I], sd[pre + "mlp.gate_proj.s2_scales"])
    else:
        assert torch.equal(full[pre + "mlp.gate_up_proj.s1_szeros"][I:], sd[pre + "mlp.up_proj.s1_szeros"])
    # the W4A8 loader drops every norm tensor (llama_w4a8_unpad.py:541-542); everything else passes through
    assert pre + "input_layernorm.weight" not in full and "model.norm.weight" not in full and "lm_head.weight" in full


def test_w4a8_loader_skips_norm_tensors_w8a8_loads_them():
    """a W4A8 checkpoint with NON-unit norm tensors must not get gamma applied twice -- the reference's W4A8 loader
    skips every tensor whose name contains 'norm' (gamma stays 1, LMQuant folded it into the next linear); the W8A8 loader keeps them."""
    H = 128
    sd = {"model.layers.0.input_layernorm.weight": torch.full((H,), 3.0), "model.layers.0.post_attention_layernorm.weight": torch.full((H,), 5.0),
          "model.norm.weight": torch.full((H,), 7.0), "model.embed_tokens.weight": torch.zeros(4, H), "lm_head.weight": torch.zeros(4, H)}
    w4 = ck.fuse_llama_state_dict(sd, 1, w_bit=4)
    assert not any("norm" in k for k in w4) and "lm_head.weight" in w4 and "model.embed_tokens.weight" in w4
    w8 = ck.fuse_llama_state_dict(sd, 1, w_bit=8)
    assert float(w8["model.layers.0.input_layernorm.weight"][0]) == 3.0 and float(w8["model.norm.weight"][0]) == 7.0

    class _Runner:  # load_into_runner only touches these attributes
        def __init__(self, wmode):
            self.wmode = wmode
            self.layers = [{"ln1": torch.ones(H), "ln2": torch.ones(H)}]
            self.embed, self.norm_w, self.lm_head = torch.ones(4, H), torch.ones(H), torch.ones(4, H)
    r4, r8 = _Runner("chn"), _Runner("w8")
    for k in ("qkv", "o", "gate_up", "down"):
        r4.layers[0][k] = object(); r8.layers[0][k] = object()
    ck.load_into_runner(r4, sd)   # even a dict that still carries the norms leaves a W4A8 runner at gamma = 1
    assert float(r4.layers[0]["ln1"][0]) == 1.0 and float(r4.layers[0]["ln2"][0]) == 1.0 and float(r4.norm_w[0]) == 1.0
    ck.load_into_runner(r8, w8)
    assert float(r8.layers[0]["ln1"][0]) == 3.0 and float(r8.layers[0]["ln2"][0]) == 5.0 and float(r8.norm_w[0]) == 7.0


def test_convert_rejects_inconsistent_group_metadata():
    fake, params = _fake_checkpoint(1, 128, 256, 1, 1, 128, -1)
    with pytest.raises(ValueError):
        ck.convert_fake_quant_checkpoint(fake, params, num_layers=1, w_bit=4, group_size=128)


# ---- property tests (hypothesis): layout identities that must hold for every shape ----
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=25, deadline=None)
@given(n32=st.integers(1, 6), k32=st.integers(1, 8), seed=st.integers(0, 2**31 - 1))
def test_pack_unpack_is_a_bijection_for_every_shape(n32, k32, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randint(0, 16, (32 * n32, 32 * k32), generator=g, dtype=torch.int16)
    packed = ck.pack_int4(q)
    assert packed.shape == (32 * n32, 16 * k32) and packed.dtype == torch.int8
    assert torch.equal(ck.unpack_int4(packed).to(torch.int16), q)
    # one 32x32 tile is one contiguous 512-byte block: changing a single code touches exactly one byte of that block
    q2 = q.clone()
    r, c = int(seed % (32 * n32)), int((seed // 7) % (32 * k32))
    q2[r, c] = (q2[r, c] + 1) % 16
    diff = (ck.pack_int4(q2) != packed).reshape(n32, k32, 512).nonzero()
    assert diff.shape[0] == 1 and int(diff[0, 0]) == r // 32 and int(diff[0, 1]) == c // 32


@settings(max_examples=15, deadline=None)
@given(tiles_n=st.integers(1, 3), blocks_k=st.integers(1, 3), size=st.sampled_from([1, 2, 4]), seed=st.integers(0, 2**31 - 1))
def test_row_and_column_shards_tile_the_unsharded_weight(tiles_n, blocks_k, size, seed):
    """tp.shard_rows / shard_columns on the packed layout == slicing the unpacked code matrix (any TP size)."""
    from qserve_b200 import tp
    N, K = 128 * size * tiles_n, 128 * size * blocks_k
    g = torch.Generator().manual_seed(seed)
    q = torch.randint(0, 16, (N, K), generator=g, dtype=torch.int16)
    packed = ck.pack_int4(q)
    for r in range(size):
        cols = ck.unpack_int4(tp.shard_columns(packed, r, size)).to(torch.int16)
        assert torch.equal(cols, q[r * N // size:(r + 1) * N // size])
        rows = ck.unpack_int4(tp.shard_rows(packed, r, size)).to(torch.int16)
        assert torch.equal(rows, q[:, r * K // size:(r + 1) * K // size])
