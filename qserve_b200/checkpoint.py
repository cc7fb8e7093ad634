"""QServe checkpoints: produce and consume the reference's real-quantised state dict.

* `quantize_w4a8` / `quantize_w8a8` -- the arithmetic of `W4A8OF16LinearDynamicInputScale.from_linear`
  (qserve/modeling/layers/quantized_linear/w4a8_linear.py:136-332) and `W8A8OF16LinearDynamicInputScale.from_linear`
  (w8a8_linear.py:112-150): a fake-quantised fp16/fp32 linear weight plus its LMQuant scales / zeros in, the buffers the kernels
  consume out (`qweight`, `s1_scales`, `s1_szeros` | `s2_scales`, `s2_zeros` | `weight`, `dequant_scale`).  The packed layout is
  written with one index permutation instead of the reference's reshape/permute chain; tests/test_checkpoint.py checks both
  against golden vectors produced by the reference's own `from_linear`.
* `convert_fake_quant_checkpoint` -- `scripts/ckpt_converter/checkpoint_converter.py:81-121` for a Llama-style state dict.
* `fuse_llama_state_dict` -- the q/k/v -> qkv_proj and gate/up -> gate_up_proj fusion and the (tensor-parallel) slicing of
  `LlamaForCausalLM.load_weights` (qserve/modeling/models/llama_w4a8_unpad.py:487-630); the reference leaves `tp_size = 1`
  (:510-514), here the slices follow qserve_b200/tp.py.
* `load_into_runner` -- installs a fused state dict into `qserve_b200.decode.DecodeRunner`.

Pure torch, device agnostic (runs on CPU tensors; nothing here launches a kernel of the library).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import tp

GROUP = 128  # the only group size the per-group kernel instantiates (w4a8_per_group/gemm_cuda.cu:652)


# ------------------------------------------------------------------------------------------------------------
# packing
# ------------------------------------------------------------------------------------------------------------
def pack_int4(q: torch.Tensor) -> torch.Tensor:
    """uint4 codes [N, K] (any integer dtype, values 0..15) -> qweight int8 [N, K/2] in the checkpoint layout.

    A 32x32 tile is 512 contiguous bytes = 32 lanes x 16 B; lane t = c*4 + e, byte j = d*8 + b*4 + f holds
    q[32 n32 + 8 b + c, 32 k32 + 16 d + 4 e + f] in its low and the same element of row +16 in its high nibble
    (the consumer side is gemm_cuda.cu:286-298)."""
    N, K = q.shape
    if N % 32 or K % 32:
        raise ValueError(f"pack_int4: N={N} and K={K} must be multiples of 32")
    q = q.to(torch.int16)
    if int(q.min()) < 0 or int(q.max()) > 15:
        raise ValueError("pack_int4: codes out of the uint4 range")
    # rows n = 32 n32 + 16 h + 8 b + c ; columns k = 32 k32 + 16 d + 4 e + f
    v = q.reshape(N // 32, 2, 2, 8, K // 32, 2, 4, 4)                 # [n32, h, b, c, k32, d, e, f]
    v = v.permute(1, 0, 4, 3, 6, 5, 2, 7)                             # [h, n32, k32, c, e, d, b, f]
    byte = v[0] | (v[1] << 4)
    return byte.to(torch.uint8).contiguous().view(torch.int8).reshape(N, K // 2)


def unpack_int4(qweight: torch.Tensor) -> torch.Tensor:
    """Inverse of `pack_int4`: qweight int8 [N, K/2] -> uint8 codes [N, K]."""
    N, K2 = qweight.shape
    K = 2 * K2
    p = qweight.contiguous().view(torch.uint8).reshape(N // 32, K // 32, 8, 4, 2, 2, 4).to(torch.int16)  # [n32, k32, c, e, d, b, f]
    lo, hi = p & 0xF, p >> 4
    v = torch.stack([lo, hi], dim=0)                                   # [h, n32, k32, c, e, d, b, f]
    v = v.permute(1, 0, 6, 3, 2, 5, 4, 7)                              # [n32, h, b, c, k32, d, e, f]
    return v.reshape(N, K).to(torch.uint8)


def _shuffle_per_32_columns(p: torch.Tensor) -> torch.Tensor:
    """Level-2 params [K/128, N]: position c*4 + j of every 32-column group <- channel j*8 + c (w4a8_linear.py:236-249)."""
    G, N = p.shape
    return p.reshape(G, N // 32, 4, 8).transpose(-2, -1).reshape(G, N).contiguous()


# ------------------------------------------------------------------------------------------------------------
# quantisation (from_linear)
# ------------------------------------------------------------------------------------------------------------
def quantize_w4a8(weight: torch.Tensor, s1_scale: torch.Tensor, zeros: torch.Tensor, s2_scale: Optional[torch.Tensor] = None,
                  group_size: int = -1) -> Dict[str, torch.Tensor]:
    """Fake-quantised weight [N, K] + scales / zero points -> the W4A8 kernel buffers.

    per-channel (group_size = -1, w4a8_linear.py:278-330):  q = round(W / s1) + z  in [0, 15];  s1_szeros = z * s1 (fp16 product)
    per-group  (group_size = 128, :166-277):  w8 = round(W / s1) in int8;  q = w8 / s2 + z  in [0, 15];
        s2_scales [K/128, N] = s2 (shuffled per 32 columns),  s2_zeros = (-z) * s2 (int32 product stored as int8, same shuffle)."""
    N, K = weight.shape
    s1 = s1_scale.reshape(N)
    w = weight.clone().div_(s1.reshape(N, 1).to(weight.dtype)).round_()   # in place on a copy, like the reference's div_ / round_
    out: Dict[str, torch.Tensor] = {"s1_scales": s1.to(torch.float16).clone()}
    if group_size == -1:
        q = w + zeros.reshape(N, 1).to(w.dtype)   # exact small integers in floating point: range-checked before the int8 cast
        if float(q.min()) < 0 or float(q.max()) > 15:
            raise ValueError("quantize_w4a8: quantised weight out of range (per-channel)")
        out["qweight"] = pack_int4(q.to(torch.int8))
        # torch type promotion: int8 * fp16 -> fp16 product (w4a8_linear.py:328-330)
        out["s1_szeros"] = (zeros.reshape(N).to(torch.int8) * s1.to(torch.float16)).to(torch.float16)
        return out
    if group_size != GROUP:
        raise ValueError("quantize_w4a8: only group_size -1 or 128 exist in the reference kernels")
    if s2_scale is None:
        raise ValueError("quantize_w4a8: per-group quantisation needs s2_scale")
    if float(w.min()) < -128 or float(w.max()) > 127:
        raise ValueError("quantize_w4a8: stage-1 quantised weight out of the int8 range")
    G = K // GROUP
    z = zeros.reshape(N, G, 1)
    s2 = s2_scale.reshape(N, G, 1)
    q = w.reshape(N, G, GROUP).div_(s2.to(torch.float16).to(w.dtype)).add_(z.to(torch.float16).to(w.dtype))
    if float(q.min()) < 0 or float(q.max()) > 15:
        raise ValueError("quantize_w4a8: stage-2 quantised weight out of range")
    out["qweight"] = pack_int4(q.reshape(N, K).to(torch.int8))
    s2_t = _shuffle_per_32_columns(s2.reshape(N, G).transpose(0, 1).contiguous())
    z_t = _shuffle_per_32_columns((-z).int().reshape(N, G).transpose(0, 1).contiguous())
    out["s2_scales"] = s2_t.to(torch.int8)
    out["s2_zeros"] = (z_t * s2_t).to(torch.int8)   # int32 x scale dtype, truncated to int8 on assignment (w4a8_linear.py:271-277)
    return out


def quantize_w8a8(weight: torch.Tensor, s1_scale: torch.Tensor) -> Dict[str, torch.Tensor]:
    """w8a8_linear.py:134-150: weight int8 [N, K] = round(W / s1) in [-127, 127], dequant_scale[N] = s1."""
    N = weight.size(0)
    w = weight.clone().div_(s1_scale.reshape(N, 1).to(weight.dtype)).round_()
    if float(w.min()) < -127 or float(w.max()) > 127:  # checked BEFORE the int8 cast (the reference's assert comes after it and wraps)
        raise ValueError("quantize_w8a8: quantised weight out of range")
    return {"weight": w.to(torch.int8).contiguous(), "dequant_scale": s1_scale.reshape(N).float().clone()}


_LINEARS = ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj")


def convert_fake_quant_checkpoint(fake_quant: Dict[str, torch.Tensor], quant_params: Dict[str, torch.Tensor], num_layers: int, w_bit: int = 4,
                                  group_size: int = -1) -> Dict[str, torch.Tensor]:
    """checkpoint_converter.py:81-121: LMQuant's `model.pt` (fake-quantised fp weights) + `scale.pt` -> real-quantised state dict
    with the reference's tensor names (`model.layers.{i}.{linear}.{qweight|s1_scales|...}`); every other tensor is copied."""
    out: Dict[str, torch.Tensor] = {}
    done = set()
    for i in range(num_layers):
        for lin in _LINEARS:
            name = f"model.layers.{i}.{lin}"
            wkey = f"{name}.weight"
            if wkey not in fake_quant:
                continue
            s1 = quant_params[f"{wkey}.scale.0"]
            if w_bit == 8:
                q = quantize_w8a8(fake_quant[wkey].float(), s1.float())
            else:
                zeros = quant_params[f"{wkey}.zero"].to(torch.int8)
                if int(zeros.min()) < 0:          # symmetric-range zero points are stored offset by -8 (checkpoint_converter.py:102-103)
                    zeros = zeros + 8
                s2 = quant_params.get(f"{wkey}.scale.1") if group_size != -1 else None
                if (group_size != -1) != (f"{wkey}.scale.1" in quant_params):
                    raise ValueError(f"{name}: level-2 scales present/absent does not match group_size={group_size}")
                q = quantize_w4a8(fake_quant[wkey].float(), s1.float(), zeros, s2, group_size)
            for k, v in q.items():
                out[f"{name}.{k}"] = v
            done.add(wkey)
    for k, v in fake_quant.items():
        if k not in done:
            out[k] = v
    return out


# ------------------------------------------------------------------------------------------------------------
# loading (LlamaForCausalLM.load_weights)
# ------------------------------------------------------------------------------------------------------------
_COLUMN = {"qweight": tp.shard_columns, "weight": tp.shard_columns, "s1_scales": tp.shard_vector, "s1_szeros": tp.shard_vector,
           "dequant_scale": tp.shard_vector, "s2_scales": tp.shard_level2_columns, "s2_zeros": tp.shard_level2_columns}
_ROW = {"qweight": tp.shard_rows, "s2_scales": tp.shard_level2_rows, "s2_zeros": tp.shard_level2_rows}


def _shard_rows_int8(w: torch.Tensor, rank: int, size: int) -> torch.Tensor:
    k = w.size(1) // size
    return w[:, rank * k:(rank + 1) * k].contiguous()


def _cat(parts, suffix: str) -> torch.Tensor:
    # level-2 params are [K/128, N]: fused layers concatenate along N = dim 1 (llama_w4a8_unpad.py:576-579, 596-600)
    return torch.cat(parts, dim=1 if suffix in ("s2_scales", "s2_zeros") else 0).contiguous()


def fuse_llama_state_dict(sd: Dict[str, torch.Tensor], num_layers: int, tp_rank: int = 0, tp_size: int = 1, w_bit: int = 4) -> Dict[str, torch.Tensor]:
    """Per-projection checkpoint tensors -> the fused, per-rank buffers the model owns: `qkv_proj` = q | k | v and
    `gate_up_proj` = gate | up along the output channels (column parallel: every part is sliced first, then concatenated),
    `o_proj` / `down_proj` sliced along K (row parallel; per-channel vectors are replicated).

    `w_bit == 4` reproduces the W4A8 loader's `if "norm" in name: continue` (llama_w4a8_unpad.py:541-542): every norm tensor of the
    checkpoint (input_layernorm, post_attention_layernorm AND the final model.norm) is dropped and gamma stays 1, because LMQuant has
    folded it into the following linear layer.  Only the W8A8 loader (`w_bit == 8`, llama_w8a8_unpad.py) loads them."""
    if w_bit not in (4, 8):
        raise ValueError(f"w_bit must be 4 or 8, got {w_bit}")
    load_norms = (w_bit == 8)
    out: Dict[str, torch.Tensor] = {}
    suffixes = sorted({k.rsplit(".", 1)[1] for k in sd if "_proj." in k})
    for i in range(num_layers):
        pre = f"model.layers.{i}."
        for fused, parts in (("self_attn.qkv_proj", ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj")),
                             ("mlp.gate_up_proj", ("mlp.gate_proj", "mlp.up_proj"))):
            for suf in suffixes:
                keys = [f"{pre}{p}.{suf}" for p in parts]
                if not all(k in sd for k in keys):
                    continue
                out[f"{pre}{fused}.{suf}"] = _cat([_COLUMN[suf](sd[k], tp_rank, tp_size) for k in keys], suf)
        for lin in ("self_attn.o_proj", "mlp.down_proj"):
            for suf in suffixes:
                key = f"{pre}{lin}.{suf}"
                if key not in sd:
                    continue
                t = sd[key]
                if suf == "weight":
                    t = _shard_rows_int8(t, tp_rank, tp_size)
                elif suf in _ROW:
                    t = _ROW[suf](t, tp_rank, tp_size)
                out[key] = t.contiguous()
        for k in (f"{pre}input_layernorm.weight", f"{pre}post_attention_layernorm.weight"):
            if k in sd and load_norms:
                out[k] = sd[k]
    for k in ("model.embed_tokens.weight", "model.norm.weight", "lm_head.weight"):
        if k in sd and (load_norms or "norm" not in k):
            out[k] = sd[k]
    return out


def load_into_runner(runner, fused: Dict[str, torch.Tensor]) -> int:
    """Install a fused (per-rank) state dict into a DecodeRunner built for the same model / precision.  Returns the number of
    tensors installed; shape or dtype mismatches raise.  A W4A8 runner never takes norm tensors (gamma = 1, see
    fuse_llama_state_dict), even if the dict still carries them."""
    take_norms = runner.wmode == "w8"
    names = {"qkv": "self_attn.qkv_proj", "o": "self_attn.o_proj", "gate_up": "mlp.gate_up_proj", "down": "mlp.down_proj"}
    attr = {"qweight": "qweight", "s1_scales": "s1", "s1_szeros": "s1z", "s2_scales": "s2_scales", "s2_zeros": "s2_zeros",
            "weight": "weight", "dequant_scale": "wscale"}
    n = 0

    def put(dst: torch.Tensor, src: torch.Tensor, what: str):
        nonlocal n
        if tuple(dst.shape) != tuple(src.shape):
            raise ValueError(f"{what}: checkpoint shape {tuple(src.shape)} != model shape {tuple(dst.shape)}")
        dst.copy_(src.to(dst.dtype))
        n += 1

    for i, ly in enumerate(runner.layers):
        for short, long in names.items():
            lin = ly[short]
            for suf, a in attr.items():
                key = f"model.layers.{i}.{long}.{suf}"
                if key in fused and hasattr(lin, a):
                    put(getattr(lin, a), fused[key], key)
        for key, slot in ((f"model.layers.{i}.input_layernorm.weight", "ln1"), (f"model.layers.{i}.post_attention_layernorm.weight", "ln2")):
            if key in fused and take_norms:
                put(ly[slot], fused[key], key)
    for key, t in (("model.embed_tokens.weight", runner.embed), ("model.norm.weight", runner.norm_w), ("lm_head.weight", runner.lm_head)):
        if key in fused and (take_norms or "norm" not in key):
            put(t, fused[key], key)
    return n
