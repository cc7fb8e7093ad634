"""Run the reference's UNMODIFIED Python model code (`qserve/modeling/**`) over this repo's `qserve_backend`.

north_star: "... keeps the qserve_backend torch-extension op signatures so qserve/modeling and the in-flight-batching engine
drop in unchanged".  This module is the proof harness for that sentence: it imports the reference's
`LlamaForCausalLM` (qserve/modeling/models/llama_w4a8_unpad.py:417-477) -- which in turn imports the reference's own
`W4A8OF16LinearDynamicInputScale`, `RMSNormGeneral`, `SiluAndMulQuant`, `InputMetadata` / `ActivationBuffer` -- with
`qserve_backend` resolving to THIS repo, fills its buffers with the same synthetic weights / KV pages a `DecodeRunner` owns
(the tensors are shared, not copied), and drives decode and prefill steps exactly as `ModelRunner.execute_model` does
(qserve/worker/model_runner.py:333-441, 445-549).

Nothing from the reference is vendored: its Python package is found at run time under `baseline/_ref/` (a
`pip install --no-deps --target baseline/_ref /root/reference`, done by `__graft_entry__.build()`; git-ignored, travels to
the GPU box) or, in the build container, directly under `/root/reference`.  The only shim is `xformers.ops.AttentionBias`
(imported by qserve/utils/input_metadata.py:12 for a type hint; xformers is not installed in this image).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import List, Optional

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CANDIDATES = (os.path.join(_ROOT, "baseline", "_ref"), "/root/reference")


def locate_reference() -> Optional[str]:
    for p in REF_CANDIDATES:
        if os.path.isfile(os.path.join(p, "qserve", "modeling", "models", "llama_w4a8_unpad.py")):
            return p
    return None


def import_reference(w8a8: bool = False):
    """Import the reference model module with `qserve_backend` = this repo.  Needs a CUDA device: the reference evaluates
    `torch.cuda.current_device()` at class-definition time (w4a8_linear.py:19)."""
    ref = locate_reference()
    if ref is None:
        raise ImportError("reference Python package not found (expected baseline/_ref/qserve: run __graft_entry__.build() where /root/reference exists)")
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)  # qserve_backend (this repo) must win over anything else of that name
    if ref not in sys.path:
        sys.path.append(ref)
    if "xformers" not in sys.modules:  # type-hint-only import (input_metadata.py:12)
        xf, xo = types.ModuleType("xformers"), types.ModuleType("xformers.ops")
        xo.AttentionBias = object
        xf.ops = xo
        sys.modules["xformers"], sys.modules["xformers.ops"] = xf, xo
    import qserve_backend  # noqa: F401
    assert os.path.dirname(os.path.abspath(qserve_backend.__file__)).startswith(_ROOT), "qserve_backend does not resolve to this repo"
    try:
        import flash_attn  # noqa: F401  (third-party prompt attention, imported at module scope by llama_w4a8_unpad.py:27)
    except Exception:  # noqa: BLE001  not installed: the reference then runs on this repo's prompt attention alone
        from qserve_b200 import backend
        fa = types.ModuleType("flash_attn")
        fa.flash_attn_varlen_func = backend.flash_attn_varlen_func
        fa.__version__ = "qserve_b200"
        sys.modules["flash_attn"] = fa
    name = "qserve.modeling.models.llama_w8a8_unpad" if w8a8 else "qserve.modeling.models.llama_w4a8_unpad"
    return importlib.import_module(name)


class RefModel:
    """The reference `LlamaForCausalLM` sharing the weights, KV pools and page tables of a `DecodeRunner`."""

    def __init__(self, runner):
        from transformers import LlamaConfig

        assert runner.tp_size == 1, "the reference model code has tp_size = 1 hard-coded (llama_w4a8_unpad.py:115)"
        self.runner = runner
        cfg = runner.cfg
        w8 = runner.wmode == "w8"
        self.mod = mod = import_reference(w8a8=w8)
        from qserve.sampling_params import SamplingParams

        hf = LlamaConfig(hidden_size=cfg.hidden, intermediate_size=cfg.intermediate, num_attention_heads=cfg.heads,
                         num_key_value_heads=cfg.kv_heads, num_hidden_layers=runner.L, vocab_size=cfg.vocab, rms_norm_eps=cfg.eps,
                         max_position_embeddings=cfg.max_pos, rope_theta=cfg.rope_theta)
        hf.rope_theta = cfg.rope_theta  # transformers 5 moves it into rope_parameters; the reference reads the attribute (:103)
        kv_cfg = {"INT4_ENABLED": runner.kv_bits == 4, "ZEROS_ENABLED": True}  # model_runner.py:126-133, arg_utils.py:422
        sp = SamplingParams(temperature=0.0, top_p=1.0, top_k=-1)  # greedy (sampler.py:87-90)
        with torch.cuda.device(runner.dev):
            if w8:
                model = mod.LlamaForCausalLM(hf, sp, kv_cache_config=kv_cfg)
            else:
                model = mod.LlamaForCausalLM(hf, -1 if runner.wmode == "chn" else 128, sp, kv_cache_config=kv_cfg)
            model = model.half().to(runner.dev)  # model_runner.py:148-150
        self.model = model
        self._share_weights()
        self._meta = None
        self._flash_attn = mod.flash_attn_varlen_func  # whatever `from flash_attn import flash_attn_varlen_func` bound (llama_w4a8_unpad.py:27)

    def use_prefill_attention(self, impl: str) -> None:
        """Choose the prompt-phase attention behind the reference layer's `flash_attn_varlen_func(...)` call (llama_w4a8_unpad.py:232-242):
        "flash_attn" = the third-party package as imported by the reference, "qserve_b200" = this repo's tcgen05 kernel.  The reference
        source is not touched: the module-level name it calls through is rebound."""
        from qserve_b200 import backend

        assert impl in ("flash_attn", "qserve_b200")
        self.mod.flash_attn_varlen_func = backend.flash_attn_varlen_func if impl == "qserve_b200" else self._flash_attn

    # -----------------------------------------------------------------------------------------------------------
    def _share_weights(self) -> None:
        run, m = self.runner, self.model
        m.model.embed_tokens.weight.data = run.embed
        m.lm_head.weight.data = run.lm_head
        m.model.norm.weight.data = run.norm_w
        for li, ly in enumerate(run.layers):
            rl = m.model.layers[li]
            for mine, theirs in (("qkv", rl.self_attn.qkv_proj), ("o", rl.self_attn.o_proj), ("gate_up", rl.mlp.gate_up_proj), ("down", rl.mlp.down_proj)):
                lin = ly[mine]
                if run.wmode == "w8":
                    assert tuple(theirs.weight.shape) == tuple(lin.weight.shape)
                    theirs.weight = lin.weight
                    theirs.dequant_scale = lin.wscale.float()  # w8a8_linear.py: fp32 buffer, .half()-ed at every call (:99-101)
                else:
                    assert tuple(theirs.qweight.shape) == tuple(lin.qweight.shape), (theirs.qweight.shape, lin.qweight.shape)
                    theirs.qweight = lin.qweight
                    theirs.s1_scales = lin.s1
                    if run.wmode == "chn":
                        theirs.s1_szeros = lin.s1z
                    else:
                        theirs.s2_scales, theirs.s2_zeros = lin.s2_scales, lin.s2_zeros
            rl.input_layernorm.weight.data = ly["ln1"]
            rl.post_attention_layernorm.weight.data = ly["ln2"]

    # -----------------------------------------------------------------------------------------------------------
    def decode_metadata(self, fresh: bool = False):
        """InputMetadata of a decode step as `_prepare_decode_*` builds it (model_runner.py:445-642): block tables
        [L, B, 2, blocks] int64, context_lens including the current token, max_seq_len = max context."""
        if self._meta is not None and not fresh:
            return self._meta
        from qserve.utils.input_metadata import InputMetadata

        run = self.runner
        meta = InputMetadata(is_prompt=False, context_lens=run.context_lens, padding_offsets=None, cu_seqlens=None, max_seq_len=run.max_seq_len,
                             max_block_table_len=run.blocks_per_seq, block_tables=run.block_tables, kv_cache_dtype="int8", kv_scales=None,
                             batched_seq_len=run.batch, model=self.model)
        self._meta = meta
        return meta

    @torch.no_grad()
    def decode_logits(self, tokens: torch.Tensor, fresh_metadata: bool = False) -> torch.Tensor:
        """One decode step through the reference model code: logits [B, vocab] (LlamaForCausalLM.forward, :464-477)."""
        return self.model(tokens, self.decode_metadata(fresh_metadata))

    @torch.no_grad()
    def decode_tokens(self, tokens: torch.Tensor, fresh_metadata: bool = False) -> torch.Tensor:
        meta = self.decode_metadata(fresh_metadata)
        logits = self.model(tokens, meta)
        return self.model.sample(tokens, logits, meta)

    # -----------------------------------------------------------------------------------------------------------
    def prefill_metadata(self, prompt_lens: List[int]):
        """InputMetadata of a prompt step as `_prepare_prompt` builds it (model_runner.py:333-441)."""
        import qserve_backend.fused_attention as fused_attention
        from qserve.utils.input_metadata import InputMetadata

        run = self.runner
        assert len(prompt_lens) <= run.batch and max(prompt_lens) <= run.blocks_per_seq * 64
        B = len(prompt_lens)
        ctx = torch.tensor(prompt_lens, dtype=torch.int, device=run.dev)
        cu = torch.nn.functional.pad(torch.cumsum(ctx, dim=0).int(), (1, 0), value=0)
        total = int(sum(prompt_lens))
        pad = fused_attention.compute_padding_offsets(cu, max(prompt_lens), total)
        return InputMetadata(is_prompt=True, context_lens=ctx, padding_offsets=pad, cu_seqlens=cu, max_seq_len=max(prompt_lens),
                             max_block_table_len=run.blocks_per_seq, block_tables=run.block_tables[:, :B].contiguous(), kv_cache_dtype="int8",
                             kv_scales=None, batched_seq_len=total, model=self.model)

    @torch.no_grad()
    def prefill_logits(self, tokens: torch.Tensor, prompt_lens: List[int], meta=None) -> torch.Tensor:
        """One prompt step (apply_bias_rope_update_kv_cache + flash_attn_varlen_func inside the reference layer code,
        llama_w4a8_unpad.py:203-242): last-token logits [B, vocab]; the KV pages of the first len(prompt_lens) sequences are rewritten."""
        meta = meta if meta is not None else self.prefill_metadata(prompt_lens)
        return self.model(tokens, meta)
