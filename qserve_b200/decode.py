"""Decode-step runner: the reference's per-layer op sequence over the drop-in `qserve_backend` API, with persistent
activation buffers and whole-step CUDA-graph capture.

This is the measurement harness for BASELINE.json's metric (tokens/s, Llama-3-8B W4A8KV4 decode) and the
"step-level CUDA-graph runner under the unchanged model code" of SURVEY.md section 8f-2.  It issues exactly the calls
`LlamaDecoderLayer.forward` issues (qserve/modeling/models/llama_w4a8_unpad.py:330-361, 69-93, 186-291; W8A8:
llama_w8a8_unpad.py) with the same argument marshalling, on synthetic weights of the named architecture
(random INT4/INT8 codes, scales chosen so that activations stay O(1); there is no network for checkpoints).

Tensor parallelism is Megatron style: qkv / gate_up column parallel, o_proj / down_proj row parallel
(split along K in multiples of 128), KV heads sharded, one NCCL sum-allreduce of [M, hidden] fp16 after each row-parallel
GEMM.  Two quantisation modes for the inputs of the row-parallel GEMMs:
  * `tp_exact=True` -- SURVEY.md 8e parity rule: the per-token amax is max-all-reduced ([M] fp32) so every rank uses the SAME
    scale, the activation sum stays the local K shard's: the INT32 partial sums of the ranks add up to the single-GPU
    accumulators bit for bit (tests/test_gpu_tp.py); costs one extra tiny collective per row-parallel GEMM;
  * `tp_exact=False` (throughput mode) -- every rank quantises its shard with its own per-token scale inside the fused
    attention / silu kernels; each partial product is a correctly de-quantised partial sum (agrees with the single-GPU run to
    quantisation noise, not bit for bit).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

import qserve_backend.activation_ops as activation_ops
import qserve_backend.fused_attention as fused_attention
import qserve_backend.fused_kernels as fused_kernels
import qserve_backend.layernorm_ops as layernorm_ops
import qserve_backend.qgemm_w4a8_per_chn as qgemm_chn
import qserve_backend.qgemm_w4a8_per_group as qgemm_grp
import qserve_backend.qgemm_w8a8 as qgemm_w8
from qserve_b200 import backend as _ext


@dataclass(frozen=True)
class OpSet:
    """The seven `qserve_backend` modules a step is made of.  Default: this repo's sm_100a library.  `bench.py --impl reference-gpu`
    and the live parity tests pass the reference's own extensions (oracle/_ref, compiled unmodified for sm_100a) instead, so the
    SAME op sequence runs on the legacy mma.sync kernels -- same buffers, same shapes, same harness."""
    layernorm_ops: object = layernorm_ops
    fused_kernels: object = fused_kernels
    activation_ops: object = activation_ops
    fused_attention: object = fused_attention
    qgemm_chn: object = qgemm_chn
    qgemm_grp: object = qgemm_grp
    qgemm_w8: object = qgemm_w8


DEFAULT_OPS = OpSet()


from qserve_b200.modelcfg import MODELS, PRECISIONS, ModelConfig  # noqa: E402,F401  (pure-Python table; re-exported)


class _Linear:
    """Weights of one quantised linear layer in the reference's buffer layout (w4a8_linear.py:38-103, w8a8_linear.py:45-54)."""

    def __init__(self, N: int, K: int, mode: str, dev, gen, ops: OpSet = DEFAULT_OPS):
        self.N, self.K, self.mode, self.ops = N, K, mode, ops
        r = lambda lo, hi, shape, dt: torch.randint(lo, hi, shape, dtype=dt, device=dev, generator=gen)
        u = lambda lo, hi, shape: (torch.rand(shape, device=dev, generator=gen) * (hi - lo) + lo)
        target = 1.0 / (K ** 0.5)  # output std ~ O(1) for unit-variance inputs
        if mode == "w8":
            self.weight = r(-127, 128, (N, K), torch.int8)
            self.wscale = (u(0.8, 1.2, (N,)) * target / 73.0).half()
        else:
            self.qweight = r(-128, 128, (N, K // 2), torch.int8)  # uniform random nibbles
            if mode == "chn":
                self.s1 = (u(0.8, 1.2, (N,)) * target / 4.6).half()
                z = r(7, 9, (N,), torch.int8).float()
                self.s1z = (z * self.s1.float()).half()
            else:
                g = K // 128
                s2 = r(1, 9, (g, N), torch.int8)
                z = r(7, 9, (g, N), torch.int8)
                self.s2_scales = s2.contiguous()
                self.s2_zeros = (-(z.int()) * s2.int()).to(torch.int8).contiguous()
                self.s1 = (u(0.8, 1.2, (N,)) * target / (4.6 * 4.5)).half()

    def __call__(self, x_q, scale, asum, out):
        if self.mode == "chn":    # w4a8_linear.py:106-115
            self.ops.qgemm_chn.gemm_forward_cuda(x_q, self.qweight, self.s1, scale, self.s1z, asum, out)
        elif self.mode == "grp":  # w4a8_linear.py:121-131
            self.ops.qgemm_grp.gemm_forward_cuda(x_q, self.qweight, self.s2_zeros, self.s2_scales, self.s1, scale, out)
        else:                     # w8a8_linear.py:98-101
            self.ops.qgemm_w8.w8a8_gemm_forward_cuda(x_q, self.weight, self.wscale, scale, out)

    def weight_bytes(self) -> int:
        return self.N * self.K if self.mode == "w8" else self.N * self.K // 2


class DecodeRunner:
    def __init__(self, model: str = "llama-3-8b", precision: str = "w4a8kv4", batch: int = 64, ctx: int = 1024,
                 device: Optional[torch.device] = None, tp_rank: int = 0, tp_size: int = 1, seed: int = 0, layers: Optional[int] = None,
                 process_group=None, fused: bool = True, ops: Optional[OpSet] = None, tp_exact: bool = False, tp_peer: bool = False, no_comm: bool = False):
        assert precision in PRECISIONS, precision
        self.ops = ops = ops or DEFAULT_OPS
        self.tp_exact = tp_exact
        self.fuse_attn_quant = True  # measured (profiles/r02_notes.md): attention + a separate per-token quant kernel is 1.3 % slower per step
        self.no_comm = no_comm  # debugging: run one rank's shard of a tensor-parallel model without the collectives (sanitizer / profiler runs)
        # tensor parallel, fused path: the all-reduce of the row-parallel GEMM outputs is folded into the following add+norm+quant kernel
        # (peer loads over NVLink symmetric memory) instead of an NCCL call
        self.tp_peer = tp_peer and tp_size > 1
        assert not (self.tp_peer and not fused), "tp_peer needs the fused path"
        assert ops is DEFAULT_OPS or not fused, "the fused extensions exist only in this repo's library"
        self.cfg = cfg = MODELS[model]
        self.precision, self.batch, self.ctx = precision, batch, ctx
        self.dev = dev = device or torch.device("cuda", torch.cuda.current_device())
        self.tp_rank, self.tp_size, self.pg = tp_rank, tp_size, process_group
        self.fused = fused  # fused residual+norm+quant and silu*mul+quant extensions (bit-identical to the reference op sequence)
        self.L = layers if layers is not None else cfg.layers
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed + 1000 * tp_rank)
        self.wmode = "w8" if precision.startswith("w8a8") else ("grp" if precision.endswith("g128") else "chn")
        self.act_sum = self.wmode == "chn"  # per-channel W4 needs the activation sum (llama_w4a8_unpad.py:69, 165-168)
        self.kv_bits = 4 if "kv4" in precision else 8
        D, H, I = cfg.head_dim, cfg.hidden, cfg.intermediate
        assert cfg.heads % tp_size == 0 and I % (128 * tp_size) == 0
        self.Hq = cfg.heads // tp_size
        self.Hkv = max(1, cfg.kv_heads // tp_size)  # llama_w4a8_unpad.py:120-129
        self.Iloc = I // tp_size
        self.q_size, self.kv_size = self.Hq * D, self.Hkv * D
        M = batch

        # ---- weights --------------------------------------------------------------------------------------
        self.layers = []
        for _ in range(self.L):
            self.layers.append({
                "qkv": _Linear(self.q_size + 2 * self.kv_size, H, self.wmode, dev, gen, ops),
                "o": _Linear(H, self.q_size, self.wmode, dev, gen, ops),
                "gate_up": _Linear(2 * self.Iloc, H, self.wmode, dev, gen, ops),
                "down": _Linear(H, self.Iloc, self.wmode, dev, gen, ops),
                # W4A8 checkpoints skip the norm weights (gamma = 1, llama_w4a8_unpad.py:541-542); W8A8 loads them
                "ln1": torch.ones(H, dtype=torch.half, device=dev),
                "ln2": torch.ones(H, dtype=torch.half, device=dev),
            })
        self.norm_w = torch.ones(H, dtype=torch.half, device=dev)
        self.embed = (torch.randn((cfg.vocab, H), device=dev, generator=gen) * 1.0).half()
        self.lm_head = (torch.randn((cfg.vocab, H), device=dev, generator=gen) * (1.0 / H ** 0.5)).half()  # fp16, cuBLAS (:432)

        # ---- paged KV cache: ctx tokens present, the step decodes token ctx (length ctx+1) ---------------------
        self.blocks_per_seq = (ctx + 1 + 63) // 64
        self.size_per_token = self.Hkv * D * self.kv_bits // 8
        code_bytes = 64 * self.size_per_token
        self.page_bytes = code_bytes + self.Hkv * 64 * 4  # cache_engine.py:62-66
        n_pages = batch * self.blocks_per_seq
        self.kpools, self.vpools, tables = [], [], []
        blk = torch.arange(n_pages, device=dev, dtype=torch.int64).view(batch, self.blocks_per_seq)
        for _ in range(self.L):
            pools = []
            for _kv in range(2):
                pool = torch.randint(0, 256, (n_pages, self.page_bytes), dtype=torch.uint8, device=dev, generator=gen)
                meta = pool[:, code_bytes:].view(torch.float16).view(n_pages, 2, self.Hkv, 64)
                meta[:, 0] = (torch.rand((n_pages, self.Hkv, 64), device=dev, generator=gen) * 0.09 + 0.01).half()
                zmax = 15.0 if self.kv_bits == 4 else 255.0
                meta[:, 1] = (torch.rand((n_pages, self.Hkv, 64), device=dev, generator=gen) * zmax).half()
                pools.append(pool)
            self.kpools.append(pools[0]); self.vpools.append(pools[1])
            # [B, 2, blocks] absolute addresses (model_runner.py:506-530)
            tables.append(torch.stack([pools[0].data_ptr() + blk * self.page_bytes, pools[1].data_ptr() + blk * self.page_bytes], dim=1))
        self.block_tables = torch.stack(tables, dim=0).contiguous()  # [L, B, 2, blocks]
        self.context_lens = torch.full((batch,), ctx + 1, dtype=torch.int32, device=dev)
        self.max_seq_len = ctx + 1

        # ---- persistent ActivationBuffer (input_metadata.py:71-109; aliasing kept) ------------------------------
        # (H is in the max for tensor parallelism: at TP = 8 a 72B model's sharded qkv / gate_up rows are narrower than the full hidden row of out_buf)
        self.act_buffer = torch.empty(M * max(self.q_size + 2 * self.kv_size, 2 * self.Iloc, H), dtype=torch.half, device=dev)
        self.qkv_buf = self.act_buffer[: M * (self.q_size + 2 * self.kv_size)].view(M, -1)
        self.out_buf = self.act_buffer[: M * H].view(M, H)
        self.gate_up_buf = self.act_buffer[: M * 2 * self.Iloc].view(M, -1)
        self.q_act = torch.empty(M * max(H, self.Iloc), dtype=torch.int8, device=dev)
        self.q_hidden = self.q_act[: M * H].view(M, H)
        self.q_attn = self.q_act[: M * self.q_size].view(M, self.q_size)
        self.q_mlp = self.q_act[: M * self.Iloc].view(M, self.Iloc)
        self.q_scale = torch.empty(M, dtype=torch.half, device=dev)
        self.q_sum = torch.empty(M, dtype=torch.half, device=dev)
        self.q_amax = torch.empty(M, dtype=torch.float32, device=dev)  # TP parity mode: per-token amax, max-all-reduced
        self.mlp_act = torch.empty((M, self.Iloc), dtype=torch.half, device=dev)  # reference: fresh torch.empty per call (activation.py:26)
        self.peer = None
        if self.tp_peer:
            self.peer = _ext.PeerContext(M, H, dev, process_group if process_group is not None else torch.distributed.group.WORLD)
        self.tokens_in = torch.zeros(M, dtype=torch.int64, device=dev)
        self.tokens_out = torch.zeros(M, dtype=torch.int64, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0

    # ---------------------------------------------------------------------------------------------------------
    def _norm_quant(self, x, gamma):
        if self.act_sum:  # layernorm.py:88
            self.ops.layernorm_ops.rms_norm_general_fuse_sum(self.q_hidden, x, gamma, self.q_sum, self.q_scale, self.cfg.eps, True)
        else:             # layernorm.py:72
            self.ops.layernorm_ops.rms_norm_general(self.q_hidden, x, gamma, self.q_scale, self.cfg.eps, True)

    def _quant(self, out_q, x):
        """Per-token quantisation of the input of a ROW-parallel GEMM (o_proj, down_proj)."""
        if self.tp_size > 1 and self.tp_exact:
            # same per-token scale on all ranks (global amax), local-K-shard activation sum
            _ext.row_absmax(self.q_amax, x)
            if not self.no_comm:
                torch.distributed.all_reduce(self.q_amax, op=torch.distributed.ReduceOp.MAX, group=self.pg)
            _ext.invoke_quant_given_amax(out_q, x, self.q_amax, self.q_sum if self.act_sum else None, self.q_scale)
        elif self.act_sum:  # llama_w4a8_unpad.py:177-183
            self.ops.fused_kernels.invoke_quant_fuse_sum(out_q, x, self.q_sum, self.q_scale)
        else:
            self.ops.fused_kernels.invoke_quant(out_q, x, self.q_scale)

    def _allreduce(self, t):
        if self.tp_size > 1 and not self.no_comm:
            torch.distributed.all_reduce(t, group=self.pg)

    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        """One decode step for `batch` sequences; returns the greedy next tokens [batch] (device)."""
        return self._forward_fused(tokens) if self.fused else self._forward_reference(tokens)

    def _attention(self, li):
        cfg, D = self.cfg, self.cfg.head_dim
        q, k, v = self.qkv_buf.split([self.q_size, self.kv_size, self.kv_size], dim=-1)  # :245-252
        q = q.reshape(q.size(0), self.Hq, D)
        k = k.reshape(k.size(0), self.Hkv, D)
        v = v.reshape(v.size(0), self.Hkv, D)
        attn = self.ops.fused_attention.single_query_attention(
            q, k, v, self.block_tables[li], self.context_lens, None, min(8192, cfg.max_pos), 64, self.size_per_token,
            self.max_seq_len, D, cfg.rope_theta, True, self.kv_bits == 4, True)  # :265-281
        return attn.reshape(q.size(0), -1)

    def _attention_quant(self, li, qsum) -> None:
        cfg, D = self.cfg, self.cfg.head_dim
        q, k, v = self.qkv_buf.split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        q = q.reshape(q.size(0), self.Hq, D)
        k = k.reshape(k.size(0), self.Hkv, D)
        v = v.reshape(v.size(0), self.Hkv, D)
        _ext.single_query_attention_quant(q, k, v, self.block_tables[li], self.context_lens, min(8192, cfg.max_pos), 64, self.size_per_token,
                                                 self.max_seq_len, D, cfg.rope_theta, self.kv_bits == 4, True, self.q_attn, self.q_scale, qsum)

    def _forward_reference(self, tokens: torch.Tensor, return_logits: bool = False) -> torch.Tensor:
        """Exactly the reference's op sequence (LlamaDecoderLayer.forward, llama_w4a8_unpad.py:330-361)."""
        cfg = self.cfg
        n = 0
        hidden = self.embed[tokens]  # LlamaModel.forward (:401-404)
        for li, ly in enumerate(self.layers):
            residual = hidden
            self._norm_quant(hidden, ly["ln1"])
            ly["qkv"](self.q_hidden, self.q_scale, self.q_sum, self.qkv_buf)
            attn = self._attention(li)
            self._quant(self.q_attn, attn)
            ly["o"](self.q_attn, self.q_scale, self.q_sum, self.out_buf)
            self._allreduce(self.out_buf)
            hidden = residual + self.out_buf  # :348
            residual = hidden
            self._norm_quant(hidden, ly["ln2"])
            ly["gate_up"](self.q_hidden, self.q_scale, self.q_sum, self.gate_up_buf)
            self.ops.activation_ops.silu_and_mul(self.mlp_act, self.gate_up_buf)  # activation.py:24-29
            self._quant(self.q_mlp, self.mlp_act)
            ly["down"](self.q_mlp, self.q_scale, self.q_sum, self.out_buf)
            self._allreduce(self.out_buf)
            hidden = residual + self.out_buf  # :360
            n += 10
        out = torch.empty_like(hidden)
        self.ops.layernorm_ops.rms_norm(out, hidden, self.norm_w, cfg.eps, False)  # final norm (:408)
        logits = torch.nn.functional.linear(out, self.lm_head)           # fp16 lm_head (:474-476)
        self.launches_per_step = n + 1
        return logits if return_logits else torch.argmax(logits, dim=-1)

    def _forward_fused(self, tokens: torch.Tensor, return_logits: bool = False) -> torch.Tensor:
        """Same arithmetic, three launches fewer per layer: the two torch residual adds are folded into the following norm
        (`add_rms_norm_general`) and silu_and_mul into the following per-token quant (`silu_and_mul_quant`)."""
        cfg = self.cfg
        qsum = self.q_sum if self.act_sum else None
        n = 0
        hidden = self.embed[tokens]
        nxt = torch.empty_like(hidden)
        self._norm_quant(hidden, self.layers[0]["ln1"])
        n += 1
        for li, ly in enumerate(self.layers):
            exact = self.tp_size > 1 and self.tp_exact
            ly["qkv"](self.q_hidden, self.q_scale, self.q_sum, self.qkv_buf)
            if exact or not self.fuse_attn_quant:
                self._quant(self.q_attn, self._attention(li))
            else:
                self._attention_quant(li, qsum)
            if self.tp_peer:
                ly["o"](self.q_attn, self.q_scale, self.q_sum, self.peer.partial[0])
                _ext.add_rms_norm_general_peer(self.q_hidden, nxt, hidden, self.peer, 0, ly["ln2"], qsum, self.q_scale, cfg.eps)
            else:
                ly["o"](self.q_attn, self.q_scale, self.q_sum, self.out_buf)
                self._allreduce(self.out_buf)
                _ext.add_rms_norm_general(self.q_hidden, nxt, hidden, self.out_buf, ly["ln2"], qsum, self.q_scale, cfg.eps)
            hidden, nxt = nxt, hidden
            ly["gate_up"](self.q_hidden, self.q_scale, self.q_sum, self.gate_up_buf)
            if exact:
                activation_ops.silu_and_mul(self.mlp_act, self.gate_up_buf)
                self._quant(self.q_mlp, self.mlp_act)
            else:
                _ext.silu_and_mul_quant(self.q_mlp, self.gate_up_buf, qsum, self.q_scale)
            if self.tp_peer:
                ly["down"](self.q_mlp, self.q_scale, self.q_sum, self.peer.partial[1])
            else:
                ly["down"](self.q_mlp, self.q_scale, self.q_sum, self.out_buf)
                self._allreduce(self.out_buf)
            n += 11 if exact else 7
            if self.tp_peer:
                # the last layer has no following norm: the same kernel delivers hidden + sum(partials) (its quantised output is not used)
                gam = self.layers[li + 1]["ln1"] if li + 1 < len(self.layers) else ly["ln1"]
                _ext.add_rms_norm_general_peer(self.q_hidden, nxt, hidden, self.peer, 1, gam, qsum, self.q_scale, cfg.eps)
                hidden, nxt = nxt, hidden
                n += 1
            elif li + 1 < len(self.layers):
                _ext.add_rms_norm_general(self.q_hidden, nxt, hidden, self.out_buf, self.layers[li + 1]["ln1"], qsum, self.q_scale, cfg.eps)
                hidden, nxt = nxt, hidden
                n += 1
            else:
                hidden = hidden + self.out_buf
        out = torch.empty_like(hidden)
        layernorm_ops.rms_norm(out, hidden, self.norm_w, cfg.eps, False)
        logits = torch.nn.functional.linear(out, self.lm_head)
        self.launches_per_step = n + 2
        return logits if return_logits else _ext.argmax_rows(logits)  # one launch instead of torch's two-pass reduction

    # ---------------------------------------------------------------------------------------------------------
    def load_shard_of(self, full: "DecodeRunner") -> None:
        """Make this tensor-parallel runner (rank tp_rank of tp_size) the shard of `full` (a tp_size = 1 runner of the same model /
        precision / batch / ctx living on the same or another device): weights by qserve_b200.tp (column parallel: the q | k | v and
        gate | up parts are sliced separately; row parallel: K slices of every band), KV pages by kv head, everything else replicated."""
        from qserve_b200 import tp

        r, n, dev = self.tp_rank, self.tp_size, self.dev
        assert full.tp_size == 1 and full.cfg == self.cfg and full.precision == self.precision and full.batch == self.batch and full.ctx == self.ctx
        assert self.cfg.kv_heads % n == 0, "KV-head replication (tp_size > kv_heads) is not needed by the benchmark models"
        D = self.cfg.head_dim

        def col_parts(lin, sizes):  # column-parallel shard of a fused layer: slice every part, then concatenate
            out, o = {}, 0
            attrs = ("weight", "wscale") if lin.mode == "w8" else ("qweight", "s1") + (("s1z",) if lin.mode == "chn" else ("s2_scales", "s2_zeros"))
            for a in attrs:
                t, pieces, o = getattr(lin, a), [], 0
                for sz in sizes:
                    if a in ("s2_scales", "s2_zeros"):
                        pieces.append(tp.shard_level2_columns(t[:, o:o + sz], r, n))
                    elif a in ("qweight", "weight"):
                        pieces.append(tp.shard_columns(t[o:o + sz], r, n))
                    else:
                        pieces.append(tp.shard_vector(t[o:o + sz], r, n))
                    o += sz
                out[a] = torch.cat(pieces, dim=1 if a in ("s2_scales", "s2_zeros") else 0).contiguous().to(dev)
            return out

        def row_parts(lin):
            out = {}
            if lin.mode == "w8":
                k = lin.K // n
                out["weight"] = lin.weight[:, r * k:(r + 1) * k].contiguous().to(dev)
                out["wscale"] = lin.wscale.to(dev)
            else:
                out["qweight"] = tp.shard_rows(lin.qweight, r, n).to(dev)
                out["s1"] = lin.s1.to(dev)
                if lin.mode == "chn":
                    out["s1z"] = lin.s1z.to(dev)
                else:
                    out["s2_scales"] = tp.shard_level2_rows(lin.s2_scales, r, n).to(dev)
                    out["s2_zeros"] = tp.shard_level2_rows(lin.s2_zeros, r, n).to(dev)
            return out

        def put(lin, parts):
            for a, t in parts.items():
                dst = getattr(lin, a)
                assert tuple(dst.shape) == tuple(t.shape), (a, dst.shape, t.shape)
                dst.copy_(t)

        cfg = self.cfg
        for mine, theirs in zip(self.layers, full.layers):
            put(mine["qkv"], col_parts(theirs["qkv"], (cfg.heads * D, cfg.kv_heads * D, cfg.kv_heads * D)))
            put(mine["gate_up"], col_parts(theirs["gate_up"], (cfg.intermediate, cfg.intermediate)))
            put(mine["o"], row_parts(theirs["o"]))
            put(mine["down"], row_parts(theirs["down"]))
            mine["ln1"].copy_(theirs["ln1"]); mine["ln2"].copy_(theirs["ln2"])
        self.norm_w.copy_(full.norm_w); self.embed.copy_(full.embed); self.lm_head.copy_(full.lm_head)
        # KV pages: [Hkv][64][D*bits/8] codes, then scales [Hkv][64], then zeros [Hkv][64]: this rank's kv heads
        Hf, Hl = full.Hkv, self.Hkv
        cb_f, cb_l = 64 * full.size_per_token, 64 * self.size_per_token
        for mine_p, full_p in zip(self.kpools + self.vpools, full.kpools + full.vpools):
            fp = full_p.to(dev)
            mine_p[:, :cb_l] = fp[:, :cb_f].reshape(-1, Hf, cb_f // Hf)[:, r * Hl:(r + 1) * Hl].reshape(-1, cb_l)
            meta_f = fp[:, cb_f:].reshape(-1, 2, Hf, 128)
            mine_p[:, cb_l:] = meta_f[:, :, r * Hl:(r + 1) * Hl].reshape(-1, 2 * Hl * 128)
        self.context_lens.copy_(full.context_lens)

    # ---------------------------------------------------------------------------------------------------------
    def capture(self, warmup: int = 2) -> None:
        """Warm up eagerly (allocates workspaces, sets kernel attributes) and capture the whole step in a CUDA graph."""
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                self.tokens_out.copy_(self.forward(self.tokens_in))
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g):
            self.tokens_out.copy_(self.forward(self.tokens_in))
        self.graph = g

    def step(self) -> None:
        """Replay the captured step: tokens_in -> tokens_out (both device resident)."""
        self.graph.replay()

    def weight_bytes_per_step(self) -> int:
        return sum(l.weight_bytes() for ly in self.layers for l in (ly["qkv"], ly["o"], ly["gate_up"], ly["down"]))

    def kv_bytes_per_step(self) -> int:
        """Algorithmic KV traffic: codes + scale/zero of K and V for ctx tokens, + q in / o out."""
        B, D = self.batch, self.cfg.head_dim
        per_layer = B * self.Hkv * self.ctx * D * self.kv_bits // 8 * 2 + B * self.Hkv * self.ctx * 8 + 4 * B * self.Hq * D
        return per_layer * self.L
