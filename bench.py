#!/usr/bin/env python
"""bench.py -- tokens/s of the Llama-3-8B W4A8KV4 decode step (batch 64, ctx 1024) on B200 in the driver's JSON contract,
with the per-kernel roofline, the reference's own GPU kernels on the same box, the reference's unmodified model code over
this backend, the tensor-parallel 72B record and the CPU baseline.

  python bench.py --gpus N --steps K --warmup W                 our sm_100a kernels behind the qserve_backend API
  python bench.py --impl reference --gpus N --steps K ...       the reference's CPU path (torch-CPU dequant-then-matmul restatement,
                                                                oracle/cpu_baseline.py) on the host cores; each step = a bounded sample
  python bench.py --impl reference-gpu --steps K ...            the reference's UNMODIFIED CUDA kernels (oracle/_ref, legacy mma.sync
                                                                compiled for sm_100a) driven through the same op sequence / buffers

A "step" is one decode step of the whole model for `batch` sequences (one new token each).  N > 1: the headline is N
data-parallel replicas of the 8B model (it fits one GPU: no collective on the data path), and the `tp` record of the same line
is Qwen1.5-72B (BASELINE config 5, all 80 layers) tensor-parallel over the N GPUs with NCCL all-reduces.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "tokens/s Llama-3-8B W4A8KV4 decode b64"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--precision", default="w4a8kv4")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=None, help="debug only: truncating the model invalidates the number")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-GPU-kernel block (oracle/_ref)")
    ap.add_argument("--no-refmodel", action="store_true", help="skip the block that runs the reference's unmodified model code over this backend")
    ap.add_argument("--no-tp", action="store_true", help="skip the Qwen1.5-72B tensor-parallel record")
    ap.add_argument("--tp-model", default="qwen1.5-72b")
    ap.add_argument("--tp-layers", type=int, default=None, help="debug only")
    ap.add_argument("--tp-only", action="store_true", help="print only the tensor-parallel record (tuning runs)")
    ap.add_argument("--tp-allreduce", default="peer", choices=["peer", "nccl"], help="TP record: all-reduce fused into add+norm+quant over peer memory, or NCCL")
    ap.add_argument("--tp-timeout", type=float, default=420.0, help="N > 1: seconds after which the tensor-parallel side record is abandoned (the headline line is still printed)")
    ap.add_argument("--tp-exact", action="store_true", help="TP record with the bit-exact parity rule (global per-token amax) instead of the throughput mode")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-fused", action="store_true", help="use exactly the reference op sequence (no fused add+norm / silu+quant extensions)")
    ap.add_argument("--kernel-reps", type=int, default=5)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
# shared by all arms: the workload description (identical `config` on every arm) and the algorithmic bytes
# ------------------------------------------------------------------------------------------------------------
def model_cfg(name):
    from qserve_b200.modelcfg import MODELS  # pure Python: does not load the CUDA library

    return MODELS[name]


def weight_bytes(cfg, precision, tp=1):
    D, H, I = cfg.head_dim, cfg.hidden, cfg.intermediate
    per = 1.0 if precision.startswith("w8a8") else 0.5
    q, kv = cfg.heads * D // tp, max(1, cfg.kv_heads // tp) * D
    elems = (q + 2 * kv) * H + H * q + 2 * (I // tp) * H + H * (I // tp)
    return int(elems * per)


def kv_bytes_per_layer(cfg, precision, batch, ctx, tp=1):
    """codes + scale/zero of K and V for ctx tokens + q in / o out (unique bytes)."""
    bits = 4 if "kv4" in precision else 8
    D, Hkv, Hq = cfg.head_dim, max(1, cfg.kv_heads // tp), cfg.heads // tp
    return batch * Hkv * ctx * D * bits // 8 * 2 + batch * Hkv * ctx * 8 + 4 * batch * Hq * D


def step_bytes(cfg, precision, batch, ctx, layers=None, tp=1):
    L = layers if layers is not None else cfg.layers
    return L * (weight_bytes(cfg, precision, tp) + kv_bytes_per_layer(cfg, precision, batch, ctx, tp)) + 2 * cfg.vocab * cfg.hidden


def make_config(args, world):
    cfg = model_cfg(args.model)
    L = cfg.layers if args.layers is None else args.layers
    sb = step_bytes(cfg, args.precision, args.batch, args.ctx, L)
    return {"workload": f"{cfg.name} {args.precision} decode batch={args.batch} ctx={args.ctx} ({L} layers + final norm + fp16 lm_head + greedy sampling)",
            "precision": args.precision, "batch": args.batch, "ctx": args.ctx, "layers": L, "parallelism": f"dp{world}",
            "l2": f"per-step working set {sb / 1e9:.2f} GB (weights + KV pages + lm_head) >> 126 MB L2: inputs larger than L2, no flush needed"}


def metric_name(args):
    cfg = model_cfg(args.model)
    return METRIC if (args.model, args.precision, args.batch) == ("llama-3-8b", "w4a8kv4", 64) else f"tokens/s {cfg.name} {args.precision} decode b{args.batch}"


# ------------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md: clocks line)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def ncu_traffic(kernel: str):
    """dram bytes per launch of `kernel` from the newest committed ncu capture summary (profiles/rNN_ncu_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return None
    d = json.load(open(files[-1])).get(kernel)
    return None if not d else int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
# arm: --impl reference  (CPU port of the reference path on the host cores; every step = one bounded sample)
# ------------------------------------------------------------------------------------------------------------
def reference_arm(args, rank: int):
    """The reference's CPU path on the host cores (kind = "port": the reference has no CPU code of its own).
    A step is a BOUNDED SAMPLE of the workload: one of the model's decoder layers (4 W4A8 GEMMs at M = batch: unpack + dequant +
    fp32 matmul; KV4 attention: dequant + fp32 SDPA) on inputs prepared once; the fp16 lm_head is timed once before the steps.
    tokens/s = batch / (layers x mean layer time + lm_head time).  Exactly --steps steps are timed after --warmup warm-up steps."""
    if rank != 0:
        return
    from oracle import cpu_baseline as cb  # the ONLY place besides tests/ and smoke() that executes oracle/ code

    cfg = model_cfg(args.model)
    threads = os.cpu_count() or 1
    sample = cb.LayerSample(cfg.hidden, cfg.intermediate, cfg.heads, cfg.kv_heads, args.batch, args.ctx, threads)
    t_lm = cb.lm_head_sample(cfg.hidden, cfg.vocab, args.batch, threads)
    for _ in range(args.warmup):
        sample.run()
    t0 = time.perf_counter()
    runs = [sample.run() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    layer_s = sum(r["layer_s"] for r in runs) / len(runs)
    mean = lambda k: sum(r[k] for r in runs) / len(runs)
    L = cfg.layers if args.layers is None else args.layers
    v = args.batch / (L * layer_s + t_lm)
    desc = (f"{args.steps} timed step(s) after {args.warmup} warm-up; each step = 1 of {L} decoder layers of the workload (4 W4A8 GEMMs at M={args.batch}: "
            f"unpack+dequant+fp32 matmul; KV4 attention B={args.batch} ctx={args.ctx}: dequant+fp32 SDPA) on inputs prepared once; fp16 lm_head timed once "
            f"({t_lm:.2f} s); tokens/s = batch / ({L} x {layer_s:.2f} s + lm_head)")
    line = {
        "impl": "reference", "metric": metric_name(args), "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s8", "data": "synthetic",
        "config": make_config(args, args.gpus),
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": desc, "layer_s": layer_s, "lm_head_s": t_lm,
                         "dequant_s": mean("dequant_s"), "matmul_s": mean("matmul_s"), "attention_s": mean("attention_s"),
                         "full_step_s_extrapolated": L * layer_s + t_lm},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# per-kernel timing (ours and the reference kernels go through the SAME function: only the OpSet differs)
# ------------------------------------------------------------------------------------------------------------
def time_kernel(fn, reps: int, stream, graph: bool = True):
    """Seconds per launch of the launches `fn()` issues (it returns their number).  graph=True: the launches are captured ONCE into a
    CUDA graph and the graph is replayed -- the host (ctypes marshalling, ~5-8 us per call, more than the short kernels themselves) is out of
    the timed region and the number is the device-side cost of a launch in a dependent chain, as in the captured decode step.
    graph=False (the reference extensions: their GEMMs launch on the legacy default stream and cannot be captured): eager back-to-back."""
    import torch

    fn()  # warm (attributes, workspaces)
    torch.cuda.synchronize()
    g = None
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            n_per = fn()
        run = g.replay
        run()
    else:
        n_per = None
        run = fn
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    n = 0
    for _ in range(reps):
        r = run()
        n += n_per if graph else r
    e1.record(stream)
    torch.cuda.synchronize()
    del g
    return e0.elapsed_time(e1) * 1e-3 / n  # seconds per launch


def kernel_table(run, reps, hbm_gbs, prefill: bool = True, fused_ops: bool = True, graph: bool = True):
    """Per-kernel launch duration and achieved HBM bandwidth, timed live with CUDA events on the launching stream (graph replay of the
    launch loop: see time_kernel).  Each timed
    loop walks the weights / KV pages of ALL layers (3.5 GB / 2.4 GB >> 126 MB L2), so every launch streams from HBM.
    `run.ops` decides whose kernels run (this library or the reference's)."""
    import torch

    stream = torch.cuda.current_stream()
    M, cfg, ops, L = run.batch, run.cfg, run.ops, run.L
    H, D = cfg.hidden, cfg.head_dim
    out = {}

    def gemm_bytes(lin):  # M*K + N*K/2 + 2*M*N + scales
        return M * lin.K + lin.weight_bytes() + 2 * M * lin.N + 4 * lin.N + 4 * M

    for name, xq, buf in (("qkv", run.q_hidden, run.qkv_buf), ("o", run.q_attn, run.out_buf), ("gate_up", run.q_hidden, run.gate_up_buf),
                          ("down", run.q_mlp, run.out_buf)):
        def fn(name=name, xq=xq, buf=buf):
            for ly in run.layers:
                ly[name](xq, run.q_scale, run.q_sum, buf)
            return len(run.layers)
        t = time_kernel(fn, reps, stream, graph)
        lin = run.layers[0][name]
        b = gemm_bytes(lin)
        out[f"gemm_{name}"] = {"M": M, "N": lin.N, "K": lin.K, "us": t * 1e6, "bytes": b, "GBps": b / t / 1e9, "frac": b / t / 1e9 / hbm_gbs,
                               "int8_TOPS": 2.0 * M * lin.N * lin.K / t / 1e12, "launches_per_step": L}

    q, k, v = run.qkv_buf.split([run.q_size, run.kv_size, run.kv_size], dim=-1)
    q, k, v = q.reshape(M, run.Hq, D), k.reshape(M, run.Hkv, D), v.reshape(M, run.Hkv, D)

    def attn():
        for li in range(L):
            ops.fused_attention.single_query_attention(q, k, v, run.block_tables[li], run.context_lens, None, min(8192, cfg.max_pos), 64, run.size_per_token,
                                                       run.max_seq_len, D, cfg.rope_theta, True, run.kv_bits == 4, True)
        return L
    t = time_kernel(attn, reps, stream, graph)
    b = run.kv_bytes_per_step() // L
    out["attention"] = {"B": M, "Hq": run.Hq, "Hkv": run.Hkv, "ctx": run.ctx, "us": t * 1e6, "bytes": b, "GBps": b / t / 1e9, "frac": b / t / 1e9 / hbm_gbs,
                        "launches_per_step": L}

    # the small ops of the reference sequence (latency-bound at M = 64; bytes = what they read + write)
    x = torch.randn((M, H), device=run.dev).half()
    gu = run.gate_up_buf
    attn_out = torch.randn((M, run.q_size), device=run.dev).half()
    nrep = 64

    def small(name, fn, nbytes, per_step):
        def loop():
            for _ in range(nrep):
                fn()
            return nrep
        t = time_kernel(loop, max(1, reps // 2), stream, graph)
        out[name] = {"us": t * 1e6, "bytes": nbytes, "GBps": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / hbm_gbs, "launches_per_step": per_step}

    gam = run.layers[0]["ln1"]
    if run.act_sum:
        small("norm_quant", lambda: ops.layernorm_ops.rms_norm_general_fuse_sum(run.q_hidden, x, gam, run.q_sum, run.q_scale, cfg.eps, True), 3 * M * H, 2 * L)
        small("quant_attn_out", lambda: ops.fused_kernels.invoke_quant_fuse_sum(run.q_attn, attn_out, run.q_sum, run.q_scale), 3 * M * run.q_size, L)
        small("quant_mlp", lambda: ops.fused_kernels.invoke_quant_fuse_sum(run.q_mlp, run.mlp_act, run.q_sum, run.q_scale), 3 * M * run.Iloc, L)
    else:
        small("norm_quant", lambda: ops.layernorm_ops.rms_norm_general(run.q_hidden, x, gam, run.q_scale, cfg.eps, True), 3 * M * H, 2 * L)
        small("quant_attn_out", lambda: ops.fused_kernels.invoke_quant(run.q_attn, attn_out, run.q_scale), 3 * M * run.q_size, L)
        small("quant_mlp", lambda: ops.fused_kernels.invoke_quant(run.q_mlp, run.mlp_act, run.q_scale), 3 * M * run.Iloc, L)
    small("silu_and_mul", lambda: ops.activation_ops.silu_and_mul(run.mlp_act, gu), 6 * M * run.Iloc, L)
    if fused_ops:
        from qserve_b200 import backend as ext
        qsum = run.q_sum if run.act_sum else None
        nxt = torch.empty_like(x)
        small("add_norm_quant(fused)", lambda: ext.add_rms_norm_general(run.q_hidden, nxt, x, run.out_buf, gam, qsum, run.q_scale, cfg.eps), 7 * M * H, 2 * L)
        small("silu_mul_quant(fused)", lambda: ext.silu_and_mul_quant(run.q_mlp, gu, qsum, run.q_scale), 5 * M * run.Iloc, L)

        def attn_q():
            for li in range(L):
                ext.single_query_attention_quant(q, k, v, run.block_tables[li], run.context_lens, min(8192, cfg.max_pos), 64, run.size_per_token, run.max_seq_len,
                                                 D, cfg.rope_theta, run.kv_bits == 4, True, run.q_attn, run.q_scale, qsum)
            return L
        t = time_kernel(attn_q, reps, stream, graph)
        out["attention_quant(fused)"] = {"us": t * 1e6, "bytes": b, "GBps": b / t / 1e9, "frac": b / t / 1e9 / hbm_gbs, "launches_per_step": L}

    if prefill:
        # prefill-sized GEMM (tensor-pipe bound, SURVEY.md 8d "GEMM INT8-TC util"): 4096 tokens through gate_up_proj
        Mp = 4096
        lin = run.layers[0]["gate_up"]
        xq = torch.randint(-127, 128, (Mp, lin.K), dtype=torch.int8, device=run.dev)
        sc = torch.full((Mp,), 0.01, dtype=torch.half, device=run.dev)
        sm = torch.zeros((Mp,), dtype=torch.half, device=run.dev)
        big = torch.empty((Mp, lin.N), dtype=torch.half, device=run.dev)

        def pf():
            for ly in run.layers[:4]:
                ly["gate_up"](xq, sc, sm, big)
            return min(4, len(run.layers))
        t = time_kernel(pf, max(2, reps // 2), stream, graph)
        opsn = 2.0 * Mp * lin.N * lin.K
        out["gemm_prefill_gate_up"] = {"M": Mp, "N": lin.N, "K": lin.K, "us": t * 1e6, "int8_TOPS": opsn / t / 1e12,
                                       "frac_of_nominal_int8_dense": opsn / t / 1e12 / 4500.0, "frac_of_measured_umma_i8_peak": opsn / t / 1e12 / 4760.0,
                                       "bound": "tensor (INT8 dense: nominal 4.5 POP/s; measured tcgen05 kind::i8 issue peak 4.76 POP/s, tools/ubench/umma.cu)"}
        del big
        if fused_ops:  # (this repo's table only: the reference-gpu arm reuses kernel_table for the reference extensions)
            # prompt-phase attention: BASELINE config 3's prompt batch, 8 x ctx tokens, causal; flash-attn (the third-party
            # kernel the reference calls, a library) is timed beside it when the image has it
            from qserve_b200 import backend as _be
            cfg = run.cfg
            n_p, plen = 8, run.ctx
            qkv = torch.randn(n_p * plen, (cfg.heads + 2 * cfg.kv_heads) * 128, device=run.dev).half()
            q, k, v = qkv.split([cfg.heads * 128, cfg.kv_heads * 128, cfg.kv_heads * 128], dim=-1)
            q, k, v = q.reshape(-1, cfg.heads, 128), k.reshape(-1, cfg.kv_heads, 128), v.reshape(-1, cfg.kv_heads, 128)
            cu = torch.arange(0, (n_p + 1) * plen, plen, dtype=torch.int32, device=run.dev)
            flops = n_p * 4.0 * 128 * cfg.heads * plen * (plen + 1) / 2
            impls = {"qserve_b200": _be.flash_attn_varlen_func}
            try:
                from flash_attn import flash_attn_varlen_func as _fa
                impls["flash_attn"] = _fa
            except Exception:  # noqa: BLE001
                pass
            rec = {"prompts": n_p, "prompt_len": plen, "heads": cfg.heads, "kv_heads": cfg.kv_heads, "causal_flops": flops,
                   "bound": "tensor (fp16 dense: nominal 2.25 PFLOP/s)"}
            for name, fn in impls.items():
                def pa(fn=fn):
                    fn(q, k, v, cu, cu, plen, plen, dropout_p=0.0, causal=True)
                    return 1
                t = time_kernel(pa, max(2, reps // 2), stream, graph)
                rec[f"us_{name}" if name != "qserve_b200" else "us"] = t * 1e6
                rec[f"TFLOPS_{name}" if name != "qserve_b200" else "TFLOPS"] = flops / t / 1e12
            rec["frac_of_nominal_fp16_dense"] = rec["TFLOPS"] / 2250.0
            out["prefill_attention"] = rec
    return out


def roofline_from(kern, hbm_gbs, peak_src, fused: bool):
    """The dominant kernel FUNCTION of the step by measured time share: launches per step x isolated launch duration, with all
    launch shapes of one template instantiation aggregated (the four M = 64 GEMM shapes are ONE instantiation of gemm_kernel)."""
    gemm_keys = [k for k in ("gemm_qkv", "gemm_o", "gemm_gate_up", "gemm_down") if k in kern]
    attn_key = "attention_quant(fused)" if fused and "attention_quant(fused)" in kern else "attention"
    groups = {"gemm_kernel (W4A8/W8A8 tcgen05 GEMM, all four decode shapes: one template instantiation)": gemm_keys,
              "decode_attention_kernel (KV4/KV8 paged single-query attention)": [attn_key]}
    small = [k for k in (("add_norm_quant(fused)", "silu_mul_quant(fused)") if fused else ("norm_quant", "quant_attn_out", "quant_mlp", "silu_and_mul")) if k in kern]
    step_us = sum(kern[k]["us"] * kern[k]["launches_per_step"] for ks in groups.values() for k in ks) + sum(kern[k]["us"] * kern[k]["launches_per_step"] for k in small)
    best, share = None, {}
    for name, ks in groups.items():
        t = sum(kern[k]["us"] * kern[k]["launches_per_step"] for k in ks)
        share[name] = t / step_us
        if best is None or t > best[1]:
            best = (name, t, ks)
    name, t_us, ks = best
    launches = sum(kern[k]["launches_per_step"] for k in ks)
    bytes_step = sum(kern[k]["bytes"] * kern[k]["launches_per_step"] for k in ks)
    achieved = bytes_step / (t_us * 1e-6) / 1e9
    traffic = None
    parts = [ncu_traffic(k if k != "attention_quant(fused)" else "attention") for k in ks]
    if all(p is not None for p in parts):
        traffic = sum(p * kern[k]["launches_per_step"] for p, k in zip(parts, ks)) / launches
    return {"bound": "hbm", "kernel": name, "selection": "largest share of the step's kernel time (launches per step x measured launch duration)",
            "time_share_of_step_kernels": share[name], "time_shares": share,
            "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs, "peak_source": peak_src,
            "algorithmic_bytes": bytes_step / launches, "us_per_launch": t_us / launches, "launches_per_step": launches,
            "traffic": traffic, "traffic_unit": "dram bytes per launch, averaged over the launch shapes (ncu --set full)",
            "by_launch_shape": {k: {"us": kern[k]["us"], "frac": kern[k]["frac"], "bytes": kern[k]["bytes"]} for k in ks}}


# ------------------------------------------------------------------------------------------------------------
# arm: --impl reference-gpu  (the reference's own CUDA kernels on this B200: the "legacy mma.sync on B200" row)
# ------------------------------------------------------------------------------------------------------------
def reference_gpu_arm(args):
    import torch

    from oracle import refmods  # the reference itself (compiled unmodified by oracle/build_ref.py); this arm exists to time it
    from qserve_b200.decode import DecodeRunner, OpSet

    names = {"layernorm_ops": "layernorm_ops", "fused_kernels": "fused_kernels", "activation_ops": "activation_ops", "fused_attention": "fused_attention",
             "qgemm_chn": "qgemm_w4a8_per_chn", "qgemm_grp": "qgemm_w4a8_per_group", "qgemm_w8": "qgemm_w8a8"}
    mods = {k: refmods.load(v) for k, v in names.items()}
    missing = [names[k] for k, m in mods.items() if m is None]
    if missing:
        print(json.dumps({"impl": "reference-gpu", "unavailable": f"oracle/_ref lacks {missing} (oracle/build_ref.py needs /root/reference)"}), flush=True)
        return
    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = model_cfg(args.model)
    run = DecodeRunner(args.model, args.precision, args.batch, args.ctx, dev, layers=args.layers, fused=False, ops=OpSet(**mods))
    hbm_gbs, _ = peaks()
    stream = torch.cuda.current_stream()  # the reference GEMMs launch on the legacy default stream (gemm_cuda.cu:53): stay on it
    tok = torch.randint(0, cfg.vocab, (args.batch,), device=dev)
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            run.tokens_out.copy_(run.forward(tok))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            run.tokens_out.copy_(run.forward(tok))
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1) / args.steps
        kern = kernel_table(run, args.kernel_reps, hbm_gbs, prefill=True, fused_ops=False, graph=False)
    seq = ("gemm_qkv", "gemm_o", "gemm_gate_up", "gemm_down", "attention", "norm_quant", "quant_attn_out", "quant_mlp", "silu_and_mul")
    kern_ms = sum(kern[k]["us"] * kern[k]["launches_per_step"] for k in seq) * 1e-3
    line = {"impl": "reference-gpu", "metric": metric_name(args), "value": args.batch / (ms * 1e-3), "unit": "tokens/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s8", "data": "synthetic",
            "config": make_config(args, 1),
            "what": "the reference's UNMODIFIED CUDA kernels (kernels/csrc/**, legacy mma.sync / CUDA-core attention) compiled for sm_100a by oracle/build_ref.py, driven "
                    "through the reference op sequence of LlamaDecoderLayer.forward on the same buffers; eager launches (the reference GEMMs use the legacy "
                    "default stream and cannot be graph-captured), so `value` includes host launch overhead; `kernel_ms_per_step` is the GPU-time-only lower bound",
            "host_ms_per_step": 1000.0 * wall / args.steps,
            "kernel_ms_per_step": kern_ms, "tokens_per_s_kernel_time_only": args.batch / (kern_ms * 1e-3 + 0.25e-3),
            "kernels": kern, "gpu_launches": run.launches_per_step * args.steps}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# block: the reference's unmodified Python model code over this backend (eager / graph decode, config-3 prefill + decode)
# ------------------------------------------------------------------------------------------------------------
def refmodel_block(run, args, steps: int):
    import torch

    from qserve_b200 import refmodel
    from qserve_b200.decode import DecodeRunner

    if refmodel.locate_reference() is None:
        return {"unavailable": "reference Python package not found (baseline/_ref is installed by __graft_entry__.build() where /root/reference exists)"}
    out = {"what": "qserve.modeling.models.llama_w4a8_unpad.LlamaForCausalLM (the reference's own layer / model classes, unmodified, from baseline/_ref) with "
                   "qserve_backend = this repo; synthetic weights shared with the DecodeRunner"}
    stream = torch.cuda.current_stream()

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, 1000.0 * (time.perf_counter() - t0) / n

    def decode_numbers(ref, r, tag):
        tok = torch.randint(0, r.cfg.vocab, (r.batch,), device=r.dev)
        for _ in range(3):
            ref.decode_tokens(tok)
        ms, host_ms = timed(lambda: ref.decode_tokens(tok), steps)
        out[f"{tag}_eager"] = {"tokens_per_s": r.batch / (ms * 1e-3), "ms_per_step": ms, "host_ms_per_step": host_ms,
                               "note": "eager Python loop of the reference model: what a drop-in user of qserve_benchmark.py gets"}
        # the same unmodified forward captured in a CUDA graph (the ops are stream-ordered and allocation-free apart from torch's own)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ref.decode_tokens(tok)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ref.decode_tokens(tok)
        for _ in range(3):
            g.replay()
        ms, _ = timed(g.replay, steps)
        out[f"{tag}_graph"] = {"tokens_per_s": r.batch / (ms * 1e-3), "ms_per_step": ms}
        del g

    ref = refmodel.RefModel(run)
    decode_numbers(ref, run, "decode")
    del ref
    # BASELINE config 3: Llama-3-8B g128, a prompt step (in-place RoPE + KV4 quant/append + flash_attn_varlen_func inside the reference
    # layer code, llama_w4a8_unpad.py:203-242) followed by decode steps over the pages it wrote
    try:
        r3 = DecodeRunner(args.model, "w4a8kv4-g128", args.batch, args.ctx, run.dev, layers=args.layers, fused=False, seed=3)
        ref3 = refmodel.RefModel(r3)
        n_prompts, plen = 8, args.ctx
        lens = [plen] * n_prompts
        toks = torch.randint(0, r3.cfg.vocab, (n_prompts * plen,), device=r3.dev)
        meta = ref3.prefill_metadata(lens)
        per_impl = {}
        for impl in ("flash_attn", "qserve_b200"):  # the prompt attention behind the reference layer's flash_attn_varlen_func call
            ref3.use_prefill_attention(impl)
            for _ in range(2):
                ref3.prefill_logits(toks, lens, meta)
            per_impl[impl] = timed(lambda: ref3.prefill_logits(toks, lens, meta), 3)
        ms, host_ms = per_impl["qserve_b200"]
        out["config3_prefill"] = {"prompts": n_prompts, "prompt_len": plen, "tokens": n_prompts * plen, "ms_per_step": ms, "host_ms_per_step": host_ms,
                                  "tokens_per_s": n_prompts * plen / (ms * 1e-3), "int8_TOPS_gemm_only": 2.0 * n_prompts * plen * weight_bytes(r3.cfg, "w4a8", 1) * 2 * r3.L / (ms * 1e-3) / 1e12,
                                  "prompt_attention": "qserve_b200 tcgen05 kernel (qs_prefill_attention)",
                                  "ms_per_step_with_flash_attn": per_impl["flash_attn"][0],
                                  "tokens_per_s_with_flash_attn": n_prompts * plen / (per_impl["flash_attn"][0] * 1e-3)}
        decode_numbers(ref3, r3, "config3_decode")
        pf, dg = out["config3_prefill"], out["config3_decode_graph"]
        # in-flight batching round: one prompt step admits 8 sequences, then the batch decodes until they finish 512 tokens (qserve_benchmark protocol 1024 in / 512 out)
        out["config3_ifb_round"] = {"what": f"1 prompt step ({n_prompts} x {plen} tokens) + {args.batch // n_prompts} decode steps at batch {args.batch} per admitted group, "
                                            "steady state of 1024-in/512-out in-flight batching: generated tokens/s",
                                    "tokens_per_s": (n_prompts * 512) / ((pf["ms_per_step"] + 512 * dg["ms_per_step"] * n_prompts / args.batch) * 1e-3)}
        del ref3, r3
    except Exception as e:  # noqa: BLE001
        out["config3_error"] = repr(e)
    return out


# ------------------------------------------------------------------------------------------------------------
# block: BASELINE config 5 -- Qwen1.5-72B W4A8KV4, all 80 layers, tensor parallel over the N GPUs of the run (TP = 1 at N = 1)
# ------------------------------------------------------------------------------------------------------------
def tp_block(args, rank, world, dev, hbm_gbs):
    import torch
    import torch.distributed as dist

    from qserve_b200.decode import DecodeRunner

    cfg = model_cfg(args.tp_model)
    want_peer = args.tp_allreduce == "peer" and world > 1 and not args.tp_exact
    peer_err = None
    if want_peer:
        # symmetric memory needs CUDA VMM handle exchange between the ranks: agree on success before building the model on it
        ok = torch.ones(1, device=dev)
        try:
            from qserve_b200 import backend as _b
            probe = _b.PeerContext(8, 128, dev, dist.group.WORLD)
            del probe
        except Exception as e:  # noqa: BLE001
            ok.zero_()
            peer_err = repr(e)[:300]
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        want_peer = bool(ok.item() > 0)
    run = DecodeRunner(args.tp_model, "w4a8kv4", args.batch, args.ctx, dev, tp_rank=rank, tp_size=world, seed=rank, layers=args.tp_layers,
                       fused=True, tp_exact=args.tp_exact, tp_peer=want_peer)
    run.capture()
    step_fn = run.step
    stream = torch.cuda.current_stream()
    steps = max(5, min(args.steps, 20))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        step_fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step_fn()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    ar_us = None
    if world > 1:
        buf = torch.zeros((args.batch, cfg.hidden), dtype=torch.half, device=dev)
        for _ in range(5):
            dist.all_reduce(buf)
        barrier()
        e0.record(stream)
        for _ in range(200):
            dist.all_reduce(buf)
        e1.record(stream)
        barrier()
        ar_us = e0.elapsed_time(e1) * 1e3 / 200
        t = torch.tensor([ms, ar_us], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ar_us = t.tolist()
    L = run.L
    sb = step_bytes(cfg, "w4a8kv4", args.batch, args.ctx, L, world)
    rec = {"model": cfg.name, "precision": "w4a8kv4", "batch": args.batch, "ctx": args.ctx, "layers": L, "tp": world, "steps": steps,
           "quant_mode": "exact (global per-token amax, SURVEY 8e parity rule)" if args.tp_exact else "throughput (per-rank local amax)",
           "tokens_per_s": args.batch / (ms * 1e-3), "ms_per_step": ms, "per_rank_hbm_bytes_per_step": sb,
           "per_rank_hbm_frac": sb / (ms * 1e-3) / 1e9 / hbm_gbs,
           "allreduce": None if world == 1 else {"calls_per_step": 2 * L, "message_bytes": args.batch * cfg.hidden * 2, "us_per_call_isolated": ar_us,
                                                 "ms_per_step_isolated": 2 * L * ar_us * 1e-3,
                                                 "nvlink_bytes_algorithmic_per_step": 2 * L * args.batch * cfg.hidden * 2 * 2 * (world - 1) // world,
                                                 "us_per_call_isolated_is": "NCCL all-reduce of the same message, back to back, for reference",
                                                 "backend": ("fused into add+norm+quant: flag exchange + 128-bit peer loads over NVLink symmetric memory (qs_add_rms_norm_general_peer), no NCCL on the data path"
                                                             if run.tp_peer else "NCCL all-reduce inside the captured CUDA graph")},
           "peer_fallback_reason": peer_err,
           "note": "efficiency vs TP=1 = this tokens_per_s / (N x the tp=1 record of the --gpus 1 line); 36 GB of W4 weights fit one GPU, TP is for throughput"}
    run.graph = None
    del run
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return rec


# ------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if args.impl == "reference-gpu":
        if rank == 0:
            reference_gpu_arm(args)
        return

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: qserve_b200 has no CPU path (use --impl reference for the CPU baseline)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from qserve_b200 import backend
    from qserve_b200.decode import DecodeRunner

    backend.set_pdl(not args.no_pdl)
    cfg = model_cfg(args.model)
    if args.tp_only:
        with torch.no_grad():
            rec = tp_block(args, rank, world, dev, peaks()[0])
        if rank == 0:
            print(json.dumps({"tp": rec}), flush=True)
        sys.stdout.flush()
        if world > 1:
            dist.barrier()
            os._exit(0)
        return
    run = DecodeRunner(args.model, args.precision, args.batch, args.ctx, dev, seed=rank, layers=args.layers, fused=not args.no_fused)
    hbm_gbs, peak_src = peaks()

    # ---- pinned host buffers for the end-to-end leg ------------------------------------------------------
    tok_host = torch.randint(0, cfg.vocab, (args.batch,), dtype=torch.int64).pin_memory()
    out_host = torch.zeros(args.batch, dtype=torch.int64).pin_memory()
    run.tokens_in.copy_(tok_host)
    if args.no_graph:
        with torch.no_grad():
            for _ in range(2):
                run.tokens_out.copy_(run.forward(run.tokens_in))
        step = lambda: run.tokens_out.copy_(run.forward(run.tokens_in))
    else:
        run.capture()
        step = run.step
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step()
        if os.environ.get("QS_PROFILE_STEP"):  # ncu --profile-from-start off: exactly one eager step between profiler start / stop
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            run.tokens_out.copy_(run.forward(run.tokens_in))
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        # ---- device-resident timing: exactly K steps, CUDA events, max over ranks -------------------------
        barrier()
        with ClockSampler(local_rank) as clk:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            barrier()
        ms_dev = e0.elapsed_time(e1)
        # ---- end to end: host tokens in (pinned) -> H2D -> step -> D2H sampled tokens (pinned), every step ----
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(args.steps):
            run.tokens_in.copy_(tok_host, non_blocking=True)
            step()
            out_host.copy_(run.tokens_out, non_blocking=True)
            stream.synchronize()          # the engine consumes the sampled tokens on the host every step (llm_engine.py:595)
            tok_host.copy_(out_host)      # feed them back, as the generation loop does
        e3.record(stream)
        barrier()
        ms_e2e = e2.elapsed_time(e3)

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = t.tolist()
    tokens_per_step = args.batch * world
    value = tokens_per_step * args.steps / (ms_dev * 1e-3)
    e2e_value = tokens_per_step * args.steps / (ms_e2e * 1e-3)
    launches_per_step = run.launches_per_step

    kern, refm, cpu, ref_gpu = None, None, None, None
    if rank == 0:
        with torch.no_grad():
            kern = kernel_table(run, args.kernel_reps, hbm_gbs)
        if world == 1 and not args.no_refmodel:
            try:
                refm = refmodel_block(run, args, max(5, min(args.steps, 20)))
            except Exception as e:  # noqa: BLE001  (a broken side block must not lose the headline)
                refm = {"error": repr(e)}
    # free the 8B replica before the 72B record
    run.graph = None
    del run
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    fused = not args.no_fused
    clocks = clk.summary()

    def make_line(tp_rec):
        """The result line (rank 0).  Built from the data-parallel measurement alone, so that a failure of the tensor-parallel side record at
        N > 1 can still be reported next to a valid headline."""
        roof = roofline_from(kern, hbm_gbs, peak_src, fused)
        sb = step_bytes(cfg, args.precision, args.batch, args.ctx, args.layers)
        line = {
            "metric": metric_name(args), "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s8", "data": "synthetic", "config": make_config(args, world),
            "run": {"graph": not args.no_graph, "pdl": not args.no_pdl, "fused_small_ops": fused,
                    "timing": "CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": args.batch * 8 * world, "d2h_bytes_per_step": args.batch * 8 * world,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": roof,
            "kernels": kern,
            "step_hbm_frac": (sb / (ms_dev / args.steps * 1e-3)) / 1e9 / hbm_gbs,
        }
        if tp_rec is not None:
            line["tp"] = tp_rec
        return line

    tp_rec = None
    if not args.no_tp and (args.model, args.precision) == ("llama-3-8b", "w4a8kv4"):
        import threading

        tp_done = threading.Event()

        def bail(why):
            # N > 1 and the tensor-parallel record failed or hangs on some rank: the data-parallel headline is complete and valid -- print it with
            # the error and leave without touching the (possibly wedged) communicator again
            if rank == 0:
                print(json.dumps(make_line({"error": why})), flush=True)
            sys.stdout.flush()
            os._exit(0)

        def watchdog():
            if not tp_done.wait(args.tp_timeout):
                bail(f"tensor-parallel record did not finish within {args.tp_timeout} s")

        if world > 1:
            dist.barrier()  # rank 0 has just spent seconds on the kernel table: enter the collective phase together
            threading.Thread(target=watchdog, daemon=True).start()
        try:
            with torch.no_grad():
                tp_rec = tp_block(args, rank, world, dev, hbm_gbs)
        except Exception as e:  # noqa: BLE001
            tp_rec = {"error": repr(e)}
            if world > 1:
                if rank != 0:
                    time.sleep(5)  # give rank 0 a moment to fail on its own (it prints); otherwise its watchdog prints
                bail(repr(e)[:400])
        tp_done.set()

    def finish():
        # do not let a slow communicator teardown keep the launcher alive after the result line is out
        sys.stdout.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if world > 1:
        dist.barrier()
    if rank != 0:
        finish()
        return

    if world == 1 and not args.no_ref_gpu:
        # the reference's own CUDA kernels on this box, same shapes / harness, in a separate process (keeps this arm's process free of oracle/_ref)
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-gpu", "--steps", str(max(5, min(args.steps, 10))), "--warmup", "3",
                   "--model", args.model, "--precision", args.precision, "--batch", str(args.batch), "--ctx", str(args.ctx), "--kernel-reps", "3"]
            if args.layers is not None:
                cmd += ["--layers", str(args.layers)]
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
            lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
            ref_gpu = json.loads(lines[-1]) if lines else {"error": (res.stderr or res.stdout)[-400:]}
        except Exception as e:  # noqa: BLE001
            ref_gpu = {"error": repr(e)}
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline as cb

        threads = os.cpu_count() or 1
        sample = cb.LayerSample(cfg.hidden, cfg.intermediate, cfg.heads, cfg.kv_heads, args.batch, args.ctx, threads)
        t_lm = cb.lm_head_sample(cfg.hidden, cfg.vocab, args.batch, threads)
        sample.run()
        rs = [sample.run() for _ in range(3)]
        layer_s = sum(r["layer_s"] for r in rs) / len(rs)
        v = args.batch / (cfg.layers * layer_s + t_lm)
        cpu = {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port",
               "sample": f"3 timed runs (after 1 warm-up) of 1 of {cfg.layers} decoder layers (4 W4A8 GEMMs M={args.batch} + KV4 attention B={args.batch} ctx={args.ctx}) on "
                         f"torch-CPU dequant-then-matmul + fp16 lm_head once, extrapolated x{cfg.layers}; {layer_s:.2f} s/layer, lm_head {t_lm:.2f} s"}

    line = make_line(tp_rec)
    if refm is not None:
        line["refmodel"] = refm
    if ref_gpu is not None:
        keep = {k: ref_gpu[k] for k in ("value", "unit", "ms_per_step", "host_ms_per_step", "kernel_ms_per_step", "tokens_per_s_kernel_time_only", "what",
                                        "unavailable", "error") if k in ref_gpu}
        if "kernels" in ref_gpu:
            keep["kernel_us"] = {k: v["us"] for k, v in ref_gpu["kernels"].items()}
            keep["speedup_per_kernel"] = {k: ref_gpu["kernels"][k]["us"] / kern[k]["us"] for k in ref_gpu["kernels"] if k in kern}
            keep["speedup_step_vs_reference_eager"] = value / ref_gpu["value"]
            keep["speedup_step_vs_reference_kernel_time_only"] = value / ref_gpu["tokens_per_s_kernel_time_only"]
        line["ref_gpu"] = keep
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
