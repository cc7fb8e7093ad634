#!/usr/bin/env python
"""bench.py -- tokens/s of the Llama-3-8B W4A8KV4 decode step (batch 64, ctx 1024) on B200, with the per-kernel
roofline and the CPU baseline, in the driver's JSON contract.

  python bench.py --gpus N --steps K --warmup W            our sm_100a kernels behind the qserve_backend API
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (torch-CPU dequant-then-matmul
                                                           restatement, oracle/cpu_baseline.py) on the host cores

A "step" is one decode step of the whole model for `batch` sequences (one new token each).  N > 1 runs N
data-parallel replicas (the path does not need a collective for a model that fits one GPU; north_star reserves
tensor parallelism for 70B/72B: `--model qwen1.5-72b` switches to TP = N with an NCCL all-reduce).
"""
from __future__ import annotations

import argparse
import json
import os
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--precision", default="w4a8kv4")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=None, help="debug only: truncating the model invalidates the number")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-fused", action="store_true", help="use exactly the reference op sequence (no fused add+norm / silu+quant extensions)")
    ap.add_argument("--kernel-reps", type=int, default=5)
    return ap.parse_args()


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
def reference_arm(args, rank: int):
    """The reference's CPU path on the host cores (kind = "port": the reference has no CPU code of its own)."""
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    from qserve_b200.decode import MODELS  # config table only

    cfg = MODELS[args.model]
    threads = os.cpu_count() or 1
    vals = []
    t_all = time.perf_counter()
    n = args.warmup + args.steps
    last = None
    for i in range(n):
        r = cb.decode_tokens_per_s(cfg.hidden, cfg.intermediate, cfg.heads, cfg.kv_heads, cfg.layers, cfg.vocab, args.batch, args.ctx, threads)
        if i >= args.warmup:
            vals.append(r["tokens_per_s"])
        last = r
        if time.perf_counter() - t_all > 240 and len(vals) >= 1:  # keep the arm within a few minutes
            break
    v = sum(vals) / len(vals)
    sample = (f"{len(vals)} step(s); each step = 1 of {cfg.layers} decoder layers (4 W4A8 GEMMs at M={args.batch}: unpack+dequant+fp32 matmul, "
              f"KV4 attention B={args.batch} ctx={args.ctx}: dequant+fp32 SDPA) + fp16 lm_head, extrapolated x{cfg.layers}")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": 1000.0 * args.batch / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s8", "data": "synthetic",
        "config": {"workload": f"{cfg.name} {args.precision} decode batch={args.batch} ctx={args.ctx}", "precision": args.precision,
                   "batch": args.batch, "ctx": args.ctx, "parallelism": "host cores"},
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample,
                         "dequant_s": last["dequant_s"], "matmul_s": last["matmul_s"], "attention_s": last["attention_s"]},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
def time_kernel(fn, reps: int, stream):
    import torch

    fn()  # warm
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    n = 0
    for _ in range(reps):
        n += fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / n  # seconds per launch


def kernel_rooflines(run, reps, hbm_gbs):
    """Per-kernel achieved HBM bandwidth, timed live with CUDA events on the launching stream.  Each timed loop walks
    the weights / KV pages of ALL layers (3.5 GB / 2.4 GB >> 126 MB L2), so every launch streams from HBM."""
    import torch

    stream = torch.cuda.current_stream()
    M = run.batch
    out = {}

    def gemm_bytes(lin):  # M*K + N*K/2 + 2*M*N + scales
        w = lin.weight_bytes()
        return M * lin.K + w + 2 * M * lin.N + 4 * lin.N + 4 * M

    for name, xq, buf in (("qkv", run.q_hidden, run.qkv_buf), ("o", run.q_attn, run.out_buf), ("gate_up", run.q_hidden, run.gate_up_buf),
                          ("down", run.q_mlp, run.out_buf)):
        def fn(name=name, xq=xq, buf=buf):
            for ly in run.layers:
                ly[name](xq, run.q_scale, run.q_sum, buf)
            return len(run.layers)
        t = time_kernel(fn, reps, stream)
        lin = run.layers[0][name]
        b = gemm_bytes(lin)
        out[f"gemm_{name}"] = {"M": M, "N": lin.N, "K": lin.K, "us": t * 1e6, "bytes": b, "GBps": b / t / 1e9, "frac": b / t / 1e9 / hbm_gbs,
                               "int8_TOPS": 2.0 * M * lin.N * lin.K / t / 1e12}

    import qserve_backend.fused_attention as fa
    D = run.cfg.head_dim
    q, k, v = run.qkv_buf.split([run.q_size, run.kv_size, run.kv_size], dim=-1)
    q, k, v = q.reshape(M, run.Hq, D), k.reshape(M, run.Hkv, D), v.reshape(M, run.Hkv, D)

    def attn():
        for li in range(run.L):
            fa.single_query_attention(q, k, v, run.block_tables[li], run.context_lens, None, 8192, 64, run.size_per_token, run.max_seq_len, D,
                                      run.cfg.rope_theta, True, run.kv_bits == 4, True)
        return run.L
    t = time_kernel(attn, reps, stream)
    b = run.kv_bytes_per_step() // run.L
    out["attention"] = {"B": M, "Hq": run.Hq, "Hkv": run.Hkv, "ctx": run.ctx, "us": t * 1e6, "bytes": b, "GBps": b / t / 1e9, "frac": b / t / 1e9 / hbm_gbs}

    # prefill-sized GEMM (tensor-pipe bound, SURVEY.md 8d "GEMM INT8-TC util"): 4096 tokens through gate_up_proj
    Mp = 4096
    lin = run.layers[0]["gate_up"]
    xq = torch.randint(-127, 128, (Mp, lin.K), dtype=torch.int8, device=run.q_hidden.device)
    sc = torch.full((Mp,), 0.01, dtype=torch.half, device=xq.device)
    sm = torch.zeros((Mp,), dtype=torch.half, device=xq.device)
    big = torch.empty((Mp, lin.N), dtype=torch.half, device=xq.device)

    def prefill():
        for ly in run.layers[:4]:
            ly["gate_up"](xq, sc, sm, big)
        return 4
    t = time_kernel(prefill, max(2, reps // 8), stream)
    ops = 2.0 * Mp * lin.N * lin.K
    out["gemm_prefill_gate_up"] = {"M": Mp, "N": lin.N, "K": lin.K, "us": t * 1e6, "int8_TOPS": ops / t / 1e12,
                                   "frac_of_nominal_int8_dense": ops / t / 1e12 / 4500.0, "bound": "tensor (INT8 dense nominal 4.5 POP/s)"}
    del big
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
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
    from qserve_b200.decode import MODELS, DecodeRunner

    backend.set_pdl(not args.no_pdl)
    cfg = MODELS[args.model]
    tp = world if args.model in ("qwen1.5-72b",) and world > 1 else 1
    run = DecodeRunner(args.model, args.precision, args.batch, args.ctx, dev, tp_rank=rank if tp > 1 else 0, tp_size=tp, seed=rank, layers=args.layers, fused=not args.no_fused)
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
    replicas = world if tp == 1 else 1
    tokens_per_step = args.batch * replicas
    value = tokens_per_step * args.steps / (ms_dev * 1e-3)
    e2e_value = tokens_per_step * args.steps / (ms_e2e * 1e-3)

    kern, cpu = None, None
    if rank == 0:
        kern = kernel_rooflines(run, args.kernel_reps, hbm_gbs)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_baseline as cb

            threads = os.cpu_count() or 1
            r = cb.decode_tokens_per_s(cfg.hidden, cfg.intermediate, cfg.heads, cfg.kv_heads, cfg.layers, cfg.vocab, args.batch, args.ctx, threads)
            cpu = {"value": r["tokens_per_s"], "unit": "tokens/s", "cores": threads, "kind": "port",
                   "sample": f"1 of {cfg.layers} decoder layers (4 W4A8 GEMMs M={args.batch} + KV4 attention B={args.batch} ctx={args.ctx}) + lm_head on torch-CPU "
                             f"dequant-then-matmul, extrapolated x{cfg.layers}; {r['layer_s']:.2f} s/layer (dequant {r['dequant_s']:.2f}, matmul {r['matmul_s']:.2f}, attention {r['attention_s']:.2f})"}
    def finish():
        # tear down in a fixed order: drop the CUDA graph (it holds NCCL kernels under TP) before the communicator, and do not
        # let a slow communicator teardown keep the launcher alive after the result line is out
        sys.stdout.flush()
        if world > 1:
            run.graph = None
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if world > 1:
        dist.barrier()
    if rank != 0:
        finish()
        return

    dom = kern["gemm_gate_up"]
    step_bytes = run.weight_bytes_per_step() + run.kv_bytes_per_step() + 2 * cfg.vocab * cfg.hidden
    line = {
        "metric": METRIC if (args.model, args.precision, args.batch) == ("llama-3-8b", "w4a8kv4", 64) else f"tokens/s {cfg.name} {args.precision} decode b{args.batch}",
        "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak" if tp == 1 else "strong", "vs_baseline": None,
        "dtype": "s8" if args.precision.startswith("w4a8") or args.precision.startswith("w8a8") else "f16", "data": "synthetic",
        "config": {"workload": f"{cfg.name} {args.precision} decode batch={args.batch} ctx={args.ctx} ({'CUDA-graph replay' if not args.no_graph else 'eager'}; "
                               f"{cfg.layers if args.layers is None else args.layers} layers + final norm + fp16 lm_head + greedy sampling)",
                   "precision": args.precision, "batch": args.batch, "ctx": args.ctx, "layers": run.L,
                   "parallelism": (f"dp{world}" if tp == 1 else f"tp{tp}"),
                   "l2": f"per-step working set {step_bytes / 1e9:.2f} GB (weights + KV pages + lm_head) >> 126 MB L2: inputs larger than L2, no flush needed",
                   "pdl": not args.no_pdl, "fused_small_ops": not args.no_fused},
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": args.batch * 8 * replicas, "d2h_bytes_per_step": args.batch * 8 * replicas,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": run.launches_per_step * args.steps,
        "clocks": clk.summary(),
        "roofline": {"bound": "hbm", "kernel": f"gemm_kernel W4A8 gate_up_proj M={dom['M']} N={dom['N']} K={dom['K']} (tcgen05, dominant: 54% of the GEMM bytes)",
                     "achieved": dom["GBps"], "peak": hbm_gbs, "unit": "GB/s", "frac": dom["frac"], "peak_source": peak_src,
                     "traffic": ncu_traffic("gemm_gate_up") if (dom["M"], dom["N"], dom["K"]) == (64, 28672, 4096) else None,
                     "traffic_unit": "dram bytes per launch (ncu --set full)", "algorithmic_bytes": dom["bytes"],
                     "us_per_launch": dom["us"]},
        "kernels": kern,
        "step_hbm_frac": (step_bytes / (ms_dev / args.steps * 1e-3)) / 1e9 / hbm_gbs,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
