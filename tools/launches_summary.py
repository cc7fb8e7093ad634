#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list of one decode step into per-kernel counts / times / shares.
usage: python tools/launches_summary.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import collections
import csv
import sys

GROUPS = [("gemm_kernel", "gemm_kernel (W4A8 tcgen05)"), ("decode_attention", "decode_attention_kernel"), ("add_layernorm", "add_layernorm_quant_kernel"),
          ("layernorm_quant", "layernorm_quant_kernel"), ("silu_mul_quant", "silu_mul_quant_kernel"), ("quant_per_token", "quant_per_token_kernel"),
          ("silu_and_mul", "silu_and_mul_kernel"), ("rms_norm", "rms_norm_kernel"), ("argmax_rows", "argmax_rows_kernel"), ("gemv", "cuBLAS (lm_head)"),
          ("cutlass", "cuBLAS (lm_head)"), ("nvjet", "cuBLAS (lm_head)"), ("gemm", "cuBLAS (lm_head)")]


def group(name):
    for key, label in GROUPS:
        if key in name:
            return label
    return "torch: " + name.split("<")[0][-60:]


rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
n, t = collections.Counter(), collections.Counter()
for r in rows:
    g = group(r[4])
    n[g] += 1
    t[g] += float(r[-1]) / 1e3
total = sum(t.values())
print(f"# {len(rows)} launches, {total:.1f} us of kernel time (ncu gpu__time_duration: cold-cache, serialised -- compare SHARES, not absolutes)")
for g, us in t.most_common():
    print(f"{g:44s} n={n[g]:4d}  total {us:9.1f} us  mean {us / n[g]:8.2f} us  share {100 * us / total:5.1f} %")
