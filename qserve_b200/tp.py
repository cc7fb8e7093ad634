"""Tensor-parallel sharding of QServe checkpoints (host-side logic).

The packed INT4 layout (w4a8_linear.py:292-322) stores one 32x32 tile per 512 contiguous bytes, tiles K-major inside
a 32-channel band, so
  * column parallel (split N: qkv_proj, gate_up_proj) = a contiguous slice of bands, any multiple of 32 channels
    (the sm_100a GEMM wants multiples of 128);
  * row parallel (split K: o_proj, down_proj) = a slice of every band's tile range, K/TP a multiple of 128 so that the
    g128 groups and the kernel's 128-wide K blocks stay whole.
Level-2 params [K/128, N] split along dim 0 (row parallel) or along N in whole 32-column shuffle groups (column parallel).
"""
from __future__ import annotations

import torch


def shard_columns(qweight: torch.Tensor, rank: int, size: int) -> torch.Tensor:
    """Column-parallel shard of qweight [N, K/2]: output channels [rank*N/size, (rank+1)*N/size)."""
    N = qweight.size(0)
    assert N % (128 * size) == 0, "N/TP must be a multiple of 128"
    n = N // size
    return qweight[rank * n:(rank + 1) * n].contiguous()


def shard_rows(qweight: torch.Tensor, rank: int, size: int) -> torch.Tensor:
    """Row-parallel shard of qweight [N, K/2]: input channels [rank*K/size, (rank+1)*K/size) of every output channel."""
    N, K = qweight.size(0), qweight.size(1) * 2
    assert K % (128 * size) == 0, "K/TP must be a multiple of 128"
    tiles = qweight.reshape(N // 32, K // 32, 512)
    t = (K // 32) // size
    return tiles[:, rank * t:(rank + 1) * t].reshape(N, (K // size) // 2).contiguous()


def shard_vector(v: torch.Tensor, rank: int, size: int) -> torch.Tensor:
    n = v.size(-1) // size
    return v[..., rank * n:(rank + 1) * n].contiguous()


def shard_level2_columns(p: torch.Tensor, rank: int, size: int) -> torch.Tensor:
    """s2_scales / s2_zeros [K/128, N], column parallel: whole 32-column shuffle groups move together."""
    return shard_vector(p, rank, size)


def shard_level2_rows(p: torch.Tensor, rank: int, size: int) -> torch.Tensor:
    g = p.size(0) // size
    return p[rank * g:(rank + 1) * g].contiguous()
