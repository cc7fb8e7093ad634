"""CPU oracle (test infrastructure, never on the product path): causal variable-length self-attention of the prompt step.

The reference does not implement this itself: `llama_w4a8_unpad.py:232-242` calls the third-party
`flash_attn.flash_attn_varlen_func(q, k, v, cu_seqlens_q=cu, cu_seqlens_k=cu, max_seqlen_q=L, max_seqlen_k=L, dropout_p=0.0, causal=True)`
(flash-attn 2.x, not under /root/reference; pyproject does not pin it, README suggests v2.5.8).  Its published semantics, restated here in
float64: per sequence b and query head h (kv head h // (Hq / Hkv)),  out[i] = softmax_j<=i(q[i] . k[j] / sqrt(D)) @ v.
flash-attn computes the same in fp32 with P rounded to fp16 before the second product; the parity tests state the resulting tolerance and,
where flash-attn is importable (it is in the GPU image), also compare against its actual output.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def causal_varlen_attention(q: np.ndarray, k: np.ndarray, v: np.ndarray, cu_seqlens: Sequence[int], softmax_scale: Optional[float] = None) -> np.ndarray:
    """q [T, Hq, D], k / v [T, Hkv, D] (any float dtype) -> float64 [T, Hq, D]."""
    T, Hq, D = q.shape
    Hkv = k.shape[1]
    assert Hq % Hkv == 0 and k.shape == v.shape and k.shape[0] == T
    g = Hq // Hkv
    scale = float(softmax_scale) if softmax_scale is not None else D ** -0.5
    out = np.zeros((T, Hq, D), dtype=np.float64)
    for b in range(len(cu_seqlens) - 1):
        s, e = int(cu_seqlens[b]), int(cu_seqlens[b + 1])
        L = e - s
        if L == 0:
            continue
        mask = np.triu(np.ones((L, L), dtype=bool), 1)
        for h in range(Hq):
            qq = q[s:e, h].astype(np.float64)
            kk = k[s:e, h // g].astype(np.float64)
            vv = v[s:e, h // g].astype(np.float64)
            sc = qq @ kk.T * scale
            sc[mask] = -np.inf
            sc -= sc.max(axis=1, keepdims=True)
            pr = np.exp(sc)
            out[s:e, h] = (pr / pr.sum(axis=1, keepdims=True)) @ vv
    return out
