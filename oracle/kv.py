"""Oracle: paged KV4/KV8 cache format, RoPE, KV quantisation, prefill append and the
single-query decode attention (TEST INFRASTRUCTURE, not product).

Restates
  * page layout            kernels/csrc/fused_attention/kvCacheUtils.h:47-126, cache_engine.py:60-66
  * KV quant / dequant     decoderMaskedMultiheadAttentionUtils.h:1838-1852, 2055-2077, 2125-2213
  * NeoX RoPE              decoderMaskedMultiheadAttentionUtils.h:1147-1169, 2536-2612
  * decode attention       decoderMaskedMultiheadAttentionTemplate.hpp:717-2222 (ZINT4 / ZINT8, Dh=128, 256 threads)
  * prefill append         applyBiasRopeUpdateKVCache.h:94-455, update_kv_cache.cu:20-108
  * padding offsets        input_metadata_helper.cu:11-31

The pool is modelled as a numpy uint8 array [num_pages, page_bytes] plus integer block tables
(page indices); GPU tests turn the indices into the int64 device-address table the ops take.

Floating-point fidelity: every fp16 rounding the reference performs is reproduced (fp16 dequant FMA,
fp16 per-thread QK partial dot products, fp16 logits, fp16 tree reduction of the output); fp32
reductions whose order only moves the last fp32 bit are evaluated in float64 and rounded once.
"""
from __future__ import annotations

import numpy as np

from .ops import cvt_rni_sat_u8, f16

TOKENS_PER_PAGE = 64


# ------------------------------------------------------------------------------------------------
# K0: page geometry   (cache_engine.py:60-66, Template.hpp:924-930)
# ------------------------------------------------------------------------------------------------


def code_bytes(num_kv_heads: int, head_dim: int, bits: int) -> int:
    return num_kv_heads * TOKENS_PER_PAGE * head_dim * bits // 8


def page_bytes(num_kv_heads: int, head_dim: int, bits: int) -> int:
    """codes [Hkv][64][D*bits/8] | scales fp16 [Hkv][64] | zeros fp16 [Hkv][64]."""
    return code_bytes(num_kv_heads, head_dim, bits) + num_kv_heads * TOKENS_PER_PAGE * 4


class PagePool:
    """One layer's K (or V) pool: uint8 [num_pages, page_bytes]."""

    def __init__(self, num_pages: int, num_kv_heads: int, head_dim: int, bits: int, rng=None):
        self.Hkv, self.D, self.bits = num_kv_heads, head_dim, bits
        self.cb = code_bytes(num_kv_heads, head_dim, bits)
        self.pb = page_bytes(num_kv_heads, head_dim, bits)
        self.data = np.zeros((num_pages, self.pb), dtype=np.uint8)
        if rng is not None:
            self.randomize(rng)

    def randomize(self, rng, scale_range=(0.01, 0.1), zero_range=(0.0, None)):
        """random codes, scale~U(0.01,0.1), zero~U(0,15|255) fp16."""
        n = self.data.shape[0]
        self.data[:, : self.cb] = rng.integers(0, 256, size=(n, self.cb), dtype=np.uint8)
        zmax = zero_range[1] if zero_range[1] is not None else (15.0 if self.bits == 4 else 255.0)
        s = rng.uniform(*scale_range, size=(n, self.Hkv, TOKENS_PER_PAGE)).astype(np.float16)
        z = rng.uniform(zero_range[0], zmax, size=(n, self.Hkv, TOKENS_PER_PAGE)).astype(np.float16)
        self.scales()[:] = s
        self.zeros()[:] = z

    def codes(self):
        """uint8 view [pages, Hkv, 64, D*bits/8]."""
        return self.data[:, : self.cb].reshape(-1, self.Hkv, TOKENS_PER_PAGE, self.D * self.bits // 8)

    def _meta(self):
        return self.data[:, self.cb:].view(np.float16).reshape(-1, 2, self.Hkv, TOKENS_PER_PAGE)

    def scales(self):
        return self._meta()[:, 0]

    def zeros(self):
        return self._meta()[:, 1]


# ------------------------------------------------------------------------------------------------
# A4: KV quant / dequant
# ------------------------------------------------------------------------------------------------


def kv_quant_params(x, bits: int):
    """x fp16 [..., D] -> (scale fp16, zero fp16) per (token, kv head)    Template.hpp:1243, 1067.

    scale = half((max-min)/L) ; zero = half(-L*min/(max-min)) with L = 15 | 255 ; all in fp32.
    """
    L = np.float32(15.0 if bits == 4 else 255.0)
    xf = np.asarray(x, np.float16).astype(np.float32)
    mx, mn = xf.max(axis=-1), xf.min(axis=-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        d = (mx - mn).astype(np.float32)
        scale = (d / L).astype(np.float32).astype(np.float16)
        zero = (((-L) * mn).astype(np.float32) / d).astype(np.float32).astype(np.float16)
    return scale, zero


def kv_quant_codes(x, scale, zero, bits: int):
    """u = cvt.rni.sat.u8(x*(1/float(scale)) + float(zero)) [& 0xF for 4 bit]   Utils.h:2047-2077, 1838-1852."""
    xf = np.asarray(x, np.float16).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv = (np.float32(1.0) / np.asarray(scale, np.float16).astype(np.float32)).astype(np.float32)
        t = (xf * inv[..., None]).astype(np.float32)
        t = (t + np.asarray(zero, np.float16).astype(np.float32)[..., None]).astype(np.float32)
    u = cvt_rni_sat_u8(t)
    return (u & 0xF) if bits == 4 else u


def pack_nibbles(u):
    """[..., D] uint4 -> [..., D/2] bytes, even element in the low nibble (Utils.h:1838-1852)."""
    u = np.asarray(u, np.uint8)
    return (u[..., 0::2] & 0xF) | ((u[..., 1::2] & 0xF) << 4)


def unpack_nibbles(b):
    b = np.asarray(b, np.uint8)
    out = np.empty(b.shape[:-1] + (b.shape[-1] * 2,), np.uint8)
    out[..., 0::2] = b & 0xF
    out[..., 1::2] = b >> 4
    return out


def kv_dequant(u, scale, zero, bits: int):
    """codes uint8 [..., D] -> fp16.
    4 bit: x = fma_f16(half(u), s, half(-float(s)*float(z)))          Utils.h:2190-2213
    8 bit: x = half( float(s) * (float(u) - float(z)) )               Utils.h:2095-2107
    """
    s = np.asarray(scale, np.float16)
    z = np.asarray(zero, np.float16)
    if bits == 4:
        c = (-(s.astype(np.float32)) * z.astype(np.float32)).astype(np.float32).astype(np.float16)
        return f16(np.asarray(u).astype(np.float64) * s.astype(np.float64)[..., None] + c.astype(np.float64)[..., None])
    d = (np.asarray(u).astype(np.float32) - z.astype(np.float32)[..., None]).astype(np.float32)
    return (s.astype(np.float32)[..., None] * d).astype(np.float32).astype(np.float16)


# ------------------------------------------------------------------------------------------------
# RoPE (NeoX)
# ------------------------------------------------------------------------------------------------


def rope_neox(x, pos, base: float, rot_dim: int | None = None):
    """x fp16 [..., D], pos int broadcastable to x.shape[:-1] -> fp16.

    inv_freq = pos / base^(2i/rot_dim) ; x'[i] = c*x[i] - s*x[i+R/2] ; x'[i+R/2] = c*x[i+R/2] + s*x[i]
    evaluated in fp32 and rounded to fp16   (Utils.h:1147-1169, 2536-2558).
    """
    x = np.asarray(x, np.float16)
    D = x.shape[-1]
    R = D if rot_dim is None else rot_dim
    half = R // 2
    i = np.arange(half, dtype=np.float32)
    denom = np.power(np.float32(base), (2.0 * i / np.float32(R)).astype(np.float32)).astype(np.float32)
    ang = (np.asarray(pos, np.float32)[..., None] / denom).astype(np.float32)
    c, s = np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)
    xf = x.astype(np.float32)
    a, b = xf[..., :half], xf[..., half:R]
    out = xf.copy()
    out[..., :half] = ((c * a).astype(np.float32) - (s * b).astype(np.float32)).astype(np.float32)
    out[..., half:R] = ((c * b).astype(np.float32) + (s * a).astype(np.float32)).astype(np.float32)
    return out.astype(np.float16)


# ------------------------------------------------------------------------------------------------
# writing one token into the pool
# ------------------------------------------------------------------------------------------------


def pool_write_token(pool: PagePool, page: int, slot: int, x):
    """x fp16 [Hkv, D] (post-RoPE for K): quantise per kv head and store codes, scale, zero."""
    s, z = kv_quant_params(x, pool.bits)
    u = kv_quant_codes(x, s, z, pool.bits)
    pool.codes()[page, :, slot, :] = pack_nibbles(u) if pool.bits == 4 else u
    pool.scales()[page, :, slot] = s
    pool.zeros()[page, :, slot] = z


def pool_read_tokens(pool: PagePool, block_row, n_tokens: int):
    """-> (codes uint8 [Hkv, n, D], scale fp16 [Hkv, n], zero fp16 [Hkv, n]) for tokens 0..n-1."""
    t = np.arange(n_tokens)
    pages = np.asarray(block_row)[t // TOKENS_PER_PAGE]
    slots = t % TOKENS_PER_PAGE
    raw = pool.codes()[pages, :, slots, :]  # [n, Hkv, Db]
    codes = unpack_nibbles(raw) if pool.bits == 4 else raw
    return (np.ascontiguousarray(codes.transpose(1, 0, 2)), pool.scales()[pages, :, slots].T.copy(),
            pool.zeros()[pages, :, slots].T.copy())


# ------------------------------------------------------------------------------------------------
# K3: compute_padding_offsets     input_metadata_helper.cu:11-31
# ------------------------------------------------------------------------------------------------


def compute_padding_offsets(cu_seqlens, max_seqlen: int, total_tokens: int):
    cu = np.asarray(cu_seqlens, np.int64)
    out = np.zeros(total_tokens, np.int32)
    for b in range(len(cu) - 1):
        out[cu[b]: cu[b + 1]] = b * max_seqlen - cu[b]
    return out


# ------------------------------------------------------------------------------------------------
# K2: apply_bias_rope_update_kv_cache (prefill)     applyBiasRopeUpdateKVCache.h:94-455
# ------------------------------------------------------------------------------------------------


def prefill_rope_append(qkv, seq_lens, padding_offset, kpool: PagePool, vpool: PagePool, block_tables,
                        num_heads: int, num_kv_heads: int, max_seq_len: int, rope_base: float,
                        max_positions: int):
    """qkv fp16 [T, (Hq+2Hkv)*D] modified IN PLACE (q and k rotated; v untouched); pages appended.

    token -> (batch, position):  g = t + padding_offset[t]; b = g // max_seq_len; pos = g % max_seq_len  (:188-201)
    kv stored only when pos >= max(len[b] - max_positions, 0)                                           (:268-271)
    """
    qkv = np.asarray(qkv)
    T = qkv.shape[0]
    D = kpool.D
    Hq, Hkv = num_heads, num_kv_heads
    g = np.arange(T) + np.asarray(padding_offset, np.int64)
    b = g // max_seq_len
    pos = g % max_seq_len
    q = qkv[:, : Hq * D].reshape(T, Hq, D)
    k = qkv[:, Hq * D: (Hq + Hkv) * D].reshape(T, Hkv, D)
    v = qkv[:, (Hq + Hkv) * D:].reshape(T, Hkv, D)
    q[:] = rope_neox(q, pos[:, None], rope_base)
    k[:] = rope_neox(k, pos[:, None], rope_base)
    lens = np.asarray(seq_lens, np.int64)
    for t in range(T):
        if pos[t] >= max(lens[b[t]] - max_positions, 0) and pos[t] < lens[b[t]]:
            page = int(np.asarray(block_tables)[b[t], pos[t] // TOKENS_PER_PAGE])
            pool_write_token(kpool, page, int(pos[t] % TOKENS_PER_PAGE), k[t])
            pool_write_token(vpool, page, int(pos[t] % TOKENS_PER_PAGE), v[t])
    return qkv


# ------------------------------------------------------------------------------------------------
# K1: single_query_attention (decode)      Template.hpp:717-2222
# ------------------------------------------------------------------------------------------------

_THREADS = 256
_THREADS_PER_KEY = 16  # 8 fp16 elements (16 B of fp16 / 4 B of int4) per thread   Template.hpp:813-820
_V_PER_ITER = 16  # 256 threads / THREADS_PER_VALUE(16)


def _qk_fp16_partial(q_h, k_h):
    """Per-thread dot of 8 fp16 elements exactly as qk_hmma_dot_simple (Template.hpp:445-467) after the
    04152637 reorder (Utils.h:1943-1954): lanes (e0..e3) and (e4..e7) accumulate with fp16 FMAs, then hadd.
    q_h: [..., 16, 8] fp16, k_h: [..., 16, 8] fp16 -> fp32 [..., 16]."""
    q64, k64 = q_h.astype(np.float64), k_h.astype(np.float64)
    lo = f16(q64[..., 0] * k64[..., 0])
    hi = f16(q64[..., 4] * k64[..., 4])
    for j in (1, 2, 3):
        lo = f16(q64[..., j] * k64[..., j] + lo.astype(np.float64))
        hi = f16(q64[..., 4 + j] * k64[..., 4 + j] + hi.astype(np.float64))
    return f16(lo.astype(np.float64) + hi.astype(np.float64)).astype(np.float32)


def _butterfly_sum16(p):
    """fp32 shfl_xor sum over 16 lanes with masks 8,4,2,1 (lane 0 result)."""
    p = p.astype(np.float32)
    for m in (8, 4, 2, 1):
        idx = np.arange(16) ^ m
        p = (p + p[..., idx]).astype(np.float32)
    return p[..., 0]


def decode_attention(q, k, v, kpool: PagePool, vpool: PagePool, block_tables, lengths, rope_base: float,
                     faithful: bool = True):
    """q fp16 [B,Hq,D]; k,v fp16 [B,Hkv,D] (new token, pre-RoPE).  lengths[b] counts the new token (A6).

    Side effect: the new K (post-RoPE) and V are quantised and appended at index lengths[b]-1.
    Returns out fp16 [B,Hq,D].  With faithful=False the arithmetic after dequantisation is float64
    ("exact given the cache"), used to measure how far each implementation is from the truth.
    """
    q = np.asarray(q, np.float16)
    k = np.asarray(k, np.float16)
    v = np.asarray(v, np.float16)
    B, Hq, D = q.shape
    Hkv = k.shape[1]
    G = Hq // Hkv
    inv_sqrt = np.float32(1.0) / np.sqrt(np.float32(D))
    out = np.zeros((B, Hq, D), np.float16)
    for b in range(B):
        tlen = int(lengths[b]) - 1
        qr = rope_neox(q[b], tlen, rope_base)  # [Hq, D]
        kr = rope_neox(k[b], tlen, rope_base)  # [Hkv, D]
        page = int(np.asarray(block_tables)[b, tlen // TOKENS_PER_PAGE])
        slot = tlen % TOKENS_PER_PAGE
        pool_write_token(kpool, page, slot, kr)
        pool_write_token(vpool, page, slot, v[b])
        if tlen > 0:
            kc, ks, kz = pool_read_tokens(kpool, block_tables[b], tlen)
            vc, vs, vz = pool_read_tokens(vpool, block_tables[b], tlen)
            kd = kv_dequant(kc, ks, kz, kpool.bits)  # [Hkv, tlen, D] fp16
            vd = kv_dequant(vc, vs, vz, vpool.bits)
        for h in range(Hq):
            hk = h // G
            if not faithful:
                logits = np.empty(tlen + 1, np.float64)
                if tlen > 0:
                    logits[:tlen] = kd[hk].astype(np.float64) @ qr[h].astype(np.float64)
                logits[tlen] = kr[hk].astype(np.float64) @ qr[h].astype(np.float64)
                logits *= float(inv_sqrt)
                p = np.exp(logits - logits.max())
                p /= p.sum()
                o = p[tlen] * v[b, hk].astype(np.float64)
                if tlen > 0:
                    o = o + p[:tlen] @ vd[hk].astype(np.float64)
                out[b, h] = o.astype(np.float16)
                continue
            # current token: fp32 dot of the un-quantised (rotated) q,k   (:1410-1418)
            qk = np.empty(tlen + 1, np.float32)
            qk[tlen] = np.float32(qr[h].astype(np.float64) @ kr[hk].astype(np.float64)) * inv_sqrt
            if tlen > 0:
                part = _qk_fp16_partial(np.broadcast_to(qr[h].reshape(1, 16, 8), (tlen, 16, 8)),
                                        kd[hk].reshape(tlen, 16, 8))
                qk[:tlen] = (_butterfly_sum16(part) * inv_sqrt).astype(np.float32)
            mx = qk.max()
            e = np.exp((qk - mx).astype(np.float32)).astype(np.float32)  # __expf
            ssum = np.float32(e.astype(np.float64).sum())
            inv_sum = np.float32(1.0) / (ssum + np.float32(1e-6))
            p = (e * inv_sum).astype(np.float32).astype(np.float16)  # logits stored in fp16 (:1831)
            # 16 value groups accumulate their tokens in fp32; vo = token % 16   (:1892-1977)
            acc = np.zeros((_V_PER_ITER, D), np.float64)  # holds fp32 values; fp32 FMA = one rounding of the exact a*b+c
            for j in range((tlen + _V_PER_ITER - 1) // _V_PER_ITER):
                ti = j * _V_PER_ITER + np.arange(_V_PER_ITER)
                ok = ti < tlen
                tt = ti[ok]
                acc[ok] = (p[tt].astype(np.float64)[:, None] * vd[hk, tt].astype(np.float64) + acc[ok]).astype(np.float32)
            vo = tlen % _V_PER_ITER
            acc[vo] = (float(p[tlen]) * v[b, hk].astype(np.float64) + acc[vo]).astype(np.float32)
            acc = acc.astype(np.float32)
            # tree reduction through fp16 shared memory (:2163-2187)
            active = _V_PER_ITER
            while active >= 2:
                mid = active // 2
                upper = acc[mid:active].astype(np.float16).astype(np.float32)
                acc[:mid] = (upper + acc[:mid]).astype(np.float32)
                active = mid
            out[b, h] = acc[0].astype(np.float16)
    return out


# ------------------------------------------------------------------------------------------------
#  prefill attention: what `flash_attn_varlen_func(q, k, v, cu_seqlens, ..., causal=True)` computes at
# llama_w4a8_unpad.py:232 on the rotated q / k of the packed qkv buffer -- un-quantised K / V, causal, grouped-query.
# Reference for the B200 prefill kernel of the next round; float64, no rounding model (flash-attn is third-party and unpinned).
# ------------------------------------------------------------------------------------------------


def prefill_attention(qkv, cu_seqlens, num_heads: int, num_kv_heads: int, head_dim: int = 128):
    """qkv fp16 [T, (Hq+2Hkv)*D] (q and k already rotated by `prefill_rope_append`), cu_seqlens int [B+1]
    -> out float64 [T, Hq*D]: softmax(q k^T / sqrt(D), causal within each sequence) v, query head h reads kv head h // (Hq/Hkv)."""
    qkv = np.asarray(qkv, dtype=np.float64)
    T = qkv.shape[0]
    Hq, Hkv, D = num_heads, num_kv_heads, head_dim
    G = Hq // Hkv
    q = qkv[:, : Hq * D].reshape(T, Hq, D)
    k = qkv[:, Hq * D: (Hq + Hkv) * D].reshape(T, Hkv, D)
    v = qkv[:, (Hq + Hkv) * D:].reshape(T, Hkv, D)
    out = np.zeros((T, Hq, D))
    cu = np.asarray(cu_seqlens, np.int64)
    for b in range(len(cu) - 1):
        s, e = int(cu[b]), int(cu[b + 1])
        L = e - s
        if L == 0:
            continue
        mask = np.tril(np.ones((L, L), dtype=bool))
        for h in range(Hq):
            logits = (q[s:e, h] @ k[s:e, h // G].T) / np.sqrt(D)
            logits = np.where(mask, logits, -np.inf)
            p = np.exp(logits - logits.max(axis=1, keepdims=True))
            p /= p.sum(axis=1, keepdims=True)
            out[s:e, h] = p @ v[s:e, h // G]
    return out.reshape(T, Hq * D)
