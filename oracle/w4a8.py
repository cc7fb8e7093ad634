"""Oracle: W4A8 / W8A8 weight formats and GEMMs (TEST INFRASTRUCTURE, not product).

Restates, in numpy:
  * the offline weight packer  qserve/modeling/layers/quantized_linear/w4a8_linear.py:166-330
  * the per-channel GEMM       kernels/csrc/qgemm/w4a8_per_chn/gemm_cuda.cu:276-301 (unpack), :564-593 (epilogue)
  * the per-group  GEMM        kernels/csrc/qgemm/w4a8_per_group/gemm_cuda.cu:271-326 (level-2 dequant), :600-625 (epilogue)
  * the W8A8 GEMM              kernels/csrc/qgemm/w8a8/w8a8_gemm_cuda.cu:503-529 (epilogue)

Integer accumulators are computed exactly (float64 BLAS matmul of small integers is
exact: |acc| <= 14336*128*255 < 2^53) so they can be compared bit-for-bit with the
INT32 accumulators of the tcgen05 kernel.  The FP16 epilogues are evaluated as
individually rounded IEEE fp32 operations in the reference's source order.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# G0: weight packing  (w4a8_linear.py:196-225 / :292-322 -- identical code in both branches)
# --------------------------------------------------------------------------------------


def pack_w4(q: np.ndarray) -> np.ndarray:
    """q: uint4 codes [N, K] (values 0..15) -> qweight int8 [N, K/2] in the reference layout.

    Follows w4a8_linear.py:292-322 literally:
      reshape (N/32, 2, 2, 8, K/32, 2, 4, 4) -> permute (0,4,3,6,1,5,2,7) -> permute (0,1,2,3,5,6,7,4)
      -> byte = (x[...,1] << 4) + x[...,0] -> view [N/32, K/32, 32, 16] -> [N, K/2]
    """
    q = np.asarray(q)
    N, K = q.shape
    assert N % 32 == 0 and K % 32 == 0
    assert q.min() >= 0 and q.max() <= 15
    x = q.astype(np.int8).reshape(N // 32, 2, 2, 8, K // 32, 2, 4, 4)
    x = np.ascontiguousarray(x.transpose(0, 4, 3, 6, 1, 5, 2, 7))
    x = np.ascontiguousarray(x.transpose(0, 1, 2, 3, 5, 6, 7, 4))
    # int8 arithmetic wraps exactly like torch.int8 does: (hi << 4) + lo
    packed = ((x[..., 1].astype(np.int16) << 4) + x[..., 0].astype(np.int16)).astype(np.uint8).view(np.int8)
    return np.ascontiguousarray(packed.reshape(N // 32, K // 32, 32, 16).reshape(N, K // 2))


def unpack_w4(qweight: np.ndarray) -> np.ndarray:
    """Inverse of pack_w4 (consumer gemm_cuda.cu:286-298).

    View qweight as P[N/32][K/32][32 lanes][16 B]; for lane t = c*4+e and byte j = d*8+b*4+f:
      low  nibble = q[n32*32 + b*8 + c     ][k32*32 + d*16 + e*4 + f]
      high nibble = q[n32*32 + b*8 + c + 16][same k]
    """
    qw = np.ascontiguousarray(qweight).view(np.uint8)
    N, K2 = qw.shape
    K = K2 * 2
    P = qw.reshape(N // 32, K // 32, 8, 4, 2, 2, 4)  # [n32, k32, c, e, d, b, f]
    lo = P & 0xF
    hi = P >> 4
    q = np.empty((N // 32, 32, K // 32, 32), dtype=np.uint8)  # [n32, n_in, k32, k_in]
    qv = q.reshape(N // 32, 2, 2, 8, K // 32, 2, 4, 4)  # n_in = h*16 + b*8 + c ; k_in = d*16 + e*4 + f
    # target index order [n32, h, b, c, k32, d, e, f]; source order [n32, k32, c, e, d, b, f]
    qv[:, 0] = lo.transpose(0, 5, 2, 1, 4, 3, 6)
    qv[:, 1] = hi.transpose(0, 5, 2, 1, 4, 3, 6)
    return q.reshape(N, K)


def shuffle_n32(x: np.ndarray) -> np.ndarray:
    """Per-32-column shuffle of the level-2 params (w4a8_linear.py:236-249, 260-274).

    x: [G, N] -> [G, N] with position c*4+j <- channel j*8+c inside every group of 32 columns.
    """
    G, N = x.shape
    return np.ascontiguousarray(x.reshape(G, N // 32, 4, 8).transpose(0, 1, 3, 2)).reshape(G, N)


def unshuffle_n32(x: np.ndarray) -> np.ndarray:
    G, N = x.shape
    return np.ascontiguousarray(x.reshape(G, N // 32, 8, 4).transpose(0, 1, 3, 2)).reshape(G, N)


def pack_per_channel(q: np.ndarray, s1: np.ndarray, z: np.ndarray):
    """(q uint4 [N,K], s1 fp16 [N], z int [N]) -> (qweight, s1_scales, s1_szeros)  w4a8_linear.py:322-330."""
    s1 = np.asarray(s1, dtype=np.float16)
    # torch: zeros(int8).reshape(N) * s1(fp16) -> fp16 product (type promotion to fp16), stored fp16
    szeros = (np.asarray(z).astype(np.float32) * s1.astype(np.float32)).astype(np.float16)
    return pack_w4(q), s1.copy(), szeros


def pack_per_group(q: np.ndarray, s1: np.ndarray, s2: np.ndarray, z: np.ndarray, group: int = 128):
    """(q uint4 [N,K], s1 fp16 [N], s2 int [N,K/g] in 1..17, z int [N,K/g] in 0..15)
    -> (qweight, s1_scales, s2_scales int8 [K/g, N], s2_zeros int8 [K/g, N])   w4a8_linear.py:226-277.

    s2_zeros = (-z) * s2 computed in int32 then truncated to int8 (two's complement), both shuffled per 32 columns.
    """
    N, K = q.shape
    s2t = np.ascontiguousarray(np.asarray(s2).reshape(N, K // group).T)  # [K/g, N]
    zt = np.ascontiguousarray((-np.asarray(z).astype(np.int32)).reshape(N, K // group).T)
    s2_shuf = shuffle_n32(s2t.astype(np.int32))
    z_shuf = shuffle_n32(zt)
    s2_scales = s2_shuf.astype(np.int8)
    # reference multiplies the shuffled int32 zeros with the already-shuffled s2 (w4a8_linear.py:271-277)
    s2_zeros = (z_shuf * s2_shuf).astype(np.int8)
    return pack_w4(q), np.asarray(s1, dtype=np.float16).copy(), s2_scales, s2_zeros


# --------------------------------------------------------------------------------------
# level-2 dequant of the per-group kernel, bit faithful (w4a8_per_group/gemm_cuda.cu:286-324)
# --------------------------------------------------------------------------------------


def dequant_level2(q: np.ndarray, s2_scales: np.ndarray, s2_zeros: np.ndarray, group: int = 128) -> np.ndarray:
    """q uint4 [N,K]; s2_scales/s2_zeros int8 [K/g, N] (shuffled, as stored) -> int8 weights [N,K].

    The kernel multiplies a 32-bit word of four nibble-bytes (4 consecutive k of ONE channel) by the
    u8 scale with a plain 32-bit multiply and then adds the s8 zero with `__vadd4`:
        w = vadd4((q0 | q1<<8 | q2<<16 | q3<<24) * s2_u8, z2 replicated)
    Carries of the 32-bit multiply propagate between the 4 bytes exactly as on the GPU (they only occur
    outside QoQ's protective range 15*s2 <= 255); vadd4 wraps per byte.
    """
    N, K = q.shape
    s2 = unshuffle_n32(np.asarray(s2_scales).view(np.uint8).astype(np.uint64))  # [K/g, N] channel order
    z2 = unshuffle_n32(np.asarray(s2_zeros).view(np.uint8).astype(np.uint32))
    s2 = np.repeat(s2.T, group // 4, axis=1)  # [N, K/4] one entry per 4-k word
    z2 = np.repeat(z2.T, group // 4, axis=1)
    qq = q.astype(np.uint64).reshape(N, K // 4, 4)
    word = qq[..., 0] | (qq[..., 1] << 8) | (qq[..., 2] << 16) | (qq[..., 3] << 24)
    prod = (word * s2) & 0xFFFFFFFF
    out = np.empty((N, K // 4, 4), dtype=np.uint8)
    for b in range(4):
        out[..., b] = (((prod >> (8 * b)) & 0xFF) + z2) & 0xFF
    return out.reshape(N, K).view(np.int8)


# --------------------------------------------------------------------------------------
# exact integer accumulators
# --------------------------------------------------------------------------------------


def int_matmul(a: np.ndarray, w: np.ndarray) -> np.ndarray:
    """a int [M,K], w int [N,K] -> int32 [M,N], exact (float64 BLAS on small integers)."""
    acc = a.astype(np.float64) @ w.astype(np.float64).T
    assert np.abs(acc).max(initial=0) < 2**31
    return acc.astype(np.int64).astype(np.int32)


def _f32(x):
    return np.asarray(x).astype(np.float32)


# --------------------------------------------------------------------------------------
# G1: per-channel GEMM   out = half( float(acc)*s1[n]*sa[m] - s1z[n]*asum[m] )
# --------------------------------------------------------------------------------------


def gemm_w4a8_per_chn(a_q, qweight, wscales, ascales, w_szs, a_ssums, return_acc: bool = False):
    """w4a8_per_chn/gemm_cuda.cu:581-588:  psum*wscale*ascale - w_sz*a_ssum, left to right in fp32."""
    q = unpack_w4(qweight)
    acc = int_matmul(np.asarray(a_q, dtype=np.int8), q)
    psum = acc.astype(np.float32)  # __int2float_rn
    t = (psum * _f32(wscales)[None, :]).astype(np.float32)
    t = (t * _f32(ascales)[:, None]).astype(np.float32)
    u = (_f32(w_szs)[None, :] * _f32(a_ssums)[:, None]).astype(np.float32)
    out = (t - u).astype(np.float32).astype(np.float16)
    return (out, acc) if return_acc else out


# --------------------------------------------------------------------------------------
# G2: per-group GEMM    out = half( float(acc) * (s1[n]*sa[m]) )
# --------------------------------------------------------------------------------------


def gemm_w4a8_per_group(a_q, qweight, s2_zeros, s2_scales, wscales, ascales, group: int = 128,
                        return_acc: bool = False):
    """w4a8_per_group/gemm_cuda.cu:617-621: psum *= wscale*ascale (note the association)."""
    q = unpack_w4(qweight)
    w8 = dequant_level2(q, s2_scales, s2_zeros, group)
    acc = int_matmul(np.asarray(a_q, dtype=np.int8), w8)
    psum = acc.astype(np.float32)
    sc = (_f32(wscales)[None, :] * _f32(ascales)[:, None]).astype(np.float32)
    out = (psum * sc).astype(np.float32).astype(np.float16)
    return (out, acc) if return_acc else out


# --------------------------------------------------------------------------------------
# G3: W8A8 GEMM         out = half( float(acc) * (sw[n]*sa[m]) )
# --------------------------------------------------------------------------------------


def gemm_w8a8(a_q, weight, wscales, ascales, return_acc: bool = False):
    """w8a8/w8a8_gemm_cuda.cu:517-525."""
    acc = int_matmul(np.asarray(a_q, dtype=np.int8), np.asarray(weight, dtype=np.int8))
    psum = acc.astype(np.float32)
    sc = (_f32(wscales)[None, :] * _f32(ascales)[:, None]).astype(np.float32)
    out = (psum * sc).astype(np.float32).astype(np.float16)
    return (out, acc) if return_acc else out


# --------------------------------------------------------------------------------------
# "torch-CPU dequant-then-FP16-matmul" restatement -- the CPU baseline leg
# --------------------------------------------------------------------------------------


def dequant_then_matmul_per_chn(a_q, qweight, s1, s1z, ascales):
    """Y = (a_q*sa) @ ((q - z)*s1)^T, fp32 accumulate, fp16 result; z recovered as s1z/s1."""
    q = unpack_w4(qweight).astype(np.float32)
    w = q * _f32(s1)[:, None] - _f32(s1z)[:, None]
    x = np.asarray(a_q).astype(np.float32) * _f32(ascales)[:, None]
    return (x @ w.T).astype(np.float16)


# --------------------------------------------------------------------------------------
# synthetic weight generators used by tests and bench
# --------------------------------------------------------------------------------------


def synth_per_channel(rng: np.random.Generator, N: int, K: int):
    q = rng.integers(0, 16, size=(N, K), dtype=np.uint8)
    s1 = rng.uniform(0.005, 0.02, size=N).astype(np.float16)
    z = rng.integers(0, 16, size=N)
    return (q,) + pack_per_channel(q, s1, z)


def synth_per_group(rng: np.random.Generator, N: int, K: int, group: int = 128):
    q = rng.integers(0, 16, size=(N, K), dtype=np.uint8)
    s1 = rng.uniform(0.0005, 0.002, size=N).astype(np.float16)
    s2 = rng.integers(1, 17, size=(N, K // group))  # protective range: 15*s2 <= 255
    z = rng.integers(0, 16, size=(N, K // group))
    return (q,) + pack_per_group(q, s1, s2, z, group)


def fake_quant_per_channel(w: np.ndarray):
    """pseudo_quantize_tensor(n_bit=4, zero_point=True, q_group_size=-1) semantics
    (scripts/ckpt_converter/quant_utils.py:96-138) -> (q uint4, s1 fp16, z int)."""
    w = np.asarray(w, dtype=np.float32)
    mx = w.max(axis=1, keepdims=True)
    mn = w.min(axis=1, keepdims=True)
    scales = np.maximum(mx - mn, 1e-5) / 15
    zeros = np.clip(-np.rint(mn / scales), 0, 15)
    s1 = scales.astype(np.float16)
    q = np.clip(np.rint(w / s1.astype(np.float32)) + zeros, 0, 15).astype(np.uint8)
    return q, s1.reshape(-1), zeros.reshape(-1).astype(np.int64)
