"""GPU parity: the tcgen05 prompt-phase attention (qs_prefill_attention / backend.flash_attn_varlen_func) against the float64 oracle and,
where importable, against flash-attn itself (the third-party kernel the reference calls at llama_w4a8_unpad.py:232-242).

Stated tolerance (floating point): P is rounded to fp16 before the second matrix product -- as in flash-attn -- and the output is fp16, so
    |out - exact| <= 2 fp16 ulps of |exact| + 1.5e-3 * max|v|
per element (measured on a B200: 2.1e-3 at max|v| = 6, flash-attn 2.8.3 on the same inputs: 2.0e-3), and the maximum error may not exceed
1.25 x flash-attn's own maximum error against the oracle + 1e-3.
"""
import numpy as np
import pytest
import torch

from oracle import prefill_attention as oracle_pa

pytestmark = pytest.mark.gpu


def _inputs(lens, hq, hkv, seed, dev, sigma=1.0):
    T = sum(lens)
    g = torch.Generator(device="cpu").manual_seed(seed)
    # q, k, v as the reference passes them: strided views of one packed qkv buffer (llama_w4a8_unpad.py:224-231)
    qkv = (torch.randn(T, (hq + 2 * hkv) * 128, generator=g) * sigma).half().to(dev)
    q, k, v = qkv.split([hq * 128, hkv * 128, hkv * 128], dim=-1)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    return q.reshape(T, hq, 128), k.reshape(T, hkv, 128), v.reshape(T, hkv, 128), cu, torch.from_numpy(cu).to(dev)


def _tol(exact, vmax):
    a = np.abs(exact)
    ulp = np.where(a >= 2.0 ** -14, 2.0 ** (np.floor(np.log2(np.maximum(a, 2.0 ** -14))) - 10), 2.0 ** -24)
    return 2 * ulp + 1.5e-3 * vmax


def _flash_attn():
    try:
        from flash_attn import flash_attn_varlen_func
        return flash_attn_varlen_func
    except Exception:  # noqa: BLE001
        return None


@pytest.mark.parametrize("lens,hq,hkv", [([128], 1, 1), ([1], 2, 1), ([127], 4, 2), ([129], 4, 1), ([256, 1, 130, 64], 8, 2), ([333, 200], 4, 4),
                                         ([640], 32, 8)])
def test_matches_oracle(dev, lens, hq, hkv):
    from qserve_b200 import backend
    q, k, v, cu, cu_d = _inputs(lens, hq, hkv, sum(lens) + hq, dev, sigma=1.5)
    out = backend.flash_attn_varlen_func(q, k, v, cu_d, cu_d, max(lens), max(lens), dropout_p=0.0, causal=True)
    torch.cuda.synchronize()
    assert out.shape == q.shape and out.dtype == torch.half and out.is_contiguous()
    exact = oracle_pa.causal_varlen_attention(q.cpu().numpy(), k.cpu().numpy(), v.cpu().numpy(), cu)
    got = out.cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    err = np.abs(got - exact)
    assert (err <= _tol(exact, float(v.abs().max()))).all(), f"max err {err.max():.3e}"
    fa = _flash_attn()
    if fa is not None:
        ref = fa(q, k, v, cu_d, cu_d, max(lens), max(lens), dropout_p=0.0, causal=True).cpu().numpy().astype(np.float64)
        assert err.max() <= 1.25 * np.abs(ref - exact).max() + 1e-3


def test_softmax_scale_and_first_token(dev):
    """A non-default scale; and position 0 of every sequence attends to itself only: out == v exactly (P = 1.0 in fp16)."""
    from qserve_b200 import backend
    lens = [70, 200, 5]
    q, k, v, cu, cu_d = _inputs(lens, 8, 2, 5, dev)
    out = backend.flash_attn_varlen_func(q, k, v, cu_d, cu_d, 200, 200, dropout_p=0.0, softmax_scale=0.05, causal=True)
    exact = oracle_pa.causal_varlen_attention(q.cpu().numpy(), k.cpu().numpy(), v.cpu().numpy(), cu, softmax_scale=0.05)
    assert (np.abs(out.cpu().numpy().astype(np.float64) - exact) <= _tol(exact, float(v.abs().max()))).all()
    for b in range(len(lens)):
        first = int(cu[b])
        assert torch.equal(out[first], v[first].repeat_interleave(4, dim=0))


def test_large_logits_do_not_overflow(dev):
    """Row maxima that grow from key block to key block (the lazy-rescale path) and logits of a few hundred: finite and within tolerance."""
    from qserve_b200 import backend
    lens = [512]
    q, k, v, cu, cu_d = _inputs(lens, 2, 1, 9, dev, sigma=1.0)
    ramp = torch.linspace(0.5, 6.0, 512, device=dev).half()[:, None, None]
    k = (k * ramp).contiguous()  # later keys have larger norms: the running maximum keeps growing
    q = (q * 4).contiguous()
    out = backend.flash_attn_varlen_func(q, k, v, cu_d, cu_d, 512, 512, dropout_p=0.0, causal=True)
    exact = oracle_pa.causal_varlen_attention(q.cpu().numpy(), k.cpu().numpy(), v.cpu().numpy(), cu)
    got = out.cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    assert (np.abs(got - exact) <= _tol(exact, float(v.abs().max()))).all()


def test_config3_prompt_batch_full_size(dev):
    """BASELINE config 3 prompt step: 8 prompts x 1024 tokens, Llama-3-8B heads (32 / 8).  Too large for the numpy oracle: a float32 torch
    reference on the device, every head and row."""
    from qserve_b200 import backend
    lens = [1024] * 8
    q, k, v, cu, cu_d = _inputs(lens, 32, 8, 3, dev)
    out = backend.flash_attn_varlen_func(q, k, v, cu_d, cu_d, 1024, 1024, dropout_p=0.0, causal=True)
    mask = torch.triu(torch.ones(1024, 1024, dtype=torch.bool, device=dev), 1)
    worst = 0.0
    for b in range(8):
        s = slice(int(cu[b]), int(cu[b + 1]))
        qq = q[s].float().transpose(0, 1)
        kk = k[s].float().repeat_interleave(4, dim=1).transpose(0, 1)
        vv = v[s].float().repeat_interleave(4, dim=1).transpose(0, 1)
        sc = (qq @ kk.transpose(1, 2)) * 128 ** -0.5
        ref = (torch.softmax(sc.masked_fill(mask, float("-inf")), dim=-1) @ vv).transpose(0, 1)
        worst = max(worst, float((out[s].float() - ref).abs().max()))
    assert worst <= 3e-3, worst


def test_empty_and_argument_errors(dev):
    from qserve_b200 import backend
    q, k, v, cu, cu_d = _inputs([64], 4, 2, 1, dev)
    empty = backend.flash_attn_varlen_func(q[:0], k[:0], v[:0], cu_d[:1], cu_d[:1], 0, 0, dropout_p=0.0, causal=True)
    assert empty.shape == (0, 4, 128)
    with pytest.raises(RuntimeError):
        backend.flash_attn_varlen_func(q, k, v, cu_d, cu_d, 64, 64, dropout_p=0.0, causal=False)
    with pytest.raises(RuntimeError):
        backend.flash_attn_varlen_func(q, k, v, cu_d, cu_d, 64, 64, dropout_p=0.1, causal=True)
    with pytest.raises(RuntimeError):
        backend.flash_attn_varlen_func(q.float(), k, v, cu_d, cu_d, 64, 64, causal=True)
    with pytest.raises(RuntimeError):
        backend.flash_attn_varlen_func(q[:, :, :64], k[:, :, :64], v[:, :, :64], cu_d, cu_d, 64, 64, causal=True)
    with pytest.raises(RuntimeError):
        backend.flash_attn_varlen_func(q, k, v, cu_d, cu_d, 64, 64, causal=True, window_size=(128, 0))
