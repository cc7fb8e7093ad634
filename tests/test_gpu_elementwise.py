"""GPU parity: fused norm / activation / per-token quant kernels vs the CPU oracle, through the drop-in API.

Stated tolerances
  * INT8 outputs of per-token quantisers: bit-exact wherever the fp32 product x*(127/amax) is further than 2e-3 from a
    rounding boundary; elsewhere |diff| <= 1 LSB (the row mean / variance / sum are fp32 reductions whose association
    differs between the GPU tree and the oracle's float64 sum, moving y by ~1e-7 relative).  For invoke_quant (no
    reduction feeding the codes) the INT8 output and the scale are required to be bit-exact.
  * fp16 scales: bit-exact for invoke_quant, <= 1 fp16 ulp for the norm (amax of half(y)).
  * fp16 row sums: invoke_quant_fuse_sum bit-exact (the kernel accumulates the fp16 inputs in 2^-24 fixed point, i.e. the
    exact sum rounded once, like the oracle's float64 sum); the norm's sum within 2e-3*sqrt(H) (its addends half(y) inherit
    the 1-ulp freedom of y).
  * silu_and_mul: <= 2 fp16 ulp, < 0.1% of the elements differ at all (expf implementations differ in the last fp32 bit;
    a 1-ulp difference of half(silu) can become 2 ulp after the fp16 product).
"""
import numpy as np
import pytest
import torch

from oracle import ops
from tests.util import bits16, np_of, to_dev, ulp16_diff

pytestmark = pytest.mark.gpu


def _codes_close(q_gpu, q_ref, exact_product, name):
    d = np.abs(q_gpu.astype(np.int32) - q_ref.astype(np.int32))
    assert d.max() <= 1, f"{name}: int8 differs by more than 1 LSB"
    frac = exact_product - np.floor(exact_product)
    near = np.abs(frac - 0.5) < 2e-3
    assert not np.any((d > 0) & ~near), f"{name}: mismatch away from a rounding boundary"
    assert (d > 0).mean() < 2e-3, f"{name}: too many boundary flips ({(d > 0).mean():.2e})"


@pytest.mark.parametrize("M,H", [(1, 128), (64, 4096), (7, 14336), (300, 4096), (5, 8192), (3, 1000)])
@pytest.mark.parametrize("fuse_sum", [True, False])
def test_invoke_quant(dev, M, H, fuse_sum):
    import qserve_backend.fused_kernels as fk
    rng = np.random.default_rng(M + H)
    x = (rng.standard_normal((M, H)) * rng.uniform(0.1, 4, size=(M, 1))).astype(np.float16)
    q_o, s_o, sum_o = ops.quant_per_token(x, fuse_sum)
    xd = to_dev(x, dev)
    q = torch.empty((M, H), dtype=torch.int8, device=dev)
    s = torch.empty(M, dtype=torch.half, device=dev)
    if fuse_sum:
        sm = torch.empty(M, dtype=torch.half, device=dev)
        fk.invoke_quant_fuse_sum(q, xd, sm, s)
    else:
        fk.invoke_quant(q, xd, s)
    torch.cuda.synchronize()
    assert np.array_equal(np_of(q), q_o), "int8 codes must be bit-exact"
    assert np.array_equal(bits16(np_of(s)), bits16(s_o)), "scales must be bit-exact"
    if fuse_sum:
        assert np.array_equal(bits16(np_of(sm)), bits16(sum_o)), "row sums are exact (fixed-point accumulation) and must be bit-exact"


@pytest.mark.parametrize("M,H", [(1, 128), (64, 4096), (300, 4096), (5, 8192), (9, 512), (4, 1000)])
@pytest.mark.parametrize("fuse_sum", [True, False])
def test_rms_norm_general(dev, M, H, fuse_sum):
    import qserve_backend.layernorm_ops as ln
    rng = np.random.default_rng(M * 3 + H)
    x = (rng.standard_normal((M, H)) * 2 + rng.uniform(-1, 1, size=(M, 1))).astype(np.float16)
    gamma = (1 + 0.1 * rng.standard_normal(H)).astype(np.float16)
    eps = 1e-5
    q_o, s_o, sum_o, y = ops.layernorm_general_quant(x, gamma, eps, fuse_sum)
    q = torch.empty((M, H), dtype=torch.int8, device=dev)
    s = torch.empty(M, dtype=torch.half, device=dev)
    if fuse_sum:
        sm = torch.empty(M, dtype=torch.half, device=dev)
        ln.rms_norm_general_fuse_sum(q, to_dev(x, dev), to_dev(gamma, dev), sm, s, eps, True)
    else:
        ln.rms_norm_general(q, to_dev(x, dev), to_dev(gamma, dev), s, eps, True)
    torch.cuda.synchronize()
    assert ulp16_diff(np_of(s), s_o).max() <= 1
    amax = s_o.astype(np.float32) * 127
    _codes_close(np_of(q), q_o, y.astype(np.float64) * (127.0 / amax.astype(np.float64))[:, None], "rms_norm_general")
    if fuse_sum:
        got, want = np_of(sm).astype(np.float32), sum_o.astype(np.float32)
        # the sum of a mean-free row is a cancellation: compare against the magnitude of what is summed
        assert np.abs(got - want).max() <= 2e-3 * np.sqrt(H)


@pytest.mark.parametrize("M,H", [(64, 4096), (3, 8192), (2, 128)])
def test_rms_norm_final(dev, M, H):
    import qserve_backend.layernorm_ops as ln
    rng = np.random.default_rng(H)
    x = rng.standard_normal((M, H)).astype(np.float16)
    w = (1 + 0.1 * rng.standard_normal(H)).astype(np.float16)
    want = ops.rms_norm(x, w, 1e-6)
    out = torch.empty((M, H), dtype=torch.half, device=dev)
    ln.rms_norm(out, to_dev(x, dev), to_dev(w, dev), 1e-6)
    torch.cuda.synchronize()
    assert ulp16_diff(np_of(out), want).max() <= 1
    assert (ulp16_diff(np_of(out), want) > 0).mean() < 1e-2
    outq = torch.empty((M, H), dtype=torch.int8, device=dev)
    ln.rms_norm(outq, to_dev(x, dev), to_dev(w, dev), 1e-6, True)
    torch.cuda.synchronize()
    assert np.abs(np_of(outq).astype(np.int32) - ops.rms_norm(x, w, 1e-6, True).astype(np.int32)).max() <= 1


@pytest.mark.parametrize("M,d", [(64, 14336), (1, 128), (300, 1024), (5, 24576)])
def test_silu_and_mul(dev, M, d):
    import qserve_backend.activation_ops as act
    rng = np.random.default_rng(d)
    x = (rng.standard_normal((M, 2 * d)) * 2).astype(np.float16)
    want = ops.silu_and_mul(x)
    out = torch.empty((M, d), dtype=torch.half, device=dev)
    act.silu_and_mul(out, to_dev(x, dev))
    torch.cuda.synchronize()
    diff = ulp16_diff(np_of(out), want)
    assert diff.max() <= 2 and (diff > 0).mean() < 1e-3


def test_legacy_exports(dev):
    import qserve_backend.activation_ops as act
    import qserve_backend.fused_kernels as fk
    import qserve_backend.layernorm_ops as ln
    rng = np.random.default_rng(0)
    M, H = 6, 512
    x = rng.standard_normal((M, H)).astype(np.float16)
    for fn, ref in ((act.gelu_new, ops.gelu_new), (act.gelu_fast, ops.gelu_fast)):
        out = torch.empty((M, H), dtype=torch.half, device=dev)
        fn(out, to_dev(x, dev))
        assert np.abs(np_of(out).astype(np.float32) - ref(x).astype(np.float32)).max() <= 4e-3  # (1 + tanh) cancels near -1
    acc = rng.integers(-30000, 30000, size=(M, H)).astype(np.int32)
    res = rng.standard_normal((M, H)).astype(np.float16)
    sc = rng.uniform(1e-4, 1e-3, size=M).astype(np.float16)
    out = torch.empty((M, H), dtype=torch.half, device=dev)
    fk.invoke_dequant_add_residual(out, to_dev(acc, dev), to_dev(res, dev), to_dev(sc, dev))
    assert np.array_equal(bits16(np_of(out)), bits16(ops.dequant_add_residual(acc, res, sc)))
    fk.invoke_dequant_add_residual(out, to_dev(acc, dev), to_dev(res, dev), 0.00075)
    assert np.array_equal(bits16(np_of(out)), bits16(ops.dequant_add_residual(acc, res, np.float16(0.00075))))
    fk.invoke_dequant(out, to_dev(acc, dev), 0.00075)
    assert np.array_equal(bits16(np_of(out)), bits16(ops.dequant(acc, 0.00075)))
    q = torch.empty((M, H), dtype=torch.int8, device=dev)
    fk.invoke_quant(q, to_dev(x, dev), 0.05)
    assert np.array_equal(np_of(q), ops.quant_scalar_scale(x, 0.05))
    gam = (1 + 0.1 * rng.standard_normal(H)).astype(np.float16)
    rd = to_dev(res, dev)
    ln.invoke_dequant_add_residual_rms_norm_quant(q, to_dev(acc, dev), rd, to_dev(gam, dev), to_dev(sc, dev), 1e-5)
    q_o, res_o = ops.dequant_add_residual_rms_norm_quant(acc, res, gam, sc, 1e-5)
    assert np.array_equal(bits16(np_of(rd)), bits16(res_o))
    assert np.abs(np_of(q).astype(np.int32) - q_o.astype(np.int32)).max() <= 1
    acc2 = rng.integers(-3000, 3000, size=(M, 2 * H)).astype(np.int32)
    act.invoke_dequant_silu_and_mul_quant(q, to_dev(acc2, dev), 1e-3, 2e-3, 0.05)
    assert np.abs(np_of(q).astype(np.int32) - ops.dequant_silu_and_mul_quant(acc2, 1e-3, 2e-3, 0.05).astype(np.int32)).max() <= 1


def test_empty_and_errors(dev):
    import qserve_backend.fused_kernels as fk
    fk.invoke_quant(torch.empty((0, 128), dtype=torch.int8, device=dev), torch.empty((0, 128), dtype=torch.half, device=dev),
                    torch.empty(0, dtype=torch.half, device=dev))
    with pytest.raises(RuntimeError):  # hidden not a multiple of 8
        fk.invoke_quant(torch.empty((2, 100), dtype=torch.int8, device=dev), torch.zeros((2, 100), dtype=torch.half, device=dev),
                        torch.empty(2, dtype=torch.half, device=dev))
    with pytest.raises(RuntimeError):  # bf16 is not part of the models' path
        fk.invoke_quant(torch.empty((2, 128), dtype=torch.int8, device=dev), torch.zeros((2, 128), dtype=torch.bfloat16, device=dev),
                        torch.empty(2, dtype=torch.half, device=dev))
