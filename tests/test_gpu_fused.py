"""GPU: the fused extensions are bit-identical to the reference op sequences they replace."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,H", [(64, 4096), (5, 8192), (3, 512)])
@pytest.mark.parametrize("with_sum", [True, False])
def test_add_rms_norm_general_equals_add_then_norm(dev, M, H, with_sum):
    from qserve_b200 import backend as ext
    import qserve_backend.layernorm_ops as ln
    g = torch.Generator(device="cpu").manual_seed(M + H)
    x = (torch.randn((M, H), generator=g) * 3).half().to(dev)
    delta = torch.randn((M, H), generator=g).half().to(dev)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).half().to(dev)
    q1 = torch.empty((M, H), dtype=torch.int8, device=dev); s1 = torch.empty(M, dtype=torch.half, device=dev); m1 = torch.zeros(M, dtype=torch.half, device=dev)
    q2 = torch.empty_like(q1); s2 = torch.empty_like(s1); m2 = torch.zeros_like(m1)
    hidden = x + delta
    if with_sum:
        ln.rms_norm_general_fuse_sum(q1, hidden, gamma, m1, s1, 1e-5, True)
    else:
        ln.rms_norm_general(q1, hidden, gamma, s1, 1e-5, True)
    h2 = torch.empty_like(x)
    ext.add_rms_norm_general(q2, h2, x, delta, gamma, m2 if with_sum else None, s2, 1e-5)
    torch.cuda.synchronize()
    assert torch.equal(h2, hidden) and torch.equal(q1, q2) and torch.equal(s1, s2) and torch.equal(m1, m2)


@pytest.mark.parametrize("M,d", [(64, 14336), (7, 1024), (2, 24576)])
@pytest.mark.parametrize("with_sum", [True, False])
def test_silu_and_mul_quant_equals_silu_then_quant(dev, M, d, with_sum):
    from qserve_b200 import backend as ext
    import qserve_backend.activation_ops as act
    import qserve_backend.fused_kernels as fk
    g = torch.Generator(device="cpu").manual_seed(M + d)
    x = (torch.randn((M, 2 * d), generator=g) * 2).half().to(dev)
    a = torch.empty((M, d), dtype=torch.half, device=dev)
    act.silu_and_mul(a, x)
    q1 = torch.empty((M, d), dtype=torch.int8, device=dev); s1 = torch.empty(M, dtype=torch.half, device=dev); m1 = torch.zeros(M, dtype=torch.half, device=dev)
    q2 = torch.empty_like(q1); s2 = torch.empty_like(s1); m2 = torch.zeros_like(m1)
    if with_sum:
        fk.invoke_quant_fuse_sum(q1, a, m1, s1)
    else:
        fk.invoke_quant(q1, a, s1)
    ext.silu_and_mul_quant(q2, x, m2 if with_sum else None, s2)
    torch.cuda.synchronize()
    assert torch.equal(q1, q2) and torch.equal(s1, s2) and torch.equal(m1, m2)


@pytest.mark.parametrize("precision", ["w4a8kv4", "w4a8kv4-g128", "w8a8kv8"])
def test_fused_runner_matches_reference_sequence(dev, precision):
    from qserve_b200.decode import DecodeRunner
    outs = []
    for fused in (False, True):
        run = DecodeRunner("tiny", precision, batch=6, ctx=100, device=dev, seed=5, fused=fused)
        run.tokens_in.copy_(torch.arange(6, device=dev) * 11)
        with torch.no_grad():
            outs.append(run.forward(run.tokens_in).clone())
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
