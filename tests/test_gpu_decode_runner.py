"""GPU: the decode-step runner (reference op sequence over the drop-in API) -- eager and CUDA-graph replay agree bit for
bit, PDL on/off agree bit for bit, outputs are finite, for every precision the BASELINE configs name."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["w4a8kv4", "w4a8kv4-g128", "w8a8kv8"])
def test_tiny_model_eager_vs_graph_vs_pdl(dev, precision):
    from qserve_b200 import backend
    from qserve_b200.decode import DecodeRunner

    outs = []
    for pdl in (True, False):
        backend.set_pdl(pdl)
        run = DecodeRunner("tiny", precision, batch=5, ctx=130, device=dev, seed=3)
        run.tokens_in.copy_(torch.arange(5, device=dev) * 7)
        with torch.no_grad():
            eager = run.forward(run.tokens_in).clone()
            eager2 = run.forward(run.tokens_in).clone()  # idempotent KV append: same slot, same bytes
        run.capture()
        run.step(); run.step()
        torch.cuda.synchronize()
        assert torch.equal(eager, eager2)
        assert torch.equal(eager, run.tokens_out)
        assert int(eager.min()) >= 0 and int(eager.max()) < run.cfg.vocab
        outs.append(eager)
    backend.set_pdl(True)
    assert torch.equal(outs[0], outs[1])


def test_hidden_state_is_finite(dev):
    from qserve_b200.decode import DecodeRunner

    run = DecodeRunner("llama-3-8b", "w4a8kv4", batch=8, ctx=100, device=dev, layers=4)
    with torch.no_grad():
        tok = run.forward(run.tokens_in)
    torch.cuda.synchronize()
    assert tok.shape == (8,)
    assert torch.isfinite(run.out_buf.float()).all()


@pytest.mark.parametrize("tp_size", [4, 8])
def test_one_rank_of_a_tensor_parallel_72b_layer_runs(dev, tp_size):
    """BASELINE config 5 shapes on ONE GPU: rank 3's shard of a Qwen1.5-72B layer at TP = 4 / 8 without the collectives (`no_comm`).  At TP = 8 the
    sharded qkv (3072) and gate_up (6144) rows are NARROWER than the hidden row (8192): the activation buffer must be sized by the widest of the three
    (the first 8-GPU run of the bench died here).  Fused graph path and reference op sequence, finite outputs."""
    from qserve_b200.decode import DecodeRunner

    for fused in (True, False):
        run = DecodeRunner("qwen1.5-72b", "w4a8kv4", batch=8, ctx=130, device=dev, layers=1, seed=5, tp_rank=3, tp_size=tp_size, no_comm=True, fused=fused)
        with torch.no_grad():
            tok = run.forward(run.tokens_in).clone()
        if fused:
            run.capture()
            run.step()
            torch.cuda.synchronize()
            assert torch.equal(tok, run.tokens_out)
        torch.cuda.synchronize()
        assert tok.shape == (8,) and int(tok.min()) >= 0 and int(tok.max()) < run.cfg.vocab
        assert torch.isfinite(run.out_buf.float()).all()
        del run
        torch.cuda.empty_cache()


@pytest.mark.parametrize("precision,group", [("w4a8kv4", -1), ("w4a8kv4-g128", 128)])
def test_converted_checkpoint_loads_and_runs(dev, precision, group):
    """fake-quant checkpoint -> convert -> fuse -> load_into_runner -> the runner's qkv GEMM equals the oracle GEMM on the
    converted buffers (bit-exact), and a decode step runs."""
    import numpy as np
    from oracle import ops, w4a8
    from qserve_b200 import checkpoint as ck
    from qserve_b200.decode import MODELS, DecodeRunner
    from tests.test_checkpoint import _fake_checkpoint
    from tests.util import bits16, np_of, to_dev
    cfg = MODELS["tiny"]
    fake, params = _fake_checkpoint(cfg.layers, cfg.hidden, cfg.intermediate, cfg.heads, cfg.kv_heads, cfg.head_dim, group, seed=3)
    fake["model.embed_tokens.weight"] = torch.randn(cfg.vocab, cfg.hidden).half()
    fake["lm_head.weight"] = torch.randn(cfg.vocab, cfg.hidden).half()
    sd = ck.convert_fake_quant_checkpoint(fake, params, cfg.layers, 4, group)
    fused = ck.fuse_llama_state_dict(sd, cfg.layers)
    run = DecodeRunner("tiny", precision, batch=4, ctx=70, device=dev, seed=1)
    n = ck.load_into_runner(run, fused)
    assert n >= cfg.layers * 4 * 3 + 2  # embed + lm_head; the W4A8 loader skips the norms
    pre = "model.layers.0.self_attn.qkv_proj."
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, cfg.hidden)).astype(np.float16)
    aq, sa, asum = ops.quant_per_token(x)
    out = torch.empty((4, fused[pre + "qweight"].size(0)), dtype=torch.half, device=dev)
    run.layers[0]["qkv"](to_dev(aq, dev), to_dev(sa, dev), to_dev(asum, dev), out)
    torch.cuda.synchronize()
    if group == -1:
        want = w4a8.gemm_w4a8_per_chn(aq, fused[pre + "qweight"].numpy(), fused[pre + "s1_scales"].numpy(), sa, fused[pre + "s1_szeros"].numpy(), asum)
    else:
        want = w4a8.gemm_w4a8_per_group(aq, fused[pre + "qweight"].numpy(), fused[pre + "s2_zeros"].numpy(), fused[pre + "s2_scales"].numpy(),
                                        fused[pre + "s1_scales"].numpy(), sa)
    assert np.array_equal(bits16(np_of(out)), bits16(want))
    with torch.no_grad():
        tok = run.forward(run.tokens_in)
    torch.cuda.synchronize()
    assert tok.shape == (4,) and int(tok.min()) >= 0 and int(tok.max()) < cfg.vocab
