"""The reference's UNMODIFIED Python model code running over this repo's `qserve_backend` (VERDICT r1 row b2,
north_star: "so qserve/modeling ... drop in unchanged").

`qserve_b200.refmodel.RefModel` imports `qserve.modeling.models.llama_w4a8_unpad.LlamaForCausalLM` (and with it the
reference's W4A8 linear / RMSNormGeneral / SiluAndMulQuant / InputMetadata / ActivationBuffer classes) from the
reference package under `baseline/_ref` (installed by `__graft_entry__.build()`), shares a DecodeRunner's synthetic
weights and KV pages with it, and drives it as ModelRunner does.  Bars:
  * decode logits through the reference code == `DecodeRunner._forward_reference` (this repo's restatement of the
    op sequence) BIT-identical, and == the fused CUDA-graph path's greedy tokens;
  * prefill (apply_bias_rope_update_kv_cache + flash_attn_varlen_func inside the reference layer) followed by decode
    (single_query_attention over the pages the prefill wrote) is self-consistent.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_or_skip():
    from qserve_b200 import refmodel

    if refmodel.locate_reference() is None:
        pytest.skip("reference Python package not present (baseline/_ref is created by __graft_entry__.build() where /root/reference exists)")
    try:
        import flash_attn  # noqa: F401  (imported at module scope by llama_w4a8_unpad.py:27)
    except Exception as e:  # pragma: no cover
        pytest.skip(f"flash_attn not importable: {e}")
    return refmodel


@pytest.mark.parametrize("precision", ["w4a8kv4", "w4a8kv4-g128", "w4a8kv8", "w8a8kv8"])
def test_reference_decoder_layers_bit_identical_to_runner(dev, precision):
    refmodel = _ref_or_skip()
    from qserve_b200.decode import DecodeRunner

    run = DecodeRunner("tiny", precision, batch=6, ctx=150, device=dev, seed=5, fused=False)
    ref = refmodel.RefModel(run)
    # the class that ran is the reference's own, bound to this repo's backend
    assert type(ref.model).__module__.startswith("qserve.modeling.models.llama_w")
    import qserve_backend
    assert qserve_backend.__file__.startswith(refmodel._ROOT)
    tokens = torch.randint(0, run.cfg.vocab, (run.batch,), device=dev)
    # both paths append the new token's K/V to the pages: snapshot the pools so that each sees the same cache
    snap = [(k.clone(), v.clone()) for k, v in zip(run.kpools, run.vpools)]
    logits_ref = ref.decode_logits(tokens).clone()
    torch.cuda.synchronize()
    pages_ref = [(k.clone(), v.clone()) for k, v in zip(run.kpools, run.vpools)]
    for (k, v), (k0, v0) in zip(zip(run.kpools, run.vpools), snap):
        k.copy_(k0); v.copy_(v0)
    with torch.no_grad():
        logits_run = run._forward_reference(tokens, return_logits=True)
    torch.cuda.synchronize()
    assert logits_ref.shape == (run.batch, run.cfg.vocab)
    assert torch.isfinite(logits_ref.float()).all()
    assert torch.equal(logits_ref, logits_run), f"max |diff| {float((logits_ref.float() - logits_run.float()).abs().max())}"
    for (k, v), (kr, vr) in zip(zip(run.kpools, run.vpools), pages_ref):
        assert torch.equal(k, kr) and torch.equal(v, vr)
    # the fused + CUDA-graph path decodes the same greedy tokens
    for (k, v), (k0, v0) in zip(zip(run.kpools, run.vpools), snap):
        k.copy_(k0); v.copy_(v0)
    run.fused = True
    with torch.no_grad():
        tok_fused = run.forward(tokens)
    torch.cuda.synchronize()
    assert torch.equal(tok_fused, torch.argmax(logits_ref, dim=-1))
    assert torch.equal(ref.model.sample(tokens, logits_ref, ref.decode_metadata()), tok_fused)


def test_reference_model_forward_is_cuda_graph_capturable(dev):
    """The reference model code over this backend captures into a CUDA graph as is (ops run on the current stream,
    no host sync, allocations go to the graph pool): replay reproduces the eager logits bit for bit."""
    refmodel = _ref_or_skip()
    from qserve_b200.decode import DecodeRunner

    run = DecodeRunner("tiny", "w4a8kv4", batch=4, ctx=100, device=dev, seed=2, fused=False)
    ref = refmodel.RefModel(run)
    tokens = torch.randint(0, run.cfg.vocab, (run.batch,), device=dev)
    snap = [(k.clone(), v.clone()) for k, v in zip(run.kpools, run.vpools)]
    eager = ref.decode_logits(tokens).clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ref.decode_logits(tokens)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = ref.decode_logits(tokens)
    for (k, v), (k0, v0) in zip(zip(run.kpools, run.vpools), snap):
        k.copy_(k0); v.copy_(v0)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)


@pytest.mark.parametrize("precision,tol", [("w4a8kv8", 0.10), ("w4a8kv4-g128", 0.50)])
def test_reference_prefill_then_decode_consistent(dev, precision, tol):
    """Config-3 shape of work at test size: prompts go through the reference prefill path (in-place RoPE + KV quant/append by
    `apply_bias_rope_update_kv_cache`, attention by flash_attn on the fp16 q/k/v), then one decode step reads those pages.
    Consistency: prefilling P tokens gives (up to KV quantisation noise, hence the KV8 / KV4 tolerances) the same last-token
    logits as prefilling P-1 tokens and decoding token P over the cache."""
    refmodel = _ref_or_skip()
    from qserve_b200.decode import DecodeRunner

    lens = [97, 64, 130, 33]
    run = DecodeRunner("tiny", precision, batch=len(lens), ctx=191, device=dev, seed=9, fused=False)
    ref = refmodel.RefModel(run)
    g = torch.Generator(device="cpu").manual_seed(0)
    prompts = [torch.randint(0, run.cfg.vocab, (n,), generator=g) for n in lens]
    full = torch.cat(prompts).to(dev)
    logits_full = ref.prefill_logits(full, lens).float()
    torch.cuda.synchronize()
    assert logits_full.shape == (len(lens), run.cfg.vocab) and torch.isfinite(logits_full).all()
    # now P-1 tokens by prefill, token P by decode
    short = [n - 1 for n in lens]
    ref.prefill_logits(torch.cat([p[:-1] for p in prompts]).to(dev), short)
    run.context_lens.copy_(torch.tensor(lens, dtype=torch.int32))
    run.max_seq_len = max(lens)
    last = torch.stack([p[-1] for p in prompts]).to(dev)
    logits_dec = ref.decode_logits(last, fresh_metadata=True).float()
    torch.cuda.synchronize()
    assert torch.isfinite(logits_dec).all()
    rel = float((logits_dec - logits_full).norm() / logits_full.norm())
    print(f"[refmodel prefill/decode consistency] {precision}: relative L2 {rel:.4f}")
    assert rel < tol, f"prefill vs prefill+decode logits differ by {rel:.3f} (relative L2)"
    cos = torch.nn.functional.cosine_similarity(logits_dec, logits_full, dim=-1)
    assert float(cos.min()) > 1.0 - tol


def test_reference_prefill_over_this_repos_prompt_attention(dev):
    """The reference layer's `flash_attn_varlen_func(...)` call (llama_w4a8_unpad.py:232-242) served by qs_prefill_attention instead of the
    flash-attn package.  The last-token logits of a 4-bit model amplify the fp16 rounding of the attention output (INT8 re-quantisation of
    every activation), so the bar is relative: against the logits obtained with an EXACT (float32 softmax) attention plugged into the same
    call site, this repo's kernel may not be further away than 1.5 x flash-attn's own distance + 1e-3."""
    refmodel = _ref_or_skip()
    from qserve_b200.decode import DecodeRunner

    lens = [130, 64, 257, 1]
    run = DecodeRunner("tiny", "w4a8kv4", batch=len(lens), ctx=300, device=dev, seed=4, fused=False)
    ref = refmodel.RefModel(run)
    g = torch.Generator(device="cpu").manual_seed(1)
    toks = torch.cat([torch.randint(0, run.cfg.vocab, (n,), generator=g) for n in lens]).to(dev)

    def exact_attention(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, dropout_p=0.0, causal=True):
        out = torch.empty_like(q)
        grp = q.size(1) // k.size(1)
        cu = cu_seqlens_q.tolist()
        for b in range(len(cu) - 1):
            s = slice(cu[b], cu[b + 1])
            qq, kk, vv = (t[s].float().transpose(0, 1) for t in (q, k.repeat_interleave(grp, dim=1), v.repeat_interleave(grp, dim=1)))
            out[s] = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=True).transpose(0, 1).half()
        return out

    ref.use_prefill_attention("flash_attn")
    a = ref.prefill_logits(toks, lens).float()
    ref.use_prefill_attention("qserve_b200")
    b = ref.prefill_logits(toks, lens).float()
    ref.mod.flash_attn_varlen_func = exact_attention
    c = ref.prefill_logits(toks, lens).float()
    ref.use_prefill_attention("qserve_b200")
    torch.cuda.synchronize()
    assert torch.isfinite(b).all()
    d_ours, d_fa = float((b - c).norm() / c.norm()), float((a - c).norm() / c.norm())
    print(f"[refmodel prefill] relative L2 of the logits against exact attention: qserve_b200 {d_ours:.2e}, flash_attn {d_fa:.2e}")
    assert d_ours <= 1.5 * d_fa + 1e-3
    assert float(torch.nn.functional.cosine_similarity(b, c, dim=-1).min()) > 0.99
