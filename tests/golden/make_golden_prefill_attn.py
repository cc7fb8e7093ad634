#!/usr/bin/env python
"""Golden vector for the prompt-phase attention: the output of flash-attn itself (the third-party kernel the reference calls at
llama_w4a8_unpad.py:232-242) on a small ragged batch.  Needs a GPU and the image's flash_attn; run on the GPU box:
    python tests/golden/make_golden_prefill_attn.py gpurun_out/prefill_attn_flash.npz
and copy the result to tests/golden/.  tests/test_oracle_prefill_attention.py pins oracle/prefill_attention.py against it."""
import sys

import flash_attn
import numpy as np
import torch
from flash_attn import flash_attn_varlen_func

lens, hq, hkv = [40, 130, 3], 4, 2
T = sum(lens)
g = torch.Generator(device="cpu").manual_seed(2024)
qkv = torch.randn(T, (hq + 2 * hkv) * 128, generator=g).half().cuda()
q, k, v = qkv.split([hq * 128, hkv * 128, hkv * 128], dim=-1)
q, k, v = q.reshape(T, hq, 128), k.reshape(T, hkv, 128), v.reshape(T, hkv, 128)
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
out = flash_attn_varlen_func(q, k, v, cu_seqlens_q=cu, cu_seqlens_k=cu, max_seqlen_q=max(lens), max_seqlen_k=max(lens), dropout_p=0.0, causal=True)
np.savez_compressed(sys.argv[1], q=q.cpu().numpy(), k=k.cpu().numpy(), v=v.cpu().numpy(), cu_seqlens=cu.cpu().numpy(), out=out.cpu().numpy(),
                    flash_attn_version=np.array(flash_attn.__version__))
print("wrote", sys.argv[1], "flash_attn", flash_attn.__version__)
