import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qserve_b200._lib import lib
from qserve_b200.decode import DecodeRunner
run = DecodeRunner("llama-3-8b", "w4a8kv4", 64, 1024, torch.device("cuda:0"), layers=8)
run.q_scale.fill_(0.01); run.q_sum.fill_(0.1)
for mode in (0, -1):
    lib.qs_gemm_force_tile_tokens(mode)
    for i in range(8):
        run.layers[i]["gate_up"](run.q_hidden, run.q_scale, run.q_sum, run.gate_up_buf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(64):
        run.layers[i % 8]["gate_up"](run.q_hidden, run.q_scale, run.q_sum, run.gate_up_buf)
    e1.record(); torch.cuda.synchronize()
    print("gate_up", "wide" if mode else "tiled", e0.elapsed_time(e1) * 1e3 / 64, "us")
lib.qs_gemm_force_tile_tokens(0)
