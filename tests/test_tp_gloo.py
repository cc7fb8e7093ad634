"""CPU, world_size = 2 (gloo): the tensor-parallel sharding rules on the packed checkpoint layout reproduce the
single-device INT32 accumulators exactly (column parallel: concatenation; row parallel: all-reduce sum)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import w4a8
from qserve_b200 import tp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # same data on every rank
    M, N, K = 8, 256, 512
    qc, qw, s1, s2s, s2z = w4a8.synth_per_group(rng, N, K)
    aq = rng.integers(-127, 128, size=(M, K), dtype=np.int8)
    w8 = w4a8.dequant_level2(qc, s2s, s2z)
    full = w4a8.int_matmul(aq, w8)
    # row parallel: split K, partial INT32 accumulators are all-reduced
    qw_r = tp.shard_rows(torch.from_numpy(qw), rank, world).numpy()
    s2s_r = tp.shard_level2_rows(torch.from_numpy(s2s), rank, world).numpy()
    s2z_r = tp.shard_level2_rows(torch.from_numpy(s2z), rank, world).numpy()
    k = K // world
    w8_r = w4a8.dequant_level2(w4a8.unpack_w4(qw_r), s2s_r, s2z_r)
    part = torch.from_numpy(w4a8.int_matmul(aq[:, rank * k:(rank + 1) * k], w8_r))
    dist.all_reduce(part)
    ok_row = bool(np.array_equal(part.numpy(), full))
    # column parallel: split N, results are concatenated
    qw_c = tp.shard_columns(torch.from_numpy(qw), rank, world).numpy()
    s2s_c = tp.shard_level2_columns(torch.from_numpy(s2s), rank, world).numpy()
    s2z_c = tp.shard_level2_columns(torch.from_numpy(s2z), rank, world).numpy()
    w8_c = w4a8.dequant_level2(w4a8.unpack_w4(qw_c), s2s_c, s2z_c)
    mine = torch.from_numpy(w4a8.int_matmul(aq, w8_c))
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    ok_col = bool(np.array_equal(torch.cat(parts, dim=1).numpy(), full))
    q.put((rank, ok_row, ok_col))
    dist.destroy_process_group()


def test_tp_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok_row and ok_col for _, ok_row, ok_col in res), res
