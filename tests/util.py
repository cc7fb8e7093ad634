"""Shared helpers for the GPU parity tests (numpy oracle <-> torch CUDA tensors)."""
import numpy as np
import torch


def to_dev(a, dev):
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint8:
        return torch.from_numpy(a).to(dev)
    return torch.from_numpy(a).to(dev)


def np_of(t):
    return t.detach().cpu().numpy()


def bits16(a):
    return np.ascontiguousarray(a).view(np.uint16)


def ulp16_diff(a, b):
    """Distance in fp16 ulps between two fp16 arrays (sign-magnitude ordered)."""
    def key(x):
        u = bits16(np.asarray(x, np.float16)).astype(np.int32)
        return np.where(u & 0x8000, -(u & 0x7FFF), u & 0x7FFF)
    return np.abs(key(a) - key(b))


class GpuPool:
    """Device copy of an oracle PagePool plus the int64 address table the ops take."""

    def __init__(self, pool, dev):
        self.pool = pool
        self.t = torch.from_numpy(pool.data.copy()).to(dev)

    def table(self, block_tables):
        bt = np.asarray(block_tables, np.int64)
        return self.t.data_ptr() + bt * self.pool.pb

    def download(self):
        return self.t.cpu().numpy()


def kv_pointer_table(kpool: "GpuPool", vpool: "GpuPool", block_tables, dev):
    """[B, 2, max_blocks] int64 absolute addresses: row 0 = K pages, row 1 = V pages (kvCacheUtils.h:84-90)."""
    k = kpool.table(block_tables)
    v = vpool.table(block_tables)
    return torch.from_numpy(np.stack([k, v], axis=1)).to(dev)
