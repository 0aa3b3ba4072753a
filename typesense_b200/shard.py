"""Query-batch sharding across ranks: one process per GPU, a full index replica each, the multi_search
batch split into contiguous slices, and the per-query top-k records gathered on rank 0. No collective inside a query;
the only exchange is this gather (NCCL over NVLink on GPUs, gloo in the CPU tests). Harness-side plumbing."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .structs import KV_DTYPE


def shard_range(n_queries: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_queries, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_topk(kv: np.ndarray, cnt: np.ndarray, found: np.ndarray, n_queries: int, device: Optional[torch.device] = None):
    """kv [n_local, stride] KV records, cnt/found [n_local] of this rank's slice -> on rank 0 the arrays of the whole
    batch in query order, elsewhere None. Works with any backend (gloo on CPU, nccl with device tensors)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    stride = kv.shape[1]
    sizes = [shard_range(n_queries, world, r) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    dev = device if device is not None else torch.device("cpu")

    def pad(a: np.ndarray, rows: int) -> torch.Tensor:
        t = torch.zeros((rows,) + a.shape[1:], dtype=torch.uint8 if a.dtype == np.uint8 else torch.from_numpy(a[:0]).dtype, device=dev)
        t[: a.shape[0]] = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        return t

    kv_b = pad(kv.view(np.uint8).reshape(kv.shape[0], stride * KV_DTYPE.itemsize), max_n)
    cf = pad(np.stack([cnt.astype(np.int64), found.astype(np.int64)], 1), max_n)
    out_kv = [torch.empty_like(kv_b) for _ in range(world)] if rank == 0 else None
    out_cf = [torch.empty_like(cf) for _ in range(world)] if rank == 0 else None
    dist.gather(kv_b, out_kv, dst=0)
    dist.gather(cf, out_cf, dst=0)
    if rank != 0:
        return None
    kvs, cnts, founds = [], [], []
    for r, (lo, hi) in enumerate(sizes):
        n = hi - lo
        kvs.append(out_kv[r][:n].cpu().numpy().reshape(n, stride * KV_DTYPE.itemsize).view(KV_DTYPE).reshape(n, stride))
        c = out_cf[r][:n].cpu().numpy()
        cnts.append(c[:, 0].astype(np.uint32)); founds.append(c[:, 1].astype(np.uint32))
    return np.concatenate(kvs), np.concatenate(cnts), np.concatenate(founds)
