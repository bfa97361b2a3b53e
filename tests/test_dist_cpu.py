"""world_size-2 gloo test (CPU) of the multi-GPU host logic: edge-balanced row partition + uneven row exchange.

The local aggregation is played by the oracle here (no GPU): what is checked is that partition bounds,
the per-owner broadcasts and the row-offset slicing reproduce the single-process result exactly.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import aggregate as A
from oracle import structure as S


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, src, dst, x, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stgraph_b200.dist.partition import edge_balanced_bounds, exchange_rows

        f = S.forward_csr(src, dst, n)
        bounds = edge_balanced_bounds(torch.from_numpy(f.row_offset), world)
        lo, hi = bounds[rank], bounds[rank + 1]
        # every rank starts with only its own block of the feature matrix
        full = torch.zeros_like(x)
        full[lo:hi] = x[lo:hi]
        exchange_rows(full, bounds)
        assert torch.equal(full, x)
        # local rows of the CSR: offset row pointer, global column ids
        ro = f.row_offset[lo:hi + 1]
        local = A.scaled_sum(ro - ro[0], f.column_indices[ro[0]:ro[-1]], None, full)
        ret[rank] = (lo, hi, local)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_exchange_reproduce_single_process_result(world):
    rng = np.random.default_rng(0)
    n, e = 300, 4000
    key = rng.choice(n * n, size=e, replace=False)
    src, dst = (key // n).astype(np.int32), (key % n).astype(np.int32)
    dst[:600] = 7                                   # a hub so that edge balance != vertex balance
    k = np.unique(src.astype(np.int64) * n + dst)
    src, dst = (k // n).astype(np.int32), (k % n).astype(np.int32)
    x = torch.randn(n, 12, generator=torch.Generator().manual_seed(1))
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, src, dst, x, ret), nprocs=world, join=True)
    f = S.forward_csr(src, dst, n)
    ref = A.scaled_sum(f.row_offset, f.column_indices, None, x)
    covered = 0
    edges = []
    for r in range(world):
        lo, hi, local = ret[r]
        assert torch.equal(local, ref[lo:hi])
        covered += hi - lo
        edges.append(int(f.row_offset[hi] - f.row_offset[lo]))
    assert covered == n
    assert max(edges) - min(edges) <= 700           # balanced by edges up to one (hub) row


def test_bounds_are_monotone_and_cover():
    from stgraph_b200.dist.partition import edge_balanced_bounds

    ro = torch.tensor([0, 0, 0, 10, 10, 11, 50, 50, 51], dtype=torch.int32)
    for p in (1, 2, 4, 8, 16):
        b = edge_balanced_bounds(ro, p)
        assert b[0] == 0 and b[-1] == 8 and len(b) == p + 1
        assert all(b[i] <= b[i + 1] for i in range(p))
