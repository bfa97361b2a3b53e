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


def _halo_worker(rank, world, port, n, src, dst, x, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stgraph_b200.dist.halo import HaloPlan
        from stgraph_b200.dist.partition import edge_balanced_bounds

        class Csr:          # CPU stand-in for the GPU CSR object (same fields)
            pass

        out = {}
        for tag, c in (("fwd", S.forward_csr(src, dst, n)), ("bwd", S.backward_csr(src, dst, n))):
            csr = Csr()
            csr.row_offset = torch.from_numpy(c.row_offset)
            csr.column_indices = torch.from_numpy(c.column_indices)
            csr.eids = torch.from_numpy(c.eids)
            csr.eids_identity = tag == "fwd"
            csr.eid_base = 0
            csr.num_edges = int(c.column_indices.shape[0])
            out[tag] = csr
        fb = edge_balanced_bounds(out["fwd"].row_offset, world)
        bb = edge_balanced_bounds(out["bwd"].row_offset, world)
        res = {}
        for tag, rb in (("fwd", fb), ("bwd", bb)):
            plan = HaloPlan(out[tag], rb, fb, rank, world)
            buf = plan.new_buffer(x.shape[1], x)
            buf[:plan.n_own] = x[plan.own_lo:plan.own_hi]
            buf[plan.n_own:] = float("nan")
            plan.exchange(buf)
            assert torch.equal(buf[plan.n_own:], x[plan.halo_ids])            # halo rows arrived, in id order
            w = torch.arange(out[tag].num_edges, dtype=torch.float32) * 0.001 + 0.5   # edge weights by GLOBAL eid
            local = A.scaled_sum(plan.local_row_offset.numpy(), plan.local_cols.numpy(), plan.local_eids.numpy(), buf,
                                 edge_scale=w)
            # two-pass form: edges to own sources + edges to halo sources == all edges of the rows
            own = x[plan.own_lo:plan.own_hi]
            halo = buf[plan.n_own:]
            two = A.scaled_sum(plan.own_ro.numpy(), plan.own_cols.numpy(), plan.own_eids.numpy(), own, edge_scale=w,
                               dtype=torch.float64)
            if plan.n_halo:
                two = two + A.scaled_sum(plan.halo_ro.numpy(), plan.halo_cols.numpy(), plan.halo_eids.numpy(), halo,
                                         edge_scale=w, dtype=torch.float64)
            assert torch.allclose(two, local, rtol=1e-5, atol=1e-6)
            assert int(plan.own_cols.shape[0]) + int(plan.halo_cols.shape[0]) == plan.num_local_edges
            # compacted halo view: only rows with a remote neighbour, same edges, rows strictly increasing
            hdeg = plan.halo_ro[1:] - plan.halo_ro[:-1]
            assert torch.equal(plan.halo_out_rows.long(), torch.nonzero(hdeg > 0).reshape(-1))
            cdeg = plan.halo_compact_ro[1:] - plan.halo_compact_ro[:-1]
            assert torch.equal(cdeg, hdeg[plan.halo_out_rows.long()]) and int(plan.halo_compact_ro[-1]) == plan.halo_cols.shape[0]
            if plan.n_halo:
                comp = A.scaled_sum(plan.halo_compact_ro.numpy(), plan.halo_cols.numpy(), plan.halo_eids.numpy(), halo,
                                    edge_scale=w, dtype=torch.float64)
                wide = A.scaled_sum(plan.halo_ro.numpy(), plan.halo_cols.numpy(), plan.halo_eids.numpy(), halo,
                                    edge_scale=w, dtype=torch.float64)
                assert torch.equal(comp, wide[plan.halo_out_rows.long()])
            res[tag] = (plan.row_lo, plan.row_hi, local, plan.n_halo)
        ret[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_plan_reproduces_single_process_result(world):
    rng = np.random.default_rng(3)
    n, e = 400, 5000
    # banded graph (neighbours mostly within +-30 ids) plus 5 % random long edges: small halos
    s = rng.integers(0, n, e)
    d = np.clip(s + rng.integers(-30, 31, e), 0, n - 1)
    far = rng.random(e) < 0.05
    d[far] = rng.integers(0, n, far.sum())
    k = np.unique(s.astype(np.int64) * n + d)
    k = k[(k // n) != (k % n)]
    src, dst = (k // n).astype(np.int32), (k % n).astype(np.int32)
    x = torch.randn(n, 6, generator=torch.Generator().manual_seed(2))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_halo_worker, args=(world, _free_port(), n, src, dst, x, ret), nprocs=world, join=True)
    for tag, c in (("fwd", S.forward_csr(src, dst, n)), ("bwd", S.backward_csr(src, dst, n))):
        w = torch.arange(c.column_indices.shape[0], dtype=torch.float32) * 0.001 + 0.5
        ref = A.scaled_sum(c.row_offset, c.column_indices, c.eids, x, edge_scale=w)
        rows = 0
        for r in range(world):
            lo, hi, local, n_halo = ret[r][tag]
            assert torch.equal(local, ref[lo:hi]), (tag, r)
            rows += hi - lo
            assert n_halo < n - (hi - lo)               # strictly fewer rows than a full all-gather
        assert rows == n
