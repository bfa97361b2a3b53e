"""Row-partitioned aggregation on ONE GPU: the P local slices, run one after another, reproduce the full result."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_slices_reproduce_full_aggregation(cuda, world):
    from stgraph_b200 import kernels
    from stgraph_b200.dist import PartitionedGraph
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.utils import synthetic

    n = 20000
    src, dst = synthetic.power_law_graph(n, 400000, alpha=2.1, locality=0.5, window=256, max_degree=5000, seed=4, device=cuda)
    g = StaticGraph(torch.stack([src, dst], 1), None, n)
    norm = g.degree_norm().reshape(-1).contiguous()
    x = torch.randn(n, 100, device=cuda)
    w = torch.rand(src.shape[0], device=cuda) + 0.1
    full_f = kernels.agg_scaled_sum(g.fwd_view(), x, norm, w, norm)
    full_b = kernels.agg_scaled_sum(g.bwd_view(), x, norm, w, norm)
    got_f, got_b, edges = [], [], []
    for rank in range(world):
        pg = PartitionedGraph(g, rank, world)
        lo, hi = pg.local_rows("fwd")
        out = torch.empty(hi - lo, 100, device=cuda)
        kernels.agg_scaled_sum(pg.fwd.view, x, norm, w, norm[lo:hi], out=out)
        got_f.append(out)
        edges.append(pg.fwd.num_local_edges)
        lo, hi = pg.local_rows("bwd")
        out = torch.empty(hi - lo, 100, device=cuda)
        kernels.agg_scaled_sum(pg.bwd.view, x, norm, w, norm[lo:hi], out=out)
        got_b.append(out)
    assert torch.equal(torch.cat(got_f), full_f)          # same kernel, same row order: bit-identical
    assert torch.equal(torch.cat(got_b), full_b)
    assert sum(edges) == src.shape[0]
    assert max(edges) <= 1.2 * (src.shape[0] / world) + 5000   # edge-balanced up to one hub row


@pytest.mark.parametrize("world", [2, 5, 16])
def test_partitioned_source_blocks_reproduce_full_aggregation(cuda, world):
    """The peer-memory kernel variant with the blocks placed in separate allocations of ONE GPU."""
    from stgraph_b200 import kernels
    from stgraph_b200.dist import PartitionedGraph
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.utils import synthetic

    n = 12000
    src, dst = synthetic.power_law_graph(n, 240000, alpha=2.1, locality=0.6, window=128, max_degree=4000, seed=6, device=cuda)
    g = StaticGraph(torch.stack([src, dst], 1), None, n)
    norm = g.degree_norm().reshape(-1).contiguous()
    for feat in (100, 16, 7):
        x = torch.randn(n, feat, device=cuda)
        full = kernels.agg_scaled_sum(g.fwd_view(), x, norm, None, norm)
        pg0 = PartitionedGraph(g, 0, world)
        blocks = [x[pg0.fwd_bounds[q]:pg0.fwd_bounds[q + 1]].clone() for q in range(world)]   # separate allocations
        ptrs = [b.data_ptr() if b.numel() else 0 for b in blocks]
        outs = []
        for rank in range(world):
            pg = PartitionedGraph(g, rank, world)
            lo, hi = pg.local_rows("fwd")
            out = torch.empty(hi - lo, feat, device=cuda)
            kernels.agg_scaled_sum_parts(pg.fwd.view, ptrs, pg.fwd_bounds, feat, norm, None, norm[lo:hi].contiguous(), out=out)
            outs.append(out)
        assert torch.equal(torch.cat(outs), full), feat


def test_two_concurrent_red_passes_are_deterministic_and_correct(cuda):
    """Split every row's edges in two CSRs, run both passes concurrently with red.global.add into a zeroed buffer."""
    from stgraph_b200 import _lib, kernels
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.utils import synthetic

    n = 30000
    src, dst = synthetic.power_law_graph(n, 600000, alpha=2.1, locality=0.5, window=512, max_degree=6000, seed=8, device=cuda)
    g = StaticGraph(torch.stack([src, dst], 1), None, n)
    F_ = g._forward_graph
    norm = g.degree_norm().reshape(-1).contiguous()
    x = torch.randn(n, 100, device=cuda)
    full = kernels.agg_scaled_sum(g.fwd_view(), x, norm, None, norm)
    mag = kernels.agg_scaled_sum(g.fwd_view(), x.abs(), norm, None, norm)
    ro = F_.row_offset.long()
    rows = torch.repeat_interleave(torch.arange(n, device=cuda), ro[1:] - ro[:-1])
    mask = (torch.arange(rows.shape[0], device=cuda) % 3) != 0

    def sub(m):
        cnt = torch.bincount(rows[m], minlength=n)
        sro = torch.zeros(n + 1, dtype=torch.int32, device=cuda)
        sro[1:] = torch.cumsum(cnt, 0).int()
        cols = F_.column_indices[m].contiguous()
        v = _lib.StgCsrView()
        v.row_offset, v.column_indices, v.eids, v.node_ids = sro.data_ptr(), cols.data_ptr(), None, None
        v.num_nodes, v.num_edges, v.eid_base, v.eids_identity = n, int(cols.shape[0]), 0, 1
        v.hub_rows = v.hub_count = None
        v.hub_threshold = v.hub_capacity = 0
        return v, (sro, cols)

    va, ka = sub(mask)
    vb, kb = sub(~mask)
    side = torch.cuda.Stream()
    results = []
    for _ in range(4):
        out = torch.zeros(n, 100, device=cuda)
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            kernels.agg_scaled_sum(vb, x, norm, None, norm, out=out, accumulate="red", stream=side.cuda_stream)
        kernels.agg_scaled_sum(va, x, norm, None, norm, out=out, accumulate="red")
        cur.wait_stream(side)
        torch.cuda.synchronize()
        results.append(out)
    for r in results[1:]:
        assert torch.equal(r, results[0])                 # two addends per element: order-independent
    assert bool(((results[0] - full).abs() <= 2e-6 * mag + 1e-30).all())
    # read-modify-write accumulate form, sequential
    out = kernels.agg_scaled_sum(va, x, norm, None, norm)
    kernels.agg_scaled_sum(vb, x, norm, None, norm, out=out, accumulate=True)
    assert torch.equal(out, results[0])
