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
        got = torch.cat(outs)
        if feat > 64:      # the single-matrix kernel sums F = 68..128 in the half-warp pair order, the partitioned one serially
            mag = kernels.agg_scaled_sum(g.fwd_view(), x.abs(), norm, None, norm)
            assert bool(((got - full).abs() <= 2e-6 * mag).all()), feat
        else:
            assert torch.equal(got, full), feat


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


@pytest.mark.gpu
@pytest.mark.parametrize("feat", [100, 16, 192])
def test_row_subset_pass_accumulates_into_the_mapped_rows(cuda, feat):
    """stg_agg_scaled_sum_rows_f32: own-source pass writes all rows, the compacted second pass += only its rows."""
    from stgraph_b200 import _lib, kernels
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.utils import synthetic

    n = 20000
    src, dst = synthetic.power_law_graph(n, 400000, alpha=2.1, locality=0.5, window=512, max_degree=5000, seed=3, device=cuda)
    g = StaticGraph(torch.stack([src, dst], 1), None, n)
    F_ = g._forward_graph
    norm = g.degree_norm().reshape(-1).contiguous()
    x = torch.randn(n, feat, device=cuda)
    full = kernels.agg_scaled_sum(g.fwd_view(), x, norm, None, norm)
    mag = kernels.agg_scaled_sum(g.fwd_view(), x.abs(), norm, None, norm)
    ro = F_.row_offset.long()
    deg = ro[1:] - ro[:-1]
    rows = torch.repeat_interleave(torch.arange(n, device=cuda), deg)
    second = (F_.column_indices.long() % 7) == 0          # "remote" sources: a sparse subset of the edges, hub rows included
    keep = []

    def view(ro32, cols, n_rows, hub=False):
        v = _lib.StgCsrView()
        v.row_offset, v.column_indices, v.eids, v.node_ids = ro32.data_ptr(), cols.data_ptr(), None, None
        v.num_nodes, v.num_edges, v.eid_base, v.eids_identity = n_rows, int(cols.shape[0]), 0, 1
        v.hub_rows = v.hub_count = None
        v.hub_threshold = v.hub_capacity = 0
        if hub:
            cap = int(cols.shape[0]) // 64 + 1
            hr = torch.empty(cap, dtype=torch.int32, device=cuda)
            hc = torch.zeros(1, dtype=torch.int32, device=cuda)
            _lib.call("stg_csr_hub_rows", ro32.data_ptr(), n_rows, 64, hr.data_ptr(), cap, hc.data_ptr(), _lib.current_stream_ptr())
            assert int(hc.item()) > 0
            keep.append((hr, hc))
            v.hub_rows, v.hub_count, v.hub_threshold, v.hub_capacity = hr.data_ptr(), hc.data_ptr(), 64, cap
        keep.append((ro32, cols))
        return v

    cnt1 = torch.bincount(rows[~second], minlength=n)
    ro1 = torch.zeros(n + 1, dtype=torch.int32, device=cuda)
    ro1[1:] = torch.cumsum(cnt1, 0).int()
    cnt2 = torch.bincount(rows[second], minlength=n)
    sel = torch.nonzero(cnt2 > 0).reshape(-1)
    assert 0 < sel.numel() < n
    ro2 = torch.zeros(sel.numel() + 1, dtype=torch.int32, device=cuda)
    ro2[1:] = torch.cumsum(cnt2[sel], 0).int()
    v1 = view(ro1, F_.column_indices[~second].contiguous(), n)
    for hub in (False, True):
        v2 = view(ro2, F_.column_indices[second].contiguous(), int(sel.numel()), hub=hub)
        out = kernels.agg_scaled_sum(v1, x, norm, None, norm)
        kernels.agg_scaled_sum(v2, x, norm, None, norm, out=out, accumulate=True, out_rows=sel.int().contiguous())
        assert bool(((out - full).abs() <= 2e-6 * mag + 1e-30).all())
        # assign form touches only the mapped rows
        out2 = torch.full((n, feat), 7.0, device=cuda)
        kernels.agg_scaled_sum(v2, x, norm, None, norm, out=out2, out_rows=sel.int().contiguous())
        untouched = torch.ones(n, dtype=torch.bool, device=cuda)
        untouched[sel] = False
        assert bool((out2[untouched] == 7.0).all()) and bool((out2[sel] != 7.0).any())
