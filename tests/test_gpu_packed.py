"""GPU parity of the packed-metadata aggregation path (``stg_csr_pack_edge_meta_f32`` + ``stg_agg_packed_sum_f32``).

The packed kernel must give the sums of the plain kernel BIT FOR BIT (same products, same order), and both must
agree with the torch-CPU index_add oracle to rel 1e-5 (BASELINE.json north_star; SURVEY.md trap T9).
"""
import numpy as np
import pytest
import torch

from oracle import aggregate as A
from oracle import structure as S

pytestmark = pytest.mark.gpu


def _graph(n, e, seed, hub=None):
    from test_gpu_agg import _rand_graph

    return _rand_graph(n, e, seed, hub)


def _inputs(n, e_count, feat, seed, weighted, cuda):
    tg = torch.Generator().manual_seed(seed)
    x = torch.randn(n, feat, generator=tg)
    norm = torch.rand(n, generator=tg) + 0.5
    w = (torch.rand(e_count, generator=tg) + 0.1) if weighted else None
    dev = lambda t: None if t is None else t.to(cuda)
    return x, norm, w, dev


@pytest.mark.parametrize("feat", [1, 2, 3, 4, 7, 16, 20, 47, 64, 100, 128, 200, 256, 300, 600])
@pytest.mark.parametrize("weighted", [False, True])
def test_packed_equals_plain_and_oracle(cuda, feat, weighted):
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph

    n, e = 400, 5000
    src, dst = _graph(n, e, seed=feat + 1)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    x, norm, w, dev = _inputs(n, src.shape[0], feat, 7 * feat + weighted, weighted, cuda)
    f, b = S.forward_csr(src, dst, n), S.backward_csr(src, dst, n)
    for view, csr in ((g.fwd_view(), f), (g.bwd_view(), b)):
        plain = kernels.agg_scaled_sum(view, dev(x), dev(norm), dev(w), dev(norm))
        meta = kernels.pack_edge_meta(view, dev(norm), dev(w))
        assert meta.shape == (view.num_edges, 2) and meta.dtype == torch.int32
        packed = kernels.agg_packed_sum(view, meta, dev(x), dev(norm))
        assert torch.equal(plain, packed), f"packed and plain kernels differ (F={feat})"
        ref = A.scaled_sum(csr.row_offset, csr.column_indices, csr.eids, x, norm, w, norm)
        mag = A.scaled_sum(csr.row_offset, csr.column_indices, csr.eids, x.abs(), norm, w, norm)
        A.assert_close_rel(packed.cpu(), ref, rel=1e-5, abs_terms=mag, what=f"packed agg F={feat}")


def test_pack_contents_bit_exact(cuda):
    """meta[e] = {col[e], ns[col[e]] * es[eid[e]]} for both directions (identity and permuted eids)."""
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph

    n = 300
    src, dst = _graph(n, 4000, seed=3)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    _, norm, w, dev = _inputs(n, src.shape[0], 4, 11, True, cuda)
    for view, csr in ((g.fwd_view(), S.forward_csr(src, dst, n)), (g.bwd_view(), S.backward_csr(src, dst, n))):
        meta = kernels.pack_edge_meta(view, dev(norm), dev(w)).cpu()
        col = torch.as_tensor(np.asarray(csr.column_indices)).long()
        eid = torch.as_tensor(np.asarray(csr.eids)).long()
        assert torch.equal(meta[:, 0].long(), col)
        exp = (norm[col] * w[eid]).to(torch.float32)
        assert torch.equal(meta[:, 1].contiguous().view(torch.float32), exp)
        only_w = kernels.pack_edge_meta(view, None, dev(w)).cpu()
        assert torch.equal(only_w[:, 1].contiguous().view(torch.float32), w[eid])


@pytest.mark.parametrize("feat", [4, 16, 100, 7])
def test_packed_hub_rows(cuda, feat):
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.graph.static import csr as csr_mod

    n = 12000
    src, dst = _graph(n, 30000, seed=21, hub=9 * csr_mod.HUB_THRESHOLD)   # > 8192 edges: cluster-wide hub tier too
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    assert g.fwd_view().hub_threshold > 0
    x, norm, w, dev = _inputs(n, src.shape[0], feat, 5, True, cuda)
    f = S.forward_csr(src, dst, n)
    for view in (g.fwd_view(), g.bwd_view()):
        plain = kernels.agg_scaled_sum(view, dev(x), dev(norm), dev(w), dev(norm))
        packed = kernels.agg_packed_sum(view, kernels.pack_edge_meta(view, dev(norm), dev(w)), dev(x), dev(norm))
        assert torch.equal(plain, packed)
    ref = A.scaled_sum(f.row_offset, f.column_indices, f.eids, x, norm, w, norm)
    mag = A.scaled_sum(f.row_offset, f.column_indices, f.eids, x.abs(), norm, w, norm)
    got = kernels.agg_packed_sum(g.fwd_view(), kernels.pack_edge_meta(g.fwd_view(), dev(norm), dev(w)), dev(x), dev(norm))
    A.assert_close_rel(got.cpu(), ref, rel=1e-5, abs_terms=mag, what="packed hub rows")


def test_packed_accumulate_modes_and_empty(cuda):
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph

    n = 500
    src, dst = _graph(n, 3000, seed=8)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    x, norm, _, dev = _inputs(n, src.shape[0], 100, 2, False, cuda)
    view = g.fwd_view()
    meta = kernels.pack_edge_meta(view, dev(norm), None)
    base = kernels.agg_packed_sum(view, meta, dev(x), dev(norm))
    init = torch.randn(n, 100, device=cuda)
    acc = kernels.agg_packed_sum(view, meta, dev(x), dev(norm), out=init.clone(), accumulate=True)
    assert torch.equal(acc, init + base)
    red = kernels.agg_packed_sum(view, meta, dev(x), dev(norm), out=torch.zeros(n, 100, device=cuda), accumulate="red")
    assert torch.equal(red, base)
    # empty graph: rows are written as zeros, meta is never read
    g0 = StaticGraph(torch.zeros(0, 2, dtype=torch.int64), None, 10)
    m0 = kernels.pack_edge_meta(g0.fwd_view(), dev(norm)[:10].contiguous(), None)
    assert m0.shape == (0, 2)
    out = kernels.agg_packed_sum(g0.fwd_view(), m0, torch.randn(10, 16, device=cuda))
    assert torch.count_nonzero(out) == 0


def test_static_graph_cache_hits_and_repacks(cuda):
    """``CSR.packed_meta``: one pack per (scale tensors, version); an in-place update of the norm repacks."""
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph

    n = 600
    src, dst = _graph(n, 6000, seed=12)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    csr = g._forward_graph
    assert csr.pack_enabled
    x = torch.randn(n, 16, device=cuda)
    norm = g.degree_norm().reshape(-1)
    a = kernels.agg_scaled_sum_graph(csr, x, norm, None, norm)
    m1 = csr.packed_meta(norm.reshape(-1), None)
    assert m1 is not None and csr._meta_misses == 1
    b = kernels.agg_scaled_sum_graph(csr, x, norm.reshape(-1), None, norm)     # a fresh view of the same storage: hit
    assert csr._meta_misses == 1 and torch.equal(a, b)
    assert torch.equal(a, kernels.agg_scaled_sum(g.fwd_view(), x, norm, None, norm))
    norm.mul_(2.0)                                                              # version bump -> repack
    c = kernels.agg_scaled_sum_graph(csr, x, norm, None, norm)
    assert csr._meta_misses == 2
    assert torch.equal(c, kernels.agg_scaled_sum(g.fwd_view(), x, norm, None, norm))
    assert torch.allclose(c, 4.0 * a, rtol=1e-6, atol=0)
    # unscaled sums have nothing to pack
    assert csr.packed_meta(None, None) is None


def test_gcnconv_layer_uses_packed_kernel_and_matches_oracle(cuda):
    """Drop-in layer on a static graph: second call hits the cache; forward and gradient match the oracle."""
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.nn.pytorch import GCNConv

    n, fin, fout = 800, 24, 100
    src, dst = _graph(n, 9000, seed=4)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    norm = g.degree_norm()
    g.set_ndata("norm", norm)
    torch.manual_seed(0)
    layer = GCNConv(fin, fout).to(cuda)
    x = torch.randn(n, fin, device=cuda, requires_grad=True)
    outs = []
    for _ in range(2):
        x.grad = None
        out = layer(g, x)
        gout = torch.ones_like(out)
        out.backward(gout)
        outs.append((out.detach().clone(), x.grad.clone()))
    assert g._forward_graph._meta_misses == 1 and g._backward_graph._meta_misses == 1
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    f, b = S.forward_csr(src, dst, n), S.backward_csr(src, dst, n)
    h = (x.detach() @ layer.weight.detach()).cpu()
    nr = norm.cpu().reshape(-1)
    ref = A.gcn_forward(f, h, nr) + layer.bias.detach().cpu()
    A.assert_close_rel(outs[0][0].cpu(), ref, rel=1e-5, abs_terms=A.gcn_forward(f, h.abs(), nr) + 1e-6, what="GCNConv packed fwd")
    gh = A.gcn_backward(b, torch.ones(n, fout), nr)
    gx_ref = gh.double() @ layer.weight.detach().cpu().double().t()
    gx_mag = A.gcn_backward(b, torch.ones(n, fout), nr).double() @ layer.weight.detach().cpu().double().abs().t()
    A.assert_close_rel(outs[0][1].cpu(), gx_ref, rel=5e-5, abs_terms=gx_mag + 1e-6, what="GCNConv packed bwd")


def test_packed_linearity_and_adjoint_large(cuda):
    """Size-independent properties at a larger size through the packed kernel (power-law graph with hub rows)."""
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.utils import synthetic

    n = 200_000
    src, dst = synthetic.power_law_graph(n, 4_000_000, alpha=2.2, locality=0.8, seed=1, device=cuda)
    g = StaticGraph(torch.stack([src, dst], 1), None, n)
    norm = g.degree_norm().reshape(-1)
    x = torch.randn(n, 100, device=cuda)
    y = torch.randn(n, 100, device=cuda)
    fw, bw = g._forward_graph, g._backward_graph
    ax = kernels.agg_scaled_sum_graph(fw, x, norm, None, norm)
    assert torch.equal(ax, kernels.agg_scaled_sum(g.fwd_view(), x, norm, None, norm))
    by = kernels.agg_scaled_sum_graph(bw, y, norm, None, norm)
    assert torch.equal(by, kernels.agg_scaled_sum(g.bwd_view(), y, norm, None, norm))
    lhs = (ax.double() * y.double()).sum()
    rhs = (x.double() * by.double()).sum()
    assert abs(lhs - rhs) <= 5e-5 * max(abs(lhs), abs(rhs), 1.0)


@pytest.mark.parametrize("feat", [68, 72, 100, 124, 128])
@pytest.mark.parametrize("n_e", [(700, 9000), (12000, 150000), (20000, 900000)])   # >= 9472 rows: row-queue kernel, pair form
def test_pair_form_widths_and_row_lengths(cuda, feat, n_e):
    """The half-warp pair form of the edge loop (F = 68..128): odd/even row lengths, rows longer than one 32-edge
    batch, empty rows; plain and packed bit-identical, both within 1e-5 of the oracle."""
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph

    n, e = n_e
    src, dst = _graph(n, e, seed=feat + n)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    x, norm, w, dev = _inputs(n, src.shape[0], feat, feat, True, cuda)
    f = S.forward_csr(src, dst, n)
    view = g.fwd_view()
    plain = kernels.agg_scaled_sum(view, dev(x), dev(norm), dev(w), dev(norm))
    packed = kernels.agg_packed_sum(view, kernels.pack_edge_meta(view, dev(norm), dev(w)), dev(x), dev(norm))
    assert torch.equal(plain, packed)
    ref = A.scaled_sum(f.row_offset, f.column_indices, f.eids, x, norm, w, norm)
    mag = A.scaled_sum(f.row_offset, f.column_indices, f.eids, x.abs(), norm, w, norm)
    A.assert_close_rel(packed.cpu(), ref, rel=1e-5, abs_terms=mag, what=f"pair form F={feat}")


def test_pair_form_ignores_nonfinite_rows_it_does_not_sum(cuda):
    """Tail lanes of the pair form must not touch x[0]: a NaN/Inf there may only reach rows that have vertex 0 as a neighbour."""
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph

    n, feat = 12000, 100                                 # >= 9472 rows: row-queue kernel, pair form
    src, dst = _graph(n, 130000, seed=77)
    keep = src != 0
    src, dst = src[keep], dst[keep]                      # vertex 0 is nobody's in-neighbour
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    x = torch.randn(n, feat, device=cuda)
    x[0] = float("nan")
    out = kernels.agg_scaled_sum(g.fwd_view(), x)
    assert bool(torch.isfinite(out).all())


@pytest.mark.parametrize("feat", [7, 16, 100, 128, 200])
def test_padded_row_layout(cuda, feat):
    """x / out as row-padded views (rows on 128-byte lines): same bits as the dense layout."""
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph

    n = 900
    src, dst = _graph(n, 12000, seed=feat)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    x, norm, _, dev = _inputs(n, src.shape[0], feat, 31, False, cuda)
    for view in (g.fwd_view(), g.bwd_view()):
        meta = kernels.pack_edge_meta(view, dev(norm), None)
        dense = kernels.agg_packed_sum(view, meta, dev(x), dev(norm))
        xp = kernels.padded_rows(n, feat, cuda)
        assert xp.stride(0) % 32 == 0 and xp.data_ptr() % 128 == 0
        xp.copy_(dev(x))
        op = kernels.padded_rows(n, feat, cuda)
        torch.as_strided(op, (n, op.stride(0)), (op.stride(0), 1)).fill_(-7.0)
        got = kernels.agg_packed_sum(view, meta, xp, dev(norm), out=op)
        assert got.data_ptr() == op.data_ptr() and torch.equal(got, dense)
        if op.stride(0) != feat:          # the padding columns are never written
            full = torch.as_strided(op, (n, op.stride(0)), (op.stride(0), 1))
            assert bool((full[:, feat:] == -7.0).all())
        assert torch.equal(kernels.agg_packed_sum(view, meta, xp, dev(norm)), dense)      # out allocated padded
