"""GPU parity: the fused gather-scale-sum kernel against the torch-CPU index_add oracle.

Tolerance: rel 1e-5 of max(|ref|, sum|terms|) (BASELINE.json north_star; SURVEY.md trap T9).
"""
import numpy as np
import pytest
import torch

from oracle import aggregate as A
from oracle import structure as S

pytestmark = pytest.mark.gpu


def _rand_graph(n, e, seed, hub=None):
    rng = np.random.default_rng(seed)
    key = rng.choice(n * n, size=min(e, n * n), replace=False)
    src, dst = (key // n).astype(np.int32), (key % n).astype(np.int32)
    if hub is not None:   # make vertex 0 a hub destination and source
        extra = np.arange(1, hub + 1, dtype=np.int32) % n
        extra = extra[extra != 0]
        src = np.concatenate([src, extra, np.zeros_like(extra)])
        dst = np.concatenate([dst, np.zeros_like(extra), extra])
        k = np.unique(src.astype(np.int64) * n + dst)
        src, dst = (k // n).astype(np.int32), (k % n).astype(np.int32)
    return src, dst


def _run(cuda, n, e, feat, seed, weighted, scales=True, hub=None):
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph

    src, dst = _rand_graph(n, e, seed, hub)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    tg = torch.Generator().manual_seed(seed)
    x = torch.randn(n, feat, generator=tg)
    norm = torch.rand(n, generator=tg) + 0.5 if scales else None
    w = (torch.rand(src.shape[0], generator=tg) + 0.1) if weighted else None
    f = S.forward_csr(src, dst, n)
    b = S.backward_csr(src, dst, n)
    dev = lambda t: None if t is None else t.to(cuda)
    for view, csr in ((g.fwd_view(), f), (g.bwd_view(), b)):
        got = kernels.agg_scaled_sum(view, dev(x), dev(norm), dev(w), dev(norm)).cpu()
        ref = A.scaled_sum(csr.row_offset, csr.column_indices, csr.eids, x, norm, w, norm)
        mag = A.scaled_sum(csr.row_offset, csr.column_indices, csr.eids, x.abs(), norm, w, norm)
        A.assert_close_rel(got, ref, rel=1e-5, abs_terms=mag, what=f"agg n={n} e={e} F={feat}")


@pytest.mark.parametrize("feat", [1, 2, 3, 4, 7, 8, 16, 20, 47, 64, 100, 128, 200, 256, 300, 512, 1000])
def test_feature_widths(cuda, feat):
    _run(cuda, 300, 3000, feat, seed=feat, weighted=False)


@pytest.mark.parametrize("feat", [7, 16, 100])
def test_edge_weighted(cuda, feat):
    _run(cuda, 500, 6000, feat, seed=100 + feat, weighted=True)


def test_no_scales(cuda):
    _run(cuda, 200, 1500, 32, seed=9, weighted=False, scales=False)


@pytest.mark.parametrize("feat", [4, 16, 100, 7])
def test_hub_rows_use_block_kernel(cuda, feat):
    from stgraph_b200.graph.static import csr as csr_mod

    _run(cuda, 4000, 20000, feat, seed=17, weighted=True, hub=3 * csr_mod.HUB_THRESHOLD)


def test_empty_rows_and_empty_graph(cuda):
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph

    g = StaticGraph(torch.zeros(0, 2, dtype=torch.int64), None, 10)
    x = torch.randn(10, 16, device=cuda)
    out = kernels.agg_scaled_sum(g.fwd_view(), x)
    assert torch.count_nonzero(out) == 0
    # single edge, most rows empty
    g = StaticGraph([(3, 7)], None, 10)
    out = kernels.agg_scaled_sum(g.fwd_view(), x).cpu()
    exp = torch.zeros(10, 16)
    exp[7] = x[3].cpu()
    assert torch.equal(out, exp)


def test_cora_shape_two_layers(cuda):
    """Config 1 shapes: F=16 then F=7 (all 7 columns are computed -- reference trap T1 drops 4..6)."""
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.utils import synthetic

    d = synthetic.cora_shaped(seed=0)
    src, dst = d["src"].numpy(), d["dst"].numpy()
    n = d["num_nodes"]
    g = StaticGraph(torch.stack([d["src"], d["dst"]], 1), None, n)
    assert g.get_num_edges() == 10556
    norm = g.degree_norm()
    f = S.forward_csr(src, dst, n)
    for feat in (16, 7):
        x = torch.randn(n, feat, generator=torch.Generator().manual_seed(feat))
        got = kernels.agg_scaled_sum(g.fwd_view(), x.to(cuda), norm.reshape(-1), None, norm.reshape(-1)).cpu()
        ref = A.gcn_forward(f, x, norm.cpu().reshape(-1))
        A.assert_close_rel(got, ref, rel=1e-5, abs_terms=A.gcn_forward(f, x.abs(), norm.cpu().reshape(-1)))
        assert torch.count_nonzero(got[:, -1]) > 0


def test_linearity_property_large(cuda):
    """Size-independent property at a larger size: agg(a*x + y) == a*agg(x) + agg(y) up to fp32 rounding."""
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.utils import synthetic

    src, dst = synthetic.power_law_graph(200_000, 4_000_000, alpha=2.2, locality=0.8, seed=1, device=cuda)
    g = StaticGraph(torch.stack([src, dst], 1), None, 200_000)
    x = torch.randn(200_000, 100, device=cuda)
    y = torch.randn(200_000, 100, device=cuda)
    norm = g.degree_norm().reshape(-1)
    ax = kernels.agg_scaled_sum(g.fwd_view(), x, norm, None, norm)
    ay = kernels.agg_scaled_sum(g.fwd_view(), y, norm, None, norm)
    axy = kernels.agg_scaled_sum(g.fwd_view(), 2.5 * x + y, norm, None, norm)
    mag = kernels.agg_scaled_sum(g.fwd_view(), 2.5 * x.abs() + y.abs(), norm, None, norm)
    err = (axy - (2.5 * ax + ay)).abs()
    assert bool((err <= 4e-6 * mag + 1e-30).all())
    # transpose property: <agg_fwd(x), y> == <x, agg_bwd(y)>  (backward kernel is the adjoint)
    lhs = (ax.double() * y.double()).sum()
    bx = kernels.agg_scaled_sum(g.bwd_view(), y, norm, None, norm)
    rhs = (x.double() * bx.double()).sum()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs), 1.0) * 50
