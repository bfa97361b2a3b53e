"""``stg_gemm_tn_f32`` (the weight-gradient GEMM of the TGCN cell, ``csrc/gemm_tn.cu``) against a float64 product."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


def _check(a, b, colsum=True):
    from stgraph_b200 import kernels

    got = kernels.gemm_tn(a, b, colsum=colsum)
    c, cs = got if colsum else (got, None)
    ref = a.double().t() @ b.double()
    scale = (a.double().abs().t() @ b.double().abs()).clamp_min(1e-30)      # sum of |terms|: the fp32 error scale
    assert ((c.double() - ref).abs() <= 2e-6 * scale).all(), float(((c.double() - ref).abs() / scale).max())
    if colsum:
        rs = b.double().sum(0)
        sc = b.double().abs().sum(0).clamp_min(1e-30)
        assert ((cs.double() - rs).abs() <= 2e-6 * sc).all()
    return c


@pytest.mark.parametrize("m,k,nc", [(1, 1, 1), (5, 3, 7), (129, 8, 48), (1068, 16, 16), (4097, 64, 64), (20000, 32, 192),
                                     (33333, 64, 128), (70001, 100, 47), (3000, 130, 70)])
def test_gemm_tn_matches_float64(cuda, m, k, nc):
    g = torch.Generator(device=cuda).manual_seed(m + k + nc)
    a = torch.randn(m, k, device=cuda, generator=g)
    b = torch.randn(m, nc, device=cuda, generator=g)
    _check(a, b)
    _check(a, b, colsum=False)


def test_gemm_tn_on_column_blocks_and_unaligned_views(cuda):
    """Operands that are column blocks of wider matrices (leading dimension 3H) and views that start at an odd element
    (no 128-bit loads)."""
    g = torch.Generator(device=cuda).manual_seed(7)
    hid, m = 64, 50000
    h = torch.randn(m, 3 * hid, device=cuda, generator=g)
    dp = torch.randn(m, 3 * hid, device=cuda, generator=g)
    for blk in range(3):
        sl = slice(blk * hid, (blk + 1) * hid)
        _check(h[:, sl], dp[:, sl])
    _check(h[:, :hid], dp[:, : 2 * hid])
    wide = torch.randn(9001, 77, device=cuda, generator=g)
    _check(wide[:, 1:34], wide[:, 35:77])            # odd offsets: scalar loads
    _check(wide[:, 3:4], wide[:, 5:6])


def test_gemm_tn_is_deterministic_and_handles_empty(cuda):
    from stgraph_b200 import kernels

    g = torch.Generator(device=cuda).manual_seed(9)
    a = torch.randn(200000, 64, device=cuda, generator=g)
    b = torch.randn(200000, 64, device=cuda, generator=g)
    c1, s1 = kernels.gemm_tn(a, b, colsum=True)
    c2, s2 = kernels.gemm_tn(a, b, colsum=True)
    assert torch.equal(c1, c2) and torch.equal(s1, s2)
    c, s = kernels.gemm_tn(a[:0], b[:0], colsum=True)
    assert c.shape == (64, 64) and float(c.abs().sum()) == 0.0 and float(s.abs().sum()) == 0.0
    out = torch.empty(64, 64, device=cuda)
    assert kernels.gemm_tn(a, b, out=out) is out and torch.equal(out, c1)
    with pytest.raises(ValueError):
        kernels.gemm_tn(a, b[:-1])
    with pytest.raises(ValueError):
        kernels.gemm_tn(a.t().contiguous().t(), b)          # column-major view: no unit stride along the rows


def test_tgcn_cell_large_graph_uses_gemm_tn_and_matches_pieces(cuda):
    """Above ops_tgcn.TALL_ROWS vertices the cell's weight gradients go through gemm_tn: same gradients as the piecewise
    fused cell (torch autograd + cuBLAS)."""
    from stgraph_b200 import ops_tgcn
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.nn.pytorch import TGCN

    n, e = ops_tgcn.TALL_ROWS + 1234, 400000
    g = torch.Generator(device=cuda).manual_seed(3)
    src = torch.randint(0, n, (e,), device=cuda, generator=g)
    dst = torch.randint(0, n, (e,), device=cuda, generator=g)
    graph = StaticGraph(torch.stack([src, dst], 1), None, n)
    graph.set_ndata("norm", graph.degree_norm())
    torch.manual_seed(1)
    a = TGCN(32, 64, fused="pieces").to(cuda)
    b = TGCN(32, 64).to(cuda)
    b.load_state_dict(a.state_dict())
    xs = [torch.randn(n, 32, device=cuda, generator=g) for _ in range(3)]
    res = []
    for cell in (a, b):
        H, cost = None, 0
        for x in xs:
            H = cell(graph, x, None, H)
            cost = cost + (H ** 2).mean()
        cost.backward()
        res.append((H.detach(), {k: p.grad.clone() for k, p in cell.named_parameters()}))
    torch.testing.assert_close(res[0][0], res[1][0], rtol=1e-5, atol=1e-6)
    for k in res[0][1]:
        ga, gb = res[0][1][k], res[1][1][k]
        assert (ga - gb).abs().max() <= 5e-5 * ga.abs().max() + 1e-8, (k, float((ga - gb).abs().max()), float(ga.abs().max()))


def test_gcnconv_weight_gradient_through_gemm_tn(cuda):
    """GCNConv on a graph above ops_gcn.TALL_ROWS vertices: the weight gradient (gemm_tn) equals the cuBLAS one."""
    from stgraph_b200 import ops_gcn
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.nn.pytorch import GCNConv

    n, e = ops_gcn.TALL_ROWS + 777, 300000
    g = torch.Generator(device=cuda).manual_seed(11)
    src = torch.randint(0, n, (e,), device=cuda, generator=g)
    dst = torch.randint(0, n, (e,), device=cuda, generator=g)
    graph = StaticGraph(torch.stack([src, dst], 1), None, n)
    graph.set_ndata("norm", graph.degree_norm())
    torch.manual_seed(4)
    layer = GCNConv(24, 10).to(cuda)
    x = torch.randn(n, 24, device=cuda, generator=g, requires_grad=True)
    gout = torch.randn(n, 10, device=cuda, generator=g)
    layer(graph, x).backward(gout)
    got_w, got_x = layer.weight.grad.clone(), x.grad.clone()
    layer.zero_grad()
    x.grad = None
    old = ops_gcn.TALL_ROWS
    ops_gcn.TALL_ROWS = 1 << 62          # cuBLAS path
    try:
        layer(graph, x).backward(gout)
    finally:
        ops_gcn.TALL_ROWS = old
    assert (got_w - layer.weight.grad).abs().max() <= 2e-5 * layer.weight.grad.abs().max()
    torch.testing.assert_close(got_x, x.grad, rtol=1e-6, atol=1e-6)
