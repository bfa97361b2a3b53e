"""GPU parity of the layers (GCNConv / GATConv / TGCN) through the public API.

Oracles: closed forms in oracle/aggregate.py evaluated with torch-CPU index_add in fp64 and the
IR interpreter oracle/ir_interp.py.  Tolerance: rel 1e-5 of max(|ref|, sum|terms|).
"""
import numpy as np
import pytest
import torch

from oracle import aggregate as A
from oracle import structure as S
from oracle.ir_interp import Interp

pytestmark = pytest.mark.gpu


def _graph(n, e, seed, cuda):
    from stgraph_b200.graph import StaticGraph

    rng = np.random.default_rng(seed)
    key = rng.choice(n * n, size=e, replace=False)
    src, dst = (key // n).astype(np.int32), (key % n).astype(np.int32)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    return g, src, dst


def _gcn_oracle(src, dst, n, x, W, b, norm, w, gout):
    """fp64 torch-CPU restatement of gcn_conv.py:158-188 with autograd."""
    f = S.forward_csr(src, dst, n)
    rows, cols = A.csr_to_coo(f.row_offset, f.column_indices)
    x64 = x.double().requires_grad_(True)
    W64 = W.double().requires_grad_(True)
    b64 = b.double().requires_grad_(True)
    h = x64 @ W64
    msg = h[cols] * norm.double()[cols]
    if w is not None:
        msg = msg * w.double().reshape(-1, 1)          # CSR slot i has eid i in the forward graph
    out = torch.zeros(n, W.shape[1], dtype=torch.float64).index_add(0, rows, msg) * norm.double() + b64
    out.backward(gout.double())
    mag = torch.zeros(n, W.shape[1], dtype=torch.float64).index_add(0, rows, msg.detach().abs()) * norm.double()
    return out.detach(), x64.grad, W64.grad, b64.grad, mag


@pytest.mark.parametrize("fin,fout,weighted", [(32, 16, False), (32, 7, False), (20, 16, True), (20, 7, True), (64, 100, False)])
def test_gcnconv_forward_backward(cuda, fin, fout, weighted):
    from stgraph_b200.nn.pytorch import GCNConv

    n, e = 400, 5000
    g, src, dst = _graph(n, e, seed=fin + fout, cuda=cuda)
    norm = g.degree_norm()
    g.set_ndata("norm", norm)
    torch.manual_seed(2)
    layer = GCNConv(fin, fout).to(cuda)
    with torch.no_grad():
        layer.bias.uniform_(-0.1, 0.1)
    x = torch.randn(n, fin, device=cuda, requires_grad=True)
    w = (torch.rand(e, 1, device=cuda) + 0.1) if weighted else None
    gout = torch.randn(n, fout, device=cuda)
    out = layer(g, x, edge_weight=w)
    out.backward(gout)
    ref, gx, gW, gb, mag = _gcn_oracle(src, dst, n, x.detach().cpu(), layer.weight.detach().cpu(),
                                       layer.bias.detach().cpu(), norm.cpu(), None if w is None else w.cpu(), gout.cpu())
    A.assert_close_rel(out.detach().cpu(), ref, rel=1e-5, abs_terms=mag + 1e-3, what="GCNConv out")
    A.assert_close_rel(x.grad.cpu(), gx, rel=2e-5, abs_terms=gx.abs().mean() * torch.ones_like(gx), what="dX")
    A.assert_close_rel(layer.weight.grad.cpu(), gW, rel=2e-5, abs_terms=gW.abs().mean() * torch.ones_like(gW), what="dW")
    A.assert_close_rel(layer.bias.grad.cpu(), gb, rel=1e-5, abs_terms=gout.abs().sum(0).cpu().double(), what="db")


def test_gcnconv_validates_norm(cuda):
    from stgraph_b200.nn.pytorch import GCNConv

    g, _, _ = _graph(50, 200, 0, cuda)
    layer = GCNConv(4, 4).to(cuda)
    with pytest.raises(KeyError):
        layer(g, torch.randn(50, 4, device=cuda))
    g.set_ndata("norm", torch.rand(50, device=cuda))
    with pytest.raises(ValueError):
        layer(g, torch.randn(50, 4, device=cuda))


def test_gcn_two_layer_model_state_stack(cuda):
    """Two layers + the same layer applied twice (BPTT-style LIFO use of the executor's state stack)."""
    from stgraph_b200.nn.pytorch import GCNConv

    n, e = 300, 3000
    g, src, dst = _graph(n, e, 7, cuda)
    norm = g.degree_norm()
    g.set_ndata("norm", norm)
    torch.manual_seed(0)
    l1 = GCNConv(16, 16, activation=torch.relu).to(cuda)
    x = torch.randn(n, 16, device=cuda, requires_grad=True)
    h = l1(g, x)
    h = l1(g, h)          # same executor, second entry on the stack
    h = l1(g, h)
    loss = (h * h).sum()
    loss.backward()
    assert len(l1.stgraph._ctx_map) == 1
    ctx = next(iter(l1.stgraph._ctx_map.values()))
    assert len(ctx._executor_cache.ts.tensor_map_stack) == 0      # every forward entry was popped
    # oracle
    f = S.forward_csr(src, dst, n)
    rows, cols = A.csr_to_coo(f.row_offset, f.column_indices)
    x64 = x.detach().cpu().double().requires_grad_(True)
    W = l1.weight.detach().cpu().double().requires_grad_(True)
    b = l1.bias.detach().cpu().double()
    nr = norm.cpu().double()
    hh = x64
    for _ in range(3):
        t = hh @ W
        hh = torch.relu(torch.zeros(n, 16, dtype=torch.float64).index_add(0, rows, t[cols] * nr[cols]) * nr + b)
    (hh * hh).sum().backward()
    A.assert_close_rel(h.detach().cpu(), hh.detach(), rel=1e-5, abs_terms=hh.detach().abs().mean() * torch.ones_like(hh))
    A.assert_close_rel(x.grad.cpu(), x64.grad, rel=5e-5, abs_terms=x64.grad.abs().mean() * torch.ones_like(x64.grad))
    A.assert_close_rel(l1.weight.grad.cpu(), W.grad, rel=5e-5, abs_terms=W.grad.abs().mean() * torch.ones_like(W.grad))


@pytest.mark.parametrize("heads,dim", [(8, 16), (2, 4), (1, 32), (3, 5)])
def test_gatconv_stock_matches_reference_semantics(cuda, heads, dim):
    """Stock GATConv == mean over in-neighbours, with the reference's el/er gradients (trap T2)."""
    from stgraph_b200.nn.pytorch import GATConv

    n, e = 300, 2500
    g, src, dst = _graph(n, e, seed=heads * 10 + dim, cuda=cuda)
    torch.manual_seed(1)
    layer = GATConv(24, dim, heads).to(cuda)
    x = torch.randn(n, 24, device=cuda, requires_grad=True)
    gout = torch.randn(n, heads, dim, device=cuda)
    out = layer(g, x)
    out.backward(gout)
    # oracle through torch-CPU: fc / el / er are plain torch; the vertex program uses the closed forms
    f = S.forward_csr(src, dst, n)
    xc = x.detach().cpu().double().requires_grad_(True)
    Wfc = layer.fc.weight.detach().cpu().double().requires_grad_(True)
    al = layer.attn_l.detach().cpu().double().requires_grad_(True)
    ar = layer.attn_r.detach().cpu().double().requires_grad_(True)
    feat = (xc @ Wfc.t()).view(-1, heads, dim)
    el = (feat * al).sum(-1).unsqueeze(-1)
    er = (feat * ar).sum(-1).unsqueeze(-1)
    ref, _, _ = A.gat_stock_forward(f, el.detach(), er.detach(), feat.detach())
    d_feat, d_el, d_er = A.gat_stock_backward(f, el.detach(), er.detach(), feat.detach(), gout.cpu().double())
    torch.autograd.backward([feat, el, er], [d_feat, d_el, d_er])
    scale = lambda t: t.abs().mean() * torch.ones_like(t) + 1e-12
    A.assert_close_rel(out.detach().cpu(), ref, rel=1e-5, abs_terms=scale(ref), what="GAT out")
    A.assert_close_rel(x.grad.cpu(), xc.grad, rel=1e-4, abs_terms=scale(xc.grad), what="GAT dX")
    A.assert_close_rel(layer.fc.weight.grad.cpu(), Wfc.grad, rel=1e-4, abs_terms=scale(Wfc.grad), what="GAT dWfc")
    A.assert_close_rel(layer.attn_l.grad.cpu(), al.grad, rel=1e-4, abs_terms=scale(al.grad), what="GAT d_attn_l")
    # d_er (hence d_attn_r) is an exact zero in real arithmetic (softmax weights sum to one): compare against
    # the magnitude of the terms, for which d_attn_l is representative
    A.assert_close_rel(layer.attn_r.grad.cpu(), ar.grad, rel=1e-4, abs_terms=scale(al.grad) * 10, what="GAT d_attn_r")
    # forward really is the neighbour mean (independent of el / er)
    rows, cols = A.csr_to_coo(f.row_offset, f.column_indices)
    deg = torch.from_numpy(f.row_degrees).double().clamp(min=1).reshape(-1, 1, 1)
    mean = torch.zeros(n, heads, dim, dtype=torch.float64).index_add(0, rows, feat.detach()[cols]) / deg
    A.assert_close_rel(out.detach().cpu(), mean, rel=1e-5, abs_terms=scale(mean), what="GAT mean")


def test_gat_units_match_ir_interpreter(cuda):
    """Every unit of the stock GAT (forward and backward) against the IR-level oracle."""
    from stgraph_b200.nn.pytorch import GATConv

    n, e, heads, dim = 120, 900, 4, 8
    g, src, dst = _graph(n, e, 3, cuda)
    torch.manual_seed(4)
    layer = GATConv(10, dim, heads).to(cuda)
    x = torch.randn(n, 10, device=cuda, requires_grad=True)
    out = layer(g, x)
    gout = torch.randn_like(out)
    ctx = next(iter(layer.stgraph._ctx_map.values()))
    ex = ctx._executor_cache
    saved = dict(ex.ts.tensor_map_stack.top())
    out.backward(gout)
    f = S.forward_csr(src, dst, n)
    order = np.lexsort((src, dst))
    it = Interp(src[order], dst[order], n)           # edges in eid order
    feat = layer.fc(x).view(-1, heads, dim).detach()
    el = (feat * layer.attn_l).sum(-1).unsqueeze(-1).detach()
    er = (feat * layer.attn_r).sum(-1).unsqueeze(-1).detach()
    env = {"Velinb": el.cpu(), "Velcen": el.cpu(), "Vercen": er.cpu(), "Verinb": er.cpu(),
           "Vfeat_srcinb": feat.cpu(), "Vfeat_srccen": feat.cpu()}
    env = it.run_units(ctx.forward_units, env)
    A.assert_close_rel(out.detach().cpu(), env[ex._rets[0].id], rel=1e-5,
                       abs_terms=env[ex._rets[0].id].abs().mean() * torch.ones_like(env[ex._rets[0].id]))
    for k, t in saved.items():
        if k in env and not k.endswith(("inb", "cen")):
            A.assert_close_rel(t.cpu(), env[k], rel=1e-5, abs_terms=env[k].abs().mean() * torch.ones_like(env[k]) + 1e-12,
                               what=f"saved {k}")
    for var, gv in ex.grad_in.items():
        env[gv.id] = gout.cpu()
    env = it.run_units(ctx.backward_units, env)
    assert set(v.id for v in ex.grad_out) == {"Velinb", "Vercen", "Vfeat_srcinb"}


@pytest.mark.parametrize("fused", [None, "pieces", False])
def test_tgcn_cell_matches_torch(cuda, fused):
    """Default construction (the one-op fused cell is picked automatically), the piecewise fused cell and the
    reference-structured path."""
    from stgraph_b200.nn.pytorch import TGCN

    n, e = 200, 2000
    g, src, dst = _graph(n, e, 11, cuda)
    norm = g.degree_norm()
    g.set_ndata("norm", norm)
    torch.manual_seed(3)
    cell = TGCN(8, 16, fused=fused).to(cuda)
    w = torch.rand(e, 1, device=cuda) + 0.1
    xs = [torch.randn(n, 8, device=cuda) for _ in range(4)]
    H = None
    cost = 0
    for x in xs:
        H = cell(g, x, w, H)
        cost = cost + (H ** 2).mean()
    cost.backward()
    got = {k: p.grad.detach().cpu().double() for k, p in cell.named_parameters()}
    # torch-CPU fp64 restatement of temporal/tgcn.py:21-55
    f = S.forward_csr(src, dst, n)
    rows, cols = A.csr_to_coo(f.row_offset, f.column_indices)
    P = {k: p.detach().cpu().double().requires_grad_(True) for k, p in cell.named_parameters()}
    nr, wc = norm.cpu().double(), w.cpu().double()

    def conv(x, name):
        t = x @ P[f"conv_{name}.weight"]
        agg = torch.zeros(n, 16, dtype=torch.float64).index_add(0, rows, t[cols] * nr[cols] * wc) * nr
        return torch.clamp(agg + P[f"conv_{name}.bias"], -1e6, 1e6)

    def lin(t, name):
        return t @ P[f"linear_{name}.weight"].t() + P[f"linear_{name}.bias"]

    Hc = torch.zeros(n, 16, dtype=torch.float64)
    cost_c = 0
    for x in xs:
        xc = x.cpu().double()
        Z = torch.sigmoid(lin(torch.cat((conv(xc, "z"), Hc), 1), "z"))
        R = torch.sigmoid(lin(torch.cat((conv(xc, "r"), Hc), 1), "r"))
        Ht = torch.tanh(lin(torch.cat((conv(xc, "h"), Hc * R), 1), "h"))
        Hc = Z * Hc + (1 - Z) * Ht
        cost_c = cost_c + (Hc ** 2).mean()
    cost_c.backward()
    A.assert_close_rel(H.detach().cpu(), Hc.detach(), rel=2e-5, abs_terms=torch.ones_like(Hc))
    for k in got:
        ref = P[k].grad
        A.assert_close_rel(got[k], ref, rel=2e-4, abs_terms=ref.abs().mean() * torch.ones_like(ref) + 1e-9, what=k)


@pytest.mark.parametrize("heads,dim", [(8, 16), (2, 4), (1, 32), (3, 8), (4, 2), (8, 64), (5, 1)])
def test_fused_edge_softmax_matches_closed_form(cuda, heads, dim):
    """Genuine edge softmax (online softmax forward, recompute-alpha backward) vs the fp64 torch oracle."""
    from stgraph_b200.ops_gat import gat_edge_softmax_aggregate

    n, e = 500, 6000
    g, src, dst = _graph(n, e, seed=heads + dim, cuda=cuda)
    f = S.forward_csr(src, dst, n)
    tg = torch.Generator().manual_seed(heads * 100 + dim)
    el = (torch.randn(n, heads, 1, generator=tg) * 2).to(cuda).requires_grad_(True)
    er = (torch.randn(n, heads, 1, generator=tg) * 2).to(cuda).requires_grad_(True)
    feat = torch.randn(n, heads, dim, generator=tg).to(cuda).requires_grad_(True)
    gout = torch.randn(n, heads, dim, generator=tg).to(cuda)
    out = gat_edge_softmax_aggregate(g, el, er, feat, 0.2)
    out.backward(gout)
    ref, _, _ = A.gat_softmax_forward(f, el.detach().cpu(), er.detach().cpu(), feat.detach().cpu())
    d_feat, d_el, d_er = A.gat_softmax_backward(f, el.detach().cpu(), er.detach().cpu(), feat.detach().cpu(), gout.cpu())
    sc = lambda t: t.abs().mean() * torch.ones_like(t) + 1e-12
    A.assert_close_rel(out.detach().cpu(), ref, rel=1e-5, abs_terms=sc(ref), what="out")
    A.assert_close_rel(feat.grad.cpu(), d_feat, rel=2e-5, abs_terms=sc(d_feat), what="d_feat")
    A.assert_close_rel(el.grad.cpu().reshape(n, heads), d_el, rel=5e-5, abs_terms=sc(d_el), what="d_el")
    A.assert_close_rel(er.grad.cpu().reshape(n, heads), d_er, rel=5e-5, abs_terms=sc(d_el), what="d_er")


def test_gatconv_fused_layer_runs_and_differs_from_stock_mean(cuda):
    from stgraph_b200.nn.pytorch import GATConv

    n, e = 300, 3000
    g, src, dst = _graph(n, e, 21, cuda)
    torch.manual_seed(5)
    fused = GATConv(16, 8, 4, softmax="fused").to(cuda)
    stock = GATConv(16, 8, 4).to(cuda)
    stock.load_state_dict(fused.state_dict())          # same parameter names as the reference layer
    x = torch.randn(n, 16, device=cuda)
    a, b = fused(g, x), stock(g, x)
    assert a.shape == b.shape == (n, 4, 8)
    assert not torch.allclose(a, b, atol=1e-3)         # the stock trace is a plain mean (trap T2)
    a.sum().backward()
    assert fused.attn_l.grad is not None and torch.isfinite(fused.attn_l.grad).all()


def test_fused_edge_softmax_rejects_unsupported_dim(cuda):
    from stgraph_b200.ops_gat import gat_edge_softmax_aggregate

    g, _, _ = _graph(50, 200, 2, cuda)
    el = torch.randn(50, 2, 1, device=cuda)
    feat = torch.randn(50, 2, 6, device=cuda)          # 6 lanes per head: not a power of two
    with pytest.raises(RuntimeError, match="power of two"):
        gat_edge_softmax_aggregate(g, el, el, feat)


@pytest.mark.parametrize("mode", [True, "pieces"])
def test_tgcn_fused_cell_equals_reference_structure(cuda, mode):
    """fused=True (one GEMM + one aggregation of width 3H) gives the same values and gradients."""
    from stgraph_b200.nn.pytorch import TGCN

    n, e = 300, 3500
    g, src, dst = _graph(n, e, 13, cuda)
    g.set_ndata("norm", g.degree_norm())
    torch.manual_seed(7)
    a = TGCN(8, 16, fused=False).to(cuda)
    b = TGCN(8, 16, fused=mode).to(cuda)
    b.load_state_dict(a.state_dict())
    w = torch.rand(e, 1, device=cuda) + 0.1
    xs = [torch.randn(n, 8, device=cuda) for _ in range(5)]
    outs = []
    for cell in (a, b):
        H, cost = None, 0
        for x in xs:
            H = cell(g, x, w, H)
            cost = cost + (H ** 2).mean()
        cost.backward()
        outs.append((H.detach(), {k: p.grad.detach().clone() for k, p in cell.named_parameters()}))
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-6)
    for k in outs[0][1]:
        ga, gb = outs[0][1][k], outs[1][1][k]
        assert (ga - gb).abs().max() <= 1e-5 * ga.abs().max() + 1e-7, k


@pytest.mark.parametrize("heads,dim", [(8, 16), (2, 4), (4, 64)])
def test_fused_edge_softmax_hub_rows(cuda, heads, dim):
    """Rows longer than the hub threshold take the block-per-row kernels (partials merged in shared memory)."""
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.graph.static import csr as csr_mod
    from stgraph_b200.ops_gat import gat_edge_softmax_aggregate

    n = 5000
    hub = 3 * csr_mod.HUB_THRESHOLD
    rng = np.random.default_rng(heads)
    key = rng.choice(n * n, size=20000, replace=False)
    src, dst = (key // n).astype(np.int64), (key % n).astype(np.int64)
    extra = np.arange(1, hub + 1) % n
    extra = extra[extra != 0]
    src = np.concatenate([src, extra, np.zeros_like(extra)])       # vertex 0: hub destination AND hub source
    dst = np.concatenate([dst, np.zeros_like(extra), extra])
    for v, deg in ((1, 300), (2, 700), (3, 129)):                  # medium hub rows: one CTA each (attention kernels)
        other = np.arange(10, 10 + deg)
        src = np.concatenate([src, other, np.full(deg, v)])
        dst = np.concatenate([dst, np.full(deg, v), other])
    k = np.unique(src * n + dst)
    src, dst = (k // n).astype(np.int32), (k % n).astype(np.int32)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    assert int(g.in_degrees_tensor().max()) > csr_mod.HUB_THRESHOLD
    from stgraph_b200 import ops_gat
    deg_in = g.in_degrees_tensor()
    assert int(((deg_in > ops_gat.GAT_HUB_THRESHOLD) & (deg_in <= 1024)).sum()) >= 3      # both hub tiers are exercised
    f = S.forward_csr(src, dst, n)
    tg = torch.Generator().manual_seed(3)
    el = (torch.randn(n, heads, 1, generator=tg) * 2).to(cuda).requires_grad_(True)
    er = (torch.randn(n, heads, 1, generator=tg) * 2).to(cuda).requires_grad_(True)
    feat = torch.randn(n, heads, dim, generator=tg).to(cuda).requires_grad_(True)
    gout = torch.randn(n, heads, dim, generator=tg).to(cuda)
    out = gat_edge_softmax_aggregate(g, el, er, feat, 0.2)
    out.backward(gout)
    ref, _, _ = A.gat_softmax_forward(f, el.detach().cpu(), er.detach().cpu(), feat.detach().cpu())
    d_feat, d_el, d_er = A.gat_softmax_backward(f, el.detach().cpu(), er.detach().cpu(), feat.detach().cpu(), gout.cpu())
    sc = lambda t: t.abs().mean() * torch.ones_like(t) + 1e-12
    A.assert_close_rel(out.detach().cpu(), ref, rel=1e-5, abs_terms=sc(ref), what="out")
    A.assert_close_rel(feat.grad.cpu(), d_feat, rel=2e-5, abs_terms=sc(d_feat), what="d_feat")
    A.assert_close_rel(el.grad.cpu().reshape(n, heads), d_el, rel=1e-4, abs_terms=sc(d_el), what="d_el")
    A.assert_close_rel(er.grad.cpu().reshape(n, heads), d_er, rel=1e-4, abs_terms=sc(d_el), what="d_er")
    # deterministic: a second run is bit-identical
    out2 = gat_edge_softmax_aggregate(g, el.detach(), er.detach(), feat.detach(), 0.2)
    assert torch.equal(out2, out.detach())


def test_stock_gat_vm_kernel_hub_rows(cuda):
    """The generic VM kernel splits hub rows over a block and merges accumulators by kind."""
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.graph.static import csr as csr_mod
    from stgraph_b200.nn.pytorch import GATConv

    n, heads, dim = 4000, 4, 8
    hub = 2 * csr_mod.HUB_THRESHOLD + 37
    rng = np.random.default_rng(5)
    key = rng.choice(n * n, size=15000, replace=False)
    src, dst = (key // n).astype(np.int64), (key % n).astype(np.int64)
    extra = np.arange(1, hub + 1) % n
    extra = extra[extra != 0]
    src = np.concatenate([src, extra, np.zeros_like(extra)])
    dst = np.concatenate([dst, np.zeros_like(extra), extra])
    k = np.unique(src * n + dst)
    src, dst = (k // n).astype(np.int32), (k % n).astype(np.int32)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    f = S.forward_csr(src, dst, n)
    torch.manual_seed(9)
    layer = GATConv(12, dim, heads).to(cuda)
    x = torch.randn(n, 12, device=cuda, requires_grad=True)
    gout = torch.randn(n, heads, dim, device=cuda)
    out = layer(g, x)
    out.backward(gout)
    feat = layer.fc(x.detach()).view(-1, heads, dim).detach().cpu()
    el = (feat * layer.attn_l.detach().cpu()).sum(-1, keepdim=True)
    er = (feat * layer.attn_r.detach().cpu()).sum(-1, keepdim=True)
    ref, _, _ = A.gat_stock_forward(f, el, er, feat)
    sc = lambda t: t.abs().mean() * torch.ones_like(t) + 1e-12
    A.assert_close_rel(out.detach().cpu(), ref, rel=1e-5, abs_terms=sc(ref), what="stock GAT out (hub)")
    d_feat, d_el, d_er, m_feat, m_el, m_er = A.gat_stock_backward(f, el, er, feat, gout.cpu(), return_mag=True)
    # d_feat reaches x through fc: compare the fc-weight gradient contribution of d_feat + d_el + d_er
    featg = (x.detach().cpu().double() @ layer.fc.weight.detach().cpu().double().t()).view(-1, heads, dim).requires_grad_(True)
    al, ar = layer.attn_l.detach().cpu().double(), layer.attn_r.detach().cpu().double()
    elg, erg = (featg * al).sum(-1, keepdim=True), (featg * ar).sum(-1, keepdim=True)
    torch.autograd.backward([featg, elg, erg], [d_feat.double(), d_el.double(), d_er.double()])
    gx_ref = featg.grad.reshape(n, -1) @ layer.fc.weight.detach().cpu().double()
    A.assert_close_rel(x.grad.cpu(), gx_ref, rel=1e-4, abs_terms=sc(gx_ref), what="stock GAT dX (hub)")


# ---------------------------------------------------------------------------------------------------------
# fused element-wise pieces of the TGCN cell (csrc/gates.cu, ops_gru.py) against plain torch
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (300, 16), (1000, 64), (4099, 48)])
def test_gru_gate_ops_match_torch(cuda, shape):
    from stgraph_b200.ops_gru import bias_clamp, gru_reset, gru_update

    torch.manual_seed(sum(shape))
    n, hdim = shape
    mk = lambda scale=1.0: (scale * torch.randn(n, hdim, device=cuda)).requires_grad_()
    # reset gate
    pr, h = mk(3.0), mk()
    ref = h * torch.sigmoid(pr)
    got = gru_reset(pr, h)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    go = torch.randn_like(ref)
    g_ref = torch.autograd.grad(ref, (pr, h), go)
    g_got = torch.autograd.grad(got, (pr, h), go)
    for a, b in zip(g_got, g_ref):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    # update gate + candidate state
    pz, ph, h = mk(3.0), mk(2.0), mk()
    z = torch.sigmoid(pz)
    ref = z * h + (1 - z) * torch.tanh(ph)
    got = gru_update(pz, ph, h)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    g_ref = torch.autograd.grad(ref, (pz, ph, h), go)
    g_got = torch.autograd.grad(got, (pz, ph, h), go)
    for a, b in zip(g_got, g_ref):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=2e-6)
    # bias + clamp, in place, with values on both sides of the bounds
    a0 = (4.0 * torch.randn(n, hdim, device=cuda)).requires_grad_()
    bias = torch.randn(hdim, device=cuda, requires_grad=True)
    ref = torch.clamp(a0 + bias, min=-5.0, max=5.0)
    got = bias_clamp(a0 * 1.0, bias, -5.0, 5.0)        # `* 1.0`: the op overwrites its (intermediate) input
    assert torch.equal(got, ref)
    g_ref = torch.autograd.grad(ref, (a0, bias), go)
    g_got = torch.autograd.grad(got, (a0, bias), go)
    torch.testing.assert_close(g_got[0], g_ref[0], rtol=0, atol=0)
    torch.testing.assert_close(g_got[1], g_ref[1], rtol=1e-5, atol=1e-5)


def test_gru_gate_ops_extreme_inputs(cuda):
    """Saturated gates stay finite: sigmoid(+-100), tanh(+-50), and the +-1e6 clamp of the cell."""
    from stgraph_b200.ops_gru import bias_clamp, gru_reset, gru_update

    big = torch.tensor([[-100.0, -20.0, 0.0, 20.0, 100.0]], device=cuda)
    h = torch.ones_like(big)
    assert torch.allclose(gru_reset(big, h), torch.sigmoid(big), rtol=1e-5, atol=1e-30)
    out = gru_update(big, 0.5 * big, h)
    z = torch.sigmoid(big)
    assert torch.isfinite(out).all() and torch.allclose(out, z * h + (1 - z) * torch.tanh(0.5 * big), rtol=1e-5, atol=1e-7)
    a = torch.tensor([[-3e6, -1e6, 0.5, 1e6, 3e6]], device=cuda)
    assert torch.equal(bias_clamp(a.clone(), None, -1e6, 1e6), torch.clamp(a, -1e6, 1e6))


def test_state_stack_does_not_grow_without_a_backward(cuda):
    """forward calls under no_grad / on inputs that need no gradient must not push on the LIFO stacks."""
    from stgraph_b200.nn.pytorch import GCNConv

    g, _, _ = _graph(200, 1500, 5, cuda)
    g.set_ndata("norm", g.degree_norm())
    layer = GCNConv(8, 8).to(cuda)
    x = torch.randn(200, 8, device=cuda)
    with torch.no_grad():
        for _ in range(5):
            layer(g, x)
    ctx = next(iter(layer.stgraph._ctx_map.values()))
    assert len(ctx._executor_cache.ts.tensor_map_stack) == 0
    y = layer(g, x)                       # weights need a gradient: one entry, popped by backward
    stack = ctx._executor_cache.ts.tensor_map_stack      # (executors are keyed by the inputs' requires_grad signature)
    assert len(stack) == 1
    with torch.no_grad():
        layer(g, x @ layer.weight.detach() * 0 + x)      # a no-grad call in between leaves the pending entry alone
    assert len(stack) == 1
    y.sum().backward()
    assert len(stack) == 0


def test_fused_clamp_propagates_nan_like_torch(cuda):
    from stgraph_b200 import ops_gru

    a = torch.tensor([[1.0, float("nan"), 2e6, -3e6]], device=cuda)
    ref = torch.clamp(a.clone(), -1e6, 1e6)
    a2 = a.clone().requires_grad_(True)
    got = ops_gru.bias_clamp(a2 * 1.0, None, -1e6, 1e6)
    assert torch.equal(torch.isnan(got), torch.isnan(ref)) and torch.equal(got[~torch.isnan(got)], ref[~torch.isnan(ref)])
    got.backward(torch.ones_like(got))
    a3 = a.clone().requires_grad_(True)
    torch.clamp(a3 * 1.0, -1e6, 1e6).backward(torch.ones_like(a3))
    assert torch.equal(a2.grad, a3.grad)            # torch's mask: no gradient where the value is NaN or clamped


def _tgcn_pair(cuda, n=260, e=3000, seed=21):
    from stgraph_b200.nn.pytorch import TGCN

    g, src, dst = _graph(n, e, seed, cuda)
    g.set_ndata("norm", g.degree_norm())
    torch.manual_seed(seed)
    a = TGCN(8, 16, fused=False).to(cuda)
    b = TGCN(8, 16).to(cuda)
    b.load_state_dict(a.state_dict())
    w = torch.rand(e, 1, device=cuda) + 0.1
    return g, a, b, w


def _window(cell, g, xs, w, H=None):
    cost = 0
    for x in xs:
        H = cell(g, x, w, H)
        cost = cost + (H ** 2).mean()
    return cost, H


def test_tgcn_packed_parameters_follow_optimizer_steps_and_repeated_backwards(cuda):
    """The one-op cell packs its parameters once per BPTT window: the pack must be rebuilt after an optimizer step
    (in-place parameter update) and after a backward pass (gradient accumulation over two windows without a step)."""
    g, a, b, w = _tgcn_pair(cuda)
    xs = [torch.randn(260, 8, device=cuda) for _ in range(6)]
    opts = [torch.optim.SGD(m.parameters(), lr=0.05) for m in (a, b)]
    for cell, opt in zip((a, b), opts):
        # window 1 -> step; windows 2 and 3 accumulate into .grad without a step in between (truncated BPTT, H detached)
        cost, H = _window(cell, g, xs[:2], w)
        opt.zero_grad()
        cost.backward()
        opt.step()
        opt.zero_grad()
        cost, H = _window(cell, g, xs[2:4], w, H.detach())
        cost.backward()
        cost, H = _window(cell, g, xs[4:], w, H.detach())
        cost.backward()
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert (pa - pb).abs().max() <= 1e-5 * pa.abs().max() + 1e-7, k
        assert (pa.grad - pb.grad).abs().max() <= 2e-5 * pa.grad.abs().max() + 1e-7, k
    # the pack is reused inside a window and dropped once a backward has run through it
    b.zero_grad()
    cost, _ = _window(b, g, xs[:3], w)
    first = b._pack_cache[1][0]
    assert b._packed_parameters()[0] is first
    cost.backward()
    assert b._packed_parameters()[0] is not first


def test_tgcn_one_op_cell_input_gradient_and_no_grad(cuda):
    """d/dX and d/dH of the one-op cell against the piecewise fused cell; inference under no_grad."""
    from stgraph_b200.nn.pytorch import TGCN

    g, a, b, w = _tgcn_pair(cuda, seed=33)
    c = TGCN(8, 16, fused="pieces").to(cuda)
    c.load_state_dict(b.state_dict())
    x0 = torch.randn(260, 8, device=cuda)
    h0 = torch.randn(260, 16, device=cuda)
    grads = []
    for cell in (b, c):
        x, h = x0.clone().requires_grad_(True), h0.clone().requires_grad_(True)
        out = cell(g, x, w, h)
        (out * torch.linspace(-1, 1, 16, device=cuda)).sum().backward()
        grads.append((out.detach(), x.grad, h.grad))
    torch.testing.assert_close(grads[0][0], grads[1][0], rtol=1e-5, atol=1e-6)
    for u, v in zip(grads[0][1:], grads[1][1:]):
        assert (u - v).abs().max() <= 1e-5 * v.abs().max() + 1e-7
    with torch.no_grad():
        out = b(g, x0, w, h0)
    assert not out.requires_grad
    torch.testing.assert_close(out, grads[0][0], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("n,hid", [(1, 1), (37, 6), (300, 16), (2049, 64)])
def test_tgcn_block_layout_gate_kernels_match_torch(cuda, n, hid):
    """stg_tgcn_{reset,update}_{fwd,bwd}_f32 on the (z | r | h) column blocks of one [N, 3H] matrix (float4 and scalar
    forms) against torch autograd on the separate gates."""
    from stgraph_b200 import _lib

    torch.manual_seed(n + hid)
    P = (2.0 * torch.randn(n, 3 * hid, device=cuda)).requires_grad_()
    H = torch.randn(n, hid, device=cuda, requires_grad=True)
    st = _lib.current_stream_ptr()
    pz, pr, ph = P[:, :hid], P[:, hid:2 * hid], P[:, 2 * hid:]
    hr_ref = H * torch.sigmoid(pr)
    z = torch.sigmoid(pz)
    out_ref = z * H + (1 - z) * torch.tanh(ph)
    hr, out = torch.empty_like(H), torch.empty_like(H)
    _lib.call("stg_tgcn_reset_fwd_f32", P.data_ptr(), H.data_ptr(), hr.data_ptr(), n, hid, st)
    _lib.call("stg_tgcn_update_fwd_f32", P.data_ptr(), H.data_ptr(), out.data_ptr(), n, hid, st)
    torch.testing.assert_close(hr, hr_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out, out_ref, rtol=1e-5, atol=1e-6)
    d_out, d_hr = torch.randn_like(H), torch.randn_like(H)
    gP, gH = torch.autograd.grad((out_ref, hr_ref), (P, H), (d_out, d_hr))
    dP = torch.full_like(P, float("nan"))
    dH = torch.full_like(H, float("nan"))
    _lib.call("stg_tgcn_update_bwd_f32", P.data_ptr(), H.data_ptr(), d_out.data_ptr(), dP.data_ptr(), dH.data_ptr(), n, hid, st)
    _lib.call("stg_tgcn_reset_bwd_f32", P.data_ptr(), H.data_ptr(), d_hr.data_ptr(), dP.data_ptr(), dH.data_ptr(), n, hid, st)
    torch.testing.assert_close(dP, gP, rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(dH, gH, rtol=1e-5, atol=2e-6)


def test_tgcn_one_op_cell_odd_hidden_size(cuda):
    """Hidden size that is not a multiple of 4 (scalar kernels, unaligned column blocks in the GEMMs)."""
    from stgraph_b200.nn.pytorch import TGCN

    n, e = 150, 1400
    g, src, dst = _graph(n, e, 17, cuda)
    g.set_ndata("norm", g.degree_norm())
    torch.manual_seed(2)
    a = TGCN(5, 6, fused=False).to(cuda)
    b = TGCN(5, 6).to(cuda)
    b.load_state_dict(a.state_dict())
    xs = [torch.randn(n, 5, device=cuda) for _ in range(3)]
    res = []
    for cell in (a, b):
        H, cost = None, 0
        for x in xs:
            H = cell(g, x, None, H)
            cost = cost + (H ** 2).mean()
        cost.backward()
        res.append((H.detach(), {k: p.grad.clone() for k, p in cell.named_parameters()}))
    torch.testing.assert_close(res[0][0], res[1][0], rtol=1e-5, atol=1e-6)
    for k in res[0][1]:
        ga, gb = res[0][1][k], res[1][1][k]
        assert (ga - gb).abs().max() <= 2e-5 * ga.abs().max() + 1e-7, k
