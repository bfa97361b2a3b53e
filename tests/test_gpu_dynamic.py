"""GPU parity of the dynamic graph containers (NaiveGraph / PCSRGraph / GPMAGraph).

Structure is bit-exact against oracle/structure.py (compacted views, labels, degrees, diffs); the
three containers must give the same aggregation as a StaticGraph of the same snapshot (SURVEY.md
trap T3 decision), and BPTT over snapshots through the executor's state stack must match a
torch-CPU restatement.
"""
import numpy as np
import pytest
import torch

from oracle import aggregate as A
from oracle import structure as S

pytestmark = pytest.mark.gpu


def _stream(n, t_count, base, churn, seed, dup=True):
    """Snapshot edge lists: start from `base` random edges, then drop/add `churn` per step (with some duplicates)."""
    rng = np.random.default_rng(seed)
    cur = set()
    while len(cur) < base:
        a, b = rng.integers(0, n, 2)
        if a != b:
            cur.add((int(a), int(b)))
    snaps = []
    for _ in range(t_count):
        lst = list(cur)
        rng.shuffle(lst)
        if dup and lst:
            lst = lst + lst[: max(1, len(lst) // 10)]          # duplicates are collapsed (dynamic_graph.py:58-63)
        snaps.append(lst)
        drop = rng.choice(len(cur), size=min(churn, len(cur)), replace=False)
        curl = sorted(cur)
        for i in drop:
            cur.discard(curl[i])
        while len(cur) < base:
            a, b = rng.integers(0, n, 2)
            if a != b:
                cur.add((int(a), int(b)))
    return snaps


def _classes():
    from stgraph_b200.graph import GPMAGraph, NaiveGraph, PCSRGraph

    return {"naive": NaiveGraph, "pcsr": PCSRGraph, "gpma": GPMAGraph}


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("kind", ["naive", "pcsr", "gpma"])
def test_structure_forward_roll_and_rewind(cuda, kind):
    n, T = 60, 7
    snaps = _stream(n, T, base=300, churn=40, seed=3)
    G = _classes()[kind](snaps, n)
    keys = S.snapshot_edge_sets(snaps)
    ups = S.snapshot_updates(snaps)
    assert G.graph_type() == {"naive": "csr", "pcsr": "pcsr", "gpma": "gpma"}[kind]
    for t in range(T):
        got_add = _np(G.graph_updates[str(t)]["add"]).astype(np.int64)
        got_del = _np(G.graph_updates[str(t)]["delete"]).astype(np.int64)
        np.testing.assert_array_equal(got_add & 0xFFFFFFFF, ups[t]["add"][0])
        np.testing.assert_array_equal(got_add >> 32, ups[t]["add"][1])
        np.testing.assert_array_equal(got_del >> 32, ups[t]["delete"][1])
    desc = kind == "pcsr"
    base = 0 if kind == "naive" else 1

    def check_forward(t):
        exp = S.labelled_forward_view(keys[t], n, descending_rows=desc)
        F = G._forward_graph
        np.testing.assert_array_equal(_np(F.row_offset), exp.row_offset)
        np.testing.assert_array_equal(_np(F.column_indices), exp.column_indices)
        np.testing.assert_array_equal(_np(F.eids), exp.eids - (1 - base))
        np.testing.assert_array_equal(G.in_degrees(), exp.row_degrees)
        np.testing.assert_array_equal(G.out_degrees(), exp.col_degrees)
        assert G.get_num_edges() == keys[t].shape[0] and G.get_num_nodes() == n

    def check_backward(t):
        exp = S.labelled_backward_view(keys[t], n, descending_rows=desc)
        B = G._backward_graph
        np.testing.assert_array_equal(_np(B.row_offset), exp.row_offset)
        np.testing.assert_array_equal(_np(B.column_indices), exp.column_indices)
        np.testing.assert_array_equal(_np(B.eids), exp.eids - (1 - base))

    check_forward(0)
    for t in range(T):
        G.get_graph(t)
        check_forward(t)
    # backprop state: rewind 6 -> 2, checking both views at every stop
    for t in range(T - 1, 1, -1):
        G.get_backward_graph(t)
        check_forward(t)
        check_backward(t)
    # forward again resumes from the cached state at T-1 ... here we restart instead
    G.reset_graph()
    G.get_graph(3)
    check_forward(3)
    with pytest.raises(RuntimeError):
        G.get_graph(1)


def test_three_containers_agree_with_static_graph(cuda):
    from stgraph_b200 import kernels
    from stgraph_b200.graph import StaticGraph

    n, T = 200, 4
    snaps = _stream(n, T, base=2500, churn=300, seed=5)
    x = torch.randn(n, 20, device=cuda)
    graphs = {k: c(snaps, n) for k, c in _classes().items()}
    keys = S.snapshot_edge_sets(snaps)
    for t in range(T):
        src, dst = (keys[t] & 0xFFFFFFFF), (keys[t] >> 32)
        sg = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
        w = torch.rand(keys[t].shape[0], device=cuda) + 0.1      # indexed by edge id = rank in (dst,src) order
        norm = sg.degree_norm().reshape(-1)
        ref_f = kernels.agg_scaled_sum(sg.fwd_view(), x, norm, w, norm)
        ref_b = kernels.agg_scaled_sum(sg.bwd_view(), x, norm, w, norm)
        exp = A.scaled_sum(_np(sg._forward_graph.row_offset), _np(sg._forward_graph.column_indices),
                           _np(sg._forward_graph.eids), x.cpu(), norm.cpu(), w.cpu(), norm.cpu())
        mag = A.scaled_sum(_np(sg._forward_graph.row_offset), _np(sg._forward_graph.column_indices),
                           _np(sg._forward_graph.eids), x.cpu().abs(), norm.cpu(), w.cpu(), norm.cpu())
        mag_b = kernels.agg_scaled_sum(sg.bwd_view(), x.abs(), norm, w, norm).cpu()
        A.assert_close_rel(ref_f.cpu(), exp, rel=1e-5, abs_terms=mag)
        for k, G in graphs.items():
            G.get_graph(t)
            torch.testing.assert_close(G.degree_norm().reshape(-1), norm, rtol=0, atol=0)
            out_f = kernels.agg_scaled_sum(G.fwd_view(), x, norm, w, norm)
            # PCSR walks every row back to front: same terms, different fp32 summation order
            A.assert_close_rel(out_f.cpu(), ref_f.cpu(), rel=2e-6, abs_terms=mag, what=f"{k} fwd t={t}")
            out_b = kernels.agg_scaled_sum(G.bwd_view(), x, norm, w, norm)
            A.assert_close_rel(out_b.cpu(), ref_b.cpu(), rel=2e-6, abs_terms=mag_b, what=f"{k} bwd t={t}")


def test_empty_and_single_edge_snapshots(cuda):
    from stgraph_b200.graph import GPMAGraph, PCSRGraph

    snaps = [[(0, 1)], [], [(2, 3), (0, 1)], [(2, 3)]]
    for cls in (GPMAGraph, PCSRGraph):
        G = cls(snaps, 5)
        sizes = []
        for t in range(4):
            G.get_graph(t)
            sizes.append(G.get_num_edges())
            assert int(G._forward_graph.row_offset[-1]) == sizes[-1]
        assert sizes == [1, 0, 2, 1]
        G.get_backward_graph(1)
        assert G.get_num_edges() == 0 and int(G._backward_graph.row_offset[-1]) == 0


@pytest.mark.parametrize("kind", ["naive", "pcsr", "gpma"])
def test_bptt_over_snapshots_through_state_stack(cuda, kind):
    """GCNConv applied on snapshots t=0..T-1, one backward at the end: the executor rewinds the graph."""
    from stgraph_b200.nn.pytorch import GCNConv

    n, T, F = 80, 5, 8
    snaps = _stream(n, T, base=500, churn=80, seed=9, dup=False)
    keys = S.snapshot_edge_sets(snaps)
    G = _classes()[kind](snaps, n)
    torch.manual_seed(0)
    layer = GCNConv(F, F).to(cuda)
    x = torch.randn(n, F, device=cuda, requires_grad=True)
    h = x
    cost = 0
    for t in range(T):
        G.get_graph(t)
        G.set_ndata("norm", G.degree_norm())
        h = torch.tanh(layer(G, h))
        cost = cost + (h ** 2).mean()
    cost.backward()
    assert G.current_timestamp == 0 and G._is_backprop_state
    # torch-CPU restatement
    xc = x.detach().cpu().double().requires_grad_(True)
    W = layer.weight.detach().cpu().double().requires_grad_(True)
    b = layer.bias.detach().cpu().double().requires_grad_(True)
    hc, cost_c = xc, 0
    for t in range(T):
        src = torch.from_numpy((keys[t] & 0xFFFFFFFF).astype(np.int64))
        dst = torch.from_numpy((keys[t] >> 32).astype(np.int64))
        deg = torch.zeros(n, dtype=torch.float64).index_add(0, dst, torch.ones(dst.shape[0], dtype=torch.float64))
        nr = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg)).float().double().reshape(-1, 1)
        tt = hc @ W
        hc = torch.tanh(torch.zeros(n, F, dtype=torch.float64).index_add(0, dst, tt[src] * nr[src]) * nr + b)
        cost_c = cost_c + (hc ** 2).mean()
    cost_c.backward()
    sc = lambda t: t.abs().mean() * torch.ones_like(t) + 1e-12
    A.assert_close_rel(h.detach().cpu(), hc.detach(), rel=2e-5, abs_terms=torch.ones_like(hc))
    A.assert_close_rel(x.grad.cpu(), xc.grad, rel=1e-4, abs_terms=sc(xc.grad), what="dX")
    A.assert_close_rel(layer.weight.grad.cpu(), W.grad, rel=1e-4, abs_terms=sc(W.grad), what="dW")
    A.assert_close_rel(layer.bias.grad.cpu(), b.grad, rel=1e-4, abs_terms=sc(b.grad), what="db")
    # a second epoch after reset_graph works (per-epoch reset, dynamic-temporal-tgcn/seastar/train.py)
    G.reset_graph()
    G.get_graph(0)
    assert G.current_timestamp == 0 and not G._is_backprop_state


@pytest.mark.parametrize("kind", ["naive", "pcsr", "gpma"])
def test_tgcn_one_op_cell_on_dynamic_snapshots(cuda, kind):
    """The one-op TGCN cell on snapshots: the views are captured at forward time (no rewind in backward); values and
    gradients equal the reference-structured cell on a NaiveGraph of the same snapshots."""
    from stgraph_b200.graph import NaiveGraph
    from stgraph_b200.nn.pytorch import TGCN

    n, T = 120, 4
    snaps = _stream(n, T, base=900, churn=150, seed=4, dup=False)
    torch.manual_seed(5)
    cells = [TGCN(8, 16, fused=f).to(cuda) for f in (False, None)]
    cells[1].load_state_dict(cells[0].state_dict())
    xs = [torch.randn(n, 8, device=cuda) for _ in range(T)]
    res = []
    for cell, G in zip(cells, (NaiveGraph(snaps, n), _classes()[kind](snaps, n))):
        H, cost = None, 0
        for t, x in enumerate(xs):
            G.get_graph(t)
            G.set_ndata("norm", G.degree_norm())
            H = cell(G, x, None, H)
            cost = cost + (H ** 2).mean()
        cost.backward()
        res.append((H.detach(), {k: p.grad.clone() for k, p in cell.named_parameters()}))
    torch.testing.assert_close(res[0][0], res[1][0], rtol=1e-5, atol=1e-6)
    for k in res[0][1]:
        ga, gb = res[0][1][k], res[1][1][k]
        assert (ga - gb).abs().max() <= 2e-5 * ga.abs().max() + 1e-7, k
