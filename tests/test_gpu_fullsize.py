"""Parity at the BASELINE.json sizes: the CUDA path against the ORACLE (not against itself).

* config 5 (2,449,029 V / 61,859,140 E): forward and backward aggregation, F = 100 and 47, plain and packed entry
  points, checked against oracle/aggregate.py on a row sample that holds every hub row (> 1024 edges) plus 50,000
  random rows (the oracle needs ~seconds for the ~4 M sampled edges; the whole graph would need an hour);
* config 3 (169,343 V / 1,166,243 E, rows up to 8.5 K edges): the fused edge-softmax kernels against the closed
  form and the stock GATConv program (VM kernel) against the closed form of its degenerate trace, whole graph;
* dynamic containers at 100 K vertices / 1 M edges / 10 snapshots: structure bit-exact, forward roll and rewind;
* AggMax / AggMin / AggMean through the VM kernel against oracle/ir_interp.py.

Tolerance (SURVEY.md trap T9): |x - ref| <= 1e-5 * sum|terms| for features, bit-exact for structure.
"""
import numpy as np
import pytest
import torch

from oracle import aggregate as A
from oracle import structure as S

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ config 5
@pytest.fixture(scope="module")
def config5(cuda):
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.utils import synthetic

    d = synthetic.products_shaped(seed=0, device=cuda)
    n = d["num_nodes"]
    g = StaticGraph(torch.stack([d["src"], d["dst"]], 1), None, n)
    assert n == 2449029 and g.get_num_edges() == 61859140
    return g, g.degree_norm().reshape(-1).contiguous()


def _oracle_rows(csr, rows, x_cpu, ns_cpu, rs_cpu, block_edges=1_000_000):
    """(oracle rows, sum of |terms|) for the sampled rows of one CSR direction (fp64 accumulate), in edge blocks."""
    ro = csr.row_offset.long()
    beg, end = ro[rows], ro[rows + 1]
    deg = (end - beg)
    sub_ro = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=rows.device)
    sub_ro[1:] = torch.cumsum(deg, 0)
    idx = torch.repeat_interleave(beg - sub_ro[:-1], deg) + torch.arange(int(sub_ro[-1]), device=rows.device)
    cols = csr.column_indices[idx].cpu().numpy()
    sub_ro = sub_ro.cpu().numpy()
    rows_c = rows.cpu()
    out, mag = [], []
    a = 0
    while a < rows.numel():
        b = int(np.searchsorted(sub_ro, sub_ro[a] + block_edges, side="left"))
        b = min(max(b, a + 1), rows.numel())
        e0, e1 = int(sub_ro[a]), int(sub_ro[b])
        rs = rs_cpu[rows_c[a:b]]
        out.append(A.scaled_sum(sub_ro[a:b + 1] - e0, cols[e0:e1], None, x_cpu, ns_cpu, None, rs))
        mag.append(A.scaled_sum(sub_ro[a:b + 1] - e0, cols[e0:e1], None, x_cpu.abs(), ns_cpu, None, rs))
        a = b
    return torch.cat(out), torch.cat(mag)


@pytest.mark.parametrize("feat", [100, 47])
def test_config5_aggregation_against_oracle_on_hub_and_random_rows(cuda, config5, feat):
    from stgraph_b200 import kernels

    g, norm = config5
    n = g.get_num_nodes()
    gen = torch.Generator(device=cuda).manual_seed(feat)
    x = torch.randn(n, feat, device=cuda, generator=gen)
    x_cpu, norm_cpu = x.cpu(), norm.cpu()
    for name, csr in (("fwd", g._forward_graph), ("bwd", g._backward_graph)):
        plain = kernels.agg_scaled_sum(csr.view(), x, norm, None, norm)
        packed = kernels.agg_scaled_sum_graph(csr, x, norm, None, norm)
        assert csr._meta_cache, "the static graph did not take the packed path"
        assert torch.equal(plain, packed), f"{name}: packed and plain entry points differ"
        deg = csr.row_degrees.long()
        hubs = torch.nonzero(deg > 1024).reshape(-1)
        assert hubs.numel() > 500
        rnd = torch.randperm(n, device=cuda, generator=gen)[:50000]
        rows = torch.unique(torch.cat([hubs, rnd, torch.tensor([0, n - 1], device=cuda)]))
        ref, mag = _oracle_rows(csr, rows, x_cpu, norm_cpu, norm_cpu)
        got = plain[rows].cpu()
        err = (got.double() - ref.double()).abs()
        bound = 1e-5 * mag.double() + 1e-30
        assert bool((err <= bound).all()), (name, feat, float((err / bound).max()))
        del plain, packed


# ------------------------------------------------------------------------------------------------ config 3
@pytest.fixture(scope="module")
def config3(cuda):
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.utils import synthetic

    d = synthetic.arxiv_shaped(seed=0, device=cuda)
    n = d["num_nodes"]
    g = StaticGraph(torch.stack([d["src"], d["dst"]], 1), None, n)
    src, dst = d["src"].cpu().numpy(), d["dst"].cpu().numpy()
    assert n == 169343 and src.shape[0] == 1166243 and int(g.in_degrees_tensor().max()) > 8000
    return g, S.forward_csr(src, dst, n), n


def test_config3_fused_edge_softmax_against_closed_form(cuda, config3):
    from stgraph_b200.ops_gat import gat_edge_softmax_aggregate

    g, f, n = config3
    heads, dim = 8, 16
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
    A.assert_close_rel(out.detach().cpu(), ref, rel=1e-5, abs_terms=sc(ref), what="config3 fused out")
    # hub sources / destinations sum thousands of terms (fp32 rounding grows with sum|terms|, not with |result|):
    # bound by a multiple of the mean magnitude of the gradient
    A.assert_close_rel(feat.grad.cpu(), d_feat, rel=1e-4, abs_terms=sc(d_feat), what="config3 fused d_feat")
    A.assert_close_rel(el.grad.cpu().reshape(n, heads), d_el, rel=2e-4, abs_terms=sc(d_el), what="config3 fused d_el")
    A.assert_close_rel(er.grad.cpu().reshape(n, heads), d_er, rel=2e-4, abs_terms=sc(d_el), what="config3 fused d_er")


def test_config3_stock_gatconv_against_closed_form(cuda, config3):
    """The reference-faithful stock program (trap T2) through the VM kernel, rows of 8.5 K edges included."""
    from stgraph_b200.nn.pytorch import GATConv

    g, f, n = config3
    heads, dim = 8, 16
    torch.manual_seed(9)
    layer = GATConv(128, dim, heads).to(cuda)
    x = torch.randn(n, 128, device=cuda, requires_grad=True)
    gout = torch.randn(n, heads, dim, device=cuda)
    out = layer(g, x)
    out.backward(gout)
    feat = layer.fc(x.detach()).view(-1, heads, dim).detach().cpu()
    el = (feat * layer.attn_l.detach().cpu()).sum(-1, keepdim=True)
    er = (feat * layer.attn_r.detach().cpu()).sum(-1, keepdim=True)
    ref, _, _ = A.gat_stock_forward(f, el, er, feat)
    sc = lambda t: t.abs().mean() * torch.ones_like(t) + 1e-12
    A.assert_close_rel(out.detach().cpu(), ref, rel=1e-5, abs_terms=sc(ref), what="config3 stock out")
    d_feat, d_el, d_er = A.gat_stock_backward(f, el, er, feat, gout.cpu())
    featg = (x.detach().cpu().double() @ layer.fc.weight.detach().cpu().double().t()).view(-1, heads, dim).requires_grad_(True)
    al, ar = layer.attn_l.detach().cpu().double(), layer.attn_r.detach().cpu().double()
    elg, erg = (featg * al).sum(-1, keepdim=True), (featg * ar).sum(-1, keepdim=True)
    torch.autograd.backward([featg, elg, erg], [d_feat.double(), d_el.double(), d_er.double()])
    gx_ref = featg.grad.reshape(n, -1) @ layer.fc.weight.detach().cpu().double()
    A.assert_close_rel(x.grad.cpu(), gx_ref, rel=2e-4, abs_terms=sc(gx_ref), what="config3 stock dX")


# ------------------------------------------------------------------------------------------------ dynamic, large
def _large_stream(n, base, churn, T, seed):
    """T snapshots as [E,2] int32 arrays: `base` distinct edges, `churn` of them replaced per step (numpy, vectorised)."""
    rng = np.random.default_rng(seed)

    def fresh(k, exclude):
        out = np.zeros(0, dtype=np.int64)
        while out.shape[0] < k:
            s = rng.integers(0, n, 2 * k)
            d = rng.integers(0, n, 2 * k)
            key = (d[s != d] << 32) | s[s != d]
            key = np.setdiff1d(np.unique(key), exclude, assume_unique=True)
            out = np.union1d(out, key)
        return rng.permutation(out)[:k]

    cur = fresh(base, np.zeros(0, dtype=np.int64))
    snaps = []
    for t in range(T):
        edges = np.stack([cur & 0xFFFFFFFF, cur >> 32], 1).astype(np.int32)
        dup = edges[rng.integers(0, edges.shape[0], edges.shape[0] // 20)]         # duplicates collapse (dynamic_graph.py:58-63)
        snaps.append(rng.permutation(np.concatenate([edges, dup])))
        keep = np.sort(rng.permutation(cur)[churn:])
        cur = np.union1d(keep, fresh(churn, cur))
    return snaps


@pytest.mark.parametrize("kind", ["naive", "pcsr", "gpma"])
def test_dynamic_structure_at_100k_vertices_1m_edges(cuda, kind):
    from stgraph_b200.graph import GPMAGraph, NaiveGraph, PCSRGraph

    n, base, churn, T = 100_000, 1_000_000, 50_000, 10
    snaps = _large_stream(n, base, churn, T, seed=21)
    keys = S.snapshot_edge_sets(snaps)
    ups = S.snapshot_updates(snaps)
    G = {"naive": NaiveGraph, "pcsr": PCSRGraph, "gpma": GPMAGraph}[kind](snaps, n)
    desc = kind == "pcsr"
    base_lab = 0 if kind == "naive" else 1
    _np = lambda t: t.cpu().numpy()
    for t in range(T):                                     # a9: add / delete lists, bit-exact
        got_add = _np(G.graph_updates[str(t)]["add"]).astype(np.int64)
        got_del = _np(G.graph_updates[str(t)]["delete"]).astype(np.int64)
        np.testing.assert_array_equal(got_add & 0xFFFFFFFF, ups[t]["add"][0])
        np.testing.assert_array_equal(got_add >> 32, ups[t]["add"][1])
        np.testing.assert_array_equal(got_del & 0xFFFFFFFF, ups[t]["delete"][0])
        np.testing.assert_array_equal(got_del >> 32, ups[t]["delete"][1])
        if t > 0:
            assert ups[t]["add"][0].shape[0] == churn and ups[t]["delete"][0].shape[0] == churn

    def check(t, backward):
        exp = S.labelled_forward_view(keys[t], n, descending_rows=desc)
        F = G._forward_graph
        np.testing.assert_array_equal(_np(F.row_offset), exp.row_offset)
        np.testing.assert_array_equal(_np(F.column_indices), exp.column_indices)
        np.testing.assert_array_equal(_np(F.eids), exp.eids - (1 - base_lab))
        np.testing.assert_array_equal(G.in_degrees(), exp.row_degrees)
        np.testing.assert_array_equal(G.out_degrees(), exp.col_degrees)
        assert G.get_num_edges() == keys[t].shape[0] == base
        if backward:
            expb = S.labelled_backward_view(keys[t], n, descending_rows=desc)
            B = G._backward_graph
            np.testing.assert_array_equal(_np(B.row_offset), expb.row_offset)
            np.testing.assert_array_equal(_np(B.column_indices), expb.column_indices)
            np.testing.assert_array_equal(_np(B.eids), expb.eids - (1 - base_lab))

    for t in range(T):
        G.get_graph(t)
        check(t, backward=False)
    for t in range(T - 1, 0, -1):
        G.get_backward_graph(t)
        check(t, backward=True)


# ------------------------------------------------------------------------------------------------ AggMax / Min / Mean
@pytest.mark.parametrize("kind", ["max", "min", "mean"])
@pytest.mark.parametrize("feat", [3, 16, 100])
def test_agg_max_min_mean_on_gpu_against_ir_interpreter(cuda, kind, feat):
    """The VM kernel's max / min accumulators and the mean read-out (hub rows merged by accumulator kind) against
    oracle/ir_interp.py; AggMax also backward (BackwardAMax + AggSum), against torch autograd of the closed form."""
    from oracle.ir_interp import Interp
    from stgraph_b200.compiler import STGraph
    from stgraph_b200.compiler.backend.pytorch.torch_callback import STGraphBackendTorch
    from stgraph_b200.compiler.op.agg import agg_max, agg_mean, agg_min
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.graph.static import csr as csr_mod

    agg = {"max": agg_max, "min": agg_min, "mean": agg_mean}[kind]

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.stgraph = STGraph(STGraphBackendTorch())

    m = M()

    @m.stgraph.compile(gnn_module=m)
    def f(v):
        return agg([nb.h * v.s for nb in v.innbs])

    n = 6000
    rng = np.random.default_rng(17)
    key = rng.choice(n * n, size=40000, replace=False)
    src, dst = (key // n).astype(np.int64), (key % n).astype(np.int64)
    hub = 2 * csr_mod.HUB_THRESHOLD + 11                       # vertex 0 is a hub destination: split row, merged by kind
    extra = np.arange(1, hub + 1) % n
    extra = extra[extra != 0]
    k = np.unique(np.concatenate([src, extra]) * n + np.concatenate([dst, np.zeros_like(extra)]))
    src, dst = (k // n).astype(np.int32), (k % n).astype(np.int32)
    g = StaticGraph(torch.from_numpy(np.stack([src, dst], 1)), None, n)
    tg = torch.Generator().manual_seed(feat)
    h = torch.randn(n, feat, generator=tg).to(cuda).requires_grad_(kind == "max")
    s = (torch.rand(n, 1, generator=tg) + 0.5).to(cuda)
    out = f(g=g, n_feats={"h": h, "s": s})
    order = np.lexsort((src, dst))
    env = Interp(src[order], dst[order], n).run_units(f.forward_units, {"Vhinb": h.detach().cpu(), "Vscen": s.cpu()})
    ref = env[f.forward_units[0].unit_rets()[0].id]
    has = torch.from_numpy(np.bincount(dst, minlength=n) > 0)
    got = out.detach().cpu().double()
    assert int(has.sum()) < n                                  # some rows have no in-edge: the accumulator's start value
    tol = 1e-6 if kind != "mean" else 1e-5
    torch.testing.assert_close(got[has], ref[has], rtol=tol, atol=tol)
    if kind == "mean":
        assert bool((got[~has] == 0).all())
    else:
        assert bool(torch.isinf(got[~has]).all())
    if kind == "max":
        gout = torch.randn(n, feat, generator=tg).to(cuda)
        gout[~has.to(cuda)] = 0
        out.backward(gout)
        # closed form: the gradient of a row's max goes to the neighbour(s) that attain it, times s[v]
        h64 = h.detach().cpu().double().requires_grad_(True)
        srct, dstt = torch.from_numpy(src[order].astype(np.int64)), torch.from_numpy(dst[order].astype(np.int64))
        vals = h64[srct] * s.cpu().double()[dstt]
        mx = torch.full((n, feat), -float("inf"), dtype=torch.float64).index_reduce(0, dstt, vals.detach(), "amax")
        hit = (vals.detach() == mx[dstt]).double()
        gh = torch.zeros(n, feat, dtype=torch.float64).index_add(0, srct, hit * gout.cpu().double()[dstt] * s.cpu().double()[dstt])
        torch.testing.assert_close(h.grad.cpu().double(), gh, rtol=1e-5, atol=1e-6)
