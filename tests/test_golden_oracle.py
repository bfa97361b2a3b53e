"""Pin the oracle to the reference: fixtures in tests/golden/ were produced by the reference's own
CSR builder and by the CUDA kernels its code generator emits (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import aggregate as A
from oracle import structure as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gs():
    return np.load(os.path.join(GOLD, "ref_structure.npz"))


@pytest.fixture(scope="module")
def gk():
    return np.load(os.path.join(GOLD, "ref_kernels.npz"))


def _csr(g, tag):
    return S.CsrArrays(g[f"{tag}_row_offset"], g[f"{tag}_column_indices"], g[f"{tag}_eids"], None, None, None)


def test_structure_oracle_equals_reference_csr(gs):
    src, dst, n = gs["src"], gs["dst"], int(gs["num_nodes"])
    f, b = S.forward_csr(src, dst, n), S.backward_csr(src, dst, n)
    for name in ("row_offset", "column_indices", "eids"):
        np.testing.assert_array_equal(getattr(f, name), gs[f"fwd_{name}"])
        np.testing.assert_array_equal(getattr(b, name), gs[f"bwd_{name}"])
    # CSR::out_degrees of the forward (reversed) graph is the in-degree (static_graph.py:115-117)
    np.testing.assert_array_equal(f.row_degrees, gs["fwd_out_degrees"])
    np.testing.assert_array_equal(f.col_degrees, gs["fwd_in_degrees"])
    np.testing.assert_array_equal(b.row_degrees, gs["bwd_out_degrees"])
    # node_ids: only "non-increasing row length" is contractual (std::sort tie order is unspecified)
    for ids, deg in ((gs["fwd_node_ids"], f.row_degrees), (gs["bwd_node_ids"], b.row_degrees)):
        assert sorted(ids.tolist()) == list(range(n))
        assert np.all(np.diff(deg[ids]) <= 0)
    for ids, deg in ((f.node_ids, f.row_degrees), (b.node_ids, b.row_degrees)):
        assert np.all(np.diff(deg[ids]) <= 0)
    w = gs["edge_weight_by_eid"]
    np.testing.assert_array_equal(S.weighted_in_degrees(src, dst, w, n), gs["fwd_weighted_out_degrees"].astype(np.int32))


def test_structure_oracle_sparse_graph_with_empty_rows(gs):
    src, dst, n = gs["sparse_src"], gs["sparse_dst"], int(gs["sparse_num_nodes"])
    f, b = S.forward_csr(src, dst, n), S.backward_csr(src, dst, n)
    for name in ("row_offset", "column_indices", "eids"):
        np.testing.assert_array_equal(getattr(f, name), gs[f"sparse_fwd_{name}"])
        np.testing.assert_array_equal(getattr(b, name), gs[f"sparse_bwd_{name}"])


def _t(gk, case, name):
    return torch.from_numpy(gk[f"{case}/tensor/{name}"])


def test_gcn_oracle_equals_reference_kernels(gs, gk):
    fwd, bwd = _csr(gs, "fwd"), _csr(gs, "bwd")
    case = "gcn_f16"
    h, norm = _t(gk, case, "Vhinb"), _t(gk, case, "Vnormcen").reshape(-1)
    k0, k1 = gk[f"{case}/kernels"]
    out_name, gin_name = gk[f"{case}/{k0}/rets"][0], gk[f"{case}/{k1}/args"][0]
    gout_name = gk[f"{case}/{k1}/rets"][0]
    ref_out, gin, ref_gh = _t(gk, case, out_name), _t(gk, case, gin_name), _t(gk, case, gout_name)
    A.assert_close_rel(A.gcn_forward(fwd, h, norm), ref_out, rel=2e-6, abs_terms=A.gcn_forward(fwd, h.abs(), norm))
    A.assert_close_rel(A.gcn_backward(bwd, gin, norm), ref_gh, rel=2e-6, abs_terms=A.gcn_backward(bwd, gin.abs(), norm))
    assert tuple(gk[f"{case}/{k0}/launch"]) == (10, 64, 16, 4)      # SURVEY.md A.1 launch tuple for N=40


def test_weighted_gcn_oracle_and_reference_trap_t1(gs, gk):
    """F=7: the reference launches 4 lanes per node, so columns 4..6 are never written (SURVEY.md trap T1)."""
    fwd, bwd = _csr(gs, "fwd"), _csr(gs, "bwd")
    case = "gcnw_f7"
    h, norm, w = _t(gk, case, "Vhinb"), _t(gk, case, "Vnormcen").reshape(-1), _t(gk, case, "Vedge_weight").reshape(-1)
    k0, k1 = gk[f"{case}/kernels"]
    assert tuple(gk[f"{case}/{k0}/launch"])[2] == 4
    ref_out = _t(gk, case, gk[f"{case}/{k0}/rets"][0])
    gin = _t(gk, case, gk[f"{case}/{k1}/args"][0])
    ref_gh = _t(gk, case, gk[f"{case}/{k1}/rets"][0])
    assert torch.count_nonzero(ref_out[:, 4:]) == 0 and torch.count_nonzero(ref_gh[:, 4:]) == 0
    mine = A.gcn_forward(fwd, h, norm, w)
    A.assert_close_rel(mine[:, :4], ref_out[:, :4], rel=2e-6, abs_terms=A.gcn_forward(fwd, h.abs(), norm, w)[:, :4])
    mine_b = A.gcn_backward(bwd, gin, norm, w)
    A.assert_close_rel(mine_b[:, :4], ref_gh[:, :4], rel=2e-6, abs_terms=A.gcn_backward(bwd, gin.abs(), norm, w)[:, :4])
    assert torch.count_nonzero(mine[:, 4:]) > 0                       # the oracle (and our kernels) compute all columns


@pytest.mark.parametrize("case", ["gat_h8d16", "gat_h2d4"])
def test_stock_gat_oracle_equals_reference_kernels(gs, gk, case):
    fwd = _csr(gs, "fwd")
    el, er, feat = _t(gk, case, "Velinb"), _t(gk, case, "Vercen"), _t(gk, case, "Vfeat_srcinb")
    k0, k1, k2 = gk[f"{case}/kernels"]
    rets0 = list(gk[f"{case}/{k0}/rets"])
    v3 = _t(gk, case, rets0[0])
    v4 = _t(gk, case, rets0[1])
    out_ref = _t(gk, case, gk[f"{case}/{k1}/rets"][0])
    out, o3, o4 = A.gat_stock_forward(fwd, el, er, feat)
    scale = lambda t: t.abs().mean() * torch.ones_like(t)
    A.assert_close_rel(o3, v3, rel=1e-6, abs_terms=scale(v3))
    A.assert_close_rel(o4, v4, rel=1e-6, abs_terms=scale(v4))
    A.assert_close_rel(out, out_ref, rel=3e-6, abs_terms=scale(out_ref))
    assert torch.all(v3 == 1.0)                                       # trap T2: exp(lrelu(x - x)) == 1
    # backward K2: args [..., grad], rets [d_feat, d_el, d_er] (ids sorted)
    args2, rets2 = list(gk[f"{case}/{k2}/args"]), list(gk[f"{case}/{k2}/rets"])
    known = set(rets0) | set(gk[f"{case}/{k1}/rets"]) | {"Velinb", "Vercen", "Vfeat_srcinb"} | set(rets2)
    gname = [a for a in args2 if a not in known][0]
    gout = _t(gk, case, gname)
    d_feat, d_el, d_er, m_feat, m_el, m_er = A.gat_stock_backward(fwd, el, er, feat, gout, return_mag=True)
    by_shape = {tuple(_t(gk, case, r).shape[1:]) + (r,): _t(gk, case, r) for r in rets2}
    ref_dfeat = [t for k, t in by_shape.items() if k[1] == feat.shape[2]][0]
    small = [(k[-1], t) for k, t in by_shape.items() if k[1] == 1]
    A.assert_close_rel(d_feat, ref_dfeat, rel=5e-6, abs_terms=scale(ref_dfeat), what="d_feat")
    # of the two [N,H,1] outputs one is typed SRC (d_el) and one DEST (d_er): tell them apart by value
    # rets are sorted by id: [d_feat (SRC,[H,D]), d_el (SRC,[H,1]), d_er (DEST,[H,1])]; d_er is an exact zero in
    # real arithmetic (softmax weights sum to one), the reference holds fp32 rounding noise there
    cand = [t for _, t in small]
    ref_der = min(cand, key=lambda t: float(t.abs().max()))
    ref_del = max(cand, key=lambda t: float(t.abs().max()))
    A.assert_close_rel(d_el, ref_del, rel=2e-5, abs_terms=m_el, what="d_el")
    A.assert_close_rel(d_er, ref_der, rel=2e-5, abs_terms=m_er, what="d_er")
    assert float(ref_der.abs().max()) < 1e-5 * float(m_er.max())


def test_reference_ir_matches_our_tracer(gk):
    """Op sequence of every reference unit (forward) equals what our tracer + fusion produce."""
    import re

    def ops(case, k):
        return [re.search(r"=(\w+)\(", s).group(1) for s in gk[f"{case}/{k}/program"]]

    assert ops("gcn_f16", gk["gcn_f16/kernels"][0]) == ["Mul", "AggSum", "Mul"]
    assert ops("gcn_f16", gk["gcn_f16/kernels"][1]) == ["Mul", "AggSum", "Mul"]
    assert ops("gcnw_f7", gk["gcnw_f7/kernels"][0]) == ["Mul", "Mul", "AggSum", "Mul"]
    k0, k1, _ = gk["gat_h8d16/kernels"]
    assert ops("gat_h8d16", k0) == ["Add", "Sub", "LeakyReLU", "exp", "AggSum"]
    assert ops("gat_h8d16", k1) == ["TrueDiv", "Mul", "AggSum"]


# ---------------------------------------------------------------------------------------------------------
# PCSR: fixtures recorded from the reference's own host-side PCSR (pcsr.cu compiled from where it lies,
# oracle/build_ref.py), driven like PCSRGraph drives it (pcsr_graph.py:45-166)
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gp():
    return np.load(os.path.join(GOLD, "ref_pcsr.npz"))


def _snapshots(gp, tag):
    flat, sizes = gp[f"{tag}/snap_edges"], gp[f"{tag}/snap_sizes"]
    cuts = np.concatenate([[0], np.cumsum(sizes)])
    return [[(int(a), int(b)) for a, b in flat[cuts[t]:cuts[t + 1]]] for t in range(len(sizes))]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_pcsr_oracle_equals_reference_pcsr(gp, tag):
    """labelled_forward_view / labelled_backward_view (descending rows, 1-based labels) == what the reference's PMA
    builds at every timestamp, rolling forward and rewinding; degrees too."""
    n = int(gp[f"{tag}/num_nodes"])
    snaps = _snapshots(gp, tag)
    keys = S.snapshot_edge_sets(snaps)
    for t, k in enumerate(keys):
        f = S.labelled_forward_view(k, n, descending_rows=True)
        b = S.labelled_backward_view(k, n, descending_rows=True)
        for name in ("row_offset", "column_indices", "eids"):
            np.testing.assert_array_equal(getattr(f, name), gp[f"{tag}/fwd/{t}/{name}"].astype(np.int32), err_msg=f"fwd {t} {name}")
            np.testing.assert_array_equal(getattr(b, name), gp[f"{tag}/bwd/{t}/{name}"].astype(np.int32), err_msg=f"bwd {t} {name}")
        # PCSRGraph.in_degrees() returns the reversed structure's out_degrees (pcsr_graph.py:108-114)
        np.testing.assert_array_equal(f.row_degrees, gp[f"{tag}/fwd/{t}/pcsr_out_degrees"].astype(np.int32))
        np.testing.assert_array_equal(f.col_degrees, gp[f"{tag}/fwd/{t}/pcsr_in_degrees"].astype(np.int32))
        ids = gp[f"{tag}/fwd/{t}/node_ids"].astype(np.int64)          # tie order unspecified (std::sort)
        assert sorted(ids.tolist()) == list(range(n)) and np.all(np.diff(f.row_degrees[ids]) <= 0)
        if t < len(keys) - 1:                                           # the backward roll lands on the same snapshot
            for name in ("row_offset", "column_indices", "eids"):
                np.testing.assert_array_equal(getattr(b, name), gp[f"{tag}/rewind/{t}/{name}"].astype(np.int32),
                                              err_msg=f"rewind {t} {name}")


def test_pcsr_fixture_exercises_readd_and_mass_delete(gp):
    snaps = _snapshots(gp, "a")
    sets = [set(s) for s in snaps]
    gone = (sets[0] - sets[1]) & sets[3]
    assert gone, "an edge is deleted and re-added later"
    assert min(len(s) for s in sets) <= 3 < max(len(s) for s in sets)


# ---- a9: the reference's own DynamicGraph._preprocess_graph_structure (dynamic_graph.py:56-79) -----------------
def _update_streams():
    g = np.load(os.path.join(GOLD, "ref_updates.npz"))
    for tag in ("a", "b", "c"):
        sizes, flat, n = g[f"{tag}/snap_sizes"], g[f"{tag}/snap_edges"], int(g[f"{tag}/num_nodes"])
        snaps, o = [], 0
        for sz in sizes:
            snaps.append(flat[o:o + sz])
            o += sz
        yield tag, g, snaps, n


def test_snapshot_update_oracle_equals_reference_preprocessing():
    """oracle/structure.py:snapshot_updates == the add / delete lists the reference's pure-Python preprocessing builds."""
    from oracle import structure as S

    for tag, g, snaps, n in _update_streams():
        ups = S.snapshot_updates(snaps)
        for t in range(len(snaps)):
            for kind in ("add", "delete"):
                ref = g[f"{tag}/{t}/{kind}"]
                np.testing.assert_array_equal(ref[:, 0], ups[t][kind][0], err_msg=f"{tag} t={t} {kind} src")
                np.testing.assert_array_equal(ref[:, 1], ups[t][kind][1], err_msg=f"{tag} t={t} {kind} dst")
        assert sum(g[f"{tag}/{t}/delete"].shape[0] for t in range(1, len(snaps))) > 0
