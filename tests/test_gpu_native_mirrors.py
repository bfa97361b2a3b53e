"""GPU parity of the native-module mirrors (``stgraph.graph.dynamic.pcsr.pcsr`` / ``...gpma.gpma`` surfaces).

PCSR is checked against what the reference's own ``pcsr.cu`` builds (``tests/golden/ref_pcsr.npz``), driven through the
same calls ``PCSRGraph`` makes (``pcsr_graph.py:45-166``); GPMA against the oracle's labelled views, driven like
``GPMAGraph`` (``gpma_graph.py:56-152``).  Bit-exact integer arrays.
"""
import os

import numpy as np
import pytest

from oracle import structure as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gp():
    return np.load(os.path.join(GOLD, "ref_pcsr.npz"))


def _snapshots(gp, tag):
    flat, sizes = gp[f"{tag}/snap_edges"], gp[f"{tag}/snap_sizes"]
    cuts = np.concatenate([[0], np.cumsum(sizes)])
    return [[(int(a), int(b)) for a, b in flat[cuts[t]:cuts[t + 1]]] for t in range(len(sizes))]


def _tuples(pair):
    return list(zip(pair[0].tolist(), pair[1].tolist()))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_pcsr_mirror_equals_reference_pcsr(cuda, gp, tag):
    import stgraph_b200.compat  # noqa: F401
    from stgraph.graph.dynamic.pcsr.pcsr import PCSR, read_gpu_csr

    n = int(gp[f"{tag}/num_nodes"])
    snaps = _snapshots(gp, tag)
    ups = S.snapshot_updates(snaps)
    T = len(snaps)
    p = PCSR(n, len({e for s in snaps for e in s}))
    assert p.get_n() == n and p.edge_count == 0

    def same(prefix, rev):
        (p.build_reverse_csr if rev else p.build_csr)()
        ro, col, eid, nid = read_gpu_csr(p)
        np.testing.assert_array_equal(ro, gp[f"{prefix}/row_offset"].astype(np.int64), err_msg=prefix)
        np.testing.assert_array_equal(col, gp[f"{prefix}/column_indices"].astype(np.int64), err_msg=prefix)
        np.testing.assert_array_equal(eid, gp[f"{prefix}/eids"].astype(np.int64), err_msg=prefix)
        deg = np.diff(np.asarray(ro))
        assert sorted(nid) == list(range(n)) and np.all(np.diff(deg[np.asarray(nid)]) <= 0)
        assert all(isinstance(v, int) for v in p.get_csr_ptrs()) and len(p.get_csr_ptrs()) == 4

    for t in range(T):                      # PCSRGraph.__init__ / _update_graph_forward
        p.edge_update_list(_tuples(ups[t]["add"]), is_reverse_edge=True)
        p.edge_update_list(_tuples(ups[t]["delete"]), is_delete=True, is_reverse_edge=True)
        p.label_edges()
        same(f"{tag}/fwd/{t}", False)
        same(f"{tag}/bwd/{t}", True)
        np.testing.assert_array_equal(p.in_degrees, gp[f"{tag}/fwd/{t}/pcsr_in_degrees"])
        np.testing.assert_array_equal(p.out_degrees, gp[f"{tag}/fwd/{t}/pcsr_out_degrees"])
        assert p.edge_count == len(set(snaps[t]))
    import copy
    keep = copy.deepcopy(p)
    for t in range(T - 1, 0, -1):           # _update_graph_backward
        p.edge_update_list(_tuples(ups[t]["delete"]), is_reverse_edge=True)
        p.edge_update_list(_tuples(ups[t]["add"]), is_delete=True, is_reverse_edge=True)
        p.label_edges()
        same(f"{tag}/rewind/{t - 1}", True)
    assert keep.edge_count == len(set(snaps[T - 1]))          # the copy kept the state it was taken in
    lab = [e[2] for e in p.get_edges()]
    assert lab == list(range(1, p.edge_count + 1))


def test_gpma_mirror_matches_oracle_and_gpma_graph(cuda):
    import stgraph_b200.compat  # noqa: F401
    from stgraph.graph.dynamic.gpma import gpma as M
    from stgraph.graph.static.csr import get_array
    from stgraph_b200.graph import GPMAGraph
    from test_gpu_dynamic import _stream

    n, T = 50, 6
    snaps = _stream(n, T, base=220, churn=35, seed=5)
    keys = S.snapshot_edge_sets(snaps)
    ups = S.snapshot_updates(snaps)
    updates = {str(t): {"add": _tuples(ups[t]["add"]), "delete": _tuples(ups[t]["delete"])} for t in range(T)}
    g = M.GPMA()
    M.init_gpma(g, n)
    M.init_graph_updates(g, updates, reverse_edges=True)
    G = GPMAGraph(snaps, n)

    def arrays(is_backward, e):
        ro, col, eid, nid = M.get_csr_ptrs(g, is_backward=is_backward)
        return get_array(ro, n + 1), get_array(col, e), get_array(eid, e), get_array(nid, n)

    def check(t, with_backward):
        e = keys[t].shape[0]
        assert M.get_graph_attr(g) == [n, e]
        f = S.labelled_forward_view(keys[t], n)
        ro, col, eid, nid = arrays(False, e)
        np.testing.assert_array_equal(ro, f.row_offset)
        np.testing.assert_array_equal(col, f.column_indices)
        np.testing.assert_array_equal(eid, f.eids)
        assert np.all(np.diff(f.row_degrees[np.asarray(nid)]) <= 0)
        np.testing.assert_array_equal(M.get_out_degrees(g), f.row_degrees)      # rows = destinations (reverse_edges)
        np.testing.assert_array_equal(M.get_in_degrees(g), f.col_degrees)
        if with_backward:
            b = S.labelled_backward_view(keys[t], n)
            ro, col, eid, _ = arrays(True, e)
            np.testing.assert_array_equal(ro, b.row_offset)
            got = S.rows_as_sorted_pairs(ro, col, eid)                          # intra-row order is unspecified in the reference
            exp = S.rows_as_sorted_pairs(b.row_offset, b.column_indices, b.eids)
            for x, y in zip(got, exp):
                np.testing.assert_array_equal(x, y)

    for t in range(T):                      # GPMAGraph.__init__ / _update_graph_forward
        M.edge_update_t(g, t)
        M.label_edges(g)
        check(t, False)
        G.get_graph(t)
        np.testing.assert_array_equal(arrays(False, keys[t].shape[0])[1], G._forward_graph.column_indices.cpu().numpy())
    M.free_backward_csr(g)
    M.build_backward_csr(g)
    check(T - 1, True)
    for t in range(T - 1, 0, -1):           # _update_graph_backward
        M.free_backward_csr(g)
        M.edge_update_t(g, t, revert_update=True)
        M.label_edges(g)
        M.build_backward_csr(g)
        check(t - 1, True)
    assert sorted(M.get_gpma_edge_list(g)) == sorted((int(k >> 32), int(k & 0xFFFFFFFF)) for k in keys[0])
    with pytest.raises(RuntimeError):
        M.free_backward_csr(g)
        M.get_csr_ptrs(g, is_backward=True)
