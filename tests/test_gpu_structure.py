"""GPU parity: CSR/CSC build, degrees, eids -- bit-exact against the structure oracle."""
import numpy as np
import pytest
import torch

from oracle import structure as S

pytestmark = pytest.mark.gpu


def _graph(src, dst, n, w=None):
    from stgraph_b200.graph import StaticGraph

    edges = torch.stack([torch.as_tensor(src), torch.as_tensor(dst)], dim=1)
    return StaticGraph(edges, w, n)


def _check_static(g, src, dst, n):
    f = S.forward_csr(src, dst, n)
    b = S.backward_csr(src, dst, n)
    F, B = g._forward_graph, g._backward_graph
    np.testing.assert_array_equal(F.row_offset.cpu().numpy(), f.row_offset)
    np.testing.assert_array_equal(F.column_indices.cpu().numpy(), f.column_indices)
    np.testing.assert_array_equal(F.eids.cpu().numpy(), f.eids)
    np.testing.assert_array_equal(B.row_offset.cpu().numpy(), b.row_offset)
    np.testing.assert_array_equal(B.column_indices.cpu().numpy(), b.column_indices)
    np.testing.assert_array_equal(B.eids.cpu().numpy(), b.eids)
    np.testing.assert_array_equal(F.node_ids.cpu().numpy(), f.node_ids)
    np.testing.assert_array_equal(B.node_ids.cpu().numpy(), b.node_ids)
    np.testing.assert_array_equal(g.in_degrees(), f.row_degrees)
    np.testing.assert_array_equal(g.out_degrees(), f.col_degrees)
    assert g.in_degrees().dtype == np.int32


@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (5, 0, 0), (7, 12, 1), (100, 1000, 2), (2708, 10556, 0), (5000, 200000, 3)])
def test_csr_build_matches_oracle(cuda, n, e, seed):
    rng = np.random.default_rng(seed)
    if e:
        key = rng.choice(n * n, size=min(e, n * n), replace=False)
        src, dst = (key // n).astype(np.int32), (key % n).astype(np.int32)
    else:
        src = dst = np.zeros(0, dtype=np.int32)
    g = _graph(src, dst, n)
    _check_static(g, src, dst, n)
    assert g.get_num_edges() == src.shape[0]
    assert g.get_num_nodes() == n
    assert g.graph_type() == "csr_unsorted"


def test_pointer_fields_and_get_array(cuda):
    from stgraph_b200.graph.static.csr import get_array

    rng = np.random.default_rng(5)
    n = 50
    key = rng.choice(n * n, size=300, replace=False)
    src, dst = (key // n).astype(np.int32), (key % n).astype(np.int32)
    g = _graph(src, dst, n)
    f = S.forward_csr(src, dst, n)
    assert get_array(g.fwd_row_offset_ptr, n + 1) == f.row_offset.tolist()
    assert get_array(g.fwd_column_indices_ptr, 300) == f.column_indices.tolist()
    assert get_array(g.bwd_eids_ptr, 300) == S.backward_csr(src, dst, n).eids.tolist()
    for name in ("fwd_row_offset_ptr", "fwd_column_indices_ptr", "fwd_eids_ptr", "fwd_node_ids_ptr",
                 "bwd_row_offset_ptr", "bwd_column_indices_ptr", "bwd_eids_ptr", "bwd_node_ids_ptr"):
        assert isinstance(getattr(g, name), int) and getattr(g, name) != 0


def test_duplicates_and_caller_list_untouched(cuda):
    edges = [(2, 0), (0, 1), (1, 0), (2, 1), (0, 2), (0, 1)]
    snapshot = list(edges)
    from stgraph_b200.graph import StaticGraph

    g = StaticGraph(edges, [1.0] * 6, 3)
    assert edges == snapshot                      # trap T8: the reference sorts the caller's list in place
    assert g.get_num_edges() == 5                 # len(set(edge_list)), static_graph.py:49
    src = np.array([e[0] for e in edges]); dst = np.array([e[1] for e in edges])
    _check_static(g, src, dst, 3)


def test_weighted_in_degrees_reference_probe(cuda):
    """SURVEY.md trap T8 probe: weights 1..5 in post-sort eid order give [3,7,5]."""
    from stgraph_b200.graph import StaticGraph

    edges = [(2, 0), (0, 1), (1, 0), (2, 1), (0, 2)]
    g = StaticGraph(edges, [1, 2, 3, 4, 5], 3)
    assert g.weighted_in_degrees().tolist() == [3, 7, 5]
    assert g.weighted_in_degrees().dtype == np.int32
    src = np.array([e[0] for e in edges]); dst = np.array([e[1] for e in edges])
    w = np.array([1, 2, 3, 4, 5], dtype=np.float32)
    np.testing.assert_array_equal(g.weighted_in_degrees(), S.weighted_in_degrees(src, dst, w, 3))


def test_weighted_degree_bit_exact_fp32(cuda):
    rng = np.random.default_rng(11)
    n = 300
    key = rng.choice(n * n, size=6000, replace=False)
    src, dst = (key // n).astype(np.int32), (key % n).astype(np.int32)
    w = rng.uniform(0.1, 1.0, size=6000).astype(np.float32)
    g = _graph(src, dst, n, w)
    order = np.lexsort((src, dst))
    acc = np.zeros(n, dtype=np.float32)
    for i in range(6000):
        d = dst[order][i]
        acc[d] = np.float32(acc[d] + w[i])
    got = g._forward_graph.weighted_row_degrees.cpu().numpy()
    np.testing.assert_array_equal(got.view(np.uint32), acc.view(np.uint32))


def test_degree_norm(cuda):
    rng = np.random.default_rng(3)
    n = 200
    key = rng.choice(n * n, size=900, replace=False)
    src, dst = (key // n).astype(np.int32), (key % n).astype(np.int32)
    g = _graph(src, dst, n)
    deg = torch.from_numpy(g.in_degrees())
    ref = torch.pow(deg, -0.5)
    ref[torch.isinf(ref)] = 0
    got = g.degree_norm().cpu().reshape(-1)
    assert got.shape == (n,)
    torch.testing.assert_close(got, ref, rtol=2e-7, atol=0)


def test_out_of_range_vertex_ids_are_rejected(cuda):
    """ids are validated against num_nodes before the int32 cast (the sort covers valid key bits only)."""
    from stgraph_b200.graph import GPMAGraph, StaticGraph

    with pytest.raises(ValueError, match="vertex ids"):
        StaticGraph([(0, 1), (2, 5)], None, 5)                    # num_nodes given as the max id
    with pytest.raises(ValueError, match="vertex ids"):
        StaticGraph(torch.tensor([[0, 1], [-1, 2]]), None, 5)
    with pytest.raises(ValueError, match="vertex ids"):
        StaticGraph(torch.tensor([[0, 1], [(1 << 32) + 1, 2]], dtype=torch.int64), None, 5)     # would wrap to 1 as int32
    with pytest.raises(ValueError, match="vertex ids"):
        GPMAGraph([[(0, 1)], [(0, 7)]], 5)
    with pytest.raises(TypeError):
        StaticGraph(torch.tensor([[0.0, 1.0]]), None, 5)
    g = StaticGraph([(0, 1), (2, 4)], None, 5)
    assert g.get_num_edges() == 2
