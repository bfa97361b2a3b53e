"""CPU check: the mirrors of the reference's three pybind modules export the reference's names
(``csr.cu:181-200``, ``pcsr.cu:917-940``, ``gpma.cu:1435-1465``); no GPU call is made."""
import inspect


def test_csr_module_surface():
    import stgraph_b200.compat  # noqa: F401
    from stgraph.graph.static import csr

    for name in ("CSR", "get_array"):
        assert hasattr(csr, name), name
    for attr in ("row_offset_ptr", "column_indices_ptr", "eids_ptr", "node_ids_ptr", "out_degrees", "in_degrees",
                 "weighted_out_degrees"):
        assert hasattr(csr.CSR, attr), attr


def test_pcsr_module_surface():
    import stgraph_b200.compat  # noqa: F401
    from stgraph.graph.dynamic.pcsr import pcsr

    assert callable(pcsr.read_gpu_csr)
    for name in ("get_n", "edge_update_list", "label_edges", "get_edges", "build_csr", "build_reverse_csr", "get_csr_ptrs",
                 "__copy__", "__deepcopy__", "in_degrees", "out_degrees", "edge_count"):
        assert hasattr(pcsr.PCSR, name), name
    sig = inspect.signature(pcsr.PCSR.edge_update_list)
    assert list(sig.parameters)[1:] == ["edge_list", "is_delete", "is_reverse_edge"]
    assert sig.parameters["is_delete"].default is False and sig.parameters["is_reverse_edge"].default is False
    assert list(inspect.signature(pcsr.PCSR.__init__).parameters)[1:3] == ["init_n", "max_edge_count"]


def test_gpma_module_surface():
    import stgraph_b200.compat  # noqa: F401
    from stgraph.graph.dynamic.gpma import gpma

    expected = {
        "init_gpma": ["gpma", "num_nodes"],
        "init_graph_updates": ["gpma", "updates", "reverse_edges"],
        "edge_update_t": ["gpma", "timestamp", "revert_update"],
        "label_edges": ["gpma"],
        "build_backward_csr": ["gpma"],
        "free_backward_csr": ["gpma"],
        "get_csr_ptrs": ["gpma", "is_backward"],
        "get_in_degrees": ["gpma"],
        "get_out_degrees": ["gpma"],
        "get_graph_attr": ["gpma"],
        "get_gpma_edge_list": ["gpma"],
        "get_reverse_csr_edge_list": ["gpma"],
        "get_node_ids": ["gpma"],
    }
    for name, params in expected.items():
        fn = getattr(gpma, name)
        assert list(inspect.signature(fn).parameters) == params, name
    assert inspect.signature(gpma.init_graph_updates).parameters["reverse_edges"].default is False
    assert inspect.signature(gpma.edge_update_t).parameters["revert_update"].default is False
    assert inspect.signature(gpma.get_csr_ptrs).parameters["is_backward"].default is False
    g = gpma.GPMA()
    import copy
    assert isinstance(copy.deepcopy(g), gpma.GPMA)
