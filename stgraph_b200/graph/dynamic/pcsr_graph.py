"""PCSRGraph (mirror of ``stgraph/graph/dynamic/pcsr/pcsr_graph.py:28-166``).

The reference keeps a Wheatman-Xu packed CSR on the HOST (``pcsr.cu:404-717``), relabels and
rebuilds a dense CSR into pinned memory per timestamp and copies it to the GPU
(``pcsr.cu:748-883``).  Its compacted CSR lists every row back to front (descending neighbour id,
``pcsr.cu:842-855``) with 1-based labels; that view is reproduced bit for bit from the sorted key
array by GPU kernels (``csrc/snapshot.cu``, ``descending_rows=1``).  Pinned against the reference itself:
``tests/golden/ref_pcsr.npz`` holds what the reference's ``pcsr.cu`` (compiled from where it lies,
``oracle/build_ref.py``) builds at every timestamp of two streams, rolling forward and backward
(``tests/test_gpu_golden.py::test_pcsr_graph_equals_reference_pcsr``).
"""
from .dynamic_graph import KeyedDynamicGraph


class PCSRGraph(KeyedDynamicGraph):
    _descending_rows = True
    _label_base = 1

    def __init__(self, edge_list, max_num_nodes: int, device=None) -> None:
        super().__init__(edge_list, max_num_nodes, device)
        self._get_max_num_edges()

    def _get_max_num_edges(self) -> None:
        """Number of distinct edges ever added (``pcsr_graph.py:98-106``)."""
        import torch

        allk = torch.cat([self.graph_updates[str(t)]["add"] for t in range(len(self.graph_updates))])
        self.max_num_edges = int(torch.unique(allk).shape[0])

    def graph_type(self) -> str:
        return "pcsr"
