"""NaiveGraph (mirror of ``stgraph/graph/dynamic/naive/naive_graph.py:45-151``).

One forward and one backward CSR per timestamp, all resident on the GPU; switching timestamps is a
pointer swap.  ``graph_type() == "csr"``; edge ids are 0-based ranks in (dst, src) order of each
snapshot's de-duplicated edge set.  Built by GPU kernels instead of one host ``CSR::CSR`` loop per
snapshot and direction.  (Fixes reference trap T6: ``_get_cached_graph`` takes the timestamp the base
class passes.)  This class is also the structural oracle for PCSRGraph / GPMAGraph in the tests.
"""
from .dynamic_graph import DynamicGraph, build_views


class NaiveGraph(DynamicGraph):
    _descending_rows = False
    _label_base = 0

    def __init__(self, edge_list, max_num_nodes: int, device=None) -> None:
        self._snapshots = []
        super().__init__(edge_list, max_num_nodes, device)
        self._pairs = []
        for keys in self._snapshots:
            fwd, bwd = build_views(keys, self.max_num_nodes, False, 0, True, want_node_ids=True)
            fwd.prepare_hub_schedule(sync=True)
            bwd.prepare_hub_schedule(sync=True)
            self._pairs.append((fwd, bwd))
        self._snapshots = None
        self._refresh_views()

    def _keep_snapshot(self, t, keys) -> None:
        self._snapshots.append(keys)

    def graph_type(self) -> str:
        return "csr"

    def _refresh_views(self) -> None:
        self._forward_graph, self._backward_graph = self._pairs[self.current_timestamp]
        self._get_graph_csr_ptrs()

    def _ensure_views(self, need_backward: bool = False) -> None:
        self._forward_graph, self._backward_graph = self._pairs[self.current_timestamp]

    def _cache_graph(self) -> None:
        pass

    def _get_cached_graph(self, timestamp=None) -> bool:
        return False

    def _update_graph_forward(self) -> None:
        if str(self.current_timestamp + 1) not in self.graph_updates:
            raise RuntimeError("⏰ Invalid timestamp during STGraphBase.update_graph_forward()")

    def _init_reverse_graph(self) -> None:
        pass

    def _update_graph_backward(self) -> None:
        if self.current_timestamp <= 0:
            raise RuntimeError("⏰ Invalid timestamp during STGraphBase.update_graph_backward()")
