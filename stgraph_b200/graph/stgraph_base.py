"""Abstract graph interface handed to the compiled vertex programs.

Same contract as ``stgraph/graph/stgraph_base.py:46-90``: eight integers holding
raw device addresses of the forward (in-edge) and backward (out-edge) CSR arrays
plus ``get_num_nodes / get_num_edges / get_ndata / set_ndata / graph_type``.
Here the arrays are torch int32 tensors (torch's caching allocator owns the
memory; the reference leaks raw ``cudaMalloc`` blocks) and the pointer fields are
their ``data_ptr()``.  ``fwd_view()`` / ``bwd_view()`` package the same arrays as
the ``StgCsrView`` struct of the C ABI.
"""
from __future__ import annotations

from abc import ABC, abstractmethod


class STGraphBase(ABC):
    def __init__(self) -> None:
        self._ndata = {}
        self._forward_graph = None
        self._backward_graph = None

        self.fwd_row_offset_ptr = None
        self.fwd_column_indices_ptr = None
        self.fwd_eids_ptr = None
        self.fwd_node_ids_ptr = None

        self.bwd_row_offset_ptr = None
        self.bwd_column_indices_ptr = None
        self.bwd_eids_ptr = None
        self.bwd_node_ids_ptr = None

    @abstractmethod
    def _get_graph_csr_ptrs(self) -> None:
        """Refresh the eight pointer fields from the current snapshot."""

    @abstractmethod
    def get_num_nodes(self) -> int:
        """Number of nodes of the current snapshot."""

    @abstractmethod
    def get_num_edges(self) -> int:
        """Number of edges of the current snapshot."""

    @abstractmethod
    def get_ndata(self, field: str):
        """Node data registered under ``field`` (None if absent)."""

    @abstractmethod
    def set_ndata(self, field: str, val) -> None:
        """Register node data."""

    @abstractmethod
    def graph_type(self) -> str:
        """One of csr / csr_unsorted / pcsr / pcsr_unsorted / gpma / gpma_unsorted (code_gen.py:94-109)."""

    # ---- C-ABI views (new) -------------------------------------------------
    @abstractmethod
    def fwd_view(self):
        """``StgCsrView`` of the in-edge CSR of the current snapshot."""

    @abstractmethod
    def bwd_view(self):
        """``StgCsrView`` of the out-edge CSR of the current snapshot."""
