"""Mirror of the reference's pybind module ``stgraph.graph.dynamic.pcsr.pcsr`` (``pcsr.cu:917-940``).

Same class, methods, attributes and meanings as the reference's host-side packed CSR (``pcsr.cu:325-891``); the state
is a sorted key array on the GPU instead of a packed memory array on the host (see ``_native_compat.py``).  What the
reference defines and this reproduces bit for bit (``tests/golden/ref_pcsr.npz`` was recorded from the reference's own
``pcsr.cu``): ``build_csr()`` lists every row back to front (descending neighbour id, ``pcsr.cu:842-855``),
``label_edges()`` numbers the live edges 1.. in (row, neighbour) order (``pcsr.cu:748-760``), ``build_reverse_csr()``
is the transpose carrying those labels (``pcsr.cu:783-840``), ``in_degrees`` / ``out_degrees`` count by the caller's
``dst`` / ``src`` after the optional ``is_reverse_edge`` swap (``pcsr.cu:762-781``).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native_compat as C


class PCSR:
    def __init__(self, init_n: int, max_edge_count: int, device=None) -> None:
        self._dev = C.device_of(device)
        self._n = int(init_n)
        self.max_edge_count = int(max_edge_count)
        self._keys = torch.empty(0, dtype=torch.int64, device=self._dev)     # (src << 32) | dst after the swap
        self._csr = None
        self._labelled = False

    # ---- attributes of the pybind class (def_readwrite: plain lists) ----------------------------------
    @property
    def edge_count(self) -> int:
        return int(self._keys.shape[0])

    @property
    def out_degrees(self):
        src, _ = C.edges_of(self._keys)
        return np.bincount(src, minlength=self._n).astype(np.uint32).tolist()

    @property
    def in_degrees(self):
        _, dst = C.edges_of(self._keys)
        return np.bincount(dst, minlength=self._n).astype(np.uint32).tolist()

    def get_n(self) -> int:
        return self._n

    # ---- exposed APIs (pcsr.cu:313-320) -----------------------------------------------------------------
    def edge_update_list(self, edge_list, is_delete: bool = False, is_reverse_edge: bool = False) -> None:
        """Insert (or delete) the edges of ``edge_list``; ``is_reverse_edge`` swaps the endpoints first."""
        if isinstance(edge_list, torch.Tensor) and edge_list.dtype == torch.int64 and edge_list.dim() == 1:
            upd = edge_list.to(self._dev)          # already (row << 32) | col keys (DynamicGraph.graph_updates)
        else:
            if isinstance(edge_list, torch.Tensor):
                edge_list = edge_list.cpu().numpy()
            upd = self._pack(edge_list, bool(is_reverse_edge))
        empty = torch.empty(0, dtype=torch.int64, device=self._dev)
        self._keys = C.apply_update(self._keys, empty, upd) if is_delete else C.apply_update(self._keys, upd, empty)
        self._csr = None

    def _pack(self, edge_list, swap: bool) -> torch.Tensor:
        e = np.asarray(edge_list, dtype=np.int64).reshape(-1, 2)
        if swap:
            e = e[:, ::-1].copy()
        # rows = first column after the swap: keys_of(reverse=False) packs (first << 32) | second
        return C.keys_of(e, reverse=False, num_nodes=self._n, dev=self._dev)

    def label_edges(self) -> None:
        """Labels are 1 + rank among the live edges in (row, neighbour) order: implicit in the sorted key array."""
        self._labelled = True

    def build_csr(self) -> float:
        self._csr, _ = C.views(self._keys, self._n, descending=True, want_backward=False)
        return 0.0          # the reference returns its pinned->device copy time; nothing is copied here

    def build_reverse_csr(self) -> float:
        _, self._csr = C.views(self._keys, self._n, descending=True, want_backward=True)
        return 0.0

    def get_csr_ptrs(self):
        if self._csr is None:
            raise RuntimeError("PCSR.get_csr_ptrs() before build_csr() / build_reverse_csr()")
        return C.csr_ptrs(self._csr)

    def get_edges(self):
        src, dst = C.edges_of(self._keys)
        lab = np.arange(1, src.shape[0] + 1) if self._labelled else np.ones(src.shape[0], dtype=np.int64)
        return list(zip(src.tolist(), dst.tolist(), lab.tolist()))

    def move_pinned_to_gpu(self) -> None:
        """No-op: the CSR arrays are built on the device."""

    # key arrays are never modified in place, so copies share them (the reference copies host vectors and SHARES its
    # raw device pointers, pcsr.cu:936-939)
    def __copy__(self):
        c = PCSR.__new__(PCSR)
        c.__dict__.update(self.__dict__)
        return c

    def __deepcopy__(self, memo):
        return self.__copy__()


def read_gpu_csr(pcsr: PCSR):
    """``[row_offset, column_indices, eids, node_ids]`` of the last built CSR as host lists (``pcsr.cu:897-915``)."""
    c = pcsr._csr
    if c is None:
        raise RuntimeError("read_gpu_csr() before build_csr() / build_reverse_csr()")
    return [c.row_offset.cpu().tolist(), c.column_indices.cpu().tolist(), c.eids.cpu().tolist(), c.node_ids.cpu().tolist()]
