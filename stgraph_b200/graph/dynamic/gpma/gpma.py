"""Mirror of the reference's pybind module ``stgraph.graph.dynamic.gpma.gpma`` (``gpma.cu:1435-1465``).

Same names and argument meanings as the reference's free functions over a ``GPMA`` handle (``gpma.cu:947-1282``).  The
reference keeps a gapped array of ``(row << 32) + col`` keys with lazily deleted slots; here the handle holds the
sorted array of LIVE keys (a packed memory array with zero gaps, ``_native_compat.py``), which is what every consumer
observes: ``get_csr_ptrs`` hands out the compacted forward CSR (labels = 1 + rank among live keys,
``gpma.cu:1121-1146``) and, after ``build_backward_csr``, its transpose carrying those labels (``gpma.cu:1165-1231``;
rows sorted by column here, unordered in the reference).  The reference's ``gpma.cu`` cannot be built for sm_100
(device-side ``cudaDeviceSynchronize``, SURVEY.md trap T4), so this mirror is checked against ``oracle/structure.py``
and against ``PCSR`` (same labelled view, ascending rows).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native_compat as C


class GPMA:
    def __init__(self) -> None:
        self.num_nodes = 0
        self.keys = None
        self.updates = {}
        self.reverse_edges = False
        self.fwd = None
        self.bwd = None

    def __copy__(self):
        c = GPMA()
        c.__dict__.update(self.__dict__)      # key arrays / update lists are immutable: shared, like thrust copies of values
        return c

    def __deepcopy__(self, memo):
        return self.__copy__()


def init_gpma(gpma: GPMA, num_nodes: int) -> None:
    gpma.num_nodes = int(num_nodes)
    gpma.keys = torch.empty(0, dtype=torch.int64, device=C.device_of())
    gpma.fwd = gpma.bwd = None


def init_graph_updates(gpma: GPMA, updates, reverse_edges: bool = False) -> None:
    """``updates[str(t)] = {"add": [...], "delete": [...]}`` (tuples, ``[E,2]`` arrays or device key tensors)."""
    dev = gpma.keys.device
    gpma.reverse_edges = bool(reverse_edges)
    gpma.updates = {
        str(t): {k: C.keys_of(u[k], gpma.reverse_edges, gpma.num_nodes, dev) for k in ("add", "delete")}
        for t, u in updates.items()
    }


def edge_update_t(gpma: GPMA, timestamp: int, revert_update: bool = False) -> None:
    u = gpma.updates[str(timestamp)]
    add, delete = (u["delete"], u["add"]) if revert_update else (u["add"], u["delete"])
    gpma.keys = C.apply_update(gpma.keys, add, delete)
    gpma.fwd = gpma.bwd = None


def label_edges(gpma: GPMA) -> None:
    """Build the labelled forward view (labels = 1 + rank among the live keys)."""
    gpma.fwd, _ = C.views(gpma.keys, gpma.num_nodes, descending=False, want_backward=False)


def build_backward_csr(gpma: GPMA) -> None:
    gpma.fwd, gpma.bwd = C.views(gpma.keys, gpma.num_nodes, descending=False, want_backward=True)


def free_backward_csr(gpma: GPMA) -> None:
    gpma.bwd = None


def get_csr_ptrs(gpma: GPMA, is_backward: bool = False):
    csr = gpma.bwd if is_backward else gpma.fwd
    if csr is None:
        raise RuntimeError("get_csr_ptrs() before label_edges() / build_backward_csr()")
    return C.csr_ptrs(csr)


def _degrees(gpma: GPMA):
    rows, cols = C.edges_of(gpma.keys)
    n = gpma.num_nodes
    return np.bincount(rows, minlength=n).astype(np.uint32), np.bincount(cols, minlength=n).astype(np.uint32)


def get_out_degrees(gpma: GPMA):
    """Row lengths of the structure (with ``reverse_edges`` the graph's in-degrees, ``gpma_graph.py:101-103``)."""
    return _degrees(gpma)[0].tolist()


def get_in_degrees(gpma: GPMA):
    return _degrees(gpma)[1].tolist()


# ---- logging APIs (gpma.cu:1451-1455) --------------------------------------------------------------------
def get_graph_attr(gpma: GPMA):
    return [gpma.num_nodes, int(gpma.keys.shape[0])]


def get_gpma_edge_list(gpma: GPMA):
    rows, cols = C.edges_of(gpma.keys)
    return list(zip(rows.tolist(), cols.tolist()))


def get_reverse_csr_edge_list(gpma: GPMA):
    if gpma.bwd is None:
        raise RuntimeError("get_reverse_csr_edge_list() before build_backward_csr()")
    ro = gpma.bwd.row_offset.cpu().numpy()
    col = gpma.bwd.column_indices.cpu().numpy()
    rows = np.repeat(np.arange(gpma.num_nodes), np.diff(ro))
    return list(zip(rows.tolist(), col.tolist()))


def get_node_ids(gpma: GPMA):
    if gpma.fwd is None:
        label_edges(gpma)
    return gpma.fwd.node_ids.cpu().tolist()


def print_gpma_info(gpma: GPMA, node: int) -> None:
    rows, cols = C.edges_of(gpma.keys)
    print(f"node {node}: neighbours {cols[rows == node].tolist()}")
