"""Shared machinery of the two native-module mirrors (``pcsr/pcsr.py``, ``gpma/gpma.py``).

The reference's graph classes talk to two pybind modules, ``stgraph.graph.dynamic.pcsr.pcsr`` (class ``PCSR``,
``pcsr.cu:917-940``) and ``stgraph.graph.dynamic.gpma.gpma`` (class ``GPMA`` + free functions, ``gpma.cu:1435-1465``).
SURVEY.md section 8(b) keeps their names and meanings; here both are thin host-side views over the same device
state the B200 graph classes use: ONE sorted array of live ``(row << 32) | col`` keys per snapshot, updated by the
merge-path kernels of ``csrc/snapshot.cu`` and turned into labelled CSR views on demand.  Unlike the graph classes
(which know every snapshot size from preprocessing) these mirrors accept arbitrary update lists, so each update reads
one count back from the device.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import _lib
from ..static.csr import get_array  # noqa: F401  (re-exported by the mirrors)
from .dynamic_graph import _ws, build_views, keys_diff, keys_from_edges


def device_of(dev=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("stgraph_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device(dev) if dev is not None else torch.device("cuda", torch.cuda.current_device())


def keys_of(edges, reverse: bool, num_nodes: int, dev) -> torch.Tensor:
    """Sorted unique keys of an update list: a list of ``(src, dst)`` tuples, an ``[E, 2]`` array, or a device key tensor
    already in ``(row << 32) | col`` form (what ``DynamicGraph.graph_updates`` holds).  ``reverse`` = rows are ``dst``."""
    if isinstance(edges, torch.Tensor) and edges.dtype == torch.int64 and edges.dim() == 1:
        return edges.to(dev)
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    if e.shape[0] == 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    if int(e.min()) < 0 or int(e.max()) >= int(num_nodes):       # host array: free to check (see csr._edges_to_device)
        raise ValueError(f"vertex ids must lie in [0, {int(num_nodes)}): found ids in [{int(e.min())}, {int(e.max())}]")
    src = torch.from_numpy(e[:, 0].copy()).to(device=dev, dtype=torch.int32)      # .copy(): fresh, positively strided
    dst = torch.from_numpy(e[:, 1].copy()).to(device=dev, dtype=torch.int32)
    # keys_from_edges packs (dst << 32) | src, i.e. rows = second endpoint
    return keys_from_edges(src, dst, num_nodes) if reverse else keys_from_edges(dst, src, num_nodes)


def apply_update(keys: torch.Tensor, add: torch.Tensor, delete: torch.Tensor) -> torch.Tensor:
    """``(keys \\ delete) U add`` with the new size read back (one small D2H copy)."""
    n, na, nd = int(keys.shape[0]), int(add.shape[0]), int(delete.shape[0])
    if na == 0 and nd == 0:
        return keys
    dev = keys.device
    out = torch.empty(max(n + na, 1), dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    ws, nb = _ws(max(n, na, 1), dev)
    _lib.call("stg_snapshot_apply", keys.data_ptr(), n, add.data_ptr(), na, delete.data_ptr(), nd, out.data_ptr(),
              cnt.data_ptr(), ws.data_ptr(), nb, _lib.current_stream_ptr())
    return out[: int(cnt.item())].clone()


def views(keys, num_nodes: int, descending: bool, want_backward: bool):
    """(forward CSR, backward CSR or None) with 1-based labels and degree-sorted ``node_ids``."""
    return build_views(keys, num_nodes, descending, 1, want_backward, want_node_ids=True)


def csr_ptrs(csr):
    return (csr.row_offset_ptr, csr.column_indices_ptr, csr.eids_ptr, csr.node_ids_ptr)


def edges_of(keys: torch.Tensor):
    k = keys.cpu().numpy().astype(np.uint64)
    return (k >> np.uint64(32)).astype(np.int64), (k & np.uint64(0xFFFFFFFF)).astype(np.int64)


__all__ = ["apply_update", "csr_ptrs", "device_of", "edges_of", "get_array", "keys_diff", "keys_of", "views"]
