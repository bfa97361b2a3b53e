"""Dynamic graphs: snapshots, per-timestamp diffs, forward roll / backward rewind.

API kept from ``stgraph/graph/dynamic/dynamic_graph.py:14-188``: the constructor takes one edge list
per timestamp; ``graph_updates[str(t)] = {"add", "delete"}`` are the set differences of consecutive
snapshots, each ordered by (dst, src) (``dynamic_graph.py:56-79``); ``get_graph(t)`` rolls forward,
``get_backward_graph(t)`` caches the forward state, switches to the backprop state and rolls back
(``90-128``); ``get_num_nodes()`` is ``max_num_nodes`` for every t, ``get_num_edges()`` the size of the
snapshot's edge *set*; node data is stored per timestamp (``138-153``).

Everything the reference does with Python ``set`` objects of tuples and host loops runs as GPU
kernels here (``csrc/snapshot.cu``): snapshot keys = sort + unique, diffs = sorted set difference,
updates = merge-path insert/delete, views = boundary fill + transpose.  Sizes are read back once,
at construction; rolling forward or backward inside a training loop never synchronises.
"""
from __future__ import annotations

import time
from abc import abstractmethod

import numpy as np
import torch

from ... import _lib
from ..static.csr import CSR, _edges_to_device
from ..stgraph_base import STGraphBase


class _Updates(dict):
    """``{"add": keys, "delete": keys}`` with device key tensors; ``edges(kind)`` gives host (src, dst) tuples."""

    def edges(self, kind):
        k = self[kind].cpu().numpy().astype(np.uint64)
        return list(zip((k & 0xFFFFFFFF).astype(np.int64).tolist(), (k >> 32).astype(np.int64).tolist()))


def _ws(items, device):
    nbytes = _lib.load().stg_snapshot_workspace_bytes(int(items))
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def keys_from_edges(src, dst, num_nodes):
    """Sorted, de-duplicated ``(dst<<32)|src`` keys of an edge list (device int64 tensor)."""
    n = int(src.shape[0])
    dev = src.device
    out = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    ws, nb = _ws(n, dev)
    _lib.call("stg_snapshot_keys_from_edges", src.data_ptr(), dst.data_ptr(), n, int(num_nodes), out.data_ptr(),
              cnt.data_ptr(), ws.data_ptr(), nb, _lib.current_stream_ptr())
    return out[: int(cnt.item())].clone()


def keys_diff(a, b):
    """``a \\ b`` for sorted unique key tensors."""
    na, nbk = int(a.shape[0]), int(b.shape[0])
    dev = a.device
    out = torch.empty(max(na, 1), dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    ws, nb = _ws(na, dev)
    _lib.call("stg_snapshot_diff", a.data_ptr(), na, b.data_ptr(), nbk, out.data_ptr(), cnt.data_ptr(), ws.data_ptr(), nb,
              _lib.current_stream_ptr())
    return out[: int(cnt.item())].clone()


def keys_apply(keys, add, delete, new_count):
    """``(keys \\ delete) U add``; ``new_count`` is known from preprocessing, so nothing is read back."""
    n, na, nd = int(keys.shape[0]), int(add.shape[0]), int(delete.shape[0])
    dev = keys.device
    out = torch.empty(max(n + na, 1), dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    ws, nb = _ws(max(n, na), dev)
    _lib.call("stg_snapshot_apply", keys.data_ptr(), n, add.data_ptr(), na, delete.data_ptr(), nd, out.data_ptr(),
              cnt.data_ptr(), ws.data_ptr(), nb, _lib.current_stream_ptr())
    return out[:new_count]


def build_views(keys, num_nodes, descending, label_base, want_backward, want_node_ids=False):
    """Labelled CSR views of a snapshot -> (forward CSR, backward CSR or None)."""
    n = int(keys.shape[0])
    dev = keys.device
    i32 = dict(dtype=torch.int32, device=dev)
    f_ro, f_col, f_lab = torch.empty(num_nodes + 1, **i32), torch.empty(n, **i32), torch.empty(n, **i32)
    in_deg = torch.empty(num_nodes, **i32)
    f_nid = torch.empty(num_nodes, **i32) if want_node_ids else None
    if want_backward:
        b_ro, b_col, b_lab = torch.empty(num_nodes + 1, **i32), torch.empty(n, **i32), torch.empty(n, **i32)
        out_deg = torch.empty(num_nodes, **i32)
        b_nid = torch.empty(num_nodes, **i32) if want_node_ids else None
    else:
        b_ro = b_col = b_lab = out_deg = b_nid = None
    ws, nb = _ws(max(n, num_nodes), dev)
    _lib.call("stg_snapshot_views", keys.data_ptr(), n, int(num_nodes), 1 if descending else 0, int(label_base),
              f_ro.data_ptr(), _lib.ptr(f_col), _lib.ptr(f_lab), _lib.ptr(f_nid),
              _lib.ptr(b_ro), _lib.ptr(b_col), _lib.ptr(b_lab), _lib.ptr(b_nid),
              in_deg.data_ptr(), _lib.ptr(out_deg), ws.data_ptr(), nb, _lib.current_stream_ptr())
    identity = not descending
    fwd = CSR(f_ro, f_col, f_lab, f_nid, in_deg, None, eid_base=label_base, eids_identity=identity, num_edges=n)
    bwd = None
    if want_backward:
        bwd = CSR(b_ro, b_col, b_lab, b_nid, out_deg, in_deg, eid_base=label_base, eids_identity=False, num_edges=n)
        fwd.col_degrees = out_deg
    return fwd, bwd


class DynamicGraph(STGraphBase):
    #: rows emitted back to front (PCSR) / label base (1 for PCSR & GPMA, 0 for the per-snapshot CSR)
    _descending_rows = False
    _label_base = 1

    def __init__(self, edge_list, max_num_nodes: int, device=None) -> None:
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("stgraph_b200 dynamic graphs need a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.graph_updates = {}
        self.max_num_nodes = int(max_num_nodes)
        self._is_backprop_state = False
        self.current_timestamp = 0
        self.get_fwd_graph_time = 0
        self.get_bwd_graph_time = 0
        self.move_to_gpu_time = 0
        self._snapshot_sizes = []
        self._preprocess_graph_structure(edge_list)
        self.graph_attr = {str(t): (self.max_num_nodes, self._snapshot_sizes[t]) for t in range(len(self._snapshot_sizes))}
        self._hub_sync = False

    # ------------------------------------------------------------- preprocessing
    def _preprocess_graph_structure(self, edge_list) -> None:
        """Per-timestamp edge sets and their add/delete differences, on the GPU (``dynamic_graph.py:56-79``)."""
        prev = torch.empty(0, dtype=torch.int64, device=self.device)
        self._base_keys = None
        for t in range(len(edge_list)):
            src, dst = _edges_to_device(edge_list[t], self.device, self.max_num_nodes)
            cur = keys_from_edges(src, dst, self.max_num_nodes)
            if t == 0:
                add, dele = cur, torch.empty(0, dtype=torch.int64, device=self.device)
                self._base_keys = cur
            else:
                add, dele = keys_diff(cur, prev), keys_diff(prev, cur)
            self.graph_updates[str(t)] = _Updates(add=add, delete=dele)
            self._snapshot_sizes.append(int(cur.shape[0]))
            self._keep_snapshot(t, cur)
            prev = cur

    def _keep_snapshot(self, t, keys) -> None:
        """Hook for subclasses that keep every snapshot resident (NaiveGraph)."""

    # ------------------------------------------------------------------ API
    def reset_graph(self) -> None:
        self._get_cached_graph("base")
        self.current_timestamp = 0
        self._is_backprop_state = False
        self.get_fwd_graph_time = 0
        self.get_bwd_graph_time = 0
        self.move_to_gpu_time = 0

    def get_graph(self, timestamp: int) -> None:
        t0 = time.time()
        self._is_backprop_state = False
        if timestamp < self.current_timestamp:
            raise RuntimeError("⏰ Invalid timestamp during STGraphBase.update_graph_forward()")
        if self._get_cached_graph(timestamp - 1):
            self.current_timestamp = timestamp - 1
        while self.current_timestamp < timestamp:
            self._update_graph_forward()
            self.current_timestamp += 1
        self._refresh_views()
        self.get_fwd_graph_time += time.time() - t0

    def get_backward_graph(self, timestamp: int) -> None:
        t0 = time.time()
        if not self._is_backprop_state:
            self._cache_graph()
            self._is_backprop_state = True
            self._init_reverse_graph()
        if timestamp > self.current_timestamp:
            raise RuntimeError("⏰ Invalid timestamp during STGraphBase.update_graph_backward()")
        while self.current_timestamp > timestamp:
            self._update_graph_backward()
            self.current_timestamp -= 1
        self._refresh_views()
        self.get_bwd_graph_time += time.time() - t0

    def get_num_nodes(self) -> int:
        return self.graph_attr[str(self.current_timestamp)][0]

    def get_num_edges(self) -> int:
        return self.graph_attr[str(self.current_timestamp)][1]

    def get_ndata(self, field: str):
        return self._ndata.get(str(self.current_timestamp), {}).get(field)

    def set_ndata(self, field: str, val) -> None:
        self._ndata.setdefault(str(self.current_timestamp), {})[field] = val

    def in_degrees(self) -> np.ndarray:
        return self.in_degrees_tensor().cpu().numpy().astype("int32")

    def out_degrees(self) -> np.ndarray:
        return self.out_degrees_tensor().cpu().numpy().astype("int32")

    def in_degrees_tensor(self) -> torch.Tensor:
        self._ensure_views()
        return self._forward_graph.row_degrees

    def out_degrees_tensor(self) -> torch.Tensor:
        self._ensure_views(need_backward=True)
        return self._backward_graph.row_degrees

    def degree_norm(self) -> torch.Tensor:
        """``in_degree^-0.5`` (inf -> 0) as ``[N,1]`` on the device.

        The reference benchmark copies the degrees to the host, runs ``torch.pow`` on the CPU and
        copies back every timestamp (``benchmarking/dynamic-temporal-tgcn/seastar/train.py:213-218``).
        """
        deg = self.in_degrees_tensor()
        norm = torch.empty(self.max_num_nodes, dtype=torch.float32, device=self.device)
        _lib.call("stg_degree_norm_f32", deg.data_ptr(), self.max_num_nodes, norm.data_ptr(), _lib.current_stream_ptr())
        return norm.unsqueeze(1)

    # ------------------------------------------------------------ C-ABI views
    def fwd_view(self):
        self._ensure_views()
        return self._forward_graph.view()

    def bwd_view(self):
        self._ensure_views(need_backward=True)
        return self._backward_graph.view()

    def _get_graph_csr_ptrs(self) -> None:
        f, b = self._forward_graph, self._backward_graph
        if f is not None:
            self.fwd_row_offset_ptr = f.row_offset_ptr
            self.fwd_column_indices_ptr = f.column_indices_ptr
            self.fwd_eids_ptr = f.eids_ptr
            self.fwd_node_ids_ptr = f.node_ids_ptr if f.node_ids is not None else None
        if b is not None:
            self.bwd_row_offset_ptr = b.row_offset_ptr
            self.bwd_column_indices_ptr = b.column_indices_ptr
            self.bwd_eids_ptr = b.eids_ptr
            self.bwd_node_ids_ptr = b.node_ids_ptr if b.node_ids is not None else None

    # -------------------------------------------------- subclass responsibilities
    @abstractmethod
    def _refresh_views(self) -> None:
        """Make the CSR views (and pointer fields) describe ``current_timestamp``."""

    @abstractmethod
    def _ensure_views(self, need_backward: bool = False) -> None:
        pass

    @abstractmethod
    def _cache_graph(self) -> None:
        pass

    @abstractmethod
    def _get_cached_graph(self, timestamp) -> bool:
        pass

    @abstractmethod
    def _update_graph_forward(self) -> None:
        pass

    @abstractmethod
    def _init_reverse_graph(self) -> None:
        pass

    @abstractmethod
    def _update_graph_backward(self) -> None:
        pass


class KeyedDynamicGraph(DynamicGraph):
    """Shared machinery of PCSRGraph / GPMAGraph: the live snapshot is one sorted key array."""

    def __init__(self, edge_list, max_num_nodes: int, device=None) -> None:
        super().__init__(edge_list, max_num_nodes, device)
        self._keys = self._base_keys
        self._views_valid = False
        self._views_have_backward = False
        self.graph_cache = {"base": self._base_keys}
        self._refresh_views()

    def _cache_graph(self) -> None:
        self.graph_cache[str(self.current_timestamp)] = self._keys     # key arrays are never modified in place

    def _get_cached_graph(self, timestamp) -> bool:
        if timestamp == "base":
            self._keys = self.graph_cache["base"]
            self._views_valid = False
            return True
        if str(timestamp) in self.graph_cache:
            self._keys = self.graph_cache.pop(str(timestamp))
            self._views_valid = False
            return True
        return False

    def _update_graph_forward(self) -> None:
        t = self.current_timestamp + 1
        if str(t) not in self.graph_updates:
            raise RuntimeError("⏰ Invalid timestamp during STGraphBase.update_graph_forward()")
        up = self.graph_updates[str(t)]
        self._keys = keys_apply(self._keys, up["add"], up["delete"], self._snapshot_sizes[t])
        self._views_valid = False

    def _init_reverse_graph(self) -> None:
        self._views_valid = False

    def _update_graph_backward(self) -> None:
        t = self.current_timestamp
        if t <= 0:
            raise RuntimeError("⏰ Invalid timestamp during STGraphBase.update_graph_backward()")
        up = self.graph_updates[str(t)]
        self._keys = keys_apply(self._keys, up["delete"], up["add"], self._snapshot_sizes[t - 1])
        self._views_valid = False

    def _refresh_views(self) -> None:
        self._views_valid = False
        self._ensure_views(need_backward=self._is_backprop_state)

    def _ensure_views(self, need_backward: bool = False) -> None:
        if self._views_valid and (self._views_have_backward or not need_backward):
            return
        want_bwd = need_backward or self._is_backprop_state
        # views already handed out (StgCsrView structs hold raw pointers) must stay valid until the snapshot
        # changes: park the objects being replaced instead of dropping them
        self._retired = [self._forward_graph, self._backward_graph] if self._views_valid else []
        self._forward_graph, bwd = build_views(self._keys, self.max_num_nodes, self._descending_rows, self._label_base,
                                               want_bwd)
        self._backward_graph = bwd
        self._forward_graph.prepare_hub_schedule(sync=False)
        if bwd is not None:
            bwd.prepare_hub_schedule(sync=False)
        self._views_valid = True
        self._views_have_backward = want_bwd
        self._get_graph_csr_ptrs()
