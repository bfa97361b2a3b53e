"""GPU-built CSR arrays (replacement of the pybind ``csr`` module).

The reference's ``CSR`` (``stgraph/graph/static/csr.cu:35-66,181-200``) is a host
loop over a pre-sorted Python list followed by four ``cudaMemcpy``; it exposes
``row_offset_ptr / column_indices_ptr / eids_ptr / node_ids_ptr`` and the host
vectors ``out_degrees / in_degrees / weighted_out_degrees``.  Here both
directions are produced by one call into ``stg_csr_build`` (radix sort + boundary
fill on the GPU) and live in torch int32 tensors.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from ... import _lib

#: rows longer than this are processed by the block-per-row kernel (agg.cu)
HUB_THRESHOLD = int(os.environ.get("STG_HUB_THRESHOLD", "1024"))
#: STG_PACK_META=0 keeps static graphs on the plain (unpacked) aggregation kernel (A/B runs)
PACK_META = os.environ.get("STG_PACK_META", "1") != "0"
#: a CSR that had to repack more often than this (its scales are per-call temporaries) falls back to the plain kernel
MAX_META_REPACKS = 16


def _edges_to_device(edge_list, device, num_nodes=None):
    """Accept a list of (src,dst) tuples, an [E,2] numpy/torch array or a (src,dst) pair of arrays.

    With ``num_nodes`` every id is checked against ``[0, num_nodes)`` BEFORE the cast to int32 (one host read-back,
    at construction): the radix sort of ``stg_csr_build`` / ``stg_snapshot_keys_from_edges`` only covers the key
    bits of valid ids, so an id >= num_nodes (``num_nodes`` passed as max id instead of max id + 1), a negative id
    or an int64 id that does not fit would otherwise corrupt the row-offset fill silently."""
    if isinstance(edge_list, tuple) and len(edge_list) == 2 and not np.isscalar(edge_list[0]) \
            and len(np.shape(edge_list[0])) == 1 and len(edge_list[0]) != 2:
        src, dst = edge_list
        src = torch.as_tensor(src)
        dst = torch.as_tensor(dst)
    else:
        if isinstance(edge_list, torch.Tensor):
            e = edge_list
        else:
            e = torch.from_numpy(np.asarray(edge_list, dtype=np.int64).reshape(-1, 2))
        e = e.reshape(-1, 2)
        src, dst = e[:, 0], e[:, 1]
    if src.is_floating_point() or dst.is_floating_point() or src.dtype == torch.bool:
        raise TypeError("vertex ids must be integers")
    src, dst = src.to(device=device), dst.to(device=device)
    if num_nodes is not None and src.numel() > 0:
        lo = int(torch.minimum(src.min(), dst.min()))
        hi = int(torch.maximum(src.max(), dst.max()))
        if lo < 0 or hi >= int(num_nodes):
            raise ValueError(f"vertex ids must lie in [0, {int(num_nodes)}): found ids in [{lo}, {hi}]")
    src = src.to(dtype=torch.int32).contiguous()
    dst = dst.to(dtype=torch.int32).contiguous()
    return src, dst


class CSR:
    """One direction of a graph in CSR form, resident on the GPU.

    Attribute names follow the pybind class (``csr.cu:181-200``); the ``*_ptr``
    fields are raw device addresses, the tensors keep the memory alive.
    ``out_degrees`` = row lengths, ``in_degrees`` = column occurrence counts.
    """

    def __init__(self, row_offset, column_indices, eids, node_ids, row_degrees, col_degrees,
                 eid_base=0, eids_identity=False, num_edges=None):
        self.row_offset = row_offset
        self.column_indices = column_indices
        self.eids = eids
        self.node_ids = node_ids
        self.row_degrees = row_degrees
        self.col_degrees = col_degrees
        self.eid_base = eid_base
        self.eids_identity = eids_identity
        self.num_nodes = int(row_offset.shape[0] - 1)
        self.num_edges = int(column_indices.shape[0]) if num_edges is None else int(num_edges)
        self.weighted_row_degrees = None
        self._hub_rows = None
        self._hub_count = None
        self._hub_enabled = None
        self._view = None
        #: static graphs cache the packed {col, scale} array per scale tensor (see packed_meta)
        self.pack_enabled = False
        self._meta_cache = []
        self._meta_misses = 0

    # -- reference-compatible surface ---------------------------------------
    @property
    def row_offset_ptr(self):
        return self.row_offset.data_ptr()

    @property
    def column_indices_ptr(self):
        return self.column_indices.data_ptr()

    @property
    def eids_ptr(self):
        return self.eids.data_ptr()

    @property
    def node_ids_ptr(self):
        return self.node_ids.data_ptr()

    @property
    def out_degrees(self):
        return self.row_degrees.cpu().tolist()

    @property
    def in_degrees(self):
        return self.col_degrees.cpu().tolist()

    @property
    def weighted_out_degrees(self):
        if self.weighted_row_degrees is None:
            return [0.0] * self.num_nodes
        return self.weighted_row_degrees.cpu().tolist()

    # -- packed edge metadata (stg_csr_pack_edge_meta_f32) -------------------------
    def packed_meta(self, nbr_scale, edge_scale):
        """``{col, nbr_scale[col] * edge_scale[eid]}`` per CSR slot for these scale tensors, packed on first use.

        The cache key is (address, in-place version counter) of each scale tensor, and the entry keeps the
        tensors alive, so an address cannot be recycled under it; an in-place update through torch bumps the
        version and repacks.  Returns None (-> plain kernel) when there is nothing to pack, when packing is
        disabled, or on a miss during CUDA-graph capture (the cache must not own capture-pool memory).
        A write that torch does not see -- ``t.data = ...`` / ``t.data.copy_()`` on some builds, a raw-pointer kernel --
        does not move the version counter: call :meth:`invalidate_packed_meta` after it.
        """
        if not PACK_META or (nbr_scale is None and edge_scale is None) or self.num_edges == 0:
            return None
        key = tuple((t.data_ptr(), t._version, t.numel()) if t is not None else None for t in (nbr_scale, edge_scale))
        for k, _, meta in self._meta_cache:
            if k == key:
                return meta
        if torch.cuda.is_current_stream_capturing():
            return None
        self._meta_misses += 1
        if self._meta_misses > MAX_META_REPACKS:   # scales that change every call (an intermediate tensor): stop packing
            self.pack_enabled = False
            self._meta_cache = []
            return None
        from ... import kernels

        meta = kernels.pack_edge_meta(self.view(), nbr_scale, edge_scale, device=self.row_offset.device)
        self._meta_cache = [(key, (nbr_scale, edge_scale), meta)] + self._meta_cache[:1]   # at most two live entries
        return meta

    def invalidate_packed_meta(self):
        """Drop the packed ``{col, scale}`` arrays (and re-enable packing): the next aggregation packs again from the
        current contents of its scale tensors."""
        if self._meta_misses > MAX_META_REPACKS:      # packing had been switched off by scales that changed every call
            self.pack_enabled = bool(PACK_META)
        self._meta_cache = []
        self._meta_misses = 0

    # -- C-ABI view ------------------------------------------------------------
    def prepare_hub_schedule(self, sync: bool = True):
        """Collect rows longer than HUB_THRESHOLD for the block-per-row kernel.

        ``sync=True`` (static graphs) reads the count once so that graphs without
        hubs skip the extra launch; ``sync=False`` keeps everything on the stream.
        """
        n, e = self.num_nodes, self.num_edges
        cap = e // max(HUB_THRESHOLD, 1) + 1
        dev = self.row_offset.device
        self._hub_rows = torch.empty(cap, dtype=torch.int32, device=dev)
        self._hub_count = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.call("stg_csr_hub_rows", self.row_offset.data_ptr(), n, HUB_THRESHOLD, self._hub_rows.data_ptr(),
                  cap, self._hub_count.data_ptr(), _lib.current_stream_ptr())
        self._hub_enabled = True
        self._hub_sync = sync
        if sync:
            self._hub_enabled = int(self._hub_count.item()) > 0
        self._view = None
        self._alt_views = {}

    def view(self) -> _lib.StgCsrView:
        if self._view is None:
            v = _lib.StgCsrView()
            v.row_offset = self.row_offset.data_ptr()
            v.column_indices = self.column_indices.data_ptr()
            v.eids = self.eids.data_ptr() if self.eids is not None else None
            v.node_ids = self.node_ids.data_ptr() if self.node_ids is not None else None
            v.num_nodes = self.num_nodes
            v.num_edges = self.num_edges
            v.eid_base = self.eid_base
            v.eids_identity = 1 if self.eids_identity else 0
            if self._hub_enabled:
                v.hub_rows = self._hub_rows.data_ptr()
                v.hub_count = self._hub_count.data_ptr()
                v.hub_threshold = HUB_THRESHOLD
                v.hub_capacity = int(self._hub_rows.shape[0])
            else:
                v.hub_rows = None
                v.hub_count = None
                v.hub_threshold = 0
                v.hub_capacity = 0
            # global row queue of the aggregation kernel (StgCsrView::work_queue): owned by this direction of the graph
            self._work_queue = torch.zeros(2, dtype=torch.int32, device=self.row_offset.device)
            v.work_queue = self._work_queue.data_ptr()
            self._view = v
        return self._view


def _alt_view(self, threshold: int) -> _lib.StgCsrView:
    """The same CSR with a hub-row list for another threshold (the attention kernels split rows much earlier than
    the plain sum: their per-edge work is a longer dependent chain, ``ops_gat.GAT_HUB_THRESHOLD``).  Cached."""
    threshold = int(threshold)
    if threshold <= 0 or threshold == HUB_THRESHOLD:
        return self.view()
    alt = getattr(self, "_alt_views", None)
    if alt is None:
        alt = self._alt_views = {}
    if threshold not in alt:
        base = self.view()
        n, e = self.num_nodes, self.num_edges
        cap = e // threshold + 1
        dev = self.row_offset.device
        rows = torch.empty(cap, dtype=torch.int32, device=dev)
        count = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.call("stg_csr_hub_rows", self.row_offset.data_ptr(), n, threshold, rows.data_ptr(), cap, count.data_ptr(),
                  _lib.current_stream_ptr())
        enabled = True
        if getattr(self, "_hub_sync", False):      # static graph: one read-back, skip the launch when there are none
            cnt = min(int(count.item()), cap)
            enabled = cnt > 0
            if cnt > 1:     # longest rows first: blocks take the list round-robin, so the big rows start at once
                r = rows[:cnt].long()
                length = (self.row_offset[r + 1] - self.row_offset[r]).long()
                order = torch.sort(length * self.num_nodes + (self.num_nodes - 1 - r), descending=True).indices
                rows[:cnt] = rows[:cnt][order]
        v = _lib.StgCsrView()
        for name, _ in _lib.StgCsrView._fields_:
            setattr(v, name, getattr(base, name))
        v.hub_rows = rows.data_ptr() if enabled else None
        v.hub_count = count.data_ptr() if enabled else None
        v.hub_threshold = threshold if enabled else 0
        v.hub_capacity = cap if enabled else 0
        alt[threshold] = (v, rows, count)
    return alt[threshold][0]


CSR.view_with_hub_threshold = _alt_view


def build_csr_pair(src: torch.Tensor, dst: torch.Tensor, num_nodes: int, want_perm: bool = False):
    """Build (forward, backward, edge_perm, num_unique) on the GPU from device int32 edge arrays."""
    assert src.is_cuda and dst.is_cuda, "stgraph_b200 builds graphs on the GPU only (no CPU fallback)"
    dev = src.device
    e = int(src.shape[0])
    n = int(num_nodes)
    i32 = dict(dtype=torch.int32, device=dev)
    f_ro = torch.empty(n + 1, **i32)
    f_col = torch.empty(e, **i32)
    f_eid = torch.empty(e, **i32)
    f_nid = torch.empty(n, **i32)
    b_ro = torch.empty(n + 1, **i32)
    b_col = torch.empty(e, **i32)
    b_eid = torch.empty(e, **i32)
    b_nid = torch.empty(n, **i32)
    in_deg = torch.empty(n, **i32)
    out_deg = torch.empty(n, **i32)
    perm = torch.empty(e, **i32) if want_perm else None
    n_unique = torch.zeros(1, **i32)
    ws_bytes = _lib.load().stg_csr_build_workspace_bytes(e, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.call("stg_csr_build", src.data_ptr(), dst.data_ptr(), e, n,
              f_ro.data_ptr(), f_col.data_ptr(), f_eid.data_ptr(), f_nid.data_ptr(),
              b_ro.data_ptr(), b_col.data_ptr(), b_eid.data_ptr(), b_nid.data_ptr(),
              in_deg.data_ptr(), out_deg.data_ptr(), _lib.ptr(perm), n_unique.data_ptr(),
              ws.data_ptr(), ws_bytes, _lib.current_stream_ptr())
    fwd = CSR(f_ro, f_col, f_eid, f_nid, in_deg, out_deg, eid_base=0, eids_identity=True)
    bwd = CSR(b_ro, b_col, b_eid, b_nid, out_deg, in_deg, eid_base=0, eids_identity=False)
    return fwd, bwd, perm, n_unique


def get_array(ptr: int, size: int):
    """D2H copy of an int32 device array by raw address (``csr.cu:172-179``)."""
    out = (ctypes.c_int32 * size)()
    _lib.call("stg_get_array_i32", ptr, size, ctypes.addressof(out), _lib.current_stream_ptr())
    return list(out)
