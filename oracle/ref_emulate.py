"""ctypes runners for the reference-derived checkers built by ``oracle/build_ref.py``.

TEST INFRASTRUCTURE ONLY.  ``RefCsr`` drives the reference's own ``CSR::CSR``
(``/root/reference/stgraph/graph/static/csr.cu:68-157``); ``run_reference_kernel`` executes one
kernel emitted by the reference's code generator on the CPU through the SIMT shim, with the launch
geometry the reference itself computes (``execution_unit.py:92-106``).
"""
from __future__ import annotations

import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ref_csr.so")) and os.path.exists(os.path.join(REF_DIR, "gcn_f16.so"))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class RefCsr:
    """Reference ``CSR(edge_list, edge_weight, num_nodes, is_edge_reverse)`` -- host vectors only."""

    def __init__(self, a, b, eid, weights, num_nodes, is_edge_reverse):
        lib = ctypes.CDLL(os.path.join(REF_DIR, "ref_csr.so"))
        lib.ref_csr_new.restype = ctypes.c_void_p
        lib.ref_csr_new.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.ref_csr_get.argtypes = [ctypes.c_void_p] * 8
        lib.ref_csr_free.argtypes = [ctypes.c_void_p]
        a, b, eid = _i32(a), _i32(b), _i32(eid)
        w = np.ascontiguousarray(np.asarray(weights, dtype=np.float32))
        e = a.shape[0]
        h = lib.ref_csr_new(a.ctypes.data, b.ctypes.data, eid.ctypes.data, e, w.ctypes.data, num_nodes, int(is_edge_reverse))
        self.row_offset = np.zeros(num_nodes + 1, np.int32)
        self.column_indices = np.zeros(e, np.int32)
        self.eids = np.zeros(e, np.int32)
        self.node_ids = np.zeros(num_nodes, np.int32)
        self.in_degrees = np.zeros(num_nodes, np.int32)
        self.out_degrees = np.zeros(num_nodes, np.int32)
        self.weighted_out_degrees = np.zeros(num_nodes, np.float32)
        lib.ref_csr_get(h, *[x.ctypes.data for x in (self.row_offset, self.column_indices, self.eids, self.node_ids,
                                                      self.in_degrees, self.out_degrees, self.weighted_out_degrees)])
        lib.ref_csr_free(h)


def reference_static_graph(src, dst, weights_by_eid, num_nodes):
    """What ``StaticGraph.__init__`` builds (``static_graph.py:40-78``): forward + backward reference CSRs."""
    src, dst = np.asarray(src), np.asarray(dst)
    order = np.lexsort((src, dst))                      # edge_list.sort(key=lambda x: (x[1], x[0]))
    s, d = src[order], dst[order]
    eid = np.arange(s.shape[0])
    fwd = RefCsr(s, d, eid, weights_by_eid, num_nodes, is_edge_reverse=True)
    border = np.lexsort((eid, d, s))                    # sorted() of (src, dst, eid) triples
    bwd = RefCsr(s[border], d[border], eid[border], weights_by_eid, num_nodes, is_edge_reverse=False)
    return fwd, bwd


def load_case(case):
    with open(os.path.join(REF_DIR, case + ".json")) as f:
        meta = json.load(f)
    lib = ctypes.CDLL(os.path.join(REF_DIR, case + ".so"))
    return meta["kernels"], lib


def reference_launch_params(feat_size, num_nodes):
    """Restatement of ``ExecutionUnit.calculate_kernel_params_fa`` (``execution_unit.py:92-116``)."""
    if feat_size >= 64:
        nthrs = min(256, feat_size)
        return num_nodes, nthrs, nthrs, 1
    nthrs = 64
    g = nthrs
    while g > feat_size:
        g //= 2
    g = max(1, g)
    npb = max(2, nthrs // g)
    return (num_nodes + npb - 1) // npb, nthrs, g, npb


def run_reference_kernel(lib, kernel, tensors, csr, num_nodes):
    """Run one emitted kernel.  ``tensors``: {var id: float32 ndarray} (rets are written in place)."""
    arr = (ctypes.c_void_p * len(kernel["args"]))()
    for i, name in enumerate(kernel["args"]):
        t = tensors[name]
        assert t.dtype == np.float32 and t.flags["C_CONTIGUOUS"]
        arr[i] = t.ctypes.data
    md = kernel["max_dims"]
    max_dims = [1, md[-1]] if len(md) == 1 else md
    feat = int(np.prod(md))
    nblks, nthrs, group, npb = reference_launch_params(feat, num_nodes)
    fn = getattr(lib, "run_" + kernel["name"])
    fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 7
    ro, ei, ci, ni = _i32(csr.row_offset), _i32(csr.eids), _i32(csr.column_indices), _i32(csr.node_ids)
    fn(arr, ro.ctypes.data, ei.ctypes.data, ci.ctypes.data, ni.ctypes.data, num_nodes, max_dims[1], max_dims[0],
       group, npb, nblks, nthrs)
    return (nblks, nthrs, group, npb)


class RefPcsr:
    """The reference's host-side ``PCSR`` (``/root/reference/stgraph/graph/dynamic/pcsr/pcsr.cu:325-891``) driven the way
    ``PCSRGraph`` drives it (``pcsr_graph.py:45-166``): ``edge_update_list(add, is_reverse_edge=True)``,
    ``edge_update_list(delete, is_delete=True, is_reverse_edge=True)``, ``label_edges()``, then ``build_csr()``
    (forward) or ``build_reverse_csr()`` (backward)."""

    def __init__(self, num_nodes: int, max_edges: int):
        lib = ctypes.CDLL(os.path.join(REF_DIR, "ref_pcsr.so"))
        lib.ref_pcsr_new.restype = ctypes.c_void_p
        lib.ref_pcsr_new.argtypes = [ctypes.c_int, ctypes.c_int]
        lib.ref_pcsr_update.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.ref_pcsr_label.argtypes = [ctypes.c_void_p]
        lib.ref_pcsr_build.restype = ctypes.c_int
        lib.ref_pcsr_build.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6
        lib.ref_pcsr_free.argtypes = [ctypes.c_void_p]
        self.lib, self.n, self.max_edges = lib, int(num_nodes), max(int(max_edges), 1)
        self.h = lib.ref_pcsr_new(self.n, self.max_edges)

    def update(self, src, dst, delete: bool):
        s = np.ascontiguousarray(np.asarray(src, dtype=np.uint32))
        d = np.ascontiguousarray(np.asarray(dst, dtype=np.uint32))
        self.lib.ref_pcsr_update(self.h, s.ctypes.data, d.ctypes.data, int(s.shape[0]), int(delete), 1)

    def step(self, add, delete):
        """One ``_update_graph_forward`` (or, with the two lists swapped, ``_update_graph_backward``)."""
        self.update(add[0], add[1], False)
        self.update(delete[0], delete[1], True)
        self.lib.ref_pcsr_label(self.h)

    def build(self, reverse: bool):
        """(row_offset, column_indices, eids, node_ids, in_degrees, out_degrees) as the reference fills them."""
        n, m = self.n, self.max_edges
        ro, col, eid = np.zeros(n + 1, np.uint32), np.zeros(m, np.uint32), np.zeros(m, np.uint32)
        nid, ind, outd = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        e = self.lib.ref_pcsr_build(self.h, int(reverse), *[x.ctypes.data for x in (ro, col, eid, nid, ind, outd)])
        return ro, col[:e].copy(), eid[:e].copy(), nid, ind, outd

    def close(self):
        if self.h:
            self.lib.ref_pcsr_free(self.h)
            self.h = None
