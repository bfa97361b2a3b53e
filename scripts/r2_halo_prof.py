"""Profiling driver: the halo-source pass (and the own-source pass) of rank 0's share of the 8-way partition, on one GPU
(no communication: the halo buffer is filled with the right rows locally).  Usage: python scripts/r2_halo_prof.py [P]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stgraph_b200 import _lib, kernels  # noqa: E402
from stgraph_b200.dist.partition import cost_balanced_bounds  # noqa: E402
from stgraph_b200.graph import StaticGraph  # noqa: E402
from stgraph_b200.graph.static.csr import HUB_THRESHOLD  # noqa: E402
from stgraph_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
F = 100
cache = "/tmp/config5_edges.pt"
if os.path.exists(cache):
    src, dst = [t.to(dev) for t in torch.load(cache)]
    n = 2449029
else:
    d = synthetic.products_shaped(seed=0, device=dev)
    src, dst, n = d["src"], d["dst"], d["num_nodes"]
    torch.save((src.cpu(), dst.cpu()), cache)
g = StaticGraph(torch.stack([src, dst], 1), None, n)
norm = g.degree_norm().reshape(-1).contiguous()
x = torch.randn(n, F, device=dev)
csr = g._forward_graph
bounds = cost_balanced_bounds(csr.row_offset, g._backward_graph.row_offset, P)
keep = []


def mkview(ro, cols, n_rows):
    q = torch.zeros(2, dtype=torch.int32, device=dev)
    keep.append((ro, cols, q))
    v = _lib.StgCsrView()
    v.row_offset, v.column_indices, v.eids, v.node_ids = ro.data_ptr(), cols.data_ptr(), None, None
    v.num_nodes, v.num_edges, v.eid_base, v.eids_identity = n_rows, int(cols.shape[0]), 0, 1
    v.hub_rows = v.hub_count = None
    v.hub_threshold = v.hub_capacity = 0
    v.work_queue = q.data_ptr()
    return v


def mkview_hub(ro, cols, n_rows, threshold):
    v = mkview(ro, cols, n_rows)
    if threshold > 0:
        cap = int(cols.shape[0]) // threshold + 1
        hr = torch.empty(cap, dtype=torch.int32, device=dev)
        hc = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.call("stg_csr_hub_rows", ro.data_ptr(), n_rows, threshold, hr.data_ptr(), cap, hc.data_ptr(), _lib.current_stream_ptr())
        nh = int(hc.item())
        keep.append((hr, hc))
        if nh > 0:
            v.hub_rows, v.hub_count, v.hub_threshold, v.hub_capacity = hr.data_ptr(), hc.data_ptr(), threshold, cap
        return v, nh
    return v, 0


def timeit(fn, reps=20):
    for _ in range(4):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rank = 0
lo, hi = bounds[rank], bounds[rank + 1]
ro = csr.row_offset[lo:hi + 1].long()
e0, e1 = int(ro[0]), int(ro[-1])
cols = csr.column_indices[e0:e1].long()
local = (cols >= lo) & (cols < hi)
halo_ids = torch.unique(cols[~local])
nr = hi - lo
rows = torch.repeat_interleave(torch.arange(nr, device=dev), ro[1:] - ro[:-1])
cnt = torch.bincount(rows[~local], minlength=nr)
sel = torch.nonzero(cnt > 0).reshape(-1)
cro = torch.zeros(sel.numel() + 1, dtype=torch.int32, device=dev)
cro[1:] = torch.cumsum(cnt[sel], 0).int()
hcols = torch.searchsorted(halo_ids, cols[~local]).int().contiguous()
ocnt = torch.bincount(rows[local], minlength=nr)
oro = torch.zeros(nr + 1, dtype=torch.int32, device=dev)
oro[1:] = torch.cumsum(ocnt, 0).int()
ocols = (cols[local] - lo).int().contiguous()
halo = x[halo_ids].contiguous()
x_own = x[lo:hi].contiguous()
ns_halo = norm[halo_ids].contiguous()
ns_own = norm[lo:hi].contiguous()
rs = ns_own
out = torch.randn(nr, F, device=dev)
out_rows = sel.int().contiguous()
print("P", P, "rows", nr, "rows with halo", int(sel.numel()), "halo edges", int(hcols.numel()), "halo rows", int(halo_ids.numel()),
      "own edges", int(ocols.numel()), "max halo deg", int(cnt.max()), "chunk", os.environ.get("STG_AGG_CHUNK", "4"))
for t in (0, 32, 64, 128, 256, 1024):
    v_halo, nh = mkview_hub(cro, hcols, int(sel.numel()), t)
    meta = kernels.pack_edge_meta(v_halo, ns_halo, None, device=dev)
    ms = timeit(lambda: kernels.agg_packed_sum_rows(v_halo, meta, out_rows, halo, rs, out, accumulate=True))
    print(f"halo pass: hub threshold {t:5d} ({nh} hub rows): {ms:.4f} ms", flush=True)
for t in (64, 128, 256, 512, 1024):
    v_own, nh = mkview_hub(oro, ocols, nr, t)
    meta = kernels.pack_edge_meta(v_own, ns_own, None, device=dev)
    ms = timeit(lambda: kernels.agg_packed_sum_rows(v_own, meta, None, x_own, rs, out, accumulate=False))
    print(f"own pass : hub threshold {t:5d} ({nh} hub rows): {ms:.4f} ms", flush=True)
