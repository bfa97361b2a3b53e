"""Round-2 A/B of the global-row-queue aggregation kernel against the static block ranges (one GPU).

Usage: STG_AGG_CHUNK=<c> python scripts/r2_agg_sweep.py [out.json]   (the chunk size is read once per process)
Times config 5 (F=100 and 47) forward/backward with the packed metadata, checks the queue result bit for bit
against the static schedule, and times one rank's share of the 8-way partition (own-source edges).
"""
import copy
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stgraph_b200 import _lib, kernels  # noqa: E402
from stgraph_b200.graph import StaticGraph  # noqa: E402
from stgraph_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda")
cache = "/tmp/config5_edges.pt"
if os.path.exists(cache):
    src, dst = torch.load(cache)
    src, dst = src.to(dev), dst.to(dev)
    n = 2449029
else:
    d = synthetic.products_shaped(seed=0, device=dev)
    src, dst, n = d["src"], d["dst"], d["num_nodes"]
    torch.save((src.cpu(), dst.cpu()), cache)
g = StaticGraph(torch.stack([src, dst], 1), None, n)
norm = g.degree_norm().reshape(-1).contiguous()
res = {"chunk": os.environ.get("STG_AGG_CHUNK", "8")}


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def no_queue(view):
    v = _lib.StgCsrView()
    for name, _ in _lib.StgCsrView._fields_:
        setattr(v, name, getattr(view, name))
    v.work_queue = None
    return v


for F in [int(v) for v in os.environ.get("STG_SWEEP_F", "100,47,128,64").split(",")]:
    x = torch.randn(n, F, device=dev)
    out = torch.empty_like(x)
    out2 = torch.empty_like(x)
    for name, csr in (("fwd", g._forward_graph), ("bwd", g._backward_graph)):
        meta = csr.packed_meta(norm, None)
        vq = csr.view()
        vs = no_queue(vq)
        t_q = timeit(lambda: kernels.agg_packed_sum(vq, meta, x, norm, out=out))
        t_s = timeit(lambda: kernels.agg_packed_sum(vs, meta, x, norm, out=out2))
        same = bool(torch.equal(out, out2))
        t_qp = timeit(lambda: kernels.agg_scaled_sum(vq, x, norm, None, norm, out=out))
        t_sp = timeit(lambda: kernels.agg_scaled_sum(vs, x, norm, None, norm, out=out2))
        same_p = bool(torch.equal(out, out2))
        res[f"F{F}_{name}"] = {"queue_packed_ms": t_q, "static_packed_ms": t_s, "queue_plain_ms": t_qp,
                               "static_plain_ms": t_sp, "bit_equal": same and same_p}
        print(F, name, res[f"F{F}_{name}"], flush=True)
    del x, out, out2

print(json.dumps(res))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
