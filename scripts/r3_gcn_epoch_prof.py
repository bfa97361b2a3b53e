"""Where does the 2-layer GCN epoch of bench.py go (config 5, one GPU)?  torch.profiler kernel table of one epoch."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stgraph_b200.graph import StaticGraph  # noqa: E402
from stgraph_b200.nn.pytorch import GCNConv  # noqa: E402
from stgraph_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda")
d = synthetic.products_shaped(seed=0, device=dev, scale=float(os.environ.get("STG_BENCH_SCALE", "1.0")), locality=0.9, window=8192)
n = d["num_nodes"]
graph = StaticGraph(torch.stack([d["src"], d["dst"]], 1), None, n)
graph.set_ndata("norm", graph.degree_norm())
x = torch.randn(n, 100, device=dev)
torch.manual_seed(7)
l1, l2 = GCNConv(100, 100, activation=torch.relu).to(dev), GCNConv(100, 47).to(dev)
params = list(l1.parameters()) + list(l2.parameters())
labels = torch.randint(0, 47, (n,), device=dev)
opt = torch.optim.Adam(params, lr=1e-2)


def epoch():
    opt.zero_grad()
    loss = torch.nn.functional.cross_entropy(l2(graph, l1(graph, x)), labels, reduction=os.environ.get("STG_LOSS_REDUCTION", "none")).sum() / n
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    epoch()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    epoch()
b.record()
torch.cuda.synchronize()
res = {"ms_per_epoch": a.elapsed_time(b) / 3}
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    epoch()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
hist = {}
for e in evs:
    c = hist.setdefault(e.name[:110], [0, 0.0])
    c[0] += 1
    c[1] += e.device_time / 1e3
res["kernels"] = len(evs)
res["kernel_ms"] = sum(c[1] for c in hist.values())
res["top"] = [{"name": k, "count": c[0], "ms": round(c[1], 3)} for k, c in sorted(hist.items(), key=lambda kv: -kv[1][1])[:25]]
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
