"""Where does a config-4 timestep go?  One BPTT window (20 snapshots) of the dynamic TGCN loop under torch.profiler:
top CUDA kernels by total time + host-side segment times.  python scripts/r2_config4_prof.py [out.json]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = sys.argv[:1] + sys.argv[1:]
import scripts.bench_configs as BC  # noqa: E402
from stgraph_b200.graph import GPMAGraph  # noqa: E402
from stgraph_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda")
n, base, slide, T = 1_000_000, 10_000_000, 100_000, 24
src, dst = synthetic.temporal_stream(n, base + slide * (T - 1), alpha=1.8, seed=0, device=dev, distinct=True, max_frac=2e-4)
snaps = [torch.stack([s, d_], 1) for s, d_ in synthetic.sliding_window_snapshots(src, dst, base, slide, T)]
G = GPMAGraph(snaps, n)
del snaps, src, dst
gen = torch.Generator(device=dev).manual_seed(5)
pairs, targets = [], []
for t in range(T - 1):
    add = G.graph_updates[str(t + 1)]["add"]
    pos = torch.stack([add & 0xFFFFFFFF, add >> 32])
    neg = torch.randint(0, n, (2, pos.shape[1]), device=dev, generator=gen)
    pairs.append(torch.cat([pos, neg], 1).contiguous())
    targets.append(torch.cat([torch.ones(pos.shape[1], device=dev), torch.zeros(pos.shape[1], device=dev)]))
torch.manual_seed(0)
model = BC.DynTGCN(32, 64).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-2)
crit = torch.nn.BCEWithLogitsLoss()
seg = {"get_graph": 0.0, "norm": 0.0, "model_fwd": 0.0, "decode_loss": 0.0, "backward": 0.0, "opt": 0.0}


def window(profile_segments):
    G.reset_graph()
    opt.zero_grad()
    cost, h = 0, None
    y_hat = torch.randn(n, 32, device=dev)

    def tick(name, t0):
        if profile_segments:
            torch.cuda.synchronize()
            seg[name] += time.perf_counter() - t0
        return time.perf_counter()

    for t in range(20):
        t0 = time.perf_counter()
        G.get_graph(t)
        t0 = tick("get_graph", t0)
        G.set_ndata("norm", G.degree_norm())
        t0 = tick("norm", t0)
        y_hat, h = model(G, y_hat, None, h)
        t0 = tick("model_fwd", t0)
        cost = cost + crit(model.decode(y_hat, pairs[t]), targets[t])
        t0 = tick("decode_loss", t0)
    t0 = time.perf_counter()
    (cost / 21).backward()
    t0 = tick("backward", t0)
    opt.step()
    tick("opt", t0)


window(False)
torch.cuda.synchronize()
t0 = time.perf_counter()
window(False)
torch.cuda.synchronize()
total = time.perf_counter() - t0
window(True)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    window(False)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in evs:
    k = e.name[:90]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.device_time / 1e3
top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]
res = {"window_ms_20_steps": total * 1e3, "host_segments_ms_with_syncs": {k: v * 1e3 for k, v in seg.items()},
       "kernels": len(evs), "kernel_time_ms": sum(e.device_time for e in evs) / 1e3,
       "top_kernels": [{"name": k, "count": c, "ms": ms} for k, (c, ms) in top]}
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
