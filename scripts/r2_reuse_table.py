"""How much on-chip reuse does the config-5 graph offer?  

For tiles of T consecutive destination rows of the in-edge CSR: edges per tile, unique source rows per tile
(the bytes a perfect per-tile cache would still have to fetch), and the share of the tile's edges whose source
falls inside a window of W source rows centred on the tile (what a shared-memory / DSMEM staging window of W
rows would serve).  W = 512 ~ one SM's shared memory at F=100 (228 KB / 400 B = 570 rows), 4096 ~ an 8-CTA
cluster, 8192 ~ a 16-CTA cluster (the non-portable maximum), 65536 = 26 MB (L2 scale, for context).
Also: the share of edges whose source is among the K most popular sources (a replicated hot set).

    python scripts/r2_reuse_table.py [out.json]      # GPU if available (same generator stream as bench.py), else CPU
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stgraph_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
scale = float(os.environ.get("STG_REUSE_SCALE", "1.0"))
d = synthetic.products_shaped(seed=0, device=dev, scale=scale)
n = d["num_nodes"]
src, dst = d["src"].long(), d["dst"].long()
e = int(src.shape[0])
res = {"device": str(dev), "num_nodes": n, "num_edges": e, "tiles": [], "hot_sources": []}
for T in (128, 512, 2048, 4096, 8192, 32768):
    tile = dst // T
    key = torch.unique(tile * n + src)
    uniq = int(key.shape[0])
    row = {"tile_rows": T, "unique_sources_per_edge": uniq / e, "unique_rows_bytes_GB_at_F100": uniq * 400 / 1e9}
    centre = tile * T + T // 2
    for W in (512, 4096, 8192, 65536):
        inside = ((src - centre).abs() <= W // 2).float().mean().item()
        row[f"edges_inside_window_{W}"] = inside
    res["tiles"].append(row)
    print(row, flush=True)
deg = torch.bincount(src, minlength=n)
top = torch.sort(deg, descending=True).values.cumsum(0)
for K in (512, 4096, 32768, 262144):
    res["hot_sources"].append({"top_k": K, "edge_share": float(top[K - 1]) / e})
print(res["hot_sources"])
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
