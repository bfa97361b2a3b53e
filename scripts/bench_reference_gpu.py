"""Reference's OWN GPU kernels vs ours on the config-5 graph (secondary baseline).

The kernels under ``oracle/_ref/gcn_f*_gpu.so`` are the CUDA text the reference's Seastar code generator
emits for ``GCNConv`` (captured by ``oracle/build_ref.py``), compiled as generic ``compute_100`` code the
way the reference's JIT would (``stgraph/compiler/code_gen/compiler.py:19-21``) and launched with the
geometry the reference computes (``execution_unit.py:92-106``).  Results are checked against our kernel
first, then both are timed with CUDA events.  Baseline tooling: not part of the product path.
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_emulate as RE  # noqa: E402
from stgraph_b200 import kernels  # noqa: E402
from stgraph_b200.graph import StaticGraph  # noqa: E402
from stgraph_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda")
out = {}


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(graph_name, d):
    n = d["num_nodes"]
    e = int(d["src"].shape[0])
    g = StaticGraph(torch.stack([d["src"], d["dst"]], 1), None, n)
    norm2 = g.degree_norm().contiguous()                 # [N,1] as the reference expects
    norm = norm2.reshape(-1)
    F_, B_ = g._forward_graph, g._backward_graph
    for feat in (100, 128, 16):
        case = f"gcn_f{feat}"
        so = os.path.join(RE.REF_DIR, case + "_gpu.so")
        if not os.path.exists(so):
            continue
        kernels_meta, _ = RE.load_case(case)
        lib = ctypes.CDLL(so)
        x = torch.randn(n, feat, device=dev)
        ours = torch.empty_like(x)
        res = {}
        for k in kernels_meta:
            csr = F_ if k["parallel_mode"] == "DstParallel" else B_
            view = g.fwd_view() if k["parallel_mode"] == "DstParallel" else g.bwd_view()
            ref_out = torch.zeros(n, feat, device=dev)   # executor.new_zeros
            # kernel args sorted by id: data tensor (Vhinb or grad), Vnormcen, Vnorminb, output
            tensors = {}
            for name in k["args"]:
                if name in k["rets"]:
                    tensors[name] = ref_out
                elif "norm" in name:
                    tensors[name] = norm2
                else:
                    tensors[name] = x
            arr = (ctypes.c_void_p * len(k["args"]))(*[ctypes.c_void_p(tensors[a].data_ptr()) for a in k["args"]])
            nblks, nthrs, group, npb = RE.reference_launch_params(feat, n)
            fn = getattr(lib, "launch_" + k["name"])
            fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 7 + [ctypes.c_void_p]
            stream = torch.cuda.current_stream().cuda_stream

            def launch_ref():
                rc = fn(arr, csr.row_offset.data_ptr(), csr.eids.data_ptr(), csr.column_indices.data_ptr(),
                        csr.node_ids.data_ptr(), n, feat, 1, group, npb, nblks, nthrs, stream)
                assert rc == 0, rc

            launch_ref()
            kernels.agg_scaled_sum(view, x, norm, None, norm, out=ours)
            torch.cuda.synchronize()
            mag = kernels.agg_scaled_sum(view, x.abs(), norm, None, norm)
            ok = bool(((ours - ref_out).abs() <= 1e-5 * mag + 1e-30).all())
            t_ref = timeit(launch_ref)
            t_ours = timeit(lambda: kernels.agg_scaled_sum(view, x, norm, None, norm, out=ours))
            res[k["direction"]] = {"reference_kernel_ms": t_ref, "our_kernel_ms": t_ours, "speedup": t_ref / t_ours,
                                   "results_agree_1e-5": ok, "reference_launch": [nblks, nthrs, group, npb]}
        out[f"{graph_name}_F{feat}"] = res
        print(graph_name, feat, json.dumps(res), flush=True)


def run_gat():
    """Config 3: the three kernels the reference emits for GATConv(.,16,8) (K0 scores + row sums, K1 weighted sum,
    K2 backward with its atomics) on the arxiv-shaped graph, against our stock-program path and the fused kernels."""
    import numpy as np
    from stgraph_b200.nn.pytorch import GATConv
    from stgraph_b200.ops_gat import gat_edge_softmax_aggregate

    case = "gat_h8d16"
    so = os.path.join(RE.REF_DIR, case + "_gpu.so")
    if not os.path.exists(so):
        return
    d = synthetic.arxiv_shaped(seed=0, device=dev)
    n, e = d["num_nodes"], int(d["src"].shape[0])
    g = StaticGraph(torch.stack([d["src"], d["dst"]], 1), None, n)
    F_, B_ = g._forward_graph, g._backward_graph
    kernels_meta, _ = RE.load_case(case)
    lib = ctypes.CDLL(so)
    tensors, res = {}, {}
    stream = torch.cuda.current_stream().cuda_stream
    for k in kernels_meta:
        csr = F_ if k["parallel_mode"] == "DstParallel" else B_
        for name, vt, shp in zip(k["args"], k["arg_types"], k["arg_shapes"]):
            if name not in tensors:
                lead = e if vt == "EDGE" else n
                tensors[name] = torch.zeros([lead] + shp, device=dev) if name in k["rets"] else torch.randn([lead] + shp, device=dev)
        arr = (ctypes.c_void_p * len(k["args"]))(*[ctypes.c_void_p(tensors[a].data_ptr()) for a in k["args"]])
        md = k["max_dims"]
        max_dims = [1, md[-1]] if len(md) == 1 else md
        feat = int(np.prod(md))
        nblks, nthrs, group, npb = RE.reference_launch_params(feat, n)
        fn = getattr(lib, "launch_" + k["name"])
        fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 7 + [ctypes.c_void_p]

        def launch_ref(fn=fn, arr=arr, csr=csr, max_dims=max_dims, group=group, npb=npb, nblks=nblks, nthrs=nthrs):
            rc = fn(arr, csr.row_offset.data_ptr(), csr.eids.data_ptr(), csr.column_indices.data_ptr(),
                    csr.node_ids.data_ptr(), n, max_dims[1], max_dims[0], group, npb, nblks, nthrs, stream)
            assert rc == 0, rc

        res[k["name"] + "_" + k["direction"] + "_ms"] = timeit(launch_ref)
    res["reference_fwd_ms"] = sum(v for k_, v in res.items() if "_forward_" in k_)
    res["reference_bwd_ms"] = sum(v for k_, v in res.items() if "_backward_" in k_)
    x = torch.randn(n, 128, device=dev)
    gout = torch.randn(n, 8, 16, device=dev)
    for mode in ("stock", "fused"):
        torch.manual_seed(0)
        layer = GATConv(128, 16, 8, softmax=mode).to(dev)

        def step():
            layer.zero_grad()
            layer(g, x).backward(gout)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        res[f"our_layer_{mode}_fwd_bwd_ms"] = (time.perf_counter() - t0) / 10 * 1e3
    out["config3_arxiv_gat_h8d16"] = res
    print("config3_arxiv_gat", json.dumps(res), flush=True)


if __name__ == "__main__":
    run_gat()
    torch.cuda.empty_cache()

    run("config5_locality0.9", synthetic.products_shaped(seed=0, device=dev))
    torch.cuda.empty_cache()
    run("config5_locality0.0", synthetic.products_shaped(seed=0, device=dev, locality=0.0))
    torch.cuda.empty_cache()
    c = synthetic.cora_shaped(seed=0, device=dev)
    run("config1_cora", c)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/reference_gpu.json", "w"), indent=1)
