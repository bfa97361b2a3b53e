"""torchrun worker of tests/test_gpu_multi.py: the row-partitioned aggregation and GCNConv on >= 2 GPUs.

Every rank holds the whole (small) graph, so it can compute the single-GPU result of its own rows with the
single-GPU kernels AND with the CPU oracle, and compare the distributed path (halo exchange over peer memory +
the two aggregation passes) against both.  Exit code 0 = all checks passed on this rank.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import aggregate as A
    from oracle import structure as S
    from stgraph_b200 import kernels
    from stgraph_b200.dist import PartitionedGraph, all_reduce_gradients
    from stgraph_b200.graph import StaticGraph
    from stgraph_b200.nn.pytorch import GCNConv
    from stgraph_b200.utils import synthetic

    mode = os.environ.get("STG_HALO_MODE", "ce")
    n, e = 30000, 600000
    src, dst = synthetic.power_law_graph(n, e, alpha=2.2, locality=0.8, window=512, max_degree=3000, seed=11, device=dev)
    dist.broadcast(src, 0)
    dist.broadcast(dst, 0)
    g = StaticGraph(torch.stack([src, dst], 1), None, n)
    norm = g.degree_norm()
    g.set_ndata("norm", norm)
    pg = PartitionedGraph(g, rank, world)
    lo, hi = pg.own_lo, pg.own_hi
    nl = pg.local(norm).contiguous()
    pg.set_ndata("norm", nl)
    fo = S.forward_csr(src.cpu().numpy(), dst.cpu().numpy(), n)
    bo = S.backward_csr(src.cpu().numpy(), dst.cpu().numpy(), n)
    nrm = norm.reshape(-1).cpu()
    for feat in (100, 47, 16):
        gen = torch.Generator(device=dev).manual_seed(5 + feat)
        x = torch.randn(n, feat, device=dev, generator=gen)
        for direction, view, csr_o in (("fwd", g.fwd_view(), fo), ("bwd", g.bwd_view(), bo)):
            single = kernels.agg_scaled_sum(view, x, norm.reshape(-1), None, norm.reshape(-1))
            for rep in range(3):          # several steps: both halo buffers and the growing flags are exercised
                xs = x * (1.0 + rep)
                got = pg.aggregate(direction, xs[lo:hi].contiguous(), nl, nl)
                pg.exchange(direction, feat, nl).check()
                ref = A.scaled_sum(csr_o.row_offset, csr_o.column_indices, csr_o.eids, xs.cpu(), nrm, None, nrm)[lo:hi]
                mag = A.scaled_sum(csr_o.row_offset, csr_o.column_indices, csr_o.eids, xs.abs().cpu(), nrm, None, nrm)[lo:hi]
                err = (got.cpu().double() - ref.double()).abs()
                assert bool((err <= 1e-5 * mag.double() + 1e-30).all()), (mode, feat, direction, rep, float(err.max()))
                if rep == 0:
                    d1 = (got - single[lo:hi]).abs()
                    assert bool((d1.cpu().double() <= 2e-6 * mag.double() + 1e-30).all()), ("vs single GPU", feat, direction)
    # ---- edge weights: one value per global edge id, replicated; aggregation and layer against the single-GPU path
    w = torch.rand(e, 1, device=dev, generator=torch.Generator(device=dev).manual_seed(21)) + 0.25
    dist.broadcast(w, 0)
    x = torch.randn(n, 48, device=dev, generator=torch.Generator(device=dev).manual_seed(22))
    for direction, view in (("fwd", g.fwd_view()), ("bwd", g.bwd_view())):
        single = kernels.agg_scaled_sum(view, x, norm.reshape(-1), w.reshape(-1), norm.reshape(-1))
        mag = kernels.agg_scaled_sum(view, x.abs(), norm.reshape(-1), w.reshape(-1), norm.reshape(-1))
        got = pg.aggregate(direction, x[lo:hi].contiguous(), nl, nl, edge_scale=w)
        assert bool(((got - single[lo:hi]).abs() <= 2e-6 * mag[lo:hi] + 1e-30).all()), ("edge weights", direction)
    torch.manual_seed(4)
    lw = GCNConv(48, 12).to(dev)
    for p in lw.parameters():
        dist.broadcast(p.data, 0)
    gw = torch.randn(n, 12, device=dev, generator=torch.Generator(device=dev).manual_seed(23))
    xf = x.clone().requires_grad_(True)
    full = lw(g, xf, edge_weight=w)
    full.backward(gw)
    ref_w, ref_x = lw.weight.grad.clone(), xf.grad.clone()
    lw.zero_grad()
    xl = x[lo:hi].clone().requires_grad_(True)
    part = lw(pg, xl, edge_weight=w)
    part.backward(gw[lo:hi])
    all_reduce_gradients(lw)
    assert float((part - full[lo:hi]).detach().abs().max()) <= 2e-5 * float(full.detach().abs().max()), "weighted partitioned GCN forward differs"
    assert float((xl.grad - ref_x[lo:hi]).abs().max()) <= 2e-5 * float(ref_x.abs().max()) + 1e-7, "weighted dX differs"
    assert float((lw.weight.grad - ref_w).abs().max()) <= 5e-4 * float(ref_w.abs().max()) + 1e-6, "weighted dW differs"
    # ---- the layer API: 2-layer GCN on the partitioned graph == the same model on the whole graph
    torch.manual_seed(3)
    l1, l2 = GCNConv(64, 32, activation=torch.relu).to(dev), GCNConv(32, 7).to(dev)
    for p in list(l1.parameters()) + list(l2.parameters()):
        dist.broadcast(p.data, 0)
    x = torch.randn(n, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(9))
    gout = torch.randn(n, 7, device=dev, generator=torch.Generator(device=dev).manual_seed(10))
    full = l2(g, l1(g, x))
    full.backward(gout)
    ref_grads = [p.grad.clone() for p in list(l1.parameters()) + list(l2.parameters())]
    for p in list(l1.parameters()) + list(l2.parameters()):
        p.grad = None
    xl = x[lo:hi].clone().requires_grad_(True)
    part = l2(pg, l1(pg, xl))
    part.backward(gout[lo:hi])
    all_reduce_gradients(list(l1.parameters()) + list(l2.parameters()))
    scale = float(full.abs().max())
    assert float((part - full[lo:hi]).abs().max()) <= 2e-5 * scale, "partitioned GCN forward differs"
    for got, ref in zip([p.grad for p in list(l1.parameters()) + list(l2.parameters())], ref_grads):
        assert float((got - ref).abs().max()) <= 5e-4 * float(ref.abs().max()) + 1e-6, "weight gradients differ after all-reduce"
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"dist_worker ok: world={world} mode={mode}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
