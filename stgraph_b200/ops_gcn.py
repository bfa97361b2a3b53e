"""GCN aggregation as a direct torch autograd op (no tracing, no executor): the temporal-loop fast path.

``out = norm * sum_{u in in(v)} norm[u] * w[eid] * h[u]`` forward on the in-edge CSR, the adjoint
on the out-edge CSR -- the same two launches the compiler produces for ``GCNConv``'s vertex program
(``stgraph/nn/pytorch/static/gcn_conv.py:162-182``), minus the per-call Python of the executor
(~25 tiny kernels + 3 executor round trips per TGCN step dominate configs 1-2).
Sync-free and allocation-free on the C side, so a whole BPTT window can be captured in a CUDA graph.
For dynamic graphs the snapshot's views are captured at forward time (they are immutable tensors), so
no rewind is needed in backward.
"""
from __future__ import annotations

import torch

from . import kernels


class _GcnAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fwd_view, bwd_view, keepalive, h, norm, edge_weight):
        h = h.contiguous()
        nflat = norm.reshape(-1)
        wflat = edge_weight.reshape(-1) if edge_weight is not None else None
        if getattr(keepalive[0], "pack_enabled", False):   # static graph: packed {col, scale} array, built on first use
            out = kernels.agg_scaled_sum_graph(keepalive[0], h, nflat, wflat, nflat)
        else:
            out = kernels.agg_scaled_sum(fwd_view, h, nflat, wflat, nflat)
        ctx.bwd_view, ctx.keepalive = bwd_view, keepalive
        ctx.save_for_backward(nflat, wflat if wflat is not None else nflat)
        ctx.weighted = wflat is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        nflat, wflat = ctx.saved_tensors
        wb = wflat if ctx.weighted else None
        if getattr(ctx.keepalive[1], "pack_enabled", False):
            gh = kernels.agg_scaled_sum_graph(ctx.keepalive[1], gout.contiguous(), nflat, wb, nflat)
        else:
            gh = kernels.agg_scaled_sum(ctx.bwd_view, gout.contiguous(), nflat, wb, nflat)
        return None, None, None, gh, None, None


def gcn_aggregate(graph, h, norm, edge_weight=None):
    """Differentiable w.r.t. ``h`` only (like the reference: no gradient for ``norm`` / ``edge_weight``)."""
    if not h.is_cuda:
        raise RuntimeError("h must live on a CUDA device (stgraph_b200 has no CPU path)")
    bwd = graph.bwd_view()            # backward first: a dynamic graph builds both views in one go
    fwd = graph.fwd_view()
    keep = (graph._forward_graph, graph._backward_graph)
    return _GcnAggregate.apply(fwd, bwd, keep, h, norm, edge_weight)


#: rows from which the weight gradient of :func:`dense_transform` goes through ``kernels.gemm_tn`` instead of cuBLAS
TALL_ROWS = 32768


class _DenseTransform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, weight):
        ctx.save_for_backward(h, weight)
        return torch.mm(h, weight)

    @staticmethod
    def backward(ctx, g):
        h, weight = ctx.saved_tensors
        d_h = torch.mm(g, weight.t()) if ctx.needs_input_grad[0] else None
        d_w = None
        if ctx.needs_input_grad[1]:
            hh = h if h.stride(-1) == 1 else h.contiguous()
            gg = g if g.stride(-1) == 1 else g.contiguous()
            d_w = kernels.gemm_tn(hh, gg)
        return d_h, d_w


def dense_transform(h, weight):
    """``h @ weight`` (GCNConv's ``X . W``, ``gcn_conv.py:160``): a library GEMM forward and for the input gradient; on
    graphs of ``TALL_ROWS`` vertices and more the weight gradient ``h^T @ g`` -- a ``[K, N] x [N, Nc]`` product that
    cuBLAS runs at a fraction of the machine -- goes through ``stg_gemm_tn_f32`` (exact fp32, deterministic)."""
    if (h.is_cuda and h.dtype == torch.float32 and weight.dtype == torch.float32 and h.dim() == 2 and h.shape[0] >= TALL_ROWS
            and not torch.is_autocast_enabled("cuda")):          # under autocast the GEMM runs in the autocast dtype: torch's own path
        return _DenseTransform.apply(h, weight)
    return torch.mm(h, weight)
