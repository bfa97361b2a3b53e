"""Row-partitioned aggregation across GPUs: one process per GPU, NCCL over NVLink.

SURVEY.md section 8(e): destination vertices are split into P contiguous ranges with ~E/P
in-edges each (edge-balanced); rank p owns rows ``[v_p, v_{p+1})`` of the in-edge CSR and of
``h / out / norm``.  Forward: every rank needs the source rows its edges reference, so the row
blocks of ``h`` are exchanged (one broadcast per owner, i.e. an all-gather with uneven blocks),
then the local gather kernel runs on the rank's slice of the CSR.  Backward is the same shape on
the out-edge CSR (partition by source, exchange ``grad_out``): no reduce-scatter, no atomics.

The local slice needs no re-indexing: ``row_offset[v_p : v_{p+1}+1]`` still indexes the global
``column_indices`` array, so the C-ABI view is the same arrays with an offset row pointer.
The reference has no multi-GPU path at all.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..graph.static.csr import HUB_THRESHOLD


def edge_balanced_bounds(row_offset: torch.Tensor, world: int) -> list[int]:
    """Row boundaries ``v_0=0 <= ... <= v_P=N`` such that every range holds ~E/P edges."""
    n = int(row_offset.shape[0] - 1)
    e = int(row_offset[-1].item())
    targets = torch.arange(1, world, dtype=torch.int64, device=row_offset.device) * e // max(world, 1)
    cuts = torch.searchsorted(row_offset.to(torch.int64), targets, right=False).clamp_(0, n)
    b = [0] + [int(c) for c in cuts.cpu()] + [n]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


class _LocalSlice:
    """Rows [lo, hi) of one CSR direction as a StgCsrView (global column ids, local row ids)."""

    def __init__(self, csr, lo: int, hi: int):
        self.lo, self.hi = lo, hi
        self.csr = csr
        n_local = hi - lo
        ro = csr.row_offset[lo:hi + 1]
        dev = ro.device
        e_local = int((ro[-1] - ro[0]).item()) if n_local > 0 else 0
        cap = e_local // max(HUB_THRESHOLD, 1) + 1
        self.hub_rows = torch.empty(cap, dtype=torch.int32, device=dev)
        self.hub_count = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.call("stg_csr_hub_rows", ro.data_ptr(), n_local, HUB_THRESHOLD, self.hub_rows.data_ptr(), cap,
                  self.hub_count.data_ptr(), _lib.current_stream_ptr())
        has_hubs = int(self.hub_count.item()) > 0
        v = _lib.StgCsrView()
        v.row_offset = ro.data_ptr()
        v.column_indices = csr.column_indices.data_ptr()
        v.eids = csr.eids.data_ptr() if csr.eids is not None else None
        v.node_ids = None
        v.num_nodes = n_local
        v.num_edges = csr.num_edges
        v.eid_base = csr.eid_base
        v.eids_identity = 1 if csr.eids_identity else 0
        v.hub_rows = self.hub_rows.data_ptr() if has_hubs else None
        v.hub_count = self.hub_count.data_ptr() if has_hubs else None
        v.hub_threshold = HUB_THRESHOLD if has_hubs else 0
        v.hub_capacity = cap if has_hubs else 0
        self.view = v
        self.num_local_edges = e_local


class PartitionedGraph:
    """A StaticGraph replicated on every rank, with this rank's row slices of both directions."""

    def __init__(self, graph, rank: int, world: int):
        self.graph = graph
        self.rank, self.world = rank, world
        self.fwd_bounds = edge_balanced_bounds(graph._forward_graph.row_offset, world)
        self.bwd_bounds = edge_balanced_bounds(graph._backward_graph.row_offset, world)
        self.fwd = _LocalSlice(graph._forward_graph, self.fwd_bounds[rank], self.fwd_bounds[rank + 1])
        self.bwd = _LocalSlice(graph._backward_graph, self.bwd_bounds[rank], self.bwd_bounds[rank + 1])

    def halo_plans(self, group=None):
        """Halo-only exchange plans: (forward, backward).

        Feature rows are owned by the forward (destination) partition -- layer outputs are produced
        there -- and gradient rows likewise; the backward aggregation computes the source rows of
        ``bwd_bounds`` and fetches the gradient rows it needs from their forward owners.
        """
        from .halo import HaloPlan

        g = self.graph
        fwd = HaloPlan(g._forward_graph, self.fwd_bounds, self.fwd_bounds, self.rank, self.world, group)
        bwd = HaloPlan(g._backward_graph, self.bwd_bounds, self.fwd_bounds, self.rank, self.world, group)
        return fwd, bwd

    def local_rows(self, direction: str):
        b = self.fwd_bounds if direction == "fwd" else self.bwd_bounds
        return b[self.rank], b[self.rank + 1]


def exchange_rows(full: torch.Tensor, bounds: list[int], group=None):
    """All-gather with uneven blocks: afterwards ``full[bounds[p]:bounds[p+1]]`` holds rank p's rows on every rank.

    ``full`` is the ``[N, F]`` buffer whose own block ``[bounds[rank], bounds[rank+1])`` is already valid.
    One broadcast per owner (NCCL coalesces them on its stream; NVSwitch gives every pair full bandwidth).
    """
    import torch.distributed as dist

    works = []
    for p in range(len(bounds) - 1):
        blk = full[bounds[p]:bounds[p + 1]]
        if blk.numel() > 0:
            works.append(dist.broadcast(blk, src=p, group=group, async_op=True))
    for w in works:
        w.wait()
    return full
