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


def cost_balanced_bounds(fwd_row_offset: torch.Tensor, bwd_row_offset: torch.Tensor, world: int, row_cost: int = 4) -> list[int]:
    """ONE set of vertex boundaries for both directions: every range holds ~1/P of ``in_deg + out_deg + 2*row_cost``.

    A vertex is a row of the forward CSR (cost ~ in-degree) and of the backward CSR (cost ~ out-degree); a row also
    has a fixed cost (offsets, scale, output row: ``row_cost`` edge-equivalents, measured ~4 at F=100), so ranges
    of many short rows are not under-estimated.  One partition means one owner per feature row AND per gradient
    row, which is what a layer stack needs (the output rows of a layer are the source rows of the next).
    """
    n = int(fwd_row_offset.shape[0] - 1)
    cost = (fwd_row_offset[1:] - fwd_row_offset[:-1]).to(torch.int64) + (bwd_row_offset[1:] - bwd_row_offset[:-1]).to(torch.int64)
    cost = cost + 2 * int(row_cost)
    pre = torch.zeros(n + 1, dtype=torch.int64, device=cost.device)
    pre[1:] = torch.cumsum(cost, 0)
    total = int(pre[-1].item())
    targets = torch.arange(1, world, dtype=torch.int64, device=cost.device) * total // max(world, 1)
    cuts = torch.searchsorted(pre, targets, right=False).clamp_(0, n)
    b = [0] + [int(c) for c in cuts.cpu()] + [n]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


class PartitionedGraph:
    """A StaticGraph row-partitioned over ``world`` GPUs (structure replicated, rows of every node tensor sharded).

    Rank r owns the vertices ``[bounds[r], bounds[r+1])``: their feature rows, their output rows (forward, in-edge
    CSR) and their gradient rows (backward, out-edge CSR).  ``GCNConv.forward(pg, h_local)`` takes and returns the
    local rows only; :meth:`aggregate` is the distributed form of ``stg_agg_scaled_sum_f32``.
    """

    def __init__(self, graph, rank: int, world: int, group=None, bounds=None, row_cost: int = 4):
        self.graph = graph
        self.rank, self.world, self.group = rank, world, group
        f, b = graph._forward_graph, graph._backward_graph
        self.bounds = list(bounds) if bounds is not None else cost_balanced_bounds(f.row_offset, b.row_offset, world, row_cost)
        self.fwd_bounds = self.bwd_bounds = self.bounds
        self.own_lo, self.own_hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_own = self.own_hi - self.own_lo
        self.fwd = _LocalSlice(f, self.own_lo, self.own_hi)
        self.bwd = _LocalSlice(b, self.own_lo, self.own_hi)
        self._plans = None
        self._exchanges = {}
        self._halo_scalars = {}
        self._ndata = {}

    # ---- graph-object surface the layers use (stgraph_base.py:51-59 names) ------------------------------
    def get_num_nodes(self) -> int:
        return self.graph.get_num_nodes()

    def get_num_edges(self) -> int:
        return self.graph.get_num_edges()

    def num_local_nodes(self) -> int:
        return self.n_own

    def local(self, t: torch.Tensor) -> torch.Tensor:
        """My rows of a replicated ``[N, ...]`` tensor."""
        return t[self.own_lo:self.own_hi]

    def set_ndata(self, name: str, value: torch.Tensor):
        """Node data of MY vertices (``[n_own, ...]``)."""
        if value.shape[0] != self.n_own:
            raise ValueError(f"ndata '{name}' must have {self.n_own} rows (the local vertices), got {value.shape[0]}")
        self._ndata[name] = value

    def get_ndata(self, name: str):
        return self._ndata.get(name)

    def degree_norm(self) -> torch.Tensor:
        """``in_degree^-0.5`` of my vertices, ``[n_own, 1]`` (``benchmarking/gcn/seastar/train.py:53-57``)."""
        return self.local(self.graph.degree_norm()).contiguous()

    # ---- exchange plans / engines ---------------------------------------------------------------------------
    def halo_plans(self, group=None):
        """Halo-only exchange plans: (forward, backward); built once (collective: every rank must call it)."""
        from .halo import HaloPlan

        if self._plans is None:
            g = self.graph
            grp = group if group is not None else self.group
            self._plans = (HaloPlan(g._forward_graph, self.bounds, self.bounds, self.rank, self.world, grp),
                           HaloPlan(g._backward_graph, self.bounds, self.bounds, self.rank, self.world, grp))
        return self._plans

    def halo_scalars(self, direction: str, own_values: torch.Tensor) -> torch.Tensor:
        """Per-vertex scalars (e.g. ``norm``) of the halo vertices of one direction: one exchange, then cached."""
        key = (direction, own_values.data_ptr(), own_values._version)
        if key not in self._halo_scalars:
            plan = self.halo_plans()[0 if direction == "fwd" else 1]
            self._halo_scalars = {k: v for k, v in self._halo_scalars.items() if k[0] != direction}
            self._halo_scalars[key] = (plan.halo_vector(own_values.reshape(-1).contiguous()), own_values)
        return self._halo_scalars[key][0]

    def exchange(self, direction: str, feat: int, nbr_scale: torch.Tensor | None, edge_scale: torch.Tensor | None = None):
        """The :class:`HaloExchange` of (direction, width, source scale, edge scale); collective on first use."""
        from .exchange import HaloExchange

        def ident(t):
            return None if t is None else (t.data_ptr(), t._version)

        key = (direction, int(feat), ident(nbr_scale), ident(edge_scale))
        ex = self._exchanges.get(key)
        if ex is None:
            plan = self.halo_plans()[0 if direction == "fwd" else 1]
            ns_own = None if nbr_scale is None else nbr_scale.reshape(-1).contiguous()
            ns_halo = None if nbr_scale is None else self.halo_scalars(direction, ns_own)
            es = None
            if edge_scale is not None:
                es = edge_scale.reshape(-1).contiguous()
                if es.numel() != self.graph.get_num_edges():
                    raise ValueError(f"edge_scale must hold one value per edge of the WHOLE graph ({self.graph.get_num_edges()}), "
                                     f"got {es.numel()}")
            ex = HaloExchange(plan, feat, ns_own, ns_halo, edge_scale=es)
            ex._scale_ref = (nbr_scale, edge_scale)
            self._exchanges[key] = ex
        return ex

    def aggregate(self, direction: str, x_own: torch.Tensor, nbr_scale=None, row_scale=None, out=None,
                  edge_scale=None) -> torch.Tensor:
        """``out[v] = row_scale[v] * sum_{u in nbrs(v)} nbr_scale[u] * edge_scale[eid(u, v)] * x[u]`` for my vertices ``v``;
        the vertex tensors hold local rows only (``[n_own, ...]``), ``edge_scale`` holds the WHOLE graph's edge values in
        edge-id order (replicated, like the structure); remote source rows travel as halo (see ``dist/exchange.py``)."""
        if x_own.shape[0] != self.n_own:
            raise ValueError(f"x_own must hold the {self.n_own} local rows, got {x_own.shape[0]}")
        x_own = x_own.contiguous()
        feat = x_own.numel() // max(self.n_own, 1)
        if out is None:
            out = torch.empty_like(x_own)
        rs = None if row_scale is None else row_scale.reshape(-1).contiguous()
        return self.exchange(direction, feat, nbr_scale, edge_scale).aggregate(x_own.reshape(self.n_own, feat), rs, out)

    def local_rows(self, direction: str):
        return self.own_lo, self.own_hi


class _PartitionedGcnAggregate(torch.autograd.Function):
    """Autograd bridge of the distributed GCN aggregation: forward on the in-edge CSR, backward on the out-edge CSR
    (a gather both ways, like the reference's K0/K1 pair, ``gcn_conv.py:162-166``)."""

    @staticmethod
    def forward(ctx, pg, h, norm, edge_weight):
        ctx.pg, ctx.norm, ctx.edge_weight = pg, norm, edge_weight
        return pg.aggregate("fwd", h, norm, norm, edge_scale=edge_weight)

    @staticmethod
    def backward(ctx, grad_out):
        # the out-edge CSR carries the forward edge ids: the same weights apply (no gradient for them, like the reference)
        return None, ctx.pg.aggregate("bwd", grad_out.contiguous(), ctx.norm, ctx.norm, edge_scale=ctx.edge_weight), None, None


def partitioned_gcn_aggregate(pg: PartitionedGraph, h: torch.Tensor, norm: torch.Tensor, edge_weight=None) -> torch.Tensor:
    if edge_weight is not None and edge_weight.requires_grad:
        raise RuntimeError("edge_weight gets no gradient on a PartitionedGraph (nor in the reference's GCNConv)")
    return _PartitionedGcnAggregate.apply(pg, h, norm, edge_weight)


def all_reduce_gradients(module_or_params, group=None):
    """Sum the (replicated) parameters' gradients over the ranks: each rank back-propagates through its own rows only
    ("dense weights are replicated; their gradients need one tiny all-reduce")."""
    import torch.distributed as dist

    params = module_or_params.parameters() if hasattr(module_or_params, "parameters") else module_or_params
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


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
