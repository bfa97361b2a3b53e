"""Vertex-centric GCN layer (mirror of ``stgraph/nn/pytorch/static/gcn_conv.py:78-189``).

Same constructor, parameter names (``weight``, ``bias``) and forward contract, so state_dicts
interchange with the reference layer:

    h' = act( norm * sum_{u in in(v)} (h W)[u] * norm[u] [* w_e] + b )

``X . W`` stays a library GEMM (it is not the hot path; only its weight gradient, a reduction over all
vertices, runs in our split-M kernel on large graphs: ``ops_gcn.dense_transform``); the aggregation is the
traced vertex program, lowered to the fused sm_100a gather kernel (``csrc/agg.cu``).
"""
from __future__ import annotations

from typing import Callable

import torch
from torch import Tensor, nn

from ....compiler import STGraph
from ....compiler.backend.pytorch.torch_callback import STGraphBackendTorch
from ....ops_gcn import dense_transform
from ....utils.constants import SizeConstants


class GCNConv(nn.Module):
    def __init__(self, in_channels: int, out_channels: int,
                 activation: Callable[..., Tensor] | None = None, bias: bool = True) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.bias = None
        self.activation = activation
        self.stgraph = STGraph(STGraphBackendTorch())
        self.reset_parameters()

    def reset_parameters(self) -> None:
        nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def _forward_partitioned(self, graph, h: Tensor, edge_weight: Tensor | None) -> Tensor:
        """``graph`` is a :class:`stgraph_b200.dist.PartitionedGraph`: ``h`` and the result hold this rank's rows."""
        from ....dist.partition import partitioned_gcn_aggregate

        norm = graph.get_ndata("norm")
        if norm is None:
            raise KeyError("PartitionedGraph passed to GCNConv forward pass does not contain 'norm' node data")
        if (len(norm.shape) != SizeConstants.NODE_NORM_SIZE.value or norm.shape[1] != 1
                or norm.shape[0] != graph.num_local_nodes()):
            raise ValueError("Node data 'norm' passed to GCNConv should be of shape (num_local_nodes, 1)")
        h = dense_transform(h, self.weight)
        h = partitioned_gcn_aggregate(graph, h, norm, edge_weight)      # edge_weight: [E, 1] of the WHOLE graph, edge-id order
        if self.bias is not None:
            h = h + self.bias
        if self.activation:
            h = self.activation(h)
        return h

    def forward(self, graph, h: Tensor, edge_weight: Tensor | None = None) -> Tensor:
        if hasattr(graph, "num_local_nodes"):          # row-partitioned graph (multi-GPU): local rows in, local rows out
            return self._forward_partitioned(graph, h, edge_weight)
        norm = graph.get_ndata("norm")
        if norm is None:
            raise KeyError("StaticGraph passed to GCNConv forward pass does not contain 'norm' node data")
        if (len(norm.shape) != SizeConstants.NODE_NORM_SIZE.value or norm.shape[1] != 1
                or norm.shape[0] != graph.get_num_nodes()):
            raise ValueError("Node data 'norm' passed to GCNConv should be of shape (num_nodes, 1)")

        h = dense_transform(h, self.weight)

        if edge_weight is None:

            @self.stgraph.compile(gnn_module=self)
            def nb_compute(v):
                return sum([nb.h * nb.norm for nb in v.innbs]) * v.norm

            h = nb_compute(g=graph, n_feats={"norm": norm, "h": h})
        else:

            @self.stgraph.compile(gnn_module=self)
            def nb_compute(v):
                return sum([nb_edge.src.norm * nb_edge.src.h * nb_edge.edge_weight
                            for nb_edge in v.inedges]) * v.norm

            h = nb_compute(g=graph, n_feats={"norm": norm, "h": h}, e_feats={"edge_weight": edge_weight})

        if self.bias is not None:
            h = h + self.bias
        if self.activation:
            h = self.activation(h)
        return h
