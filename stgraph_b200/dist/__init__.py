"""1D vertex partitioning + halo exchange for multi-GPU aggregation (new; the reference is single-GPU)."""
from .exchange import HaloExchange
from .halo import HaloPlan
from .partition import (PartitionedGraph, all_reduce_gradients, cost_balanced_bounds, edge_balanced_bounds, exchange_rows,
                        partitioned_gcn_aggregate)
from .peer import PeerBlocks

__all__ = ["PartitionedGraph", "HaloPlan", "HaloExchange", "PeerBlocks", "edge_balanced_bounds", "cost_balanced_bounds",
           "exchange_rows", "partitioned_gcn_aggregate", "all_reduce_gradients"]
