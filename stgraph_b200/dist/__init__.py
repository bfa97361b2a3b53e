"""1D destination-row partitioning + halo exchange for multi-GPU aggregation (new; the reference is single-GPU)."""
from .halo import HaloPlan
from .peer import PeerBlocks
from .partition import PartitionedGraph, edge_balanced_bounds, exchange_rows

__all__ = ["PartitionedGraph", "HaloPlan", "PeerBlocks", "edge_balanced_bounds", "exchange_rows"]
