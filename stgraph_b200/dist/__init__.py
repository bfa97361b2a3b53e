"""1D destination-row partitioning + halo exchange for multi-GPU aggregation (new; the reference is single-GPU)."""
from .partition import PartitionedGraph, edge_balanced_bounds, exchange_rows

__all__ = ["PartitionedGraph", "edge_balanced_bounds", "exchange_rows"]
