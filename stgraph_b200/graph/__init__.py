"""Graph containers (mirror of ``stgraph.graph``)."""
from .stgraph_base import STGraphBase
from .static.static_graph import StaticGraph

__all__ = ["STGraphBase", "StaticGraph"]
