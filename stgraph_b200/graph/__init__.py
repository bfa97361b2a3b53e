"""Graph containers (mirror of ``stgraph.graph``)."""
from .dynamic.gpma_graph import GPMAGraph
from .dynamic.naive_graph import NaiveGraph
from .dynamic.pcsr_graph import PCSRGraph
from .static.static_graph import StaticGraph
from .stgraph_base import STGraphBase

__all__ = ["STGraphBase", "StaticGraph", "NaiveGraph", "PCSRGraph", "GPMAGraph"]
