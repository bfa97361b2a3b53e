from ..pcsr_graph import *  # noqa: F401,F403  (reference module path stgraph.graph.dynamic.pcsr.pcsr_graph)
