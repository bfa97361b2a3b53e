from ..gpma_graph import *  # noqa: F401,F403  (reference module path stgraph.graph.dynamic.gpma.gpma_graph)
