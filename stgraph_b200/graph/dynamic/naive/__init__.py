from ..naive_graph import *  # noqa: F401,F403  (reference module path stgraph.graph.dynamic.naive.naive_graph)
