from ..gpma_graph import *  # noqa: F401,F403
