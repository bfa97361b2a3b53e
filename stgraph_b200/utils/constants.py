"""Constants used by the layers (mirrors ``stgraph/utils/constants.py:1-17``)."""
from enum import Enum


class SizeConstants(Enum):
    """Expected ranks of well-known tensors."""

    NODE_NORM_SIZE = 2   # graph.get_ndata("norm") must be [num_nodes, 1] (gcn_conv.py:151-156)
