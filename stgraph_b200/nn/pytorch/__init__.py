from .static.gat_conv import GATConv
from .static.gcn_conv import GCNConv
from .temporal.tgcn import TGCN

__all__ = ["GCNConv", "GATConv", "TGCN"]
