"""Base class of static datasets (``stgraph/dataset/static/stgraph_static_dataset.py:11-24``)."""
from ..stgraph_dataset import STGraphDataset


class STGraphStaticDataset(STGraphDataset):
    def _init_graph_data(self) -> None:
        self.gdata = {"num_nodes": 0, "num_edges": 0}
