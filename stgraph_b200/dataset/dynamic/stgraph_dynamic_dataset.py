"""Base class of dynamic datasets (``stgraph/dataset/dynamic/stgraph_dynamic_dataset.py``): per-timestamp meta data."""
from ..stgraph_dataset import STGraphDataset


class STGraphDynamicDataset(STGraphDataset):
    def _init_graph_data(self) -> None:
        self.gdata = {"num_nodes": {}, "num_edges": {}, "total_timestamps": 0}
