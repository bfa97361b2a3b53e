"""Base class of static-temporal datasets (``stgraph/dataset/temporal/stgraph_temporal_dataset.py:19-33``)."""
from ..stgraph_dataset import STGraphDataset


class STGraphTemporalDataset(STGraphDataset):
    def _init_graph_data(self) -> None:
        self.gdata = {"num_nodes": 0, "num_edges": 0, "total_timestamps": 0}
