"""Cora-shaped static dataset (API of ``stgraph/dataset/static/cora_dataloader.py:67-142``)."""
from __future__ import annotations

import numpy as np

from ...utils import synthetic
from .stgraph_static_dataset import STGraphStaticDataset


class CoraDataLoader(STGraphStaticDataset):
    """2708 nodes, 10556 directed edges (5278 undirected pairs), 1433 sparse row-normalised features, 7 classes."""

    def __init__(self, verbose: bool = False, redownload: bool = False, seed: int = 0) -> None:
        super().__init__()
        self.name = "Cora (synthetic)"
        self._verbose = verbose
        self._log("generating (no network: synthetic data of the dataset's shape)" + (" again" if redownload else ""))
        d = synthetic.cora_shaped(seed=seed, device="cpu")
        src, dst = d["src"].numpy(), d["dst"].numpy()
        self._edge_list = list(zip(src.tolist(), dst.tolist()))
        self._all_features = d["features"].numpy()
        self._all_targets = d["labels"].numpy().astype(np.int64)
        self.gdata["num_nodes"] = int(d["num_nodes"])
        self.gdata["num_edges"] = len(self._edge_list)
        self.gdata["num_feats"] = int(self._all_features.shape[1])
        self.gdata["num_classes"] = int(len(set(self._all_targets.tolist())))

    def get_edges(self) -> list:
        return self._edge_list

    def get_all_features(self) -> np.ndarray:
        return self._all_features

    def get_all_targets(self) -> np.ndarray:
        return self._all_targets
