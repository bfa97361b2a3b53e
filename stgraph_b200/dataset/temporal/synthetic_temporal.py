"""Shared machinery of the synthetic static-temporal loaders (the reference's loaders differ only in their JSON
layout; ``stgraph/dataset/temporal/*_dataloader.py``): a seeded directed graph with the dataset's node / edge counts,
edge weights ordered by (dst, src) like the reference sorts them (e.g. ``pedalme_dataloader.py:150-158``; SURVEY.md
trap T8) and a ``[T, N]`` signal with per-node seasonality that the subclasses window into features / targets."""
from __future__ import annotations

import numpy as np

from ..stgraph_dataset import check_lags_and_cutoff
from .stgraph_temporal_dataset import STGraphTemporalDataset


def synthetic_edges(num_nodes: int, num_edges: int, rng, full: bool = False):
    """``num_edges`` distinct directed pairs (self loops allowed when the dataset is a complete graph with loops:
    PedalMe 15 x 15, Windmill 319 x 319 / 26 x 26 / 11 x 11)."""
    if full:
        a, b = np.divmod(np.arange(num_nodes * num_nodes), num_nodes)
        assert a.shape[0] == num_edges
        return a.astype(np.int64), b.astype(np.int64)
    key = rng.choice(num_nodes * num_nodes, size=min(num_nodes * num_nodes, 2 * num_edges + num_nodes), replace=False)
    key = key[(key // num_nodes) != (key % num_nodes)][:num_edges]
    assert key.shape[0] == num_edges
    return key // num_nodes, key % num_nodes


def synthetic_signal(timestamps: int, num_nodes: int, rng) -> np.ndarray:
    t = np.arange(timestamps, dtype=np.float64)[:, None]
    period = rng.uniform(5, 60, size=(1, num_nodes))
    return np.sin(2 * np.pi * t / period + rng.uniform(0, 6.28, size=(1, num_nodes))) + 0.3 * rng.standard_normal((timestamps, num_nodes))


def standardize(x: np.ndarray) -> np.ndarray:
    return (x - x.mean(axis=0)) / (x.std(axis=0) + 10 ** -10)


class SyntheticTemporalLoader(STGraphTemporalDataset):
    NUM_NODES = 0
    NUM_EDGES = 0
    TIME_PERIODS = 0
    FULL_GRAPH = False
    UNIT_WEIGHTS = False

    def _build(self, name: str, verbose: bool, lags: int, cutoff_time, redownload: bool, seed: int) -> np.ndarray:
        """Common constructor body; returns the raw ``[total_timestamps, N]`` signal."""
        check_lags_and_cutoff(lags, cutoff_time)
        self.name = name + " (synthetic)"
        self._verbose, self._lags, self._cutoff_time = verbose, lags, cutoff_time
        self._log("generating (no network: synthetic data of the dataset's shape)" + (" again" if redownload else ""))
        rng = np.random.default_rng(seed)
        total = min(self.TIME_PERIODS, cutoff_time) if cutoff_time is not None else self.TIME_PERIODS
        self.gdata["total_timestamps"] = total
        self.gdata["num_nodes"], self.gdata["num_edges"] = self.NUM_NODES, self.NUM_EDGES
        src, dst = synthetic_edges(self.NUM_NODES, self.NUM_EDGES, rng, self.FULL_GRAPH)
        self._edge_list = list(zip(src.tolist(), dst.tolist()))
        w = np.ones(self.NUM_EDGES) if self.UNIT_WEIGHTS else rng.uniform(0.1, 1.0, size=self.NUM_EDGES)
        self._edge_weights = w[np.lexsort((src, dst))]
        return synthetic_signal(self.TIME_PERIODS, self.NUM_NODES, rng)[:total]

    def get_edges(self) -> list:
        return self._edge_list

    def get_edge_weights(self) -> np.ndarray:
        return self._edge_weights

    def get_all_targets(self):
        return self._all_targets
