"""EnglandCovid-shaped dynamic dataset (API of ``stgraph/dataset/dynamic/england_covid_dataloader.py``).

61 daily mobility graphs over 129 regions whose edge sets change from day to day; features are the ``lags`` previous
values of a per-node signal, targets the next value -- so ``get_all_features()`` / ``get_all_targets()`` have
``total_timestamps - lags`` entries of shape ``(129, lags)`` / ``(129,)`` like the reference
(``tests/dataset/dynamic/test_EnglandCovidDataLoader.py:6-50``)."""
from __future__ import annotations

import numpy as np

from ..stgraph_dataset import check_lags_and_cutoff
from .stgraph_dynamic_dataset import STGraphDynamicDataset


class EnglandCovidDataLoader(STGraphDynamicDataset):
    TIME_PERIODS = 61
    NUM_NODES = 129

    def __init__(self, verbose: bool = False, lags: int = 8, cutoff_time: int | None = None, redownload: bool = False,
                 seed: int = 0) -> None:
        super().__init__()
        check_lags_and_cutoff(lags, cutoff_time)
        self.name = "EnglandCOVID (synthetic)"
        self._verbose = verbose
        self._lags = lags
        self._cutoff_time = cutoff_time
        self._log("generating (no network: synthetic data of the dataset's shape)" + (" again" if redownload else ""))
        rng = np.random.default_rng(seed)
        total = min(self.TIME_PERIODS, cutoff_time) if cutoff_time is not None else self.TIME_PERIODS
        n = self.NUM_NODES
        self.gdata["total_timestamps"] = total
        # a mobility graph that drifts: ~1500 directed edges, about a tenth replaced every day
        live = set()
        while len(live) < 1500:
            a, b = rng.integers(0, n, 2)
            live.add((int(a), int(b)))
        self._edge_list, self._edge_weights = [], []
        for t in range(total):
            edges = sorted(live)
            self._edge_list.append(edges)
            self._edge_weights.append(rng.uniform(0.05, 1.0, size=len(edges)))
            self.gdata["num_nodes"][str(t)] = n
            self.gdata["num_edges"][str(t)] = len(edges)
            for i in rng.choice(len(edges), size=150, replace=False):
                live.discard(edges[i])
            while len(live) < 1500:
                a, b = rng.integers(0, n, 2)
                live.add((int(a), int(b)))
        signal = rng.standard_normal((total, n)).cumsum(axis=0)
        signal = (signal - signal.mean(axis=0)) / (signal.std(axis=0) + 10 ** -10)
        self._all_features = [signal[t:t + lags].T.copy() for t in range(total - lags)]
        self._all_targets = [signal[t + lags].copy() for t in range(total - lags)]

    def get_edges(self) -> list:
        return self._edge_list

    def get_edge_weights(self) -> list:
        return self._edge_weights

    def get_all_features(self) -> list:
        return self._all_features

    def get_all_targets(self) -> list:
        return self._all_targets
