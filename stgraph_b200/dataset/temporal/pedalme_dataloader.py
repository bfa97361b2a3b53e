"""PedalMe (London bicycle deliveries) shaped loader (API of ``stgraph/dataset/temporal/pedalme_dataloader.py:64-189``)."""
from __future__ import annotations

import numpy as np

from .synthetic_temporal import SyntheticTemporalLoader


class PedalMeDataLoader(SyntheticTemporalLoader):
    """15 localities, the complete 15 x 15 weighted graph (225 edges), 36 weekly counts; targets are an array of shape
    ``(total_timestamps - lags, 15)`` (``pedalme_dataloader.py:160-174``)."""

    NUM_NODES, NUM_EDGES, TIME_PERIODS, FULL_GRAPH = 15, 225, 36, True

    def __init__(self, verbose: bool = False, lags: int = 4, cutoff_time: int | None = None, redownload: bool = False,
                 seed: int = 0) -> None:
        super().__init__()
        x = self._build("PedalMe", verbose, lags, cutoff_time, redownload, seed)
        self._all_targets = np.array([x[i + lags, :].T for i in range(x.shape[0] - lags)])
