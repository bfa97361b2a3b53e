"""Montevideo bus inflow shaped loader (API of ``stgraph/dataset/temporal/montevideobus_dataloader.py:72-221``)."""
from __future__ import annotations

import numpy as np

from .synthetic_temporal import SyntheticTemporalLoader, standardize


class MontevideoBusDataLoader(SyntheticTemporalLoader):
    """675 bus stops, 690 weighted edges, 744 hourly inflows; features ``(T - lags, 675, lags)`` are the ``lags``
    previous standardised values, targets ``(T - lags, 675)`` the next one (``montevideobus_dataloader.py:168-205``)."""

    NUM_NODES, NUM_EDGES, TIME_PERIODS = 675, 690, 744

    def __init__(self, verbose: bool = False, lags: int = 4, cutoff_time: int | None = None, redownload: bool = False,
                 seed: int = 0) -> None:
        super().__init__()
        y = standardize(self._build("Montevideo_Bus", verbose, lags, cutoff_time, redownload, seed))
        self._all_features = np.array([y[i:i + lags, :].T for i in range(len(y) - lags)])
        self._all_targets = np.array([y[i + lags, :].T for i in range(len(y) - lags)])

    def get_all_features(self) -> np.ndarray:
        return self._all_features
