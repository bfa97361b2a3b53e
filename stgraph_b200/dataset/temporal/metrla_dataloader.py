"""METR-LA traffic shaped loader (API of ``stgraph/dataset/temporal/metrla_dataloader.py:73-226``)."""
from __future__ import annotations

import numpy as np

from .synthetic_temporal import SyntheticTemporalLoader, synthetic_signal


class METRLADataLoader(SyntheticTemporalLoader):
    """207 loop detectors, 1722 weighted edges, 100 time periods of two channels (speed, time of day); sliding windows of
    ``num_timesteps_in`` inputs ``[207, 2, in]`` and ``num_timesteps_out`` speed targets ``[207, out]``, z-scored per
    channel (``metrla_dataloader.py:177-210``)."""

    NUM_NODES, NUM_EDGES, TIME_PERIODS = 207, 1722, 100

    def __init__(self, verbose: bool = False, num_timesteps_in: int = 12, num_timesteps_out: int = 12,
                 cutoff_time: int | None = None, redownload: bool = False, seed: int = 0) -> None:
        super().__init__()
        for name, v in (("num_timesteps_in", num_timesteps_in), ("num_timesteps_out", num_timesteps_out)):
            if not isinstance(v, int):
                raise TypeError(f"{name} must be of type int")
            if v < 0:
                raise ValueError(f"{name} must be a positive integer")
        self._num_timesteps_in, self._num_timesteps_out = num_timesteps_in, num_timesteps_out
        speed = self._build("METRLA", verbose, 0, cutoff_time, redownload, seed)            # [T, N]
        total = self.gdata["total_timestamps"]
        tod = np.tile((np.arange(total) % 288 / 288.0)[:, None], (1, self.NUM_NODES))
        x = np.stack([speed, tod], axis=1).transpose(2, 1, 0).astype(np.float32)               # [N, 2, T]
        x = (x - x.mean(axis=(0, 2)).reshape(1, -1, 1)) / x.std(axis=(0, 2)).reshape(1, -1, 1)
        span = num_timesteps_in + num_timesteps_out
        feats, targs = [], []
        for i in range(max(x.shape[2] - span + 1, 0)):
            feats.append(x[:, :, i:i + num_timesteps_in])
            targs.append(x[:, 0, i + num_timesteps_in:i + span])
        self._all_features, self._all_targets = np.array(feats), np.array(targs)

    def get_all_features(self) -> np.ndarray:
        return self._all_features
