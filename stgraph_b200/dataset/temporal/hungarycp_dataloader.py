"""Hungary chicken-pox shaped loader (API of ``stgraph/dataset/temporal/hungarycp_dataloader.py:64-173``)."""
from __future__ import annotations

from .synthetic_temporal import SyntheticTemporalLoader


class HungaryCPDataLoader(SyntheticTemporalLoader):
    """20 counties, 102 unit-weight edges, 521 weekly case counts; ``get_all_targets()`` is a list of
    ``total_timestamps - lags`` arrays of shape ``(20,)``: the raw count at ``t + lags`` (``hungarycp_dataloader.py:154-161``)."""

    NUM_NODES, NUM_EDGES, TIME_PERIODS, UNIT_WEIGHTS = 20, 102, 521, True

    def __init__(self, verbose: bool = False, lags: int = 4, cutoff_time: int | None = None, redownload: bool = False,
                 seed: int = 0) -> None:
        super().__init__()
        fx = self._build("Hungary_Chickenpox", verbose, lags, cutoff_time, redownload, seed)
        total = self.gdata["total_timestamps"]
        self._all_targets = [fx[i + lags, :].T for i in range(total - lags)]
