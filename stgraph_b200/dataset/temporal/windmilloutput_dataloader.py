"""Windmill energy output shaped loader (API of ``stgraph/dataset/temporal/windmilloutput_dataloader.py:88-221``)."""
from __future__ import annotations

from .synthetic_temporal import SyntheticTemporalLoader, standardize


class WindmillOutputDataLoader(SyntheticTemporalLoader):
    """Hourly output of 319 / 26 / 11 windmills (``size`` = large / medium / small) over 17472 hours on the complete
    weighted graph with self loops; targets: a list of ``total_timestamps`` standardised arrays of shape ``(N,)``
    (``windmilloutput_dataloader.py:198-206``: no lag offset)."""

    TIME_PERIODS, FULL_GRAPH = 17472, True
    SIZES = {"large": (319, 101761), "medium": (26, 676), "small": (11, 121)}

    def __init__(self, verbose: bool = False, lags: int = 8, cutoff_time: int | None = None, size: str = "large",
                 redownload: bool = False, seed: int = 0) -> None:
        super().__init__()
        if not isinstance(size, str):
            raise TypeError("size must be of type string")
        if size not in self.SIZES:
            raise ValueError("size must take either of the following values : large, medium or small")
        self._size = size
        self.NUM_NODES, self.NUM_EDGES = self.SIZES[size]
        block = self._build("WindMill_" + size, verbose, lags, cutoff_time, redownload, seed)
        z = standardize(block)
        self._all_targets = [z[i, :].T for i in range(self.gdata["total_timestamps"])]
