"""WikiMath-shaped static-temporal dataset (API of ``stgraph/dataset/temporal/wikimath_dataloader.py:71-193``)."""
from __future__ import annotations

import numpy as np

from ...utils import synthetic
from ..stgraph_dataset import check_lags_and_cutoff
from .stgraph_temporal_dataset import STGraphTemporalDataset


class WikiMathDataLoader(STGraphTemporalDataset):
    """1068 nodes, 27079 directed weighted edges, 731 daily snapshots of one scalar per node.

    ``get_edge_weights()`` is ordered like the reference's: by (dst, src), the order ``StaticGraph`` numbers its edges
    in (``wikimath_dataloader.py:155-163``; SURVEY.md trap T8).  Targets are standardised per node over time
    (``165-178``)."""

    TIME_PERIODS = 731

    def __init__(self, verbose: bool = False, lags: int = 8, cutoff_time: int | None = None, redownload: bool = False,
                 seed: int = 0) -> None:
        super().__init__()
        check_lags_and_cutoff(lags, cutoff_time)
        self.name = "WikiMath (synthetic)"
        self._verbose = verbose
        self._lags = lags
        self._cutoff_time = cutoff_time
        self._log("generating (no network: synthetic data of the dataset's shape)" + (" again" if redownload else ""))
        d = synthetic.wikimaths_shaped(seed=seed, device="cpu", num_timestamps=self.TIME_PERIODS, lags=lags)
        self.gdata["total_timestamps"] = min(self.TIME_PERIODS, cutoff_time) if cutoff_time is not None else self.TIME_PERIODS
        self.gdata["num_nodes"] = int(d["num_nodes"])
        src, dst = d["src"].numpy(), d["dst"].numpy()
        self.gdata["num_edges"] = int(src.shape[0])
        self._edge_list = list(zip(src.tolist(), dst.tolist()))
        order = np.lexsort((src, dst))                      # (dst, src): the order the edge ids are assigned in
        self._edge_weights = d["edge_weight"].numpy()[order]
        raw = d["targets"].numpy()[: self.gdata["total_timestamps"]]
        self._all_targets = (raw - raw.mean(axis=0)) / (raw.std(axis=0) + 10 ** -10)

    def get_edges(self) -> list:
        return self._edge_list

    def get_edge_weights(self) -> np.ndarray:
        return self._edge_weights

    def get_all_targets(self) -> np.ndarray:
        return self._all_targets
