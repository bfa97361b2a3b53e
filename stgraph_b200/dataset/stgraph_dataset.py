"""Dataset loaders with the reference's surface on SYNTHETIC data.

The reference's loaders (``stgraph/dataset/stgraph_dataset.py:19-215``) download JSON files from GitHub and cache
them under ``~/.stgraph``; this environment has no network, and the measured path only needs inputs of the right
SHAPE.  Every loader here keeps the reference's constructor arguments, validation messages, ``gdata`` keys and getters
(``get_edges / get_edge_weights / get_all_features / get_all_targets``) and fills them from the seeded generators of
``stgraph_b200/utils/synthetic.py``, sized exactly like the datasets the reference's own tests pin
(``tests/dataset/**``: Cora 2708 / 10556 / 1433 / 7, WikiMath 1068 / 27079 / 731, EnglandCovid 129 x 61).
``name`` carries a ``(synthetic)`` suffix so nobody mistakes the numbers for the real data.
"""
from __future__ import annotations


class STGraphDataset:
    """Common state of every loader: ``name``, ``gdata`` (graph meta data), verbosity."""

    def __init__(self) -> None:
        self.name = ""
        self.gdata = {}
        self._dataset = {}
        self._verbose = False
        self._init_graph_data()

    def _init_graph_data(self) -> None:
        self.gdata = {}

    def _log(self, message: str) -> None:
        if self._verbose:
            print(f"[stgraph_b200.dataset] {self.name}: {message}")

    # the reference's cache / download hooks: nothing is downloaded or cached here
    def _has_dataset_cache(self) -> bool:
        return False

    def _delete_cached_dataset(self) -> None:
        pass


def check_lags_and_cutoff(lags, cutoff_time) -> None:
    """The argument checks every temporal / dynamic loader of the reference performs, with its messages
    (e.g. ``stgraph/dataset/temporal/wikimath_dataloader.py:81-89``)."""
    if not isinstance(lags, int):
        raise TypeError("lags must be of type int")
    if lags < 0:
        raise ValueError("lags must be a positive integer")
    if cutoff_time is not None and not isinstance(cutoff_time, int):
        raise TypeError("cutoff_time must be of type int")
    if cutoff_time is not None and cutoff_time < 0:
        raise ValueError("cutoff_time must be a positive integer")
