"""The reference's own dataset tests (``tests/dataset/static/test_CoraDataLoader.py``,
``tests/dataset/temporal/test_WikiMathDataLoader.py``, ``tests/dataset/dynamic/test_EnglandCovidDataLoader.py``)
restated against the synthetic loaders: same shapes, same ``gdata`` keys, same validation messages."""
import numpy as np
import pytest

import stgraph_b200.compat  # noqa: F401
from stgraph.dataset import CoraDataLoader, EnglandCovidDataLoader, WikiMathDataLoader


def _cora_check(cora):
    assert len(cora._edge_list) == 10556
    assert cora._all_features.shape == (2708, 1433)
    assert cora._all_targets.shape == (2708,)
    assert cora.gdata["num_nodes"] == 2708 and cora.gdata["num_edges"] == 10556
    assert cora.gdata["num_feats"] == 1433 and cora.gdata["num_classes"] == 7
    edge_list = cora.get_edges()
    assert len(edge_list) == 10556 and len(edge_list[0]) == 2
    assert cora.get_all_features().shape == (2708, 1433) and cora.get_all_targets().shape == (2708,)
    assert len(set(edge_list)) == 10556 and all((b, a) in set(edge_list) for a, b in edge_list[:200])   # symmetric, no duplicates
    assert np.allclose(cora.get_all_features().sum(1)[cora.get_all_features().sum(1) > 0], 1.0)           # row-normalised


def test_cora_loader():
    _cora_check(CoraDataLoader(verbose=False))
    _cora_check(CoraDataLoader(redownload=True))


def _wiki_check(wiki):
    assert wiki.gdata["total_timestamps"] == (731 if not wiki._cutoff_time else wiki._cutoff_time)
    assert wiki.gdata["num_nodes"] == 1068 and wiki.gdata["num_edges"] == 27079
    edges, w, y = wiki.get_edges(), wiki.get_edge_weights(), wiki.get_all_targets()
    assert len(edges) == 27079 and all(len(e) == 2 for e in edges) and len(w) == 27079
    assert y.shape == (wiki.gdata["total_timestamps"], 1068)
    assert np.allclose(y.mean(axis=0), 0, atol=1e-5) and np.allclose(y.std(axis=0), 1, atol=1e-4)


def test_wikimath_loader():
    for kw in ({}, {"redownload": True}, {"lags": 4}, {"cutoff_time": 500}):
        _wiki_check(WikiMathDataLoader(**kw))
    for kw, exc, msg in (({"lags": "lags"}, TypeError, "lags must be of type int"),
                         ({"lags": -1}, ValueError, "lags must be a positive integer"),
                         ({"cutoff_time": "time"}, TypeError, "cutoff_time must be of type int"),
                         ({"cutoff_time": -1}, ValueError, "cutoff_time must be a positive integer")):
        with pytest.raises(exc) as info:
            WikiMathDataLoader(**kw)
        assert str(info.value) == msg


def test_wikimath_weights_follow_edge_id_order():
    """Weights are stored in (dst, src) order -- the order StaticGraph assigns edge ids in (SURVEY.md trap T8)."""
    wiki = WikiMathDataLoader(cutoff_time=10)
    from stgraph_b200.utils import synthetic

    d = synthetic.wikimaths_shaped(seed=0, device="cpu")
    src, dst, w = d["src"].numpy(), d["dst"].numpy(), d["edge_weight"].numpy()
    order = np.lexsort((src, dst))
    np.testing.assert_array_equal(wiki.get_edge_weights(), w[order])


def _covid_check(ec):
    total = ec.gdata["total_timestamps"]
    assert total == (61 if not ec._cutoff_time else ec._cutoff_time)
    assert len(ec.gdata["num_nodes"]) == total and all(v == 129 for v in ec.gdata["num_nodes"].values())
    assert len(ec.gdata["num_edges"]) == total
    edges, w = ec.get_edges(), ec.get_edge_weights()
    assert len(edges) == total and len(edges[0][0]) == 2 and len(w) == total
    assert all(len(edges[i]) == len(w[i]) for i in range(total))
    feats, targets = ec.get_all_features(), ec.get_all_targets()
    assert len(feats) == total - ec._lags and feats[0].shape == (129, ec._lags)
    assert len(targets) == total - ec._lags and targets[0].shape == (129,)
    assert set(edges[0]) != set(edges[1])                 # the graph really changes over time


def test_england_covid_loader():
    for kw in ({}, {"cutoff_time": 30}, {"lags": 12}, {"redownload": True}):
        _covid_check(EnglandCovidDataLoader(**kw))
