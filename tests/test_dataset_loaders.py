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


# ---- the remaining static-temporal loaders (reference: tests/dataset/temporal/test_*DataLoader.py) -----------------
def _lag_errors(cls):
    for bad, exc, msg in (("lags", TypeError, "lags must be of type int"), (-1, ValueError, "lags must be a positive integer")):
        with pytest.raises(exc) as e:
            cls(lags=bad)
        assert str(e.value) == msg
    for bad, exc, msg in (("time", TypeError, "cutoff_time must be of type int"),
                          (-1, ValueError, "cutoff_time must be a positive integer")):
        with pytest.raises(exc) as e:
            cls(cutoff_time=bad)
        assert str(e.value) == msg


def test_hungarycp_loader_shapes_and_errors():
    from stgraph_b200.dataset import HungaryCPDataLoader

    for kw in ({"verbose": True}, {"lags": 6}, {"cutoff_time": 100}, {"redownload": True}):
        h = HungaryCPDataLoader(**kw)
        assert h.gdata["total_timestamps"] == (521 if not h._cutoff_time else h._cutoff_time)
        assert h.gdata["num_nodes"] == 20 and h.gdata["num_edges"] == 102
        assert len(h.get_edges()) == 102 and len(h.get_edges()[0]) == 2 and len(h.get_edge_weights()) == 102
        assert len(h.get_all_targets()) == h.gdata["total_timestamps"] - h._lags
        assert h.get_all_targets()[0].shape == (20,)
    _lag_errors(HungaryCPDataLoader)


def test_pedalme_loader_shapes_and_errors():
    from stgraph_b200.dataset import PedalMeDataLoader

    for kw in ({"verbose": True}, {"redownload": True}, {"lags": 6}, {"cutoff_time": 20}):
        p = PedalMeDataLoader(**kw)
        assert p.gdata["total_timestamps"] == (36 if not p._cutoff_time else p._cutoff_time)
        assert p.gdata["num_nodes"] == 15 and p.gdata["num_edges"] == 225
        assert len(p.get_edges()) == 225 and all(len(e) == 2 for e in p.get_edges()) and len(p.get_edge_weights()) == 225
        assert p.get_all_targets().shape == (p.gdata["total_timestamps"] - p._lags, 15)
    _lag_errors(PedalMeDataLoader)


def test_windmill_loader_sizes_and_errors():
    from stgraph_b200.dataset import WindmillOutputDataLoader

    for size, (n, e) in {"large": (319, 101761), "medium": (26, 676), "small": (11, 121)}.items():
        for kw in ({"verbose": True}, {"lags": 4}, {"cutoff_time": 100}):
            w = WindmillOutputDataLoader(size=size, **kw)
            assert w.gdata["total_timestamps"] == (17472 if not w._cutoff_time else w._cutoff_time)
            assert w.gdata["num_nodes"] == n and w.gdata["num_edges"] == e
            assert len(w.get_edges()) == e and len(w.get_edge_weights()) == e
            assert len(w.get_all_targets()) == w.gdata["total_timestamps"] and w.get_all_targets()[0].shape == (n,)
    with pytest.raises(TypeError) as ex:
        WindmillOutputDataLoader(size=1)
    assert str(ex.value) == "size must be of type string"
    with pytest.raises(ValueError):
        WindmillOutputDataLoader(size="tiny")
    _lag_errors(WindmillOutputDataLoader)


def test_montevideobus_loader_shapes_and_errors():
    from stgraph_b200.dataset import MontevideoBusDataLoader

    for kw in ({"verbose": True}, {"redownload": True}, {"lags": 6}, {"cutoff_time": 50}):
        m = MontevideoBusDataLoader(**kw)
        t = m.gdata["total_timestamps"]
        assert t == (744 if not m._cutoff_time else m._cutoff_time)
        assert m.gdata["num_nodes"] == 675 and m.gdata["num_edges"] == 690
        assert len(m.get_edges()) == 690 and len(m.get_edge_weights()) == 690
        assert m.get_all_features().shape == (t - m._lags, 675, m._lags)
        assert m.get_all_targets().shape == (t - m._lags, 675)
        # the feature window of sample i ends right before its target
        assert np.allclose(m.get_all_features()[1][:, -1], m.get_all_targets()[0])
    _lag_errors(MontevideoBusDataLoader)


def test_metrla_loader_shapes_and_errors():
    from stgraph_b200.dataset import METRLADataLoader

    for kw in ({"verbose": True}, {"redownload": True}, {"num_timesteps_in": 8, "num_timesteps_out": 8}, {"cutoff_time": 50}):
        r = METRLADataLoader(**kw)
        t = r.gdata["total_timestamps"]
        assert t == (100 if not r._cutoff_time else r._cutoff_time)
        assert r.gdata["num_nodes"] == 207 and r.gdata["num_edges"] == 1722
        assert len(r.get_edges()) == 1722 and len(r.get_edges()[0]) == 2 and len(r.get_edge_weights()) == 1722
        k = t - (r._num_timesteps_in + r._num_timesteps_out) + 1
        assert r.get_all_features().shape == (k, 207, 2, r._num_timesteps_in)
        assert r.get_all_targets().shape == (k, 207, r._num_timesteps_out)
    for name in ("num_timesteps_in", "num_timesteps_out"):
        with pytest.raises(TypeError) as ex:
            METRLADataLoader(**{name: name})
        assert str(ex.value) == f"{name} must be of type int"
        with pytest.raises(ValueError) as ex:
            METRLADataLoader(**{name: -1})
        assert str(ex.value) == f"{name} must be a positive integer"
