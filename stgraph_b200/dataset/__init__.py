"""Dataset loaders with the reference's surface on synthetic data (see ``stgraph_dataset.py``).

The temporal loaders the reference ships besides WikiMath (HungaryCP, METR-LA, MontevideoBus, PedalMe, WindmillOutput)
are not mirrored: no BASELINE.json config uses them."""
from .dynamic.england_covid_dataloader import EnglandCovidDataLoader
from .dynamic.stgraph_dynamic_dataset import STGraphDynamicDataset
from .static.cora_dataloader import CoraDataLoader
from .static.stgraph_static_dataset import STGraphStaticDataset
from .stgraph_dataset import STGraphDataset
from .temporal.stgraph_temporal_dataset import STGraphTemporalDataset
from .temporal.wikimath_dataloader import WikiMathDataLoader

__all__ = ["CoraDataLoader", "EnglandCovidDataLoader", "STGraphDataset", "STGraphDynamicDataset", "STGraphStaticDataset",
           "STGraphTemporalDataset", "WikiMathDataLoader"]
