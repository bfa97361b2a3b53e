"""Dataset loaders with the reference's surface on synthetic data (see ``stgraph_dataset.py``): every loader the
reference ships (``stgraph/dataset/__init__.py``), sized like the datasets its own tests pin."""
from .dynamic.england_covid_dataloader import EnglandCovidDataLoader
from .dynamic.stgraph_dynamic_dataset import STGraphDynamicDataset
from .static.cora_dataloader import CoraDataLoader
from .static.stgraph_static_dataset import STGraphStaticDataset
from .stgraph_dataset import STGraphDataset
from .temporal.hungarycp_dataloader import HungaryCPDataLoader
from .temporal.metrla_dataloader import METRLADataLoader
from .temporal.montevideobus_dataloader import MontevideoBusDataLoader
from .temporal.pedalme_dataloader import PedalMeDataLoader
from .temporal.stgraph_temporal_dataset import STGraphTemporalDataset
from .temporal.wikimath_dataloader import WikiMathDataLoader
from .temporal.windmilloutput_dataloader import WindmillOutputDataLoader

__all__ = ["CoraDataLoader", "EnglandCovidDataLoader", "HungaryCPDataLoader", "METRLADataLoader", "MontevideoBusDataLoader",
           "PedalMeDataLoader", "STGraphDataset", "STGraphDynamicDataset", "STGraphStaticDataset", "STGraphTemporalDataset",
           "WikiMathDataLoader", "WindmillOutputDataLoader"]
